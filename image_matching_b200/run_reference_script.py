"""Run one of the reference's own entry scripts (e.g. superpoint_glue_test.py) UNCHANGED on top of the
B200 path:

    python -m image_matching_b200.run_reference_script /path/to/image-matching superpoint_glue_test.py [script args]

`superglue.models.matching_test` (and the two model modules it imports) are pre-seeded in sys.modules
with this package's drop-in classes, so the script's `from superglue.models.matching_test import Matching`
(superpoint_glue_test.py:10) resolves to the CUDA implementation.  `matplotlib.cm` is stubbed only if
matplotlib is not installed (the script uses cm.jet for overlay colours, :115).
"""
from __future__ import annotations

import os
import runpy
import sys
import types


def install_shims():
    from . import matching as _m
    for name, attrs in (("superglue.models.matching_test", {"Matching": _m.Matching}),
                        ("superpoint.models.superpoint_test", {"SuperPoint": _m.SuperPoint}),
                        ("superglue.models.superglue_test", {"SuperGlue": _m.SuperGlue}),
                        # the "official" variant used by superpoint_glue_official_test.py:10
                        ("superglue.models.matching", {"Matching": _m.MatchingOfficial}),
                        ("superglue.models.superpoint", {"SuperPoint": _m.SuperPointOfficial})):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        sys.modules[name] = mod
    try:
        import matplotlib.cm  # noqa: F401
    except Exception:
        import numpy as np
        mpl = types.ModuleType("matplotlib")
        cm = types.ModuleType("matplotlib.cm")
        cm.jet = lambda x: np.stack([np.asarray(x)] * 3 + [np.ones_like(np.asarray(x))], -1)
        pyplot = types.ModuleType("matplotlib.pyplot")
        mpl.cm, mpl.pyplot = cm, pyplot
        mpl.use = lambda *a, **k: None
        sys.modules.update({"matplotlib": mpl, "matplotlib.cm": cm, "matplotlib.pyplot": pyplot})


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit(__doc__)
    ref_root, script = os.path.abspath(argv[0]), argv[1]
    install_shims()
    sys.path.insert(0, ref_root)
    os.chdir(ref_root)
    sys.argv = [script] + argv[2:]
    runpy.run_path(os.path.join(ref_root, script), run_name="__main__")


if __name__ == "__main__":
    main()
