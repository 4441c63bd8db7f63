"""Recipe for ``oracle/_ref``: the reference's OWN hot-path modules, unmodified, as the CPU baseline.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product path).

The reference (PH8411/image-matching) is a pure-Python script collection: the whole ``Matching.forward`` path lives in
four files.  This recipe copies exactly those files from the reference checkout (``/root/reference`` in the build
container, or ``$B200M_REF``) into ``oracle/_ref/`` with their package layout, byte for byte:

    superglue/models/matching_test.py     Matching          (matching_test.py:47-82)
    superglue/models/superglue_test.py    SuperGlue         (superglue_test.py:177-285)
    superpoint/models/superpoint_test.py  SuperPoint        (superpoint_test.py:55-161)
    superpoint/models/unet_parts.py       double_conv/down  (unet_parts.py:10-48)

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box like the built ``.so``; ``bench.py --impl reference`` and the ``cpu_baseline`` leg import the
modules from there (``cpu_baseline.kind = "reference"``) and fall back to the torch-CPU oracle port
(``oracle/matching_oracle_torch.py``, ``kind = "port"``) only when the directory is absent.
``__graft_entry__.build()`` runs this whenever the reference checkout is present.

    python oracle/make_ref.py            # (re)create oracle/_ref
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["superglue/models/matching_test.py", "superglue/models/superglue_test.py",
         "superpoint/models/superpoint_test.py", "superpoint/models/unet_parts.py"]
# the reference's regular packages (superglue/ and superpoint/ themselves are namespace packages there)
INITS = ["superglue/models/__init__.py", "superpoint/models/__init__.py"]


def reference_root():
    for cand in (os.environ.get("B200M_REF"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, FILES[0])):
            return cand
    return None


def make(verbose: bool = True) -> str | None:
    """Copy the four files (+ the two package markers) into oracle/_ref; returns the path, or None when no reference
    checkout is available (the existing copy, if any, is left alone)."""
    ref = reference_root()
    if ref is None:
        return DST if available() else None
    manifest = {}
    for rel in FILES + INITS:
        src, dst = os.path.join(ref, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        else:                     # package marker absent in the reference: an empty one keeps the import path identical
            open(dst, "w").close()
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": ref, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("oracle/_ref: copied", len(FILES), "reference modules from", ref)
    return DST


def available() -> bool:
    return all(os.path.isfile(os.path.join(DST, rel)) for rel in FILES)


def load_matching(cfg, sp_sd, sg_sd):
    """Build the reference's ``Matching`` from oracle/_ref with the given state dicts (numpy), following SURVEY.md
    Appendix B: weights=None / '' at construction, then ``load_state_dict`` (the checkpoints' own ``torch.load`` path
    needs CUDA-saved pickles).  Returns the eval-mode module (CPU)."""
    import numpy as np
    import torch
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where the reference checkout exists")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from superglue.models.matching_test import Matching      # the reference's module, unmodified
    torch.set_grad_enabled(False)
    c = {"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}
    m = Matching(c).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp_sd.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg_sd.items()})
    return m


if __name__ == "__main__":
    p = make()
    print(p if p else "no reference checkout found")
