import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_cfg(D=128, kenc=(32, 64, 128), max_kp=1024, iters=30, mthr=0.2, kthr=0.005):
    """Same configuration dict tests/golden/make_golden.py used."""
    return {"superpoint": {"descriptor_dim": D, "nms_radius": 4, "keypoint_threshold": kthr,
                           "max_keypoints": max_kp, "remove_borders": 4},
            "superglue": {"descriptor_dim": D, "keypoint_encoder": list(kenc),
                          "GNN_layers": ["self", "cross"] * 9,
                          "sinkhorn_iterations": iters, "match_threshold": mthr}}


def real_superpoint_weights():
    return dict(np.load(os.path.join(GOLDEN, "superpoint_allss128_weights.npz")))


def kp_set(kp):
    return set(map(tuple, np.asarray(kp).astype(np.int64).tolist()))


def match_pairs(kp0, kp1, m0):
    """matches as a set of coordinate pairs (order-canonical; SURVEY.md 8d correctness gates)."""
    kp0 = np.asarray(kp0).astype(np.int64)
    kp1 = np.asarray(kp1).astype(np.int64)
    m0 = np.asarray(m0)
    return {(tuple(kp0[i]), tuple(kp1[j])) for i, j in enumerate(m0.tolist()) if j >= 0}
