"""GPU: the small-batch path of Matching.forward.  The reference's caller runs ONE pair per call
(superpoint_glue_test.py:65-78); for such calls the library runs the two images' SuperPoint passes side by side on a
forked stream, launches every kernel with programmatic dependent launch, and replays a captured CUDA graph from the third
identical call on.  None of this may change a single bit of the results: scheduling is the only thing that differs."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_cfg, real_superpoint_weights

pytestmark = pytest.mark.gpu

KEYS_T = ("matches0", "matches1", "matching_scores0", "matching_scores1", "counts")
KEYS_K = ("keypoints0", "scores0", "descriptors0", "keypoints1", "scores1", "descriptors1")


def _model(env):
    """A Matching whose library handle is created under `env` (the switches are read when the handle is created)."""
    from image_matching_b200 import Matching, synth
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        cfg = golden_cfg(max_kp=256)
        m = Matching({"superpoint": dict(cfg["superpoint"], weights=None),
                      "superglue": dict(cfg["superglue"], weights="")}).eval()
        m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in real_superpoint_weights().items()})
        m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v))
                                     for k, v in synth.superglue_weights(0, 128).items()})
        m = m.to("cuda:0")
        # the handle is created lazily: force it now, while the environment is set
        a, b = synth.make_pair_batch([7], 120, 160)
        m.forward_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        torch.cuda.synchronize()
        return m
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _snapshot(out):
    return {k: out[k].clone() for k in KEYS_T + KEYS_K}


@pytest.mark.parametrize("pairs", [1, 3])
def test_side_by_side_and_pdl_do_not_change_results(pairs):
    from image_matching_b200 import synth
    a, b = synth.make_pair_batch(list(range(1, pairs + 1)), 240, 320)
    d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    # (both switches are per handle; B200M_PDL / B200M_PDL_MASK are read once per process and stay at their defaults)
    plain = _model({"B200M_SP_DUAL_MAX": "0", "B200M_PDL_MAX_PAIRS": "0"})
    fast = _model({"B200M_SP_DUAL_MAX": "16", "B200M_PDL_MAX_PAIRS": "8"})
    ref = _snapshot(plain.forward_device(d0, d1))
    assert int((ref["matches0"] > -1).sum()) > 20 * pairs
    out = None
    replays0 = fast._engine.graph_replays()
    for call in range(4):                 # eager, eager (key seen) / captured, replayed, replayed
        out = fast.forward_device(d0, d1, out=out)
        got = _snapshot(out)
        for k in KEYS_T + KEYS_K:
            assert torch.equal(ref[k], got[k]), (k, call)
    assert fast._engine.graph_replays() > replays0          # the later calls really were graph replays


def test_single_side_workspace_is_accepted():
    """b200m_matching_forward with the single-side workspace size (what a caller sized before the side-by-side path
    existed) runs the two SuperPoint passes one after the other and returns the same results."""
    import ctypes as C
    from image_matching_b200 import lib, synth
    m = _model({})
    L = lib.load()
    h = m._engine.handle
    B, H, W = 2, 120, 160
    a, b = synth.make_pair_batch([3, 4], H, W)
    d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ref = _snapshot(m.forward_device(d0, d1))
    cap = L.b200m_keypoint_capacity(h, H, W)
    full = L.b200m_matching_workspace_bytes(h, B, H, W)
    small = full - (full - L.b200m_superglue_workspace_bytes(h, B, cap, cap)) // 2 + 1024   # ~ one SuperPoint side less
    assert small < full
    D = 128
    dev = d0.device
    bufs = {"kp": [torch.zeros(B, cap, 2, device=dev) for _ in range(2)],
            "sc": [torch.zeros(B, cap, device=dev) for _ in range(2)],
            "de": [torch.zeros(B, D, cap, device=dev) for _ in range(2)],
            "cn": [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(2)],
            "m": [torch.zeros(B, cap, dtype=torch.int64, device=dev) for _ in range(2)],
            "ms": [torch.zeros(B, cap, device=dev) for _ in range(2)]}
    ws = torch.empty(small, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    for _ in range(3):
        rc = L.b200m_matching_forward(h, p(d0), p(d1), B, H, W, p(bufs["kp"][0]), p(bufs["sc"][0]), p(bufs["de"][0]),
                                      p(bufs["cn"][0]), p(bufs["kp"][1]), p(bufs["sc"][1]), p(bufs["de"][1]),
                                      p(bufs["cn"][1]), cap, p(bufs["m"][0]), p(bufs["m"][1]), p(bufs["ms"][0]),
                                      p(bufs["ms"][1]), p(ws), C.c_size_t(small),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.last_error()
    torch.cuda.synchronize()
    n = int(ref["counts"][0].max())
    assert torch.equal(bufs["m"][0][:, :n], ref["matches0"][:, :n])
    assert torch.equal(bufs["kp"][0][:, :n], ref["keypoints0"][:, :n])
