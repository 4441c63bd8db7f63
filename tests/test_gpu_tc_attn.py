"""GPU: tcgen05 flash attention against a float64 PyTorch attention and the fp32 CUDA-core kernel."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(D, kenc):
    from image_matching_b200 import Matching, synth
    cfg = golden_cfg(D=D, kenc=kenc, max_kp=128)
    sp, sg = synth.superpoint_weights(3, D), synth.superglue_weights(3, D, kenc)
    m = Matching({"superpoint": dict(cfg["superpoint"], weights=None),
                  "superglue": dict(cfg["superglue"], weights="")}).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    return m.to(DEV)


def _ref(qkv, B, Np, D, n0, n1, cross):
    heads, d = 4, D // 4
    x = qkv.double().view(2, B, Np, 3, heads, d)
    out = torch.zeros(2, B, Np, heads, d, dtype=torch.double, device=qkv.device)
    for side in range(2):
        src = 1 - side if cross else side
        nk = n0 if src == 0 else n1
        q = x[side, :, :, 0]                        # (B, Np, h, d)
        k = x[src, :, :nk, 1]
        v = x[src, :, :nk, 2]
        s = torch.einsum("bnhd,bmhd->bhnm", q, k) / d ** 0.5
        p = torch.softmax(s, -1)
        out[side] = torch.einsum("bhnm,bmhd->bnhd", p, v)
    return out.view(2 * B * Np, D)


@pytest.mark.parametrize("D,kenc", [(128, (32, 64, 128)), (256, (32, 64, 128, 256)), (64, (32, 64))])
@pytest.mark.parametrize("case", ["ones_v", "zero_k", "general", "ragged_cross", "ascending"])
def test_tc_attention(D, kenc, case):
    from image_matching_b200 import stages
    m = _model(D, kenc)
    B, Np = 2, 192
    n0, n1, cross = Np, Np, False
    g = torch.Generator(device=DEV).manual_seed(7)
    qkv = torch.randn((2 * B * Np, 3 * D), device=DEV, generator=g)
    qkv[:, :2 * D] *= 1.5
    if case == "ones_v":
        qkv[:, 2 * D:] = 1.0
    elif case == "zero_k":
        qkv[:, D:2 * D] = 0.0
    elif case == "ragged_cross":
        n0, n1, cross = 150, 77, True
    elif case == "ascending":
        # scores that keep growing along the keys: the running maximum runs away from the reference maximum several
        # times per row, which exercises the in-TMEM rescale of the lazily-rescaled accumulator
        ramp = torch.linspace(0.2, 6.0, Np, device=DEV).repeat(2 * B)[:, None]
        qkv[:, D:2 * D] = qkv[:, D:2 * D].abs() * ramp
        qkv[:, :D] = qkv[:, :D].abs()
    ref = _ref(qkv, B, Np, D, n0, n1, cross)
    simt = stages.debug_attention(m, qkv, B, Np, n0, n1, cross, False)
    tc = stages.debug_attention(m, qkv, B, Np, n0, n1, cross, True)
    e_s = float((simt.double() - ref).abs().max())
    e_t = float((tc.double() - ref).abs().max())
    print(f"D={D} {case}: max abs err simt {e_s:.2e} tc {e_t:.2e}")
    if e_t > 1e-4:
        d = (tc.double() - ref).abs().view(2, B, Np, 4, D // 4)
        print("  err by side", d.amax((1, 2, 3, 4)).tolist(), "by head", d.amax((0, 1, 2, 4)).tolist())
        print("  err by dim ", [round(v, 3) for v in d.amax((0, 1, 2, 3)).tolist()])
        print("  err by row%16", [round(v, 3) for v in d.view(2, B, Np // 16, 16, 4, D // 4).amax((0, 1, 2, 4, 5)).tolist()])
        print("  tc row0 head0", tc.view(2, B, Np, 4, D // 4)[0, 0, 0, 0, :8].tolist())
        print("  ref row0 head0", ref.view(2, B, Np, 4, D // 4)[0, 0, 0, 0, :8].tolist())
    tol = 1e-4 if case == "ascending" else 2e-5     # |scores| ~ 50 there: fp32 rounding of the logits itself is ~1e-5
    assert e_s < tol
    assert e_t < tol
