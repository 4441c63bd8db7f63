"""GPU: the fused GNN layer kernel (tc_gnn.cu: merge -> mlp -> residual -> next q|k|v in one launch) against the
unfused four-GEMM path of the same library and a float64 PyTorch restatement of the layer
(superglue/models/superglue_test.py:92-138)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D = 128


def _model(seed=3):
    from image_matching_b200 import Matching, synth
    cfg = golden_cfg(D=D, kenc=(32, 64, 128), max_kp=128)
    sp, sg = synth.superpoint_weights(seed, D), synth.superglue_weights(seed, D, (32, 64, 128))
    m = Matching({"superpoint": dict(cfg["superpoint"], weights=None),
                  "superglue": dict(cfg["superglue"], weights="")}).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    return m.to(DEV), sg


def _gnn(m, d0, d1, lb, le, c0=None, c1=None, unfused=False):
    """stages.gnn with the engine created under B200M_GNN_IMPL (read by b200m_create)."""
    from image_matching_b200 import stages
    old = os.environ.get("B200M_GNN_IMPL")
    if unfused:
        os.environ["B200M_GNN_IMPL"] = "unfused"
    try:
        return stages.gnn(m, d0, d1, lb, le, c0, c1)
    finally:
        if unfused:
            if old is None:
                os.environ.pop("B200M_GNN_IMPL", None)
            else:
                os.environ["B200M_GNN_IMPL"] = old


def _ref_layers(sg, d0, d1, lb, le, names):
    """float64 restatement: desc (B,D,N) each."""
    W = {k: torch.from_numpy(np.asarray(v)).double().to(d0.device) for k, v in sg.items()}
    x0, x1 = d0.double(), d1.double()

    def conv(x, name):
        w = W[name + ".weight"]
        return torch.einsum("oi,bin->bon", w.view(w.shape[0], -1), x) + W[name + ".bias"][None, :, None]

    def attn_prop(l, x, src):
        p = f"gnn.layers.{l}"
        B = x.shape[0]
        q, k, v = conv(x, p + ".attn.proj.0"), conv(src, p + ".attn.proj.1"), conv(src, p + ".attn.proj.2")
        q, k, v = (t.view(B, D // 4, 4, -1) for t in (q, k, v))
        s = torch.einsum("bdhn,bdhm->bhnm", q, k) / (D // 4) ** 0.5
        msg = torch.einsum("bhnm,bdhm->bdhn", torch.softmax(s, -1), v).reshape(B, D, -1)
        msg = conv(msg, p + ".attn.merge")
        h = conv(torch.cat([x, msg], 1), p + ".mlp.0")
        g, b = W[p + ".mlp.1.weight"], W[p + ".mlp.1.bias"]
        mu, var = W[p + ".mlp.1.running_mean"], W[p + ".mlp.1.running_var"]
        h = (h - mu[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5) * g[None, :, None] + b[None, :, None]
        return conv(torch.relu(h), p + ".mlp.3")

    for l in range(lb, le):
        if names[l] == "cross":
            dl0, dl1 = attn_prop(l, x0, x1), attn_prop(l, x1, x0)
        else:
            dl0, dl1 = attn_prop(l, x0, x0), attn_prop(l, x1, x1)
        x0, x1 = x0 + dl0, x1 + dl1
    return x0, x1


@pytest.mark.parametrize("lb,le", [(0, 1), (0, 2), (3, 6), (16, 18), (0, 18)])
def test_fused_layers_vs_float64(lb, le):
    m, sg = _model()
    names = m.superglue.config["GNN_layers"]
    g = torch.Generator(device=DEV).manual_seed(11)
    B, N, M = 2, 320, 256
    d0 = torch.nn.functional.normalize(torch.randn((B, D, N), device=DEV, generator=g), dim=1)
    d1 = torch.nn.functional.normalize(torch.randn((B, D, M), device=DEV, generator=g), dim=1)
    r0, r1 = _ref_layers(sg, d0, d1, lb, le, names)
    f0, f1 = _gnn(m, d0, d1, lb, le)
    scale = float(max(r0.abs().max(), r1.abs().max()))
    e = max(float((f0.double() - r0).abs().max()), float((f1.double() - r1).abs().max())) / scale
    print(f"layers [{lb},{le}): fused vs float64 max rel err {e:.2e} (scale {scale:.2f})")
    assert e < 2e-5


def test_fused_equals_unfused_ragged_many_tiles():
    """More 128-token tiles than SMs (persistent loop takes a second tile), ragged per-pair counts."""
    m, _ = _model(5)
    m2, _ = _model(5)
    g = torch.Generator(device=DEV).manual_seed(12)
    B, N = 10, 1024                                   # 2 * 10 * 1024 rows = 160 tiles > 148 SMs
    d0 = torch.nn.functional.normalize(torch.randn((B, D, N), device=DEV, generator=g), dim=1)
    d1 = torch.nn.functional.normalize(torch.randn((B, D, N), device=DEV, generator=g), dim=1)
    c0 = torch.tensor([1024, 1000, 513, 512, 511, 129, 128, 127, 1, 777], dtype=torch.int32, device=DEV)
    c1 = torch.tensor([1024, 31, 1024, 640, 64, 65, 1023, 300, 1024, 2], dtype=torch.int32, device=DEV)
    f0, f1 = _gnn(m, d0, d1, 0, 4, c0, c1)
    u0, u1 = _gnn(m2, d0, d1, 0, 4, c0, c1, unfused=True)
    for b in range(B):
        n, mm = int(c0[b]), int(c1[b])
        e0 = float((f0[b, :, :n] - u0[b, :, :n]).abs().max()) / float(u0[b, :, :n].abs().max())
        e1 = float((f1[b, :, :mm] - u1[b, :, :mm]).abs().max()) / float(u1[b, :, :mm].abs().max())
        assert e0 < 2e-5 and e1 < 2e-5, (b, e0, e1)
