"""Stage-level entry points of the C ABI as torch-tensor functions (used by the parity tests:
one per reference function, so a mismatch is localised to a kernel).  All tensors are CUDA."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib
from .matching import Matching, _ptr, _stream


def _prep(m: Matching, device):
    L = m._ensure(device)
    return L, m._engine


def superpoint_dense(m: Matching, images: torch.Tensor):
    """images (n,1,H,W) -> semi (n,65,h,w), desc (n,D,h,w)   [superpoint_test.py:113-126]"""
    images = images.contiguous().float()
    L, e = _prep(m, images.device)
    n, _, H, W = images.shape
    hc, wc = H // 2 // 2 // 2, W // 2 // 2 // 2
    D = m.superpoint.config["descriptor_dim"]
    semi = torch.empty((n, 65, hc, wc), device=images.device)
    desc = torch.empty((n, D, hc, wc), device=images.device)
    ws = e.workspace(L.b200m_superpoint_workspace_bytes(e.handle, n, H, W), images.device)
    _lib.check(L.b200m_superpoint_dense(e.handle, _ptr(images), n, H, W, _ptr(semi), _ptr(desc), _ptr(ws),
                                        ws.numel(), _stream(images.device)), "b200m_superpoint_dense")
    return semi, desc


def detector_post(m: Matching, semi: torch.Tensor):
    """semi (n,65,h,w) -> heat, nms (n,8h,8w), keypoints (n,cap,2), scores (n,cap), counts (n)
    [superpoint_test.py:128-151]"""
    semi = semi.contiguous().float()
    L, e = _prep(m, semi.device)
    n, _, hc, wc = semi.shape
    dev = semi.device
    cap = int(L.b200m_keypoint_capacity(e.handle, hc * 8, wc * 8))
    heat = torch.empty((n, hc * 8, wc * 8), device=dev)
    nms = torch.empty((n, hc * 8, wc * 8), device=dev)
    kp = torch.empty((n, cap, 2), device=dev)
    sc = torch.empty((n, cap), device=dev)
    cnt = torch.empty((n,), dtype=torch.int32, device=dev)
    ws = e.workspace(L.b200m_superpoint_workspace_bytes(e.handle, n, hc * 8, wc * 8), dev)
    _lib.check(L.b200m_detector_post(e.handle, _ptr(semi), n, hc, wc, _ptr(heat), _ptr(nms), _ptr(kp), _ptr(sc),
                                     _ptr(cnt), cap, _ptr(ws), ws.numel(), _stream(semi.device)), "b200m_detector_post")
    return heat, nms, kp, sc, cnt


def sample_descriptors(m: Matching, keypoints: torch.Tensor, counts, desc: torch.Tensor):
    """keypoints (n,cap,2), counts (n) or None, desc (n,D,h,w) -> (n,D,cap)   [superpoint_test.py:40-52]"""
    keypoints = keypoints.contiguous().float()
    desc = desc.contiguous().float()
    L, e = _prep(m, desc.device)
    n, D, hc, wc = desc.shape
    cap = keypoints.shape[1]
    out = torch.empty((n, D, cap), device=desc.device)
    _lib.check(L.b200m_sample_descriptors(e.handle, _ptr(keypoints), _ptr(counts), _ptr(desc), n, hc, wc, cap,
                                          _ptr(out), _stream(desc.device)), "b200m_sample_descriptors")
    return out


def _sg_ws(L, e, B, N, M, dev):
    return e.workspace(L.b200m_superglue_workspace_bytes(e.handle, B, N, M), dev)


def keypoint_encode(m: Matching, kpts, scores, desc, H, W):
    """desc + kenc(normalize_keypoints(kpts), scores)   [superglue_test.py:63-82, 249-250]"""
    kpts, scores, desc = kpts.contiguous().float(), scores.contiguous().float(), desc.contiguous().float()
    L, e = _prep(m, desc.device)
    B, D, N = desc.shape
    out = torch.empty_like(desc)
    ws = _sg_ws(L, e, B, N, N, desc.device)
    _lib.check(L.b200m_keypoint_encode(e.handle, _ptr(kpts), _ptr(scores), _ptr(desc), B, N, H, W, _ptr(out),
                                       _ptr(ws), ws.numel(), _stream(desc.device)), "b200m_keypoint_encode")
    return out


def gnn(m: Matching, desc0, desc1, layer_begin=0, layer_end=None, counts0=None, counts1=None):
    """AttentionalGNN layers [layer_begin, layer_end)   [superglue_test.py:85-138]"""
    desc0, desc1 = desc0.contiguous().float(), desc1.contiguous().float()
    L, e = _prep(m, desc0.device)
    B, D, N = desc0.shape
    M = desc1.shape[2]
    if layer_end is None:
        layer_end = len(m.superglue.config["GNN_layers"])
    o0, o1 = torch.empty_like(desc0), torch.empty_like(desc1)
    ws = _sg_ws(L, e, B, N, M, desc0.device)
    _lib.check(L.b200m_gnn(e.handle, _ptr(desc0), _ptr(desc1), _ptr(counts0), _ptr(counts1), B, N, M,
                           layer_begin, layer_end, _ptr(o0), _ptr(o1), _ptr(ws), ws.numel(), _stream(desc0.device)),
               "b200m_gnn")
    return o0, o1


def score_matrix(m: Matching, desc0, desc1):
    """final_proj + einsum / sqrt(D)   [superglue_test.py:256-260]"""
    desc0, desc1 = desc0.contiguous().float(), desc1.contiguous().float()
    L, e = _prep(m, desc0.device)
    B, D, N = desc0.shape
    M = desc1.shape[2]
    S = torch.empty((B, N, M), device=desc0.device)
    ws = _sg_ws(L, e, B, N, M, desc0.device)
    _lib.check(L.b200m_score_matrix(e.handle, _ptr(desc0), _ptr(desc1), B, N, M, _ptr(S), _ptr(ws), ws.numel(),
                                    _stream(desc0.device)), "b200m_score_matrix")
    return S


def sinkhorn(m: Matching, S, iters=None):
    """log_optimal_transport   [superglue_test.py:141-170]"""
    S = S.contiguous().float()
    L, e = _prep(m, S.device)
    B, N, M = S.shape
    if iters is None:
        iters = m.superglue.config["sinkhorn_iterations"]
    Z = torch.empty((B, N + 1, M + 1), device=S.device)
    ws = _sg_ws(L, e, B, N, M, S.device)
    _lib.check(L.b200m_sinkhorn(e.handle, _ptr(S), B, N, M, iters, _ptr(Z), _ptr(ws), ws.numel(), _stream(S.device)),
               "b200m_sinkhorn")
    return Z


def match_select(m: Matching, Z):
    """mutual check + threshold   [superglue_test.py:268-285]"""
    Z = Z.contiguous().float()
    L, e = _prep(m, Z.device)
    B, N1, M1 = Z.shape
    N, M = N1 - 1, M1 - 1
    dev = Z.device
    m0 = torch.empty((B, N), dtype=torch.int64, device=dev)
    m1 = torch.empty((B, M), dtype=torch.int64, device=dev)
    s0 = torch.empty((B, N), device=dev)
    s1 = torch.empty((B, M), device=dev)
    ws = _sg_ws(L, e, B, N, M, dev)
    _lib.check(L.b200m_match_select(e.handle, _ptr(Z), B, N, M, _ptr(m0), _ptr(m1), _ptr(s0), _ptr(s1), _ptr(ws),
                                    ws.numel(), _stream(Z.device)), "b200m_match_select")
    return m0, m1, s0, s1


def debug_conv_layer(m: Matching, layer: int, use_tc: bool, x: torch.Tensor):
    """One packed 3x3 SuperPoint layer (0..7) with the tcgen05 (use_tc) or the fp32 CUDA-core kernel."""
    x = x.contiguous().float()
    L, e = _prep(m, x.device)
    n, cin, H, W = x.shape
    cout = [64, 64, 64, 128, 128, 128, 128, 512][layer]
    pool = layer in (0, 2, 4)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.empty((n, cout, Ho, Wo), device=x.device)
    _lib.check(L.b200m_debug_conv_layer(e.handle, layer, int(use_tc), _ptr(x), _ptr(out), n, H, W, _stream(x.device)),
               "b200m_debug_conv_layer")
    return out


def debug_attention(m: Matching, qkv: torch.Tensor, B: int, Np: int, n0: int, n1: int, cross: bool, use_tc: bool):
    """Attention on a fused q|k|v buffer (2*B*Np, 3D) with head-major columns -> (2*B*Np, D)."""
    qkv = qkv.contiguous().float()
    L, e = _prep(m, qkv.device)
    D = qkv.shape[1] // 3
    out = torch.zeros((qkv.shape[0], D), device=qkv.device)
    _lib.check(L.b200m_debug_attention(e.handle, _ptr(qkv), _ptr(out), B, Np, n0, n1, int(cross), int(use_tc),
                                       _stream(qkv.device)), "b200m_debug_attention")
    return out
