"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A numpy/fp32 restatement of the reference's SuperPoint+SuperGlue ``Matching.forward``
hot path (PH8411/image-matching @ 26bdfe05).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / CPU baseline.

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 4, 8c).  The oracle is instead pinned against outputs of the reference
modules themselves, imported from /root/reference in the build container by
``tests/golden/make_golden.py`` (committed, with the fixtures it wrote under
``tests/golden/``); ``tests/test_oracle_golden.py`` re-checks the oracle against
those fixtures on every run.

All arithmetic of the reference is torch (un-vendored, no version pin in the
reference; torch 2.11.0 in this image).  The torch ops on the path are restated
from their published semantics: conv2d/conv1d (cross-correlation, zero padding),
BatchNorm eval, max_pool2d (-inf padding), softmax, topk (descending), grid_sample
(bilinear, zeros padding, align_corners as selected by the reference's version
test), normalize (eps 1e-12), logsumexp (max-shifted).

Each function cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- helpers
def fold_bn(w, b, sd, bn, eps=1e-5):
    """conv+BatchNorm(eval) -> conv.  torch.nn.BatchNorm{1,2}d eval semantics
    (superpoint/models/unet_parts.py:15-20; superglue/models/superglue_test.py:56-58).
    A state_dict without the BatchNorm tensors (the "official" SuperPoint, superglue/models/superpoint.py) is used as is."""
    if bn + ".weight" not in sd:
        return w.astype(F32), b.astype(F32)
    g = sd[bn + ".weight"].astype(np.float64)
    beta = sd[bn + ".bias"].astype(np.float64)
    mu = sd[bn + ".running_mean"].astype(np.float64)
    var = sd[bn + ".running_var"].astype(np.float64)
    s = g / np.sqrt(var + eps)
    wf = w.astype(np.float64) * s.reshape((-1,) + (1,) * (w.ndim - 1))
    bf = (b.astype(np.float64) - mu) * s + beta
    return wf.astype(F32), bf.astype(F32)


def conv3x3_hwc(x, w, b, relu):
    """3x3 stride-1 zero-pad-1 cross-correlation; x (H,W,C) channel-last, w (O,C,3,3) -> (H,W,O)
    (unet_parts.py:15,18).  im2col over strips of rows + one BLAS sgemm per strip (pixels x 9C @ 9C x O;
    the pixel-major operand order is the one OpenBLAS runs at full speed)."""
    H, W, C = x.shape
    O = w.shape[0]
    xp = np.zeros((H + 2, W + 2, C), F32)
    xp[1:-1, 1:-1] = x
    w2t = np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(9 * C, O))    # (ky, kx, c) x O
    out = np.empty((H, W, O), F32)
    R = max(1, min(H, (1 << 22) // max(1, 9 * C * W)))                       # ~16 MB im2col strip
    for r0 in range(0, H, R):
        r1 = min(H, r0 + R)
        col = np.empty((r1 - r0, W, 9, C), F32)
        for ky in range(3):
            for kx in range(3):
                col[:, :, ky * 3 + kx, :] = xp[ky + r0:ky + r1, kx:kx + W, :]
        out[r0:r1] = (col.reshape(-1, 9 * C) @ w2t).reshape(r1 - r0, W, O)
    out += b
    if relu:
        np.maximum(out, 0, out=out)
    return out


def conv3x3(x, w, b, relu):
    """CHW wrapper of conv3x3_hwc (used by the tests)."""
    return np.ascontiguousarray(conv3x3_hwc(np.ascontiguousarray(x.transpose(1, 2, 0)), w, b, relu).transpose(2, 0, 1))


def conv1x1(x, w, b, relu=False):
    """1x1 conv / Conv1d(k=1); x (C, ...) ; w (O,C) or (O,C,1[,1])."""
    w2 = w.reshape(w.shape[0], w.shape[1])
    sh = x.shape[1:]
    out = w2 @ x.reshape(x.shape[0], -1) + b[:, None]
    if relu:
        np.maximum(out, 0, out=out)
    return out.reshape((w2.shape[0],) + sh).astype(F32)


def maxpool2_hwc(x):
    """nn.MaxPool2d(2) floor mode (unet_parts.py:42), channel-last."""
    H, W, C = x.shape
    h, w = H // 2, W // 2
    v = x[:2 * h, :2 * w]
    return np.maximum(np.maximum(v[0::2, 0::2], v[0::2, 1::2]), np.maximum(v[1::2, 0::2], v[1::2, 1::2]))


# ----------------------------------------------------------------------------- SuperPoint
def superpoint_dense(img, sd):
    """Encoder + heads -> (semi (65,h,w), desc (D,h,w) channel-L2-normalised).
    superpoint/models/superpoint_test.py:113-126, unet_parts.py:10-48."""
    x = img.reshape(img.shape[-2], img.shape[-1], 1).astype(F32)       # channel-last internally
    if "conv1a.weight" in sd:
        # "official" variant (superglue/models/superpoint.py:116-132, :151-162): same topology, no BatchNorm, other names
        ren = {"conv1a": "inc.conv.conv.0", "conv1b": "inc.conv.conv.3", "conv2a": "down1.mpconv.1.conv.0",
               "conv2b": "down1.mpconv.1.conv.3", "conv3a": "down2.mpconv.1.conv.0", "conv3b": "down2.mpconv.1.conv.3",
               "conv4a": "down3.mpconv.1.conv.0", "conv4b": "down3.mpconv.1.conv.3"}
        sd = {(ren.get(k.rsplit(".", 1)[0], k.rsplit(".", 1)[0]) + "." + k.rsplit(".", 1)[1]): v for k, v in sd.items()}

    def dconv(x, p):
        w, b = fold_bn(sd[p + ".0.weight"], sd[p + ".0.bias"], sd, p + ".1")
        x = conv3x3_hwc(x, w, b, True)
        w, b = fold_bn(sd[p + ".3.weight"], sd[p + ".3.bias"], sd, p + ".4")
        return conv3x3_hwc(x, w, b, True)

    def c1(x, w, b):                                                    # 1x1 conv, channel-last
        return (x.reshape(-1, x.shape[-1]) @ np.ascontiguousarray(w.reshape(w.shape[0], -1).T) + b).reshape(
            x.shape[0], x.shape[1], -1).astype(F32)

    x1 = dconv(x, "inc.conv.conv")
    x2 = dconv(maxpool2_hwc(x1), "down1.mpconv.1.conv")
    x3 = dconv(maxpool2_hwc(x2), "down2.mpconv.1.conv")
    x4 = dconv(maxpool2_hwc(x3), "down3.mpconv.1.conv")
    w, b = fold_bn(sd["convPa.weight"], sd["convPa.bias"], sd, "bnPa")
    cPa = conv3x3_hwc(x4, w, b, True)
    w, b = fold_bn(sd["convPb.weight"], sd["convPb.bias"], sd, "bnPb")
    semi = np.ascontiguousarray(c1(cPa, w, b).transpose(2, 0, 1))
    w, b = fold_bn(sd["convDa.weight"], sd["convDa.bias"], sd, "bnDa")
    cDa = conv3x3_hwc(x4, w, b, True)
    w, b = fold_bn(sd["convDb.weight"], sd["convDb.bias"], sd, "bnDb")
    desc = np.ascontiguousarray(c1(cDa, w, b).transpose(2, 0, 1))
    dn = np.sqrt((desc.astype(F32) ** 2).sum(0, dtype=F32))
    desc = desc / dn[None]                      # no eps (superpoint_test.py:125-126)
    return semi.astype(F32), desc.astype(F32)


def heatmap(semi):
    """softmax over 65 channels, drop dustbin, depth-to-space x8 (superpoint_test.py:128-131)."""
    m = semi.max(0, keepdims=True)
    e = np.exp((semi - m).astype(F32))
    p = (e / e.sum(0, keepdims=True, dtype=F32))[:-1]
    _, h, w = p.shape
    p = p.transpose(1, 2, 0).reshape(h, w, 8, 8)
    return np.ascontiguousarray(p.transpose(0, 2, 1, 3)).reshape(h * 8, w * 8).astype(F32)


def _maxpool_same(x, r):
    """max_pool2d(kernel 2r+1, stride 1, padding r) with -inf padding (superpoint_test.py:11-13)."""
    H, W = x.shape
    xp = np.full((H + 2 * r, W + 2 * r), -np.inf, F32)
    xp[r:r + H, r:r + W] = x
    t = xp[:, 0:W].copy()
    for d in range(1, 2 * r + 1):
        np.maximum(t, xp[:, d:d + W], out=t)
    o = t[0:H].copy()
    for d in range(1, 2 * r + 1):
        np.maximum(o, t[d:d + H], out=o)
    return o


def simple_nms(scores, r):
    """superpoint_test.py:7-22."""
    assert r >= 0
    zeros = np.zeros_like(scores)
    max_mask = scores == _maxpool_same(scores, r)
    for _ in range(2):
        supp_mask = _maxpool_same(max_mask.astype(F32), r) > 0
        supp_scores = np.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == _maxpool_same(supp_scores, r)
        max_mask = max_mask | (new_max_mask & (~supp_mask))
    return np.where(max_mask, scores, zeros)


def extract_keypoints(nms, thr, border, max_kp):
    """nonzero(s>thr) row-major, remove_borders, top-k descending (superpoint_test.py:25-37,135-151).
    Returns keypoints (n,2) float32 in (x,y) order and scores (n,)."""
    H, W = nms.shape
    ys, xs = np.nonzero(nms > F32(thr))
    sc = nms[ys, xs]
    keep = (ys >= border) & (ys < H - border) & (xs >= border) & (xs < W - border)
    ys, xs, sc = ys[keep], xs[keep], sc[keep]
    if max_kp >= 0 and max_kp < len(sc):
        order = np.argsort(-sc, kind="stable")[:max_kp]
        ys, xs, sc = ys[order], xs[order], sc[order]
    kp = np.stack([xs, ys], 1).astype(F32).reshape(-1, 2)
    return kp, sc.astype(F32)


def sample_descriptors(kp, desc, align_corners=False, s=8):
    """grid_sample(bilinear, zeros) + L2 normalise (superpoint_test.py:40-52).
    kp (n,2) xy; desc (D,h,w) -> (D,n)."""
    D, h, w = desc.shape
    n = kp.shape[0]
    if n == 0:
        return np.zeros((D, 0), F32)
    k = kp.astype(F32) - F32(s / 2) + F32(0.5)
    k = k / np.array([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5], F32)[None]
    g = k * F32(2) - F32(1)
    if align_corners:
        px = (g[:, 0] + 1) / 2 * (w - 1)
        py = (g[:, 1] + 1) / 2 * (h - 1)
    else:
        px = ((g[:, 0] + 1) * w - 1) / 2
        py = ((g[:, 1] + 1) * h - 1) / 2
    px = px.astype(F32)
    py = py.astype(F32)
    x0 = np.floor(px).astype(np.int64)
    y0 = np.floor(py).astype(np.int64)
    fx = (px - x0).astype(F32)
    fy = (py - y0).astype(F32)
    out = np.zeros((D, n), F32)
    for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                        (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xi, yi = x0 + dx, y0 + dy
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        v = desc[:, np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)]
        out += v * (wgt * ok).astype(F32)[None]
    nrm = np.sqrt((out * out).sum(0, dtype=F32))
    return (out / np.maximum(nrm, F32(1e-12))[None]).astype(F32)


def superpoint_forward(img, sd, cfg, align_corners=False):
    """One image (H,W) -> dict(keypoints (n,2), scores (n,), descriptors (D,n)).
    superpoint/models/superpoint_test.py:103-161."""
    semi, desc = superpoint_dense(img, sd)
    heat = heatmap(semi)
    nms = simple_nms(heat, cfg.get("nms_radius", 4))
    kp, sc = extract_keypoints(nms, cfg.get("keypoint_threshold", 0.005),
                               cfg.get("remove_borders", 4), cfg.get("max_keypoints", -1))
    return {"keypoints": kp, "scores": sc,
            "descriptors": sample_descriptors(kp, desc, align_corners),
            "semi": semi, "desc": desc, "heat": heat, "nms": nms}


# ----------------------------------------------------------------------------- SuperGlue
def normalize_keypoints(kp, H, W):
    """superglue_test.py:63-70."""
    size = np.array([W, H], F32)
    center = size / 2
    scaling = F32(size.max() * F32(0.7))
    return ((kp - center[None]) / scaling).astype(F32)


def _mlp(x, sd, prefix, n_convs):
    """Conv1d(k=1) [+BN+ReLU] chain (superglue_test.py:49-60). x (C,N)."""
    idx = 0
    for i in range(n_convs):
        w, b = sd[f"{prefix}.{idx}.weight"], sd[f"{prefix}.{idx}.bias"]
        last = i == n_convs - 1
        if not last:
            w, b = fold_bn(w, b, sd, f"{prefix}.{idx + 1}")
        x = conv1x1(x, w, b, relu=not last)
        idx += 1 if last else 3
    return x


def keypoint_encoder(kpn, scores, sd, n_convs):
    """superglue_test.py:73-82: cat(kpts^T, scores) -> MLP."""
    inp = np.concatenate([kpn.T, scores[None]], 0).astype(F32)
    return _mlp(inp, sd, "kenc.encoder", n_convs)


def attentional_propagation(x, src, sd, p, heads=4):
    """superglue_test.py:85-119.  x (D,N), src (D,M) -> delta (D,N)."""
    D = x.shape[0]
    d = D // heads
    q = conv1x1(x, sd[p + ".attn.proj.0.weight"], sd[p + ".attn.proj.0.bias"])
    k = conv1x1(src, sd[p + ".attn.proj.1.weight"], sd[p + ".attn.proj.1.bias"])
    v = conv1x1(src, sd[p + ".attn.proj.2.weight"], sd[p + ".attn.proj.2.bias"])
    q = q.reshape(d, heads, -1)
    k = k.reshape(d, heads, -1)
    v = v.reshape(d, heads, -1)
    out = np.zeros((d, heads, x.shape[1]), F32)
    for h in range(heads):
        s = (q[:, h].T @ k[:, h]) / F32(d ** 0.5)
        s = s - s.max(1, keepdims=True)
        e = np.exp(s, dtype=F32)
        prob = e / e.sum(1, keepdims=True, dtype=F32)
        out[:, h] = (prob @ v[:, h].T).T
    msg = conv1x1(out.reshape(D, -1), sd[p + ".attn.merge.weight"], sd[p + ".attn.merge.bias"])
    y = np.concatenate([x, msg], 0)
    w, b = fold_bn(sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"], sd, p + ".mlp.1")
    y = conv1x1(y, w, b, relu=True)
    return conv1x1(y, sd[p + ".mlp.3.weight"], sd[p + ".mlp.3.bias"])


def log_optimal_transport(S, alpha, iters):
    """superglue_test.py:141-170 for one pair; S (N,M) -> Z (N+1,M+1)."""
    N, M = S.shape
    C = np.full((N + 1, M + 1), F32(alpha), F32)
    C[:N, :M] = S
    norm = F32(-np.log(F32(N + M)))
    log_mu = np.concatenate([np.full(N, norm, F32), [np.log(F32(M)) + norm]]).astype(F32)
    log_nu = np.concatenate([np.full(M, norm, F32), [np.log(F32(N)) + norm]]).astype(F32)

    def lse(a, axis):
        m = a.max(axis, keepdims=True)
        return (np.log(np.exp(a - m, dtype=F32).sum(axis, dtype=F32)) + m.squeeze(axis)).astype(F32)

    u = np.zeros(N + 1, F32)
    v = np.zeros(M + 1, F32)
    for _ in range(iters):
        u = log_mu - lse(C + v[None], 1)
        v = log_nu - lse(C + u[:, None], 0)
    return (C + u[:, None] + v[None] - norm).astype(F32)


def match_select(Z, thr):
    """superglue_test.py:268-285 for one pair."""
    Zi = Z[:-1, :-1]
    i0 = Zi.argmax(1)
    i1 = Zi.argmax(0)
    m0v = Zi.max(1)
    mutual0 = np.arange(Zi.shape[0]) == i1[i0]
    mutual1 = np.arange(Zi.shape[1]) == i0[i1]
    ms0 = np.where(mutual0, np.exp(m0v, dtype=F32), F32(0)).astype(F32)
    ms1 = np.where(mutual1, ms0[i1], F32(0)).astype(F32)
    valid0 = mutual0 & (ms0 > F32(thr))
    valid1 = mutual1 & valid0[i1]
    return (np.where(valid0, i0, -1).astype(np.int64), np.where(valid1, i1, -1).astype(np.int64),
            ms0, ms1)


def superglue_forward(kp0, sc0, de0, kp1, sc1, de1, H, W, sd, cfg, want=()):
    """One pair. kp (n,2), sc (n,), de (D,n).  superglue_test.py:230-285."""
    if kp0.shape[0] == 0 or kp1.shape[0] == 0:
        return {"matches0": np.full(kp0.shape[0], -1, np.int32),
                "matches1": np.full(kp1.shape[0], -1, np.int32),
                "matching_scores0": np.zeros(kp0.shape[0], F32),
                "matching_scores1": np.zeros(kp1.shape[0], F32)}
    D = de0.shape[0]
    n_kenc = len(cfg.get("keypoint_encoder", [32, 64, 128])) + 1
    names = cfg.get("GNN_layers", ["self", "cross"] * 9)
    d0 = de0 + keypoint_encoder(normalize_keypoints(kp0, H, W), sc0, sd, n_kenc)
    d1 = de1 + keypoint_encoder(normalize_keypoints(kp1, H, W), sc1, sd, n_kenc)
    out = {}
    if "kenc" in want:
        out["kenc0"], out["kenc1"] = d0.copy(), d1.copy()
    for l, name in enumerate(names):
        p = f"gnn.layers.{l}"
        s0, s1 = (d1, d0) if name == "cross" else (d0, d1)
        delta0 = attentional_propagation(d0, s0, sd, p)
        delta1 = attentional_propagation(d1, s1, sd, p)
        d0, d1 = d0 + delta0, d1 + delta1
    if "gnn" in want:
        out["gnn0"], out["gnn1"] = d0.copy(), d1.copy()
    m0 = conv1x1(d0, sd["final_proj.weight"], sd["final_proj.bias"])
    m1 = conv1x1(d1, sd["final_proj.weight"], sd["final_proj.bias"])
    S = ((m0.T @ m1) / F32(D ** 0.5)).astype(F32)
    Z = log_optimal_transport(S, sd["bin_score"], cfg.get("sinkhorn_iterations", 100))
    a, b, c, d = match_select(Z, cfg.get("match_threshold", 0.2))
    out.update({"matches0": a, "matches1": b, "matching_scores0": c, "matching_scores1": d})
    if "S" in want:
        out["S"] = S
    if "Z" in want:
        out["Z"] = Z
    return out


def matching_forward(img0, img1, sp_sd, sg_sd, cfg, align_corners=False, want=()):
    """superglue/models/matching_test.py:54-82 for one pair of (H,W) images."""
    p0 = superpoint_forward(img0, sp_sd, cfg["superpoint"], align_corners)
    p1 = superpoint_forward(img1, sp_sd, cfg["superpoint"], align_corners)
    H, W = img0.shape[-2:]
    r = superglue_forward(p0["keypoints"], p0["scores"], p0["descriptors"],
                          p1["keypoints"], p1["scores"], p1["descriptors"],
                          H, W, sg_sd, cfg["superglue"], want)
    r.update({"keypoints0": p0["keypoints"], "scores0": p0["scores"], "descriptors0": p0["descriptors"],
              "keypoints1": p1["keypoints"], "scores1": p1["scores"], "descriptors1": p1["descriptors"]})
    return r
