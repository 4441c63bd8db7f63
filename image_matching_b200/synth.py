"""Seeded synthetic inputs and weights for the SuperPoint+SuperGlue hot path.

Everything here is numpy-only and deterministic across machines, so the same
call produces the same bytes in the build container (where the golden vectors
are generated against the reference) and on the GPU box (where the reference
does not exist).

* ``make_pair``           -- SURVEY.md section 8(d) "random filled rectangles" image pair.
* ``superpoint_weights``  -- state_dict with the key names/shapes the reference's
                             ``SuperPoint`` owns (superpoint/models/superpoint_test.py:64-84,
                             superpoint/models/unet_parts.py:10-48).
* ``superglue_weights``   -- state_dict for the reference's ``SuperGlue``
                             (superglue/models/superglue_test.py:204-219).
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_image", "make_pair", "make_pair_batch", "superpoint_weights",
           "superglue_weights", "random_features"]


# --------------------------------------------------------------------------- images
def _gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gaussian blur with reflect-101 borders (numpy only, fp32)."""
    r = max(1, int(np.ceil(3.0 * sigma)))
    t = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k = (k / k.sum()).astype(np.float32)
    pad = np.pad(img, ((r, r), (r, r)), mode="reflect")
    tmp = np.zeros((img.shape[0] + 2 * r, img.shape[1]), np.float32)
    for i, kv in enumerate(k):
        tmp += kv * pad[:, i:i + img.shape[1]]
    out = np.zeros_like(img, dtype=np.float32)
    for i, kv in enumerate(k):
        out += kv * tmp[i:i + img.shape[0], :]
    return out


def make_image(seed: int, H: int = 480, W: int = 640) -> np.ndarray:
    """Random filled rectangles on a 0.5 canvas, blurred, clipped to [0,1]; float32 (H,W)."""
    rng = np.random.default_rng(seed)
    n = max(8, int(round(600 * (H * W) / (480.0 * 640.0))))
    img = np.full((H, W), 0.5, np.float32)
    x0 = rng.integers(0, max(1, W - 20), n)
    y0 = rng.integers(0, max(1, H - 20), n)
    ws = rng.integers(6, 40, n)
    hs = rng.integers(6, 40, n)
    g = rng.random(n).astype(np.float32)
    for i in range(n):
        img[y0[i]:y0[i] + hs[i], x0[i]:x0[i] + ws[i]] = g[i]
    img = _gauss_blur(img, 0.7)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def _homography(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Direct linear transform for 4 point pairs (float64)."""
    A = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y, -u])
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y, -v])
    _, _, vt = np.linalg.svd(np.asarray(A, np.float64))
    Hm = vt[-1].reshape(3, 3)
    return Hm / Hm[2, 2]


def _warp_perspective(img: np.ndarray, Hm: np.ndarray) -> np.ndarray:
    """Inverse-map bilinear warp with reflect borders (numpy only)."""
    H, W = img.shape
    Hi = np.linalg.inv(Hm)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    den = Hi[2, 0] * xs + Hi[2, 1] * ys + Hi[2, 2]
    sx = (Hi[0, 0] * xs + Hi[0, 1] * ys + Hi[0, 2]) / den
    sy = (Hi[1, 0] * xs + Hi[1, 1] * ys + Hi[1, 2]) / den

    def refl(v, n):
        v = np.abs(v)
        period = 2 * (n - 1)
        v = np.mod(v, period)
        return np.where(v > n - 1, period - v, v)

    sx = refl(sx, W)
    sy = refl(sy, H)
    x0 = np.clip(np.floor(sx).astype(np.int64), 0, W - 1)
    y0 = np.clip(np.floor(sy).astype(np.int64), 0, H - 1)
    x1 = np.minimum(x0 + 1, W - 1)
    y1 = np.minimum(y0 + 1, H - 1)
    fx = (sx - x0).astype(np.float32)
    fy = (sy - y0).astype(np.float32)
    out = (img[y0, x0] * (1 - fx) * (1 - fy) + img[y0, x1] * fx * (1 - fy)
           + img[y1, x0] * (1 - fx) * fy + img[y1, x1] * fx * fy)
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def make_pair(seed: int, H: int = 480, W: int = 640):
    """(image0, image1): image1 is a perspective warp of image0 (corner jitter in [-30,30))."""
    img0 = make_image(seed, H, W)
    rng = np.random.default_rng(seed + 10000)
    jit = min(30, max(2, min(H, W) // 8))
    src = np.array([[0, 0], [W - 1, 0], [W - 1, H - 1], [0, H - 1]], np.float64)
    dst = src + rng.integers(-jit, jit, (4, 2)).astype(np.float64)
    img1 = _warp_perspective(img0, _homography(src, dst))
    return img0, img1


def make_pair_batch(seeds, H: int = 480, W: int = 640):
    """Stack pairs into two float32 arrays shaped (B,1,H,W) -- the Matching.forward layout."""
    a, b = zip(*[make_pair(int(s), H, W) for s in seeds])
    return np.stack(a)[:, None], np.stack(b)[:, None]


# --------------------------------------------------------------------------- weights
def _bn(rng, c, prefix, sd):
    sd[prefix + ".weight"] = rng.uniform(0.6, 1.4, c).astype(np.float32)
    sd[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(1000, np.int64)


def _conv(rng, cout, cin, k, prefix, sd, gain=2.0, ndim=2):
    fan_in = cin * k * k
    std = np.sqrt(gain / fan_in)
    shape = (cout, cin, k, k) if ndim == 2 else (cout, cin, 1)
    sd[prefix + ".weight"] = (std * rng.standard_normal(shape)).astype(np.float32)
    sd[prefix + ".bias"] = (0.05 * rng.standard_normal(cout)).astype(np.float32)


def superpoint_weights(seed: int = 0, descriptor_dim: int = 128) -> dict:
    """Synthetic SuperPoint state_dict (84 tensors, same names as the reference module)."""
    rng = np.random.default_rng(1_000_003 + seed)
    sd: dict = {}
    c1, c2, c3, c4, c5 = 64, 64, 128, 128, 256

    def dconv(prefix, cin, cout):
        _conv(rng, cout, cin, 3, prefix + ".0", sd)
        _bn(rng, cout, prefix + ".1", sd)
        _conv(rng, cout, cout, 3, prefix + ".3", sd)
        _bn(rng, cout, prefix + ".4", sd)

    dconv("inc.conv.conv", 1, c1)
    dconv("down1.mpconv.1.conv", c1, c2)
    dconv("down2.mpconv.1.conv", c2, c3)
    dconv("down3.mpconv.1.conv", c3, c4)
    _conv(rng, c5, c4, 3, "convPa", sd)
    _bn(rng, c5, "bnPa", sd)
    # a larger gain on the detector logits gives a peaky heat-map like a trained detector
    _conv(rng, 65, c5, 1, "convPb", sd, gain=16.0)
    _bn(rng, 65, "bnPb", sd)
    _conv(rng, c5, c4, 3, "convDa", sd)
    _bn(rng, c5, "bnDa", sd)
    _conv(rng, descriptor_dim, c5, 1, "convDb", sd, gain=1.0)
    _bn(rng, descriptor_dim, "bnDb", sd)
    return sd


_OFFICIAL_NAMES = {"inc.conv.conv.0": "conv1a", "inc.conv.conv.3": "conv1b",
                   "down1.mpconv.1.conv.0": "conv2a", "down1.mpconv.1.conv.3": "conv2b",
                   "down2.mpconv.1.conv.0": "conv3a", "down2.mpconv.1.conv.3": "conv3b",
                   "down3.mpconv.1.conv.0": "conv4a", "down3.mpconv.1.conv.3": "conv4b",
                   "convPa": "convPa", "convPb": "convPb", "convDa": "convDa", "convDb": "convDb"}
_OFFICIAL_BN = {"inc.conv.conv.0": "inc.conv.conv.1", "inc.conv.conv.3": "inc.conv.conv.4",
                "down1.mpconv.1.conv.0": "down1.mpconv.1.conv.1", "down1.mpconv.1.conv.3": "down1.mpconv.1.conv.4",
                "down2.mpconv.1.conv.0": "down2.mpconv.1.conv.1", "down2.mpconv.1.conv.3": "down2.mpconv.1.conv.4",
                "down3.mpconv.1.conv.0": "down3.mpconv.1.conv.1", "down3.mpconv.1.conv.3": "down3.mpconv.1.conv.4",
                "convPa": "bnPa", "convPb": "bnPb", "convDa": "bnDa", "convDb": "bnDb"}


def superpoint_official_weights(seed: int = 0, descriptor_dim: int = 256) -> dict:
    """Synthetic state_dict of the BatchNorm-free "official" SuperPoint (reference superglue/models/superpoint.py:
    keys conv1a .. conv4b, convPa/Pb/Da/Db): the synthetic weights above with their BatchNorm folded into the
    convolutions, so the activations keep sane scales without normalisation layers."""
    sd = superpoint_weights(seed, descriptor_dim)
    out = {}
    for conv, name in _OFFICIAL_NAMES.items():
        bn = _OFFICIAL_BN[conv]
        s = sd[bn + ".weight"].astype(np.float64) / np.sqrt(sd[bn + ".running_var"].astype(np.float64) + 1e-5)
        w = sd[conv + ".weight"].astype(np.float64) * s.reshape(-1, 1, 1, 1)
        b = (sd[conv + ".bias"].astype(np.float64) - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]
        out[name + ".weight"] = w.astype(np.float32)
        out[name + ".bias"] = b.astype(np.float32)
    return out


def superglue_weights(seed: int = 0, descriptor_dim: int = 128,
                      keypoint_encoder=(32, 64, 128), n_layers: int = 18,
                      sharpen: float = 16.0) -> dict:
    """Synthetic SuperGlue state_dict.  The residual MLPs get a small gain and
    final_proj is ``sharpen * (I + 0.1 * noise)`` so that the optimal-transport scores
    are peaky enough to produce hundreds of real matches (a raw default
    init gives a flat score matrix and zero valid matches, which would make
    match-index parity vacuous)."""
    rng = np.random.default_rng(2_000_003 + seed)
    D = descriptor_dim
    sd: dict = {"bin_score": np.array(1.0, np.float32)}
    ch = [3] + list(keypoint_encoder) + [D]
    idx = 0
    for i in range(1, len(ch)):
        _conv(rng, ch[i], ch[i - 1], 1, f"kenc.encoder.{idx}", sd, gain=1.0, ndim=1)
        idx += 1
        if i < len(ch) - 1:
            _bn(rng, ch[i], f"kenc.encoder.{idx}", sd)
            idx += 2  # BN then ReLU
    sd[f"kenc.encoder.{idx - 1}.bias"][:] = 0.0
    for l in range(n_layers):
        p = f"gnn.layers.{l}"
        _conv(rng, D, D, 1, p + ".attn.merge", sd, gain=1.0, ndim=1)
        # q/k projections get a larger gain so the attention softmax is not flat
        _conv(rng, D, D, 1, p + ".attn.proj.0", sd, gain=4.0, ndim=1)
        _conv(rng, D, D, 1, p + ".attn.proj.1", sd, gain=4.0, ndim=1)
        _conv(rng, D, D, 1, p + ".attn.proj.2", sd, gain=1.0, ndim=1)
        _conv(rng, 2 * D, 2 * D, 1, p + ".mlp.0", sd, gain=1.0, ndim=1)
        _bn(rng, 2 * D, p + ".mlp.1", sd)
        _conv(rng, D, 2 * D, 1, p + ".mlp.3", sd, gain=0.05, ndim=1)
        sd[p + ".mlp.3.bias"][:] = 0.0
    _conv(rng, D, D, 1, "final_proj", sd, gain=1.0, ndim=1)
    sd["final_proj.weight"] = (np.float32(sharpen) * (np.eye(D, dtype=np.float32)[:, :, None]
                                                      + np.float32(0.1) * sd["final_proj.weight"])).astype(np.float32)
    return sd


def random_features(seed: int, B: int, N: int, D: int, H: int, W: int):
    """Config-5 style inputs: uniform keypoints, U(0,1) scores, unit-norm descriptors.

    Returns (keypoints (B,N,2) xy float32, scores (B,N), descriptors (B,D,N))."""
    rng = np.random.default_rng(3_000_003 + seed)
    kp = np.stack([rng.integers(4, W - 4, (B, N)), rng.integers(4, H - 4, (B, N))], -1)
    sc = rng.random((B, N)).astype(np.float32)
    de = rng.standard_normal((B, D, N)).astype(np.float32)
    de /= np.linalg.norm(de, axis=1, keepdims=True)
    return kp.astype(np.float32), sc, de.astype(np.float32)
