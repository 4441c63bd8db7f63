"""Registration step that follows ``Matching.forward`` in the reference's caller:

    superpoint_glue_test.py:79-92   kpts/matches -> cv2.estimateAffinePartial2D(RANSAC, 7 px) -> Matrix, mask
    superpoint_glue_test.py:101     cv2.warpAffine(source_original, Matrix, (w, h))

The reference pulls four tensors to the host and runs OpenCV per pair inside its timed region; here both steps run on
the device for the whole batch, straight from the padded outputs of ``Matching.forward_device`` (no host sync), through
``b200m_estimate_affine_partial`` / ``b200m_warp_affine``.  The inlier mask is identical to cv2 4.13's and the matrix
agrees to ~1e-12 (see csrc/sp_register.cu); the warp is bit-identical.  CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import torch

from . import lib as _lib
from .matching import _ptr, _stream

_DTYPES = {torch.uint8: 0, torch.float32: 1, torch.float64: 2}


def _handle(owner, device):
    """`owner` holds the library handle: a Matching (or MatchingOfficial) or a stand-alone SuperPoint of this package."""
    if hasattr(owner, "_ensure"):
        L = owner._ensure(device)
    else:
        L = owner._engine.ensure(device, owner, None)
    return L, owner._engine.handle


def estimate_affine_partial_2d(owner, kpts0: torch.Tensor, kpts1: torch.Tensor, matches0: torch.Tensor,
                               counts0: torch.Tensor | None = None, ransac_reproj_threshold: float = 3.0,
                               max_iters: int = 2000, confidence: float = 0.99, refine_iters: int = 10):
    """Batched ``cv2.estimateAffinePartial2D(kpts0[valid], kpts1[matches0[valid]], method=cv2.RANSAC, ...)``.

    kpts0 (B,N,2) / kpts1 (B,M,2) float32 xy, matches0 (B,N) integer with -1 = unmatched, counts0 (B) int32 valid
    keypoints per pair (padded batches) or None.  Returns device tensors, no host synchronisation:
    matrices (B,2,3) float64, inlier0 (B,N) uint8 (cv2's mask scattered to keypoint indices of image0),
    info (B,4) int32 = [correspondences, inliers, RANSAC iterations, found]."""
    if kpts0.device.type != "cuda":
        raise RuntimeError("image_matching_b200 runs on CUDA (sm_100a) only -- there is no CPU fallback")
    k0 = kpts0.contiguous().float()
    k1 = kpts1.contiguous().float()
    m0 = matches0.contiguous().to(torch.int64)
    if k0.dim() != 3 or k1.dim() != 3 or k0.shape[2] != 2 or k1.shape[2] != 2 or m0.shape != k0.shape[:2] \
            or k1.shape[0] != k0.shape[0]:
        raise ValueError("expected kpts0 (B,N,2), kpts1 (B,M,2), matches0 (B,N)")
    dev = k0.device
    B, N, M = k0.shape[0], k0.shape[1], k1.shape[1]
    L, h = _handle(owner, dev)
    mats = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
    inl = torch.empty((B, N), dtype=torch.uint8, device=dev)
    info = torch.empty((B, 4), dtype=torch.int32, device=dev)
    if B == 0 or N == 0 or M == 0:      # nothing to estimate from (cv2: Matrix None, empty / all-zero mask)
        return mats.zero_(), inl.zero_(), info.zero_()
    if counts0 is not None:
        counts0 = counts0.contiguous().to(torch.int32)
    _lib.check(L.b200m_estimate_affine_partial(h, _ptr(k0), _ptr(k1), _ptr(m0), _ptr(counts0), B, N, M,
                                               float(ransac_reproj_threshold), int(max_iters), float(confidence),
                                               int(refine_iters), _ptr(mats), _ptr(inl), _ptr(info), _stream(dev)),
               "b200m_estimate_affine_partial")
    return mats, inl, info


def warp_affine(owner, src: torch.Tensor, matrices: torch.Tensor, dsize: tuple[int, int] | None = None):
    """Batched ``cv2.warpAffine(src[b], matrices[b], dsize)`` (INTER_LINEAR, constant 0 border): src (B,H,W) or (H,W)
    uint8 / float32 / float64 on the device, matrices (B,2,3) or (2,3); dsize = (width, height) like cv2 (default: the
    source size).  Returns the warped images in the source dtype."""
    if src.device.type != "cuda":
        raise RuntimeError("image_matching_b200 runs on CUDA (sm_100a) only -- there is no CPU fallback")
    if src.dtype not in _DTYPES:
        raise ValueError(f"warp_affine supports uint8 / float32 / float64 images, got {src.dtype}")
    single = src.dim() == 2
    s = (src[None] if single else src).contiguous()
    m = matrices.to(device=s.device, dtype=torch.float64).reshape(-1, 2, 3).contiguous()
    if s.dim() != 3 or m.shape[0] != s.shape[0]:
        raise ValueError("expected src (B,H,W) and one 2x3 matrix per image")
    B, H, W = s.shape
    dw, dh = (W, H) if dsize is None else (int(dsize[0]), int(dsize[1]))
    L, h = _handle(owner, s.device)
    dst = torch.empty((B, dh, dw), dtype=s.dtype, device=s.device)
    _lib.check(L.b200m_warp_affine(h, _ptr(s), _DTYPES[s.dtype], B, H, W, _ptr(m), _ptr(dst), dh, dw, _stream(s.device)),
               "b200m_warp_affine")
    return dst[0] if single else dst


def register_pairs(owner, pred: dict, ransac_reproj_threshold: float = 7.0, resize_scale: float | None = None):
    """The caller's post-processing, superpoint_glue_test.py:79-92, for every pair of a ``Matching.forward`` result
    (or a ``Matching.forward_device`` result).  Returns {"pairs": [...], "matrices", "inlier0", "info"}: one dict per
    pair with the names the script uses -- ``Matrix`` (2,3) float64 numpy or None (fewer than 4 matches: the script
    skips the estimate), ``mask`` (n,1) uint8, ``mkpts0`` / ``mkpts1`` (the inlier correspondences), ``valid`` -- plus
    the device tensors of ``estimate_affine_partial_2d`` for a following ``warp_affine``."""
    k0, k1, m0 = pred["keypoints0"], pred["keypoints1"], pred["matches0"]
    counts = pred.get("counts")
    if isinstance(k0, (list, tuple)):
        k0, k1 = torch.stack(list(k0)), torch.stack(list(k1))
    mats, inl, info = estimate_affine_partial_2d(owner, k0, k1, m0, None if counts is None else counts[0],
                                                 ransac_reproj_threshold)
    cnt = None if counts is None else counts[0].cpu().tolist()
    if resize_scale is not None:                       # Matrix[:,2] = Matrix[:,2] / resize_scale (:89-90)
        mats = mats.clone()
        mats[:, :, 2] /= resize_scale
    k0h, k1h, m0h = k0.cpu().numpy(), k1.cpu().numpy(), m0.cpu().numpy()
    math_, inlh, infoh = mats.cpu().numpy(), inl.cpu().numpy(), info.cpu().numpy()
    out = []
    for b in range(k0h.shape[0]):
        n = k0h.shape[1] if cnt is None else cnt[b]
        matches = m0h[b, :n]
        valid = matches > -1
        mk0, mk1 = k0h[b, :n][valid], k1h[b][matches[valid]]
        r = {"valid": valid, "Matrix": None, "mask": None, "mkpts0": mk0, "mkpts1": mk1,
             "iterations": int(infoh[b, 2])}
        if len(mk0) > 3 and infoh[b, 3]:
            mask = inlh[b, :n][valid].reshape(-1, 1)
            flag = mask.ravel() > 0
            r.update({"Matrix": math_[b].copy(), "mask": mask, "mkpts0": mk0[flag], "mkpts1": mk1[flag]})
        out.append(r)
    return {"pairs": out, "matrices": mats, "inlier0": inl, "info": info}


__all__ = ["estimate_affine_partial_2d", "warp_affine", "register_pairs"]


def resize_u8(owner, images: torch.Tensor, resize_scale: float | None = None, dsize: tuple[int, int] | None = None):
    """The data loader's resize on the device (datasets/SSHIDataset.py:20-22):
    ``cv2.resize(img, (int(resize_scale * W), int(resize_scale * H)))`` for uint8 grayscale images (H,W), (B,H,W) or
    (B,1,H,W) -- bit-identical to OpenCV 4.13's INTER_LINEAR.  ``dsize`` = (width, height) overrides the scale.  The
    uint8 result can be passed straight to ``Matching`` / ``SuperPoint`` (their first kernel divides by 255 while loading
    the pixels, as SSHIDataset.py:26-28 + the caller's ``.float()`` do).  CUDA only -- no CPU fallback."""
    if images.device.type != "cuda":
        raise RuntimeError("image_matching_b200 runs on CUDA (sm_100a) only -- there is no CPU fallback")
    if images.dtype != torch.uint8:
        raise ValueError(f"resize_u8 takes uint8 images (the loader resizes before normalising), got {images.dtype}")
    shape = images.shape
    if images.dim() == 4 and shape[1] != 1:
        raise ValueError("expected single-channel images")
    x = images.reshape(-1, shape[-2], shape[-1]).contiguous()
    B, H, W = x.shape
    if dsize is None:
        if resize_scale is None:
            return images
        dsize = (int(resize_scale * W), int(resize_scale * H))
    dw, dh = int(dsize[0]), int(dsize[1])
    if dw <= 0 or dh <= 0:
        raise ValueError("empty destination size")
    L, h = _handle(owner, x.device)
    dst = torch.empty((B, dh, dw), dtype=torch.uint8, device=x.device)
    _lib.check(L.b200m_resize_linear_u8(h, _ptr(x), B, H, W, _ptr(dst), dh, dw, _stream(x.device)), "b200m_resize_linear_u8")
    return dst.reshape(tuple(shape[:-2]) + (dh, dw))
