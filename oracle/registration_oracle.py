"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the registration step that follows Matching.forward in the
reference's caller, superpoint_glue_test.py:83-92,101:

    Matrix, mask = cv2.estimateAffinePartial2D(mkpts0, mkpts1, method=cv2.RANSAC, ransacReprojThreshold=7)
    Transform    = cv2.warpAffine(source_original, Matrix, (w, h))

The arithmetic lives in a third-party dependency that the reference does not vendor or pin (`cv2`; `requirements.txt`
pins nothing; this image has opencv-python 4.13.0).  Its published algorithm (modules/calib3d/src/ptsetreg.cpp,
modules/imgproc/src/imgwarp.cpp) is restated below and PINNED against outputs of cv2 itself:
tests/golden/make_golden_registration.py runs cv2 4.13.0 in the build container and commits
tests/golden/registration.npz; tests/test_oracle_golden.py checks this file against those vectors (inlier masks and
warped images bit-identical, matrices <= 1e-9), and live against cv2 when it is importable.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(image_matching_b200.registration -> libb200match.so) never does.
"""
import math

import numpy as np

f32 = np.float32
_DBL_MIN = 2.2250738585072014e-308


class CvRNG:
    """cv::RNG (core/operations.hpp): multiply-with-carry, `RNG rng((uint64)-1)` in RANSACPointSetRegistrator::run."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def similarity_from_two(fr, to):
    """AffinePartial2DEstimatorCallback::runKernel: [[a,-b,tx],[b,a,ty]] through two correspondences (doubles)."""
    x1, y1, x2, y2 = (float(v) for v in (fr[0, 0], fr[0, 1], fr[1, 0], fr[1, 1]))
    X1, Y1, X2, Y2 = (float(v) for v in (to[0, 0], to[0, 1], to[1, 0], to[1, 1]))
    d = 1.0 / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2))
    S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
    S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
    S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2))
    S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return np.array([S0, -S1, S2, S1, S0, S3])


def reprojection_errors(fr, to, M):
    """Affine2DEstimatorCallback::computeError: squared distances in fp32, evaluated left to right."""
    F = M.astype(f32)
    a = F[0] * fr[:, 0] + F[1] * fr[:, 1] + F[2] - to[:, 0]
    b = F[3] * fr[:, 0] + F[4] * fr[:, 1] + F[5] - to[:, 1]
    return a * a + b * b


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """cv::RANSACUpdateNumIters."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, _DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < _DBL_MIN:
        return 0
    num, denom = math.log(num), math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def least_squares_similarity(fr, to):
    """Fixed point of cv2's Levenberg-Marquardt refinement: the 4-parameter similarity minimising the summed squared
    reprojection error (a linear problem), in doubles, mean-centred."""
    x, X = fr.astype(np.float64), to.astype(np.float64)
    mx, mX = x.mean(0), X.mean(0)
    xc, Xc = x - mx, X - mX
    den = (xc ** 2).sum()
    a = (xc[:, 0] * Xc[:, 0] + xc[:, 1] * Xc[:, 1]).sum() / den
    b = (xc[:, 0] * Xc[:, 1] - xc[:, 1] * Xc[:, 0]).sum() / den
    return np.array([[a, -b, mX[0] - (a * mx[0] - b * mx[1])], [b, a, mX[1] - (b * mx[0] + a * mx[1])]])


def estimate_affine_partial_2d(fr, to, ransac_reproj_threshold=3.0, max_iters=2000, confidence=0.99, refine_iters=10):
    """cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ...) -> (Matrix (2,3) float64 | None, mask (n,1) uint8,
    iterations).  RANSACPointSetRegistrator::run with modelPoints = 2, then the refinement over the inliers."""
    fr = np.asarray(fr, f32).reshape(-1, 2)
    to = np.asarray(to, f32).reshape(-1, 2)
    count = len(fr)
    zero = np.zeros((count, 1), np.uint8)
    if count < 2:
        return None, zero, 0
    if count == 2:
        return similarity_from_two(fr, to).reshape(2, 3), np.ones((2, 1), np.uint8), 0
    rng = CvRNG()
    niters, best, best_model, best_mask = max(max_iters, 1), 0, None, None
    t = f32(ransac_reproj_threshold * ransac_reproj_threshold)
    it = 0
    while it < niters:
        idx = []
        for _ in range(2):                      # getSubset: draw again on a duplicate; 2 points are never degenerate
            v = rng.uniform(0, count)
            while v in idx:
                v = rng.uniform(0, count)
            idx.append(v)
        model = similarity_from_two(fr[idx], to[idx])
        mask = reprojection_errors(fr, to, model) <= t
        good = int(mask.sum())
        if good > max(best, 1):
            best, best_model, best_mask = good, model, mask
            niters = ransac_update_num_iters(confidence, (count - good) / count, 2, niters)
        it += 1
    if best <= 0:
        return None, zero, it
    M = best_model.reshape(2, 3)
    if refine_iters:
        M = least_squares_similarity(fr[best_mask], to[best_mask])
    return M, best_mask.astype(np.uint8).reshape(-1, 1), it


def register_pair(kpts0, kpts1, matches0, ransac_reproj_threshold=7.0, resize_scale=None):
    """superpoint_glue_test.py:83-92 for one pair."""
    matches0 = np.asarray(matches0)
    valid = matches0 > -1
    mk0, mk1 = np.asarray(kpts0)[valid], np.asarray(kpts1)[matches0[valid]]
    if len(mk0) <= 3:
        return None, None, mk0, mk1
    M, mask, _ = estimate_affine_partial_2d(mk0, mk1, ransac_reproj_threshold)
    if M is None:
        return None, mask, mk0, mk1
    if resize_scale is not None:
        M = M.copy()
        M[:, 2] = M[:, 2] / resize_scale
    flag = mask.ravel() > 0
    return M, mask, mk0[flag], mk1[flag]


# ------------------------------------------------------------------------------------------------ warpAffine
def invert_affine(M):
    """cv::warpAffine without WARP_INVERSE_MAP: the 2x3 inverse in doubles, in OpenCV's operation order."""
    M = np.asarray(M, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def _sat_int(v):
    return np.clip(np.rint(v), -2147483648, 2147483647).astype(np.int64)


def warp_affine(src, M, dsize=None):
    """cv2.warpAffine(src, M, dsize) with the defaults (INTER_LINEAR, BORDER_CONSTANT 0) for one single-channel uint8 /
    float32 / float64 image: AB_BITS = 10 fixed-point source coordinates, INTER_BITS = 5 interpolation grid, bilinear
    table weights in fp32 (uint8: 15-bit integer weights, rounded)."""
    src = np.asarray(src)
    sh, sw = src.shape
    W, H = (sw, sh) if dsize is None else dsize
    Mi = invert_affine(M)
    x, y = np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64)
    adelta, bdelta = _sat_int(Mi[0] * x * 1024), _sat_int(Mi[3] * x * 1024)
    X0 = _sat_int((Mi[1] * y + Mi[2]) * 1024) + 16
    Y0 = _sat_int((Mi[4] * y + Mi[5]) * 1024) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    ax, ay = X & 31, Y & 31

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        return np.where(ok, src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)], 0)

    taps = (tap(sy, sx), tap(sy, sx + 1), tap(sy + 1, sx), tap(sy + 1, sx + 1))
    if src.dtype == np.uint8:
        w = ((32 - ay) * (32 - ax) * 32, (32 - ay) * ax * 32, ay * (32 - ax) * 32, ay * ax * 32)
        s = sum(t.astype(np.int64) * wi for t, wi in zip(taps, w))
        return np.clip((s + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    fx, fy = ax.astype(f32) * f32(1 / 32), ay.astype(f32) * f32(1 / 32)
    gx, gy = f32(1) - fx, f32(1) - fy
    w = (gy * gx, gy * fx, fy * gx, fy * fx)
    wt = np.float64 if src.dtype == np.float64 else np.float32
    out = taps[0].astype(wt) * w[0].astype(wt)
    for t, wi in zip(taps[1:], w[1:]):
        out = out + t.astype(wt) * wi.astype(wt)
    return out.astype(src.dtype)
