"""CPU: the registration oracle (oracle/registration_oracle.py: numpy restatement of cv2.estimateAffinePartial2D(RANSAC)
and cv2.warpAffine as the reference's caller uses them, superpoint_glue_test.py:88,101) against golden vectors generated
with cv2 4.13.0 (tests/golden/make_golden_registration.py) and, when cv2 is importable, against cv2 live."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import registration_oracle as R


def test_ransac_matches_cv2_golden():
    g = load_golden("registration")
    for i in range(int(g["n_ransac"])):
        fr, to = g[f"ransac{i}_from"], g[f"ransac{i}_to"]
        M, mask, iters = R.estimate_affine_partial_2d(fr, to, 7.0)
        gM, gmask = g[f"ransac{i}_M"], g[f"ransac{i}_mask"]
        if gM.size == 0:
            assert M is None and not mask.any()
            continue
        assert np.array_equal(mask, gmask), i          # inlier set: identical
        assert np.abs(M - gM).max() < 1e-9, i          # matrix: closed-form fixed point of cv2's LM refinement
        assert iters >= 1


def test_warp_affine_matches_cv2_golden_bit_exact():
    g = load_golden("registration")
    for k in range(int(g["n_warp"])):
        src, M, dst = g[f"warp{k}_src"], g[f"warp{k}_M"], g[f"warp{k}_dst"]
        out = R.warp_affine(src, M, (dst.shape[1], dst.shape[0]))
        assert out.dtype == dst.dtype and np.array_equal(out, dst), k


def test_edge_cases():
    fr = np.array([[10, 10], [50, 20]], np.float32)
    to = fr * 2 + 3
    M, mask, it = R.estimate_affine_partial_2d(fr, to, 7.0)          # exactly two points: the exact model, no RANSAC
    assert it == 0 and mask.all() and np.allclose(M, [[2, 0, 3], [0, 2, 3]])
    M, mask, it = R.estimate_affine_partial_2d(fr[:1], to[:1], 7.0)
    assert M is None and mask.shape == (1, 1) and not mask.any()
    # the caller skips the estimate below four matches (superpoint_glue_test.py:86)
    k0 = np.array([[1, 1], [2, 5], [7, 3]], np.float32)
    M, mask, mk0, mk1 = R.register_pair(k0, k0 + 1, np.array([0, 1, 2]))
    assert M is None and len(mk0) == 3
    # resize_scale divides the translation only (:89-90)
    rs = np.random.default_rng(0)
    k0 = rs.integers(0, 600, (50, 2)).astype(np.float32)
    M1, _, _, _ = R.register_pair(k0, k0 + np.float32([8, -4]), np.arange(50))
    M2, _, _, _ = R.register_pair(k0, k0 + np.float32([8, -4]), np.arange(50), resize_scale=0.5)
    assert np.allclose(M1[:, :2], M2[:, :2]) and np.allclose(M2[:, 2], M1[:, 2] / 0.5)
    assert np.allclose(M1, [[1, 0, 8], [0, 1, -4]], atol=1e-9)


def test_live_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.default_rng(7)
    for trial in range(40):
        n = int(rs.integers(4, 600))
        fr = np.stack([rs.integers(4, 636, n), rs.integers(4, 476, n)], 1).astype(np.float32)
        ang, s = rs.uniform(-0.5, 0.5), rs.uniform(0.7, 1.3)
        A = np.array([[s * np.cos(ang), -s * np.sin(ang), 11.0], [s * np.sin(ang), s * np.cos(ang), -23.0]])
        to = fr @ A[:, :2].T + A[:, 2] + rs.normal(0, rs.uniform(0, 3), (n, 2))
        o = rs.random(n) < rs.uniform(0, 0.8)
        to[o] = rs.uniform(0, 480, (int(o.sum()), 2))
        to = np.round(to).astype(np.float32)
        thr = float(rs.choice([3.0, 7.0]))
        cM, cmask = cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ransacReprojThreshold=thr)
        M, mask, _ = R.estimate_affine_partial_2d(fr, to, thr)
        assert (cM is None) == (M is None)
        assert np.array_equal(mask, cmask)
        if M is not None:
            assert np.abs(M - cM).max() < 1e-9
    for dt in (np.float64, np.float32, np.uint8):
        src = (rs.random((77, 101)) * 255).astype(dt)
        M = np.array([[0.9, -0.2, 5.5], [0.2, 0.9, -7.25]])
        assert np.array_equal(R.warp_affine(src, M, (90, 80)), cv2.warpAffine(src, M, (90, 80)))
