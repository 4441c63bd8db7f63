"""GPU: the registration step after Matching.forward (superpoint_glue_test.py:83-92,101) through the
C ABI (b200m_estimate_affine_partial, b200m_warp_affine) against the cv2-pinned oracle and the cv2 golden vectors:
inlier masks and RANSAC iteration counts identical, matrices <= 1e-9, warped images bit-identical."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg, load_golden
from oracle import registration_oracle as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def owner():
    from image_matching_b200 import SuperPoint, synth
    cfg = golden_cfg()
    sp = SuperPoint(dict(cfg["superpoint"], weights=None)).eval()
    sp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_weights(0, 128).items()})
    return sp.to("cuda:0")


def _problem(rs, n, m, n_matches, noise, outliers):
    """keypoints of two images + a matches0 vector with n_matches valid entries (shuffled targets)."""
    k0 = np.stack([rs.integers(4, 636, n), rs.integers(4, 476, n)], 1).astype(np.float32)
    k1 = np.stack([rs.integers(4, 636, m), rs.integers(4, 476, m)], 1).astype(np.float32)
    ang, s = rs.uniform(-0.5, 0.5), rs.uniform(0.7, 1.3)
    A = np.array([[s * np.cos(ang), -s * np.sin(ang), rs.uniform(-40, 40)],
                  [s * np.sin(ang), s * np.cos(ang), rs.uniform(-40, 40)]])
    matches = np.full(n, -1, np.int64)
    src = rs.choice(n, n_matches, replace=False)
    dst = rs.choice(m, n_matches, replace=False)
    matches[src] = dst
    t = k0[src] @ A[:, :2].T + A[:, 2] + rs.normal(0, noise, (n_matches, 2))
    o = rs.random(n_matches) < outliers
    t[o] = np.stack([rs.uniform(0, 640, int(o.sum())), rs.uniform(0, 480, int(o.sum()))], 1)
    k1[dst] = np.round(t).astype(np.float32)
    return k0, k1, matches


def _check(owner, probs, thr=7.0, counts=None):
    from image_matching_b200 import estimate_affine_partial_2d
    k0 = torch.from_numpy(np.stack([p[0] for p in probs])).cuda()
    k1 = torch.from_numpy(np.stack([p[1] for p in probs])).cuda()
    m0 = torch.from_numpy(np.stack([p[2] for p in probs])).cuda()
    c = None if counts is None else torch.tensor(counts, dtype=torch.int32, device="cuda")
    mats, inl, info = estimate_affine_partial_2d(owner, k0, k1, m0, c, thr)
    mats, inl, info = mats.cpu().numpy(), inl.cpu().numpy(), info.cpu().numpy()
    for b, (a0, a1, mm) in enumerate(probs):
        n = len(mm) if counts is None else counts[b]
        mm = np.where(np.arange(len(mm)) < n, mm, -1)
        valid = mm > -1
        M, mask, iters = R.estimate_affine_partial_2d(a0[valid], a1[mm[valid]], thr)
        assert info[b, 0] == valid.sum()
        assert not inl[b][~valid].any()
        if M is None:
            assert info[b, 3] == 0 and not inl[b].any()
            continue
        assert info[b, 3] == 1 and info[b, 2] == iters, (b, info[b], iters)
        assert np.array_equal(inl[b][valid], mask.ravel()), b
        assert info[b, 1] == mask.sum()
        assert np.abs(mats[b] - M).max() < 1e-9, (b, mats[b], M)
    return info


def test_ransac_matches_oracle_batched(owner):
    rs = np.random.default_rng(11)
    probs = [_problem(rs, 1024, 1024, int(rs.integers(4, 1025)), rs.uniform(0, 3), rs.uniform(0, 0.85))
             for _ in range(48)]
    info = _check(owner, probs)
    assert info[:, 2].max() > 16 and info[:, 2].min() < 16     # both the single-round and the multi-round paths ran


def test_ransac_edge_counts(owner):
    rs = np.random.default_rng(12)
    probs = [_problem(rs, 64, 80, k, 0.5, 0.0) for k in (0, 1, 2, 3, 4, 5, 64)]
    _check(owner, probs)
    # no consensus at all: random targets, tight threshold -> still the best 2-point model, like cv2
    _check(owner, [_problem(rs, 300, 300, 200, 0.0, 1.0)], thr=1.0)
    # all correspondences exact inliers: the adaptive bound stops after the first hypothesis
    info = _check(owner, [_problem(rs, 500, 500, 500, 0.0, 0.0)])
    assert info[0, 2] == 1 and info[0, 1] == 500


def test_ransac_empty_inputs(owner):
    from image_matching_b200 import estimate_affine_partial_2d
    k0 = torch.zeros((2, 0, 2), device="cuda")
    k1 = torch.rand((2, 5, 2), device="cuda")
    mats, inl, info = estimate_affine_partial_2d(owner, k0, k1, torch.zeros((2, 0), dtype=torch.int32, device="cuda"))
    assert mats.shape == (2, 2, 3) and inl.shape == (2, 0) and int(info.abs().sum()) == 0
    with pytest.raises(RuntimeError):
        estimate_affine_partial_2d(owner, k0.cpu(), k1.cpu(), torch.zeros((2, 0), dtype=torch.int64))


def test_ransac_padded_counts_and_large_n(owner):
    rs = np.random.default_rng(13)
    probs = [_problem(rs, 4096, 4096, 3000, 1.0, 0.5) for _ in range(3)]
    _check(owner, probs, counts=[4096, 1000, 0])
    _check(owner, [_problem(rs, 37, 4096, 20, 1.0, 0.3)])


def test_ransac_matches_cv2_golden(owner):
    from image_matching_b200 import estimate_affine_partial_2d
    g = load_golden("registration")
    for i in range(int(g["n_ransac"])):
        fr, to = g[f"ransac{i}_from"], g[f"ransac{i}_to"]
        n = len(fr)
        mats, inl, info = estimate_affine_partial_2d(owner, torch.from_numpy(fr)[None].cuda(),
                                                     torch.from_numpy(to)[None].cuda(),
                                                     torch.arange(n)[None].cuda(), None, 7.0)
        gM, gmask = g[f"ransac{i}_M"], g[f"ransac{i}_mask"]
        if gM.size == 0:
            assert int(info[0, 3]) == 0
            continue
        assert np.array_equal(inl[0].cpu().numpy(), gmask.ravel()), i
        assert np.abs(mats[0].cpu().numpy() - gM).max() < 1e-9, i


def test_warp_affine_bit_exact(owner):
    from image_matching_b200 import warp_affine
    g = load_golden("registration")
    for k in range(int(g["n_warp"])):
        src, M, dst = g[f"warp{k}_src"], g[f"warp{k}_M"], g[f"warp{k}_dst"]
        out = warp_affine(owner, torch.from_numpy(src).cuda(), torch.from_numpy(M), (dst.shape[1], dst.shape[0]))
        assert np.array_equal(out.cpu().numpy(), dst), k
    # full-size batch, one matrix per image (the caller warps the full-resolution float64 source, :98-101)
    rs = np.random.default_rng(3)
    for dt in (np.float64, np.float32, np.uint8):
        src = (rs.random((3, 480, 640)) * 255).astype(dt)
        Ms = np.stack([np.array([[np.cos(a), -np.sin(a), tx], [np.sin(a), np.cos(a), ty]])
                       for a, tx, ty in ((0.1, 12.5, -30.0), (-0.3, 100.0, 50.0), (0.0, 0.0, 0.0))])
        out = warp_affine(owner, torch.from_numpy(src).cuda(), torch.from_numpy(Ms)).cpu().numpy()
        for b in range(3):
            assert np.array_equal(out[b], R.warp_affine(src[b], Ms[b])), (dt, b)


def test_register_pairs_end_to_end():
    """Matching.forward -> register_pairs -> warp_affine, against the oracle on the SAME matches."""
    from image_matching_b200 import Matching, register_pairs, warp_affine, synth
    cfg = golden_cfg(max_kp=512)
    m = Matching({"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")})
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_weights(0, 128).items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superglue_weights(0, 128).items()})
    m = m.eval().to("cuda:0")
    pairs = [synth.make_pair(s, 240, 320) for s in (1, 2)]
    a = torch.from_numpy(np.stack([p[0] for p in pairs])[:, None]).cuda()
    b = torch.from_numpy(np.stack([p[1] for p in pairs])[:, None]).cuda()
    out = m.forward_device(a, b)
    reg = register_pairs(m, out, 7.0, resize_scale=0.5)
    cnt = out["counts"].cpu().numpy()
    n_found = 0
    for i, r in enumerate(reg["pairs"]):
        n0 = int(cnt[0, i])
        M, mask, mk0, mk1 = R.register_pair(out["keypoints0"][i, :n0].cpu().numpy(), out["keypoints1"][i].cpu().numpy(),
                                            out["matches0"][i, :n0].cpu().numpy(), 7.0, resize_scale=0.5)
        if M is None:
            assert r["Matrix"] is None
            continue
        n_found += 1
        assert np.array_equal(r["mask"], mask) and np.abs(r["Matrix"] - M).max() < 1e-9
        assert np.array_equal(r["mkpts0"], mk0) and np.array_equal(r["mkpts1"], mk1)
        src = (a[i, 0].double() * 255)
        w = warp_affine(m, src, reg["matrices"][i]).cpu().numpy()
        assert np.array_equal(w, R.warp_affine(src.cpu().numpy(), M))
    assert n_found >= 1
