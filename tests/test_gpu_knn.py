"""GPU: exact 2-NN + ratio-test descriptor matcher (replaces the FLANN step of
superpoint_flann_test.py:69-78) against a float64 brute-force restatement."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg

pytestmark = pytest.mark.gpu


def _ref(d0, d1, ratio):
    a, b = d0.T.astype(np.float64), d1.T.astype(np.float64)              # (N,D), (M,D) like Desc1 / Desc2 (:66-67)
    dist = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
    order = np.argsort(dist, axis=1, kind="stable")
    e1, e2 = dist[np.arange(len(a)), order[:, 0]], dist[np.arange(len(a)), order[:, 1]]
    return np.where(e1 < ratio * e2, order[:, 0], -1), e1, e2


@pytest.mark.parametrize("D,N,M", [(128, 1024, 1024), (128, 300, 77), (256, 513, 640), (64, 65, 1000)])
def test_knn_ratio_match(D, N, M):
    from image_matching_b200 import SuperPoint, knn_ratio_match, synth
    cfg = golden_cfg(D=D)
    sp = SuperPoint(dict(cfg["superpoint"], weights=None)).eval()
    sp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_weights(0, D).items()})
    sp = sp.to("cuda:0")
    rng = np.random.default_rng(5)
    d1 = rng.standard_normal((D, M)).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=0, keepdims=True)
    # queries: noisy copies of some train descriptors (true matches) + unrelated ones
    src = rng.integers(0, M, N)
    d0 = d1[:, src] + 0.15 * rng.standard_normal((D, N)).astype(np.float32) * (rng.random(N) < 0.6)
    d0[:, N // 2:] = rng.standard_normal((D, N - N // 2)).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=0, keepdims=True)
    m, e1, e2 = knn_ratio_match(sp, torch.from_numpy(d0).cuda(), torch.from_numpy(d1).cuda(), 0.7)
    rm, r1, r2 = _ref(d0, d1, 0.7)
    m, e1, e2 = m.cpu().numpy(), e1.cpu().numpy(), e2.cpu().numpy()
    assert np.abs(e1 - r1).max() < 2e-5 and np.abs(e2 - r2).max() < 2e-5   # sqrt near 0 amplifies fp32 rounding
    decided = np.abs(r1 - 0.7 * r2) > 1e-4                                  # away from the ratio boundary
    assert np.array_equal(m[decided], rm[decided])
    assert (rm >= 0).sum() > N // 8
