"""compute-sanitizer --tool memcheck python profiles/tools/sanitize_smoke.py: smoke() plus small invocations of the newer\nkernels (wide Sinkhorn, radix-select top-k, fused NMS on 64x64 tiles, RANSAC, warpAffine).  Result: profiles/r01_compute_sanitizer_memcheck.log"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import __graft_entry__ as g
if os.environ.get('SANITIZE_SKIP_SMOKE') != '1':
    g.smoke()
from image_matching_b200 import Matching, stages, synth, estimate_affine_partial_2d, warp_affine
cfg = {"superpoint": {"descriptor_dim": 128, "nms_radius": 4, "keypoint_threshold": 0.005, "max_keypoints": 1024, "remove_borders": 4, "weights": None},
       "superglue": {"descriptor_dim": 128, "keypoint_encoder": [32, 64, 128], "GNN_layers": ["self", "cross"] * 9, "sinkhorn_iterations": 3, "match_threshold": 0.2, "weights": ""}}
m = Matching(cfg).eval()
m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_weights(0, 128).items()})
m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superglue_weights(0, 128).items()})
m = m.to("cuda:0")
rng = np.random.default_rng(0)
for N, M in ((300, 1100), (200, 2100), (130, 4096), (50, 3000)):
    Z = stages.sinkhorn(m, torch.from_numpy(rng.standard_normal((2, N, M)).astype(np.float32)).cuda(), iters=2)
    assert torch.isfinite(Z).all()
semi = torch.from_numpy(rng.standard_normal((1, 65, 120, 160)).astype(np.float32)).cuda()
heat, nms, kp, sc, cnt = stages.detector_post(m, semi)
assert int(cnt[0]) == 1024
k0 = torch.rand(3, 500, 2, device="cuda") * 400
k1 = k0 + 5
m0 = torch.arange(500, device="cuda")[None].repeat(3, 1)
m0[:, ::3] = -1
mats, inl, info = estimate_affine_partial_2d(m, k0, k1, m0, None, 7.0)
w = warp_affine(m, torch.rand(3, 100, 130, device="cuda", dtype=torch.float64), mats)
torch.cuda.synchronize()
print("sanitize script ok", info.cpu().tolist())
