#!/usr/bin/env python
"""Timing of the other BASELINE.json configurations on one B200 (informational; bench.py's line is C2):
  C3: 1280x960 pairs, 2048 keypoints, descriptor_dim 256, keypoint_encoder [32,64,128,256], 30 Sinkhorn iterations
  C5: SuperGlue only on external 128-d descriptors, 4096 keypoints per image, 100 Sinkhorn iterations
with the per-kernel CUDA-event breakdown of b200m_profile_begin/end.  Seeded synthetic weights and inputs.

    python profiles/tools/config_times.py [--c3-pairs 8] [--c5-pairs 4]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from image_matching_b200 import Matching, lib, synth  # noqa: E402

DEV = "cuda:0"


def build(D, kenc, max_kp, iters):
    cfg = {"superpoint": {"descriptor_dim": D, "nms_radius": 4, "keypoint_threshold": 0.005, "max_keypoints": max_kp,
                          "remove_borders": 4, "weights": None},
           "superglue": {"descriptor_dim": D, "keypoint_encoder": list(kenc), "GNN_layers": ["self", "cross"] * 9,
                         "sinkhorn_iterations": iters, "match_threshold": 0.2, "weights": ""}}
    m = Matching(cfg).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_weights(1, D).items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superglue_weights(1, D, kenc).items()})
    return m.to(DEV)


def timed(m, fn, steps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    L = lib.load()
    lib.check(L.b200m_profile_begin(m._engine.handle, 20000))
    fn()
    buf = C.create_string_buffer(1 << 16)
    lib.check(L.b200m_profile_end(m._engine.handle, buf, len(buf)))
    prof = {k: round(v["ms"], 3) for k, v in sorted(json.loads(buf.value.decode()).items(), key=lambda kv: -kv[1]["ms"])}
    return ms, prof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-pairs", type=int, default=8)
    ap.add_argument("--c5-pairs", type=int, default=4)
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    out = {}
    if args.c3_pairs:
        B = args.c3_pairs
        m = build(256, (32, 64, 128, 256), 2048, 30)
        a, b = synth.make_pair_batch(list(range(40, 40 + min(B, 4))), 960, 1280)
        reps = (B + len(a) - 1) // len(a)
        a, b = np.concatenate([a] * reps)[:B], np.concatenate([b] * reps)[:B]
        d0, d1 = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)
        ms, prof = timed(m, lambda: m.forward_device(d0, d1))
        cnt = m.forward_device(d0, d1)["counts"].cpu().numpy()
        out["C3"] = {"pairs_per_step": B, "ms_per_step": ms, "pairs_per_s": B / ms * 1e3,
                     "keypoints_min": int(cnt.min()), "kernel_ms": prof}
        del m
    if args.c5_pairs:
        B, N, H, W = args.c5_pairs, 4096, 480, 640
        m = build(128, (32, 64, 128), -1, 100)
        kp0, sc0, de0 = synth.random_features(1, B, N, 128, H, W)
        kp1, sc1, de1 = synth.random_features(2, B, N, 128, H, W)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)  # noqa: E731
        data = {"image0": torch.empty(B, 1, H, W, device=DEV), "image1": torch.empty(B, 1, H, W, device=DEV),
                "keypoints0": t(kp0), "scores0": t(sc0), "descriptors0": t(de0),
                "keypoints1": t(kp1), "scores1": t(sc1), "descriptors1": t(de1)}
        ms, prof = timed(m, lambda: m(data))
        out["C5"] = {"pairs_per_step": B, "ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "keypoints": N,
                     "sinkhorn_iterations": 100, "kernel_ms": prof}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
