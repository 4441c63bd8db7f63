#!/usr/bin/env python
"""Second baseline of SURVEY.md 8(d): the reference's algorithm in EAGER torch on the B200 (cuDNN convolutions, cuBLAS
batched GEMMs, ATen kernels) -- what `superpoint_glue_test.py` gets today with `device='cuda'`, one pair per call like
its DataLoader(batch_size=1) loop.  Measurement tooling only: it drives the torch restatement of the reference path
(oracle/matching_oracle_torch.py, validated against the reference-generated goldens) with CUDA tensors; nothing in the
product path uses it.  Prints one JSON line.

    python profiles/tools/eager_gpu_baseline.py [--pairs 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20)
    ap.add_argument("--tf32", type=int, default=1, help="cudnn.allow_tf32 (torch default: 1); matmul TF32 stays off")
    args = ap.parse_args()
    import bench
    from image_matching_b200 import synth
    from oracle import matching_oracle_torch as O

    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    cache = {}

    def to_dev(x):      # weights are uploaded once (the reference's modules keep them on the device)
        k = id(x)
        if k not in cache:
            cache[k] = (x, torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(dev))
        return cache[k][1]

    O._t = to_dev
    to_host = torch.Tensor.numpy
    torch.Tensor.numpy = lambda self, *a, **k: to_host(self.detach().cpu(), *a, **k)   # the caller's .cpu().numpy()
    torch.set_default_device(dev)
    torch.set_grad_enabled(False)
    sp, sg, sp_name = bench.load_weights()
    cfg = bench.make_cfg()
    pairs = [synth.make_pair(3000 + i, bench.H, bench.W) for i in range(args.pairs + 3)]
    for a, b in pairs[:3]:
        out = O.matching_forward(a, b, sp, sg, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, b in pairs[3:]:
        out = O.matching_forward(a, b, sp, sg, cfg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"baseline": "reference algorithm, eager torch %s on %s (cuDNN/cuBLAS/ATen), batch 1 per call"
                                  % (torch.__version__, torch.cuda.get_device_name(0)),
                      "pairs_per_s": args.pairs / dt, "ms_per_pair": 1e3 * dt / args.pairs, "pairs": args.pairs,
                      "cudnn_allow_tf32": bool(args.tf32), "matmul_allow_tf32": False,
                      "includes": "H2D of the two fp32 images and D2H of keypoints / matches per pair, as the script does",
                      "keypoints": [int(len(out["keypoints0"])), int(len(out["keypoints1"]))],
                      "valid_matches": int((out["matches0"] > -1).sum()), "weights": sp_name}))


if __name__ == "__main__":
    main()
