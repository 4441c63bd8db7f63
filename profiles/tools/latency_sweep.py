"""Latency of Matching.forward_device (device-resident inputs, CUDA-graph replay) for small batches.
usage: python profiles/tools/latency_sweep.py [pairs ...]   (default 1 2 4 8 16)
Environment switches of the library apply (B200M_PDL, B200M_PDL_MAX_PAIRS, B200M_SP_DUAL_MAX, B200M_GRAPHS)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from image_matching_b200 import Matching, synth  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]
c = bench.CONFIGS["C2"]
sp, sg, _ = bench.load_weights(c)
cfg = bench.make_cfg(c)
m = Matching({"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}).eval()
m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
m = m.to("cuda:0")
a, b = synth.make_pair_batch(range(8), 480, 640)
out = []
for B in sizes:
    aa = np.concatenate([a] * ((B + 7) // 8))[:B]
    bb = np.concatenate([b] * ((B + 7) // 8))[:B]
    d0, d1 = torch.from_numpy(aa).cuda(), torch.from_numpy(bb).cuda()
    o = None
    for _ in range(4):
        o = m.forward_device(d0, d1, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n):
        o = m.forward_device(d0, d1, out=o)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"B={B}: {e0.elapsed_time(e1) / n:.3f} ms")
print("  ".join(out), f"(valid matches of the last batch: {int((o['matches0'] > -1).sum())}; graph replays {m._engine.graph_replays()})")
