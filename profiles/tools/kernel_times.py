"""Per-kernel CUDA-event times of one Matching.forward_device step (device-resident inputs).
usage: python profiles/tools/kernel_times.py [pairs]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from image_matching_b200 import Matching, synth, lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sp, sg, _ = bench.load_weights(bench.CONFIGS["C2"])
cfg = bench.make_cfg(bench.CONFIGS["C2"])
m = Matching({"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}).eval()
m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
m = m.to("cuda:0")
a, b = synth.make_pair_batch(range(min(B, 8)), 480, 640)
a = np.concatenate([a] * ((B + 7) // 8))[:B]
b = np.concatenate([b] * ((B + 7) // 8))[:B]
d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
for _ in range(3):
    m.forward_device(d0, d1)
torch.cuda.synchronize()
L = lib.load()
lib.check(L.b200m_profile_begin(m._engine.handle, 20000))
n = 3
for _ in range(n):
    m.forward_device(d0, d1)
buf = C.create_string_buffer(1 << 16)
lib.check(L.b200m_profile_end(m._engine.handle, buf, len(buf)))
prof = json.loads(buf.value.decode())
tot = 0.0
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:22s} {v['ms'] / n:9.3f} ms/step  {v['launches'] // n:5d} launches")
    tot += v["ms"] / n
print(f"{'sum':22s} {tot:9.3f} ms/step for {B} pairs -> {B / tot * 1e3:.1f} pairs/s (kernel time only)")
