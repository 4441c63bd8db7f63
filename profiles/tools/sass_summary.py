#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-only SASS instructions in the built library (evidence that the hot kernels are
hand-written tcgen05 / TMEM / TMA code, not recompiled mma.sync):

    UTCHMMA  tcgen05.mma (kind::f16)        UTCQMMA/UTCIMMA other tcgen05.mma kinds
    UTMALDG  cp.async.bulk.tensor load (TMA)        UTMASTG  TMA tensor store        UBLKCP  cp.async.bulk (1-D bulk copy)
    LDTM / STTM  tcgen05.ld / tcgen05.st (TMEM)     SYNCS    mbarrier ops            UTCBAR  tcgen05.commit
    HMMA/IMMA/QMMA  legacy mma.sync (must be absent in the tensor-core kernels)

usage: python profiles/tools/sass_summary.py [path/to/libb200match.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "image_matching_b200", "libb200match.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "MUFU.EX2",
       "HMMA", "IMMA", "QMMA"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, regs, cur = collections.OrderedDict(), {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("b200m::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                counts[cur][o] += 1
print(f"SASS instruction counts per kernel of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_total"]):
    print(f"{k[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
