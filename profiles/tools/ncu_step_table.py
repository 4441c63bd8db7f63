#!/usr/bin/env python
"""Per-kernel table of EXACTLY ONE bench step from an ncu CSV, plus the DRAM traffic per kernel that bench.py's
`roofline.traffic` reads (profiles/r02_ncu_traffic.json).

Capture (one GPU; `--ncu-step` brackets one step with cudaProfilerStart/Stop, so nothing else is profiled):

    ncu --profile-from-start off --clock-control none --csv --page raw \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active \
        --log-file gpurun_out/ncu_step_C2.csv python bench.py --config C2 --ncu-step

    python profiles/tools/ncu_step_table.py gpurun_out/ncu_step_C2.csv C2 64 > profiles/r02_ncu_step_per_kernel_C2.txt

Times under ncu are serialised and cold-cache: compare SHARES with bench.py's CUDA-event times, not absolutes.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def prof_name(kernel):
    """CUDA kernel name -> the name of the library's per-launch profiler scope (the keys of bench.py's `kernels`)."""
    k = kernel
    m = re.match(r"(?:void )?(?:b200m::)?tc_conv_kernel<(\d+), *(\w+), *(\d+), *(\w+), *(\w+)>", k)
    if m:
        ks, fuse = m.group(3), m.group(4)
        return "tc_conv1x1" if ks == "1" else ("tc_conv3x3_stem" if fuse in ("1", "true") else "tc_conv3x3")
    table = [("tc_attention", "tc_attention"), ("tc_gnn_layer", "tc_gnn_layer"), ("tc_gemm", "tc_gemm"),
             ("ot_iter", "ot_iter_fused"), ("ot_init", "ot_init"), ("nms_", "nms_candidates"),
             ("softmax_heat", "softmax_heat"), ("select_keypoints", "select_keypoints"),
             ("sample_desc", "sample_descriptors"), ("c4_l2_normalize", "c4_l2_normalize"),
             ("argmax", "argmax"), ("match_select", "match_select"), ("kenc_input", "kenc_input"),
             ("apply_flags", "apply_flags"), ("gemm_tn", "gemm"), ("u8_to_unit", "u8_to_unit_f32"),
             ("match_wire", "match_wire"), ("resize", "resize")]
    for sub, name in table:
        if sub in k:
            return name
    return None


def main():
    path, cname, pairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}

    def f(r, name):
        try:
            return float(r[col[name]].replace(",", ""))
        except (KeyError, ValueError):
            return 0.0
    unit_t = rows[1][col["gpu__time_duration.sum"]]
    t_scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit_t, 1.0)

    def bytes_of(r, name):
        u = rows[1][col[name]]
        return f(r, name) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    agg = collections.OrderedDict()
    for r in data:
        k = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("b200m::", "").replace("void ", "")
        a = agg.setdefault(k, collections.Counter())
        t = f(r, "gpu__time_duration.sum") * t_scale
        a["n"] += 1
        a["t"] += t
        a["rd"] += bytes_of(r, "dram__bytes_read.sum")
        a["wr"] += bytes_of(r, "dram__bytes_write.sum")
        for key, name in (("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                          ("sm", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                          ("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                          ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active")):
            a[key] += f(r, name) * t          # time-weighted
    tot = sum(a["t"] for a in agg.values())
    print(f"One bench step ({pairs} pairs, {cname}) under ncu: {len(data)} launches, per kernel, summed over its launches "
          f"in the step (cudaProfilerStart/Stop around one step: `bench.py --config {cname} --ncu-step`).")
    print("Times under ncu are serialised and cold-cache: compare shares.  HBM GB/s = (dram read + write) / time; "
          "percentages are time-weighted means over the launches.")
    print(f"{'kernel':44s} {'launches':>8s} {'time us':>10s} {'share %':>8s} {'HBM GB/s':>9s} {'dram %':>7s} "
          f"{'tensor %':>8s} {'sm thr %':>8s} {'issue %':>8s} {'warps %':>8s} {'MB / launch':>12s}")
    traffic = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        t = max(a["t"], 1e-9)
        print(f"{k[:44]:44s} {a['n']:8d} {a['t']:10.1f} {100 * a['t'] / tot:8.2f} {(a['rd'] + a['wr']) / t / 1e3:9.0f} "
              f"{a['dram'] / t:7.1f} {a['tensor'] / t:8.1f} {a['sm'] / t:8.1f} {a['issue'] / t:8.1f} {a['warps'] / t:8.1f} "
              f"{(a['rd'] + a['wr']) / a['n'] / 1e6:12.2f}")
        pn = prof_name(k)
        if pn:
            e = traffic.setdefault(pn, {"dram_bytes_per_step": 0.0, "launches_per_step": 0, "pairs_per_step": pairs,
                                        "ncu_time_us": 0.0})
            e["dram_bytes_per_step"] += a["rd"] + a["wr"]
            e["launches_per_step"] += a["n"]
            e["ncu_time_us"] += a["t"]
    print(f"{'total':44s} {len(data):8d} {tot:10.1f}")
    out = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    allt = json.load(open(out)) if os.path.exists(out) else {}
    allt[cname] = traffic
    json.dump(allt, open(out, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
