"""Summarise one `ncu --set full --import-source on` report: headline metrics, instruction mix and the warp-stall
samples aggregated per CUDA source line (SASS offsets mapped through `nvdisasm --print-line-info` of the in-tree .so).

usage: python profiles/tools/ncu_lines.py report.ncu-rep kernel_name_substring [top_n]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_xu.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def line_map(kernel):
    tmp = tempfile.mkdtemp()
    so = os.path.join(ROOT, "image_matching_b200", "libb200match.so")
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
        if kernel not in txt:
            continue
        off2line, line, infunc = {}, None, False
        for l in txt.splitlines():
            if l.startswith(".text."):
                infunc = kernel in l
            elif l.startswith(".section"):
                infunc = False
            if not infunc:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
            if m:
                off2line[int(m.group(1), 16)] = line
        if off2line:
            return off2line
    return {}


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {vals[i]:>16s} {units[i]}")
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.15:
            print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(vals[i]):.2f}")
    src = ncu_csv(rep, "source")
    h = src[1]
    iS, iN, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = []
    for r in src[2:]:                     # a report with several launches repeats the table: keep the first launch
        if len(r) <= max(iS, iN, iW) or not re.fullmatch(r"(0x)?[0-9a-fA-F]+", r[0]):
            break
        data.append(r)
    tot_n = sum(int(r[iN]) for r in data)
    tot_w = sum(int(r[iW]) for r in data)
    ops, opw = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
        op = m.group(2) if m else "?"
        ops[op] += int(r[iN])
        opw[op] += int(r[iW])
    print(f"instructions {tot_n}, stall samples {tot_w}; mix:")
    print("  " + ", ".join(f"{op} {100 * n / tot_n:.1f}%" for op, n in ops.most_common(14)))
    o2l = line_map(kernel)
    base = int(data[0][0], 16)
    agg, aggn = collections.Counter(), collections.Counter()
    for r in data:
        ln = o2l.get(int(r[0], 16) - base)
        agg[ln] += int(r[iW])
        aggn[ln] += int(r[iN])
    cache = {}
    for ln, w in agg.most_common(top):
        text = ""
        if ln:
            f = os.path.join(ROOT, "image_matching_b200", "csrc", ln[0])
            if os.path.exists(f):
                cache.setdefault(f, open(f).read().splitlines())
                text = cache[f][ln[1] - 1].strip()[:95]
        print(f"{w:6d} {100 * w / max(tot_w, 1):5.1f}%  inst {100 * aggn[ln] / max(tot_n, 1):5.1f}%  {ln}: {text}")


if __name__ == "__main__":
    main()
