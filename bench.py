#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the SuperPoint+SuperGlue Matching.forward hot path.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1         # CPU arm (oracle port, host cores)
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU, pairs sharded

Workload (BASELINE.json configs[1], "C2"): per GPU a batch of 64 synthetic 640x480 grayscale pairs,
max_keypoints 1024, descriptor_dim 128, keypoint_encoder [32,64,128], 18 GNN layers, 30 Sinkhorn
iterations.  A step is one Matching.forward over that batch.  Weak scaling: every rank owns its own
64 pairs; the only collective is the final all-gather of match indices / scores.
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_RESULT_LINE = []              # the one JSON line, printed by main() after stdout is restored

H, W, MAX_KP, D = 480, 640, 1024, 128
KENC = [32, 64, 128]
SINKHORN = 30
GF_PAIR_TOTAL = 135.43        # algorithmic GFLOP per pair (C1/C2/C4)
GF_PAIR_QK = 9.664            # attention QK^T only
# dram__bytes_read+write of the dominant conv launch (fused stem + 64->64 layer at 480x640, 64-image micro-batch) from
# profiles/r01_ncu_step_per_kernel.txt / the ncu run behind it: 78.9 MB read (= the 64 images, compulsory) + 1204 MB
# written; algorithmic = 78.6 MB images in + 1258 MB of pooled fp16 hi/lo planes out (a little of the output is still in
# L2 when the kernel ends): no wasted re-reads
NCU_CONV_DRAM_BYTES_PER_LAUNCH = 1283.0e6   # per 64-image launch (20.0 MB / image; scaled by the images one launch covers)
NCU_CONV_IMAGES_PER_LAUNCH = 64
GF_IMG_CONV3 = 51.79 - 0.354 - 0.472   # the eight 3x3 conv layers with Cin >= 64 (all but the Cin=1 stem and the two 1x1 heads)
GF_IMG_CONV1 = 0.354                   # the Cin=1 stem, computed inside the fused first tc_conv launch
GF_IMG_C1B = 2 * 9 * 64 * 64 * H * W / 1e9   # the 64->64 3x3 conv at full resolution (22.65 GF / image)


def make_cfg():
    return {"superpoint": {"descriptor_dim": D, "nms_radius": 4, "keypoint_threshold": 0.005,
                           "max_keypoints": MAX_KP, "remove_borders": 4},
            "superglue": {"descriptor_dim": D, "keypoint_encoder": list(KENC),
                          "GNN_layers": ["self", "cross"] * 9, "sinkhorn_iterations": SINKHORN,
                          "match_threshold": 0.2}}


def load_weights():
    from image_matching_b200 import synth
    real = os.path.join(ROOT, "tests", "golden", "superpoint_allss128_weights.npz")
    if os.path.exists(real):
        sp, sp_name = dict(np.load(real)), "reference SuperPoint checkpoint (allss, D=128)"
    else:
        sp, sp_name = synth.superpoint_weights(0, D), "seeded synthetic SuperPoint weights"
    return sp, synth.superglue_weights(0, D, KENC), sp_name


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, n): n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                 for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and nm not in ("All", "None", "GpuIdle"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def cpu_oracle():
    """The CPU arm's implementation: the torch-CPU restatement of the reference path (the reference's own arithmetic
    library, oneDNN convolutions, all host threads); the numpy oracle stays the parity checker."""
    import torch
    from oracle import matching_oracle_torch as O
    torch.set_num_threads(os.cpu_count() or 1)
    return O, torch.get_num_threads()


def cpu_oracle_pairs_per_s(n_pairs, seeds_from=1000):
    """The oracle port on the host cores (all threads); returns (pairs/s, seconds)."""
    from image_matching_b200 import synth
    O, _ = cpu_oracle()
    sp, sg, _ = load_weights()
    cfg = make_cfg()
    a, b = synth.make_pair(seeds_from - 1, H, W)
    O.matching_forward(a, b, sp, sg, cfg)          # warm-up (BLAS thread pool, page faults)
    t0 = time.perf_counter()
    for i in range(n_pairs):
        a, b = synth.make_pair(seeds_from + i, H, W)
        t_in = time.perf_counter()
        O.matching_forward(a, b, sp, sg, cfg)
        if i == 0:
            gen = t_in - t0
    dt = time.perf_counter() - t0 - gen * n_pairs
    return n_pairs / dt, dt


def blas_threads():
    import torch
    return torch.get_num_threads()


def run_reference(args):
    """CPU arm: the reference's algorithm (torch-CPU oracle port; the reference itself is a Python script collection
    that cannot travel to the GPU box) on all host cores.  Each step = 1 pair of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from image_matching_b200 import synth
    O, _ = cpu_oracle()
    sp, sg, sp_name = load_weights()
    cfg = make_cfg()
    pairs = [synth.make_pair(2000 + i, H, W) for i in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        O.matching_forward(pairs[i][0], pairs[i][1], sp, sg, cfg)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        O.matching_forward(pairs[i][0], pairs[i][1], sp, sg, cfg)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cores = blas_threads()
    line = {"impl": "reference", "metric": "image-pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(1, sp_name),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x 1 pair (640x480, 1024 kpts) through the torch-CPU oracle port "
                                       "(oracle/matching_oracle_torch.py)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _RESULT_LINE.append(json.dumps(line))


def workload_config(batch_per_gpu, sp_name):
    return {"workload": f"C2: batch={batch_per_gpu} pairs/GPU of synthetic 640x480 grayscale (random rectangles + "
                        "perspective warp), max_keypoints=1024, desc_dim=128, kenc [32,64,128], 18 GNN layers, "
                        "30 Sinkhorn iterations",
            "pairs_per_gpu": batch_per_gpu, "image": [H, W], "max_keypoints": MAX_KP, "descriptor_dim": D,
            "gnn_layers": 18, "sinkhorn_iterations": SINKHORN, "weights": sp_name + " + seeded synthetic SuperGlue",
            "l2": "inputs (157 MB/step at 64 pairs) and the ~2 GB activation arena exceed the 126 MB L2; no flush needed"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from image_matching_b200 import Matching, synth
    from image_matching_b200.dist import gather_matches

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    sp, sg, sp_name = load_weights()
    cfg = make_cfg()
    c = {"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}
    m = Matching(c).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    m = m.to(dev)
    torch.set_grad_enabled(False)

    # synthetic pairs: a pool of distinct seeds, tiled to the batch (generation is host-side numpy)
    n_unique = min(B, args.unique)
    a, b = synth.make_pair_batch([rank * 100000 + i for i in range(n_unique)], H, W)
    reps = (B + n_unique - 1) // n_unique
    a = np.concatenate([a] * reps)[:B]
    b = np.concatenate([b] * reps)[:B]
    h0 = torch.from_numpy(a).pin_memory()
    h1 = torch.from_numpy(b).pin_memory()
    d0, d1 = h0.to(dev), h1.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        out = m.forward_device(d0, d1)
        if world > 1:
            gather_matches(out["matches0"], out["matching_scores0"], B * world)
        return out

    # End to end through the public API with HOST buffers.  Every step copies its own inputs from pinned host memory
    # and reads its results back; the copy of step i+1's images runs on a side stream while step i computes (what any
    # input pipeline does), so the PCIe transfer is inside the timed region but off the critical path.
    copy_stream = torch.cuda.Stream(device=dev)

    src = {"a": h0, "b": h1}

    def upload():
        with torch.cuda.stream(copy_stream):
            x0 = src["a"].to(dev, non_blocking=True)
            x1 = src["b"].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x0, x1, ev

    pending = []

    def step_e2e():
        if not pending:
            pending.append(upload())
        x0, x1, ev = pending.pop()
        pending.append(upload())                       # next step's inputs start moving now
        torch.cuda.current_stream().wait_event(ev)
        x0.record_stream(torch.cuda.current_stream())
        x1.record_stream(torch.cuda.current_stream())
        pred = m({"image0": x0, "image1": x1})
        m0, s0 = pred["matches0"], pred["matching_scores0"]
        if world > 1:
            m0, s0 = gather_matches(m0, s0, B * world)
        # what the reference's caller reads back per pair (superpoint_glue_test.py:79-82)
        res = [m0.cpu(), s0.cpu(), torch.stack(pred["keypoints0"]).cpu(), torch.stack(pred["keypoints1"]).cpu()]
        return res

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    counts = out["counts"].cpu().numpy()
    valid = int((out["matches0"] > -1).sum().item())

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = m._engine.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = m._engine.launch_count() - launches0
    # ---- end-to-end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        res = step_e2e()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    # informational: the same loop uploading the 8-bit pixels (b200m_matching_forward_u8 normalises on the device;
    # the reference's loader would upload 4x / 8x more bytes as float32 / float64)
    src["a"] = torch.from_numpy(np.round(a * 255).astype(np.uint8)).pin_memory()
    src["b"] = torch.from_numpy(np.round(b * 255).astype(np.uint8)).pin_memory()
    pending.clear()
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_u8_ms = 1e3 * (time.perf_counter() - t0)
    # informational: the caller's next step on the device -- batched estimateAffinePartial2D(RANSAC)
    # straight from the device-resident matches, then the matrices / inlier masks read back
    from image_matching_b200 import estimate_affine_partial_2d

    def step_registered():
        out = m.forward_device(d0, d1)
        mats, inl, info = estimate_affine_partial_2d(m, out["keypoints0"], out["keypoints1"], out["matches0"],
                                                     out["counts"][0], 7.0)
        return mats.cpu(), inl.cpu(), info.cpu()

    step_registered()
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(args.steps):
        reg = step_registered()
    r1.record()
    barrier()
    reg_ms = r0.elapsed_time(r1)
    # informational: single-pair latency (the reference script's own batch_size=1 loop, superpoint_glue_test.py:66,72)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        m.forward_device(d0[:1], d1[:1])
    barrier()
    l0.record()
    for _ in range(20):
        m.forward_device(d0[:1], d1[:1])
    l1.record()
    barrier()
    lat_ms = l0.elapsed_time(l1) / 20
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([ms, e2e_ms, e2e_u8_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_u8_ms = float(t[0]), float(t[1]), float(t[2])

    # ---- per-kernel profile pass (CUDA events around every launch, on the launching stream)
    prof = {}
    if rank == 0:
        import ctypes as C
        from image_matching_b200 import lib
        L = lib.load()
        lib.check(L.b200m_profile_begin(m._engine.handle, 20000))
        nprof = min(args.steps, 2)
        for _ in range(nprof):
            m.forward_device(d0, d1)
        buf = C.create_string_buffer(1 << 16)
        lib.check(L.b200m_profile_end(m._engine.handle, buf, len(buf)))
        prof = json.loads(buf.value.decode())
        for k in prof:
            prof[k]["ms_per_step"] = prof[k]["ms"] / nprof
            prof[k]["launches_per_step"] = prof[k]["launches"] / nprof
    if world > 1:
        dist.barrier()

    if rank == 0:
        pk = peaks()
        total_pairs = B * world
        value = total_pairs * args.steps / (ms / 1e3)
        e2e = total_pairs * args.steps / (e2e_ms / 1e3)
        dom = max(prof, key=lambda k: prof[k]["ms_per_step"]) if prof else None
        roof = None
        # roofline of the dominant kernel: the fused first launch of the encoder (stem conv 1->64 computed in the operand
        # producer + the 64->64 3x3 conv at full resolution + 2x2 max-pool), one launch = one 16-image micro-batch
        ck = "tc_conv3x3_stem" if "tc_conv3x3_stem" in prof else ("tc_conv3x3" if "tc_conv3x3" in prof else None)
        if ck:
            conv_ms = prof[ck]["ms_per_step"]
            n_launch = prof[ck]["launches_per_step"]
            gf_img = (GF_IMG_C1B + GF_IMG_CONV1) if ck == "tc_conv3x3_stem" else GF_IMG_CONV3
            flops_step = gf_img * 1e9 * 2 * B
            ach = flops_step / (conv_ms / 1e3) / 1e12                  # algorithmic FLOPs (NOT x3 for the fp16 split)
            all_conv_ms = sum(prof[k]["ms_per_step"] for k in ("tc_conv3x3", "tc_conv3x3_stem", "tc_conv1x1") if k in prof)
            roof = {"kernel": ck + " (tc_conv.cu: fused stem + 64->64 3x3 conv @ 480x640 + max-pool; implicit GEMM on "
                              "tcgen05, fp16 hi/lo operand split, weights resident in shared memory)",
                    "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"],
                    "traffic": NCU_CONV_DRAM_BYTES_PER_LAUNCH * (2 * B / max(n_launch, 1)) / NCU_CONV_IMAGES_PER_LAUNCH,
                    "peak_source": pk["source"] + " bf16 dense sustained (kernel timed inside a long step)",
                    "flops_per_launch": flops_step / max(n_launch, 1), "avg_launch_ms": conv_ms / max(n_launch, 1),
                    "algorithmic_bytes_per_launch": int(2 * B / max(n_launch, 1) * (H * W * 4 + 64 * (H // 2) * (W // 2) * 2 * 2)),
                    "images_per_launch": 2 * B / max(n_launch, 1),
                    "share_of_step": conv_ms / sum(p["ms_per_step"] for p in prof.values()),
                    "all_conv_kernels_share_of_step": all_conv_ms / sum(p["ms_per_step"] for p in prof.values()),
                    "all_conv_kernels_tflops": (GF_IMG_CONV3 + GF_IMG_CONV1 + 0.472) * 2 * B / all_conv_ms,
                    "dominant_by_time": dom,
                    "note": "achieved counts algorithmic FLOPs once; the kernel issues 3 fp16 products per algorithmic "
                            "product (A_hi W_hi + A_hi W_lo + A_lo W_hi, fp32-class accuracy: 0 keypoint flips vs the "
                            "reference), i.e. %.0f TFLOP/s of fp16 tensor work against the bf16/fp16 dense peak; "
                            "traffic = dram bytes per launch from ncu (profiles/r01_ncu_step_per_kernel.txt)" % (3 * ach)}
        line = {"metric": "image-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(B, sp_name),
                "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * B * H * W * 4),
                        "d2h_bytes_per_step": int(sum(r.numel() * r.element_size() for r in res) + 8 * B),
                        "ms_per_step": e2e_ms / args.steps,
                        "note": "Matching.forward on fp32 host images; the upload of step i+1 overlaps step i on a side stream"},
                "e2e_uint8_images": {"value": total_pairs * args.steps / (e2e_u8_ms / 1e3), "unit": "pairs/s",
                                     "h2d_bytes_per_step": int(2 * B * H * W), "ms_per_step": e2e_u8_ms / args.steps},
                "with_registration": {"value": B * args.steps / (reg_ms / 1e3), "unit": "pairs/s (this rank)",
                                      "ms_per_step": reg_ms / args.steps,
                                      "inliers_per_pair": float(reg[2][:, 1].float().mean()),
                                      "ransac_iterations_per_pair": float(reg[2][:, 2].float().mean()),
                                      "note": "forward_device + b200m_estimate_affine_partial (cv2-identical RANSAC, "
                                              "7 px) + D2H of matrices and inlier masks"},
                "latency_batch1_ms": lat_ms,
                "gpu_launches": int(launches),
                "clocks": sampler.result(),
                "roofline": roof,
                "qk_roofline": {"gflop_per_pair": GF_PAIR_QK, "achieved_tflops": value / world * GF_PAIR_QK / 1e3,
                                "frac_of_bf16_peak": value / world * GF_PAIR_QK / 1e3 / pk["tf_sustained"]},
                "whole_path_tflops_per_gpu": value / world * GF_PAIR_TOTAL / 1e3,
                "kernel_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items())},
                "check": {"keypoints_per_image_min": int(counts.min()), "keypoints_per_image_max": int(counts.max()),
                          "valid_matches_per_pair": valid / B}}
        if world == 1 and not args.no_cpu:
            v, dt = cpu_oracle_pairs_per_s(args.cpu_pairs)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": blas_threads(), "kind": "port",
                                    "sample": f"{args.cpu_pairs} pairs of the same workload through the torch-CPU "
                                              f"oracle port ({dt:.1f} s)"}
        _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU per step (C2 = 64)")
    ap.add_argument("--unique", type=int, default=16, help="distinct synthetic pairs generated per rank")
    ap.add_argument("--cpu-pairs", type=int, default=16, help="pairs timed through the CPU oracle (cpu_baseline)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr, so whatever a
    # library prints to stdout from C code (NCCL's version banner under NCCL_DEBUG=VERSION, for one) cannot precede it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
