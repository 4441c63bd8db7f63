#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the SuperPoint+SuperGlue Matching.forward hot path.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo's CUDA path, config C2
    python bench.py --config C3|C4|C5 ...                         # the other BASELINE.json configurations
    python bench.py --impl reference --steps 3 --warmup 1         # CPU arm: the reference's own modules (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU, pairs sharded (default config C4)

Workloads (BASELINE.json `configs`; every one is synthetic data of the stated shape, see `config.workload`):
  C2  configs[1]  64 pairs/GPU of 640x480, 1024 keypoints, D=128, kenc [32,64,128], 18 GNN layers, 30 Sinkhorn iterations
  C3  configs[2]  256 pairs/GPU of 1280x960, 2048 keypoints, D=256, kenc [32,64,128,256] (reference COCO-256 SuperPoint weights)
  C4  configs[3]  C2's model at 128 pairs/GPU, weak scaling over 1/2/4/8 GPUs (the default when --gpus N > 1)
  C5  configs[4]  SuperGlue only on supplied 128-d descriptors, 4096 keypoints per image, 100 Sinkhorn iterations, 8 pairs/GPU
A step is one Matching.forward over that batch.  Weak scaling: every rank owns its own pairs; the only collective is
the final gather of match indices / scores (one all-gather).  Prints ONE JSON line on rank 0.
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

CONFIGS = {
    "C2": dict(H=480, W=640, K=1024, D=128, kenc=[32, 64, 128], T=30, pairs=64, sp="allss128", sg_seed=0,
               kind="matching", gf_pair=135.43, gf_qk=9.664, baseline_index=1),
    "C4": dict(H=480, W=640, K=1024, D=128, kenc=[32, 64, 128], T=30, pairs=128, sp="allss128", sg_seed=0,
               kind="matching", gf_pair=135.43, gf_qk=9.664, baseline_index=3),
    "C3": dict(H=960, W=1280, K=2048, D=256, kenc=[32, 64, 128, 256], T=30, pairs=256, sp="coco256", sg_seed=1,
               kind="matching", gf_pair=671.7, gf_qk=77.3, baseline_index=2),
    "C5": dict(H=480, W=640, K=4096, D=128, kenc=[32, 64, 128], T=100, pairs=8, sp=None, sg_seed=5,
               kind="superglue", gf_pair=362.6, gf_qk=154.6, baseline_index=4),
}
SP_FILES = {"allss128": ("superpoint_allss128_weights.npz", "reference SuperPoint checkpoint (allss, D=128)"),
            "coco256": ("superpoint_coco256_weights.npz", "reference SuperPoint checkpoint (coco, D=256)")}


def make_cfg(c):
    return {"superpoint": {"descriptor_dim": c["D"], "nms_radius": 4, "keypoint_threshold": 0.005,
                           "max_keypoints": c["K"], "remove_borders": 4},
            "superglue": {"descriptor_dim": c["D"], "keypoint_encoder": list(c["kenc"]),
                          "GNN_layers": ["self", "cross"] * 9, "sinkhorn_iterations": c["T"],
                          "match_threshold": 0.2}}


def load_weights(c):
    from image_matching_b200 import synth
    sp, sp_name = None, "no SuperPoint (features supplied)"
    if c["sp"]:
        fn, sp_name = SP_FILES[c["sp"]]
        real = os.path.join(ROOT, "tests", "golden", fn)
        if os.path.exists(real):
            sp = dict(np.load(real))
        else:
            sp, sp_name = synth.superpoint_weights(0, c["D"]), "seeded synthetic SuperPoint weights"
    return sp, synth.superglue_weights(c["sg_seed"], c["D"], c["kenc"]), sp_name


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------- algorithmic work per kernel
def kernel_work(c, B):
    """Algorithmic work of ONE step (B pairs) per kernel name of the library's per-launch profiler (SURVEY.md 8d
    formulas, stated in DESIGN.md section 5): ("tensor", FLOPs) for the GEMM-class kernels -- every algorithmic
    multiply-add counted ONCE, not x3 for the fp16 hi/lo operand split -- and ("hbm", bytes) for the streaming ones."""
    H, W, K, D, T = c["H"], c["W"], c["K"], c["D"], c["T"]
    h, w = H // 8, W // 8
    w_ = {}
    if c["kind"] == "matching":
        n_img = 2 * B
        conv = lambda cin, cout, hh, ww, k=3: 2.0 * k * k * cin * cout * hh * ww      # noqa: E731
        w_["tc_conv3x3_stem"] = ("tensor", n_img * (conv(1, 64, H, W) + conv(64, 64, H, W)))
        w_["tc_conv3x3"] = ("tensor", n_img * (2 * conv(64, 64, H // 2, W // 2) + conv(64, 128, H // 4, W // 4)
                                              + conv(128, 128, H // 4, W // 4) + 2 * conv(128, 128, h, w)
                                              + conv(128, 512, h, w)))
        w_["tc_conv1x1"] = ("tensor", n_img * (conv(256, 65, h, w, 1) + conv(256, D, h, w, 1)))
        w_["softmax_heat"] = ("hbm", n_img * (65 * h * w * 4 + H * W * 4))
        w_["nms_candidates"] = ("hbm", n_img * (8 * h * 8 * w * 4))
        w_["sample_descriptors"] = ("hbm", n_img * (K * 4 * D * 4 + K * D * 4 * 2))
    N = M = K
    tok = 2 * B * N
    w_["tc_attention"] = ("tensor", 18 * 2 * B * 4.0 * N * M * D)
    gnn = tok * 2.0 * D * D * (7 * 18 + 3 * 17)        # merge + MLP (7 D^2 per token and layer) + the next layer's q|k|v
    kch = [3] + list(c["kenc"]) + [D]
    lin = tok * 2.0 * (sum(a * b for a, b in zip(kch[:-1], kch[1:])) + 3 * D * D + 2 * D * D) + B * 2.0 * N * M * D
    if D == 128:
        w_["tc_gnn_layer"] = ("tensor", gnn)           # one fused kernel per layer (tc_gnn.cu)
        w_["tc_gemm"] = ("tensor", lin)
    else:
        w_["tc_gemm"] = ("tensor", lin + gnn)          # other widths: the layers run as tc_gemm launches
    w_["ot_iter_fused"] = ("hbm", B * T * 4.0 * N * M)
    w_["argmax"] = ("hbm", B * 2 * 4.0 * N * M)
    return w_


def traffic_table(cname):
    """dram__bytes_read.sum + dram__bytes_write.sum per kernel name and step from the committed ncu capture of this
    config (profiles/r02_ncu_traffic.json, written by profiles/tools/ncu_step_table.py), or {}."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(cname, {})
    return {}


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


# --------------------------------------------------------------------------- CPU arm
class CpuArm:
    """The reference's CPU implementation of the path on the host cores: the reference's OWN modules when
    oracle/_ref exists (oracle/make_ref.py copies them, unmodified, where the reference checkout is present; kind
    "reference"), else the torch-CPU oracle port (kind "port").  One call = one pair (the reference's own caller is
    a batch_size=1 loop, superpoint_glue_test.py:65-78)."""

    def __init__(self, c):
        import torch
        from oracle import make_ref
        torch.set_num_threads(os.cpu_count() or 1)
        torch.set_grad_enabled(False)
        self.torch, self.c, self.cfg = torch, c, make_cfg(c)
        sp, sg, self.sp_name = load_weights(c)
        self.cores = torch.get_num_threads()
        self.kind = "reference" if make_ref.available() else "port"
        if self.kind == "reference":
            from image_matching_b200 import synth
            if sp is None:      # the reference's Matching always owns a SuperPoint; it is skipped when features are given
                sp = synth.superpoint_weights(0, c["D"])
            self.m = make_ref.load_matching(self.cfg, sp, sg)
            self.what = "the reference's own modules (oracle/_ref, unmodified)"
        else:
            from oracle import matching_oracle_torch as O
            self.O, self.sp, self.sg = O, sp, sg
            self.what = "the torch-CPU oracle port (oracle/matching_oracle_torch.py)"

    def inputs(self, seed):
        from image_matching_b200 import synth
        c = self.c
        if c["kind"] == "matching":
            return synth.make_pair(seed, c["H"], c["W"])
        return c5_features(seed, 1, c)

    def run(self, inp):
        t, c = self.torch, self.c
        if c["kind"] == "matching":
            a, b = inp
            if self.kind == "reference":
                return self.m({"image0": t.from_numpy(a[None, None]), "image1": t.from_numpy(b[None, None])})
            return self.O.matching_forward(a, b, self.sp, self.sg, self.cfg)
        kp0, sc0, de0, kp1, sc1, de1 = inp
        if self.kind == "reference":
            z = t.zeros(1, 1, c["H"], c["W"])
            return self.m({"image0": z, "image1": z, "keypoints0": t.from_numpy(kp0), "scores0": t.from_numpy(sc0),
                           "descriptors0": t.from_numpy(de0), "keypoints1": t.from_numpy(kp1),
                           "scores1": t.from_numpy(sc1), "descriptors1": t.from_numpy(de1)})
        return self.O.superglue_forward(t.from_numpy(kp0[0]), t.from_numpy(sc0[0]), t.from_numpy(de0[0]),
                                        t.from_numpy(kp1[0]), t.from_numpy(sc1[0]), t.from_numpy(de1[0]),
                                        c["H"], c["W"], self.sg, self.cfg)

    def pairs_per_s(self, n_pairs, budget_s, seeds_from=1000):
        """Bounded sample: up to n_pairs pairs, stopping early once budget_s of CPU work is spent."""
        self.run(self.inputs(seeds_from - 1))                 # warm-up (thread pool, page faults)
        inps = [self.inputs(seeds_from + i) for i in range(n_pairs)]
        self.results = []                                     # (inputs, prediction) of every timed pair: parity_vs_cpu()
        done, t0 = 0, time.perf_counter()
        for inp in inps:
            self.results.append((inp, self.run(inp)))
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return done / dt, dt, done


def parity_vs_cpu(arm, m, dev):
    """The pairs the CPU arm just timed, once more through the CUDA path (one pair per call, like the reference's
    caller): keypoint-set and match-pair differences against the CPU arm's own results.  A checker riding on the
    cpu_baseline leg -- it runs after every timed region."""
    import torch

    def arr(x):
        return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x)

    def kp_set(k):
        return set(map(tuple, arr(k).astype(np.int64).tolist()))

    def pairs(k0, k1, m0):
        k0, k1 = arr(k0).astype(np.int64), arr(k1).astype(np.int64)
        return {(tuple(k0[i]), tuple(k1[j])) for i, j in enumerate(arr(m0).tolist()) if j >= 0}

    nk = kf = nm = mf = 0
    for inp, ref in arm.results:
        if arm.c["kind"] == "matching":
            a, b = inp
            got = m({"image0": torch.from_numpy(a[None, None]).to(dev), "image1": torch.from_numpy(b[None, None]).to(dev)})
            gk0, gk1 = got["keypoints0"][0], got["keypoints1"][0]
            rk0, rk1 = ref["keypoints0"], ref["keypoints1"]
            if arm.kind == "reference":               # lists of per-image tensors; the port returns the single pair's arrays
                rk0, rk1 = rk0[0], rk1[0]
            for r, g in ((rk0, gk0), (rk1, gk1)):
                nk += len(kp_set(r))
                kf += len(kp_set(r) ^ kp_set(g))
        else:
            keys = ["keypoints0", "scores0", "descriptors0", "keypoints1", "scores1", "descriptors1"]
            z = torch.zeros(1, 1, arm.c["H"], arm.c["W"], device=dev)
            got = m(dict(zip(keys, [torch.from_numpy(x).to(dev) for x in inp]), image0=z, image1=z))
            rk0 = gk0 = inp[0][0]
            rk1 = gk1 = inp[3][0]
        rm0 = arr(ref[0] if isinstance(ref, tuple) else ref["matches0"])
        rp, gp = pairs(rk0, rk1, rm0[0] if rm0.ndim == 2 else rm0), pairs(gk0, gk1, arr(got["matches0"])[0])
        nm += len(rp)
        mf += len(rp ^ gp)
    return {"pairs": len(arm.results), "keypoints": nk, "keypoint_flips": kf, "matches": nm, "match_flips": mf,
            "against": arm.what}


def c5_features(seed, B, c):
    """Config-5 inputs: uniform keypoints, U(0,1) scores, unit-norm 128-d descriptors; image1's descriptors are a
    permuted noisy copy of image0's so that real matches exist."""
    from image_matching_b200 import synth
    N, D, H, W = c["K"], c["D"], c["H"], c["W"]
    kp0, sc0, de0 = synth.random_features(2 * seed + 1, B, N, D, H, W)
    kp1, sc1, de1 = synth.random_features(2 * seed + 2, B, N, D, H, W)
    perm = np.random.default_rng(seed).permutation(N)
    de1 = (0.6 * de0[:, :, perm] + 0.4 * de1).astype(np.float32)
    de1 /= np.linalg.norm(de1, axis=1, keepdims=True)
    return kp0, sc0, de0, kp1, sc1, de1


def workload_config(cname, c, batch_per_gpu, sp_name):
    if c["kind"] == "matching":
        desc = (f"{cname} (BASELINE.json configs[{c['baseline_index']}]): batch={batch_per_gpu} pairs/GPU of synthetic "
                f"{c['W']}x{c['H']} grayscale (random rectangles + perspective warp), max_keypoints={c['K']}, "
                f"desc_dim={c['D']}, kenc {c['kenc']}, 18 GNN layers, {c['T']} Sinkhorn iterations")
        l2 = (f"inputs ({2 * batch_per_gpu * c['H'] * c['W'] * 4 / 1e6:.0f} MB/step) and the multi-GB activation arena "
              "exceed the 126 MB L2; no flush needed")
    else:
        desc = (f"{cname} (BASELINE.json configs[{c['baseline_index']}]): SuperGlue only, batch={batch_per_gpu} pairs/GPU, "
                f"{c['K']} supplied keypoints per image with unit-norm {c['D']}-d descriptors (synthetic), "
                f"18 GNN layers, {c['T']} Sinkhorn iterations")
        l2 = (f"the score matrices ({batch_per_gpu * c['K'] * c['K'] * 4 / 1e6:.0f} MB/step) and token buffers exceed "
              "the 126 MB L2; no flush needed")
    return {"workload": desc, "name": cname, "pairs_per_gpu": batch_per_gpu, "image": [c["H"], c["W"]],
            "max_keypoints": c["K"], "descriptor_dim": c["D"], "gnn_layers": 18, "sinkhorn_iterations": c["T"],
            "weights": sp_name + " + seeded synthetic SuperGlue", "l2": l2}


def run_reference(args, cname, c):
    """`--impl reference`: the reference's CPU implementation of the path on all host cores, one pair per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(c)
    inps = [arm.inputs(2000 + i) for i in range(args.warmup + args.steps)]
    for i in range(args.warmup):
        arm.run(inps[i])
    t0 = time.perf_counter()
    done = 0
    for i in range(args.warmup, args.warmup + args.steps):
        arm.run(inps[i])
        done += 1
        if time.perf_counter() - t0 > args.ref_budget:      # bounded: the whole run must end within a few minutes
            break
    dt = time.perf_counter() - t0
    v = done / dt
    sample = f"{done} steps x 1 pair of the {cname} workload through {arm.what}"
    args.steps = done
    line = {"impl": "reference", "metric": "image-pairs/sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cname, c, 1, arm.sp_name),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.gpus > 1:
        line["note"] = ("the CPU arm runs on rank 0 only: at N > 1 the driver's ratio compares N GPUs with ONE host "
                        f"process of {arm.cores} threads")
    _RESULT_LINE.append(json.dumps(line))


def run_b200(args, cname, c):
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
    B = args.batch or c["pairs"]
    H, W, D = c["H"], c["W"], c["D"]
    sp, sg, sp_name = load_weights(c)
    cfg = make_cfg(c)
    mc = {"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}
    m = Matching(mc).eval()
    if sp is not None:
        m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    m = m.to(dev)
    torch.set_grad_enabled(False)
    matching = c["kind"] == "matching"

    # ---- synthetic inputs: a pool of distinct seeds, tiled to the batch (generation is host-side numpy)
    n_unique = min(B, args.unique)
    reps = (B + n_unique - 1) // n_unique
    if matching:
        a, b = synth.make_pair_batch([rank * 100000 + i for i in range(n_unique)], H, W)
        host = [torch.from_numpy(np.concatenate([x] * reps)[:B]).pin_memory() for x in (a, b)]
        keys = ["image0", "image1"]
    else:
        feats = c5_features(rank * 1000 + 7, n_unique, c)
        host = [torch.from_numpy(np.ascontiguousarray(np.concatenate([x] * reps)[:B])).pin_memory() for x in feats]
        keys = ["keypoints0", "scores0", "descriptors0", "keypoints1", "scores1", "descriptors1"]
    resident = [t.to(dev) for t in host]
    dummy = torch.empty(1, 1, H, W, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    outs = {}                     # result buffers per batch size, reused step after step (stable pointers: the library
                                  # replays the CUDA graph it captured for the call)

    def run_model(tensors):
        """One forward on device tensors without any host synchronisation -> dict with matches0 / matching_scores0."""
        if matching:
            nb = tensors[0].shape[0]
            outs[nb] = m.forward_device(tensors[0], tensors[1], out=outs.get(nb))
            return outs[nb]
        data = dict(zip(keys, tensors), image0=dummy, image1=dummy)
        return m.superglue.forward(data, _engine=m._engine)

    def step_device():
        out = run_model(resident)
        if world > 1:
            gather_matches(out["matches0"], out["matching_scores0"], B * world, handle=m._engine.handle)
        return out

    # End to end through the public API with HOST buffers.  Every step copies its own inputs from pinned host memory
    # and reads its results back; the copy of step i+1's inputs runs on a side stream while step i computes (what any
    # input pipeline does), so the PCIe transfer is inside the timed region but off the critical path.
    copy_stream = torch.cuda.Stream(device=dev)
    src = {"t": host}

    def upload():
        with torch.cuda.stream(copy_stream):
            xs = [t.to(dev, non_blocking=True) for t in src["t"]]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xs, ev

    pending = []
    host_out = {}                 # pinned result buffers of the end-to-end leg

    def step_e2e():
        if not pending:
            pending.append(upload())
        xs, ev = pending.pop()
        pending.append(upload())                       # next step's inputs start moving now
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for x in xs:
            x.record_stream(cur)
        if matching:
            pred = m({"image0": xs[0], "image1": xs[1]})
            extra = [torch.stack(pred["keypoints0"]), torch.stack(pred["keypoints1"])]
        else:
            pred = m(dict(zip(keys, xs), image0=dummy, image1=dummy))
            extra = []
        m0, s0 = pred["matches0"], pred["matching_scores0"]
        if world > 1:
            # one all-gather; only rank 0 unpacks and reads the whole job's matches back, the others their own shard
            g0, gs0 = gather_matches(m0, s0, B * world, handle=m._engine.handle, dst=0)
            if rank == 0:
                m0, s0 = g0, gs0
        # what the reference's caller reads back per pair (superpoint_glue_test.py:79-82), into pinned host buffers (one
        # stream sync for all of them instead of one pageable copy + sync per tensor)
        res = []
        for i, t in enumerate([m0, s0] + extra):
            key = (i, tuple(t.shape), t.dtype)
            if key not in host_out:
                host_out[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host_out[key].copy_(t, non_blocking=True)
            res.append(host_out[key])
        cur.synchronize()
        return res

    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    if matching:
        counts = out["counts"].cpu().numpy()
    else:
        counts = np.full((2, B), c["K"])
    valid = int((out["matches0"] > -1).sum().item())

    if args.ncu_step:                      # exactly one step between cudaProfilerStart/Stop (ncu --profile-from-start off)
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

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
    h2d = int(sum(t.numel() * t.element_size() for t in host))
    d2h = int(sum(r.numel() * r.element_size() for r in res) + (8 * B if matching else 0))
    extra_lines = {}
    e2e_u8_ms = 0.0
    if matching and not args.quick:
        # informational: the same loop uploading the 8-bit pixels (b200m_matching_forward_u8 normalises on the device;
        # the reference's loader would upload 4x / 8x more bytes as float32 / float64)
        src["t"] = [torch.from_numpy(np.round(t.numpy() * 255).astype(np.uint8)).pin_memory() for t in host]
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
            o = m.forward_device(resident[0], resident[1])
            mats, inl, info = estimate_affine_partial_2d(m, o["keypoints0"], o["keypoints1"], o["matches0"],
                                                         o["counts"][0], 7.0)
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
        extra_lines["with_registration"] = {
            "value": B * args.steps / (reg_ms / 1e3), "unit": "pairs/s (this rank)", "ms_per_step": reg_ms / args.steps,
            "inliers_per_pair": float(reg[2][:, 1].float().mean()),
            "ransac_iterations_per_pair": float(reg[2][:, 2].float().mean()),
            "note": "forward_device + b200m_estimate_affine_partial (cv2-identical RANSAC, 7 px) + D2H of matrices "
                    "and inlier masks"}
    # informational: single-pair latency (the reference script's own batch_size=1 loop, superpoint_glue_test.py:66,72)
    one = [t[:1].contiguous() for t in resident]
    for _ in range(3):
        run_model(one)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(20):
        run_model(one)
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
        lib.check(L.b200m_profile_begin(m._engine.handle, 40000))
        nprof = min(args.steps, 2)
        for _ in range(nprof):
            run_model(resident)
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
        work = kernel_work(c, B)
        traffic = traffic_table(cname)
        sum_ms = sum(p["ms_per_step"] for p in prof.values()) or 1.0
        per_kernel = {}
        for k, p in sorted(prof.items()):
            e = {"ms_per_step": round(p["ms_per_step"], 4), "launches_per_step": p["launches_per_step"]}
            if k in work:
                bound, amount = work[k]
                if bound == "tensor":
                    e["tflops_algorithmic"] = round(amount / (p["ms_per_step"] / 1e3) / 1e12, 2)
                    e["frac_of_tensor_peak"] = round(e["tflops_algorithmic"] / pk["tf_sustained"], 4)
                else:
                    e["gbs_algorithmic"] = round(amount / (p["ms_per_step"] / 1e3) / 1e9, 1)
                    e["frac_of_hbm_peak"] = round(e["gbs_algorithmic"] / pk["hbm_gbs"], 4)
            per_kernel[k] = e
        # roofline of the DOMINANT kernel (largest share of the step among the kernels with a work model)
        roof = None
        cand = [k for k in prof if k in work]
        if cand:
            dom = max(cand, key=lambda k: prof[k]["ms_per_step"])
            bound, amount = work[dom]
            n_launch = max(prof[dom]["launches_per_step"], 1)
            k_ms = prof[dom]["ms_per_step"]
            if bound == "tensor":
                ach, peak, unit = amount / (k_ms / 1e3) / 1e12, pk["tf_sustained"], "TFLOP/s"
                note = ("achieved counts every algorithmic multiply-add once; the kernel issues 3 fp16 products per "
                        "algorithmic product where it runs the fp32-class hi/lo operand split")
            else:
                ach, peak, unit = amount / (k_ms / 1e3) / 1e9, pk["hbm_gbs"], "GB/s"
                note = "achieved = compulsory (algorithmic) bytes / time"
            tr = traffic.get(dom)
            roof = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": (tr["dram_bytes_per_step"] / n_launch * (B / tr["pairs_per_step"])) if tr else None,
                    "traffic_source": ("profiles/r02_ncu_traffic.json (ncu dram__bytes_read+write, scaled to this "
                                       "batch)") if tr else None,
                    "peak_source": pk["source"] + (", bf16 dense sustained (kernel timed inside a long step)"
                                                   if bound == "tensor" else ", copy bandwidth"),
                    "algorithmic_work_per_launch": amount / n_launch, "launches_per_step": n_launch,
                    "avg_launch_ms": k_ms / n_launch, "share_of_step": k_ms / sum_ms, "note": note}
        line = {"metric": "image-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(cname, c, B, sp_name),
                "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps,
                        "note": "Matching.forward on fp32 host buffers; the upload of step i+1 overlaps step i on a side "
                                "stream; at N > 1 one all-gather, rank 0 reads the whole job's matches back"},
                "latency_batch1_ms": lat_ms,
                "gpu_launches": int(launches),
                "cuda_graph_replays": m._engine.graph_replays(),
                "clocks": sampler.result(),
                "roofline": roof,
                "qk_roofline": {"gflop_per_pair": c["gf_qk"], "achieved_tflops": value / world * c["gf_qk"] / 1e3,
                                "frac_of_bf16_peak": value / world * c["gf_qk"] / 1e3 / pk["tf_sustained"]},
                "whole_path_tflops_per_gpu": value / world * c["gf_pair"] / 1e3,
                "kernels": per_kernel,
                "kernel_ms_sum": round(sum_ms, 3),
                "check": {"keypoints_per_image_min": int(counts.min()), "keypoints_per_image_max": int(counts.max()),
                          "valid_matches_per_pair": valid / B}}
        if e2e_u8_ms:
            line["e2e_uint8_images"] = {"value": total_pairs * args.steps / (e2e_u8_ms / 1e3), "unit": "pairs/s",
                                        "h2d_bytes_per_step": h2d // 4, "ms_per_step": e2e_u8_ms / args.steps}
        line.update(extra_lines)
        if world == 1 and not args.no_cpu:
            arm = CpuArm(c)
            v, dt, done = arm.pairs_per_s(args.cpu_pairs, args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind,
                                    "sample": f"{done} pairs of the {cname} workload through {arm.what} ({dt:.1f} s)"}
            line["parity_vs_cpu_arm"] = parity_vs_cpu(arm, m, dev)
        _RESULT_LINE.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration; default C2 on one GPU, C4 (128 pairs/GPU) under --gpus N > 1")
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU per step (default: the configuration's)")
    ap.add_argument("--unique", type=int, default=16, help="distinct synthetic pairs generated per rank")
    ap.add_argument("--cpu-pairs", type=int, default=16, help="most pairs timed through the CPU arm (cpu_baseline)")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work after which the sample stops")
    ap.add_argument("--ref-budget", type=float, default=240.0,
                    help="--impl reference: stop after this many seconds of timed steps (steps reports what ran)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the informational uint8 / registration legs")
    ap.add_argument("--ncu-step", action="store_true",
                    help="warm up, then run exactly ONE step between cudaProfilerStart/Stop and exit (no JSON line)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cname = args.config or ("C4" if max(args.gpus, world) > 1 else "C2")
    c = CONFIGS[cname]
    # stdout carries exactly ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr, so whatever a
    # library prints to stdout from C code (NCCL's version banner under NCCL_DEBUG=VERSION, for one) cannot precede it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args, cname, c)
        else:
            run_b200(args, cname, c)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _RESULT_LINE:
        print(_RESULT_LINE[0], flush=True)


if __name__ == "__main__":
    main()
