#!/usr/bin/env python
"""What would a single fp16 product (hi planes only) instead of the fp16x3 operand split cost in keypoint / match flips,
and what would it save?    One process per setting (the switch is read when the handle is
created): B200M_SINGLE = "", desc, gemm, gnn, attn, "gemm,gnn,attn", "desc,gemm,gnn,attn".

  (a) the reference-generated goldens (tests/golden): keypoint-set and match-pair differences against the reference;
  (b) 64 synthetic 640x480 pairs (the bench workload): match differences against THIS library's default (fp16x3) run,
      which itself has 0 flips on every golden -- a 30x larger sample than the goldens;
  (c) kernel times of one 64-pair step.

The library must be built with  make -C image_matching_b200/csrc clean all EXTRA=-DB200M_SINGLE_EXPERIMENT  (the product
build compiles the single-product code paths out and rejects B200M_SINGLE).

usage: python profiles/tools/single_product_experiment.py            (driver: runs every setting, prints a table)
       python profiles/tools/single_product_experiment.py --worker   (one setting, JSON on stdout)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SETTINGS = ["", "desc", "gemm", "gnn", "attn", "gemm,gnn,attn", "desc,gemm,gnn,attn"]


def worker():
    import ctypes as C
    import numpy as np
    import torch
    import bench
    import test_gpu_parity as tp
    from conftest import kp_set, match_pairs
    from image_matching_b200 import synth, lib
    res = {"golden": {}}
    for name in ["c1_real", "c1_pair", "small_stages", "d256_small", "c3_real"]:
        c = tp._case(name)
        g = c["g"]
        m = tp._matching(c["cfg"], c["sp"], c["sg"])
        a, b = synth.make_pair_batch(c["seeds"], c["H"], c["W"])
        try:
            pred = m({"image0": tp._t(a), "image1": tp._t(b)})
        except RuntimeError as e:            # ragged keypoint counts (a flip in the count itself)
            res["golden"][name] = {"error": str(e)[:80]}
            continue
        kf = mf = nm = nk = 0
        dmax = 0.0
        for i in range(len(c["seeds"])):
            for side in "01":
                ref = kp_set(g[f"keypoints{side}_{i}"])
                got = kp_set(pred["keypoints" + side][i].cpu().numpy())
                kf += len(ref ^ got)
                nk += len(ref)
                if ref == got:
                    gk = pred["keypoints" + side][i].cpu().numpy().astype(np.int64)
                    rk = g[f"keypoints{side}_{i}"].astype(np.int64)
                    pos = {tuple(k): j for j, k in enumerate(gk.tolist())}
                    idx = np.array([pos[tuple(k)] for k in rk.tolist()])
                    rd = g[f"descriptors{side}_{i}"]
                    gd = pred["descriptors" + side][i].cpu().numpy()[:, idx]
                    if rd.shape[1] != gd.shape[1]:
                        gd = gd[:, ::16]
                    dmax = max(dmax, float(np.abs(gd - rd).max()))
            rp = match_pairs(g[f"keypoints0_{i}"], g[f"keypoints1_{i}"], g["matches0"][i])
            gp = match_pairs(pred["keypoints0"][i].cpu().numpy(), pred["keypoints1"][i].cpu().numpy(),
                             pred["matches0"][i].cpu().numpy())
            mf += len(rp ^ gp)
            nm += len(rp)
        res["golden"][name] = {"keypoints": nk, "keypoint_flips": kf, "matches": nm, "match_flips": mf,
                               "descriptor_max_diff": dmax}
    # (b) + (c): the bench workload
    cfgc = bench.CONFIGS["C2"]
    sp, sg, _ = bench.load_weights(cfgc)
    cfg = bench.make_cfg(cfgc)
    m = tp._matching(cfg, sp, sg)
    a, b = synth.make_pair_batch(range(64), 480, 640)
    d0, d1 = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = m.forward_device(d0, d1)
    torch.cuda.synchronize()
    res["bench"] = {"matches0": out["matches0"].cpu().numpy().tolist(),
                    "keypoints0": out["keypoints0"].cpu().numpy().astype(int).tolist(),
                    "keypoints1": out["keypoints1"].cpu().numpy().astype(int).tolist()}
    L = lib.load()
    for _ in range(2):
        m.forward_device(d0, d1)
    torch.cuda.synchronize()
    lib.check(L.b200m_profile_begin(m._engine.handle, 20000))
    for _ in range(3):
        m.forward_device(d0, d1)
    buf = C.create_string_buffer(1 << 16)
    lib.check(L.b200m_profile_end(m._engine.handle, buf, len(buf)))
    prof = json.loads(buf.value.decode())
    res["kernel_ms"] = {k: round(v["ms"] / 3, 3) for k, v in prof.items() if v["ms"] / 3 > 0.05}
    res["step_ms"] = round(sum(v["ms"] for v in prof.values()) / 3, 3)
    print("RESULT " + json.dumps(res))


def pairs_of(r, i):
    k0, k1, m0 = r["keypoints0"][i], r["keypoints1"][i], r["matches0"][i]
    return {(tuple(k0[j]), tuple(k1[m0[j]])) for j in range(len(m0)) if m0[j] >= 0}


def main():
    if "--worker" in sys.argv:
        return worker()
    results = {}
    for s in SETTINGS:
        env = dict(os.environ, B200M_SINGLE=s)
        out = subprocess.run([sys.executable, __file__, "--worker"], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f"setting {s!r} failed:\n{out.stderr[-2000:]}")
            continue
        results[s] = json.loads(line[0][7:])
    base = results[""]
    print("| B200M_SINGLE | golden keypoint flips | golden match flips | descriptor max diff | bench match flips vs fp16x3 "
          "(64 pairs) | step ms | conv3x3 | attention | gnn layer | gemm |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for s, r in results.items():
        g = r["golden"]
        kf = sum(v.get("keypoint_flips", 0) for v in g.values())
        nk = sum(v.get("keypoints", 0) for v in g.values())
        mf = sum(v.get("match_flips", 0) for v in g.values())
        nm = sum(v.get("matches", 0) for v in g.values())
        err = [n for n, v in g.items() if "error" in v]
        dm = max(v.get("descriptor_max_diff", 0) for v in g.values())
        bf = bn = 0
        for i in range(64):
            pa, pb = pairs_of(base["bench"], i), pairs_of(r["bench"], i)
            bf += len(pa ^ pb)
            bn += len(pa)
        k = r["kernel_ms"]
        print(f"| {s or '(none: product path)'} | {kf} / {nk}{' + ragged: ' + ','.join(err) if err else ''} | {mf} / {nm} | "
              f"{dm:.1e} | {bf} / {bn} | {r['step_ms']} | {k.get('tc_conv3x3', 0)} | {k.get('tc_attention', 0)} | "
              f"{k.get('tc_gnn_layer', 0)} | {k.get('tc_gemm', 0)} |")
    for s, r in results.items():
        print(s or "(none)", json.dumps(r["golden"]))


if __name__ == "__main__":
    main()
