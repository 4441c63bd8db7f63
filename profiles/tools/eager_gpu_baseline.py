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


def sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def batched_pairs_per_s(O, to_dev, sp, sg, cfg, pairs, B, dev):
    """The same algorithm with the whole batch in every torch call, the way the reference's modules run when a caller
    stacks B pairs (Matching.forward supports it when the keypoint counts agree, matching_test.py:70-77): dense
    SuperPoint, NMS and SuperGlue on (B, ...) tensors; only the keypoint extraction loops over images like
    superpoint_test.py:135-155 does."""
    import torch.nn.functional as F
    c, g = cfg["superpoint"], cfg["superglue"]
    D = c["descriptor_dim"]
    reps = (B + len(pairs) - 1) // len(pairs)
    a = np.stack([p[0] for p in pairs] * reps)[:B]
    b = np.stack([p[1] for p in pairs] * reps)[:B]
    ha, hb = torch.from_numpy(a)[:, None], torch.from_numpy(b)[:, None]
    if dev.type == "cuda":
        ha, hb = ha.pin_memory(), hb.pin_memory()

    def superpoint(img):
        semi, desc = O.superpoint_dense(img, sp)
        p = F.softmax(semi, 1)[:, :-1]
        n, _, h, w = p.shape
        heat = p.permute(0, 2, 3, 1).reshape(n, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(n, h * 8, w * 8)
        nms = O.simple_nms(heat, c["nms_radius"])
        kps, scs, des = [], [], []
        for i in range(n):
            kp = torch.nonzero(nms[i] > c["keypoint_threshold"])
            sc = nms[i][kp[:, 0], kp[:, 1]]
            keep = (kp[:, 0] >= 4) & (kp[:, 0] < h * 8 - 4) & (kp[:, 1] >= 4) & (kp[:, 1] < w * 8 - 4)
            kp, sc = kp[keep], sc[keep]
            if c["max_keypoints"] >= 0 and c["max_keypoints"] < len(sc):
                sc, idx = torch.topk(sc, c["max_keypoints"], dim=0)
                kp = kp[idx]
            kp = torch.flip(kp, [1]).float()
            gr = (kp - 3.5) / torch.tensor([w * 8 - 4.5, h * 8 - 4.5]) * 2 - 1
            d = F.grid_sample(desc[i:i + 1], gr.view(1, 1, -1, 2), mode="bilinear", align_corners=False)
            des.append(F.normalize(d.reshape(D, -1), p=2, dim=0))
            kps.append(kp)
            scs.append(sc)
        return torch.stack(kps), torch.stack(scs), torch.stack(des)

    def mlp(x, prefix, n):
        for i in range(n):
            j = 3 * i
            x = F.conv1d(x, to_dev(sg[f"{prefix}.{j}.weight"]), to_dev(sg[f"{prefix}.{j}.bias"]))
            if i + 1 < n:
                x = F.relu(O._bn(x, sg, f"{prefix}.{j + 1}"))
        return x

    def superglue(k0, s0, d0, k1, s1, d1, H, W):
        def enc(kp, sc, de):
            kn = (kp - torch.tensor([W / 2.0, H / 2.0])) / (max(W, H) * 0.7)
            return de + mlp(torch.cat([kn.transpose(1, 2), sc[:, None]], 1), "kenc.encoder", len(g["keypoint_encoder"]) + 1)
        x0, x1 = enc(k0, s0, d0), enc(k1, s1, d1)

        def prop(l, x, src):
            p = f"gnn.layers.{l}"
            n = x.shape[0]
            q = F.conv1d(x, to_dev(sg[p + ".attn.proj.0.weight"]), to_dev(sg[p + ".attn.proj.0.bias"])).view(n, D // 4, 4, -1)
            k = F.conv1d(src, to_dev(sg[p + ".attn.proj.1.weight"]), to_dev(sg[p + ".attn.proj.1.bias"])).view(n, D // 4, 4, -1)
            v = F.conv1d(src, to_dev(sg[p + ".attn.proj.2.weight"]), to_dev(sg[p + ".attn.proj.2.bias"])).view(n, D // 4, 4, -1)
            s = torch.einsum("bdhn,bdhm->bhnm", q, k) / (D // 4) ** 0.5
            m = torch.einsum("bhnm,bdhm->bdhn", F.softmax(s, -1), v).reshape(n, D, -1)
            m = F.conv1d(m, to_dev(sg[p + ".attn.merge.weight"]), to_dev(sg[p + ".attn.merge.bias"]))
            return mlp(torch.cat([x, m], 1), p + ".mlp", 2)
        for l, name in enumerate(g["GNN_layers"]):
            a0, a1 = (x1, x0) if name == "cross" else (x0, x1)
            e0, e1 = prop(l, x0, a0), prop(l, x1, a1)
            x0, x1 = x0 + e0, x1 + e1
        m0 = F.conv1d(x0, to_dev(sg["final_proj.weight"]), to_dev(sg["final_proj.bias"]))
        m1 = F.conv1d(x1, to_dev(sg["final_proj.weight"]), to_dev(sg["final_proj.bias"]))
        S = torch.einsum("bdn,bdm->bnm", m0, m1) / D ** 0.5
        n, r, cc = S.shape
        alpha = to_dev(sg["bin_score"]).reshape(())
        C = torch.cat([torch.cat([S, alpha.expand(n, r, 1)], -1), alpha.expand(n, 1, cc + 1)], 1)
        ms, ns = torch.tensor(float(r)), torch.tensor(float(cc))
        norm = -(ms + ns).log()
        log_mu = torch.cat([norm.expand(r), ns.log()[None] + norm])[None].expand(n, -1)
        log_nu = torch.cat([norm.expand(cc), ms.log()[None] + norm])[None].expand(n, -1)
        u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
        for _ in range(g["sinkhorn_iterations"]):
            u = log_mu - torch.logsumexp(C + v.unsqueeze(1), dim=2)
            v = log_nu - torch.logsumexp(C + u.unsqueeze(2), dim=1)
        Z = C + u.unsqueeze(2) + v.unsqueeze(1) - norm
        mx0, mx1 = Z[:, :-1, :-1].max(2), Z[:, :-1, :-1].max(1)
        i0, i1 = mx0.indices, mx1.indices
        mut0 = torch.arange(r)[None] == i1.gather(1, i0)
        ms0 = torch.where(mut0, mx0.values.exp(), Z.new_tensor(0))
        v0 = mut0 & (ms0 > g["match_threshold"])
        return torch.where(v0, i0, i0.new_tensor(-1)), ms0

    def step():
        x0, x1 = ha.to(dev, non_blocking=True), hb.to(dev, non_blocking=True)
        k0, s0, d0 = superpoint(x0)
        k1, s1, d1 = superpoint(x1)
        m0, ms0 = superglue(k0, s0, d0, k1, s1, d1, x0.shape[2], x0.shape[3])
        return m0.cpu(), ms0.cpu(), k0.cpu(), k1.cpu()

    for _ in range(2):
        res = step()
    sync()
    t0 = time.perf_counter()
    n_steps = 3
    for _ in range(n_steps):
        res = step()
    sync()
    dt = time.perf_counter() - t0
    return {"pairs_per_step": B, "pairs_per_s": B * n_steps / dt, "ms_per_step": 1e3 * dt / n_steps,
            "valid_matches_per_pair": float((res[0] > -1).sum()) / B,
            "note": "needs equal keypoint counts across the batch (torch.stack), as the reference does"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20)
    ap.add_argument("--device", default="cuda:0", help="cpu only for a dry run of this script")
    ap.add_argument("--batch", type=int, default=64, help="also time a batched forward of this many pairs (0/1: skip)")
    ap.add_argument("--tf32", type=int, default=1, help="cudnn.allow_tf32 (torch default: 1); matmul TF32 stays off")
    args = ap.parse_args()
    import bench
    from image_matching_b200 import synth
    from oracle import matching_oracle_torch as O

    dev = torch.device(args.device)
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
    c2 = bench.CONFIGS["C2"]
    sp, sg, sp_name = bench.load_weights(c2)
    cfg = bench.make_cfg(c2)
    pairs = [synth.make_pair(3000 + i, c2["H"], c2["W"]) for i in range(args.pairs + 3)]
    for a, b in pairs[:3]:
        out = O.matching_forward(a, b, sp, sg, cfg)
    sync()
    t0 = time.perf_counter()
    for a, b in pairs[3:]:
        out = O.matching_forward(a, b, sp, sg, cfg)
    sync()
    dt = time.perf_counter() - t0
    batched = batched_pairs_per_s(O, to_dev, sp, sg, cfg, pairs[:16], args.batch, dev) if args.batch > 1 else None
    print(json.dumps({"baseline": "reference algorithm, eager torch %s on %s (cuDNN/cuBLAS/ATen), batch 1 per call"
                                  % (torch.__version__, torch.cuda.get_device_name(0) if dev.type == "cuda" else "cpu"),
                      "pairs_per_s": args.pairs / dt, "ms_per_pair": 1e3 * dt / args.pairs, "pairs": args.pairs,
                      "cudnn_allow_tf32": bool(args.tf32), "matmul_allow_tf32": False,
                      "includes": "H2D of the two fp32 images and D2H of keypoints / matches per pair, as the script does",
                      "keypoints": [int(len(out["keypoints0"])), int(len(out["keypoints1"]))],
                      "valid_matches": int((out["matches0"] > -1).sum()), "weights": sp_name,
                      "batched": batched}))


if __name__ == "__main__":
    main()
