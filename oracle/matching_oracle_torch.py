"""CPU ORACLE (torch-CPU flavour) -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

The same restatement as ``oracle/matching_oracle.py`` (which stays the primary, library-free checker), written
with ``torch.nn.functional`` ops on CPU tensors, because that is the arithmetic library the reference itself
runs on (all of its math is torch; README.md:23).  It exists for ONE purpose: ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` arm should time the reference's algorithm at the speed the reference's own CPU path reaches
(oneDNN convolutions, all host threads), not at the speed of a numpy port -- the numpy oracle is ~2x slower than
the real reference on the same cores, which would flatter every GPU/CPU ratio.  It is validated against the numpy
oracle and the reference-generated goldens in ``tests/test_oracle_golden.py``.

Every function cites the reference file:line it follows (paths relative to the reference root).
Parity pinning: see ``oracle/matching_oracle.py`` (reference-generated goldens under ``tests/golden/``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))


def _bn(x, sd, name):
    """BatchNorm eval (unet_parts.py:16,19; superglue_test.py:57)."""
    return F.batch_norm(x, _t(sd[name + ".running_mean"]), _t(sd[name + ".running_var"]), _t(sd[name + ".weight"]),
                        _t(sd[name + ".bias"]), False, 0.0, 1e-5)


def superpoint_dense(img, sd):
    """superpoint_test.py:113-126, unet_parts.py:10-48.  img (B,1,H,W) -> semi (B,65,h,w), desc (B,D,h,w)."""
    def double(x, p):
        x = F.relu(_bn(F.conv2d(x, _t(sd[p + ".0.weight"]), _t(sd[p + ".0.bias"]), padding=1), sd, p + ".1"))
        return F.relu(_bn(F.conv2d(x, _t(sd[p + ".3.weight"]), _t(sd[p + ".3.bias"]), padding=1), sd, p + ".4"))
    x = double(img, "inc.conv.conv")
    for d in ("down1", "down2", "down3"):
        x = double(F.max_pool2d(x, 2), d + ".mpconv.1.conv")
    cPa = F.relu(_bn(F.conv2d(x, _t(sd["convPa.weight"]), _t(sd["convPa.bias"]), padding=1), sd, "bnPa"))
    semi = _bn(F.conv2d(cPa, _t(sd["convPb.weight"]), _t(sd["convPb.bias"])), sd, "bnPb")
    cDa = F.relu(_bn(F.conv2d(x, _t(sd["convDa.weight"]), _t(sd["convDa.bias"]), padding=1), sd, "bnDa"))
    desc = _bn(F.conv2d(cDa, _t(sd["convDb.weight"]), _t(sd["convDb.bias"])), sd, "bnDb")
    return semi, desc / torch.norm(desc, p=2, dim=1, keepdim=True)


def simple_nms(scores, r):
    """superpoint_test.py:7-22."""
    def mp(x):
        return F.max_pool2d(x, kernel_size=2 * r + 1, stride=1, padding=r)
    zeros = torch.zeros_like(scores)
    mask = scores == mp(scores)
    for _ in range(2):
        supp = mp(mask.float()) > 0
        s2 = torch.where(supp, zeros, scores)
        mask = mask | ((s2 == mp(s2)) & ~supp)
    return torch.where(mask, scores, zeros)


def superpoint_forward(img, sd, cfg, align_corners=False):
    """superpoint_test.py:103-161 for ONE image (1,1,H,W)."""
    c = cfg["superpoint"]
    semi, desc = superpoint_dense(img, sd)
    p = F.softmax(semi, 1)[:, :-1]
    b, _, h, w = p.shape
    heat = p.permute(0, 2, 3, 1).reshape(b, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(b, h * 8, w * 8)
    nms = simple_nms(heat, c["nms_radius"])[0]
    kp = torch.nonzero(nms > c["keypoint_threshold"])
    sc = nms[kp[:, 0], kp[:, 1]]
    bd = c.get("remove_borders", 4)
    keep = (kp[:, 0] >= bd) & (kp[:, 0] < h * 8 - bd) & (kp[:, 1] >= bd) & (kp[:, 1] < w * 8 - bd)
    kp, sc = kp[keep], sc[keep]
    if c["max_keypoints"] >= 0 and c["max_keypoints"] < len(sc):
        sc, idx = torch.topk(sc, c["max_keypoints"], dim=0)
        kp = kp[idx]
    kp = torch.flip(kp, [1]).float()
    # sample_descriptors (:40-52)
    s = 8
    g = kp - s / 2 + 0.5
    g = g / torch.tensor([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5]) * 2 - 1
    d = F.grid_sample(desc, g.view(1, 1, -1, 2), mode="bilinear", align_corners=bool(align_corners))
    d = F.normalize(d.reshape(1, desc.shape[1], -1), p=2, dim=1)[0]
    return kp, sc, d


def _mlp(x, sd, prefix, n):
    """superglue_test.py:49-60: Conv1d(k=1) [+ BN + ReLU] blocks; layer indices 0,3,6,... (BN at +1)."""
    for i in range(n):
        j = 3 * i
        x = F.conv1d(x, _t(sd[f"{prefix}.{j}.weight"]), _t(sd[f"{prefix}.{j}.bias"]))
        if i + 1 < n:
            x = F.relu(_bn(x, sd, f"{prefix}.{j + 1}"))
    return x


def superglue_forward(kp0, sc0, de0, kp1, sc1, de1, H, W, sd, cfg):
    """superglue_test.py:230-285 for one pair; kp (N,2), sc (N,), de (D,N)."""
    c = cfg["superglue"]
    D = de0.shape[0]
    n0, n1 = kp0.shape[0], kp1.shape[0]
    if n0 == 0 or n1 == 0:
        return (np.full(n0, -1, np.int64), np.full(n1, -1, np.int64), np.zeros(n0, np.float32), np.zeros(n1, np.float32))

    def enc(kp, sc, de):
        ctr = torch.tensor([W / 2.0, H / 2.0])
        kn = (kp - ctr) / (max(W, H) * 0.7)                                            # :63-70
        x = torch.cat([kn.t(), sc[None]], 0)[None]                                      # :80-82
        return de[None] + _mlp(x, sd, "kenc.encoder", len(c["keypoint_encoder"]) + 1)

    x0, x1 = enc(kp0, sc0, de0), enc(kp1, sc1, de1)

    def prop(l, x, src):                                                                # :92-119
        p = f"gnn.layers.{l}"
        q = F.conv1d(x, _t(sd[p + ".attn.proj.0.weight"]), _t(sd[p + ".attn.proj.0.bias"])).view(1, D // 4, 4, -1)
        k = F.conv1d(src, _t(sd[p + ".attn.proj.1.weight"]), _t(sd[p + ".attn.proj.1.bias"])).view(1, D // 4, 4, -1)
        v = F.conv1d(src, _t(sd[p + ".attn.proj.2.weight"]), _t(sd[p + ".attn.proj.2.bias"])).view(1, D // 4, 4, -1)
        s = torch.einsum("bdhn,bdhm->bhnm", q, k) / (D // 4) ** 0.5
        m = torch.einsum("bhnm,bdhm->bdhn", F.softmax(s, -1), v).reshape(1, D, -1)
        m = F.conv1d(m, _t(sd[p + ".attn.merge.weight"]), _t(sd[p + ".attn.merge.bias"]))
        return _mlp(torch.cat([x, m], 1), sd, p + ".mlp", 2)

    for l, name in enumerate(c["GNN_layers"]):                                          # :127-138
        s0, s1 = (x1, x0) if name == "cross" else (x0, x1)
        d0, d1 = prop(l, x0, s0), prop(l, x1, s1)
        x0, x1 = x0 + d0, x1 + d1
    m0 = F.conv1d(x0, _t(sd["final_proj.weight"]), _t(sd["final_proj.bias"]))
    m1 = F.conv1d(x1, _t(sd["final_proj.weight"]), _t(sd["final_proj.bias"]))
    S = torch.einsum("bdn,bdm->bnm", m0, m1) / D ** 0.5
    # log_optimal_transport (:141-170)
    alpha = _t(sd["bin_score"]).reshape(())
    ms, ns = torch.tensor(float(n0)), torch.tensor(float(n1))
    C = torch.cat([torch.cat([S, alpha.expand(1, n0, 1)], -1), alpha.expand(1, 1, n1 + 1)], 1)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(n0), ns.log()[None] + norm])[None]
    log_nu = torch.cat([norm.expand(n1), ms.log()[None] + norm])[None]
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(c["sinkhorn_iterations"]):
        u = log_mu - torch.logsumexp(C + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(C + u.unsqueeze(2), dim=1)
    Z = C + u.unsqueeze(2) + v.unsqueeze(1) - norm
    # match selection (:268-278)
    mx0, mx1 = Z[:, :-1, :-1].max(2), Z[:, :-1, :-1].max(1)
    i0, i1 = mx0.indices, mx1.indices
    mut0 = torch.arange(n0)[None] == i1.gather(1, i0)
    mut1 = torch.arange(n1)[None] == i0.gather(1, i1)
    zero = Z.new_tensor(0)
    ms0 = torch.where(mut0, mx0.values.exp(), zero)
    ms1 = torch.where(mut1, ms0.gather(1, i1), zero)
    v0 = mut0 & (ms0 > c["match_threshold"])
    v1 = mut1 & v0.gather(1, i1)
    return (torch.where(v0, i0, i0.new_tensor(-1))[0].numpy(), torch.where(v1, i1, i1.new_tensor(-1))[0].numpy(),
            ms0[0].numpy(), ms1[0].numpy())


@torch.no_grad()
def matching_forward(img0, img1, sp_sd, sg_sd, cfg, align_corners=False):
    """matching_test.py:54-82 for one pair of (H,W) float32 images; same result keys as the numpy oracle."""
    H, W = img0.shape
    out = {}
    feats = []
    for side, im in (("0", img0), ("1", img1)):
        kp, sc, de = superpoint_forward(_t(im)[None, None], sp_sd, cfg, align_corners)
        feats.append((kp, sc, de))
        out["keypoints" + side], out["scores" + side], out["descriptors" + side] = kp.numpy(), sc.numpy(), de.numpy()
    (k0, s0, d0), (k1, s1, d1) = feats
    m0, m1, ms0, ms1 = superglue_forward(k0, s0, d0, k1, s1, d1, H, W, sg_sd, cfg)
    out.update(matches0=m0, matches1=m1, matching_scores0=ms0, matching_scores1=ms1)
    return out
