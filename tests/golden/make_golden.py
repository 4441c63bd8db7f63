"""Generate golden vectors by running the UNMODIFIED reference modules (imported from
/root/reference, build container only) on seeded synthetic inputs/weights.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The fixtures are committed; nothing at test/bench time reads /root/reference.
Recipe follows SURVEY.md Appendix B (no edits to the reference; weights are loaded
with load_state_dict so the reference's own torch.load path is not needed).
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("B200M_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from image_matching_b200 import synth  # noqa: E402


def to_torch_sd(sd):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}


def build_reference(cfg, sp_sd, sg_sd):
    from superglue.models.matching_test import Matching  # the reference
    torch.set_grad_enabled(False)
    c = {"superpoint": dict(cfg["superpoint"], weights=None),
         "superglue": dict(cfg["superglue"], weights="")}
    m = Matching(c).eval()
    m.superpoint.load_state_dict(to_torch_sd(sp_sd))
    m.superglue.load_state_dict(to_torch_sd(sg_sd))
    return m


def make_cfg(D=128, kenc=(32, 64, 128), max_kp=1024, iters=30, mthr=0.2, kthr=0.005):
    return {"superpoint": {"descriptor_dim": D, "nms_radius": 4, "keypoint_threshold": kthr,
                           "max_keypoints": max_kp, "remove_borders": 4},
            "superglue": {"descriptor_dim": D, "keypoint_encoder": list(kenc),
                          "GNN_layers": ["self", "cross"] * 9,
                          "sinkhorn_iterations": iters, "match_threshold": mthr}}


def run_case(name, H, W, seeds, cfg, sp_sd, sg_sd, stages=False):
    m = build_reference(cfg, sp_sd, sg_sd)
    a, b = synth.make_pair_batch(seeds, H, W)
    out = {"seeds": np.asarray(seeds), "H": H, "W": W}
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    for k, v in pred.items():
        if isinstance(v, (list, tuple)):
            for i, t in enumerate(v):
                out[f"{k}_{i}"] = t.numpy()
        else:
            out[k] = v.numpy()
    if stages:
        # stage-level intermediates of pair 0
        sp = m.superpoint
        from superpoint.models.superpoint_test import simple_nms
        for side, img in (("0", a[:1]), ("1", b[:1])):
            x = torch.from_numpy(img)
            x4 = sp.down3(sp.down2(sp.down1(sp.inc(x))))
            semi = sp.bnPb(sp.convPb(sp.relu(sp.bnPa(sp.convPa(x4)))))
            desc = sp.bnDb(sp.convDb(sp.relu(sp.bnDa(sp.convDa(x4)))))
            desc = desc / torch.norm(desc, p=2, dim=1, keepdim=True)
            sc = torch.nn.functional.softmax(semi, 1)[:, :-1]
            bb, _, h, w = sc.shape
            sc = sc.permute(0, 2, 3, 1).reshape(bb, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(bb, h * 8, w * 8)
            out["semi_" + side] = semi[0].numpy()
            out["desc_" + side] = desc[0].numpy()
            out["heat_" + side] = sc[0].numpy()
            out["nms_" + side] = simple_nms(sc, 4)[0].numpy()
        sg = m.superglue
        from superglue.models.superglue_test import normalize_keypoints, log_optimal_transport
        k0, k1 = pred["keypoints0"][0][None], pred["keypoints1"][0][None]
        d0 = pred["descriptors0"][0][None] + sg.kenc(normalize_keypoints(k0, a[:1].shape), pred["scores0"][0][None])
        d1 = pred["descriptors1"][0][None] + sg.kenc(normalize_keypoints(k1, b[:1].shape), pred["scores1"][0][None])
        out["kenc0"], out["kenc1"] = d0[0].numpy(), d1[0].numpy()
        l0 = sg.gnn.layers[0]
        out["layer0_delta0"] = l0(d0, d0)[0].numpy()
        out["layer1_delta0"] = sg.gnn.layers[1](d0, d1)[0].numpy()
        g0, g1 = sg.gnn(d0, d1)
        out["gnn0"], out["gnn1"] = g0[0].numpy(), g1[0].numpy()
        m0, m1 = sg.final_proj(g0), sg.final_proj(g1)
        S = torch.einsum("bdn,bdm->bnm", m0, m1) / cfg["superglue"]["descriptor_dim"] ** .5
        out["S"] = S[0].numpy()
        out["Z"] = log_optimal_transport(S, sg.bin_score, cfg["superglue"]["sinkhorn_iterations"])[0].numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB",
          "kpts:", [int(pred["keypoints0"][i].shape[0]) for i in range(len(seeds))],
          "valid matches:", [int((pred["matches0"][i] > -1).sum()) for i in range(len(seeds))])


def run_official_case(name, H, W, seeds, cfg, sp_sd, sg_sd):
    """The "official" variant (superglue/models/matching.py + superpoint.py: SuperPoint without BatchNorm).  Its
    constructor insists on torch.load-ing weights/superpoint_v1.pth (a git-LFS stub in the reference tree), so
    torch.load is patched FOR THE CONSTRUCTOR CALL ONLY to hand it the seeded synthetic state_dict; the reference
    code itself is untouched."""
    import superglue.models.superpoint as ref_sp
    from superglue.models.matching import Matching
    torch.set_grad_enabled(False)
    real_load = torch.load
    torch.load = lambda *a, **k: to_torch_sd(sp_sd)
    try:
        m = Matching({"superpoint": dict(cfg["superpoint"]), "superglue": dict(cfg["superglue"], weights="")}).eval()
    finally:
        torch.load = real_load
    assert isinstance(m.superpoint, ref_sp.SuperPoint)
    m.superglue.load_state_dict(to_torch_sd(sg_sd))
    a, b = synth.make_pair_batch(seeds, H, W)
    out = {"seeds": np.asarray(seeds), "H": H, "W": W}
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    for k, v in pred.items():
        if isinstance(v, (list, tuple)):
            for i, t in enumerate(v):
                out[f"{k}_{i}"] = t.numpy()
        else:
            out[k] = v.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB",
          "kpts:", [int(pred["keypoints0"][i].shape[0]) for i in range(len(seeds))],
          "valid matches:", [int((pred["matches0"][i] > -1).sum()) for i in range(len(seeds))])


def convert_real_superpoint_weights():
    """The reference's trained SuperPoint checkpoint (D=128, the one BASELINE's C1 is quoted
    on) is a 15 MB CUDA-saved torch pickle with optimizer state; keep only the 84 model
    tensors, `module.` prefix stripped exactly as superpoint_test.py:87-100 does, as a
    plain .npz so tests/bench on the GPU box (no /root/reference there) can use it."""
    ck = torch.load(os.path.join(REF, "superpoint/models/weights/superPointNet_allss_descriptor_128.pth.tar"),
                    map_location="cpu")
    sd = {(k[7:] if "module" in k else k): v.numpy() for k, v in ck["model_state_dict"].items()}
    path = os.path.join(HERE, "superpoint_allss128_weights.npz")
    np.savez_compressed(path, **sd)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    return sd


def convert_coco256_superpoint_weights():
    """The reference's D=256 SuperPoint checkpoint (superPointNet_coco_descriptor_256.pth.tar, the one SURVEY.md 8(c)
    names for BASELINE config 3), converted like the D=128 one."""
    ck = torch.load(os.path.join(REF, "superpoint/models/weights/superPointNet_coco_descriptor_256.pth.tar"),
                    map_location="cpu")
    sd = {(k[7:] if "module" in k else k): v.numpy() for k, v in ck["model_state_dict"].items()}
    path = os.path.join(HERE, "superpoint_coco256_weights.npz")
    np.savez_compressed(path, **sd)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    return sd


def main_c3():
    """BASELINE config 3 at its stated size: one 1280x960 pair, the reference's COCO D=256 SuperPoint checkpoint,
    2048 keypoints, kenc [32,64,128,256], 18 layers, 30 Sinkhorn iterations.  Only the boundary outputs are kept
    (keypoints, scores, matches, matching scores, every 16th descriptor column): the full descriptors would be 4 MB."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    sp = convert_coco256_superpoint_weights()
    sg = synth.superglue_weights(1, 256, (32, 64, 128, 256))
    cfg = make_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=2048, iters=30)
    m = build_reference(cfg, sp, sg)
    a, b = synth.make_pair_batch([1], 960, 1280)
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    out = {"seeds": np.asarray([1]), "H": 960, "W": 1280}
    for k, v in pred.items():
        if isinstance(v, (list, tuple)):
            for i, t in enumerate(v):
                out[f"{k}_{i}"] = t.numpy()[:, ::16] if k.startswith("descriptors") else t.numpy()
        else:
            out[k] = v.numpy()
    path = os.path.join(HERE, "c3_real.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "kpts:", int(pred["keypoints0"][0].shape[0]),
          int(pred["keypoints1"][0].shape[0]), "valid matches:", int((pred["matches0"][0] > -1).sum()))


def convert_allss64_superpoint_weights():
    """The reference's D=64 SuperPoint checkpoint (superPointNet_allss_descriptor_64.pth.tar), converted like the others."""
    ck = torch.load(os.path.join(REF, "superpoint/models/weights/superPointNet_allss_descriptor_64.pth.tar"),
                    map_location="cpu")
    sd = {(k[7:] if "module" in k else k): v.numpy() for k, v in ck["model_state_dict"].items()}
    path = os.path.join(HERE, "superpoint_allss64_weights.npz")
    np.savez_compressed(path, **sd)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    return sd


def main_d64():
    """The third SuperPoint width the reference ships (D = 64, head_dim 16): one 640x480 pair, the reference's own
    checkpoint, 1024 keypoints, kenc [32, 64], 18 layers, 30 Sinkhorn iterations; boundary outputs only."""
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    sp = convert_allss64_superpoint_weights()
    sg = synth.superglue_weights(2, 64, (32, 64))
    cfg = make_cfg(D=64, kenc=(32, 64), max_kp=1024, iters=30)
    m = build_reference(cfg, sp, sg)
    a, b = synth.make_pair_batch([1], 480, 640)
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    out = {"seeds": np.asarray([1]), "H": 480, "W": 640}
    for k, v in pred.items():
        if isinstance(v, (list, tuple)):
            for i, t in enumerate(v):
                out[f"{k}_{i}"] = t.numpy()
        else:
            out[k] = v.numpy()
    path = os.path.join(HERE, "d64_real.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "kpts:", int(pred["keypoints0"][0].shape[0]),
          int(pred["keypoints1"][0].shape[0]), "valid matches:", int((pred["matches0"][0] > -1).sum()))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    real = convert_real_superpoint_weights()
    sp = synth.superpoint_weights(0, 128)
    sg = synth.superglue_weights(0, 128)
    # small: full stage intermediates (fast in the oracle, small on disk)
    run_case("small_stages", 120, 160, [1], make_cfg(max_kp=256), sp, sg, stages=True)
    # not a multiple of 8 / unbounded keypoints / row-major order branch
    run_case("ragged_hw", 123, 165, [2], make_cfg(max_kp=-1, iters=20), sp, sg)
    # C1: the metric's configuration (640x480, 1024 kpts, 18 layers, 30 Sinkhorn iterations)
    run_case("c1_pair", 480, 640, [1, 2], make_cfg(max_kp=1024), sp, sg)
    # C1 with the reference's trained SuperPoint weights
    run_case("c1_real", 480, 640, [1], make_cfg(max_kp=1024), real, sg)
    run_case("real_small_stages", 160, 224, [4], make_cfg(max_kp=300), real, sg, stages=True)
    # D=256 with the 4-layer keypoint encoder (config 3's model at a small image)
    sp256 = synth.superpoint_weights(1, 256)
    sg256 = synth.superglue_weights(1, 256, (32, 64, 128, 256))
    run_case("d256_small", 120, 160, [3], make_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=200, iters=50), sp256, sg256,
             stages=True)


def main_official():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    cfg = make_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=300, iters=30)
    run_official_case("official_small", 160, 224, [5], cfg, synth.superpoint_official_weights(1, 256),
                      synth.superglue_weights(1, 256, (32, 64, 128, 256)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "official":
        main_official()      # only the fixture added for SURVEY.md 8(f3); the others are unchanged
    elif len(sys.argv) > 1 and sys.argv[1] == "c3":
        main_c3()            # only the config-3 fixture (round 2); the others are unchanged
    elif len(sys.argv) > 1 and sys.argv[1] == "d64":
        main_d64()           # only the D = 64 fixture (round 2); the others are unchanged
    else:
        main()
        main_official()
        main_c3()
