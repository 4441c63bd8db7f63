"""CPU: pin the numpy oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py ran the unmodified reference modules)."""
import numpy as np
import pytest

from conftest import load_golden, golden_cfg, real_superpoint_weights, kp_set, match_pairs
from image_matching_b200 import synth
from oracle import matching_oracle as O


def _sp_stage_checks(g, sd, H, W, seed, side):
    img = synth.make_pair(seed, H, W)[int(side)]
    semi, desc = O.superpoint_dense(img, sd)
    assert np.abs(semi - g["semi_" + side]).max() < 2e-4
    assert np.abs(desc - g["desc_" + side]).max() < 2e-5
    # given the reference's semi, the post-processing is (nearly) bit exact
    heat = O.heatmap(g["semi_" + side])
    assert np.abs(heat - g["heat_" + side]).max() < 1e-6   # softmax: <= ~1 ulp vs torch
    nms = O.simple_nms(g["heat_" + side], 4)
    assert np.array_equal(nms, g["nms_" + side])          # compares only -> bit exact


@pytest.mark.parametrize("name,D,seed,hw,kw", [
    ("small_stages", 128, 1, (120, 160), dict(max_kp=256)),
    ("d256_small", 256, 3, (120, 160), dict(D=256, kenc=(32, 64, 128, 256), max_kp=200, iters=50)),
    ("real_small_stages", 128, 4, (160, 224), dict(max_kp=300)),
])
def test_stages(name, D, seed, hw, kw):
    g = load_golden(name)
    cfg = golden_cfg(**kw)
    if name.startswith("real"):
        sp = real_superpoint_weights()
    else:
        sp = synth.superpoint_weights(1 if D == 256 else 0, D)
    sg = synth.superglue_weights(1 if D == 256 else 0, D, cfg["superglue"]["keypoint_encoder"])
    H, W = hw
    for side in "01":
        _sp_stage_checks(g, sp, H, W, seed, side)
    # keypoint extraction + descriptor sampling given the reference's nms / desc maps
    for side in "01":
        kp, sc = O.extract_keypoints(g["nms_" + side], 0.005, 4, cfg["superpoint"]["max_keypoints"])
        assert np.array_equal(kp, g[f"keypoints{side}_0"])
        assert np.array_equal(sc, g[f"scores{side}_0"])
        de = O.sample_descriptors(kp, g["desc_" + side], align_corners=False)
        assert np.abs(de - g[f"descriptors{side}_0"]).max() < 2e-6
    # SuperGlue given the reference's SuperPoint outputs
    r = O.superglue_forward(g["keypoints0_0"], g["scores0_0"], g["descriptors0_0"],
                            g["keypoints1_0"], g["scores1_0"], g["descriptors1_0"],
                            H, W, sg, cfg["superglue"], want=("kenc", "gnn", "S", "Z"))
    assert np.abs(r["kenc0"] - g["kenc0"]).max() < 1e-5
    assert np.abs(r["gnn0"] - g["gnn0"]).max() < 2e-4
    assert np.abs(r["gnn1"] - g["gnn1"]).max() < 2e-4
    assert np.abs(r["S"] - g["S"]).max() < 1e-3
    assert np.abs(r["Z"] - g["Z"]).max() < 1e-3
    # optimal transport + match selection given the reference's S: exact indices
    Z = O.log_optimal_transport(g["S"], sg["bin_score"], cfg["superglue"]["sinkhorn_iterations"])
    assert np.allclose(Z, g["Z"], rtol=0, atol=1e-4)   # summation-order ulps at |Z| ~ 1e2
    m0, m1, s0, s1 = O.match_select(g["Z"], cfg["superglue"]["match_threshold"])
    assert np.array_equal(m0, g["matches0"][0]) and np.array_equal(m1, g["matches1"][0])
    assert np.abs(s0 - g["matching_scores0"][0]).max() < 1e-6
    assert np.abs(s1 - g["matching_scores1"][0]).max() < 1e-6
    # single layer deltas
    d0 = O.attentional_propagation(g["kenc0"], g["kenc0"], sg, "gnn.layers.0")
    assert np.abs(d0 - g["layer0_delta0"]).max() < 2e-5
    d1 = O.attentional_propagation(g["kenc0"], g["kenc1"], sg, "gnn.layers.1")
    assert np.abs(d1 - g["layer1_delta0"]).max() < 2e-5


def _end_to_end(name, seeds, H, W, sp, cfg, min_common=0.98):
    g = load_golden(name)
    sg = synth.superglue_weights(0, 128)
    for i, seed in enumerate(seeds):
        a, b = synth.make_pair(seed, H, W)
        r = O.matching_forward(a, b, sp, sg, cfg)
        for side in "01":
            ref = kp_set(g[f"keypoints{side}_{i}"])
            got = kp_set(r["keypoints" + side])
            assert len(ref & got) >= min_common * len(ref), (len(ref & got), len(ref))
        ref_pairs = match_pairs(g[f"keypoints0_{i}"], g[f"keypoints1_{i}"], g["matches0"][i])
        got_pairs = match_pairs(r["keypoints0"], r["keypoints1"], r["matches0"])
        assert len(ref_pairs & got_pairs) >= 0.9 * len(ref_pairs), (len(ref_pairs & got_pairs), len(ref_pairs))


def test_end_to_end_ragged():
    # H, W not multiples of 8; max_keypoints=-1 -> row-major order branch
    _end_to_end("ragged_hw", [2], 123, 165, synth.superpoint_weights(0, 128),
                golden_cfg(max_kp=-1, iters=20))


def test_end_to_end_c1_real():
    _end_to_end("c1_real", [1], 480, 640, real_superpoint_weights(), golden_cfg(max_kp=1024))


def test_empty_keypoints_branch():
    # superglue_test.py:235-242: int32 -1 matches, zero scores
    sg = synth.superglue_weights(0, 128)
    kp1, sc1, de1 = synth.random_features(0, 1, 7, 128, 120, 160)
    r = O.superglue_forward(np.zeros((0, 2), np.float32), np.zeros(0, np.float32), np.zeros((128, 0), np.float32),
                            kp1[0], sc1[0], de1[0], 120, 160, sg, golden_cfg()["superglue"])
    assert r["matches0"].shape == (0,) and r["matches0"].dtype == np.int32
    assert r["matches1"].shape == (7,) and (r["matches1"] == -1).all()
    assert (r["matching_scores1"] == 0).all()


@pytest.mark.parametrize("name,seed,hw,sp_real,kw", [
    ("small_stages", 1, (120, 160), False, dict(max_kp=256)),
    ("real_small_stages", 4, (160, 224), True, dict(max_kp=300)),
])
def test_torch_cpu_oracle_matches_reference_goldens(name, seed, hw, sp_real, kw):
    """oracle/matching_oracle_torch.py (the CPU arm bench.py times) against the reference-generated goldens:
    identical keypoints and matches, descriptors / scores to fp32 rounding."""
    from oracle import matching_oracle_torch as OT
    g = load_golden(name)
    cfg = golden_cfg(**kw)
    sp = real_superpoint_weights() if sp_real else synth.superpoint_weights(0, 128)
    sg = synth.superglue_weights(0, 128, cfg["superglue"]["keypoint_encoder"])
    a, b = synth.make_pair(seed, *hw)
    r = OT.matching_forward(a, b, sp, sg, cfg)
    for side in "01":
        assert np.array_equal(r["keypoints" + side], g[f"keypoints{side}_0"])
        assert np.abs(r["scores" + side] - g[f"scores{side}_0"]).max() < 1e-6
        assert np.abs(r["descriptors" + side] - g[f"descriptors{side}_0"]).max() < 1e-5
    assert np.array_equal(r["matches0"], g["matches0"][0]) and np.array_equal(r["matches1"], g["matches1"][0])
    assert np.abs(r["matching_scores0"] - g["matching_scores0"][0]).max() < 1e-4


def test_official_variant_oracle_matches_reference_golden():
    """the BatchNorm-free "official" SuperPoint + SuperGlue (superglue/models/matching.py) -- the
    oracle against a golden generated from the reference's own official classes (make_golden.py official)."""
    g = load_golden("official_small")
    cfg = golden_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=300, iters=30)
    sp = synth.superpoint_official_weights(1, 256)
    sg = synth.superglue_weights(1, 256, (32, 64, 128, 256))
    a, b = synth.make_pair(5, 160, 224)
    r = O.matching_forward(a, b, sp, sg, cfg)
    for side in "01":
        ref, got = kp_set(g[f"keypoints{side}_0"]), kp_set(r["keypoints" + side])
        assert len(ref & got) >= 0.98 * len(ref), (len(ref & got), len(ref))
    ref_pairs = match_pairs(g["keypoints0_0"], g["keypoints1_0"], g["matches0"][0])
    got_pairs = match_pairs(r["keypoints0"], r["keypoints1"], r["matches0"])
    assert len(ref_pairs) > 20 and len(ref_pairs & got_pairs) >= 0.9 * len(ref_pairs), (len(ref_pairs & got_pairs), len(ref_pairs))


def test_torch_cpu_oracle_config3_golden():
    """BASELINE config 3 at its stated size (1280x960, the reference's COCO D=256 SuperPoint checkpoint, 2048 keypoints):
    the torch-CPU oracle reproduces the reference-generated golden -- identical keypoints and matches."""
    import os
    from conftest import GOLDEN
    from oracle import matching_oracle_torch as OT
    g = load_golden("c3_real")
    cfg = golden_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=2048, iters=30)
    sp = dict(np.load(os.path.join(GOLDEN, "superpoint_coco256_weights.npz")))
    sg = synth.superglue_weights(1, 256, (32, 64, 128, 256))
    a, b = synth.make_pair(1, 960, 1280)
    r = OT.matching_forward(a, b, sp, sg, cfg)
    for side in "01":
        assert np.array_equal(r["keypoints" + side], g[f"keypoints{side}_0"])
        assert np.abs(r["descriptors" + side][:, ::16] - g[f"descriptors{side}_0"]).max() < 1e-5
    assert np.array_equal(r["matches0"], g["matches0"][0]) and np.array_equal(r["matches1"], g["matches1"][0])


def test_torch_cpu_oracle_d64_golden():
    """The third SuperPoint width the reference ships (D = 64, its own allss_descriptor_64 checkpoint, 640x480, 1024
    keypoints): the torch-CPU oracle reproduces the reference-generated golden -- identical keypoints and matches."""
    import os
    from conftest import GOLDEN
    from oracle import matching_oracle_torch as OT
    g = load_golden("d64_real")
    cfg = golden_cfg(D=64, kenc=(32, 64), max_kp=1024, iters=30)
    sp = dict(np.load(os.path.join(GOLDEN, "superpoint_allss64_weights.npz")))
    sg = synth.superglue_weights(2, 64, (32, 64))
    a, b = synth.make_pair(1, 480, 640)
    r = OT.matching_forward(a, b, sp, sg, cfg)
    for side in "01":
        assert np.array_equal(r["keypoints" + side], g[f"keypoints{side}_0"])
        assert np.abs(r["descriptors" + side] - g[f"descriptors{side}_0"]).max() < 1e-5
    assert np.array_equal(r["matches0"], g["matches0"][0]) and np.array_equal(r["matches1"], g["matches1"][0])


def test_reference_copy_reproduces_goldens_and_port():
    """oracle/_ref (the reference's own modules, copied by oracle/make_ref.py -- the CPU baseline bench.py times)
    reproduces the committed golden bit for bit, and the torch-CPU port agrees with it on a fresh seed."""
    import torch
    from oracle import make_ref, matching_oracle_torch as OT
    if not make_ref.available() and make_ref.make(verbose=False) is None:
        pytest.skip("no reference checkout and no oracle/_ref copy")
    g = load_golden("real_small_stages")
    cfg = golden_cfg(max_kp=300)
    sp, sg = real_superpoint_weights(), synth.superglue_weights(0, 128)
    m = make_ref.load_matching(cfg, sp, sg)
    a, b = synth.make_pair_batch([4], 160, 224)
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    assert np.array_equal(pred["keypoints0"][0].numpy(), g["keypoints0_0"])
    assert np.array_equal(pred["matches0"].numpy(), g["matches0"])
    assert np.array_equal(pred["descriptors1"][0].numpy(), g["descriptors1_0"])
    a, b = synth.make_pair_batch([77], 160, 224)
    pred = m({"image0": torch.from_numpy(a), "image1": torch.from_numpy(b)})
    r = OT.matching_forward(a[0, 0], b[0, 0], sp, sg, cfg)
    assert np.array_equal(pred["keypoints1"][0].numpy(), r["keypoints1"])
    assert np.array_equal(pred["matches0"][0].numpy(), r["matches0"])
