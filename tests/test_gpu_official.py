"""GPU: the "official" variant -- MatchingOfficial / SuperPointOfficial (no BatchNorm, MagicLeap key
names; reference superglue/models/matching.py:46-82, superglue/models/superpoint.py:95-202) against the golden
generated from the reference's own classes."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_cfg, kp_set, match_pairs

pytestmark = pytest.mark.gpu


def _model():
    from image_matching_b200 import MatchingOfficial, synth
    cfg = golden_cfg(D=256, kenc=(32, 64, 128, 256), max_kp=300, iters=30)
    m = MatchingOfficial({"superpoint": dict(cfg["superpoint"]), "superglue": dict(cfg["superglue"], weights="")}).eval()
    sp = synth.superpoint_official_weights(1, 256)
    sg = synth.superglue_weights(1, 256, (32, 64, 128, 256))
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})   # reference key names
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    return m.to("cuda:0")


def test_official_matching_against_reference_golden():
    from image_matching_b200 import synth
    g = load_golden("official_small")
    m = _model()
    assert sorted(m.superpoint.state_dict()) == sorted(
        f"{n}.{p}" for n in ("conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
                             "convPa", "convPb", "convDa", "convDb") for p in ("weight", "bias"))
    a, b = synth.make_pair(5, 160, 224)
    pred = m({"image0": torch.from_numpy(a[None, None]).cuda(), "image1": torch.from_numpy(b[None, None]).cuda()})
    for side in "01":
        kp = pred["keypoints" + side][0].cpu().numpy()
        assert np.array_equal(kp, g[f"keypoints{side}_0"])                     # same keypoints, same order
        ds = float(np.abs(pred["scores" + side][0].cpu().numpy() - g[f"scores{side}_0"]).max())
        print(f"official side {side}: max |score - ref| = {ds:.2e}")
        assert ds < 1e-4        # detector probabilities after 12 BatchNorm-free fp16x3 conv layers (measured 1.3e-5)
        assert np.abs(pred["descriptors" + side][0].cpu().numpy() - g[f"descriptors{side}_0"]).max() < 1e-3
    assert pred["matches0"].dtype == torch.int64
    assert np.array_equal(pred["matches0"][0].cpu().numpy(), g["matches0"][0])
    assert np.array_equal(pred["matches1"][0].cpu().numpy(), g["matches1"][0])
    assert np.abs(pred["matching_scores0"][0].cpu().numpy() - g["matching_scores0"][0]).max() < 1e-3


def test_official_superpoint_alone_and_validation():
    from image_matching_b200 import SuperPointOfficial, synth
    with pytest.raises(ValueError):
        SuperPointOfficial({"max_keypoints": 0})                               # superpoint.py:143-145
    sp = SuperPointOfficial({"descriptor_dim": 256, "max_keypoints": 300}).eval()
    sp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.superpoint_official_weights(1, 256).items()})
    sp = sp.to("cuda:0")
    g = load_golden("official_small")
    a, _ = synth.make_pair(5, 160, 224)
    out = sp({"image": torch.from_numpy(a[None, None]).cuda()})                # reference call signature (:149)
    assert kp_set(out["keypoints"][0].cpu().numpy()) == kp_set(g["keypoints0_0"])
