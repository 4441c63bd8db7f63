"""GPU: 8-bit image entry points.  The reference's data loader turns the decoded image into
`img / 255.` (float64, datasets/SSHIDataset.py:26-29) and the caller casts to float32 (superpoint_glue_test.py:74-75);
the uint8 entry points do that normalisation on the device, so uploading the raw pixels must give bit-identical results."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg, real_superpoint_weights

pytestmark = pytest.mark.gpu


def test_normalisation_is_the_reference_rounding():
    v = np.arange(256, dtype=np.uint8)
    ref = (v.astype(np.float64) / 255.0).astype(np.float32)          # SSHIDataset.py:28 + .float()
    assert np.array_equal(ref, v.astype(np.float32) / np.float32(255.0))   # == one correctly rounded fp32 division


def test_uint8_images_match_float_images_bit_for_bit():
    from image_matching_b200 import Matching, synth
    cfg = golden_cfg(max_kp=256)      # every image yields more candidates than that: equal counts, batch stacks
    m = Matching({"superpoint": dict(cfg["superpoint"], weights=None), "superglue": dict(cfg["superglue"], weights="")}).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in real_superpoint_weights().items()})
    sg = synth.superglue_weights(0, 128)
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    m = m.to("cuda:0")
    a, b = synth.make_pair_batch([1, 2, 3], 240, 320)
    a8, b8 = np.round(a * 255).astype(np.uint8), np.round(b * 255).astype(np.uint8)
    af = torch.from_numpy(a8.astype(np.float64) / 255.0).float()     # what the reference's loader + caller produce
    bf = torch.from_numpy(b8.astype(np.float64) / 255.0).float()
    pf = m({"image0": af.cuda(), "image1": bf.cuda()})
    p8 = m({"image0": torch.from_numpy(a8).cuda(), "image1": torch.from_numpy(b8).cuda()})
    for k in ("matches0", "matches1", "matching_scores0", "matching_scores1"):
        assert torch.equal(pf[k], p8[k]), k
    for k in ("keypoints0", "scores0", "descriptors0", "keypoints1", "scores1", "descriptors1"):
        assert all(torch.equal(x, y) for x, y in zip(pf[k], p8[k])), k
    assert int((p8["matches0"] > -1).sum()) > 50
    # SuperPoint alone
    s8 = m.superpoint(torch.from_numpy(a8).cuda())
    sf = m.superpoint(af.cuda())
    assert all(torch.equal(x, y) for x, y in zip(s8["keypoints"], sf["keypoints"]))


def test_resize_matches_cv2_golden_and_oracle():
    """b200m_resize_linear_u8 (datasets/SSHIDataset.py:20-22 on the device): bit-identical to the cv2-generated golden
    (tests/golden/resize.npz) and to the oracle at the loader's real sizes, batched, incl. the exact-2x box path."""
    from conftest import load_golden
    from image_matching_b200 import SuperPoint, resize_u8
    from oracle import input_oracle as IO
    sp = SuperPoint({"weights": None, "descriptor_dim": 128}).eval().to("cuda:0")
    g = load_golden("resize")
    for i, (h, w, sc) in enumerate(g["cases"]):
        got = resize_u8(sp, torch.from_numpy(g[f"src_{i}"]).cuda(), float(sc)).cpu().numpy()
        assert np.array_equal(got, g[f"dst_{i}"]), (i, h, w, sc)
    rng = np.random.default_rng(9)
    for H, W, sc in [(1944, 2592, 0.125), (960, 1280, 0.5), (1000, 1504, 0.32)]:
        src = rng.integers(0, 256, (3, 1, H, W), dtype=np.uint8)
        got = resize_u8(sp, torch.from_numpy(src).cuda(), sc).cpu().numpy()
        dw, dh = int(sc * W), int(sc * H)
        assert got.shape == (3, 1, dh, dw)
        for b in range(3):
            assert np.array_equal(got[b, 0], IO.resize_linear_u8(src[b, 0], dw, dh)), (H, W, sc, b)


def test_loader_pipeline_on_device_equals_reference_loader():
    """resize + /255 on the device (uint8 in) == the reference's loader on the host (oracle restatement) through SuperPoint."""
    from image_matching_b200 import SuperPoint, resize_u8, synth
    from oracle import input_oracle as IO
    sp = SuperPoint({"weights": None, "descriptor_dim": 128, "max_keypoints": 200}).eval()
    sp.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in real_superpoint_weights().items()})
    sp = sp.to("cuda:0")
    big = np.round(synth.make_image(3, 960, 1280) * 255).astype(np.uint8)
    host = IO.load_like_sshi(big, 0.25)                                   # (1, 240, 320) float32, the reference's loader
    dev8 = resize_u8(sp, torch.from_numpy(big).cuda()[None, None], 0.25)  # (1, 1, 240, 320) uint8 on the device
    a = sp(torch.from_numpy(host)[None].cuda())
    b = sp(dev8)
    assert torch.equal(a["keypoints"][0], b["keypoints"][0]) and torch.equal(a["descriptors"][0], b["descriptors"][0])
    assert a["keypoints"][0].shape[0] == 200
