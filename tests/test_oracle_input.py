"""CPU: the input-side oracle (cv2.resize restatement, SSHIDataset normalisation) against cv2-generated goldens."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import input_oracle as IO


def test_resize_oracle_matches_cv2_golden():
    g = load_golden("resize")
    for i, (h, w, sc) in enumerate(g["cases"]):
        src, ref = g[f"src_{i}"], g[f"dst_{i}"]
        got = IO.resize_linear_u8(src, int(sc * w), int(sc * h))
        assert got.shape == ref.shape and np.array_equal(got, ref), (i, h, w, sc, int((got != ref).sum()))


def test_resize_oracle_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for h, w, sc in [(480, 640, 0.125), (777, 1033, 0.3), (960, 1280, 0.5), (200, 300, 1.25), (64, 64, 3.0), (1944, 2592, 0.2)]:
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = int(sc * w), int(sc * h)
        assert np.array_equal(IO.resize_linear_u8(src, dw, dh), cv2.resize(src, (dw, dh))), (h, w, sc)


def test_load_like_sshi_normalisation():
    img = np.arange(256, dtype=np.uint8).reshape(16, 16)
    x = IO.load_like_sshi(img)
    assert x.dtype == np.float32 and x.shape == (1, 16, 16)
    assert np.array_equal(x[0], (img / 255).astype(np.float32))
    # float32 division gives the same bits as the loader's float64 division followed by .float() (what the device does)
    assert np.array_equal(x[0], img.astype(np.float32) / np.float32(255))
