"""Golden vectors for the input side: cv2.resize (the call of datasets/SSHIDataset.py:20-22) on seeded
random 8-bit images, generated with the cv2 of the build container (4.13.0).

    python tests/golden/make_golden_resize.py        # writes tests/golden/resize.npz
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(96, 128, 0.5), (96, 128, 0.25), (97, 131, 0.3), (77, 103, 0.125), (48, 64, 1.5), (60, 80, 0.75), (50, 70, 1.0),
         (120, 90, 0.4), (33, 47, 2.0)]


def main():
    rng = np.random.default_rng(123)
    out = {"cases": np.asarray(CASES, np.float64), "cv2_version": np.asarray(cv2.__version__)}
    for i, (h, w, sc) in enumerate(CASES):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        # a smooth image as well: random noise alone would hide rounding differences on gradients
        if i % 2:
            yy, xx = np.mgrid[0:h, 0:w]
            src = ((np.sin(xx / 7.0) * np.cos(yy / 5.0) * 0.5 + 0.5) * 255).astype(np.uint8)
        out[f"src_{i}"] = src
        out[f"dst_{i}"] = cv2.resize(src, (int(sc * w), int(sc * h)))      # exactly the reference's call
    path = os.path.join(HERE, "resize.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
