"""Generates tests/golden/registration.npz with cv2 (opencv-python 4.13.0 in the build container): the vectors that pin
oracle/registration_oracle.py and the CUDA registration kernels to what the reference's caller gets from OpenCV at
superpoint_glue_test.py:88,101.  Run from the repo root: python tests/golden/make_golden_registration.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def problem(rs, n, noise, outliers):
    fr = np.stack([rs.integers(4, 636, n), rs.integers(4, 476, n)], 1).astype(np.float32)
    ang, s = rs.uniform(-0.5, 0.5), rs.uniform(0.7, 1.3)
    A = np.array([[s * np.cos(ang), -s * np.sin(ang), rs.uniform(-40, 40)],
                  [s * np.sin(ang), s * np.cos(ang), rs.uniform(-40, 40)]])
    to = fr @ A[:, :2].T + A[:, 2] + rs.normal(0, noise, (n, 2))
    o = rs.random(n) < outliers
    to[o] = np.stack([rs.uniform(0, 640, o.sum()), rs.uniform(0, 480, o.sum())], 1)
    return fr, np.round(to).astype(np.float32)


def main():
    rs = np.random.default_rng(2024)
    out = {"cv2_version": np.array(cv2.__version__)}
    cases = [(4, 0.0, 0.0), (5, 1.0, 0.2), (17, 2.0, 0.5), (64, 0.5, 0.0), (200, 3.0, 0.7), (413, 1.5, 0.3),
             (1024, 2.0, 0.6), (1024, 0.0, 0.0), (700, 4.0, 0.85), (3, 0.0, 0.0), (300, 1.0, 0.95)]
    for i, (n, noise, outl) in enumerate(cases):
        fr, to = problem(rs, n, noise, outl)
        M, mask = cv2.estimateAffinePartial2D(fr, to, method=cv2.RANSAC, ransacReprojThreshold=7)
        out[f"ransac{i}_from"], out[f"ransac{i}_to"] = fr, to
        out[f"ransac{i}_M"] = np.zeros((0,)) if M is None else M
        out[f"ransac{i}_mask"] = mask
    out["n_ransac"] = np.array(len(cases))
    k = 0
    for dt in (np.float64, np.float32, np.uint8):
        for (H, W), (dw, dh) in (((96, 128), (128, 96)), ((61, 83), (100, 70))):
            src = (rs.random((H, W)) * 255).astype(dt)
            ang, s = rs.uniform(-0.4, 0.4), rs.uniform(0.7, 1.3)
            M = np.array([[s * np.cos(ang), -s * np.sin(ang), rs.uniform(-20, 20)],
                          [s * np.sin(ang), s * np.cos(ang), rs.uniform(-20, 20)]])
            out[f"warp{k}_src"], out[f"warp{k}_M"] = src, M
            out[f"warp{k}_dst"] = cv2.warpAffine(src, M, (dw, dh))
            k += 1
    # integer translation / identity: every pixel sits on the interpolation grid origin
    src = (rs.random((40, 50)) * 255).astype(np.uint8)
    M = np.array([[1.0, 0.0, 3.0], [0.0, 1.0, -2.0]])
    out[f"warp{k}_src"], out[f"warp{k}_M"], out[f"warp{k}_dst"] = src, M, cv2.warpAffine(src, M, (50, 40))
    k += 1
    out["n_warp"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "registration.npz"), **out)
    print("wrote registration.npz:", len(cases), "RANSAC problems,", k, "warps, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
