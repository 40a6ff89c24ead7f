"""Generates tests/golden/gauss_golden.npz from the REAL cv2: cv2.GaussianBlur(img, (5, 5), 0) on uint8 and the blurred Canny
(gray -> blur -> cv2.Canny) — the optional, default-off pre-stage of fie_canny_u8 (include/fie_b200.h).

Run in the build container (cv2 4.13.0):  python tests/golden/make_gauss_golden.py
Inputs are regenerated from seeds; only CRC32s of the cv2 outputs (and bit-packed edges for one case) are stored."""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.canny_oracle import synthetic_image  # noqa: E402

CASES = [  # (seed, h, w, kind, gray)
    (0, 256, 256, "shapes", 0), (1, 200, 333, "noise", 0), (2, 97, 64, "smooth", 1), (3, 5, 7, "noise", 1), (4, 3, 9, "noise", 0),
    (5, 2, 2, "noise", 1), (6, 1, 17, "noise", 1), (7, 33, 1, "noise", 0), (0, 1024, 1024, "shapes", 0), (1, 1024, 1024, "noise", 1),
]

if __name__ == "__main__":
    out = {}
    for i, (seed, h, w, kind, gray) in enumerate(CASES):
        img = synthetic_image(seed, h, w, kind)
        src = np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY).reshape(h, w)) if gray else img
        blur = cv2.GaussianBlur(src, (5, 5), 0).reshape(src.shape)
        out[f"case{i}_meta"] = np.array([seed, h, w, ["shapes", "noise", "smooth"].index(kind), gray, zlib.crc32(img.tobytes()), zlib.crc32(blur.tobytes())], np.int64)
        if gray and min(h, w) > 2:
            out[f"case{i}_edges_crc"] = np.array([zlib.crc32(cv2.Canny(blur, 100, 200).tobytes())], np.int64)
        print(i, kind, src.shape)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gauss_golden.npz"), **out)
