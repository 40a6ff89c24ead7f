"""Generates tests/golden/canny_golden.npz from the REAL cv2 (the reference's Canny, src/pipeline.py:200,205).

Run in the build container (cv2 4.13.0):  python tests/golden/make_canny_golden.py
Inputs are regenerated from seeds by oracle.canny_oracle.synthetic_image; only the (bit-packed) cv2 outputs
and a CRC of each input are stored, so the fixture stays small.
"""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.canny_oracle import synthetic_image  # noqa: E402

CASES = [  # (seed, h, w, kind, low, high)
    (0, 256, 256, "shapes", 100, 200), (1, 256, 256, "noise", 100, 200), (2, 256, 256, "smooth", 100, 200),
    (3, 129, 203, "shapes", 50, 150), (4, 97, 64, "noise", 200, 100), (5, 64, 320, "smooth", 30, 90),
    (6, 1, 17, "noise", 100, 200), (7, 33, 1, "noise", 100, 200), (8, 512, 512, "shapes", 100, 200),
    (0, 1024, 1024, "shapes", 100, 200), (1, 1024, 1024, "noise", 100, 200),
]

if __name__ == "__main__":
    out = {}
    for i, (seed, h, w, kind, lo, hi) in enumerate(CASES):
        img = synthetic_image(seed, h, w, kind)
        gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) if min(h, w) > 1 else cv2.cvtColor(img, cv2.COLOR_RGB2GRAY).reshape(h, w)
        edges = cv2.Canny(gray, lo, hi)
        out[f"case{i}_meta"] = np.array([seed, h, w, ["shapes", "noise", "smooth"].index(kind), lo, hi, zlib.crc32(img.tobytes())], np.int64)
        out[f"case{i}_gray_crc"] = np.array([zlib.crc32(gray.tobytes())], np.int64)
        out[f"case{i}_edges"] = np.packbits(edges > 0)
        print(i, kind, h, w, float((edges > 0).mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "canny_golden.npz"), **out)
