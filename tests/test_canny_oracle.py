"""CPU: the Canny oracle (numpy + plain C) against the committed cv2 golden vectors and live cv2."""
import zlib

import numpy as np
import pytest

from oracle import c_oracle
from oracle.canny_oracle import canny_gray, preprocess_image, rgb_to_gray, synthetic_image
from tests.util import canny_golden_cases


def test_numpy_oracle_matches_golden():
    n = 0
    for img, lo, hi, edges, gray_crc in canny_golden_cases():
        gray = rgb_to_gray(img)
        assert zlib.crc32(gray.tobytes()) == gray_crc
        assert np.array_equal(canny_gray(gray, lo, hi), edges)
        if img.shape[0] * img.shape[1] <= 256 * 256:   # the pure-Python flood fill only on small cases
            assert np.array_equal(canny_gray(gray, lo, hi, fast=False), edges)
        n += 1
    assert n >= 10


def test_c_oracle_matches_golden():
    for img, lo, hi, edges, _ in canny_golden_cases():
        out = c_oracle.canny_u8(img[None], lo, hi)
        assert np.array_equal(out[0], edges)
        out3 = c_oracle.canny_u8(img[None], lo, hi, replicate3=True)
        assert np.array_equal(out3[0], np.stack([edges] * 3, axis=2))


def test_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for seed, kind, (h, w) in [(11, "shapes", (200, 300)), (12, "noise", (77, 131)), (13, "smooth", (128, 128))]:
        img = synthetic_image(seed, h, w, kind)
        gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
        assert np.array_equal(gray, rgb_to_gray(img))
        for lo, hi in [(100, 200), (10, 20), (300, 50)]:
            ref = cv2.Canny(gray, lo, hi)
            assert np.array_equal(canny_gray(gray, lo, hi), ref)
            assert np.array_equal(c_oracle.canny_u8(gray[None], lo, hi)[0], ref)
        assert np.array_equal(preprocess_image(img)[..., 1], cv2.Canny(gray, 100, 200))


def test_gray_formula_exhaustive_sample():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(512, 512, 3), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), rgb_to_gray(img))


def test_empty_and_flat_inputs():
    flat = np.full((1, 40, 50, 3), 77, np.uint8)
    assert c_oracle.canny_u8(flat).sum() == 0
    assert c_oracle.canny_u8(np.zeros((0, 8, 8, 3), np.uint8)).shape == (0, 8, 8)


def test_gaussian_oracle_matches_golden_and_live_cv2():
    """The optional Gaussian pre-stage: plain-C restatement of cv2.GaussianBlur(img, (5, 5), 0) against the committed cv2 CRCs
    (incl. 1-, 2- and 3-pixel-wide images where BORDER_REFLECT_101 wraps more than once) and against cv2 itself."""
    import zlib
    from tests.util import gauss_golden_cases
    n = 0
    for src, blur_crc, edges_crc in gauss_golden_cases():
        blur = c_oracle.gaussian_blur5_u8(src[None])[0]
        assert zlib.crc32(blur.tobytes()) == blur_crc, src.shape
        if edges_crc is not None:
            assert zlib.crc32(c_oracle.canny_u8(blur[None], 100, 200)[0].tobytes()) == edges_crc
        n += 1
    assert n >= 10
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for shape in [(64, 64), (37, 53, 3), (4, 6), (300, 200, 3)]:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(c_oracle.gaussian_blur5_u8(a[None])[0], cv2.GaussianBlur(a, (5, 5), 0))
