"""CPU: the Lanczos resize oracle (oracle/resize_oracle.py) is pinned bit-exactly against Pillow itself — the implementation the
reference calls at ``src/pipeline.py:251`` — and the product's host-side weight tables equal the oracle's."""
import numpy as np
import pytest
from PIL import Image

from oracle import resize_oracle as R

CASES = [(512, 512, 1024, 1024), (480, 640, 1024, 1024), (1500, 1100, 1024, 1024), (1024, 512, 1024, 1024), (333, 777, 256, 300),
         (1, 17, 1024, 1024), (33, 1, 64, 64), (1024, 1024, 1024, 1024), (2048, 2048, 1024, 1024)]


@pytest.mark.parametrize("h,w,oh,ow", CASES)
def test_oracle_matches_pillow_bit_exactly(h, w, oh, ow):
    rng = np.random.default_rng(h * 7 + w)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(a).resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(R.resize_lanczos_u8(a, oh, ow), ref)


def test_smooth_and_extreme_images():
    yy, xx = np.mgrid[0:512, 0:512]
    grad = np.stack([(xx // 2) % 256, (yy // 2) % 256, ((xx + yy) // 4) % 256], -1).astype(np.uint8)
    checker = (((xx // 3 + yy // 5) % 2) * 255).astype(np.uint8)[..., None].repeat(3, -1)     # ringing: exercises the clip to [0, 255]
    for a in (grad, checker, np.zeros((512, 512, 3), np.uint8), np.full((512, 512, 3), 255, np.uint8)):
        assert np.array_equal(R.resize_lanczos_u8(a, 1024, 1024), np.array(Image.fromarray(a).resize((1024, 1024), Image.LANCZOS)))


@pytest.mark.parametrize("n_in,n_out", [(512, 1024), (640, 1024), (1500, 1024), (777, 300), (17, 1024)])
def test_product_tables_equal_oracle_tables(n_in, n_out):
    from fast_image_editing_with_generative_models_b200.resize import lanczos_tables
    b, k, ks = lanczos_tables(n_in, n_out)
    bo, ko, kso = R.lanczos_coeffs(n_in, n_out)
    assert ks == kso and np.array_equal(b.numpy(), bo) and np.array_equal(k.numpy(), ko)
    assert int(k.sum(1).min()) >= (1 << 22) - 8 and int(k.sum(1).max()) <= (1 << 22) + 8      # weights sum to ~1.0 in 2^22 fixed point
