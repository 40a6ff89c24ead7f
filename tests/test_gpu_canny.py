"""GPU: fie_canny_u8 (through the C-ABI) must be bit-exact with the oracle / cv2 golden vectors."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle.canny_oracle import synthetic_image
from tests.util import canny_golden_cases

pytestmark = pytest.mark.gpu


def test_canny_matches_golden(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    for img, lo, hi, edges, _ in canny_golden_cases():
        out = ops.canny(torch.from_numpy(img[None]).to(cuda_dev), lo, hi).cpu().numpy()[0]
        assert np.array_equal(out, edges), f"mismatch {img.shape} lo={lo} hi={hi}: {(out != edges).sum()} px"
        out3 = ops.canny(torch.from_numpy(img[None]).to(cuda_dev), lo, hi, out_channels=3).cpu().numpy()[0]
        assert np.array_equal(out3, np.stack([edges] * 3, axis=2))


def test_canny_batch_and_gray_input(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops
    from oracle.canny_oracle import rgb_to_gray
    imgs = np.stack([synthetic_image(s, 300, 412, k) for s, k in [(1, "shapes"), (2, "noise"), (3, "smooth"), (4, "shapes")]])
    ref = c_oracle.canny_u8(imgs, 100, 200)
    out = ops.canny(torch.from_numpy(imgs).to(cuda_dev), 100, 200).cpu().numpy()
    assert np.array_equal(out, ref)
    gray = np.stack([rgb_to_gray(i) for i in imgs])
    out_g = ops.canny(torch.from_numpy(gray).to(cuda_dev), 100, 200).cpu().numpy()
    assert np.array_equal(out_g, ref)


def test_canny_full_size_batch(cuda_dev):
    """BASELINE config sizes: 1024x1024 batch; oracle = plain-C restatement; plus idempotence-style property."""
    from fast_image_editing_with_generative_models_b200 import ops
    imgs = np.stack([synthetic_image(s, 1024, 1024, "shapes" if s % 2 == 0 else "noise") for s in range(4)])
    ref = c_oracle.canny_u8(imgs, 100, 200)
    d = torch.from_numpy(imgs).to(cuda_dev)
    out = ops.canny(d, 100, 200)
    assert np.array_equal(out.cpu().numpy(), ref)
    # determinism across launches (lock-free union-find must give the unique closure every time)
    for _ in range(3):
        assert torch.equal(ops.canny(d, 100, 200), out)
    # monotonicity property: raising the low threshold can only remove edges
    hi_low = ops.canny(d, 150, 200)
    assert bool(((hi_low > 0) <= (out > 0)).all())


def test_canny_edge_cases(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import ops, _lib
    flat = torch.full((2, 40, 50, 3), 77, dtype=torch.uint8, device=cuda_dev)
    assert int(ops.canny(flat).sum()) == 0
    empty = torch.zeros((0, 8, 8, 3), dtype=torch.uint8, device=cuda_dev)
    assert ops.canny(empty).shape == (0, 8, 8)
    with pytest.raises(_lib.FieError):
        ops.canny(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))   # CPU tensor: no fallback


def test_gaussian_prestage_matches_cv2_golden(cuda_dev):
    """Optional (default-off) integer Gaussian pre-stage: fie_gaussian_blur5_u8 == cv2.GaussianBlur(img, (5, 5), 0) bit for bit
    (committed cv2 CRCs + the plain-C oracle), and ops.canny(gaussian_blur=True) == cv2.Canny of the blurred gray image."""
    import zlib
    from fast_image_editing_with_generative_models_b200 import ops
    from oracle.canny_oracle import rgb_to_gray
    from tests.util import gauss_golden_cases
    for src, blur_crc, edges_crc in gauss_golden_cases():
        d = torch.from_numpy(src[None]).to(cuda_dev)
        blur = ops.gaussian_blur5(d).cpu().numpy()[0]
        assert zlib.crc32(blur.tobytes()) == blur_crc, src.shape
        if edges_crc is not None:
            assert zlib.crc32(ops.canny(d, 100, 200, gaussian_blur=True).cpu().numpy()[0].tobytes()) == edges_crc
    imgs = np.stack([synthetic_image(s, 1024, 1024, "shapes" if s % 2 == 0 else "noise") for s in range(3)])
    d = torch.from_numpy(imgs).to(cuda_dev)
    assert np.array_equal(ops.gaussian_blur5(d).cpu().numpy(), c_oracle.gaussian_blur5_u8(imgs))
    gray = np.stack([rgb_to_gray(i) for i in imgs])
    assert np.array_equal(ops.rgb_to_gray(d).cpu().numpy(), gray)
    ref = c_oracle.canny_u8(c_oracle.gaussian_blur5_u8(gray), 100, 200)
    assert np.array_equal(ops.canny(d, 100, 200, gaussian_blur=True).cpu().numpy(), ref)
    assert not np.array_equal(ops.canny(d, 100, 200).cpu().numpy(), ref)                       # the default path has no blur


@pytest.mark.parametrize("h,w", [(32, 64), (64, 64), (32, 128), (96, 192), (256, 64), (512, 512)])
def test_canny_tile_aligned_small_shapes(cuda_dev, h, w):
    """Shapes that take the four-pixel tile kernel (W % 64 == 0, H % 32 == 0) with few tiles, so that one CTA touches several image borders
    at once; dense (noise) and sparse (shapes) candidate maps, several threshold pairs incl. low = 0, 1- and 3-channel outputs."""
    from fast_image_editing_with_generative_models_b200 import ops
    imgs = np.stack([synthetic_image(s + h + w, h, w, kind) for s, kind in enumerate(("shapes", "noise", "smooth", "noise"))])
    d = torch.from_numpy(imgs).to(cuda_dev)
    for lo, hi in ((100, 200), (0, 40), (50, 60), (200, 100), (300, 900)):
        ref = c_oracle.canny_u8(imgs, lo, hi)
        out = ops.canny(d, lo, hi)
        assert np.array_equal(out.cpu().numpy(), ref), (h, w, lo, hi, int((out.cpu().numpy() != ref).sum()))
    rgb = ops.canny(d, 100, 200, out_channels=3).cpu().numpy()
    assert np.array_equal(rgb, np.repeat(c_oracle.canny_u8(imgs, 100, 200)[..., None], 3, axis=-1))
