"""GPU: fie_jpeg_encode_u8 (through the C-ABI) must write the SAME BYTES as Pillow's ``Image.save(f, "JPEG", quality=q)`` — the call
the reference's CLIs make on every edited image (``run_batch.py:224``, ``run_single_image.py:114``) — and as the numpy oracle."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import jpeg_oracle as J
from oracle.canny_oracle import synthetic_image

pytestmark = pytest.mark.gpu


def pil_bytes(a, q=75):
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", quality=q)
    return b.getvalue()


@pytest.mark.parametrize("seed,h,w,kind,q", [(0, 64, 64, "shapes", 75), (1, 37, 53, "smooth", 75), (2, 100, 200, "shapes", 90), (3, 9, 9, "noise", 75),
                                            (4, 1, 1, "noise", 75), (5, 17, 40, "noise", 60), (6, 40, 17, "shapes", 85), (7, 130, 70, "noise", 75),
                                            (8, 128, 96, "noise", 95), (9, 256, 256, "shapes", 50), (10, 512, 512, "noise", 75), (11, 250, 250, "smooth", 30)])
def test_gpu_jpeg_is_byte_identical_to_pillow(cuda_dev, seed, h, w, kind, q):
    from fast_image_editing_with_generative_models_b200 import ops
    a = synthetic_image(seed, h, w, kind)
    got = ops.jpeg_bytes(torch.from_numpy(a[None]).to(cuda_dev), q)[0]
    ref = pil_bytes(a, q)
    assert got == J.encode(a, q), "GPU encoder differs from the oracle"
    assert got == ref, "GPU encoder differs from Pillow"


def test_gpu_jpeg_full_size_batch(cuda_dev):
    """The path's shape: a batch of 1024 x 1024 outputs, default quality (what `edited.save("x.jpg")` uses); every file must decode
    (Pillow) to exactly what Pillow's own file decodes to — trivially, because the bytes are equal."""
    from fast_image_editing_with_generative_models_b200 import ops
    imgs = np.stack([synthetic_image(s, 1024, 1024, k) for s, k in [(0, "shapes"), (1, "noise"), (2, "smooth"), (3, "shapes")]])
    imgs[3] = 255                                                            # saturated image: long runs of identical blocks
    files = ops.jpeg_bytes(torch.from_numpy(imgs).to(cuda_dev), 75)
    for a, f in zip(imgs, files):
        assert f == pil_bytes(a, 75)
        assert np.array_equal(np.array(Image.open(io.BytesIO(f))), np.array(Image.open(io.BytesIO(pil_bytes(a, 75)))))
    assert ops.jpeg_bytes(torch.zeros((0, 16, 16, 3), dtype=torch.uint8, device=cuda_dev)) == []
    # repeated calls reuse the cached workspace: the zeroing of the bit stream must happen every call
    assert ops.jpeg_bytes(torch.from_numpy(imgs[:1]).to(cuda_dev), 75)[0] == files[0]
    assert ops.jpeg_bytes(torch.from_numpy(imgs[:1]).to(cuda_dev), 75)[0] == files[0]


def test_editor_returns_jpeg_files(cuda_dev, tmp_path):
    """FastEditor.edit_many(output="jpeg"): the edited images leave the GPU as JPEG files, equal to saving the PIL results."""
    from src.pipeline import FastEditor
    ed = FastEditor(model_name="ssd-1b", device="cuda", tiny=True, verbose=False)
    imgs = [Image.fromarray(synthetic_image(20 + i, 1024, 1024)) for i in range(3)]
    pils = ed.edit_many(imgs, "a rusty bicycle", seed=4, micro_batch=2)
    jpgs = ed.edit_many(imgs, "a rusty bicycle", seed=4, micro_batch=2, output="jpeg")
    assert all(isinstance(j, bytes) for j in jpgs)
    for p, j in zip(pils, jpgs):
        b = io.BytesIO()
        p.save(b, "JPEG")
        assert b.getvalue() == j
