"""CPU: the JPEG oracle (numpy restatement of libjpeg's baseline encoder) must produce the SAME BYTES as Pillow's
``Image.save(f, "JPEG", quality=q)`` — the call the reference makes at ``run_batch.py:224`` / ``run_single_image.py:114``."""
import io

import numpy as np
import pytest
from PIL import Image

from oracle import jpeg_oracle as J
from oracle.canny_oracle import synthetic_image


def pil_bytes(a, q):
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", quality=q)
    return b.getvalue()


CASES = [(0, 64, 64, "shapes", 75), (1, 37, 53, "smooth", 75), (2, 100, 200, "shapes", 90), (3, 9, 9, "noise", 75), (4, 1, 1, "noise", 75),
         (5, 17, 40, "noise", 60), (6, 40, 17, "shapes", 85), (7, 130, 70, "noise", 75), (8, 128, 96, "noise", 95), (9, 256, 256, "shapes", 50),
         (10, 512, 512, "noise", 75), (11, 250, 250, "smooth", 30)]


@pytest.mark.parametrize("seed,h,w,kind,q", CASES)
def test_jpeg_oracle_is_byte_identical_to_pillow(seed, h, w, kind, q):
    a = synthetic_image(seed, h, w, kind)
    assert J.encode(a, q) == pil_bytes(a, q)


def test_jpeg_oracle_extremes_and_default_quality():
    for a in (np.zeros((48, 48, 3), np.uint8), np.full((64, 80, 3), 255, np.uint8), np.random.default_rng(0).integers(0, 256, (96, 64, 3), dtype=np.uint8)):
        b = io.BytesIO()
        Image.fromarray(a).save(b, "JPEG")                                   # PIL's default options = quality 75, 4:2:0
        assert J.encode(a) == b.getvalue()


def test_jpeg_oracle_full_size():
    a = synthetic_image(0, 1024, 1024, "shapes")
    assert J.encode(a, 75) == pil_bytes(a, 75)


def test_c_abi_header_writer_matches_pillow_without_a_gpu():
    """fie_jpeg_write_header is host code of libfie_b200.so: the 623 bytes before the scan must equal Pillow's (SOI, JFIF APP0, DQT x 2,
    SOF0 4:2:0, DHT x 4, SOS) for any size and quality — checked here on the CPU box through the C-ABI."""
    import ctypes
    from fast_image_editing_with_generative_models_b200 import _lib
    L = _lib.lib()
    n = L.fie_jpeg_header_bytes()
    assert n == 623
    for (h, w, q) in [(1024, 1024, 75), (37, 53, 90), (1, 1, 30), (600, 800, 95)]:
        buf = (ctypes.c_ubyte * 1024)()
        assert L.fie_jpeg_write_header(buf, h, w, q) == n
        ours = bytes(buf[:n])
        ref = pil_bytes(np.zeros((h, w, 3), np.uint8), q)
        assert ours == ref[:n], (h, w, q)
        ql, qc = J.quant_tables(q)
        assert ours == J.header(h, w, ql, qc)
