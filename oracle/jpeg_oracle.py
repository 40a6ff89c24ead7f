"""TEST INFRASTRUCTURE — numpy restatement of the baseline JPEG encoder the reference's ``edited.save(output_path)`` runs
(reference ``run_batch.py:224``, ``run_single_image.py:114``: PIL ``Image.save`` of an RGB image to ``*.jpg`` with default options
= libjpeg(-turbo) baseline sequential DCT, quality 75, 4:2:0 chroma subsampling, standard Huffman tables, JFIF APP0 header).

Pillow / libjpeg-turbo is an un-vendored third-party dependency (``Pillow>=9.5`` ``requirements.txt``; 12.2 / libjpeg-turbo 3.x here);
its published algorithm (IJG libjpeg: jccolor.c, jcsample.c, jfdctint.c, jcdctmgr.c, jchuff.c) is restated:

* RGB -> YCbCr: 16-bit fixed point (FIX(x) = round(x * 65536)), ``+ ONE_HALF`` for Y, ``+ 128 << 16 + ONE_HALF - 1`` for Cb / Cr
* edge expansion to whole MCUs (16 x 16) by replicating the last column / row
* h2v2 chroma downsampling: ``(a + b + c + d + bias) >> 2`` with bias alternating 1, 2, 1, 2 ... along a row
* level shift by 128 and the "islow" integer forward DCT (Loeffler-Ligtenberg-Moschytz, CONST_BITS 13, PASS1_BITS 2; output x 8)
* quantisation: round-half-up of the magnitude by ``8 * q`` with the quality-scaled Annex K tables (``(base * scale + 50) / 100``, clamp 1..255)
* Huffman coding with the Annex K tables: DC difference category + bits, AC (run, size) with ZRL / EOB, byte stuffing, 1-padding

PARITY: PINNED — ``tests/test_jpeg_oracle.py`` requires the byte stream to be IDENTICAL to Pillow's for every case (qualities 50-95,
ragged sizes, noise and smooth images).  Only tests / smoke / bench's checker legs may import this module."""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
STD_LUMA_Q = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                       18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99])
STD_CHROMA_Q = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                         99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99])
DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUMA_VALS = [0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
                0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
                0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
                0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
                0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
                0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
                0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
                  0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
                  0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
                  0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
                  0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
                  0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
                  0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa]


def quant_tables(quality: int):
    """jpeg_set_quality(quality, force_baseline=TRUE): (luma, chroma) tables in natural (row-major) order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return tuple(np.clip((base * scale + 50) // 100, 1, 255).astype(np.int32) for base in (STD_LUMA_Q, STD_CHROMA_Q))


def huff_table(bits, vals):
    """-> (code[256], size[256]) from the JPEG BITS / HUFFVAL lists (Annex C)."""
    code = np.zeros(256, np.uint32)
    size = np.zeros(256, np.uint8)
    c, k = 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            code[vals[k]], size[vals[k]] = c, length
            c += 1
            k += 1
        c <<= 1
    return code, size


def rgb_to_ycbcr(rgb: np.ndarray):
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    fix = lambda x: int(x * 65536 + 0.5)
    half, off = 1 << 15, 128 << 16
    y = (fix(0.29900) * r + fix(0.58700) * g + fix(0.11400) * b + half) >> 16
    cb = (-fix(0.16874) * r - fix(0.33126) * g + fix(0.50000) * b + off + half - 1) >> 16
    cr = (fix(0.50000) * r - fix(0.41869) * g - fix(0.08131) * b + off + half - 1) >> 16
    return y.astype(np.int32), cb.astype(np.int32), cr.astype(np.int32)


def _expand(p: np.ndarray, h: int, w: int) -> np.ndarray:
    return np.pad(p, ((0, h - p.shape[0]), (0, w - p.shape[1])), mode="edge")


def h2v2_downsample(p: np.ndarray) -> np.ndarray:
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (s + bias) >> 2


def fdct_islow(blocks: np.ndarray) -> np.ndarray:
    """jfdctint.c jpeg_fdct_islow on int32 [..., 8, 8] (level-shifted samples) -> coefficients scaled by 8."""
    C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137, f1_961=16069,
             f2_053=16819, f2_562=20995, f3_072=25172)
    CB, P1 = 13, 2

    def desc(x, n):
        return (x + (1 << (n - 1))) >> n

    def pass1d(d, first):
        t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
        t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
        t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
        t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
        t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
        out = [None] * 8
        if first:
            out[0] = (t10 + t11) << P1
            out[4] = (t10 - t11) << P1
        else:
            out[0] = desc(t10 + t11, P1)
            out[4] = desc(t10 - t11, P1)
        z1 = (t12 + t13) * C["f0_541"]
        sh = CB - P1 if first else CB + P1
        out[2] = desc(z1 + t13 * C["f0_765"], sh)
        out[6] = desc(z1 + t12 * (-C["f1_847"]), sh)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * C["f1_175"]
        t4, t5, t6, t7 = t4 * C["f0_298"], t5 * C["f2_053"], t6 * C["f3_072"], t7 * C["f1_501"]
        z1, z2, z3, z4 = z1 * (-C["f0_899"]), z2 * (-C["f2_562"]), z3 * (-C["f1_961"]) + z5, z4 * (-C["f0_390"]) + z5
        out[7] = desc(t4 + z1 + z3, sh)
        out[5] = desc(t5 + z2 + z4, sh)
        out[3] = desc(t6 + z2 + z3, sh)
        out[1] = desc(t7 + z1 + z4, sh)
        return np.stack(out, axis=-1)

    d = blocks.astype(np.int64)
    d = pass1d(d, True)                                           # rows
    d = np.swapaxes(pass1d(np.swapaxes(d, -1, -2), False), -1, -2)  # columns
    return d.astype(np.int32)


def quantize(coef: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    q = (qtbl.reshape(8, 8).astype(np.int64)) << 3
    a = np.abs(coef.astype(np.int64))
    return (np.sign(coef) * ((a + (q >> 1)) // q)).astype(np.int32)


def coefficients(rgb: np.ndarray, quality: int = 75):
    """uint8 [H,W,3] -> (quantised coefficients int32 [mcu_rows, mcu_cols, 6, 64] in ZIGZAG order (Y00 Y01 Y10 Y11 Cb Cr), tables)."""
    h, w, _ = rgb.shape
    hp, wp = -(-h // 16) * 16, -(-w // 16) * 16
    y, cb, cr = rgb_to_ycbcr(rgb)
    # jcprepct.c pads in two stages: the input rows up to a whole row GROUP (2 rows) and columns up to whole blocks before downsampling,
    # then the DOWNSAMPLED rows up to the iMCU height by replicating the last chroma row
    he = h + (h & 1)
    y = _expand(y, hp, wp)
    cb, cr = (_expand(h2v2_downsample(_expand(c, he, wp)), hp // 2, wp // 2) for c in (cb, cr))
    ql, qc = quant_tables(quality)

    def blocks(p, q):
        bh, bw = p.shape[0] // 8, p.shape[1] // 8
        b = (p - 128).reshape(bh, 8, bw, 8).transpose(0, 2, 1, 3)
        return quantize(fdct_islow(b), q).reshape(bh, bw, 64)[..., ZIGZAG]

    yb, cbb, crb = blocks(y, ql), blocks(cb, qc), blocks(cr, qc)
    mr, mc = hp // 16, wp // 16
    out = np.empty((mr, mc, 6, 64), np.int32)
    out[:, :, 0], out[:, :, 1] = yb[0::2, 0::2], yb[0::2, 1::2]
    out[:, :, 2], out[:, :, 3] = yb[1::2, 0::2], yb[1::2, 1::2]
    out[:, :, 4], out[:, :, 5] = cbb, crb
    # jccoefct.c compress_data: luma blocks that lie entirely outside the image (odd block counts: the right block of the last MCU
    # column, the bottom block row of the last MCU row) are DUMMY blocks — all AC zero, DC copied from the previous block of the MCU
    by, bx = -(-h // 8), -(-w // 8)
    if bx % 2:
        for r in (0, 2):
            out[:, -1, r + 1] = 0
            out[:, -1, r + 1, 0] = out[:, -1, r, 0]
    if by % 2:
        out[-1, :, 2:4] = 0
        out[-1, :, 2, 0] = out[-1, :, 1, 0]
        out[-1, :, 3, 0] = out[-1, :, 1, 0]
    return out, (ql, qc)


def header(h: int, w: int, ql: np.ndarray, qc: np.ndarray) -> bytes:
    """SOI, JFIF APP0, two DQT, SOF0 (4:2:0), four DHT, SOS — the segment sequence Pillow / libjpeg write."""
    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    out = b"\xff\xd8" + seg(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    out += seg(0xDB, bytes([0]) + bytes(int(v) for v in ql[ZIGZAG])) + seg(0xDB, bytes([1]) + bytes(int(v) for v in qc[ZIGZAG]))
    out += seg(0xC0, bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for tc_th, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS), (0x01, DC_CHROMA_BITS, DC_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += seg(0xC4, bytes([tc_th]) + bytes(bits) + bytes(vals))
    out += seg(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def entropy_encode(coefs: np.ndarray) -> bytes:
    """Huffman-codes [mcu_rows, mcu_cols, 6, 64] zigzag coefficients (sequential, interleaved scan) -> stuffed scan bytes."""
    dcl, acl = huff_table(DC_LUMA_BITS, DC_VALS), huff_table(AC_LUMA_BITS, AC_LUMA_VALS)
    dcc, acc = huff_table(DC_CHROMA_BITS, DC_VALS), huff_table(AC_CHROMA_BITS, AC_CHROMA_VALS)
    acc_bits, nbits = 0, 0
    out = bytearray()

    def put(code, size):
        nonlocal acc_bits, nbits
        acc_bits = (acc_bits << size) | (int(code) & ((1 << size) - 1))
        nbits += size
        while nbits >= 8:
            byte = (acc_bits >> (nbits - 8)) & 0xFF
            out.append(byte)
            if byte == 0xFF:
                out.append(0)
            nbits -= 8
        acc_bits &= (1 << nbits) - 1

    pred = [0, 0, 0]
    flat = coefs.reshape(-1, 6, 64)
    for mcu in flat:
        for bi in range(6):
            comp = 0 if bi < 4 else bi - 3
            dc_t, ac_t = (dcl, acl) if comp == 0 else (dcc, acc)
            blk = mcu[bi]
            diff = int(blk[0]) - pred[comp]
            pred[comp] = int(blk[0])
            mag = abs(diff)
            cat = mag.bit_length()
            put(dc_t[0][cat], int(dc_t[1][cat]))
            if cat:
                put(diff if diff >= 0 else diff - 1, cat)
            run = 0
            nz = np.nonzero(blk[1:])[0]
            last = nz[-1] + 1 if len(nz) else 0
            for k in range(1, last + 1):
                v = int(blk[k])
                if v == 0:
                    run += 1
                    continue
                while run > 15:
                    put(ac_t[0][0xF0], int(ac_t[1][0xF0]))
                    run -= 16
                cat = abs(v).bit_length()
                sym = (run << 4) | cat
                put(ac_t[0][sym], int(ac_t[1][sym]))
                put(v if v >= 0 else v - 1, cat)
                run = 0
            if last < 63:
                put(ac_t[0][0], int(ac_t[1][0]))
    if nbits:
        put(0x7F, 8 - nbits if nbits else 0)      # pad the last byte with 1-bits
    return bytes(out)


def encode(rgb: np.ndarray, quality: int = 75) -> bytes:
    """uint8 [H,W,3] -> the JPEG file bytes ``PIL.Image.fromarray(rgb).save(f, "JPEG", quality=quality)`` writes."""
    coefs, (ql, qc) = coefficients(rgb, quality)
    return header(rgb.shape[0], rgb.shape[1], ql, qc) + entropy_encode(coefs) + b"\xff\xd9"
