"""TEST INFRASTRUCTURE — ctypes loader for the plain-C oracle (oracle/canny_oracle.c)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfie_oracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.fie_oracle_canny_u8.restype = ctypes.c_int
    lib.fie_oracle_canny_u8.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 7
    lib.fie_oracle_gaussian_blur5_u8.restype = ctypes.c_int
    lib.fie_oracle_gaussian_blur5_u8.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
    return lib


def canny_u8(img: np.ndarray, low=100, high=200, replicate3=False) -> np.ndarray:
    """img: [N,H,W,3] or [N,H,W] uint8 -> [N,H,W(,3)] uint8."""
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim in (3, 4)
    ch = 3 if img.ndim == 4 else 1
    n, h, w = img.shape[:3]
    out = np.empty((n, h, w, 3) if replicate3 else (n, h, w), np.uint8)
    rc = _lib().fie_oracle_canny_u8(img.ctypes.data, out.ctypes.data, n, h, w, ch, int(np.floor(low)), int(np.floor(high)), int(replicate3))
    if rc:
        raise RuntimeError(f"fie_oracle_canny_u8 failed rc={rc}")
    return out


def gaussian_blur5_u8(img: np.ndarray) -> np.ndarray:
    """img: [N,H,W] or [N,H,W,C] uint8 -> same shape; restates cv2.GaussianBlur(img, (5, 5), 0)."""
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim in (3, 4)
    n, h, w = img.shape[:3]
    out = np.empty_like(img)
    rc = _lib().fie_oracle_gaussian_blur5_u8(img.ctypes.data, out.ctypes.data, n, h, w, img.shape[3] if img.ndim == 4 else 1)
    if rc:
        raise RuntimeError(f"fie_oracle_gaussian_blur5_u8 failed rc={rc}")
    return out
