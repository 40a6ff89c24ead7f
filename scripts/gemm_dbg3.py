"""Which stage bounds the short GEMMs: full vs no-TMA / no-MMA / no-epilogue variants (fie_tune_gemm debug bits)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
dev = torch.device("cuda:0"); L = _lib.lib()
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
shapes = [(16384, 1280, 1280), (65536, 640, 640), (16384, 10240, 1280), (65536, 5120, 640)]
for (m, n, k) in shapes:
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    geglu = n >= 5120
    out = torch.empty((m, n // 2 if geglu else n), device=dev, dtype=torch.float16)
    res = torch.randn((m, n), device=dev).half(); bias = torch.randn((n,), device=dev)
    for bn in ((0,) if geglu else (0, 256, 160, 128)):
        for dbg, name in ((0, "full"), (1, "noTMA"), (2, "noMMA"), (4, "noEpi"), (5, "noTMA+noEpi"), (7, "sync only")):
            L.fie_tune_gemm(0 | (0 << 2) | (dbg << 4), bn)
            if geglu:
                ms = timeit(lambda: ops.gemm(a, w, out=out, col_bias=bias, act=ops.ACT_GEGLU)); ms2 = float("nan")
            else:
                ms = timeit(lambda: ops.gemm(a, w, out=out))
                ms2 = timeit(lambda: ops.gemm(a, w, out=out, col_bias=bias, residual=res))
            print(f"M={m} N={n} K={k} bn{bn} {name:12s}: {ms*1e3:8.1f} us  {2.0*m*n*k/ms/1e9:7.1f} TF/s-equiv   bias+residual: {ms2*1e3:8.1f} us {2.0*m*n*k/ms2/1e9:7.1f}", flush=True)
L.fie_tune_gemm(0, 0)
