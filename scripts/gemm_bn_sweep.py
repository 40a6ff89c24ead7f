"""Accumulator-width sweep per shape (fie_tune_gemm force_block_n) to calibrate pick_config's cost model."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3
dev = torch.device("cuda:0"); L = _lib.lib()
def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
gemms = [(16384, 1280, 1280, 1), (16384, 1280, 1280, 0), (65536, 640, 640, 1), (65536, 640, 640, 0), (16384, 1280, 5120, 1), (16384, 3840, 1280, 0), (65536, 1920, 640, 0), (65536, 640, 2560, 1),
         (2048, 1280, 1280, 1), (8192, 640, 640, 1), (2048, 3840, 1280, 0), (2048, 1280, 5120, 1)]
bns = [0, 256, 224, 192, 160, 128, 96, 64]
for (m, n, k, r) in gemms:
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    out = torch.empty((m, n), device=dev, dtype=torch.float16); res = torch.randn((m, n), device=dev).half(); bias = torch.randn((n,), device=dev)
    row = []
    for bn in bns:
        L.fie_tune_gemm(0, bn)
        row.append(timeit((lambda: ops.gemm(a, w, out=out, col_bias=bias, residual=res)) if r else (lambda: ops.gemm(a, w, out=out))))
    print(f"gemm M{m} N{n} K{k} res{r}: " + "  ".join(f"bn{b}:{t:7.1f}" for b, t in zip(bns, row)), flush=True)
convs = [(16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280), (16, 64, 64, 1280, 640), (2, 64, 64, 640, 640), (2, 32, 32, 1280, 1280), (8, 128, 128, 512, 512)]
for (nb, h, wd, cin, cout) in convs:
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    out = torch.empty((nb, h, wd, cout), device=dev, dtype=torch.float16)
    row = []
    for bn in bns:
        L.fie_tune_gemm(0, bn)
        row.append(timeit(lambda: ops.conv3x3(x, w, out=out), 4))
    print(f"conv [{nb},{h},{wd},{cin}]->{cout}: " + "  ".join(f"bn{b}:{t:7.1f}" for b, t in zip(bns, row)), flush=True)
L.fie_tune_gemm(0, 0)
