"""KPS (K blocks per pipeline stage) 1 vs 2 per shape after the uniform-issue change."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3, pack_geglu
dev = torch.device("cuda:0"); L = _lib.lib()
def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
cases = []
for (m, n, k, mode) in [(16384, 10240, 1280, "geglu"), (16384, 1280, 5120, "res"), (16384, 3840, 1280, "plain"), (16384, 1280, 1280, "res"), (65536, 5120, 640, "geglu"), (65536, 640, 2560, "res")]:
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    bias = torch.randn((n,), device=dev); res = torch.randn((m, n), device=dev).half()
    if mode == "geglu": fn = (lambda a=a, w=w, bias=bias: ops.gemm(a, w, col_bias=bias, act=ops.ACT_GEGLU))
    elif mode == "res": fn = (lambda a=a, w=w, bias=bias, res=res: ops.gemm(a, w, col_bias=bias, residual=res))
    else: fn = (lambda a=a, w=w: ops.gemm(a, w))
    cases.append((f"gemm M{m} N{n} K{k} {mode}", fn, 2.0 * m * n * k))
for (nb, h, wd, cin, cout) in [(16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280), (8, 128, 128, 512, 512), (8, 256, 256, 512, 512)]:
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    cases.append((f"conv [{nb},{h},{wd},{cin}]->{cout}", (lambda x=x, w=w: ops.conv3x3(x, w)), 2.0 * nb * h * wd * cout * 9 * cin))
for name, fn, fl in cases:
    row = []
    for kps in (2, 1):
        L.fie_tune_gemm(kps << 2, 0)
        t = timeit(fn); row.append(f"kps{kps}: {t:8.1f} us {fl/t/1e6:7.0f} TF/s")
    print(f"{name:40s} " + "   ".join(row), flush=True)
L.fie_tune_gemm(0, 0)
