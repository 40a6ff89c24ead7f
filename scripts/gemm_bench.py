"""Per-shape throughput of the tcgen05 GEMM / conv kernel (SDXL batch-8 shapes), 1-CTA vs CTA-pair form."""
import math, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3

dev = torch.device("cuda:0")
L = _lib.lib()
gemms = [(65536, 1920, 640), (65536, 640, 640), (65536, 5120, 640), (65536, 640, 2560), (16384, 3840, 1280), (16384, 1280, 1280),
         (16384, 10240, 1280), (16384, 1280, 5120), (16384, 16384, 512), (4096, 512, 16384), (2048, 1280, 1280), (2048, 10240, 1280), (8192, 640, 640)]
convs = [(16, 128, 128, 320, 320), (16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280), (16, 32, 32, 2560, 1280), (16, 64, 64, 1280, 640),
         (8, 1024, 1024, 128, 128), (8, 512, 512, 256, 256), (8, 256, 256, 512, 512), (8, 128, 128, 512, 512), (2, 128, 128, 320, 320), (2, 32, 32, 1280, 1280)]

def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

configs = [(2 | 8, 256), (2 | 8, 224), (2 | 8, 160), (2 | 8, 128), (0, 0)]
res = []
for (m, n, k) in gemms:
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    out = torch.empty((m, n), device=dev, dtype=torch.float16)
    row = []
    for cg, bn in configs:
        L.fie_tune_gemm(cg, bn)
        try:
            ms = timeit(lambda: ops.gemm(a, w, out=out)); row.append(2.0 * m * n * k / ms / 1e9)
        except Exception as e:
            row.append(float("nan"))
    print(f"gemm M={m:6d} N={n:5d} K={k:5d}: " + "  ".join(f"c{c&3}k{c>>2}b{b}:{r:6.0f}" for (c, b), r in zip(configs, row)), flush=True)
    res.append(("gemm", m, n, k, row))
for (nb, h, wd, cin, cout) in convs:
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    out = torch.empty((nb, h, wd, cout), device=dev, dtype=torch.float16)
    row = []
    for cg, bn in configs:
        L.fie_tune_gemm(cg, bn)
        try:
            ms = timeit(lambda: ops.conv3x3(x, w, out=out), 3); row.append(2.0 * nb * h * wd * cout * 9 * cin / ms / 1e9)
        except Exception as e:
            row.append(float("nan"))
    print(f"conv [{nb},{h},{wd},{cin}]->{cout}: " + "  ".join(f"c{c&3}k{c>>2}b{b}:{r:6.0f}" for (c, b), r in zip(configs, row)), flush=True)
    res.append(("conv", nb, h, wd, cin, cout, row))
L.fie_tune_gemm(0, 0)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"))
