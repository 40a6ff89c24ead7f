import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3
dev = torch.device("cuda:0"); L = _lib.lib()
def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def tune(cg, kps, mt, bn): L.fie_tune_gemm(cg | (kps << 2) | (mt << 8), bn)
for (nb, h, wd, cin, cout) in [(8, 1024, 1024, 128, 128), (8, 1024, 1024, 64, 64), (16, 128, 128, 320, 320), (8, 512, 512, 256, 256)]:
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    out = torch.empty((nb, h, wd, cout), device=dev, dtype=torch.float16)
    fl = 2.0 * nb * h * wd * cout * 9 * cin
    for (cg, kps, mt, bn) in [(2, 2, 1, 0), (2, 1, 2, 0), (2, 2, 2, 0), (2, 1, 2, 128), (2, 1, 2, 96), (2, 1, 2, 64), (0, 0, 0, 0)]:
        tune(cg, kps, mt, bn)
        try:
            ms = timeit(lambda: ops.conv3x3(x, w, out=out)); print(f"conv [{nb},{h},{wd},{cin}]->{cout} cg{cg} kps{kps} mt{mt} bn{bn}: {ms:7.3f} ms {fl/ms/1e9:7.1f} TF/s", flush=True)
        except Exception as e:
            print("fail", cg, kps, mt, bn, str(e)[:100])
for (m, n, k) in [(65536, 640, 640), (16384, 1280, 1280)]:
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    out = torch.empty((m, n), device=dev, dtype=torch.float16); res = torch.randn((m, n), device=dev).half()
    for (cg, kps, mt, bn) in [(2, 2, 1, 0), (2, 1, 2, 128), (2, 2, 2, 128), (0, 0, 0, 0)]:
        tune(cg, kps, mt, bn)
        ms = timeit(lambda: ops.gemm(a, w, out=out, residual=res), 5); print(f"gemm+res M{m} N{n} K{k} cg{cg} kps{kps} mt{mt} bn{bn}: {ms*1e3:7.1f} us {2.0*m*n*k/ms/1e9:7.1f} TF/s", flush=True)
tune(0, 0, 0, 0)
