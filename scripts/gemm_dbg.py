import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3
dev = torch.device("cuda:0"); L = _lib.lib()
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
nb, h, wd, cin, cout = 8, 256, 256, 512, 512
x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
out = torch.empty((nb, h, wd, cout), device=dev, dtype=torch.float16)
fl = 2.0 * nb * h * wd * cout * 9 * cin
kb = (nb*h*wd/128) * 72   # k-blocks per n-tile column
for cg in (1, 2):
 for kps in (1, 2):
  for bn in (256, 160, 128):
    for dbg, name in ((0, "full"), (1, "noTMA"), (2, "noMMA"), (3, "noTMA+noMMA")):
        L.fie_tune_gemm(cg | (kps << 2) | (dbg << 4), bn)
        ms = timeit(lambda: ops.conv3x3(x, w, out=out), 3)
        ntiles = (nb*h*wd/128/cg) * (cout/bn)
        per_kb_ns = ms*1e6 / (ntiles*72/ (148/cg))
        print(f"cg{cg} kps{kps} bn{bn:3d} {name:12s}: {ms:7.3f} ms  {fl/ms/1e9:7.1f} TFLOP/s-equiv  per-kblock {per_kb_ns:6.1f} ns", flush=True)
L.fie_tune_gemm(0, 0)
