import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (n, h, cout, ld, act) in [(8, 1024, 128, 128, 0), (8, 1024, 16, 64, 1), (16, 128, 320, 320, 0), (8, 128, 4, 4, 0), (8, 128, 512, 512, 0)]:
    x = torch.randn((n, h, h, 4), device=dev).half(); w = torch.randn((cout, 3, 3, 4), device=dev); b = torch.randn((cout,), device=dev)
    ms = timeit(lambda: ops.conv3x3_cin4(x, w, b, cout, ld_out=ld, act=act))
    print(f"n={n} h={h} cout={cout} ld={ld}: {ms:8.3f} ms  out {n*h*h*ld*2/ms/1e6:8.1f} GB/s", flush=True)
