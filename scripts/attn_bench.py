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
for (b, h, nq, nkv) in [(16, 10, 4096, 4096), (16, 20, 1024, 1024), (16, 10, 4096, 77), (16, 20, 1024, 77), (2, 10, 4096, 4096), (2, 20, 1024, 1024), (1, 1, 128, 128), (1, 1, 128, 4096)]:
    c = h * 64
    qkv = torch.randn((b * nq, 3 * c), device=dev).half()
    kv = torch.randn((b * nkv, 2 * c), device=dev).half()
    if nq == nkv:
        f = lambda: ops.attention_d64(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], b, h, nq, nkv)
    else:
        f = lambda: ops.attention_d64(qkv[:, :c], kv[:, :c], kv[:, c:], b, h, nq, nkv)
    ms = timeit(f)
    fl = 4.0 * b * h * nq * nkv * 64
    ctas = b * h * ((nq + 127) // 128)
    print(f"b={b} h={h} nq={nq} nkv={nkv}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  ctas={ctas}  us/cta/tile={ms*1e3/ (ctas/296) / ((nkv+127)//128):.2f}", flush=True)
