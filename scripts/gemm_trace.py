"""Per-role wait-cycle accounting of the tcgen05 GEMM (fie_gemm_trace): who waits on whom, per shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3
dev = torch.device("cuda:0"); L = _lib.lib()
trace = torch.zeros((148, 8), dtype=torch.int64, device=dev)

def run(name, fn, cfgs=((0, 0),)):
    for cg, bn in cfgs:
        L.fie_tune_gemm(cg, bn)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        trace.zero_(); L.fie_gemm_trace(trace.data_ptr()); fn(); torch.cuda.synchronize(); L.fie_gemm_trace(None)
        t = trace.float()
        lead = t[t[:, 4] > 0]            # leader CTAs (MMA issuer ran)
        f = lambda col, src=t: float(src[:, col][src[:, col - 0] >= 0].mean())
        print(f"{name:44s} cfg{cg},{bn:3d}: {us:8.1f} us | producer total {t[:,1].mean():8.0f} clk, blocked on empty {100*t[:,0].sum()/t[:,1].sum():4.1f}% | "
              f"issuer total {lead[:,4].mean():8.0f}, blocked on full {100*lead[:,2].sum()/lead[:,4].sum():4.1f}%, on acc-buffer {100*lead[:,3].sum()/lead[:,4].sum():4.1f}% | "
              f"epilogue total {t[:,6].mean():8.0f}, blocked on acc-full {100*t[:,5].sum()/t[:,6].sum():4.1f}% | first data at {lead[:,7].mean():6.0f} clk, epilogue tail {(t[:,6]-t[:,4].max()).mean() if False else (t[t[:,4]>0][:,6]-lead[:,4]).mean():6.0f} clk", flush=True)
    L.fie_tune_gemm(0, 0)

def gemm_case(m, n, k, mode):
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    bias = torch.randn((n,), device=dev); res = torch.randn((m, n), device=dev).half()
    if mode == "geglu": return lambda: ops.gemm(a, w, col_bias=bias, act=ops.ACT_GEGLU)
    if mode == "res": return lambda: ops.gemm(a, w, col_bias=bias, residual=res)
    return lambda: ops.gemm(a, w)

def conv_case(nb, h, wd, cin, cout):
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    bias = torch.randn((cout,), device=dev)
    return lambda: ops.conv3x3(x, w, col_bias=bias)

kps1, kps2 = 2 | (1 << 2), 2 | (2 << 2)
if len(sys.argv) > 1 and sys.argv[1] == "conv":          # python scripts/gemm_trace.py conv  -> the VAE / UNet convolution shapes only
    for c in [(8, 1024, 1024, 128, 128), (8, 512, 512, 256, 256), (8, 256, 256, 512, 512), (16, 128, 128, 320, 320), (16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280)]:
        run(f"conv {c}", conv_case(*c), ((0, 0),))
    sys.exit(0)
for (m, n, k, mode) in [(16384, 1280, 1280, "plain"), (16384, 1280, 1280, "res"), (16384, 10240, 1280, "geglu"), (65536, 5120, 640, "geglu"),
                        (65536, 640, 640, "res"), (16384, 1280, 5120, "res"), (16384, 3840, 1280, "plain")]:
    run(f"gemm M{m} N{n} K{k} {mode}", gemm_case(m, n, k, mode), ((0, 0),))
for c in [(16, 64, 64, 640, 640), (16, 32, 32, 1280, 1280), (8, 128, 128, 512, 512)]:
    run(f"conv {c}", conv_case(*c), ((0, 0),))
