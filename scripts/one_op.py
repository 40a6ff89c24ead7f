"""Runs one op a few times (for `ncu --set full` on a single shape).
   python scripts/one_op.py gemm M N K [geglu|res|plain]
   python scripts/one_op.py conv N H W CIN COUT
   python scripts/one_op.py attn B HEADS NQ NKV
   python scripts/one_op.py gn N HW C"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_image_editing_with_generative_models_b200 import _lib, ops
from fast_image_editing_with_generative_models_b200.weights import pack_conv3x3
dev = torch.device("cuda:0")
if os.environ.get("FIE_DBG"): _lib.lib().fie_tune_gemm(int(os.environ["FIE_DBG"]) << 4, 0)      # 1 noTMA, 2 noMMA, 4 noEpilogue
if os.environ.get("FIE_HALO_MAX"): _lib.lib().fie_tune_conv_halo(1, int(os.environ["FIE_HALO_MAX"]))
kind = sys.argv[1]
args = sys.argv[2:]
iters = 3
if kind == "gemm":
    m, n, k = map(int, args[:3]); mode = args[3] if len(args) > 3 else "plain"
    a = torch.randn((m, k), device=dev).half(); w = (torch.randn((n, k), device=dev) / math.sqrt(k)).half()
    bias = torch.randn((n,), device=dev); res = torch.randn((m, n), device=dev).half()
    if mode == "geglu": fn = lambda: ops.gemm(a, w, col_bias=bias, act=ops.ACT_GEGLU)
    elif mode == "res": fn = lambda: ops.gemm(a, w, col_bias=bias, residual=res)
    else: fn = lambda: ops.gemm(a, w)
    work = 2.0 * m * n * k
elif kind == "conv":
    nb, h, wd, cin, cout = map(int, args[:5])
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv3x3((torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin)).half())
    bias = torch.randn((cout,), device=dev)
    fn = lambda: ops.conv3x3(x, w, col_bias=bias)
    work = 2.0 * nb * h * wd * cout * 9 * cin
elif kind == "up2x":
    from fast_image_editing_with_generative_models_b200.weights import pack_conv_up2x
    nb, h, wd, cin, cout = map(int, args[:5])
    x = torch.randn((nb, h, wd, cin), device=dev).half(); w = pack_conv_up2x(torch.randn((cout, cin, 3, 3), device=dev) / math.sqrt(9 * cin))
    bias = torch.randn((cout,), device=dev)
    fn = lambda: ops.conv_up2x(x, w, col_bias=bias)
    work = 2.0 * nb * 4 * h * wd * cout * 4 * cin
elif kind == "attn":
    b, heads, nq, nkv = map(int, args[:4])
    q = torch.randn((b * nq, heads * 64), device=dev).half(); k = torch.randn((b * nkv, heads * 64), device=dev).half(); v = torch.randn((b * nkv, heads * 64), device=dev).half()
    fn = lambda: ops.attention_d64(q, k, v, b, heads, nq, nkv)
    work = 4.0 * b * heads * nq * nkv * 64
elif kind == "ln":
    pass
elif kind == "gn":
    n, hw, c = map(int, args[:3])
    x = torch.randn((n, hw, c), device=dev).half(); g = torch.ones(c, device=dev); bt = torch.zeros(c, device=dev)
    fn = lambda: ops.groupnorm(x, g, bt, 1e-5, True, 32)
    work = 4.0 * n * hw * c
if kind == "ln":
    rows, c = map(int, args[:2])
    x = torch.randn((rows, c), device=dev).half(); g = torch.ones(c, device=dev); bt = torch.zeros(c, device=dev)
    fn = lambda: ops.layernorm(x, g, bt)
    work = 4.0 * rows * c
    iters = 20
for _ in range(2): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{kind} {' '.join(args)}: {ms*1e3:.1f} us  {work/ms/1e9:.1f} G(FLOP|B)/s")
