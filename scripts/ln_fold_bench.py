"""A/B of the folded-LayerNorm epilogues on the SDXL transformer GEMM shapes (isolated launches, CUDA-event timed)."""
import math, sys, torch
sys.path.insert(0, ".")
from fast_image_editing_with_generative_models_b200 import ops
from fast_image_editing_with_generative_models_b200.weights import fold_layernorm, pack_geglu

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
R = lambda *s: torch.randn(*s, device=dev, generator=g)


def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (m, c) in [(16384, 1280), (65536, 640)]:
    x = R(m, c).half(); res = R(m, c).half()
    st = ops.zeros_i64((m, 2), dev); st2 = ops.zeros_i64((m, 2), dev)
    gamma, beta = 1 + 0.1 * R(c), 0.1 * R(c)
    for name, n, geglu in [("qkv", 3 * c, False), ("q2", c, False), ("geglu", 8 * c, True)]:
        w = R(n, c) / math.sqrt(c); b = R(n)
        w16, bias = fold_layernorm(w, b, gamma, beta)
        if geglu:
            w16, bias = pack_geglu(w16, bias)
        act = ops.ACT_GEGLU if geglu else ops.ACT_NONE
        x0 = ops.gemm(x, torch.eye(c, device=dev).half(), ln_out=st)
        t_plain = timed(lambda: ops.gemm(x, w16, col_bias=bias, act=act))
        t_ln = timed(lambda: ops.gemm(x, w16, col_bias=bias, act=act, ln_in=(st, 1e-5)))
        t_lnk = timed(lambda: ops.layernorm(x, gamma, beta))
        fl = 2.0 * m * n * c
        print(f"M{m} C{c} {name:6s} N{n}: plain {t_plain*1e3:7.1f} us ({fl/t_plain/1e9:6.0f} TF/s)  ln_in {t_ln*1e3:7.1f} us ({fl/t_ln/1e9:6.0f} TF/s)  [+layernorm kernel {t_lnk*1e3:6.1f} us]")
    for name, k in [("out", c), ("ff_out", 4 * c)]:
        a = R(m, k).half(); w = (R(c, k) / math.sqrt(k)).half(); b = R(c)
        t_plain = timed(lambda: ops.gemm(a, w, col_bias=b, residual=res))
        t_st = timed(lambda: ops.gemm(a, w, col_bias=b, residual=res, ln_out=st2))
        fl = 2.0 * m * c * k
        print(f"M{m} C{c} {name:6s} K{k}: plain {t_plain*1e3:7.1f} us ({fl/t_plain/1e9:6.0f} TF/s)  ln_out {t_st*1e3:7.1f} us ({fl/t_st/1e9:6.0f} TF/s)")
