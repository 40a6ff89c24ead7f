"""Per-role wait-cycle accounting of the self-attention kernel (fie_attention_trace)."""
import sys, torch
sys.path.insert(0, ".")
from fast_image_editing_with_generative_models_b200 import ops, _lib

dev = torch.device("cuda:0")
L = _lib.lib()
for (b, h, nq, nkv) in [(16, 10, 4096, 4096), (16, 20, 1024, 1024)]:
    c = h * 64
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = (torch.randn(b * n, c, device=dev, generator=g).half() for n in (nq, nkv, nkv))
    for _ in range(3): ops.attention_d64(q, k, v, b, h, nq, nkv)
    torch.cuda.synchronize()
    ctas = b * h * (nq // 256)
    tr = torch.zeros(ctas, 16, dtype=torch.int64, device=dev)
    L.fie_attention_trace(tr.data_ptr()); ops.attention_d64(q, k, v, b, h, nq, nkv); torch.cuda.synchronize(); L.fie_attention_trace(None)
    t = tr.double().mean(0).tolist()
    nt = nkv // 128
    print(f"b{b} h{h} nq{nq} nkv{nkv}: per K/V tile (cycles, mean over {ctas} CTAs)")
    print(f"  MMA warp   total {t[0]/nt:7.0f}  wait kv {t[1]/nt:6.0f}  wait S-in-regs {t[2]/nt:6.0f}  wait P {t[3]/nt:6.0f}")
    for qq in (0, 1):
        o = 4 + 4 * qq
        print(f"  softmax q{qq} total {t[o]/nt:7.0f}  wait S {t[o+1]/nt:6.0f}  wait PV done {t[o+2]/nt:6.0f}  wait tmem st {t[o+3]/nt:6.0f}  compute {(t[o]-t[o+1]-t[o+2]-t[o+3])/nt:6.0f}")
