"""Host-side weight transforms (no GPU): the algebra of the folded LayerNorm, the hi/lo split of the c8 conv_in weights, LoRA fuse."""
import torch
import torch.nn.functional as F

from fast_image_editing_with_generative_models_b200.weights import fold_layernorm, fuse_lora, pack_conv3x3_c8


def test_fold_layernorm_is_layernorm_then_linear():
    """LN(x) W^T + b == rstd(x) * (x W''^T) + b'' with W'' = centred(gamma * W): what the GEMM epilogue (ln_stats_in) computes."""
    g = torch.Generator().manual_seed(0)
    m, k, n = 37, 96, 40
    x = torch.randn(m, k, generator=g, dtype=torch.float64) * 3 + 1.7          # rows with a clearly non-zero mean
    w = torch.randn(n, k, generator=g, dtype=torch.float64) / k ** 0.5
    b = torch.randn(n, generator=g, dtype=torch.float64)
    gamma = 1 + 0.3 * torch.randn(k, generator=g, dtype=torch.float64)
    beta = 0.2 * torch.randn(k, generator=g, dtype=torch.float64)
    ref = F.linear(F.layer_norm(x, (k,), gamma, beta, 1e-5), w, b)
    w16, bias = fold_layernorm(w, b, gamma, beta)
    assert w16.dtype == torch.float16 and bias.dtype == torch.float32 and w16.shape == (n, k)
    assert float(w16.float().sum(1).abs().max()) < 2e-2                        # rows are centred (up to fp16 rounding)
    # the statistics the producer epilogue accumulates: row sums and sums of squares
    mean = x.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt((x * x).mean(1, keepdim=True) - mean * mean + 1e-5)
    got = rstd * (x @ w16.double().t()) + bias.double()
    assert float((got - ref).abs().max()) < 5e-3 * float(ref.abs().max())      # only the fp16 rounding of W'' separates them
    # exact in exact arithmetic: same formula with un-rounded centred weights
    wc = w * gamma[None, :]
    wc = wc - wc.mean(1, keepdim=True)
    exact = rstd * (x @ wc.t()) + (b + w @ beta)
    assert float((exact - ref).abs().max()) < 1e-9


def test_fold_layernorm_without_bias_or_beta():
    w = torch.randn(8, 16)
    w16, bias = fold_layernorm(w, None, torch.ones(16), None)
    assert float(bias.abs().max()) == 0.0 and torch.allclose(w16.float(), (w - w.mean(1, keepdim=True)), atol=2e-3)


def test_c8_weights_keep_fp32_precision_as_hi_plus_lo():
    w = torch.randn(32, 3, 3, 3) / 5
    p = pack_conv3x3_c8(w).view(32, 3, 2, 8, 8)                                 # [co][kh][hi|lo][pixel][channel]
    rebuilt = (p[:, :, 0].float() + p[:, :, 1].float())[:, :, :3, :3].permute(0, 3, 1, 2)   # -> [co][c][kh][kw]
    assert float((rebuilt - w).abs().max()) < 1e-6
    assert float(p[:, :, :, 3:].abs().max()) == 0.0 and float(p[:, :, :, :, 3:].abs().max()) == 0.0


def test_fuse_lora_linear_and_conv():
    w = torch.randn(6, 5); a = torch.randn(2, 5); b = torch.randn(6, 2)
    x = torch.randn(4, 5)
    assert torch.allclose(F.linear(x, fuse_lora(w, a, b, 0.5)), F.linear(x, w) + 0.5 * F.linear(F.linear(x, a), b), atol=1e-5)
    wc = torch.randn(6, 5, 3, 3); ac = torch.randn(2, 5, 3, 3); bc = torch.randn(6, 2, 1, 1)
    xi = torch.randn(1, 5, 7, 7)
    assert torch.allclose(F.conv2d(xi, fuse_lora(wc, ac, bc, 0.25), padding=1), F.conv2d(xi, wc, padding=1) + 0.25 * F.conv2d(F.conv2d(xi, ac, padding=1), bc), atol=1e-4)
