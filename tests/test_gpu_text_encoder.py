"""GPU: the CLIP text encoders (the stage next to the path) on the fie_b200 kernels against the REAL reference
implementation — transformers' ``CLIPTextModel`` / ``CLIPTextModelWithProjection`` (the classes the diffusers pipeline runs in
``encode_prompt``), fp32 on the CPU, loaded with the identical seeded random-init state dict.  Tolerances: fp16 activations."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _hf_model(cfg, params):
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection
    hc = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                        projection_dim=cfg.projection_dim or 512, num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                        max_position_embeddings=cfg.max_positions, hidden_act=cfg.hidden_act, layer_norm_eps=cfg.layer_norm_eps,
                        eos_token_id=2, bos_token_id=0, pad_token_id=1)          # eos_token_id 2 = the argmax pooling of the original CLIP
    m = (CLIPTextModelWithProjection if cfg.projection_dim else CLIPTextModel)(hc).eval()
    res = m.load_state_dict(params, strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys), res
    return m


def _ids(T, cfg):
    prompts = ["a rusty bicycle leaning against a red brick wall", "", "a watercolor painting of a fox in the snow , highly detailed"]
    return torch.stack([T.pseudo_token_ids(p, cfg.vocab_size, cfg.max_positions) for p in prompts])


@pytest.mark.parametrize("which", ["tiny-quick", "tiny-gelu-proj", "clip-l", "openclip-bigg"])
def test_clip_text_encoder_matches_transformers(cuda_dev, which):
    from fast_image_editing_with_generative_models_b200 import text_encoder as T
    cfg = {"tiny-quick": T.tiny_clip_config(False, "quick_gelu"), "tiny-gelu-proj": T.tiny_clip_config(True, "gelu"),
           "clip-l": T.clip_l_config(), "openclip-bigg": T.openclip_bigg_config()}[which]
    params = T.make_clip_params(cfg)
    ids = _ids(T, cfg)
    with torch.no_grad():
        ref = _hf_model(cfg, params)(input_ids=ids, output_hidden_states=True)
    enc = T.CLIPTextEncoder(params, cfg, cuda_dev)
    hid, pooled, embeds = enc.forward(ids)
    r_hid = ref.hidden_states[-2]
    err = float((hid.float().cpu() - r_hid).abs().max())
    print(f"[{which}] hidden_states[-2] max-abs {err:.4g} (ref absmax {float(r_hid.abs().max()):.3g}, std {float(r_hid.std()):.3g})")
    assert err <= 2e-2 * max(1.0, float(r_hid.abs().max()))
    r_pool = ref.pooler_output if not cfg.projection_dim else None
    if cfg.projection_dim:
        e = float((embeds.float().cpu() - ref.text_embeds).abs().max())
        print(f"[{which}] text_embeds max-abs {e:.4g} (ref absmax {float(ref.text_embeds.abs().max()):.3g})")
        assert e <= 2e-2 * max(1.0, float(ref.text_embeds.abs().max()))
    else:
        assert float((pooled.float().cpu() - r_pool).abs().max()) <= 2e-2 * max(1.0, float(r_pool.abs().max()))
    # the causal mask matters: token t must not see later tokens -> changing the tail of the prompt leaves earlier rows untouched
    ids2 = ids.clone(); ids2[0, 5:] = cfg.vocab_size - 1
    hid2, _, _ = enc.forward(ids2)
    assert torch.equal(hid2[0, :5], hid[0, :5]) and not torch.equal(hid2[0, 5:], hid[0, 5:])


def test_sdxl_encode_prompt_shapes(cuda_dev):
    from fast_image_editing_with_generative_models_b200 import text_encoder as T
    c1, c2 = T.tiny_clip_config(False, "quick_gelu"), T.tiny_clip_config(True, "gelu")
    te = T.SDXLTextEncoders(T.make_clip_params(c1), c1, T.make_clip_params(c2), c2, cuda_dev)
    ids = torch.stack([T.pseudo_token_ids("", c1.vocab_size), T.pseudo_token_ids("a rusty bicycle", c1.vocab_size)])     # [negative, positive]
    pe, pooled = te.encode(ids, ids)
    assert pe.shape == (2, 77, c1.hidden_size + c2.hidden_size) and pooled.shape == (2, c2.projection_dim)
    assert pe.dtype == torch.float16 and bool(torch.isfinite(pe.float()).all())
