"""CPU: real-weight loading — diffusers ``config.json`` contents -> configs, safetensors round trip, LoRA key styles.
The SDXL / SSD-1B dictionaries below restate the fields of the published ``unet/config.json`` files that define the topology."""
import torch

from fast_image_editing_with_generative_models_b200 import checkpoints as K
from fast_image_editing_with_generative_models_b200 import configs as C
from fast_image_editing_with_generative_models_b200 import synthetic as S

SDXL_UNET = {"block_out_channels": [320, 640, 1280], "layers_per_block": 2, "attention_head_dim": [5, 10, 20], "cross_attention_dim": 2048,
             "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"], "mid_block_type": "UNetMidBlock2DCrossAttn",
             "up_block_types": ["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"], "transformer_layers_per_block": [1, 2, 10],
             "addition_time_embed_dim": 256, "projection_class_embeddings_input_dim": 2816, "norm_num_groups": 32, "in_channels": 4, "out_channels": 4}
SSD1B_UNET = {**SDXL_UNET, "mid_block_type": "UNetMidBlock2D", "transformer_layers_per_block": [1, [2, 2], [4, 4]],
              "reverse_transformer_layers_per_block": [[4, 4, 10], [2, 1, 1], 1]}
CN_SMALL = {**SDXL_UNET, "down_block_types": ["DownBlock2D", "DownBlock2D", "DownBlock2D"], "mid_block_type": "UNetMidBlock2D",
            "transformer_layers_per_block": 1, "conditioning_embedding_out_channels": [16, 32, 96, 256]}


def _topology(c):
    return (tuple(c.block_out_channels), c.layers_per_block, tuple(map(tuple, c.down_depths)), c.mid_depth, tuple(map(tuple, c.up_depths)),
            c.cross_attention_dim, c.time_embed_dim, c.addition_time_embed_dim, c.projection_class_embeddings_input_dim)


def test_published_unet_configs_map_to_the_engine_topologies():
    assert _topology(K.unet_config_from_json(SDXL_UNET)) == _topology(C.sdxl_unet_config())
    assert _topology(K.unet_config_from_json(SSD1B_UNET)) == _topology(C.ssd1b_unet_config())
    cn = K.controlnet_config_from_json(CN_SMALL)
    ref = C.controlnet_config(False)
    assert not cn.full and tuple(cn.cond_channels) == tuple(ref.cond_channels) and _topology(cn.unet) == _topology(ref.unet)
    full = K.controlnet_config_from_json({**SDXL_UNET, "conditioning_embedding_out_channels": [16, 32, 96, 256]})
    assert full.full and _topology(full.unet) == _topology(C.controlnet_config(True).unet)


def test_checkpoint_round_trip(tmp_path):
    ucfg, ccfg, vcfg = C.tiny_unet_config(), C.tiny_controlnet_config(True), C.tiny_vae_config()
    unet, cn, vae = S.make_unet_params(ucfg), S.make_controlnet_params(ccfg), S.make_vae_params(vcfg)
    K.save_model_dir(tmp_path / "unet", K.unet_config_to_json(ucfg), unet, fp16=False)
    K.save_model_dir(tmp_path / "controlnet", {**K.unet_config_to_json(ccfg.unet), "conditioning_embedding_out_channels": list(ccfg.cond_channels)}, cn, fp16=False)
    K.save_model_dir(tmp_path / "vae", {"block_out_channels": list(vcfg.block_out_channels), "layers_per_block": vcfg.layers_per_block,
                                         "latent_channels": vcfg.latent_channels, "scaling_factor": vcfg.scaling_factor}, vae, fp16=False)
    st = K.load_state(str(tmp_path / "unet"), str(tmp_path / "controlnet"), str(tmp_path / "vae"))
    assert _topology(st["unet_cfg"]) == _topology(ucfg) and _topology(st["cn_cfg"].unet) == _topology(ccfg.unet) and st["cn_cfg"].full
    assert tuple(st["vae_cfg"].block_out_channels) == tuple(vcfg.block_out_channels)
    for got, ref in ((st["unet"], unet), (st["cn"], cn), (st["vae"], vae)):
        assert set(got) == set(ref) and all(torch.equal(got[k], ref[k].float()) for k in ref)


def test_lora_key_styles():
    ucfg = C.tiny_unet_config()
    unet = S.make_unet_params(ucfg)
    peft = S.make_lora_params(unet, rank=4)
    names = sorted({k[: -len(".lora_A.weight")] for k in peft if k.endswith(".lora_A.weight")})
    assert names
    kohya = {}
    for n in names:
        flat = "lora_unet_" + n.replace(".", "_")
        kohya[flat + ".lora_down.weight"] = peft[n + ".lora_A.weight"]
        kohya[flat + ".lora_up.weight"] = peft[n + ".lora_B.weight"]
        kohya[flat + ".alpha"] = torch.tensor(2.0)                    # alpha / rank = 0.5
    got = K.lora_to_peft(kohya, list(unet.keys()))
    assert set(got) == set(peft)
    for n in names:
        assert torch.equal(got[n + ".lora_A.weight"], peft[n + ".lora_A.weight"].float())
        assert torch.allclose(got[n + ".lora_B.weight"], peft[n + ".lora_B.weight"].float() * 0.5)
    prefixed = {"unet." + k: v for k, v in peft.items()}
    got2 = K.lora_to_peft(prefixed, list(unet.keys()))
    assert set(got2) == set(peft) and all(torch.equal(got2[k], peft[k].float()) for k in peft)


def test_lora_with_unmapped_keys_is_an_error():
    """A LoRA tensor that matches no UNet parameter must not be dropped silently."""
    import pytest
    unet = S.make_unet_params(C.tiny_unet_config())
    bad = {"lora_unet_no_such_layer.lora_down.weight": torch.zeros(4, 8), "lora_unet_no_such_layer.lora_up.weight": torch.zeros(8, 4)}
    with pytest.raises(ValueError, match="do not match any UNet parameter"):
        K.lora_to_peft(bad, list(unet.keys()))


def test_mid_block_type_null_is_rejected():
    import pytest
    with pytest.raises(ValueError, match="mid_block_type"):
        K.unet_config_from_json({**SDXL_UNET, "mid_block_type": None})


def test_clip_text_encoder_round_trip_and_env_expansion(tmp_path, monkeypatch):
    """text_encoder / text_encoder_2 folders in the transformers layout (config.json + model.fp16.safetensors) -> CLIPTextConfig + params;
    the published SDXL values: CLIP-L is a CLIPTextModel (projection_dim present but unused), bigG a CLIPTextModelWithProjection."""
    from fast_image_editing_with_generative_models_b200 import text_encoder as T
    from fast_image_editing_with_generative_models_b200.editor import checkpoints_from_env, expand_checkpoints
    c1, c2 = T.tiny_clip_config(False, "quick_gelu"), T.tiny_clip_config(True, "gelu")
    p1, p2 = T.make_clip_params(c1), T.make_clip_params(c2)
    root = tmp_path / "pipe"
    K.save_clip_dir(str(root / "text_encoder"), c1, p1, fp16=False)
    K.save_clip_dir(str(root / "text_encoder_2"), c2, p2, fp16=False)
    g1, q1 = K.load_clip_dir(str(root / "text_encoder"))
    g2, q2 = K.load_clip_dir(str(root / "text_encoder_2"))
    assert g1.projection_dim is None and g2.projection_dim == c2.projection_dim
    assert (g1.hidden_size, g1.num_layers, g1.num_heads, g1.intermediate_size, g1.hidden_act) == (c1.hidden_size, c1.num_layers, c1.num_heads, c1.intermediate_size, "quick_gelu")
    assert g2.hidden_act == "gelu" and set(q1) == set(p1) and set(q2) == set(p2)
    assert all(torch.equal(q2[k], p2[k].float()) for k in p2)
    published_l = {"architectures": ["CLIPTextModel"], "hidden_size": 768, "num_hidden_layers": 12, "num_attention_heads": 12, "intermediate_size": 3072,
                   "hidden_act": "quick_gelu", "projection_dim": 768, "vocab_size": 49408, "max_position_embeddings": 77}
    published_g = {"architectures": ["CLIPTextModelWithProjection"], "hidden_size": 1280, "num_hidden_layers": 32, "num_attention_heads": 20,
                   "intermediate_size": 5120, "hidden_act": "gelu", "projection_dim": 1280, "vocab_size": 49408, "max_position_embeddings": 77}
    a, b = K.clip_config_from_json(published_l, "l"), K.clip_config_from_json(published_g, "g")
    ref_l, ref_g = T.clip_l_config(), T.openclip_bigg_config()
    for got, ref in ((a, ref_l), (b, ref_g)):
        assert (got.hidden_size, got.num_layers, got.num_heads, got.intermediate_size, got.hidden_act, got.projection_dim) == \
               (ref.hidden_size, ref.num_layers, ref.num_heads, ref.intermediate_size, ref.hidden_act, ref.projection_dim)
    # folder expansion used by the CLIs / environment variables
    (root / "tokenizer").mkdir(); (root / "tokenizer_2").mkdir()
    ck = expand_checkpoints(str(root), str(tmp_path / "cn"), lora=str(tmp_path / "lora.safetensors"))
    assert ck["unet"].endswith("unet") and ck["vae"].endswith("vae") and ck["controlnet"].endswith("cn") and ck["lora"].endswith("lora.safetensors")
    assert {"text_encoder", "text_encoder_2", "tokenizer", "tokenizer_2"} <= set(ck)
    monkeypatch.delenv("FIE_CHECKPOINTS", raising=False)
    assert checkpoints_from_env() is None
    monkeypatch.setenv("FIE_CHECKPOINTS", str(root)); monkeypatch.setenv("FIE_CONTROLNET", str(tmp_path / "cn"))
    assert checkpoints_from_env()["controlnet"].endswith("cn")
    import pytest
    with pytest.raises(ValueError, match="ControlNet"):
        expand_checkpoints(str(root), None)
