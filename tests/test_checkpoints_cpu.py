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
