"""``FastEditor`` — drop-in mirror of the reference class (``src/pipeline.py:17-293``) on the B200 engine.

Same constructor, ``edit`` / ``preprocess_image`` / ``clear_memory`` / ``get_memory_usage`` signatures, attributes
(``model_name, device, dtype, config, controlnet, pipe``) and error behaviour (``ValueError`` for an unknown model,
Python exceptions per image, usable after a failed call) as the reference; the arithmetic runs in the hand-written
CUDA kernels behind ``include/fie_b200.h``.  Differences, all forced by the environment and stated in DESIGN.md:

* weights are seeded synthetic tensors of the named architectures (no checkpoints / network); a state dict can be
  injected with ``state=``;
* the two CLIP text encoders are outside the accelerated path: prompts are mapped to deterministic
  synthetic embeddings (seeded by the prompt text) unless ``prompt_encoder`` is supplied;
* compute is fp16 on the GPU (``use_full_precision`` is accepted and recorded but the kernels are fp16/fp32-accumulate);
  ``enable_cpu_offload`` is a no-op on a 180 GB part.
"""
from __future__ import annotations

import os
import zlib
from typing import Sequence, Callable, Dict, Optional

import numpy as np
import torch
from PIL import Image

from . import model_zoo, ops
from . import synthetic as S


class _PipeShim:
    """What callers touch on ``editor.pipe``: ``run_batch.py:157-158`` calls ``set_progress_bar_config(disable=True)``."""

    def __init__(self, engine):
        self.engine = engine
        self.progress_bar_config: Dict = {}

    def set_progress_bar_config(self, **kwargs):
        self.progress_bar_config.update(kwargs)


class FastEditor:
    """Fast image editor: SDXL / SSD-1B + ControlNet (Canny) + LCM, 4-step img2img with structure preservation."""

    MODEL_CONFIGS = {
        "sdxl": {
            "base_model": "stabilityai/stable-diffusion-xl-base-1.0",
            "lcm_lora": "latent-consistency/lcm-lora-sdxl",
            "use_full_lcm": False,
            "description": "Full SDXL (highest quality, ~6GB VRAM)",
        },
        "ssd-1b": {
            "base_model": "segmind/SSD-1B",
            "lcm_model": "latent-consistency/lcm-ssd-1b",
            "use_full_lcm": True,
            "description": "SSD-1B distilled (50% smaller, 60% faster, ~4GB VRAM)",
        },
    }

    def __init__(self, model_name="sdxl", device="cuda", dtype=torch.float16, enable_cpu_offload=True, use_full_precision=False,
                 use_full_controlnet=False, *, state: Optional[Dict] = None, prompt_encoder: Optional[Callable] = None, tiny: bool = False,
                 text_encoders: bool = False, checkpoints: Optional[Dict[str, str]] = None, tokenizers: Optional[Sequence[str]] = None,
                 verbose: bool = True):
        if model_name not in self.MODEL_CONFIGS:
            raise ValueError(f"Unknown model: {model_name}. Choose from {list(self.MODEL_CONFIGS.keys())}")
        self.model_name = model_name
        self.device = device
        self.dtype = torch.float32 if use_full_precision else dtype
        self.enable_cpu_offload = enable_cpu_offload
        self.use_full_controlnet = use_full_controlnet
        self.config = self.MODEL_CONFIGS[model_name]
        self._say = print if verbose else (lambda *a, **k: None)
        self._say(f"[FastEditor] Initializing with {model_name.upper()} (B200-native engine)")
        self._say(f"[FastEditor] {self.config['description']}")
        self._say(f"[FastEditor] Device: {device}, Dtype: {self.dtype} (kernels compute fp16 with fp32 accumulation)")
        if not str(device).startswith("cuda"):
            raise RuntimeError("FastEditor (B200-native) needs a CUDA device; there is no CPU path")
        if not torch.cuda.is_available():
            raise RuntimeError("FastEditor (B200-native): no CUDA device available")
        tiny = tiny or os.environ.get("FIE_TINY") == "1"     # small same-topology models: CLI smoke tests
        if state is None and checkpoints:
            # real weights: diffusers folders {unet, controlnet, vae} (+ optional LCM-LoRA file), see checkpoints.py
            from . import checkpoints as K
            self._say("[FastEditor] Loading checkpoints...")
            state = K.load_state(checkpoints["unet"], checkpoints["controlnet"], checkpoints["vae"], checkpoints.get("lora"))
        if state is None:
            self._say("[FastEditor] Generating seeded synthetic weights (no checkpoints available offline)...")
            state = model_zoo.synthetic_state(model_name, use_full_controlnet, tiny)
        self._engine = model_zoo.build_engine(state, device)
        self._engine.use_graphs = True      # every edit has the same shapes: replay it as one CUDA graph after the first call
        self.controlnet = self._engine.cn
        self.pipe = _PipeShim(self._engine)
        self._prompt_encoder = prompt_encoder
        self._text = None
        if text_encoders and prompt_encoder is None:
            # the two CLIP text towers of encode_prompt on the same kernels (seeded random-init weights; the BPE
            # tokenizer's vocabulary is not available offline, so prompts go through text_encoder.pseudo_token_ids)
            from . import text_encoder as T
            ucfg = self._engine.unet.cfg
            c1, c2 = (T.tiny_clip_config(False, "quick_gelu"), T.tiny_clip_config(True, "gelu")) if tiny else (T.clip_l_config(), T.openclip_bigg_config())
            pooled_dim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
            if c1.hidden_size + c2.hidden_size != ucfg.cross_attention_dim or c2.projection_dim != pooled_dim:
                raise ValueError("text encoder widths do not match the UNet's cross_attention_dim / pooled embedding size")
            self._say("[FastEditor] Building the CLIP text encoders (synthetic weights)...")
            self._text = T.SDXLTextEncoders(T.make_clip_params(c1), c1, T.make_clip_params(c2), c2, device)
            self._text_vocab = c1.vocab_size
            self._tokenizers = None
            if tokenizers:
                # (tokenizer, tokenizer_2) folders with vocab.json + merges.txt: the real CLIP BPE (tokenizer.py)
                from .tokenizer import CLIPBPETokenizer
                self._tokenizers = (CLIPBPETokenizer.from_files(tokenizers[0]), CLIPBPETokenizer.from_files(tokenizers[1], pad_token="!"))
        self._say("[FastEditor] Initialization complete!")

    # ---- prompt -> embeddings (text encoders are not on the accelerated path) ----
    def _encode_prompt(self, prompt: str, negative_prompt: str):
        ucfg = self._engine.unet.cfg
        if self._prompt_encoder is not None:
            return self._prompt_encoder(prompt, negative_prompt)
        if self._text is not None:
            if self._tokenizers is not None:
                ids1, ids2 = (torch.tensor(t([negative_prompt, prompt]), dtype=torch.int64) for t in self._tokenizers)
                return self._text.encode(ids1, ids2)
            from .text_encoder import pseudo_token_ids
            ids = torch.stack([pseudo_token_ids(negative_prompt, self._text_vocab), pseudo_token_ids(prompt, self._text_vocab)])   # the empty negative prompt is ENCODED, as in the reference
            return self._text.encode(ids, ids)          # ([neg, pos] x 77 x 2048, [neg, pos] x 1280)
        pooled_dim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
        pos = S.synthetic_prompt(zlib.crc32(prompt.encode("utf-8")) % (1 << 30), ucfg.cross_attention_dim, pooled_dim)
        neg = S.synthetic_prompt(zlib.crc32(negative_prompt.encode("utf-8")) % (1 << 30), ucfg.cross_attention_dim, pooled_dim)
        return torch.stack([neg[0][1], pos[0][1]]), torch.stack([neg[1][1], pos[1][1]])

    def preprocess_image(self, image, low_threshold=100, high_threshold=200):
        """PIL image (RGB, or 2-D gray array) -> PIL RGB Canny edge map (reference ``src/pipeline.py:183-210``)."""
        image_np = np.ascontiguousarray(np.array(image))
        if image_np.dtype != np.uint8 or image_np.ndim not in (2, 3):
            raise ValueError("preprocess_image expects an 8-bit RGB or gray image")
        if image_np.ndim == 3 and image_np.shape[2] != 3:
            image_np = np.ascontiguousarray(image_np[..., :3])
        d = torch.from_numpy(image_np[None]).to(self.device)
        edges = ops.canny(d, int(np.floor(low_threshold)), int(np.floor(high_threshold)), out_channels=3)
        return Image.fromarray(edges[0].cpu().numpy())

    def edit(self, image, prompt, negative_prompt="", strength=0.80, num_inference_steps=4, guidance_scale=1.5,
             controlnet_conditioning_scale=0.5, canny_low_threshold=100, canny_high_threshold=200, seed=None):
        """Edit an image with a text prompt, preserving structure via Canny conditioning -> PIL RGB 1024x1024."""
        gen = torch.Generator(device=self.device)
        if seed is not None:
            gen.manual_seed(int(seed))
        else:
            gen.seed()
        # image.resize((1024, 1024), Image.LANCZOS) of the reference (src/pipeline.py:251) on the GPU, bit-identical to Pillow
        arr = np.ascontiguousarray(np.array(image.convert("RGB")))
        img = torch.from_numpy(arr[None]).to(self.device)
        if img.shape[1] != 1024 or img.shape[2] != 1024:
            img = ops.resize_lanczos(img, 1024, 1024)
        pe, pl = self._encode_prompt(prompt, negative_prompt)
        n_exec = min(int(num_inference_steps * strength), num_inference_steps)
        # reference RNG order: posterior sample, init noise, then one draw per non-final executed step
        noises = [torch.randn((1, 4, 128, 128), generator=gen, device=self.device, dtype=torch.float16) for _ in range(2 + max(n_exec - 1, 0))]
        out = self._engine.edit_batch(img, pe, pl, noises, strength=strength, num_inference_steps=num_inference_steps,
                                      guidance_scale=guidance_scale, controlnet_conditioning_scale=controlnet_conditioning_scale,
                                      canny_low=int(np.floor(canny_low_threshold)), canny_high=int(np.floor(canny_high_threshold)))
        return Image.fromarray(out.images[0].cpu().numpy())

    def clear_memory(self):
        """Clear GPU memory cache."""
        if str(self.device).startswith("cuda"):
            torch.cuda.empty_cache()

    def get_memory_usage(self):
        if str(self.device).startswith("cuda"):
            return {"allocated_gb": torch.cuda.memory_allocated() / 1024 ** 3, "reserved_gb": torch.cuda.memory_reserved() / 1024 ** 3}
        return {"allocated_gb": 0, "reserved_gb": 0}
