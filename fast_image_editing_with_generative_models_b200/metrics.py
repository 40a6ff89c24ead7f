"""``MetricsCalculator`` / ``DinoDistanceMetric`` on the fie_b200 kernels — SURVEY 8(f)-4, the drop-in for reference
``src/metrics.py:24-386`` (same class names, method names, argument meaning and return types: Python floats).

| metric | reference (file:line) | here |
|---|---|---|
| SSIM | torchmetrics ``StructuralSimilarityIndexMeasure(data_range=1.0)`` on 512^2 Lanczos copies (``:175-177,214-237``) | ``fie_resample_lanczos_u8`` (bit-identical to Pillow) + ``fie_ssim_u8`` |
| PSNR / MSE | ``PeakSignalNoiseRatio(data_range=1.0)`` / ``MeanSquaredError`` (``:190-197,285-336``) | ``fie_sqdiff_u8``: exact integer sum of squared byte differences |
| LPIPS | ``LearnedPerceptualImagePatchSimilarity(net_type='squeeze')`` (``:180-182,239-262``) | :mod:`.lpips` |
| CLIP score | ``CLIPScore("openai/clip-vit-base-patch16")`` (``:185-187,264-283``) | Pillow-bicubic resize + centre crop, :mod:`.vit` image tower, :mod:`.text_encoder` text tower, ``fie_cosine_rows_f16`` |
| DINO distance | ``dino_vitb8`` keys of block 11, cosine self-similarity, MSE (``:24-148``) | ``fie_resample_f32`` (antialiased bilinear + normalise), :mod:`.vit`, ``fie_l2norm_rows_f16`` + one GEMM per image, ``fie_sqdiff_f32`` |

Weights: ``checkpoints={"clip": <transformers CLIPModel folder>, "dino": <dino_vitbase8_pretrain.pth>, "squeezenet": <torchvision
squeezenet1_1 .pth>, "lpips": <lpips squeeze.pth>}`` or ``FIE_METRIC_CHECKPOINTS=<folder holding clip-vit-base-patch16/,
dino_vitbase8_pretrain.pth, squeezenet1_1.pth, lpips_squeeze.pth>``.  No checkpoint exists offline: a network without one is built
with seeded random weights and a loud warning — its score is then a structural test value, not a quality measurement (SSIM / PSNR /
MSE need no weights).  There is no CPU path: the calculator needs a CUDA device."""
from __future__ import annotations

import math
import os
import warnings
from typing import Dict, Optional

import numpy as np
import torch

from . import lpips as lpips_mod
from . import ops, vit
from .text_encoder import CLIPTextConfig, CLIPTextEncoder, make_clip_params, pseudo_token_ids

METRIC_SIZE = 512            # "PIE-Bench resolution (512x512) for fair comparison", src/metrics.py:225-230


def clip_b16_text_config() -> CLIPTextConfig:        # openai/clip-vit-base-patch16 text tower
    return CLIPTextConfig(name="clip-vit-b16-text", hidden_size=512, num_layers=12, num_heads=8, intermediate_size=2048, projection_dim=512, seed=44)


def metric_checkpoints_from_env() -> Dict[str, str]:
    root = os.environ.get("FIE_METRIC_CHECKPOINTS")
    if not root:
        return {}
    cand = {"clip": "clip-vit-base-patch16", "dino": "dino_vitbase8_pretrain.pth", "squeezenet": "squeezenet1_1.pth", "lpips": "lpips_squeeze.pth"}
    return {k: os.path.join(root, v) for k, v in cand.items() if os.path.exists(os.path.join(root, v))}


def load_clip_model_dir(folder: str):
    """transformers ``CLIPModel`` folder (``config.json`` with ``text_config`` / ``vision_config`` + ``model.safetensors``)
    -> (ViTConfig, CLIPTextConfig, {name: fp32 CPU tensor})."""
    import json
    from safetensors.torch import load_file
    with open(os.path.join(folder, "config.json")) as f:
        d = json.load(f)
    v, t = d.get("vision_config") or {}, d.get("text_config") or {}
    proj = d.get("projection_dim", 512)
    vcfg = vit.ViTConfig(name="clip-vision", image_size=v.get("image_size", 224), patch_size=v.get("patch_size", 16), hidden_size=v.get("hidden_size", 768),
                         num_layers=v.get("num_hidden_layers", 12), num_heads=v.get("num_attention_heads", 12), intermediate_size=v.get("intermediate_size", 3072),
                         hidden_act=v.get("hidden_act", "quick_gelu"), layer_norm_eps=v.get("layer_norm_eps", 1e-5), projection_dim=proj)
    tcfg = CLIPTextConfig(name="clip-text", vocab_size=t.get("vocab_size", 49408), hidden_size=t.get("hidden_size", 512), num_layers=t.get("num_hidden_layers", 12),
                          num_heads=t.get("num_attention_heads", 8), intermediate_size=t.get("intermediate_size", 2048),
                          max_positions=t.get("max_position_embeddings", 77), hidden_act=t.get("hidden_act", "quick_gelu"),
                          layer_norm_eps=t.get("layer_norm_eps", 1e-5), projection_dim=proj)
    for fn in ("model.safetensors", "model.fp16.safetensors"):
        if os.path.exists(os.path.join(folder, fn)):
            sd = {k: x.float() for k, x in load_file(os.path.join(folder, fn)).items() if "position_ids" not in k}
            return vcfg, tcfg, sd
    raise FileNotFoundError(f"no model[.fp16].safetensors in {folder}")


def _as_u8_nhwc(img, dev) -> torch.Tensor:
    """PIL image / HWC or CHW uint8 array or tensor (or float in [0, 1]) -> uint8 CUDA [1,H,W,3]."""
    if isinstance(img, torch.Tensor):
        t = img.detach()
        if t.dim() == 4:
            t = t[0]
        if t.shape[0] == 3 and t.shape[-1] != 3:
            t = t.permute(1, 2, 0)
        if t.dtype != torch.uint8:
            t = t.float()
            t = (t * 255.0 if float(t.max()) <= 1.0 else t).round().clamp(0, 255).to(torch.uint8)
        arr = t.contiguous()
    else:
        if hasattr(img, "convert"):
            img = img.convert("RGB")
        arr = torch.from_numpy(np.array(img, dtype=np.uint8))                  # (a copy: PIL hands out read-only buffers)
    if arr.dim() != 3 or arr.shape[-1] != 3:
        raise ValueError(f"metrics: RGB image expected, got shape {tuple(arr.shape)}")
    if not arr.is_cuda:
        arr = arr.pin_memory().to(dev, non_blocking=True)
    return arr.unsqueeze(0).contiguous()


class DinoDistanceMetric:
    """DINO-based structural distance (reference ``src/metrics.py:113-148``): MSE between the cosine self-similarity maps of the block-
    ``layer`` keys of the source and the edited image."""

    def __init__(self, device: str, model_name: str = "dino_vitb8", resize_to: int = 224, layer: int = 11, *, checkpoint: Optional[str] = None,
                 params: Optional[Dict[str, torch.Tensor]] = None, config: Optional[vit.ViTConfig] = None):
        if model_name != "dino_vitb8" and config is None and params is None and checkpoint is None:
            raise ValueError("DinoDistanceMetric: without a checkpoint only dino_vitb8 is configured (pass checkpoint=, params= or config=)")
        self.device = torch.device("cuda:0" if str(device) == "cuda" else device)
        self.layer, self.resize_to = layer, resize_to
        self.synthetic_weights = params is None and checkpoint is None
        if params is None and checkpoint is not None:
            params = {k: v.float() for k, v in torch.load(checkpoint, map_location="cpu", weights_only=True).items()}
        cfg = config or (vit.dino_config_from_params(params, model_name) if params is not None else vit.dino_vitb8_config())
        if params is None:
            warnings.warn("DinoDistanceMetric: no DINO checkpoint given — the ViT runs on seeded RANDOM weights; the distance is a structural "
                          "test value, not the published metric", RuntimeWarning, stacklevel=2)
            params = vit.make_vit_params(cfg)
        with torch.cuda.device(self.device):
            self.model = vit.VisionTransformer(params, cfg, self.device)

    def _preprocess(self, img_u8: torch.Tensor) -> torch.Tensor:
        """``_to_tensor`` (reference src/metrics.py:124-136): /255, Resize(resize_to, antialias=True), ImageNet normalisation -> fp32 [n,h,w,3]."""
        n, h, w, _ = img_u8.shape
        s = self.resize_to                                                      # transforms.Resize(int): the shorter side becomes `s`
        oh, ow = (s, int(s * w / h)) if h <= w else (int(s * h / w), s)
        return ops.resize_aa_normalize(img_u8, oh, ow, vit.IMAGENET_MEAN, vit.IMAGENET_STD)

    def _self_similarity(self, img_u8: torch.Tensor, preprocessed: bool = False):
        """-> ([fp32 [T, T_padded] cosine self-similarity of the block-``layer`` keys, one per image], T)."""
        x = img_u8 if preprocessed else self._preprocess(img_u8)
        n = x.shape[0]
        keys = self.model.keys(x, self.layer)                                   # [n*T, C] view
        t = keys.shape[0] // n
        tp = (t + 31) // 32 * 32
        sims = []
        for i in range(n):
            kn = torch.zeros((tp, keys.shape[1]), dtype=torch.float16, device=keys.device)      # zero rows pad the GEMM's N to a multiple of 32
            ops.l2norm_rows(keys[i * t:(i + 1) * t], out=kn[:t])
            sims.append(ops.gemm(kn[:t], kn, out_f32=True))                     # [T, tp] fp32 cosine similarities
        return sims, t

    def _distance_dev(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """uint8 CUDA [1,H,W,3] pair -> fp64 [1] distance on the device (no host synchronisation)."""
        xa, xb = self._preprocess(a), self._preprocess(b)
        if xa.shape == xb.shape:                                 # one ViT pass over both images (rows of a batch are computed independently)
            (sa, sb), t = self._self_similarity(torch.cat([xa, xb], 0), preprocessed=True)
        else:
            (sa,), t = self._self_similarity(xa, preprocessed=True)
            (sb,), tb = self._self_similarity(xb, preprocessed=True)
            if t != tb:
                raise ValueError("DinoDistanceMetric: source and edited image give different token counts")
        return ops.sqdiff_f32(sb, sa, t, t) / float(t * t)

    def calculate_distance(self, source_img, edited_img) -> float:
        with torch.cuda.device(self.device), torch.no_grad():
            return float(self._distance_dev(_as_u8_nhwc(source_img, self.device), _as_u8_nhwc(edited_img, self.device)).item())


class MetricsCalculator:
    """Image quality and editing metrics (reference ``src/metrics.py:150-386``): SSIM, LPIPS, CLIP score, PSNR, MSE, DINO distance."""

    def __init__(self, device="cuda", *, checkpoints: Optional[Dict[str, str]] = None, networks: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("MetricsCalculator: the B200 path has no CPU fallback (a CUDA device is required)")
        self.device = torch.device("cuda:0" if str(device) == "cuda" else device)
        if self.device.type != "cuda":
            raise RuntimeError(f"MetricsCalculator: device {device!r} is not a CUDA device (no CPU fallback)")
        ck = dict(metric_checkpoints_from_env())
        ck.update(checkpoints or {})
        self.checkpoints = ck
        self.synthetic, self._notes = [], []
        self.lpips_net = self.clip_vision = self.clip_text = self.dino_metric = self._tokenizer = None
        print(f"[MetricsCalculator] Initializing on {self.device}...")
        if networks:
            with torch.cuda.device(self.device), warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)                  # one combined warning below
                self._build_networks(ck)
            for note in self._notes:
                warnings.warn(note, RuntimeWarning, stacklevel=2)
            if self.synthetic:
                warnings.warn(f"MetricsCalculator: no checkpoint for {', '.join(self.synthetic)} — those networks run on seeded RANDOM weights; "
                              "their scores are structural test values, not quality measurements (set FIE_METRIC_CHECKPOINTS or checkpoints=)",
                              RuntimeWarning, stacklevel=2)
        print("[MetricsCalculator] Initialization complete!")

    def _build_networks(self, ck: Dict[str, str]):
        dev = self.device
        # LPIPS: torchvision squeezenet1_1 backbone + lpips "lin" layers
        if "squeezenet" in ck and "lpips" in ck:
            p = {k: v.float() for k, v in torch.load(ck["squeezenet"], map_location="cpu", weights_only=True).items()}
            p.update({k: v.float() for k, v in torch.load(ck["lpips"], map_location="cpu", weights_only=True).items()})
        else:
            p = lpips_mod.make_lpips_params()
            self.synthetic.append("LPIPS (SqueezeNet 1.1)")
        self.lpips_net = lpips_mod.LPIPSSqueeze(p, dev)
        # CLIP ViT-B/16 (both towers) + tokenizer
        if "clip" in ck:
            vcfg, tcfg, sd = load_clip_model_dir(ck["clip"])
            if os.path.exists(os.path.join(ck["clip"], "vocab.json")):
                from .tokenizer import CLIPBPETokenizer
                self._tokenizer = CLIPBPETokenizer.from_files(ck["clip"])
            else:
                self._notes.append("MetricsCalculator: CLIP checkpoint without vocab.json / merges.txt; prompts are mapped to pseudo token ids")
        else:
            vcfg, tcfg = vit.clip_b16_vision_config(), clip_b16_text_config()
            sd = dict(vit.make_vit_params(vcfg))
            sd.update(make_clip_params(tcfg))
            self.synthetic.append("CLIP ViT-B/16")
        self.clip_vision = vit.VisionTransformer(sd, vcfg, dev)
        self.clip_text = CLIPTextEncoder(sd, tcfg, dev)
        self._text_vocab = tcfg.vocab_size
        # DINO ViT-B/8
        if "dino" not in ck:
            self.synthetic.append("DINO ViT-B/8")
        self.dino_metric = DinoDistanceMetric(device=dev, checkpoint=ck.get("dino"))

    # ---- helpers -----------------------------------------------------------------------------------------------------------
    def _pil_to_tensor(self, img) -> torch.Tensor:
        """-> uint8 CUDA [1,H,W,3] (the reference's float [1,3,H,W] / 255 is applied inside the kernels)."""
        return _as_u8_nhwc(img, self.device)

    def to_metric_size(self, img) -> torch.Tensor:
        """The 512^2 Lanczos copy ``evaluate.py:127-130`` makes of every image before scoring it, as a uint8 CUDA tensor."""
        with torch.cuda.device(self.device):
            return self._at_metric_size(img)

    def _at_metric_size(self, img) -> torch.Tensor:
        t = self._pil_to_tensor(img)
        if t.shape[1] != METRIC_SIZE or t.shape[2] != METRIC_SIZE:
            t = ops.resize_lanczos(t, METRIC_SIZE, METRIC_SIZE)               # == img.resize((512, 512), Image.LANCZOS), bit for bit
        return t

    def _need(self, net, name: str):
        if net is None:
            raise RuntimeError(f"MetricsCalculator(networks=False) has no {name} network")
        return net

    # ---- device-side halves: uint8 CUDA tensors in, device scalars out (no host synchronisation) -----------------------------------
    def _ssim_dev(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return ops.ssim_u8(a, b)

    def _lpips_dev(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return self._need(self.lpips_net, "LPIPS").distance(a, b)

    def clip_image_input(self, img) -> torch.Tensor:
        """The CLIP image processor on one image: Pillow-bicubic resize of the shorter side to the tower's input size, centre crop -> uint8
        CUDA [1,S,S,3] (rescale by 1/255 and mean / std normalisation happen in ``fie_patchify_f16``)."""
        t = self._pil_to_tensor(img)
        s = self._need(self.clip_vision, "CLIP").cfg.image_size
        _, h, w, _ = t.shape
        oh, ow = (s, int(s * w / h)) if h <= w else (int(s * h / w), s)
        if (oh, ow) != (h, w):
            t = ops.resize_pillow(t, oh, ow, "bicubic")
        if (oh, ow) != (s, s):
            top, left = (oh - s) // 2, (ow - s) // 2
            t = t[:, top:top + s, left:left + s].contiguous()
        return t

    def _clip_cosine_dev(self, img, text) -> torch.Tensor:
        x = self.clip_image_input(img)
        img_emb = self.clip_vision.embed(x, vit.CLIP_MEAN, vit.CLIP_STD)
        if self._tokenizer is not None:
            ids = torch.tensor(self._tokenizer([text]), dtype=torch.int64)
        else:
            ids = pseudo_token_ids(text, self._text_vocab).unsqueeze(0)
        _, _, txt_emb = self._need(self.clip_text, "CLIP").forward(ids)
        return ops.cosine_rows(img_emb, txt_emb)

    @staticmethod
    def _mse_psnr(sq_sum: int, count: int):
        mse = sq_sum / (255.0 * 255.0 * count)
        return mse, (float("inf") if mse == 0.0 else 10.0 * math.log10(1.0 / mse))

    # ---- the reference's public methods --------------------------------------------------------------------------------------
    def calculate_ssim(self, img1, img2) -> float:
        with torch.cuda.device(self.device):
            return float(self._ssim_dev(self._at_metric_size(img1), self._at_metric_size(img2)).item())

    def calculate_lpips(self, img1, img2) -> float:
        with torch.cuda.device(self.device):
            return float(self._lpips_dev(self._at_metric_size(img1), self._at_metric_size(img2)).item())

    def calculate_mse(self, img1, img2) -> float:
        with torch.cuda.device(self.device):
            a, b = self._at_metric_size(img1), self._at_metric_size(img2)
            return self._mse_psnr(int(ops.sqdiff_u8(a, b).item()), a.numel())[0]

    def calculate_psnr(self, img1, img2) -> float:
        with torch.cuda.device(self.device):
            a, b = self._at_metric_size(img1), self._at_metric_size(img2)
            return self._mse_psnr(int(ops.sqdiff_u8(a, b).item()), a.numel())[1]

    def clip_cosine(self, img, text) -> float:
        """cos(image features, text features) of the CLIP towers (CLIPScore before the x100 and the floor at 0)."""
        with torch.cuda.device(self.device), torch.no_grad():
            return float(self._clip_cosine_dev(img, text).item())

    def calculate_clip_score(self, img, text) -> float:
        return max(100.0 * self.clip_cosine(img, text), 0.0)

    def calculate_all_metrics(self, source_img, edited_img, prompt) -> Dict[str, float]:
        """All six metrics of one (source, edited, prompt) triple — reference ``src/metrics.py:338-381`` — with each image uploaded once,
        the 512^2 copies made once, every kernel queued before the first result is read back (one host synchronisation per pair)."""
        with torch.cuda.device(self.device), torch.no_grad():
            src, edt = self._pil_to_tensor(source_img), self._pil_to_tensor(edited_img)
            a, b = self._at_metric_size(src), self._at_metric_size(edt)
            ssim = self._ssim_dev(a, b)
            lp = self._lpips_dev(a, b)
            cos = self._clip_cosine_dev(edt, prompt)
            sq = ops.sqdiff_u8(a, b)
            dino = self._need(self.dino_metric, "DINO")._distance_dev(src, edt)
            mse, psnr = self._mse_psnr(int(sq.item()), a.numel())
            return {"ssim": float(ssim.item()), "lpips": float(lp.item()), "clip_score": max(100.0 * float(cos.item()), 0.0), "psnr": psnr, "mse": mse,
                    "dino_distance": float(dino.item())}

    def clear_memory(self):
        """Clear GPU memory cache."""
        with torch.cuda.device(self.device):
            torch.cuda.empty_cache()
