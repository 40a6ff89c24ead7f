"""Host-side LCMScheduler logic (timestep table and per-step scalars); the arithmetic on latents runs in
``fie_cfg_lcm_step`` / ``fie_vae_sample_add_noise``.  Mirrors diffusers ``LCMScheduler`` as configured by the
reference (``LCMScheduler.from_config(..., timestep_spacing="trailing")``, ``src/pipeline.py:138-141,158-161``):
scaled_linear betas 0.00085..0.012 over 1000 steps, epsilon prediction, original_inference_steps 50,
timestep_scaling 10, sigma_data 0.5."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


class LCMSchedule:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 original_inference_steps: int = 50, timestep_scaling: float = 10.0):
        betas = np.linspace(np.float32(beta_start) ** 0.5, np.float32(beta_end) ** 0.5, num_train_timesteps, dtype=np.float32) ** 2
        self.alphas_cumprod = np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)
        self.num_train_timesteps = num_train_timesteps
        self.original_inference_steps = original_inference_steps
        self.timestep_scaling = timestep_scaling
        self.sigma_data = 0.5
        self.timesteps: List[int] = []

    def set_timesteps(self, n: int) -> List[int]:
        k = self.num_train_timesteps // self.original_inference_steps
        origin = (np.arange(1, self.original_inference_steps + 1) * k - 1)[::-1]
        idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int64)
        self.timesteps = [int(v) for v in origin[idx]]
        return self.timesteps

    def img2img_timesteps(self, n: int, strength: float) -> Tuple[List[int], int]:
        """diffusers img2img ``get_timesteps``: keeps the last int(n*strength) steps; returns (timesteps, begin_index)."""
        self.set_timesteps(n)
        init = min(int(n * strength), n)
        t_start = max(n - init, 0)
        return self.timesteps[t_start:], t_start

    def add_noise_coeffs(self, t: int) -> Tuple[float, float]:
        a = float(self.alphas_cumprod[t])
        return a ** 0.5, (1.0 - a) ** 0.5

    def step_coeffs(self, step_index: int) -> Dict[str, float]:
        t = self.timesteps[step_index]
        last = step_index == len(self.timesteps) - 1
        prev_t = t if last else self.timesteps[step_index + 1]
        a_t, a_prev = float(self.alphas_cumprod[t]), float(self.alphas_cumprod[prev_t])
        s = t * self.timestep_scaling
        return dict(t=t, last=last, sqrt_a=a_t ** 0.5, sqrt_1ma=(1 - a_t) ** 0.5,
                    c_skip=self.sigma_data ** 2 / (s ** 2 + self.sigma_data ** 2), c_out=s / (s ** 2 + self.sigma_data ** 2) ** 0.5,
                    sqrt_a_prev=a_prev ** 0.5, sqrt_1ma_prev=(1 - a_prev) ** 0.5)
