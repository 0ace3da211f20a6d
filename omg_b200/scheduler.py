"""Euler-discrete schedule tables (host side; the per-step arithmetic itself runs inside omg_fuse_step).

Mirrors what the reference obtains from `self.scheduler` (diffusers EulerDiscreteScheduler configured by SDXL-base's
scheduler_config.json) at src/pipelines/lora_pipeline.py:391 (set_timesteps), :397-405 (init_noise_sigma),
:492 (scale_model_input), :615 (step): scaled-linear betas 0.00085..0.012, 1000 train steps, "leading" spacing,
steps_offset 1, epsilon prediction, linearly interpolated sigmas, sigma_T+1 = 0."""
import numpy as np


class EulerDiscreteSchedule:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float32) ** 2
        self.alphas_cumprod = np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.timesteps = None
        self.sigmas = None

    def set_timesteps(self, num_inference_steps: int):
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32) + self.steps_offset
        sig = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = ts
        return ts

    @property
    def init_noise_sigma(self) -> float:
        return float((self.sigmas.max() ** 2 + 1) ** 0.5)

    def input_scale(self, i: int) -> float:
        return float(1.0 / (self.sigmas[i] ** 2 + 1) ** 0.5)
