"""diffusers==0.25.0 EulerDiscreteScheduler restated for the SDXL-base scheduler_config.json [third-party; not under
/root/reference]: scaled_linear betas 0.00085 -> 0.012, 1000 train steps, timestep_spacing="leading",
steps_offset=1, epsilon prediction, linear sigma interpolation, no Karras sigmas.  TEST INFRASTRUCTURE.
Reference call sites: src/pipelines/lora_pipeline.py:391 (set_timesteps), :492 (scale_model_input), :615 (step).
"""
import numpy as np
import torch


class EulerDiscrete:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.timesteps = None
        self.sigmas = None

    def set_timesteps(self, n: int):
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.float32) + self.steps_offset
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        return self.timesteps

    @property
    def init_noise_sigma(self):
        return float((self.sigmas.max() ** 2 + 1) ** 0.5)  # "leading" spacing

    def scale_model_input(self, sample, i):
        return sample / ((self.sigmas[i] ** 2 + 1) ** 0.5)

    def step(self, model_output, i, sample):
        out_dtype = model_output.dtype
        sample = sample.to(torch.float32)
        sigma = self.sigmas[i]
        pred_original = sample - sigma * model_output.to(torch.float32)
        derivative = (sample - pred_original) / sigma
        dt = self.sigmas[i + 1] - sigma
        return (sample + derivative * dt).to(out_dtype)
