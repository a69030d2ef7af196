"""EulerDiscreteScheduler with the SVD scheduler_config — the part of diffusers==0.25.1 the reference pipelines
call (svd/pipeline_stable_video_diffusion_controlnet.py:583, :589 init_noise_sigma, :632, :709): Karras sigmas
(sigma_min 0.002, sigma_max 700, rho 7), continuous timesteps 0.25*ln(sigma), v-prediction Euler step in fp32.
Host-side scalar math only; the tensor update itself is the fused CUDA step in this_and_that_vdm_b200/sampler.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch


@dataclass
class EulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "v_prediction",
                 interpolation_type: str = "linear", use_karras_sigmas: bool = True, sigma_min: float = 0.002,
                 sigma_max: float = 700.0, timestep_spacing: str = "leading", timestep_type: str = "continuous",
                 steps_offset: int = 1, rescale_betas_zero_snr: bool = False):
        if prediction_type != "v_prediction" or not use_karras_sigmas or timestep_type != "continuous":
            raise NotImplementedError("only the SVD scheduler configuration is implemented")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, prediction_type=prediction_type, interpolation_type=interpolation_type,
            use_karras_sigmas=use_karras_sigmas, sigma_min=sigma_min, sigma_max=sigma_max,
            timestep_spacing=timestep_spacing, timestep_type=timestep_type, steps_offset=steps_offset)
        self.sigmas = None
        self.timesteps = None
        self.num_inference_steps = None
        self._step_index = None
        self.set_timesteps(25)

    @property
    def init_noise_sigma(self) -> float:
        max_sigma = float(self.sigmas.max())
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv, max_inv = self.config.sigma_min ** (1 / rho), self.config.sigma_max ** (1 / rho)
        sigmas = (max_inv + ramp * (min_inv - max_inv)) ** rho
        timesteps = np.array([0.25 * np.log(s) for s in sigmas])
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to(device=device)
        self.timesteps = torch.from_numpy(timesteps.astype(np.float32)).to(device=device)
        self._step_index = None

    def _init_step_index(self, timestep) -> None:
        t = float(timestep)
        idx = (self.timesteps.cpu() - t).abs().argmin().item()
        self._step_index = int(idx)

    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self.sigmas[self._step_index])
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self.sigmas[self._step_index])
        sigma_next = float(self.sigmas[self._step_index + 1])
        x = sample.to(torch.float32)
        x0 = model_output.to(torch.float32) * (-sigma / (sigma ** 2 + 1) ** 0.5) + x / (sigma ** 2 + 1)
        d = (x - x0) / sigma
        prev = (x + d * (sigma_next - sigma)).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return EulerDiscreteSchedulerOutput(prev_sample=prev, pred_original_sample=x0)
