"""ORACLE (test infrastructure only — never imported by the product path).

Restatement of diffusers' DDIMScheduler / LCMScheduler as Marigold drives them
(reference marigold/marigold_depth_pipeline.py:423-424 `set_timesteps`/`timesteps`, :466-468
`step(...).prev_sample`; config reads at :349,362). Follows SURVEY.md App. A.3-A.5.

PARITY UNPINNED against diffusers (absent). Known-answer checks that ARE possible offline and are
asserted in tests/test_oracle.py: trailing timestep lists for n in {1,4,10,50}; the LCM list
[999,759,499,259]; alpha_bar_999 == 0 under rescale_betas_zero_snr; t=999 v-prediction => x0 = -v.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SchedulerConfig:
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    beta_schedule: str = "scaled_linear"
    prediction_type: str = "v_prediction"
    timestep_spacing: str = "trailing"
    rescale_betas_zero_snr: bool = True
    set_alpha_to_one: bool = False
    steps_offset: int = 1
    # LCM only
    original_inference_steps: int = 50
    timestep_scaling: float = 10.0


def _betas(cfg: SchedulerConfig) -> torch.Tensor:
    assert cfg.beta_schedule == "scaled_linear"
    betas = torch.linspace(cfg.beta_start ** 0.5, cfg.beta_end ** 0.5, cfg.num_train_timesteps,
                           dtype=torch.float32) ** 2
    if cfg.rescale_betas_zero_snr:
        alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
        s0, sT = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
        alphas_bar_sqrt = (alphas_bar_sqrt - sT) * s0 / (s0 - sT)
        alphas_bar = alphas_bar_sqrt ** 2
        alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
        betas = 1.0 - alphas
    return betas


class DDIMSchedulerOracle:
    """eta = 0, no clip_sample, no thresholding."""

    def __init__(self, cfg: SchedulerConfig = SchedulerConfig()):
        self.config = cfg
        self.betas = _betas(cfg)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, n: int, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = n
        if self.config.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif self.config.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        else:  # linspace
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output: torch.Tensor, t, sample: torch.Tensor, generator=None, noise=None):
        t = int(t)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].to(sample)
        a_prev = (self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod).to(sample)
        b_t = 1 - a_t
        pt = self.config.prediction_type
        if pt == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif pt == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        else:  # v_prediction
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        direction = (1 - a_prev) ** 0.5 * eps  # eta = 0 => sigma_t = 0
        return a_prev ** 0.5 * x0 + direction


class LCMSchedulerOracle:
    def __init__(self, cfg: SchedulerConfig | None = None):
        if cfg is None:
            cfg = SchedulerConfig(timestep_spacing="leading", rescale_betas_zero_snr=False,
                                  prediction_type="v_prediction")
        self.config = cfg
        self.betas = _betas(cfg)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None
        self._step_index = None

    def set_timesteps(self, n: int, device=None):
        T, k0 = self.config.num_train_timesteps, self.config.original_inference_steps
        k = T // k0
        origin = np.asarray(list(range(1, k0 + 1))) * k - 1
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx].astype(np.int64))
        self.num_inference_steps = n
        self._step_index = 0

    def _scalings(self, t):
        s = t * self.config.timestep_scaling
        sigma_data = 0.5
        c_skip = sigma_data ** 2 / (s ** 2 + sigma_data ** 2)
        c_out = s / (s ** 2 + sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output: torch.Tensor, t, sample: torch.Tensor, generator=None, noise=None):
        i = self._step_index
        t = int(t)
        prev_t = int(self.timesteps[i + 1]) if i + 1 < len(self.timesteps) else t
        a_t = self.alphas_cumprod[t].to(sample)
        a_prev = (self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod).to(sample)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        c_skip, c_out = self._scalings(float(t))
        pt = self.config.prediction_type
        if pt == "epsilon":
            x0 = (sample - b_t.sqrt() * model_output) / a_t.sqrt()
        elif pt == "sample":
            x0 = model_output
        else:
            x0 = a_t.sqrt() * sample - b_t.sqrt() * model_output
        denoised = c_out * x0 + c_skip * sample
        if i != self.num_inference_steps - 1:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = a_prev.sqrt() * denoised + b_prev.sqrt() * noise.to(denoised)
        else:
            prev = denoised
        self._step_index += 1
        return prev
