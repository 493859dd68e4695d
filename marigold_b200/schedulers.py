"""Host-side mirrors of the two schedulers Marigold drives (diffusers DDIMScheduler / LCMScheduler):
`set_timesteps`, `timesteps`, `config`, and — instead of a per-step tensor `step()` — the per-step
coefficients of the update, which the CUDA library fuses into the UNet's conv_out epilogue:

    x_prev = kx[i] * x + kv[i] * model_output + kz[i] * noise_i

Reference call sites: marigold/marigold_depth_pipeline.py:423-424 (set_timesteps / timesteps),
:466-468 (step), :349,362 (config.timestep_spacing / rescale_betas_zero_snr).
Arithmetic follows SURVEY.md App. A.3-A.5, evaluated in float64 and rounded once to float32.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np


@dataclass
class SchedulerOutput:
    prev_sample: object


def _alphas_cumprod(cfg) -> np.ndarray:
    # float32 like diffusers (torch.linspace(..., dtype=float32) [** 2]), then cumprod in float32
    if cfg.beta_schedule == "scaled_linear":
        betas = np.linspace(np.float32(cfg.beta_start) ** np.float32(0.5), np.float32(cfg.beta_end) ** np.float32(0.5),
                            cfg.num_train_timesteps, dtype=np.float32) ** 2
    elif cfg.beta_schedule == "linear":
        betas = np.linspace(np.float32(cfg.beta_start), np.float32(cfg.beta_end), cfg.num_train_timesteps, dtype=np.float32)
    else:
        raise RuntimeError(f"Unsupported beta_schedule: {cfg.beta_schedule}")
    if cfg.rescale_betas_zero_snr:
        ab_sqrt = np.sqrt(np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32))
        s0, sT = ab_sqrt[0].copy(), ab_sqrt[-1].copy()
        ab_sqrt = (ab_sqrt - sT) * (s0 / (s0 - sT))
        ab = ab_sqrt ** 2
        alphas = np.concatenate([ab[0:1], ab[1:] / ab[:-1]]).astype(np.float32)
        betas = (1.0 - alphas).astype(np.float32)
    return np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)


class _Base:
    def __init__(self, **kw):
        cfg = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                   prediction_type="v_prediction", timestep_spacing="trailing", rescale_betas_zero_snr=True,
                   set_alpha_to_one=False, steps_offset=1, original_inference_steps=50, timestep_scaling=10.0)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        self.alphas_cumprod = _alphas_cumprod(self.config)
        self.final_alpha_cumprod = np.float32(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None

    def _x0_eps_coeffs(self, a_t: float):
        """(x0, eps) as linear maps of (x, model_output): x0 = px*x + pv*v ; eps = ex*x + ev*v."""
        b_t = 1.0 - a_t
        pt = self.config.prediction_type
        if pt == "v_prediction":
            return (a_t ** 0.5, -(b_t ** 0.5)), (b_t ** 0.5, a_t ** 0.5)
        if pt == "epsilon":
            if a_t == 0.0:
                raise RuntimeError("epsilon prediction is undefined at alpha_bar = 0 (zero-terminal-SNR schedule)")
            return (1.0 / a_t ** 0.5, -(b_t ** 0.5) / a_t ** 0.5), (0.0, 1.0)
        if pt == "sample":
            return (0.0, 1.0), (1.0 / b_t ** 0.5, -(a_t ** 0.5) / b_t ** 0.5)
        raise RuntimeError(f"Unsupported prediction_type: {pt}")


class DDIMScheduler(_Base):
    """eta = 0; no clip_sample / thresholding (Marigold's configuration)."""

    def set_timesteps(self, num_inference_steps: int, device=None):
        T = self.config.num_train_timesteps
        n = int(num_inference_steps)
        sp = self.config.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        elif sp == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        else:
            raise RuntimeError(f"Unsupported timestep_spacing: {sp}")
        self.timesteps = ts
        self.num_inference_steps = n

    def coefficients(self):
        n = self.num_inference_steps
        kx, kv, kz = np.zeros(n), np.zeros(n), np.zeros(n)
        for i, t in enumerate(self.timesteps):
            prev_t = int(t) - self.config.num_train_timesteps // n   # integer floor; NOT the next list entry
            a_t = float(self.alphas_cumprod[int(t)])
            a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
            (px, pv), (ex, ev) = self._x0_eps_coeffs(a_t)
            kx[i] = a_p ** 0.5 * px + (1.0 - a_p) ** 0.5 * ex
            kv[i] = a_p ** 0.5 * pv + (1.0 - a_p) ** 0.5 * ev
        return kx.astype(np.float32), kv.astype(np.float32), kz.astype(np.float32)


class LCMScheduler(_Base):
    def __init__(self, **kw):
        d = dict(timestep_spacing="leading", rescale_betas_zero_snr=False)
        d.update(kw)
        super().__init__(**d)

    def set_timesteps(self, num_inference_steps: int, device=None):
        T, k0 = self.config.num_train_timesteps, self.config.original_inference_steps
        n = int(num_inference_steps)
        k = T // k0
        origin = (np.arange(1, k0 + 1) * k - 1)[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=n, endpoint=False)).astype(np.int64)
        self.timesteps = origin[idx].astype(np.int64)
        self.num_inference_steps = n

    def coefficients(self):
        n = self.num_inference_steps
        kx, kv, kz = np.zeros(n), np.zeros(n), np.zeros(n)
        for i, t in enumerate(self.timesteps):
            t = int(t)
            prev_t = int(self.timesteps[i + 1]) if i + 1 < n else t
            a_t = float(self.alphas_cumprod[t])
            a_p = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
            (px, pv), _ = self._x0_eps_coeffs(a_t)
            s = t * self.config.timestep_scaling
            c_skip = 0.25 / (s * s + 0.25)
            c_out = s / (s * s + 0.25) ** 0.5
            dx, dv = c_out * px + c_skip, c_out * pv          # denoised = dx*x + dv*v
            if i != n - 1:
                kx[i], kv[i], kz[i] = a_p ** 0.5 * dx, a_p ** 0.5 * dv, (1.0 - a_p) ** 0.5
            else:
                kx[i], kv[i], kz[i] = dx, dv, 0.0
        return kx.astype(np.float32), kv.astype(np.float32), kz.astype(np.float32)
