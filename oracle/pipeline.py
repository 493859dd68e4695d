"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the control flow of
  MarigoldDepthPipeline.__call__ / single_infer / encode_rgb / decode_depth
      (marigold/marigold_depth_pipeline.py:155-338, 397-477, 479-496, 498-516)
  MarigoldNormalsPipeline.__call__ / single_infer / decode_normals
      (marigold/marigold_normals_pipeline.py:140-308, 362-442, 463-479)
  MarigoldIIDPipeline.__call__ / single_infer / decode_targets / fill_outputs          
      (marigold/marigold_iid_pipeline.py:239-411, 467-547, 568-585)
driving the restated networks in oracle/unet.py, oracle/vae.py and schedulers in oracle/schedulers.py.

The reference pipelines subclass diffusers.DiffusionPipeline and cannot be imported here
(`import marigold` -> ModuleNotFoundError: diffusers), hence the restatement. Two deliberate
additions, both keyword-only and both needed for any cross-device parity:
  noise       [E,4,h,w]      explicit initial latents instead of torch.randn(generator)   (:430-435)
  step_noise  [n-1,E,4,h,w]  explicit per-step noise for LCMScheduler.step
PARITY UNPINNED for the network part (see oracle/unet.py); the ensemble part is pinned (oracle/ensemble.py).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from .ensemble import ensemble_depth, ensemble_iid, ensemble_normals
from .schedulers import DDIMSchedulerOracle, LCMSchedulerOracle


def resize_max_res(img: torch.Tensor, max_edge_resolution: int, mode: str = "bilinear") -> torch.Tensor:
    """marigold/util/image_util.py:90-120 (int() truncation :116-117, antialias :119)."""
    assert img.dim() == 4
    h, w = img.shape[-2:]
    f = min(max_edge_resolution / w, max_edge_resolution / h)
    nw, nh = int(w * f), int(h * f)
    return _resize(img, (nh, nw), mode)


def _resize(img, size, mode):
    if mode in ("nearest", "nearest-exact"):
        return F.interpolate(img.float(), size=size, mode="nearest-exact").to(img.dtype)
    out = F.interpolate(img.float(), size=size, mode=mode, antialias=True, align_corners=False)
    if img.dtype == torch.uint8:
        out = out.round().clamp(0, 255).to(torch.uint8)
    return out


class _OraclePipelineBase:
    latent_scale_factor = 0.18215  # marigold_depth_pipeline.py:118

    def __init__(self, unet, vae, scheduler, empty_text_embed: torch.Tensor,
                 default_denoising_steps: int = 4, default_processing_resolution: int = 768):
        self.unet, self.vae, self.scheduler = unet, vae, scheduler
        self.empty_text_embed = empty_text_embed  # [1, 2, cross_dim]
        self.default_denoising_steps = default_denoising_steps
        self.default_processing_resolution = default_processing_resolution
        self.dtype = torch.float32

    @torch.no_grad()
    def encode_rgb(self, rgb_in):
        h = self.vae.encoder(rgb_in)
        moments = self.vae.quant_conv(h)
        mean, _logvar = torch.chunk(moments, 2, dim=1)
        return mean * self.latent_scale_factor

    @torch.no_grad()
    def denoise(self, rgb_latent, target_latent, n_steps, step_noise=None, trace=None):
        self.scheduler.set_timesteps(n_steps)
        ctx = self.empty_text_embed.repeat(rgb_latent.shape[0], 1, 1)
        for i, t in enumerate(self.scheduler.timesteps):
            unet_input = torch.cat([rgb_latent, target_latent], dim=1)  # rgb first (:456-458)
            pred = self.unet(unet_input, t, encoder_hidden_states=ctx)
            nz = None
            if isinstance(self.scheduler, LCMSchedulerOracle) and i < n_steps - 1:
                assert step_noise is not None, "LCM needs explicit per-step noise for parity"
                nz = step_noise[i]
            target_latent = self.scheduler.step(pred, t, target_latent, noise=nz)
            if trace is not None:
                trace.append((pred.clone(), target_latent.clone()))
        return target_latent

    def _decode_raw(self, latent):
        z = self.vae.post_quant_conv(latent / self.latent_scale_factor)
        return self.vae.decoder(z)

    def _preprocess(self, input_image, processing_res, resample_method):
        if not isinstance(input_image, torch.Tensor):
            raise TypeError(f"Unknown input type: {type(input_image) = }")
        rgb = input_image
        assert rgb.dim() == 4 and rgb.shape[-3] == 3, f"Wrong input shape {rgb.shape}, expected [1, rgb, H, W]"
        input_size = rgb.shape
        if processing_res > 0:
            rgb = resize_max_res(rgb, processing_res, resample_method)
        rgb_norm = (rgb / 255.0 * 2.0 - 1.0).to(self.dtype)
        assert rgb_norm.min() >= -1.0 and rgb_norm.max() <= 1.0
        return rgb_norm, input_size


class OracleDepthPipeline(_OraclePipelineBase):
    scale_invariant = True
    shift_invariant = True

    @torch.no_grad()
    def single_infer(self, rgb_in, n_steps, noise, step_noise=None):
        rgb_latent = self.encode_rgb(rgb_in)
        target = self.denoise(rgb_latent, noise.to(rgb_latent), n_steps, step_noise)
        depth = self._decode_raw(target).mean(dim=1, keepdim=True)  # :515
        depth = torch.clip(depth, -1.0, 1.0)                         # :473
        return (depth + 1.0) / 2.0                                   # :475

    @torch.no_grad()
    def __call__(self, input_image, denoising_steps: Optional[int] = None, ensemble_size: int = 1,
                 processing_res: Optional[int] = None, match_input_res: bool = True,
                 resample_method: str = "bilinear", batch_size: int = 0, ensemble_kwargs=None, *,
                 noise: torch.Tensor, step_noise: Optional[torch.Tensor] = None):
        if denoising_steps is None:
            denoising_steps = self.default_denoising_steps
        if processing_res is None:
            processing_res = self.default_processing_resolution
        assert processing_res >= 0 and ensemble_size >= 1 and denoising_steps >= 1
        rgb_norm, input_size = self._preprocess(input_image, processing_res, resample_method)
        bs = batch_size if batch_size > 0 else 1
        preds = []
        for s in range(0, ensemble_size, bs):
            e = min(ensemble_size, s + bs)
            sn = step_noise[:, s:e] if step_noise is not None else None
            preds.append(self.single_infer(rgb_norm.expand(e - s, -1, -1, -1), denoising_steps, noise[s:e], sn))
        target_preds = torch.cat(preds, dim=0)
        if ensemble_size > 1:
            final_pred, uncert = ensemble_depth(target_preds, scale_invariant=self.scale_invariant,
                                                shift_invariant=self.shift_invariant, **(ensemble_kwargs or {}))
        else:
            final_pred, uncert = target_preds, None
        if match_input_res:
            final_pred = _resize(final_pred, tuple(input_size[-2:]), resample_method)
        final_pred = final_pred.squeeze().cpu().numpy().clip(0, 1)
        if uncert is not None:
            uncert = uncert.squeeze().cpu().numpy()
        return final_pred, uncert, target_preds


class OracleNormalsPipeline(_OraclePipelineBase):
    @torch.no_grad()
    def single_infer(self, rgb_in, n_steps, noise, step_noise=None):
        rgb_latent = self.encode_rgb(rgb_in)
        target = self.denoise(rgb_latent, noise.to(rgb_latent), n_steps, step_noise)
        normals = torch.clip(self._decode_raw(target), -1.0, 1.0)    # :438
        norm = torch.norm(normals, dim=1, keepdim=True)
        return normals / norm.clamp(min=1e-6)                        # :439-440

    @torch.no_grad()
    def __call__(self, input_image, denoising_steps: Optional[int] = None, ensemble_size: int = 1,
                 processing_res: Optional[int] = None, match_input_res: bool = True,
                 resample_method: str = "bilinear", batch_size: int = 0, ensemble_kwargs=None, *,
                 noise: torch.Tensor):
        if isinstance(self.scheduler, LCMSchedulerOracle):
            raise RuntimeError("This pipeline implementation does not support the LCMScheduler.")  # :338-342
        if denoising_steps is None:
            denoising_steps = self.default_denoising_steps
        if processing_res is None:
            processing_res = self.default_processing_resolution
        rgb_norm, input_size = self._preprocess(input_image, processing_res, resample_method)
        bs = batch_size if batch_size > 0 else 1
        preds = []
        for s in range(0, ensemble_size, bs):
            e = min(ensemble_size, s + bs)
            preds.append(self.single_infer(rgb_norm.expand(e - s, -1, -1, -1), denoising_steps, noise[s:e]))
        target_preds = torch.cat(preds, dim=0)
        if ensemble_size > 1:
            final_pred, uncert = ensemble_normals(target_preds, **(ensemble_kwargs or {}))
        else:
            final_pred, uncert = target_preds, None
        if match_input_res:
            final_pred = _resize(final_pred, tuple(input_size[-2:]), resample_method)
        final_pred = final_pred.squeeze().cpu().numpy().clip(-1, 1)
        if uncert is not None:
            uncert = uncert.squeeze().cpu().numpy()
        return final_pred, uncert, target_preds


class OracleIIDPipeline(_OraclePipelineBase):
    """Intrinsic image decomposition: n_targets 3-channel maps from ONE UNet whose conv_in takes 4 * (n + 1) latent
    channels and whose conv_out produces 4 * n (marigold_iid_pipeline.py:491-495; src/trainer/marigold_iid_trainer.py:
    203-246). Returns (pred [1 or E, 3n, H, W] in [0,1], uncertainty or None, members [E, 3n, h, w])."""

    def __init__(self, unet, vae, scheduler, empty_text_embed, target_names,
                 default_denoising_steps: int = 4, default_processing_resolution: int = 768):
        super().__init__(unet, vae, scheduler, empty_text_embed, default_denoising_steps, default_processing_resolution)
        self.target_names = list(target_names)
        self.n_targets = len(self.target_names)

    def decode_targets(self, target_latent):
        """:568-585 — one VAE decode per 4-channel chunk, concatenated along channels."""
        outs = [self._decode_raw(target_latent[:, 4 * i: 4 * (i + 1)]) for i in range(self.n_targets)]
        return torch.cat(outs, dim=1)

    @torch.no_grad()
    def single_infer(self, rgb_in, n_steps, noise, step_noise=None):
        rgb_latent = self.encode_rgb(rgb_in)
        assert noise.shape[1] == 4 * self.n_targets, "noise must be [B, 4 * n_targets, h, w] (:493-497)"
        target = self.denoise(rgb_latent, noise.to(rgb_latent), n_steps, step_noise)
        targets = self.decode_targets(target)
        return (torch.clip(targets, -1.0, 1.0) + 1.0) / 2.0          # :543-545

    @torch.no_grad()
    def __call__(self, input_image, denoising_steps: Optional[int] = None, ensemble_size: int = 1,
                 processing_res: Optional[int] = None, match_input_res: bool = True,
                 resample_method: str = "bilinear", batch_size: int = 0, ensemble_kwargs=None, *,
                 noise: torch.Tensor, step_noise: Optional[torch.Tensor] = None):
        if denoising_steps is None:
            denoising_steps = self.default_denoising_steps
        if processing_res is None:
            processing_res = self.default_processing_resolution
        assert processing_res >= 0 and ensemble_size >= 1 and denoising_steps >= 1
        rgb_norm, input_size = self._preprocess(input_image, processing_res, resample_method)
        bs = batch_size if batch_size > 0 else 1
        preds = []
        for s0 in range(0, ensemble_size, bs):
            e = min(ensemble_size, s0 + bs)
            sn = step_noise[:, s0:e] if step_noise is not None else None
            raw = self.single_infer(rgb_norm.expand(e - s0, -1, -1, -1), denoising_steps, noise[s0:e], sn)
            assert raw.dim() == 4 and raw.shape[1] == 3 * self.n_targets          # :367-370
            preds.append(raw)
        members = torch.cat(preds, dim=0)
        if ensemble_size > 1:
            final, unc = ensemble_iid(members, **(ensemble_kwargs or {}))          # :376-380
        else:
            final, unc = members, None
        if match_input_res:
            final = _resize(final, tuple(input_size[-2:]), resample_method)        # :386-392 (uncertainty is NOT resized)
        return final, unc, members

    def split(self, final, unc=None):
        """fill_outputs :393-411: target i owns channels [3i, 3i + 3)."""
        return {name: (final[:, 3 * i: 3 * i + 3], None if unc is None else unc[:, 3 * i: 3 * i + 3])
                for i, name in enumerate(self.target_names)}
