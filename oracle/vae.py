"""ORACLE (test infrastructure only — never imported by the product path).

Plain-torch fp32 restatement of the sub-modules of diffusers' `AutoencoderKL` that Marigold calls
(reference marigold/marigold_depth_pipeline.py:491-492 `vae.encoder`, `vae.quant_conv`;
:512-513 `vae.post_quant_conv`, `vae.decoder`). SD VAE config, SURVEY.md App. A.2.

PARITY UNPINNED: diffusers is absent (see oracle/unet.py). Cross-check: parameter counts 34.2 M
(encoder) / 49.5 M (decoder) for the SD config (tests/test_oracle.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import ResnetBlock2D


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: List[int] = field(default_factory=lambda: [128, 256, 512, 512])
    layers_per_block: int = 2
    latent_channels: int = 4
    norm_num_groups: int = 32
    norm_eps: float = 1e-6

    @staticmethod
    def tiny():
        return VAEConfig(block_out_channels=[64, 64, 128, 128])


class VAEAttention(nn.Module):
    """Single-head attention over h*w tokens, dim = channels; q/k/v/out with bias."""

    def __init__(self, c, groups, eps):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).reshape(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]  # scale 1/sqrt(C)
        o = self.to_out[0](o)
        return x + o.transpose(1, 2).reshape(B, C, H, W)


class VAEMidBlock(nn.Module):
    def __init__(self, c, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, 0, groups, eps), ResnetBlock2D(c, c, 0, groups, eps)])
        self.attentions = nn.ModuleList([VAEAttention(c, groups, eps)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class VAEDownsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))  # asymmetric: right/bottom only


class VAEUpsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, eps, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, 0, groups, eps) for i in range(n)])
        self.downsamplers = nn.ModuleList([VAEDownsample(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, eps, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, 0, groups, eps) for i in range(n)])
        self.upsamplers = nn.ModuleList([VAEUpsample(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch, g, e = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        prev = ch[0]
        for i, c in enumerate(ch):
            self.down_blocks.append(_EncBlock(prev, c, cfg.layers_per_block, g, e, down=i < len(ch) - 1))
            prev = c
        self.mid_block = VAEMidBlock(ch[-1], g, e)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=e)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for b in self.down_blocks:
            h = b(h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch, g, e = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = VAEMidBlock(rev[0], g, e)
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, c in enumerate(rev):
            self.up_blocks.append(_DecBlock(prev, c, cfg.layers_per_block + 1, g, e, up=i < len(ch) - 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, rev[-1], eps=e)
        self.conv_out = nn.Conv2d(rev[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid_block(h)
        for b in self.up_blocks:
            h = b(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLOracle(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
