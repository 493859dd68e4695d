"""Freeze the CPU oracle's network outputs on the seeded tiny configuration ("its own golden
vectors"). This does NOT pin the oracle to the reference (diffusers is absent: the network arithmetic stays "parity
unpinned"); it pins the oracle to ITSELF, so that an accidental edit of oracle/*.py or of the seeded
builders in tests/helpers.py shows up as a CPU test failure instead of silently moving the GPU parity target.

    python tests/golden/make_network_golden.py        # writes tests/golden/network_golden.npz (fp32, ~60 KB)
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle.pipeline import OracleDepthPipeline, OracleNormalsPipeline  # noqa: E402
from oracle.schedulers import DDIMSchedulerOracle, LCMSchedulerOracle  # noqa: E402
from tests.helpers import oracle_models, synthetic_image  # noqa: E402


def compute():
    unet, vae, text = oracle_models("tiny")
    g = torch.Generator().manual_seed(11)
    rgb = torch.randn(2, 4, 8, 24, generator=g)
    x = torch.randn(2, 4, 8, 24, generator=g)
    out = {}
    with torch.no_grad():
        for t in (999, 499, 1):
            out[f"unet_t{t}"] = unet(torch.cat([rgb, x], 1), t, text.repeat(2, 1, 1)).numpy()
        img = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(12)) * 2 - 1
        out["vae_encode_mean_scaled"] = (vae.quant_conv(vae.encoder(img))[:, :4] * 0.18215).numpy()
        lat = torch.randn(1, 4, 8, 16, generator=torch.Generator().manual_seed(13))
        out["vae_decode_raw"] = vae.decoder(vae.post_quant_conv(lat / 0.18215)).numpy()
    image = synthetic_image(64)
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(2024))
    zs = torch.randn(3, 2, 4, 8, 8, generator=torch.Generator().manual_seed(2025))
    d, _, _ = OracleDepthPipeline(unet, vae, DDIMSchedulerOracle(), text, 4, 64)(image, ensemble_size=1, noise=z[:1])
    out["depth_ddim4"] = np.asarray(d, dtype=np.float32)
    d, _, _ = OracleDepthPipeline(unet, vae, LCMSchedulerOracle(), text, 4, 64)(image, ensemble_size=1, noise=z[:1],
                                                                                  step_noise=zs[:, :1])
    out["depth_lcm4"] = np.asarray(d, dtype=np.float32)
    n = OracleNormalsPipeline(unet, vae, DDIMSchedulerOracle(), text, 2, 64)(image, ensemble_size=1, noise=z[:1])
    out["normals_ddim2"] = np.asarray(n[0], dtype=np.float32)
    return out


if __name__ == "__main__":
    torch.set_num_threads(4)
    o = compute()
    p = Path(__file__).resolve().parent / "network_golden.npz"
    np.savez_compressed(p, **o)
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in o.items()}, p.stat().st_size)
