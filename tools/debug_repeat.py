"""Bisect helper: run the same 2-step, E=3 (batches 2+1) inference repeatedly and report NaNs per call."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from marigold_b200.schedulers import DDIMScheduler  # noqa: E402
from tests.helpers import engine_from_oracle, oracle_models  # noqa: E402

unet, vae, text = oracle_models("tiny")
eng = engine_from_oracle(unet, vae, text)
s = DDIMScheduler()
s.set_timesteps(2)
eng.set_schedule(s.timesteps, *s.coefficients())
g = torch.Generator().manual_seed(0)
img = torch.rand(1, 3, 64, 128, generator=g) * 2 - 1
z = torch.randn(3, 4, 8, 16, generator=g)
for call in range(4):
    lat = eng.encode(img.cuda())
    outs = []
    for ids in ([0, 1], [2]):
        t = eng.denoise(lat.expand(len(ids), -1, -1, -1).contiguous(), z[ids].cuda())
        d = eng.decode(t, 0)
        outs.append((torch.isnan(t).sum().item(), d.abs().max().item()))
    if call == 1:  # exercise the ensemble kernels in between, like the pipeline does
        from marigold_b200.ensemble import ensemble_depth
        ensemble_depth(torch.rand(3, 1, 64, 128, device="cuda"), engine=eng)
    print("call", call, "nan_in_latent / max_depth per batch:", outs, flush=True)
eng.close()
