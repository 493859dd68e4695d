"""Run only UNet denoising steps of the full-size model (for ncu / timing): python tools/step_only.py [steps] [res]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from marigold_b200.schedulers import DDIMScheduler  # noqa: E402
from tests.helpers import engine_from_oracle, oracle_models  # noqa: E402

_pos = [a for a in sys.argv[1:] if not a.startswith('--')]
steps = int(_pos[0]) if len(_pos) > 0 else 3
res = int(_pos[1]) if len(_pos) > 1 else 768
unet, vae, text = oracle_models("full")
eng = engine_from_oracle(unet, vae, text)
s = DDIMScheduler()
s.set_timesteps(max(steps, 1))
eng.set_schedule(s.timesteps, *s.coefficients())
lh = res // 8
g = torch.Generator().manual_seed(11)
rgb = torch.randn(1, 4, lh, lh, generator=g).cuda()
x = torch.randn(1, 4, lh, lh, generator=g).cuda()
torch.cuda.synchronize()
warm = not any(a == "--cold" for a in sys.argv)
if warm:        # first call: eager step + graph capture + one-time allocations / module loads (not timed)
    eng.denoise(rgb, x)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = eng.denoise(rgb, x)
e1.record()
torch.cuda.synchronize()
print(f"{steps} steps ({'graph replay' if warm else 'cold'}): {e0.elapsed_time(e1) / steps:.3f} ms/step, finite={bool(torch.isfinite(out).all())} "
      f"checksum mean_abs={out.abs().mean().item():.6f} sum={out.double().sum().item():.4f} x[0,0,5,7]={out[0,0,5,7].item():.6f}")
eng.close()
