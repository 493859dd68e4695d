"""Where one pipeline call spends its time at the quoted size: python tools/e2e_phases.py [res] [steps]

Device time of VAE encode, the denoising loop and VAE decode (CUDA events, warm), and the wall time of the whole
MarigoldDepthPipeline call around them (host image in, numpy map out)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from marigold_b200.pipeline import MarigoldDepthPipeline  # noqa: E402
from marigold_b200.schedulers import DDIMScheduler  # noqa: E402
from tests.helpers import engine_from_oracle, oracle_models  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
unet, vae, text = oracle_models("full")
eng = engine_from_oracle(unet, vae, text)
sched = DDIMScheduler()
pipe = MarigoldDepthPipeline(eng, sched, text, default_denoising_steps=steps, default_processing_resolution=res)
g = torch.Generator().manual_seed(5)
img = torch.randint(0, 256, (1, 3, res, res), generator=g, dtype=torch.uint8).pin_memory()
noise = torch.randn(1, 4, res // 8, res // 8, generator=g).pin_memory()
kw = dict(noise=noise, show_progress_bar=False, color_map=None)
pipe(img, **kw)
torch.cuda.synchronize()


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


rgb = (img.float() / 255 * 2 - 1).cuda()
t_enc, lat = timed(lambda: eng.encode(rgb))
z = noise.cuda()
t_den, tgt = timed(lambda: eng.denoise(lat, z.clone()), 2)
t_dec, _ = timed(lambda: eng.decode(tgt, 0))
t0 = time.perf_counter()
for _ in range(3):
    pipe(img, **kw)
torch.cuda.synchronize()
t_call = (time.perf_counter() - t0) / 3 * 1e3
print(f"res {res}, {steps} steps: encode {t_enc:.2f} ms, denoise {t_den:.2f} ms ({t_den / steps:.3f}/step), decode {t_dec:.2f} ms, "
      f"whole call {t_call:.2f} ms, host+copies+resize = {t_call - t_enc - t_den - t_dec:.2f} ms")
eng.close()
