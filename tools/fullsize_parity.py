"""Full-size (768 x 768, real SD-2 shapes, seeded random weights) parity of every stage against the fp32 CPU oracle:
VAE encode, one UNet + DDIM step, VAE decode (depth head). The unit tests do this on the tiny configuration only;
full-size output is otherwise only checked for finiteness (bench / net_check). Takes a few minutes of CPU time on
the GPU box (the oracle UNet step is ~3 s, the VAE decoder ~1 min at 768 x 768).

    python tools/fullsize_parity.py [res]        # writes gpurun_out/fullsize_parity.json

Round-2 to-do item: also the place to confirm the PlainLaunchScope fix of the VAE attention GEMMs.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from marigold_b200.schedulers import DDIMScheduler  # noqa: E402
from oracle.schedulers import DDIMSchedulerOracle  # noqa: E402
from tests.helpers import engine_from_oracle, oracle_models, rel_err, synthetic_image, usable_cores  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 768
torch.set_num_threads(usable_cores())
unet, vae, text = oracle_models("full")
eng = engine_from_oracle(unet, vae, text)
s = DDIMScheduler()
s.set_timesteps(2)
eng.set_schedule(s.timesteps, *s.coefficients())
img = (synthetic_image(res).float() / 255.0 * 2.0 - 1.0)
g = torch.Generator().manual_seed(2024)
x0 = torch.randn(1, 4, res // 8, res // 8, generator=g)
out = {"res": res}

# --- product (two different images back to back: a stale-operand race in the VAE attention would show up here)
other = (synthetic_image(res, seed=99).float() / 255.0 * 2.0 - 1.0)
eng.encode(other.cuda())
lat = eng.encode(img.cuda())
x = x0.cuda().clone()
mo = eng.unet_step(lat, x, 0, want_model_out=True)
dep = eng.decode(x, 0)
torch.cuda.synchronize()

# --- checker
t0 = time.time()
with torch.no_grad():
    rl = vae.quant_conv(vae.encoder(img))[:, :4] * 0.18215
    out["encode_rel"] = rel_err(lat, rl)
    o = DDIMSchedulerOracle()
    o.set_timesteps(2)
    t = o.timesteps[0]
    v_prod_in = unet(torch.cat([lat.cpu(), x0], 1), t, text)          # same input as the product saw
    out["unet_out_rel"] = rel_err(mo, v_prod_in)
    out["sched_rel"] = rel_err(x, o.step(mo.cpu(), t, x0))             # fused epilogue vs oracle step on the product's v
    ref_dep = (vae.decoder(vae.post_quant_conv(x.cpu() / 0.18215)).mean(1, keepdim=True).clip(-1, 1) + 1) / 2
    out["decode_rel"] = rel_err(dep, ref_dep)
    out["decode_mean_abs"] = float((dep.cpu() - ref_dep).abs().mean())
out["oracle_seconds"] = round(time.time() - t0, 1)
out["finite"] = bool(torch.isfinite(dep).all() and torch.isfinite(mo).all())
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "fullsize_parity.json").write_text(json.dumps(out, indent=1))
eng.close()
