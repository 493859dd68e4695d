"""Round-2 first action: validate MGB_GN_DETERMINISTIC=1 (norm_fx.cu) on the GPU.
Runs the tiny seeded pipeline (4-step DDIM, E=1) twice per setting in separate processes and reports
 (a) max |default - deterministic| (must be at run-to-run noise level, ~1e-2 worst pixel),
 (b) whether two deterministic runs are bit-identical (the point of the exercise),
 (c) full-size step time of both settings (tools/step_only.py).
    python tools/check_deterministic.py
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from marigold_b200.pipeline import MarigoldDepthPipeline
from marigold_b200.schedulers import DDIMScheduler
from tests.helpers import engine_from_oracle, oracle_models, synthetic_image
unet, vae, text = oracle_models("tiny")
pipe = MarigoldDepthPipeline(engine_from_oracle(unet, vae, text), DDIMScheduler(), text, default_denoising_steps=4,
                             default_processing_resolution=128)
z = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(2024))
out = pipe(synthetic_image(128), ensemble_size=1, noise=z, show_progress_bar=False)
np.save(sys.argv[1], out.depth_np)
''' % str(ROOT)


def run(tag, det):
    env = dict(os.environ)
    env["MGB_GN_DETERMINISTIC"] = "1" if det else "0"
    p = ROOT / "gpurun_out" / f"det_{tag}.npy"
    p.parent.mkdir(exist_ok=True)
    subprocess.run([sys.executable, "-c", CHILD, str(p)], env=env, check=True)
    import numpy as np

    return np.load(p)


if __name__ == "__main__":
    import numpy as np

    a0, a1 = run("default_a", False), run("default_b", False)
    d0, d1 = run("det_a", True), run("det_b", True)
    print("default vs default : max abs diff", float(np.abs(a0 - a1).max()), "bit-identical:", bool(np.array_equal(a0, a1)))
    print("det     vs det     : max abs diff", float(np.abs(d0 - d1).max()), "bit-identical:", bool(np.array_equal(d0, d1)))
    print("default vs det     : max abs diff", float(np.abs(a0 - d0).max()), "mean", float(np.abs(a0 - d0).mean()))
    for det in ("0", "1"):
        env = dict(os.environ, MGB_GN_DETERMINISTIC=det)
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "step_only.py"), "12"], env=env, capture_output=True, text=True)
        print(f"MGB_GN_DETERMINISTIC={det}:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
