"""Benchmark of the Marigold denoising hot path (BASELINE.json metric: denoise-steps/sec @768 px).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one denoising iteration (cat -> UNet -> scheduler.step, reference
marigold/marigold_depth_pipeline.py:456-468) of ONE ensemble member at 768x768 (latent 96x96), the
configuration BASELINE.json quotes the metric on (configs[1]: marigold-depth-v1-1, 768x768, E=1, 50 DDIM
steps, bf16 operands, 1 GPU). With N GPUs every rank runs one member (weak scaling, no collective in the
loop) and `value` = N * K / max-over-ranks device time.

Weights are random-init tensors of the SD-2 UNet / SD VAE architecture (no checkpoints offline) and the
image / noise are synthetic: "data": "synthetic". Inputs exceed L2: every step streams the 1.73 GB bf16
UNet weights from HBM (L2 is 126 MB), so no explicit flush is needed between iterations.

One JSON line is printed by rank 0; see DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RES = 768
SCHEDULE_STEPS = 50                     # configs[1]: 50 DDIM steps
F_UNET = {384: 0.418e12, 768: 2.138e12, 1024: 4.658e12}    # algorithmic FLOP per member-step
METRIC = "denoise-steps/sec @768px (UNet forward + scheduler step per ensemble member)"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_sustained": d.get("bf16_tflops_sustained"), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _build_models(kind="full"):
    from tests.helpers import oracle_models

    return oracle_models(kind)


def _ddim_tables(total):
    """Coefficient tables of the 50-step DDIM schedule, cycled to `total` entries."""
    import numpy as np

    from marigold_b200.schedulers import DDIMScheduler

    s = DDIMScheduler()
    s.set_timesteps(SCHEDULE_STEPS)
    kx, kv, kz = s.coefficients()
    idx = np.arange(total) % SCHEDULE_STEPS
    return s.timesteps[idx], kx[idx], kv[idx], kz[idx]


# -------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from marigold_b200 import _lib, parallel
    from marigold_b200.pipeline import MarigoldDepthPipeline
    from marigold_b200.schedulers import DDIMScheduler
    from tests.helpers import engine_from_oracle, synthetic_image

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    from tests.helpers import usable_cores

    torch.set_num_threads(max(1, usable_cores() // max(1, world)))   # N ranks share the box's usable cores
    unet, vae, text = _build_models("full")
    eng = engine_from_oracle(unet, vae, text)
    del vae                                   # weights live on the device now; only rank 0 keeps the fp32 UNet
    if rank != 0 or args.no_cpu_baseline:     # (the checker of the cpu_baseline leg)
        unet = None
    import gc

    gc.collect()
    K, W = args.steps, args.warmup
    ts, kx, kv, kz = _ddim_tables(W + K)
    eng.set_schedule(ts, kx, kv, kz)

    lh = lw = RES // 8
    g = torch.Generator().manual_seed(2024)
    noise_all = torch.randn(max(world, 1), 4, lh, lw, generator=g)           # member k uses row k on any rank
    img = synthetic_image(RES)
    rgb = (img.float() / 255.0 * 2 - 1).to(dev)
    rgb_latent = eng.encode(rgb).contiguous()
    target = noise_all[rank:rank + 1].to(dev).contiguous()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed K steps, inputs resident in HBM -------------------------------------------
    eng.denoise_range_(rgb_latent, target, 0, W)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.mgb_launch_count()
    t_cpu0 = time.perf_counter()
    e0.record()
    eng.denoise_range_(rgb_latent, target, W, K)
    e1.record()
    t_enqueue = time.perf_counter() - t_cpu0
    sync_all()
    launches = int(lib.mgb_launch_count() - l0)
    clocks = sampler.stop()
    ms_local = e0.elapsed_time(e1)
    ms = parallel.barrier_max_ms(ms_local, dev)
    assert torch.isfinite(target).all(), "non-finite latent after the timed region"
    value = world * K / (ms / 1e3)

    # ---- end to end through the public pipeline API: host image in, numpy depth out -------------
    n_e2e = min(K, SCHEDULE_STEPS)
    pipe = MarigoldDepthPipeline(eng, DDIMScheduler(), text, default_denoising_steps=n_e2e,
                                 default_processing_resolution=RES)
    img_pinned = img.pin_memory()
    noise_pinned = noise_all.pin_memory()
    E = world                                                               # 1 member per GPU
    import logging

    logging.disable(logging.WARNING)
    pipe(img_pinned, ensemble_size=E, noise=noise_pinned, color_map=None, show_progress_bar=False)   # warm-up
    sync_all()
    reps = 2
    t0 = time.perf_counter()
    for _ in range(reps):
        out = pipe(img_pinned, ensemble_size=E, noise=noise_pinned, color_map=None, show_progress_bar=False)
    torch.cuda.synchronize()
    t_e2e_local = (time.perf_counter() - t0) / reps
    t_e2e = parallel.barrier_max_ms(t_e2e_local * 1e3, dev) / 1e3
    e2e_value = E * n_e2e / t_e2e
    h2d = img_pinned.numel() * img_pinned.element_size() + (noise_pinned.numel() // max(world, 1)) * 4
    d2h = out.depth_np.size * 4

    # ---- dominant kernel alone (CUDA-graph replay => pure device time) ---------------------------
    kern = None
    if rank == 0 and not args.no_kernel_roofline:
        kern = _dominant_kernel_roofline(torch)

    # ---- CPU baseline (oracle port) on a bounded sample ------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = _cpu_baseline(unet, text, steps=1)
    eng.close()

    if rank == 0:
        pk = _peaks()
        achieved = world * K * F_UNET[RES] / (ms / 1e3) / 1e12
        peak = pk["tflops_sustained"] * world
        line = {
            "metric": METRIC, "value": value, "unit": "denoise-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "marigold-depth-v1-1 architecture (SD-2 UNet 865.9M + SD VAE), 768x768, "
                                   "ensemble_size=1 member per GPU, 50-step DDIM schedule (trailing, zero-SNR, v-pred)",
                       "members_per_gpu": 1, "latent": [lh, lw], "parallelism": f"members-dp{world}",
                       "l2": "inputs > L2: 1.73 GB of bf16 weights stream from HBM every step",
                       "weights": "random init (torch default init, seed 0)"},
            "clocks": clocks,
            "gpu_launches": launches,
            "cpu_enqueue_ms_per_step": t_enqueue * 1e3 / K,
            "e2e": {"value": e2e_value, "unit": "denoise-steps/s", "h2d_bytes_per_step": h2d / n_e2e,
                    "d2h_bytes_per_step": d2h / n_e2e, "seconds_per_image": t_e2e, "steps_per_call": n_e2e,
                    "includes": "H2D image+noise, resize/normalise, VAE encode, denoise loop, VAE decode, "
                                + ("all-gather + ensemble, " if world > 1 else "") + "resize, D2H depth"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None,
                         "what": "whole fused UNet step (all kernels), algorithmic FLOP 2.138e12 per member-step",
                         "peak_source": pk["source"] + ", bf16_tflops_sustained x n_gpus",
                         "dominant_kernel": kern},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _dominant_kernel_roofline(torch):
    """The top-level 3x3 conv (320 -> 320 @ 96x96, 16 per UNet step): algorithmic FLOP / device time,
    timed alone with a CUDA graph of 20 launches (no host gaps), burst peak as denominator."""
    from marigold_b200 import ops

    pk = _peaks()
    NB, H, W_, C = 1, 96, 96, 320
    x = torch.randn(NB, H, W_, C, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv_weight((torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5).to(torch.bfloat16))
    b = torch.randn(C, device="cuda")
    out = torch.empty(NB, H, W_, C, dtype=torch.float32, device="cuda")
    from marigold_b200 import _lib
    from marigold_b200._lib import check, ptr, stream_ptr

    lib = _lib.load()

    def launch():
        check(lib.mgb_op_conv2d(ptr(x), ptr(w), ptr(b), None, ptr(out), None, NB, H, W_, C, C, 0, 0, 0, 0, 0, None,
                                stream_ptr()), "mgb_op_conv2d")

    launch()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        launch()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(20):
                launch()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 100
    flop = 2.0 * NB * H * W_ * C * C * 9
    ach = flop / (us * 1e-6) / 1e12
    return {"kernel": "gemm_tc_kernel (implicit-GEMM conv3x3 320->320 @96x96)", "us_per_launch": us,
            "achieved": ach, "peak": pk["tflops_burst"], "unit": "TFLOP/s", "frac": ach / pk["tflops_burst"],
            "peak_source": pk["source"] + ", bf16_tflops (burst: kernel timed alone)",
            # one `ncu --set full` capture of this launch (profiles/r01f_ncu_full_summary.txt): dram__bytes_read.sum +
            # dram__bytes_write.sum. Algorithmic bytes are 19.5e6 (A 5.9e6 bf16, weights 1.8e6, fp32 output 11.8e6): the
            # operands and the output stay in the 126 MB L2 between kernels, so DRAM sees less than the algorithm moves.
            "traffic": 7791872, "traffic_source": "ncu r01f, dram read+write bytes per launch"}


def _cpu_baseline(unet, text, steps=1, res=RES):
    """Oracle port (fp32 torch on the host cores) on a bounded sample: `steps` UNet+DDIM steps at `res`."""
    import torch

    from oracle.schedulers import DDIMSchedulerOracle

    from tests.helpers import usable_cores

    torch.set_num_threads(usable_cores())
    lh = res // 8
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, lh, lh, generator=g)
    rgb = torch.randn(1, 4, lh, lh, generator=g)
    o = DDIMSchedulerOracle()
    o.set_timesteps(SCHEDULE_STEPS)
    with torch.no_grad():
        t0 = time.perf_counter()
        for i in range(steps):
            t = o.timesteps[i]
            x = o.step(unet(torch.cat([rgb, x], 1), t, text), t, x)
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "denoise-steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} UNet+DDIM step(s), 1 member, {res}x{res}, fp32 torch CPU oracle (oracle/unet.py), "
                      f"{dt:.1f} s"}


# -------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path. The reference pipeline cannot
    be imported offline (diffusers absent), so this is the oracle PORT (kind="port") on all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch

    K, W = args.steps, args.warmup
    from tests.helpers import usable_cores

    unet, vae, text = _build_models("full")
    torch.set_num_threads(usable_cores())
    # bounded sample: pick the resolution so that W + K steps fit in ~4 minutes on this host
    probe = _cpu_baseline(unet, text, steps=1, res=384)
    t384 = 1.0 / probe["value"]
    budget = 240.0
    res = 768 if (W + K) * t384 * (F_UNET[768] / F_UNET[384]) < budget else 384
    lh = res // 8
    from oracle.schedulers import DDIMSchedulerOracle

    o = DDIMSchedulerOracle()
    o.set_timesteps(SCHEDULE_STEPS)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, lh, lh, generator=g)
    rgb = torch.randn(1, 4, lh, lh, generator=g)
    n_sched = SCHEDULE_STEPS
    with torch.no_grad():
        for i in range(W):
            t = o.timesteps[i % n_sched]
            x = o.step(unet(torch.cat([rgb, x], 1), t, text), t, x)
        t0 = time.perf_counter()
        for i in range(W, W + K):
            t = o.timesteps[i % n_sched]
            x = o.step(unet(torch.cat([rgb, x], 1), t, text), t, x)
        dt = time.perf_counter() - t0
    scale = F_UNET[res] / F_UNET[768]          # FLOP-equivalent 768-px steps
    value = K * scale / dt
    sample = (f"{K} UNet+DDIM steps at {res}x{res} (1 member), fp32 torch CPU oracle port; "
              + ("measured at the metric's resolution" if res == 768 else
                 f"bounded sample: value scaled by F_unet({res})/F_unet(768) = {scale:.4f} to 768-px-equivalent steps"))
    cpu = {"value": value, "unit": "denoise-steps/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "denoise-steps/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": K, "warmup": W,
            "ms_per_step": dt * 1e3 / K / scale, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "marigold-depth-v1-1 architecture, 768x768, ensemble_size=1, 50-step DDIM schedule",
                       "note": "reference pipeline needs diffusers (absent offline): oracle port on host cores"},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "denoise-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-roofline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
