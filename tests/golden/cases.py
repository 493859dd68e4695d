"""Seeded inputs shared by make_golden.py (reference run, build container) and the tests."""
import numpy as np
import torch

DEPTH_CASES = {
    "e2_default": dict(E=2, H=40, W=56, seed=1),
    "e3_default": dict(E=3, H=40, W=56, seed=2),
    "e8_default": dict(E=8, H=48, W=64, seed=3),
    "e10_default": dict(E=10, H=48, W=64, seed=4),
    "e10_unc": dict(E=10, H=32, W=48, seed=5, kwargs=dict(output_uncertainty=True)),
    "e8_mean_unc": dict(E=8, H=32, W=48, seed=6, kwargs=dict(reduction="mean", output_uncertainty=True)),
    "e4_scale_only": dict(E=4, H=32, W=48, seed=7, kwargs=dict(shift_invariant=False)),
    "e5_maxres": dict(E=5, H=72, W=96, seed=8, kwargs=dict(max_res=48)),
    "e6_smallmin": dict(E=6, H=32, W=48, seed=9, small_min=True),
    "e10_noreg": dict(E=10, H=32, W=32, seed=10, kwargs=dict(regularizer_strength=0.0)),
}
NORMALS_CASES = {
    "e2": dict(E=2, H=32, W=48, seed=21),
    "e3_unc": dict(E=3, H=32, W=48, seed=22, kwargs=dict(output_uncertainty=True)),
    "e8": dict(E=8, H=40, W=56, seed=23),
    "e10_unc": dict(E=10, H=40, W=56, seed=24, kwargs=dict(output_uncertainty=True)),
    "e10_mean": dict(E=10, H=24, W=32, seed=25, kwargs=dict(reduction="mean", output_uncertainty=True)),
    "e4_ties": dict(E=4, H=24, W=32, seed=26, ties=True),
}
IID_CASES = {
    "e2_median": dict(E=2, C=6, H=24, W=32, seed=41),
    "e5_median_unc": dict(E=5, C=6, H=24, W=32, seed=42, kwargs=dict(output_uncertainty=True)),
    "e4_mean_unc": dict(E=4, C=9, H=16, W=24, seed=43, kwargs=dict(reduction="mean", output_uncertainty=True)),
    "e1": dict(E=1, C=6, H=8, W=8, seed=44, kwargs=dict(output_uncertainty=True)),
}
RESIZE_CASES = {
    "bilinear_down": dict(H=90, W=130, max_edge=64, method="bilinear", seed=31),
    "bilinear_up": dict(H=30, W=20, max_edge=48, method="bilinear", seed=32),
    "nearest": dict(H=50, W=70, max_edge=32, method="nearest", seed=33),
}


def depth_input(cfg) -> torch.Tensor:
    """[E,1,H,W] in [0,1]: a smooth scene seen through per-member affine maps plus noise — the
    structure ensemble_depth is designed for (ensemble.py:51-57)."""
    rng = np.random.default_rng(cfg["seed"])
    E, H, W = cfg["E"], cfg["H"], cfg["W"]
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    base = 0.5 + 0.25 * np.sin(3 * xx + 2 * yy) + 0.2 * xx * yy
    base = (base - base.min()) / (base.max() - base.min())
    s = rng.uniform(0.4, 0.9, size=(E, 1, 1))
    t = rng.uniform(0.0, 0.1, size=(E, 1, 1)) if not cfg.get("small_min") else rng.uniform(0.0, 0.004, size=(E, 1, 1))
    d = base[None] * s + t + rng.normal(0, 0.01, size=(E, H, W))
    d = np.clip(d, 0.0, 1.0).astype(np.float32)
    return torch.from_numpy(d)[:, None]


def normals_input(cfg) -> torch.Tensor:
    rng = np.random.default_rng(cfg["seed"])
    E, H, W = cfg["E"], cfg["H"], cfg["W"]
    base = rng.normal(size=(1, 3, H, W))
    base[:, 2] += 1.5
    n = base + rng.normal(0, 0.3, size=(E, 3, H, W))
    n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-6)
    n = n.astype(np.float32)
    if cfg.get("ties"):
        n[1] = n[0]          # exact duplicates: argmax must pick the lowest index
        n[3] = n[2]
    return torch.from_numpy(n)


def resize_input(cfg) -> torch.Tensor:
    rng = np.random.default_rng(cfg["seed"])
    return torch.from_numpy(rng.integers(0, 256, size=(1, 3, cfg["H"], cfg["W"]), dtype=np.uint8))


def iid_input(cfg) -> torch.Tensor:
    """[E, C, H, W] in [0,1]: per-member noisy copies of a smooth multi-channel target (albedo / material maps)."""
    rng = np.random.default_rng(cfg["seed"])
    E, C, H, W = cfg["E"], cfg["C"], cfg["H"], cfg["W"]
    base = rng.uniform(0, 1, size=(1, C, 1, 1)) * np.ones((1, C, H, W)) + 0.1 * np.sin(np.linspace(0, 6, W))[None, None, None, :]
    x = np.clip(base + rng.normal(0, 0.05, size=(E, C, H, W)), 0, 1).astype(np.float32)
    if E >= 4:
        x[1] = x[0]            # exact ties
    return torch.from_numpy(x)
