"""Intrinsic-image-decomposition output containers (host side).

Mirrors `IIDEntry` / `MarigoldIIDOutput` of the reference (marigold/marigold_iid_pipeline.py:59-160): per target a
[3,H,W] array in [0,1], an 8-bit visualisation that depends on the target's `prediction_space` ("srgb" and "stack" as
is; "linear" gamma-encoded with 1/2.2 after an optional rescale to the maximum), and the ensembling uncertainty.
Filled by `marigold_b200.pipeline.MarigoldIIDPipeline` (engine with 4 (n + 1) / 4 n latent channels, per-target
decode, `ensemble_iid`); checker side: oracle.pipeline.OracleIIDPipeline, oracle.ensemble.ensemble_iid.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import numpy as np

try:  # PIL is optional at import time (same policy as pipeline.py)
    from PIL import Image
except Exception:  # noqa: BLE001
    Image = None


@dataclass
class IIDEntry:
    name: str
    array: Optional[np.ndarray] = None          # [3, H, W] in [0, 1]
    image: Optional["Image.Image"] = None       # [H, W, 3] uint8
    uncertainty: Optional[np.ndarray] = None


def visualize_target(array: np.ndarray, properties: Optional[Dict[str, Any]]) -> np.ndarray:
    """[3,H,W] float in [0,1] -> [H,W,3] uint8 following marigold_iid_pipeline.py:121-137."""
    props = properties or {}
    space = props.get("prediction_space", "srgb")
    img = np.asarray(array)
    if space == "linear":
        if props.get("up_to_scale", False):
            img = img / max(img.max(), 1e-6)
        img = img ** (1 / 2.2)
    elif space not in ("srgb", "stack"):
        pass                                     # the reference leaves unknown spaces untouched as well
    img = (img * 255).astype(np.uint8)
    return np.moveaxis(img, 0, -1)


class MarigoldIIDOutput:
    def __init__(self, target_names: List[str]):
        self.n_targets = len(target_names)
        self.target_names = target_names
        self.entries: List[IIDEntry] = [IIDEntry(name=name) for name in target_names]
        self._entry_map = {entry.name: entry for entry in self.entries}
        self._filled_entries = set()

    def fill_entry(self, name: str, prediction, uncertainty=None,
                   target_properties: Optional[Dict[str, Any]] = None) -> None:
        if name not in self._entry_map:
            raise KeyError(f"Unknown entry name: {name}")
        if name in self._filled_entries:
            raise RuntimeError(f"Entry {name} already filled")
        entry = self._entry_map[name]
        to_np = lambda t: t.squeeze().cpu().numpy() if hasattr(t, "cpu") else np.asarray(t).squeeze()  # noqa: E731
        array = to_np(prediction)
        vis = visualize_target(array, (target_properties or {}).get(name))
        entry.array = array
        entry.image = Image.fromarray(vis) if Image is not None else vis
        entry.uncertainty = to_np(uncertainty) if uncertainty is not None else None
        self._filled_entries.add(name)

    @property
    def is_complete(self) -> bool:
        return len(self._filled_entries) == self.n_targets

    def __getitem__(self, key: str) -> IIDEntry:
        return self._entry_map[key]

    def __iter__(self):
        return iter(self.entries)


def fill_outputs(output: MarigoldIIDOutput, final_pred, pred_uncert, target_names: List[str],
                 target_properties: Optional[Dict[str, Any]]) -> None:
    """marigold_iid_pipeline.py:393-411: target i owns channels [3i, 3i+3) of the [1, 3n, H, W] prediction."""
    for i, name in enumerate(target_names):
        a, b = 3 * i, 3 * i + 3
        output.fill_entry(name, final_pred[:, a:b], pred_uncert[:, a:b] if pred_uncert is not None else None,
                          target_properties)
