"""Device-side evaluation step that follows the hot path in dataset evaluation (csrc/eval.cu): least-squares scale / shift
alignment (reference src/util/alignment.py:35-82), the clips of script/depth/eval.py:201-207 and the masked depth metrics
of src/util/metric.py:64-191 — two streaming passes and one host synchronisation per sample."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

METRIC_NAMES = ("abs_relative_difference", "squared_relative_difference", "rmse_linear", "rmse_log", "log10", "delta1_acc",
                "delta2_acc", "delta3_acc", "i_rmse", "silog_rmse")
_ws = {}


def _run(pred, gt, mask, least_squares: bool, dmin: float, dmax: float, want_aligned: bool):
    if not (pred.is_cuda and gt.is_cuda):
        raise _lib.MgbError("marigold_b200.evaluation needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    p = pred.to(torch.float32).contiguous().reshape(-1)
    g = gt.to(torch.float32).contiguous().reshape(-1)
    assert p.numel() == g.numel(), f"{tuple(pred.shape)} vs {tuple(gt.shape)}"
    m = None
    if mask is not None:
        m = mask.to(device=p.device).to(torch.uint8).contiguous().reshape(-1)
        assert m.numel() == p.numel()
    with torch.cuda.device(p.device):
        ws = _ws.get(p.device)
        if ws is None:
            ws = torch.empty(int(lib.mgb_eval_ws_bytes()), dtype=torch.uint8, device=p.device)
            _ws[p.device] = ws
        aligned = torch.empty_like(p) if want_aligned else None
        out = np.zeros(13, dtype=np.float64)
        check(lib.mgb_eval_depth(ptr(p), ptr(g), ptr(m), p.numel(), int(least_squares), float(dmin), float(dmax), ptr(aligned),
                                 ptr(ws), out.ctypes.data_as(C.c_void_p), stream_ptr()), "mgb_eval_depth")
    return out, (aligned.reshape(pred.shape) if aligned is not None else None)


def align_depth_least_square(gt: torch.Tensor, pred: torch.Tensor, valid_mask: Optional[torch.Tensor],
                             return_scale_shift: bool = True, max_resolution: Optional[int] = None):
    """src/util/alignment.py:35-82 on the device: (aligned_pred, scale, shift). `max_resolution` downsamples the three
    maps with the reference's nearest Upsample before the fit; the fitted map is always the full-resolution prediction."""
    fit_p, fit_g, fit_m = pred, gt, valid_mask
    if max_resolution is not None:
        sf = float(np.min(max_resolution / np.array(pred.shape[-2:])))
        if sf < 1:
            down = torch.nn.Upsample(scale_factor=sf, mode="nearest")
            fit_g = down(gt.reshape(1, 1, *gt.shape[-2:]).float())
            fit_p = down(pred.reshape(1, 1, *pred.shape[-2:]).float())
            fit_m = down(valid_mask.reshape(1, 1, *valid_mask.shape[-2:]).float()).bool() if valid_mask is not None else None
    out, _ = _run(fit_p, fit_g, fit_m, True, -3.0e38, 3.0e38, False)      # only the fit is used from this call
    scale, shift = out[0], out[1]
    aligned = pred.to(torch.float64) * scale + shift
    return (aligned, scale, shift) if return_scale_shift else aligned


def evaluate_depth(pred: torch.Tensor, gt: torch.Tensor, valid_mask: Optional[torch.Tensor] = None,
                   alignment: Optional[str] = "least_square", min_depth: float = 1e-6, max_depth: float = 3.0e38
                   ) -> Tuple[Dict[str, float], Dict[str, float]]:
    """One sample of script/depth/eval.py:171-217: align (or not), clip to the dataset range and to d > 1e-6, all metrics.
    Returns (metrics by the reference's function names, {"scale", "shift", "n_valid"})."""
    if alignment not in (None, "least_square"):
        raise ValueError(f"unsupported alignment {alignment!r} (least_square_disparity is not implemented on the device)")
    out, _ = _run(pred, gt, valid_mask, alignment == "least_square", min_depth, max_depth, False)
    metrics = dict(zip(METRIC_NAMES, (float(v) for v in out[3:13])))
    return metrics, {"scale": float(out[0]), "shift": float(out[1]), "n_valid": int(out[2])}
