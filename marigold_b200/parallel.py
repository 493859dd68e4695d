"""Ensemble-member sharding across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-process; the only place members meet is the ensemble
(marigold/marigold_depth_pipeline.py:281-300). So: the job list is (image, member); rank r takes members
r, r+G, r+2G, ...; every rank encodes its image locally (cheaper than a broadcast) and reads ITS rows
of the pre-drawn noise tensor; one all-gather of the decoded per-member maps (NCCL over NVLink on
GPUs, gloo on CPU in the tests) precedes the ensemble, which then runs replicated and deterministic.
There is no collective inside the denoising loop.
"""
from __future__ import annotations

from typing import List, Optional

import torch

try:
    import torch.distributed as dist
except Exception:  # noqa: BLE001
    dist = None


def world() -> tuple:
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def member_indices(ensemble_size: int, rank: int, world_size: int) -> List[int]:
    """Round-robin ownership: member m belongs to rank m % G."""
    return list(range(rank, ensemble_size, world_size))


def slots_per_rank(ensemble_size: int, world_size: int) -> int:
    return (ensemble_size + world_size - 1) // world_size


def gather_members(local: torch.Tensor, ensemble_size: int, group=None) -> torch.Tensor:
    """local: [n_local, C, H, W] predictions of this rank's members (in member_indices order).
    Returns [ensemble_size, C, H, W] in global member order on every rank: ONE all_gather of
    ceil(E/G) padded slots per rank."""
    rank, G = world()
    if G == 1:
        assert local.shape[0] == ensemble_size
        return local
    S = slots_per_rank(ensemble_size, G)
    pad = torch.zeros((S,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((G * S,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    # slot (r, s) holds member r + s*G
    out = out.reshape(G, S, *local.shape[1:])
    members = [out[m % G, m // G] for m in range(ensemble_size)]
    return torch.stack(members, dim=0)


def barrier_max_ms(ms: float, device: Optional[torch.device] = None) -> float:
    """Max over ranks of a locally measured duration (multi-GPU numbers are max-over-ranks)."""
    rank, G = world()
    if G == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
