"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch).

The path shards by WINDOW (window path) or by VIDEO (frame path) with no data-path collective
(SURVEY.md section 8e); the only exchange step is ONE sum all-reduce of the flat gradient buffer per
step, followed by the fused Adam kernel that applies the 1/world_size scale.  BatchNorm statistics stay
per rank (DDP default), so results parity is asserted at world_size 1 and per rank above it.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the process group when
    WORLD_SIZE > 1 (nccl on CUDA, gloo otherwise)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_batch(idx: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Contiguous share of a global batch for one rank.  Even split (sizes differ by at most one), so that a rank's share is
    empty only if the batch has fewer samples than there are ranks -- the loaders drop such a batch on EVERY rank
    (every step ends in a gradient all-reduce: a rank that skipped it would leave the others waiting)."""
    if world_size <= 1:
        return idx
    n = idx.numel()
    return idx[rank * n // world_size:(rank + 1) * n // world_size]


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def max_over_ranks(value: float, device=None) -> float:
    """Step time of the job = slowest rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
