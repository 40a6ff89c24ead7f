"""Data-parallel sweep over independent images: the B200 analogue of the reference's sequential loop
(``run_batch.py:176-261``).  One process per GPU (torchrun); rank r edits ``entries[r::world]`` with replicated
weights; there is NO collective on the hot path.  After the sweep, per-image results (uint8 outputs and timings)
are gathered with ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def dist_env() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard(entries: Sequence[Any], rank: int, world: int) -> List[Any]:
    """Round-robin shard: rank r owns entries r, r+world, ... (balanced to within one image)."""
    return list(entries[rank::world])


def owner_of(index: int, world: int) -> int:
    return index % world


def gather_outputs(local: torch.Tensor, counts: Optional[List[int]] = None) -> Optional[torch.Tensor]:
    """Gather per-rank uint8 outputs [n_r, ...] to every rank and restore the original interleaved order.
    Ranks may own different counts (tail imbalance): shorter shards are padded to the maximum."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    ns = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(ns, n_local)
    ns = [int(n.item()) for n in ns]
    nmax = max(ns)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    total = sum(ns)
    out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        if ns[r]:
            out[r::world][: ns[r]] = bufs[r][: ns[r]]
    return out


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
