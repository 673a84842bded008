"""Multi-GPU plumbing of the scene-sharded path (one process per GPU, torch.distributed).

The reference shards inference over scenes with `DistributedSampler(shuffle=False)`
(inference_combined.py:247-256), one process per GPU (`mp.spawn`, :343-350), and gathers per-scene results
with `dist.gather_object` (pixie/metrics.py:150). Nothing on the compute path is collective; the same holds
here — NCCL (or gloo in the CPU tests) only carries the barrier, the max-over-ranks of timings and the final
gather of small per-scene records.
"""
from __future__ import annotations

import os
from typing import Any, List, Optional

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: Optional[str] = None, device: Optional[torch.device] = None) -> None:
    """init_process_group from the torchrun environment (MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE)."""
    rank, world, _ = env_rank_world()
    if world <= 1 or dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)


def shard_scenes(n_scenes: int, rank: int, world: int) -> List[int]:
    """Indices of the scenes rank `rank` processes — identical to
    `list(DistributedSampler(range(n_scenes), num_replicas=world, rank=rank, shuffle=False))`:
    the index list is padded by wrapping around to a multiple of `world`, then strided."""
    if n_scenes <= 0:
        return []
    total = -(-n_scenes // world) * world
    idx = list(range(n_scenes))
    while len(idx) < total:
        idx += idx[: total - len(idx)]
    return idx[rank:total:world]


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Timings are reported as the max over ranks (device-side for NCCL)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_records(local: List[Any], dst: int = 0) -> Optional[List[Any]]:
    """Per-scene records of all ranks, flattened in rank order on `dst` (None elsewhere);
    mirrors InferenceMetrics.gather_all_metrics (pixie/metrics.py:132-153)."""
    if not dist.is_initialized():
        return list(local)
    rank = dist.get_rank()
    out = [None] * dist.get_world_size() if rank == dst else None
    dist.gather_object(local, out, dst=dst)
    if rank != dst:
        return None
    return [r for part in out for r in part]
