"""Data-parallel plumbing: one process per GPU, torch.distributed over NCCL (gloo on CPU for tests).

The path shards by the batch dimension only; its single exchange step is the gradient all-reduce (SUM / world) that
``torch.nn.parallel.DistributedDataParallel`` overlaps with the backward kernels (reference: main_train.py:90-96).
BatchNorm statistics stay per-GPU like the reference recipe (``batch_norm``, not ``sync_batch_norm``), so there is no
collective inside the forward.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend: str = "nccl", device=None) -> Tuple[int, int, int]:
    rank, world, local_rank = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def shard_seed(base: int, rank: int) -> int:
    """Independent synthetic shard per rank (the reference's BatchSamplerDDP gives every rank a disjoint slice)."""
    return base + 1000003 * rank


def max_over_ranks(value: float, device="cpu") -> float:
    """Multi-GPU timings are the MAX over ranks of the device-side time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def weak_scaling_throughput(per_rank_units: int, world: int, max_ms_per_step: float) -> float:
    """Whole-job units/s: every rank processes ``per_rank_units`` per step; the step takes the slowest rank's time."""
    return world * per_rank_units / (max_ms_per_step * 1e-3)


def allreduce_mean_(tensors, world: int) -> None:
    """Flat bucketed gradient all-reduce (SUM / world) -- what DDP does per bucket; used by tests and by callers that do
    not want the DDP wrapper (e.g. CUDA-graph captured steps)."""
    if world == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n


def wrap_ddp(model, local_rank: int, device_type: str = "cuda"):
    from torch.nn.parallel import DistributedDataParallel as DDP
    if device_type == "cuda":
        return DDP(model, device_ids=[local_rank], output_device=local_rank, broadcast_buffers=True, gradient_as_bucket_view=True)
    return DDP(model)
