"""Data-parallel plumbing (reference: train.py:52-85, 147-148, 186-190, 292-294).

One process per GPU; the model is replicated and the batch is sharded by rank (no operator mixes samples:
LayerNorm, windows and the loss reductions are per sample, and the loss is a *sum* over the batch).  The
only exchange step is the bucketed gradient all-reduce, issued by torch DDP on the NCCL stream from the
autograd hooks that fire when each block's backward returns its parameter gradients -- i.e. overlapped with
the backward of the preceding blocks.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """env:// rendezvous from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (reference: train.py:52-66)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pins the calling process to the CPU cores next to its GPU (NVML's ideal affinity), so that the pinned host buffers
    it allocates afterwards are first-touched on that NUMA node and the per-step host -> device copies of 8 ranks do not
    cross the socket interconnect.  The reference gets the same effect from its SLURM launch line (`--cpu-bind`,
    submit_batch.sh).  Returns False when NVML is unavailable; never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def local_batch_size(global_batch: int, world: int) -> int:
    """reference: train.py:147-148 -- the global batch is split evenly; uneven splits are rejected."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    return global_batch // world


def shard_indices(n_samples: int, rank: int, world: int):
    """Contiguous per-rank shard of a sample index range (deterministic, no overlap, covers everything; an uneven split
    is rejected like `local_batch_size` does, reference train.py:147-148)."""
    if n_samples % world != 0:
        raise ValueError(f"{n_samples} samples are not divisible by world size {world}")
    per = n_samples // world
    return range(rank * per, (rank + 1) * per)


def wrap_ddp(model: torch.nn.Module, local_rank: Optional[int] = None, bucket_cap_mb: int = 100,
             static_graph: bool = False) -> torch.nn.Module:
    """DistributedDataParallel with settings that suit this model: gradients live in the bucket memory
    (no extra copy before NCCL reads them), large buckets (NVSwitch collectives are latency- not link-bound;
    `pos_embed` alone is 199 MB), no unused-parameter search."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP
    kw = dict(gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb, find_unused_parameters=False,
              static_graph=static_graph, broadcast_buffers=False)
    if next(model.parameters()).is_cuda:
        dev = torch.cuda.current_device() if local_rank is None else local_rank
        return DDP(model, device_ids=[dev], output_device=dev, **kw)
    return DDP(model, **kw)


def all_reduce_mean_scalar(x: torch.Tensor) -> torch.Tensor:
    """reference: train.py:292-294 (loss logging)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x)
        x = x / dist.get_world_size()
    return x


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
