"""Multi-GPU plumbing: utterances are independent, so ranks take contiguous blocks of the utterance index and the
hot path moves no data between GPUs.  The one collective is an all-gather of the per-utterance score rows
(SURVEY.md 8-E).  Works with the `nccl` backend on GPUs and with `gloo` on CPU tensors (used by the tests)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the larger blocks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scores(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather score rows [n_local, K] of every rank into [n_total, K], in utterance order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    K = local.shape[1]
    base, rem = divmod(n_total, world)
    if rem == 0:
        out = torch.empty((n_total, K), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged shards: pad every block to the largest size, gather, then cut the padding out
    big = base + 1
    padded = torch.zeros((big, K), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * big, K), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(buf[r * big:r * big + (hi - lo)])
    return torch.cat(parts, dim=0)


def bind_to_gpu_numa(gpu_index: int) -> Optional[List[int]]:
    """Pin this process to the CPUs NVML reports as local to GPU `gpu_index`, so that pinned host buffers allocated
    afterwards (first touch) sit on the GPU's own NUMA node and host<->device copies do not cross sockets.  Only the
    host-buffer leg cares (the hot path never touches the host).  Returns the CPU list, or None if NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
