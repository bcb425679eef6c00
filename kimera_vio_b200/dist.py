"""Multi-GPU plumbing: camera streams shard across ranks with no data-path
collective; the only (optional) exchange is the gather of the fixed-size keypoint packets to rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).  `torch.distributed` is plumbing only."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def stream_shard(n_streams: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of global camera-stream ids owned by `rank` (one process per
    rank keeps the per-process landmark-id counter of FeatureDetector.cpp:141 per stream block)."""
    base, rem = divmod(n_streams, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_packets(local: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Gathers each rank's packed packet buffer (uint8, same size on every rank) on `dst`."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, out, dst=dst)
    return out


def max_over_ranks(value_ms: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
