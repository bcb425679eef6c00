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


class PacketGather:
    """The optional gather of the keypoint packets to rank 0, without a pack kernel: the context's
    finalize kernel assembles the packets of every step directly in this rank's send buffer
    (kvfe_frontend_bind_packets), and the NCCL gather is enqueued on the context's own CUDA stream right behind
    the step graph, so it overlaps nothing it should not and needs no host synchronisation."""

    def __init__(self, ctx, dst: int = 0):
        self.ctx, self.dst = ctx, dst
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        nbytes = ctx.B * ctx.packet_bytes
        self.send = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        self.recv = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(self.world)] if self.rank == dst else None
        ctx.bind_packets(self.send.data_ptr())
        self.stream = torch.cuda.ExternalStream(int(ctx.lib.kvfe_cuda_stream(ctx.h)))

    def gather(self):
        """Enqueue the gather of the last step's packets (returns rank 0's list of per-rank buffers, or None)."""
        if self.world == 1:
            return [self.send]
        with torch.cuda.stream(self.stream):
            dist.gather(self.send, self.recv, dst=self.dst)
        return self.recv

    def close(self):
        self.ctx.bind_packets(0)
