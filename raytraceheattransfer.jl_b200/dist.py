"""One-process-per-GPU sharding of the exchange-factor trace (torch.distributed is plumbing only).

Emitters (matrix rows) are independent (parallelRayTracing.jl:82,102-151), so rank r traces the rows
e ≡ r (mod world) — interleaving balances the cheap surface rows against the volume rows — into a full-size
UInt64 count matrix whose other rows stay zero; the per-rank matrices are then summed with ONE reduce to rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Because the RNG is counter-based and the sums are integer,
the reduced matrix is bit-identical for every world size.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def owned_emitters(n_elements: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, n_elements, world, dtype=np.int64)


def reduce_counts(counts: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor:
    """Sum per-rank count matrices onto `dst`.  `counts` is an int64 view of the UInt64 tallies (two's-complement
    addition is the same bit pattern), on the GPU for NCCL or on the CPU for gloo."""
    assert counts.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return counts


class ShardedTracer:
    """Device-resident sharded trace for one rank: owns the count / lost buffers on this rank's GPU."""

    def __init__(self, flat, device: Optional[int] = None, rank: Optional[int] = None, world: Optional[int] = None,
                 n_bins: int = 1):
        from ._lib import DeviceTracer
        self.rank = dist.get_rank() if rank is None and dist.is_initialized() else (rank or 0)
        self.world = dist.get_world_size() if world is None and dist.is_initialized() else (world or 1)
        self.device = torch.cuda.current_device() if device is None else device
        self.tracer = DeviceTracer(flat, device=self.device)
        N = self.tracer.n_elements
        self.N = N
        self.n_bins = n_bins
        dev = torch.device("cuda", self.device)
        self.counts = torch.zeros((n_bins, N, N), dtype=torch.int64, device=dev)
        self.lost = torch.zeros((n_bins, N), dtype=torch.int64, device=dev)

    def trace(self, rays_per_emitter: int, reduce: bool = True, **kw):
        """Enqueue zero + trace kernel on torch's current stream, then (optionally) the reduce to rank 0."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        st = self.tracer.trace_device(rays_per_emitter, self.counts.data_ptr(), self.lost.data_ptr(), stream=stream,
                                      zero_first=True, emitter_rank=self.rank, emitter_world=self.world, **kw)
        if reduce and self.world > 1:
            reduce_counts(self.counts)
            reduce_counts(self.lost)
        return st
