"""One-process-per-GPU sharding of the exchange-factor trace (torch.distributed is plumbing only).

Emitters (matrix rows) are independent (parallelRayTracing.jl:82,102-151), so rank r traces the rows
e ≡ r (mod world) — interleaving balances the cheap surface rows against the volume rows.  Two ways to land the
UInt64 count matrix on rank 0:

  mode="fused" (default for world > 1): rank 0 owns the matrix; every other rank maps it through CUDA IPC and its
      trace kernel flushes its own (disjoint) rows straight into rank 0's HBM with red.global.add.u64 over
      NVLink/NVSwitch, overlapped with tracing.  The "reduce" is fused into the kernel; only a barrier remains.
  mode="nccl": every rank fills a private full-size matrix (other rows zero) and ONE `reduce(SUM)` of the int64
      view sums them onto rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).

Because the RNG is counter-based and the sums are integer, both give bit-identical matrices for every world size.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ._abi import RTHX_ZERO_ALL, RTHX_ZERO_OWN_ROWS


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
    """Device-resident sharded trace for one rank.  After `trace()` returns on rank 0 (stream-ordered behind the
    barrier / reduce), `counts` [n_bins, N, N] and `lost` [n_bins, N] hold the whole job's tallies."""

    def __init__(self, flat, device: Optional[int] = None, rank: Optional[int] = None, world: Optional[int] = None,
                 n_bins: int = 1, mode: str = "fused"):
        from ._lib import DeviceTracer, SharedDeviceBuffer
        self.rank = dist.get_rank() if rank is None and dist.is_initialized() else (rank or 0)
        self.world = dist.get_world_size() if world is None and dist.is_initialized() else (world or 1)
        self.device = torch.cuda.current_device() if device is None else device
        self.tracer = DeviceTracer(flat, device=self.device)
        N = self.tracer.n_elements
        self.N = N
        self.n_bins = n_bins
        self.mode = mode if self.world > 1 else "local"
        dev = torch.device("cuda", self.device)
        n_counts, n_lost = n_bins * N * N, n_bins * N
        self.shared = None
        self.counts = self.lost = None
        if self.mode == "fused":
            hb = torch.zeros(64, dtype=torch.uint8, device=dev)
            if self.rank == 0:
                self.shared = SharedDeviceBuffer(self.device, n_counts + n_lost)
                hb.copy_(torch.tensor(list(self.shared.handle), dtype=torch.uint8))
            dist.broadcast(hb, src=0)
            if self.rank != 0:
                self.shared = SharedDeviceBuffer(self.device, n_counts + n_lost, handle=bytes(hb.cpu().tolist()))
            self.counts_ptr = self.shared.ptr
            self.lost_ptr = self.shared.ptr + 8 * n_counts
            if self.rank == 0:
                flat_t = torch.as_tensor(self.shared, device=dev)
                self.counts = flat_t[:n_counts].view(n_bins, N, N)
                self.lost = flat_t[n_counts:].view(n_bins, N)
                flat_t.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(device_ids=[self.device])
        else:
            self.counts = torch.zeros((n_bins, N, N), dtype=torch.int64, device=dev)
            self.lost = torch.zeros((n_bins, N), dtype=torch.int64, device=dev)
            self.counts_ptr = self.counts.data_ptr()
            self.lost_ptr = self.lost.data_ptr()

    def enqueue(self, rays_per_emitter: int, **kw):
        """Zero + trace kernel on torch's current stream.  In fused mode the matrix is shared, so a leading barrier
        (stream-ordered behind rank 0's pending reads of the previous result) keeps any rank from clearing its rows
        while rank 0 is still consuming them."""
        if self.mode == "fused":
            dist.barrier(device_ids=[self.device])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        zero = RTHX_ZERO_OWN_ROWS if self.mode == "fused" else RTHX_ZERO_ALL
        return self.tracer.trace_device(rays_per_emitter, self.counts_ptr, self.lost_ptr, stream=stream,
                                        zero_first=zero, emitter_rank=self.rank, emitter_world=self.world, **kw)

    def finish(self):
        """Make the whole job's tallies visible on rank 0: a barrier (fused) or the NCCL reduce."""
        if self.mode == "fused":
            dist.barrier(device_ids=[self.device])
        elif self.mode == "nccl":
            reduce_counts(self.counts)
            reduce_counts(self.lost)

    def trace(self, rays_per_emitter: int, **kw):
        st = self.enqueue(rays_per_emitter, **kw)
        self.finish()
        return st

    def close(self):
        self.counts = self.lost = None
        if self.shared is not None:
            if self.world > 1:
                torch.cuda.synchronize(self.device)
                dist.barrier(device_ids=[self.device])
            if self.rank != 0:
                self.shared.close()
            if self.world > 1:
                dist.barrier(device_ids=[self.device])
            if self.rank == 0:
                self.shared.close()
            self.shared = None
        self.tracer.close()
