"""One-process-per-GPU sharding of the exchange-factor trace (torch.distributed is plumbing only).

Emitters (matrix rows) are independent (parallelRayTracing.jl:82,102-151), so rank r traces the rows
e ≡ r (mod world) — interleaving balances the cheap surface rows against the volume rows.  Two ways to land the
UInt64 count matrix on rank 0:

  mode="fused" (default for world > 1): rank 0 owns the matrix; every other rank maps it through CUDA IPC and its
      trace kernel hands its own (disjoint) finished rows over to rank 0's HBM with plain coalesced stores over
      NVLink/NVSwitch (RTHX_DEST_PEER), overlapped with tracing.  The "reduce" is fused into the kernel; the ranks synchronise through
      device-side step flags in rank 0's memory (no host-launched collective per step).
  mode="nccl": every rank fills a private full-size matrix (other rows zero) and ONE `reduce(SUM)` of the int64
      view sums them onto rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).

Because the RNG is counter-based and the sums are integer, both give bit-identical matrices for every world size.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ._abi import RTHX_ZERO_ALL, RTHX_ZERO_OWN_ROWS, RTHX_DEST_PEER


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
    flag wait / reduce), `counts` [n_bins, N, N] and `lost` [n_bins, N] hold the whole job's tallies.

    Fused mode keeps TWO matrices on rank 0 and alternates between them, and the ranks synchronise through 64-bit step
    counters in rank 0's memory (rthx_flag_signal / rthx_flag_wait): a rank signals "my rows of step s have landed" with a
    system-scope release store behind its trace kernel, rank 0's stream spins until every rank has signalled, and a rank may
    overwrite a matrix only after rank 0 has signalled that its previous contents were consumed.  No host-launched collective
    and no host synchronisation remains inside a step; consumers of `counts` must run on the stream the trace was enqueued on."""

    N_FLAGS = 64

    def __init__(self, flat, device: Optional[int] = None, rank: Optional[int] = None, world: Optional[int] = None,
                 n_bins: int = 1, mode: str = "fused"):
        from ._lib import DeviceTracer, SharedDeviceBuffer, load_library
        self._L = load_library()
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
        self.step = 0
        if self.mode == "fused":
            assert self.world + 2 <= self.N_FLAGS
            per = n_counts + n_lost
            hb = torch.zeros(64, dtype=torch.uint8, device=dev)
            if self.rank == 0:
                self.shared = SharedDeviceBuffer(self.device, 2 * per + self.N_FLAGS)
                hb.copy_(torch.tensor(list(self.shared.handle), dtype=torch.uint8))
            dist.broadcast(hb, src=0)
            if self.rank != 0:
                self.shared = SharedDeviceBuffer(self.device, 2 * per + self.N_FLAGS, handle=bytes(hb.cpu().tolist()))
            base = self.shared.ptr
            self._counts_ptrs = [base + 8 * k * per for k in (0, 1)]
            self._lost_ptrs = [base + 8 * (k * per + n_counts) for k in (0, 1)]
            self._flags_ptr = base + 8 * 2 * per          # [0..world): done(rank), [world]: consumed, [world+1]: wait timeouts
            self._views = None
            if self.rank == 0:
                flat_t = torch.as_tensor(self.shared, device=dev)
                flat_t.zero_()
                self._views = [(flat_t[k * per:k * per + n_counts].view(n_bins, N, N),
                                flat_t[k * per + n_counts:(k + 1) * per].view(n_bins, N)) for k in (0, 1)]
                self._flags = flat_t[2 * per:]
                self.counts, self.lost = self._views[0]
            torch.cuda.synchronize(dev)
            dist.barrier(device_ids=[self.device])
        else:
            self.counts = torch.zeros((n_bins, N, N), dtype=torch.int64, device=dev)
            self.lost = torch.zeros((n_bins, N), dtype=torch.int64, device=dev)
            self.counts_ptr = self.counts.data_ptr()
            self.lost_ptr = self.lost.data_ptr()

    def _check(self, rc):
        if rc != 0:
            from ._lib import RthxError
            msg = self._L.rthx_last_error(None)
            raise RthxError(f"rthx flag call failed ({rc}): {msg.decode() if msg else ''}")

    def enqueue(self, rays_per_emitter: int, **kw):
        """Zero + trace kernel on torch's current stream (asynchronous)."""
        import ctypes as C
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.mode != "fused":
            return self.tracer.trace_device(rays_per_emitter, self.counts_ptr, self.lost_ptr, stream=stream,
                                            zero_first=RTHX_ZERO_ALL, emitter_rank=self.rank, emitter_world=self.world, **kw)
        s, W = self.step, self.world
        b = s & 1
        fl = self._flags_ptr
        if self.rank == 0 and s >= 1:
            # everything already enqueued on this stream has consumed the result of step s-1: release its matrix
            self._check(self._L.rthx_flag_signal(C.c_void_p(fl + 8 * W), s, C.c_void_p(stream)))
        if s >= 2:
            # matrix b held step s-2: wait until rank 0 has released it (consumed counter >= s-1)
            self._check(self._L.rthx_flag_wait(C.c_void_p(fl + 8 * W), 1, s - 1, 30.0, C.c_void_p(fl + 8 * (W + 1)), C.c_void_p(stream)))
        # ranks other than 0 see the matrix through a CUDA-IPC mapping: say so, the library then stages rows locally and hands them over
        st = self.tracer.trace_device(rays_per_emitter, self._counts_ptrs[b], self._lost_ptrs[b], stream=stream,
                                      zero_first=RTHX_ZERO_OWN_ROWS | (RTHX_DEST_PEER if self.rank != 0 else 0),
                                      emitter_rank=self.rank, emitter_world=self.world, **kw)
        self._check(self._L.rthx_flag_signal(C.c_void_p(fl + 8 * self.rank), s + 1, C.c_void_p(stream)))
        return st

    def finish(self):
        """Make the whole job's tallies visible on rank 0 (stream-ordered): the flag wait (fused) or the NCCL reduce."""
        import ctypes as C
        if self.mode == "fused":
            s, W = self.step, self.world
            if self.rank == 0:
                stream = torch.cuda.current_stream(self.device).cuda_stream
                fl = self._flags_ptr
                self._check(self._L.rthx_flag_wait(C.c_void_p(fl), W, s + 1, 30.0, C.c_void_p(fl + 8 * (W + 1)), C.c_void_p(stream)))
                self.counts, self.lost = self._views[s & 1]
            self.step = s + 1
        elif self.mode == "nccl":
            reduce_counts(self.counts)
            reduce_counts(self.lost)

    def wait_errors(self) -> int:
        """Number of flag waits that timed out (rank 0; synchronises the device)."""
        if self.mode != "fused" or self.rank != 0:
            return 0
        torch.cuda.synchronize(self.device)
        return int(self._flags[self.world + 1].item())

    def trace(self, rays_per_emitter: int, **kw):
        st = self.enqueue(rays_per_emitter, **kw)
        self.finish()
        return st

    def close(self):
        self.counts = self.lost = None
        self._views = None
        self._flags = None
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
