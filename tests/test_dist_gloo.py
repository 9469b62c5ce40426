"""world_size-2 gloo test of the sharding plumbing (partition by emitter row + one integer reduce), on CPU.
The per-rank trace is played by the CPU oracle here (test infrastructure); on GPUs the same partition arguments go
to rthx_trace_exchange_device and the reduce runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rthx
    from rthx.dist import owned_emitters, reduce_counts
    from oracle import oracle
    flat = rthx.flatten_domain(rthx.meshes.square_domain(6, kappa=1.0, sigma_s=0.5))
    N = flat.n_elements
    mine = owned_emitters(N, rank, world)
    assert np.array_equal(mine, np.arange(rank, N, world))
    part = oracle.trace(flat, 500, seed=77, emitter_rank=rank, emitter_world=world, n_threads=1)
    c = part["counts"]
    other = np.setdiff1d(np.arange(N), mine)
    assert c[:, other, :].sum() == 0                               # rows of other ranks stay zero
    t = torch.from_numpy(c.view(np.int64).copy())
    reduce_counts(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reduce_is_bit_identical(tmp_path, oracle_mod, rthx_mod):
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    reduced = np.load(out)
    flat = rthx_mod.flatten_domain(rthx_mod.meshes.square_domain(6, kappa=1.0, sigma_s=0.5))
    full = oracle_mod.trace(flat, 500, seed=77)["counts"]
    assert np.array_equal(reduced, full)
