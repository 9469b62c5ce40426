"""One process, every visible GPU: cfg3 rows traced on all devices and gathered on device 0 over NVLink (rthx_trace_exchange_multi with
counts_out = NULL) — the command the NVLink counters of profiles/ are taken on.  Prints the rate and checks the gathered matrix."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rthx
from rthx._lib import create_multi, trace_multi, device_count

n = device_count()
flat = rthx.flatten_domain(rthx.meshes.cfg3())
N = flat.n_elements
rpe = int(1e9) // N
trs = create_multi(flat, list(range(n)))
for rep in range(3):
    t0 = time.perf_counter()
    out = trace_multi(trs, rpe, dense=False, seed=100 + rep)
    dt = time.perf_counter() - t0
nnz, chi = trs[0].counts_stats(0)
print(f"{n} GPUs, {rpe * N} rays: {1e3 * dt:.2f} ms ({rpe * N / dt:.3e} rays/s), kernel {out['stats']['kernel_ms']:.2f} ms, nnz {nnz}, lost {int(out['lost'].sum())}")
row_ptr, cols, vals, _ = trs[0].counts_csr(0)
assert int(vals.sum()) + int(out["lost"].sum()) == rpe * N
