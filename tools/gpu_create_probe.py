"""Time the pieces of the end-to-end call on the GPU box: rthx_create (mesh preparation + upload), trace, destroy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np
import rthx

flat = rthx.flatten_domain(rthx.meshes.cfg3())
for it in range(6):
    t0 = time.perf_counter()
    tr = rthx.DeviceTracer(flat, device=0)
    t1 = time.perf_counter()
    out = tr.trace(1000, seed=it)
    t2 = time.perf_counter()
    tr.close() if hasattr(tr, "close") else None
    del tr
    t3 = time.perf_counter()
    print(f"iter {it}: create {1e3 * (t1 - t0):.2f} ms, trace(1000 rays/emitter, host counts) {1e3 * (t2 - t1):.2f} ms, destroy {1e3 * (t3 - t2):.2f} ms")
