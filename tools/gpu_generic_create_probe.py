"""First-use cost of the generic-locator tables on one GPU: a fresh mesh (tables built + uploaded), a second handle on the same
mesh (tables from the process-wide cache, upload only), and the analytic path for reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rthx

for name in ("cfg3", "cfg5"):
    flat = rthx.flatten_domain(getattr(rthx.meshes, name)())
    rthx.DeviceTracer(flat, device=0).trace(10, seed=1, dense=False)           # warm the context / resource pool
    for tag, loc in (("analytic", 0), ("generic first", 1), ("generic cached", 1)):
        t0 = time.perf_counter()
        tr = rthx.DeviceTracer(flat, device=0)
        t1 = time.perf_counter()
        tr.trace(10, seed=1, locator=loc, dense=False)
        t2 = time.perf_counter()
        tr.close()
        print(f"{name} {tag:15s} create {1e3 * (t1 - t0):7.2f} ms, first trace of 10 rays/emitter {1e3 * (t2 - t1):7.2f} ms")
