"""Not a pytest file: first-contact script for a GPU box (prints diagnostics instead of asserting)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rthx
from oracle import oracle

def nd(a, b):
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum() // 2)

for name, rtm, rpe in (("cfg1", rthx.meshes.cfg1(), 6060), ("cfg5", rthx.meshes.cfg5(), 4000),
                       ("twoq", rthx.meshes.two_quads_domain(kappa=(0.5, 3.0)), 20000),
                       ("skew", rthx.meshes.two_quads_domain(kappa=(1.0, 1.0), skew=0.3), 20000)):
    flat = rthx.flatten_domain(rtm)
    tr = rthx.DeviceTracer(flat, 0)
    ref = oracle.trace(flat, rpe, seed=5)
    for loc in (0, 1):
        got = tr.trace(rpe, seed=5, locator=loc)
        print(name, "locator", loc, "info", tr.info["n_affine_faces"], "rays", int(ref["counts"].sum()), "diff", nd(got["counts"], ref["counts"]),
              "lost gpu/ref", int(got["lost"].sum()), int(ref["lost"].sum()), "kernel_ms", round(got["stats"]["kernel_ms"], 3), flush=True)
rtm = rthx.meshes.cfg3(); flat = rthx.flatten_domain(rtm); tr = rthx.DeviceTracer(flat, 0)
print("fp64 peak TF", tr.measure_fp64_peak())
for rpe in (10000, 100000):
    for bt, ch in ((256, 0), (128, 0), (256, 1)):
        got = tr.trace(rpe, seed=1, block_threads=bt, row_chunks=ch)
        st = got["stats"]
        print("cfg3 rpe", rpe, "bt", bt, "chunks", st["row_chunks"], "blocks", st["n_blocks"], "kernel_ms", round(st["kernel_ms"], 2), "total_ms", round(st["total_ms"], 2),
              "rays/s", f"{st['rays_traced'] / st['kernel_ms'] * 1e3:.3e}", "lost", st["rays_lost"], flush=True)
