"""Not a pytest file: kernel tuning sweep on a GPU box (prints kernel times for cfg3 / cfg5 variants)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rthx
from oracle import oracle

def nd(a, b):
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum() // 2)

cases = {"cfg3": (rthx.meshes.cfg3(), 100000), "cfg5": (rthx.meshes.cfg5(), 400000), "cfg1": (rthx.meshes.cfg1(), 6060)}
for name, (rtm, rpe) in cases.items():
    flat = rthx.flatten_domain(rtm)
    tr = rthx.DeviceTracer(flat, 0)
    if name != "cfg3":
        small = min(rpe, 6060)
        ref = oracle.trace(flat, small, seed=5)
        for loc in (0, 1):
            got = tr.trace(small, seed=5, locator=loc)
            print(name, "parity locator", loc, "diff", nd(got["counts"], ref["counts"]), "lost", int(got["lost"].sum()), int(ref["lost"].sum()), flush=True)
    for minb in ("2", "3", "4"):
        os.environ["RTHX_MINB"] = minb
        for bt in (128, 256):
            best = None
            for rep in range(3):
                got = tr.trace(rpe, seed=1 + rep, block_threads=bt)
                st = got["stats"]
                best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
            print(name, "minb", minb, "bt", bt, "chunks", st["row_chunks"], "blocks", st["n_blocks"], "kernel_ms", round(best, 3),
                  "rays/s", f"{st['rays_traced'] / best * 1e3:.3e}", flush=True)
