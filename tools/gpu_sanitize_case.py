"""Not a pytest file: a small workload for compute-sanitizer (memcheck / racecheck) runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rthx
for rtm, kw in ((rthx.meshes.cfg1(), dict(rec_ids=[9, 19, 150])), (rthx.meshes.circle_domain(16, 3), {}),
                (rthx.meshes.two_quads_domain(kappa=(0.5, 3.0), skew=0.2), {}), (rthx.meshes.cfg4(Ndim=7, n_bins=3), dict(bins=[0, 1, 2]))):
    flat = rthx.flatten_domain(rtm)
    tr = rthx.DeviceTracer(flat, 0)
    for loc in (0, 1):
        for mode in (0, 1):
            out = tr.trace(300, seed=3, locator=loc, mode=mode, **kw)
            assert np.all(out["counts"].sum(axis=2) + out["lost"] == 300)
    tr.close()
print("sanitize case OK")
