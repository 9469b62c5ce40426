"""Throughput of the bilinear-lattice path: a 101 x 101 trapezoid (kappa = sigma_s = 0.5, like cfg3) with the analytic locator
(queue kernel, BILIN variant) and with the generic grid + point-in-polygon locator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rthx
from rthx import PolyVolume2D, RayTracingDomain2D

face = PolyVolume2D([(0.0, 0.0), (1.0, 0.0), (0.8, 1.0), (0.2, 1.0)], (True, True, True, True), 1, 0.5, 0.5)
face.epsilon = [1.0] * 4
face.T_in_w = [1000.0, 0.0, 0.0, 0.0]
face.T_in_g = -1.0
face.q_in_g = 0.0
rtm = RayTracingDomain2D([face], [(101, 101)])
flat = rthx.flatten_domain(rtm)
tr = rthx.DeviceTracer(flat, device=0)
print(tr.info)
for loc, rpe in ((0, 20000), (0, 94295), (1, 9430)):
    for it in range(3):
        out = tr.trace(rpe, seed=it, locator=loc, dense=False)
    st = out["stats"]
    print(f"locator={loc} rays={rpe * flat.n_elements:.3e} kernel {st['kernel_ms']:.3f} ms -> {rpe * flat.n_elements / st['kernel_ms'] / 1e-3:.4g} rays/s, lost {st['rays_lost']}, smem {st['smem_bytes']}")
