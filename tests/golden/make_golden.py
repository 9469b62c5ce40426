#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: seeded tallies of the CPU oracle (oracle/rthx_oracle.c) on small meshes.

The reference itself cannot run here (pure Julia, unseeded RNG, no golden F matrices in its tests), so these fixtures
pin the *oracle + RNG contract*: `tests/test_golden.py` checks that the oracle still reproduces them on the CPU and
that the CUDA path reproduces them on the GPU.  The reference's own golden vectors for this path (the Crosbie &
Schrenker table, the circle-centre temperature) are checked in tests/test_oracle_known_answers.py.

    python tests/golden/make_golden.py          # rewrites the fixtures (only after a deliberate contract change)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import rthx  # noqa: E402
from oracle import oracle  # noqa: E402


def cases():
    yield "cfg1_square11", rthx.meshes.cfg1(), dict(rpe=300, seed=0x5EED0001, bins=[0], mode=0, rec_ids=[9, 19, 29])
    yield "circle16x3", rthx.meshes.circle_domain(16, 3), dict(rpe=300, seed=7, bins=[0], mode=0, rec_ids=None)
    yield "two_quads_variable_beta", rthx.meshes.two_quads_domain(kappa=(0.5, 3.0)), dict(rpe=400, seed=8, bins=[0], mode=0, rec_ids=None)
    yield "spectral3_square7", rthx.meshes.cfg4(Ndim=7, n_bins=3), dict(rpe=200, seed=9, bins=[2, 0], mode=0, rec_ids=None)
    mb = rthx.meshes.square_domain(6, kappa=0.6, sigma_s=0.9, epsilon=(0.3, 0.6, 0.9, 0.5))
    yield "multibounce_square6", mb, dict(rpe=300, seed=10, bins=[0], mode=1, rec_ids=None)
    yield "multibounce_specular_square6", mb, dict(rpe=300, seed=10, bins=[0], mode=2, rec_ids=None)


def run(rtm, kw):
    flat = rthx.flatten_domain(rtm)
    return flat, oracle.trace(flat, kw["rpe"], seed=kw["seed"], bins=kw["bins"], mode=kw["mode"], rec_ids=kw["rec_ids"], n_threads=1)


if __name__ == "__main__":
    for name, rtm, kw in cases():
        flat, out = run(rtm, kw)
        b, i, j = np.nonzero(out["counts"])
        extra = {}
        if kw["rec_ids"]:
            extra = dict(origins=out["origins"], endpoints=out["endpoints"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), n_elements=flat.n_elements, band=b.astype(np.int32),
                            row=i.astype(np.int32), col=j.astype(np.int32),
                            count=out["counts"][b, i, j].astype(np.uint32), lost=out["lost"].astype(np.uint32), **extra)
        print(name, flat.n_elements, int(out["counts"].sum()), "rays,", len(b), "non-zeros")
