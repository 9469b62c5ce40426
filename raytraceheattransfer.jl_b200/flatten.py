"""Flatten a RayTracingDomain2D into the SoA arrays of `rthx_mesh` (include/rthx.h).

This is the step the Julia shim performs on `rtm.coarse_mesh` / `rtm.fine_mesh` / `rtm.surface_mapping`
(DomainStructs.jl:89-130) before the `ccall`; indices become 0-based.  The flattening is redone on every
trace call because user code edits properties between construction and tracing (test/test_2d_grey.jl:199-201).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._abi import rthx_mesh, c_f64p, c_i32p, c_u8p


class FlatMesh:
    """Owns the numpy arrays and exposes a ctypes `rthx_mesh` view on them."""

    def __init__(self, **arrays):
        self.__dict__.update(arrays)
        m = rthx_mesh()
        m.n_coarse = int(self.n_coarse)
        m.n_cells = int(self.n_cells)
        m.n_bands = int(self.n_bands)
        m.n_surfaces = int(self.n_surfaces)
        for name, ptr in (("coarse_nv", c_i32p), ("coarse_vx", c_f64p), ("coarse_vy", c_f64p),
                          ("coarse_solid", c_u8p), ("fine_off", c_i32p), ("cell_nv", c_i32p),
                          ("cell_vx", c_f64p), ("cell_vy", c_f64p), ("cell_mid", c_f64p),
                          ("cell_volume", c_f64p), ("cell_surf_id", c_i32p), ("kappa", c_f64p),
                          ("sigma_s", c_f64p), ("epsilon", c_f64p), ("uniform_beta", c_f64p)):
            a = getattr(self, name)
            assert a.flags["C_CONTIGUOUS"]
            setattr(m, name, a.ctypes.data_as(ptr))
        self.c = m

    @property
    def n_elements(self) -> int:
        return int(self.n_surfaces + self.n_cells)


def flatten_domain(rtm) -> FlatMesh:
    nc = len(rtm.coarse_mesh)
    nb = rtm.n_spectral_bins
    n_cells = rtm.num_volumes
    ns = rtm.num_surfaces
    coarse_nv = np.zeros(nc, np.int32)
    coarse_vx = np.zeros((nc, 4))
    coarse_vy = np.zeros((nc, 4))
    coarse_solid = np.zeros((nc, 4), np.uint8)
    fine_off = np.zeros(nc + 1, np.int32)
    cell_nv = np.zeros(n_cells, np.int32)
    cell_vx = np.zeros((n_cells, 4))
    cell_vy = np.zeros((n_cells, 4))
    cell_mid = np.zeros((n_cells, 2))
    cell_volume = np.zeros(n_cells)
    cell_surf_id = np.full((n_cells, 4), -1, np.int32)
    kappa = np.zeros((nb, n_cells))
    sigma_s = np.zeros((nb, n_cells))
    epsilon = np.zeros((nb, max(ns, 1)))
    g = 0
    for c, face in enumerate(rtm.coarse_mesh):
        nv = len(face.vertices)
        coarse_nv[c] = nv
        for i, v in enumerate(face.vertices):
            coarse_vx[c, i], coarse_vy[c, i] = v
            coarse_solid[c, i] = 1 if face.solidWalls[i] else 0
        fine_off[c] = g
        for f, cell in enumerate(rtm.fine_mesh[c]):
            n = len(cell.vertices)
            cell_nv[g] = n
            for i, v in enumerate(cell.vertices):
                cell_vx[g, i], cell_vy[g, i] = v
            cell_mid[g] = cell.midPoint
            cell_volume[g] = cell.volume
            for w in range(n):
                sid = rtm.surface_mapping.get((c + 1, f + 1, w + 1))
                if sid is not None:
                    cell_surf_id[g, w] = sid - 1
                    for b in range(nb):
                        epsilon[b, sid - 1] = cell.eps(w, b)
            if isinstance(cell.kappa_g, list):
                kappa[:, g] = cell.kappa_g
                sigma_s[:, g] = cell.sigma_s_g
            else:
                kappa[:, g] = cell.kappa_g
                sigma_s[:, g] = cell.sigma_s_g
            assert rtm.volume_mapping[(c + 1, f + 1)] == g + 1
            g += 1
    fine_off[nc] = g
    return FlatMesh(n_coarse=nc, n_cells=n_cells, n_bands=nb, n_surfaces=ns,
                    coarse_nv=coarse_nv, coarse_vx=coarse_vx, coarse_vy=coarse_vy, coarse_solid=coarse_solid,
                    fine_off=fine_off, cell_nv=cell_nv, cell_vx=cell_vx, cell_vy=cell_vy, cell_mid=cell_mid,
                    cell_volume=cell_volume, cell_surf_id=cell_surf_id, kappa=kappa, sigma_s=sigma_s,
                    epsilon=epsilon, uniform_beta=np.asarray(rtm.uniform_across_bin, dtype=np.float64).copy())
