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


class _Geometry:
    """The part of the flattening that only depends on the mesh geometry and numbering (vertices, midpoints, areas, solid-wall
    numbering): built once per domain and reused while the mesh is the same set of cell objects with the same vertices — what user
    code edits between construction and tracing are the PROPERTIES (kappa, sigma_s, epsilon, temperatures; test/test_2d_grey.jl:199-201),
    and those are re-read on every call."""

    def __init__(self, rtm):
        nc = len(rtm.coarse_mesh)
        n_cells = rtm.num_volumes
        self.cells = [cell for fine in rtm.fine_mesh for cell in fine]        # walk order of createIndexMapping2D.jl:7-18
        assert len(self.cells) == n_cells
        self.coarse_nv = np.zeros(nc, np.int32)
        self.coarse_vx = np.zeros((nc, 4))
        self.coarse_vy = np.zeros((nc, 4))
        self.coarse_solid = np.zeros((nc, 4), np.uint8)
        self.fine_off = np.zeros(nc + 1, np.int32)
        g = 0
        for c, face in enumerate(rtm.coarse_mesh):
            self.coarse_nv[c] = len(face.vertices)
            for i, v in enumerate(face.vertices):
                self.coarse_vx[c, i], self.coarse_vy[c, i] = v
                self.coarse_solid[c, i] = 1 if face.solidWalls[i] else 0
            self.fine_off[c] = g
            g += len(rtm.fine_mesh[c])
        self.fine_off[nc] = g
        cells = self.cells
        self.cell_nv = np.fromiter((len(c.vertices) for c in cells), np.int32, n_cells)
        pad = (0.0, 0.0)
        verts = np.array([tuple(c.vertices) if len(c.vertices) == 4 else tuple(c.vertices) + (pad,) for c in cells], dtype=np.float64)
        self.cell_vx = np.ascontiguousarray(verts[:, :, 0])
        self.cell_vy = np.ascontiguousarray(verts[:, :, 1])
        self.cell_mid = np.array([c.midPoint for c in cells], dtype=np.float64).reshape(n_cells, 2)
        self.cell_volume = np.fromiter((c.volume for c in cells), np.float64, n_cells)
        self.cell_surf_id = np.full((n_cells, 4), -1, np.int32)
        self.surfs = []                                                       # (surface index, cell, wall) of every solid fine wall
        for (c, f, w), sid in rtm.surface_mapping.items():
            gidx = int(self.fine_off[c - 1]) + f - 1
            self.cell_surf_id[gidx, w - 1] = sid - 1
            self.surfs.append((sid - 1, cells[gidx], w - 1))
        for (c, f), vid in rtm.volume_mapping.items():
            assert int(self.fine_off[c - 1]) + f == vid
        self.key = self._key(rtm)

    @staticmethod
    def _key(rtm):
        cells = [fine[k] for fine in rtm.fine_mesh for k in (0, len(fine) // 2, len(fine) - 1)]
        return (tuple(len(fine) for fine in rtm.fine_mesh), tuple(id(c) for c in cells),
                tuple(tuple(map(tuple, c.vertices)) for c in cells), len(rtm.surface_mapping))


def flatten_domain(rtm) -> FlatMesh:
    geo = getattr(rtm, "_flat_geometry", None)
    if geo is None or geo.key != _Geometry._key(rtm):
        geo = _Geometry(rtm)
        rtm._flat_geometry = geo
    nb = rtm.n_spectral_bins
    n_cells = rtm.num_volumes
    ns = rtm.num_surfaces
    cells = geo.cells
    # properties: re-read on every call
    if cells and isinstance(cells[0].kappa_g, list):
        kappa = np.ascontiguousarray(np.array([c.kappa_g for c in cells], dtype=np.float64).reshape(n_cells, nb).T)
        sigma_s = np.ascontiguousarray(np.array([c.sigma_s_g for c in cells], dtype=np.float64).reshape(n_cells, nb).T)
    else:
        kappa = np.fromiter((c.kappa_g for c in cells), np.float64, n_cells).reshape(1, n_cells).repeat(nb, axis=0)
        sigma_s = np.fromiter((c.sigma_s_g for c in cells), np.float64, n_cells).reshape(1, n_cells).repeat(nb, axis=0)
    epsilon = np.zeros((nb, max(ns, 1)))
    for sid, cell, w in geo.surfs:
        for b in range(nb):
            epsilon[b, sid] = cell.eps(w, b)
    return FlatMesh(n_coarse=len(rtm.coarse_mesh), n_cells=n_cells, n_bands=nb, n_surfaces=ns,
                    coarse_nv=geo.coarse_nv, coarse_vx=geo.coarse_vx, coarse_vy=geo.coarse_vy, coarse_solid=geo.coarse_solid,
                    fine_off=geo.fine_off, cell_nv=geo.cell_nv, cell_vx=geo.cell_vx, cell_vy=geo.cell_vy, cell_mid=geo.cell_mid,
                    cell_volume=geo.cell_volume, cell_surf_id=geo.cell_surf_id, kappa=np.ascontiguousarray(kappa),
                    sigma_s=np.ascontiguousarray(sigma_s), epsilon=epsilon,
                    uniform_beta=np.asarray(rtm.uniform_across_bin, dtype=np.float64).copy())
