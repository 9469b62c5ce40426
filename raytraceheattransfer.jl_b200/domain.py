"""Host-side data model: the Python twin of the reference's Julia structs for the `:exchange` path.

Julia is not available in this image, so the host side above the C ABI mirrors the reference's
operator interface in Python (same names, argument meaning and conventions):

  PolyVolume2D          src/Domains/domains/PolyVolume2D.jl:2-184, DomainStructs.jl:2-44
  meshQuad              src/Meshing/meshing/meshQuad.jl:75-182
  meshTriangle          src/Meshing/meshing/meshTriangle.jl:2-103
  RayTracingDomain2D    src/Domains/domains/RayTracingDomain2D.jl:2-155, IntermediateMesh2D.jl:2-56
  RayRecorder           DomainStructs.jl:176-181, parallelRayTracing.jl:194-200

Only what the tracer reads is kept (geometry, solidWalls, kappa/sigma_s/epsilon, index maps, spectral
flags) plus the boundary-condition fields the validation solver in tests needs.  Indices exposed to the
user stay 1-based like the reference (RayRecorder ids, surface/volume mappings); the flattener in
flatten.py converts to the 0-based ABI.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np


def _as_points(verts) -> List[Tuple[float, float]]:
    return [(float(v[0]), float(v[1])) for v in verts]


class PolyVolume2D:
    """A convex triangle or quadrilateral with CCW vertices (PolyVolume2D.jl:2-23 quad, :96-114 tri).

    kappa_g / sigma_s_g / epsilon are scalars-per-element in grey mode and per-band lists in
    spectral mode (n_spectral_bins > 1), like the reference's `Union{G,Vector{G}}` fields.
    """

    __slots__ = ("vertices", "solidWalls", "midPoint", "volume", "area", "subVolumes",
                 "n_spectral_bins", "kappa_g", "sigma_s_g", "epsilon",
                 "T_in_w", "T_in_g", "q_in_g", "q_in_w", "T_g", "T_w",
                 # result fields written by solveEquilibrium (DomainStructs.jl:22-43)
                 "j_g", "g_a_g", "e_g", "r_g", "g_g", "i_g", "q_g",
                 "j_w", "g_a_w", "e_w", "r_w", "g_w", "i_w", "q_w")

    def __init__(self, vertices, solidWalls, n_spectral_bins: int = 1,
                 kappa_default: float = 0.0, sigma_s_default: float = 0.0):
        p = _as_points(vertices)
        n = len(p)
        if n not in (3, 4):
            raise ValueError("Only triangles and quadrilaterals are supported.")  # IntermediateMesh2D.jl:14
        if len(solidWalls) != n:
            raise ValueError("solidWalls must have one flag per wall")
        self.vertices = p
        self.solidWalls = [bool(b) for b in solidWalls]
        if n == 4:
            # PolyVolume2D.jl:9 — (p1+p2+p3+p4)/4, summed left to right
            self.midPoint = ((((p[0][0] + p[1][0]) + p[2][0]) + p[3][0]) / 4,
                             (((p[0][1] + p[1][1]) + p[2][1]) + p[3][1]) / 4)
            # PolyVolume2D.jl:20-21 — two shoelace triangles ABC + CDA
            self.volume = (0.5 * (p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1])
                                  + p[2][0] * (p[0][1] - p[1][1]))
                           + 0.5 * (p[2][0] * (p[3][1] - p[0][1]) + p[3][0] * (p[0][1] - p[2][1])
                                    + p[0][0] * (p[2][1] - p[3][1])))
        else:
            # PolyVolume2D.jl:103,112
            self.midPoint = (((p[0][0] + p[1][0]) + p[2][0]) / 3, ((p[0][1] + p[1][1]) + p[2][1]) / 3)
            self.volume = 0.5 * (p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1])
                                 + p[2][0] * (p[0][1] - p[1][1]))
        self.area = [math.hypot(p[i][0] - p[(i + 1) % n][0], p[i][1] - p[(i + 1) % n][1])
                     for i in range(n)]
        self.subVolumes: List["PolyVolume2D"] = []
        self.n_spectral_bins = int(n_spectral_bins)
        if n_spectral_bins == 1:
            self.kappa_g = float(kappa_default)
            self.sigma_s_g = float(sigma_s_default)
            self.epsilon = [0.0] * n
        else:
            self.kappa_g = [float(kappa_default)] * n_spectral_bins
            self.sigma_s_g = [float(sigma_s_default)] * n_spectral_bins
            self.epsilon = [[0.0] * n_spectral_bins for _ in range(n)]
        self.T_in_w = [0.0] * n
        self.q_in_w = [0.0] * n
        self.T_in_g = 0.0
        self.q_in_g = 0.0
        self.T_g = 0.0
        self.T_w = [0.0] * n
        self.j_g = self.g_a_g = self.e_g = self.r_g = self.g_g = self.i_g = self.q_g = 0.0
        self.j_w, self.g_a_w, self.e_w, self.r_w = [0.0] * n, [0.0] * n, [0.0] * n, [0.0] * n
        self.g_w, self.i_w, self.q_w = [0.0] * n, [0.0] * n, [0.0] * n

    # -- helpers -------------------------------------------------------------------------------
    def beta(self, band: int = 0) -> float:
        """kappa + sigma_s of this element for a 0-based band (traceRay.jl:7-11, 96-100)."""
        if isinstance(self.kappa_g, list):
            return self.kappa_g[band] + self.sigma_s_g[band]
        return self.kappa_g + self.sigma_s_g

    def eps(self, wall: int, band: int = 0) -> float:
        e = self.epsilon[wall]
        return e[band] if isinstance(e, list) else e


def _copy_prop(v):
    return list(v) if isinstance(v, list) else v


def _containing_edge(sup: PolyVolume2D, p):
    """addSubVolume.jl:43-56 — nearest edge of the super-volume to point p."""
    nV = len(sup.vertices)
    best_d, best_k = math.inf, 0
    for k in range(nV):
        a = sup.vertices[k]
        b = sup.vertices[(k + 1) % nV]
        abx, aby = b[0] - a[0], b[1] - a[1]
        t = ((p[0] - a[0]) * abx + (p[1] - a[1]) * aby) / (abx * abx + aby * aby)
        t = min(1.0, max(0.0, t))
        d = math.hypot(p[0] - (a[0] + t * abx), p[1] - (a[1] + t * aby))
        if d < best_d:
            best_d, best_k = d, k
    return best_k, best_d


def addSubVolume(sup: PolyVolume2D, sub: PolyVolume2D) -> None:
    """Property inheritance parent -> sub-cell (addSubVolume.jl:2-40).

    Volume properties are copied (:8-19); a solid sub-wall inherits the wall properties of the
    super-volume edge it lies on (:24-35).  solidWalls themselves are NOT changed here.
    """
    sub.kappa_g = _copy_prop(sup.kappa_g)
    sub.sigma_s_g = _copy_prop(sup.sigma_s_g)
    sub.T_in_g = sup.T_in_g
    sub.q_in_g = sup.q_in_g
    nS = len(sup.vertices)
    charlen = max(sup.area)
    n = len(sub.vertices)
    for i in range(n):
        if not sub.solidWalls[i]:
            continue
        a, b = sub.vertices[i], sub.vertices[(i + 1) % n]
        m = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)
        k, d = _containing_edge(sup, m)
        if d < 1e-8 * charlen and sup.solidWalls[k]:
            sub.epsilon[i] = _copy_prop(sup.epsilon[k])
            sub.T_in_w[i] = sup.T_in_w[k]
            sub.q_in_w[i] = sup.q_in_w[k]
    sup.subVolumes.append(sub)


def _n_bins_of(vol: PolyVolume2D) -> int:
    return len(vol.kappa_g) if isinstance(vol.kappa_g, list) else 1


def meshQuad(volume: PolyVolume2D, Nx: int, Ny: int) -> PolyVolume2D:
    """Bilinear sub-meshing of a quadrilateral (meshQuad.jl:75-182).

    Lattice point formulas follow :107-134 operation by operation so the vertices are bit-identical to
    the reference's; cells are created m (A->D direction) outer, n (A->B) inner (:139,:151); vertices of a
    cell are (n,m),(n+1,m),(n+1,m+1),(n,m+1) (:167-170).  The solid-wall inheritance keeps the reference's
    `if m==1 ... elseif m==Ny` / `if n==1 ... elseif n==Nx` structure (:145-161), i.e. with Ny==1 wall 3
    and with Nx==1 wall 2 are never solid.
    """
    is_spectral = isinstance(volume.kappa_g, list)
    n_bins = len(volume.kappa_g) if is_spectral else 1
    kappa_default = volume.kappa_g[0] if is_spectral else volume.kappa_g
    sigma_default = volume.sigma_s_g[0] if is_spectral else volume.sigma_s_g
    A, B, C, D = volume.vertices
    xs = (A[0], B[0], C[0], D[0], A[0])
    ys = (A[1], B[1], C[1], D[1], A[1])
    deltaXbot = xs[1] - xs[0]
    deltaXtop = xs[3] - xs[2]
    deltaXleft = xs[4] - xs[3]
    deltaYBot = ys[0] - ys[1]
    deltaYRight = ys[1] - ys[2]
    deltaYLeft = ys[3] - ys[0]
    xP = [[0.0] * (Ny + 1) for _ in range(Nx + 1)]
    yP = [[0.0] * (Ny + 1) for _ in range(Nx + 1)]
    for m in range(1, Ny + 2):
        refmoveXleft = (m - 1) * deltaXleft / Ny
        refmoveXright = deltaXbot - (m - 1) * (deltaXbot + deltaXtop) / Ny
        for n in range(1, Nx + 2):
            refmoveYdown = (n - 1) * deltaYBot / Nx
            refmoveYup = deltaYLeft - (n - 1) * (deltaYLeft + deltaYRight) / Nx
            xP[n - 1][m - 1] = xs[0] - refmoveXleft + (n - 1) * refmoveXright / Nx
            yP[n - 1][m - 1] = ys[0] - refmoveYdown + (m - 1) * refmoveYup / Ny
    for m in range(1, Ny + 1):
        sw = [False, False, False, False]
        if m == 1:
            sw[0] = volume.solidWalls[0]
        elif m == Ny:
            sw[2] = volume.solidWalls[2]
        for n in range(1, Nx + 1):
            sw = [sw[0], False, sw[2], False]
            if n == 1:
                sw[3] = volume.solidWalls[3]
            elif n == Nx:
                sw[1] = volume.solidWalls[1]
            pts = [(xP[n - 1][m - 1], yP[n - 1][m - 1]), (xP[n][m - 1], yP[n][m - 1]),
                   (xP[n][m], yP[n][m]), (xP[n - 1][m], yP[n - 1][m])]
            sub = PolyVolume2D(pts, sw, n_bins, kappa_default, sigma_default)
            addSubVolume(volume, sub)
    return volume


def meshTriangle(face: PolyVolume2D, Ndim: int) -> PolyVolume2D:
    """Triangle sub-meshing (meshTriangle.jl:2-103): mirror the vertex opposite the longest edge through that
    edge's midpoint (:15-42), mesh the resulting parallelogram with meshQuad (:58-61), keep the cells on the
    triangle's side (:89-97) and replace each cell on the cut diagonal by a triangle (:73-86)."""
    is_spectral = isinstance(face.kappa_g, list)
    n_bins = len(face.kappa_g) if is_spectral else 1
    kappa_default = face.kappa_g[0] if is_spectral else face.kappa_g
    sigma_default = face.sigma_s_g[0] if is_spectral else face.sigma_s_g
    v = face.vertices
    tri_mid = face.midPoint
    norms = [math.hypot(v[0][0] - v[1][0], v[0][1] - v[1][1]),
             math.hypot(v[1][0] - v[2][0], v[1][1] - v[2][1]),
             math.hypot(v[2][0] - v[0][0], v[2][1] - v[0][1])]
    max_index = 1 + max(range(3), key=lambda i: (norms[i], -i))  # findmax: first maximum
    if max_index == 1:
        to_mirror, start, line = v[2], v[0], (v[1][0] - v[0][0], v[1][1] - v[0][1])
        diag_ind, mirror_ind = 1, 2
    elif max_index == 2:
        to_mirror, start, line = v[0], v[1], (v[2][0] - v[1][0], v[2][1] - v[1][1])
        diag_ind, mirror_ind = 2, 3
    else:
        to_mirror, start, line = v[1], v[2], (v[0][0] - v[2][0], v[0][1] - v[2][1])
        diag_ind, mirror_ind = 3, 4
    line_mid = (start[0] + line[0] / 2, start[1] + line[1] / 2)
    vec = (to_mirror[0] - line_mid[0], to_mirror[1] - line_mid[1])
    mirrored = (-vec[0] + line_mid[0], -vec[1] + line_mid[1])
    s = face.solidWalls
    if max_index == 1:
        new_points = [v[0], mirrored, v[1], v[2]]
        new_solid = [s[0], s[0], s[1], s[2]]
    elif max_index == 2:
        new_points = [v[0], v[1], mirrored, v[2]]
        new_solid = [s[0], s[1], s[1], s[2]]
    else:
        new_points = [v[0], v[1], v[2], mirrored]
        new_solid = [s[0], s[1], s[2], s[2]]
    tria_ids = [i for i in (1, 2, 3, 4) if i != mirror_ind]
    face2 = PolyVolume2D(new_points, new_solid, n_bins, kappa_default, sigma_default)
    face2.kappa_g = _copy_prop(face.kappa_g)      # so sub-cells inherit per-band values
    face2.sigma_s_g = _copy_prop(face.sigma_s_g)
    face2 = meshQuad(face2, Ndim, Ndim)
    pv = (tri_mid[0] - start[0], tri_mid[1] - start[1])
    t = (pv[0] * line[0] + pv[1] * line[1]) / (line[0] * line[0] + line[1] * line[1])
    t = min(1.0, max(0.0, t))
    nearest = (start[0] + t * line[0], start[1] + t * line[1])
    tm = (tri_mid[0] - nearest[0], tri_mid[1] - nearest[1])
    for sub in face2.subVolumes:
        sm = (sub.midPoint[0] - nearest[0], sub.midPoint[1] - nearest[1])
        c = tm[0] * sm[0] + tm[1] * sm[1]
        if abs(c) <= 1e-6:  # isapprox(c, 0.0, atol=1e-6): cell on the diagonal
            sub_solid_ids = [i for i in (1, 2, 3, 4) if i not in (mirror_ind - 1, mirror_ind)]
            walls_ids = sorted(sub_solid_ids + [diag_ind])
            walls_solid = [face.solidWalls[diag_ind - 1] if i == diag_ind else sub.solidWalls[i - 1]
                           for i in walls_ids]
            pts = [sub.vertices[i - 1] for i in tria_ids]
            keeper = PolyVolume2D(pts, walls_solid, n_bins, kappa_default, sigma_default)
            addSubVolume(face, keeper)
        elif c > 0.0 - 1e-6:
            # re-parented quad cell: keeps its solid flags, re-inherits properties from the triangle
            keep = PolyVolume2D(sub.vertices, sub.solidWalls, n_bins, kappa_default, sigma_default)
            addSubVolume(face, keep)
    return face


class RayRecorder:
    """RayRecorder(ids; bin=1) — parallelRayTracing.jl:194-197.  ids are 1-based global element indices."""

    def __init__(self, ids: Sequence[int], bin: int = 1, nt: int = 1):
        self.ids = [int(i) for i in ids]
        self.bin = int(bin)
        self.origins: List[np.ndarray] = [np.zeros((0, 2)) for _ in range(max(1, nt))]
        self.endpoints: List[np.ndarray] = [np.zeros((0, 2)) for _ in range(max(1, nt))]


def collect_rays(r: RayRecorder):
    """collect_rays(r) = (vcat(origins), vcat(endpoints)) — parallelRayTracing.jl:199-200."""
    return np.concatenate(r.origins, axis=0), np.concatenate(r.endpoints, axis=0)


class RayTracingDomain2D:
    """RayTracingDomain2D(faces, Ndiv; verbose) — RayTracingDomain2D.jl:114-155.

    Builds the fine mesh (IntermediateMesh2D.jl:2-23), the 1-based index maps (RayTracingDomain2D.jl:57-76,
    identical to createIndexMapping2D.jl:1-20), `uniform_across_bin` (validateDomainUniformity.jl:57-85),
    `spectral_mode` (:102-108) and `surfaces_only` (:124-131).  The object is callable like the reference's
    functor (multiDispatchRayTrace2D.jl:1-18); see tracing.py.
    """

    def __init__(self, faces: Sequence[PolyVolume2D], Ndiv: Sequence[Tuple[int, int]], verbose: bool = False):
        if len(faces) != len(Ndiv):
            raise ValueError("one (Nx, Ny) per face is required")
        self.coarse_mesh: List[PolyVolume2D] = list(faces)
        self.fine_mesh: List[List[PolyVolume2D]] = []
        verbose and print("Building intermediate mesh...")
        for face, nd in zip(faces, Ndiv):
            face.subVolumes = []
            if len(face.vertices) == 3:
                if nd[0] != nd[1]:
                    raise ValueError("Number of divisions must be equal for triangles.")  # IntermediateMesh2D.jl:7
                meshTriangle(face, int(nd[0]))
            else:
                meshQuad(face, int(nd[0]), int(nd[1]))
            self.fine_mesh.append(face.subVolumes)
        self.Ndiv = [tuple(int(x) for x in nd) for nd in Ndiv]
        verbose and print("Optimizing mesh...")
        self.surface_mapping = {}
        self.volume_mapping = {}
        self.surface_areas: List[float] = []
        self.volumes: List[float] = []
        si = vi = 1
        for c, fine in enumerate(self.fine_mesh, start=1):
            for f, cell in enumerate(fine, start=1):
                for w, solid in enumerate(cell.solidWalls, start=1):
                    if solid:
                        self.surface_mapping[(c, f, w)] = si
                        self.surface_areas.append(cell.area[w - 1])
                        si += 1
                self.volume_mapping[(c, f)] = vi
                self.volumes.append(cell.volume)
                vi += 1
        first = self.fine_mesh[0][0]
        self.is_spectral = isinstance(first.kappa_g, list)
        self.n_spectral_bins = len(first.kappa_g) if self.is_spectral else 1
        self.F_raw = None
        self.F_smooth = None
        self.energy_error = None
        self.refresh_spectral_flags(verbose=verbose)
        # surfaces_only — RayTracingDomain2D.jl:124-131 (coarse faces, mean beta over bins)
        self.surfaces_only = True
        for face in faces:
            nb = _n_bins_of(face)
            mean_beta = sum(face.beta(b) for b in range(nb)) / nb
            if face.volume * mean_beta > 1e-8:
                self.surfaces_only = False
                break
        self._device = None  # lazily created rthx handle cache (tracing.py)

    # -- spectral bookkeeping -------------------------------------------------------------------
    def refresh_spectral_flags(self, atol: float = 1e-5, verbose: bool = False) -> None:
        """validateExtinctionUniformity! (validateDomainUniformity.jl:57-85) + spectral_mode
        (RayTracingDomain2D.jl:99-108, validateSpectralUniformity! :1-55).  Re-run after editing properties."""
        uab = []
        for b in range(self.n_spectral_bins):
            first_beta = None
            broke = False
            for fine in self.fine_mesh:
                for cell in fine:
                    bt = cell.beta(b)
                    if first_beta is None:
                        first_beta = bt
                    elif abs(first_beta - bt) > atol:
                        broke = True
                        break
                if broke:
                    break
            uab.append(-1.0 if broke else first_beta)
        self.uniform_across_bin = uab
        if not self.is_spectral:
            self.spectral_mode = "grey"
        else:
            self.spectral_mode = "spectral_uniform" if self._spectrally_uniform() else "spectral_variable"

    def _spectrally_uniform(self, atol: float = 1e-10) -> bool:
        """validateSpectralUniformity! (validateDomainUniformity.jl:1-55): false as soon as a wall emissivity or a
        cell kappa/sigma_s varies across bins; otherwise true only for a black, non-scattering medium (the test
        at :49-50 uses the last wall / cell visited; the walk here is in index order)."""
        eps0 = k0 = s0 = None
        for fine in self.fine_mesh:
            for cell in fine:
                for w, solid in enumerate(cell.solidWalls):
                    if solid:
                        e = cell.epsilon[w]
                        eps0 = e[0]
                        if any(abs(x - eps0) > atol for x in e[1:]):
                            return False
        for fine in self.fine_mesh:
            for cell in fine:
                k0, s0 = cell.kappa_g[0], cell.sigma_s_g[0]
                if not all(math.isfinite(x) for x in cell.kappa_g + cell.sigma_s_g):
                    raise ValueError("Non-finite spectral properties")
                if any(abs(k - k0) > atol for k in cell.kappa_g[1:]):
                    return False
                if any(abs(s - s0) > atol for s in cell.sigma_s_g[1:]):
                    return False
        if eps0 is None or k0 is None or (k0 + s0) == 0.0:
            return False
        albedo_c = k0 / (k0 + s0)
        return abs(eps0 - albedo_c) <= 1e-10 * max(abs(eps0), abs(albedo_c)) and abs(eps0 - 1.0) < 1e-10

    # -- sizes ---------------------------------------------------------------------------------
    @property
    def num_surfaces(self) -> int:
        return len(self.surface_mapping)

    @property
    def num_volumes(self) -> int:
        return len(self.volume_mapping)

    @property
    def num_elements(self) -> int:
        return self.num_surfaces + self.num_volumes

    def __call__(self, rays_tot: int, method: str = "exchange", nudge: Optional[float] = None,
                 k_dykstra: Optional[int] = None, max_iters: int = 1000, verbose: bool = True,
                 rec: Optional[RayRecorder] = None, **kw):
        from .tracing import dispatch_ray_trace
        return dispatch_ray_trace(self, int(rays_tot), method=method, nudge=nudge, k_dykstra=k_dykstra,
                                  max_iters=max_iters, verbose=verbose, rec=rec, **kw)
