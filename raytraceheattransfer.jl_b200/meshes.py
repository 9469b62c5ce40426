"""Synthetic domains of the shapes named in BASELINE.json `configs`, built with the same
constructors a user of the reference would call."""
from __future__ import annotations

import math
from typing import Sequence

from .domain import PolyVolume2D, RayTracingDomain2D


def square_domain(Ndim: int = 11, kappa: float = 1.0, sigma_s: float = 0.0, epsilon: Sequence[float] = (1, 1, 1, 1),
                  T_walls: Sequence[float] = (1000.0, 0.0, 0.0, 0.0), n_bins: int = 1, kappa_bins=None,
                  rotation_angle: float = 0.0, size: Sequence[float] = (1.0, 1.0), Ndiv=None,
                  solid=(True, True, True, True)) -> RayTracingDomain2D:
    """1x1 (or size[0] x size[1]) enclosure like createSquareDomain2D in test/test_2d_grey.jl:43-95."""
    w, h = size
    verts = [(0.0, 0.0), (w, 0.0), (w, h), (0.0, h)]
    if rotation_angle != 0.0:
        cx, cy = w / 2, h / 2
        ct, st = math.cos(rotation_angle), math.sin(rotation_angle)
        verts = [((x - cx) * ct - (y - cy) * st + cx, (x - cx) * st + (y - cy) * ct + cy) for x, y in verts]
    face = PolyVolume2D(verts, solid, n_bins, kappa, sigma_s)
    if n_bins > 1:
        if kappa_bins is not None:
            face.kappa_g = [float(k) for k in kappa_bins]
        face.epsilon = [[float(e)] * n_bins for e in epsilon]
    else:
        face.epsilon = [float(e) for e in epsilon]
    face.T_in_w = [float(t) for t in T_walls]
    face.T_in_g = -1.0
    face.q_in_g = 0.0
    return RayTracingDomain2D([face], [Ndiv or (Ndim, Ndim)])


def cfg1():
    """README Example 1: 1x1 m, 11x11, kappa=1, black walls, bottom 1000 K."""
    return square_domain(11, kappa=1.0)


def cfg2():
    """test_2d_grey_reflecting: 41x41, eps=(1,.5,.5,.5)."""
    return square_domain(41, kappa=1.0, epsilon=(1, .5, .5, .5))


def cfg3(Ndim: int = 101):
    """Grey absorbing+scattering kappa=sigma_s=0.5, 101x101."""
    return square_domain(Ndim, kappa=0.5, sigma_s=0.5)


def cfg4(Ndim: int = 51, n_bins: int = 8):
    """Spectral multi-band gas: kappa_b = 1 + 0.01 (b-1)/(n_bins-1) => :spectral_variable, every bin uniform."""
    kb = [1.0 * (1 + 0.01 * b / (n_bins - 1)) for b in range(n_bins)]
    return square_domain(Ndim, kappa=1.0, n_bins=n_bins, kappa_bins=kb)


def circle_domain(N_seg: int = 16, Ndim: int = 11, kappa: float = 1.0, R: float = 1.0, T_hot: float = 1000.0,
                  half_hot: bool = True) -> RayTracingDomain2D:
    """Circle of N_seg triangular wedges, rim solid, spokes open (test/test_triangle_mesh.jl:19-63)."""
    def rim(j):
        return (R * math.cos(2 * math.pi * (j - 1) / N_seg), R * math.sin(2 * math.pi * (j - 1) / N_seg))
    faces, div = [], []
    for j in range(1, N_seg + 1):
        face = PolyVolume2D([(0.0, 0.0), rim(j), rim(j + 1)], (False, True, False), 1, kappa, 0.0)
        hot = T_hot if (not half_hot or j <= N_seg // 2) else 0.0
        face.T_in_w = [0.0, hot, 0.0]
        face.epsilon = [1.0, 1.0, 1.0]
        face.T_in_g = -1.0
        face.q_in_g = 0.0
        faces.append(face)
        div.append((Ndim, Ndim))
    return RayTracingDomain2D(faces, div)


def cfg5():
    """Multi-volume geometry with transparent interfaces and triangle sub-meshes (16 wedges x (11,11))."""
    return circle_domain(16, 11)


def two_quads_domain(Ndiv=((4, 3), (5, 3)), kappa=(1.0, 2.0), skew: float = 0.0) -> RayTracingDomain2D:
    """Two quadrilaterals sharing a transparent interface with different kappa (variable-beta path);
    skew != 0 makes the right one a non-parallelogram (generic locator path)."""
    f1 = PolyVolume2D([(0, 0), (1, 0), (1, 1), (0, 1)], (True, False, True, True), 1, kappa[0], 0.0)
    f2 = PolyVolume2D([(1, 0), (2.0, 0.0 - skew), (2.0 + skew, 1.0), (1, 1)], (True, True, True, False), 1, kappa[1], 0.0)
    for f in (f1, f2):
        f.epsilon = [1.0] * 4
        f.T_in_g = -1.0
    f1.T_in_w = [1000.0, 0, 0, 0]
    return RayTracingDomain2D([f1, f2], list(Ndiv))


def parallel_plates_domain(W: float = 100.0, H: float = 1.0, Ndiv=(21, 2), eps_plates: float = 0.5, T_hot: float = 1000.0,
                           kappa: float = 1e-3) -> RayTracingDomain2D:
    """Wide thin enclosure of test/test_2d_grey_reflecting.jl:96-121 ('Parallel Plates vs Textbook'): grey-diffuse plates of
    emissivity eps_plates at the bottom (hot) and the top (cold), black cold side walls, an almost transparent medium."""
    face = PolyVolume2D([(0.0, 0.0), (W, 0.0), (W, H), (0.0, H)], (True, True, True, True), 1, kappa, 0.0)
    face.T_in_w = [T_hot, 0.0, 0.0, 0.0]
    face.q_in_w = [0.0, 0.0, 0.0, 0.0]
    face.epsilon = [eps_plates, 1.0, eps_plates, 1.0]
    face.T_in_g = -1.0
    face.q_in_g = 0.0
    return RayTracingDomain2D([face], [tuple(Ndiv)])


def diffusion_slab_domain(N_side: int = 31, beta: float = 25.0, aspect: float = 1000.0, T_hot: float = 1000.0) -> RayTracingDomain2D:
    """Optically thick 1000:1 slab of test/test_2d_diffusion.jl:28-39 (build_diffusion_grey): black walls, bottom hot."""
    face = PolyVolume2D([(0.0, 0.0), (aspect, 0.0), (aspect, 1.0), (0.0, 1.0)], (True, True, True, True), 1, beta, 0.0)
    face.T_in_w = [T_hot, 0.0, 0.0, 0.0]
    face.epsilon = [1.0, 1.0, 1.0, 1.0]
    face.T_in_g = -1.0
    face.q_in_g = 0.0
    return RayTracingDomain2D([face], [(N_side, N_side)])
