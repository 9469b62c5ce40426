"""Host-side mirror of the reference's grey equilibrium solver with its linear algebra on the GPU.

  solveEquilibrium      HeatTransfer/equilibrium/solveEquilibrium.jl:1-25
  equilibriumGrey2D     HeatTransfer/equilibrium/equilibriumGrey2D.jl:80-211
  populateWorkspace     HeatTransfer/equilibrium/WorkspaceStructs.jl:68-117
  buildSystemMatrix     HeatTransfer/equilibrium/buildSystemMatrix.jl:1-82
  write-back            HeatTransfer/writeResults/writeResultsToDomain3D.jl:122-144

Everything that is O(N) — the workspace, the boundary-condition vectors, the temperature recovery, the write-back —
stays on the host exactly as in the reference.  The O(N^2)-per-iteration part, `j = M \\ h` with
`M = I - Diagonal(coeff) * F'` and the incident power `g = F' j`, runs on the device through `rthx_solve_grey`
(restarted GMRES(50), rtol 1e-12, as the reference's sparse branch :152-155 with Krylov.jl's atol = sqrt(eps); for a dense
F — solved exactly by `\\` at :157 — the stop is purely relative, atol = 0, so j agrees with the direct solve to ~1e-12 |h|).  When `F` is the matrix the device-side smoothing of the last `mesh(...)` call left resident,
it is read where it lies — no 900 MB host round trip for cfg3.  There is no CPU fallback: without the CUDA library or a
B200 the call raises `RthxError`.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import scipy.sparse as sp

from .flatten import flatten_domain
from ._lib import DeviceTracer

STEFAN_BOLTZMANN = 5.670374419e-8


def populateWorkspace(rtm, spectral_bin: int = 1) -> dict:
    """populateWorkspace!(ws, mesh, spectral_bin) — WorkspaceStructs.jl:68-117.  Surfaces and volumes are numbered in the
    walk order coarse -> fine -> wall, which is the order of surface_mapping / volume_mapping."""
    b = spectral_bin - 1
    ns = rtm.num_surfaces
    nv = 0 if rtm.surfaces_only else rtm.num_volumes
    ws = dict(Area=np.zeros(ns), epsw=np.zeros(ns), Tw=np.zeros(ns), qw=np.zeros(ns), Qw_known=np.zeros(ns, bool),
              Volume=np.zeros(nv), kappa_g=np.zeros(nv), omega_g=np.zeros(nv), Tg=np.zeros(nv), qg=np.zeros(nv),
              Qg_known=np.zeros(nv, bool))
    s = v = 0
    for fine in rtm.fine_mesh:
        for cell in fine:
            if not rtm.surfaces_only:
                k = cell.kappa_g[b] if isinstance(cell.kappa_g, list) else cell.kappa_g
                sg = cell.sigma_s_g[b] if isinstance(cell.sigma_s_g, list) else cell.sigma_s_g
                ws["Volume"][v] = cell.volume
                ws["kappa_g"][v] = k
                ws["omega_g"][v] = sg / (k + sg) if (k + sg) > 0.0 else 0.0
                ws["Tg"][v] = cell.T_in_g
                ws["qg"][v] = cell.q_in_g
                ws["Qg_known"][v] = cell.T_in_g < 0.0
                v += 1
            for w, solid in enumerate(cell.solidWalls):
                if solid:
                    ws["Area"][s] = cell.area[w]
                    ws["epsw"][s] = cell.eps(w, b)
                    ws["Tw"][s] = cell.T_in_w[w]
                    ws["qw"][s] = cell.q_in_w[w]
                    ws["Qw_known"][s] = cell.T_in_w[w] < 0.0
                    s += 1
    return ws


def _system_vectors(rtm, ws):
    """computeEmissivePowersVariable! + b + coeff + h — equilibriumGrey2D.jl:4-39,104-149."""
    Qk = np.concatenate([ws["Qw_known"], ws["Qg_known"]])
    E = np.concatenate([ws["epsw"] * STEFAN_BOLTZMANN * ws["Area"] * ws["Tw"] ** 4,
                        4 * ws["kappa_g"] * STEFAN_BOLTZMANN * ws["Volume"] * ws["Tg"] ** 4])
    Q = np.concatenate([ws["qw"], ws["qg"]])
    n = len(Qk)
    b = np.zeros(n)
    has_scattering = bool(np.any(ws["omega_g"] > 1e-6))
    has_reflection = float(np.sum(ws["epsw"])) < n                                  # :109 (as written in the reference)
    if has_scattering or has_reflection:
        b = np.concatenate([1.0 - ws["epsw"], ws["omega_g"]])
    h = np.where(Qk, Q, E)
    coeff = np.where(Qk, 1.0, b)
    return Qk, b, coeff, h


def buildSystemMatrix(rtm, F, spectral_bin: int = 1) -> np.ndarray:
    """buildSystemMatrix(domain, F; spectral_bin) — buildSystemMatrix.jl:1-82: M = I - Diagonal(coeff) * F' (dense, host;
    a diagnostic of the reference, not part of the solve)."""
    from .tracing import get_b
    ws = populateWorkspace(rtm, spectral_bin)
    Qk = np.concatenate([ws["Qw_known"], ws["Qg_known"]])
    n = len(Qk)
    b = get_b(rtm)[:n, spectral_bin - 1]
    coeff = np.where(Qk, 1.0, b)
    Fd = F.toarray() if sp.issparse(F) else np.asarray(F, dtype=np.float64)
    return np.eye(n) - coeff[:, None] * Fd.T


def _device_for(rtm, device: int) -> DeviceTracer:
    tr = getattr(rtm, "_device", None)
    if tr is None or getattr(tr, "_h", None) is None:
        tr = DeviceTracer(flatten_domain(rtm), device=device)      # raises RthxError without the library / a B200
        rtm._device = tr
    return tr


def equilibriumGrey2D(rtm, F, spectral_bin: int = 1, device: int = 0, verbose: bool = True) -> None:
    """equilibriumGrey2D!(mesh, F; spectral_bin) — equilibriumGrey2D.jl:80-211."""
    say = print if verbose else (lambda *a, **k: None)
    say("=== Variable Extinction Memory-Optimized Steady State Solver ===")
    ns = rtm.num_surfaces
    nv = 0 if rtm.surfaces_only else rtm.num_volumes
    n = ns + nv
    if F.shape != (n, n):
        raise ValueError(f"F is {F.shape}, expected {(n, n)}")
    ws = populateWorkspace(rtm, spectral_bin)
    Qk, b, coeff, h = _system_vectors(rtm, ws)
    say("Solving linear system...")
    tr = _device_for(rtm, device)
    resident = getattr(tr, "_resident_F", None)
    # A dense F is solved exactly by the reference (`M \\ h`, :157): no absolute tolerance, the residual target is purely
    # relative (1e-12 |h|).  Only the sparse branch inherits Krylov.jl's atol = sqrt(eps) (gmres!, :152-155).
    if resident is not None and resident[0] is F and _same_sample(F, resident[1]):
        j, g, st = tr.solve_grey(coeff, h, atol=0.0)               # F_smooth is still on the device
    else:
        j, g, st = tr.solve_grey(coeff, h, F=F, atol=-1.0 if sp.issparse(F) else 0.0)
    rtm.last_solve_stats = st
    if not st["converged"]:
        print(f"Warning: GMRES stopped at residual {st['residual']:.3e} after {st['iterations']} iterations")
    say(f"GMRES({50}): {st['iterations']} iterations, residual {st['residual']:.3e}, {st['total_ms']:.2f} ms on the device")
    # receiver-indexed split (:168-194)
    r = b * g
    Abs = (1.0 - b) * g
    # computeTemperaturesVariable! (:42-77)
    e = np.maximum(j - r, 0.0)
    T = np.zeros(n)
    with np.errstate(divide="ignore", invalid="ignore"):
        T[:ns] = np.where((ws["epsw"] > 0.0) & (ws["Area"] > 0.0), (e[:ns] / (ws["epsw"] * STEFAN_BOLTZMANN * ws["Area"])) ** 0.25, 0.0)
        if nv:
            T[ns:] = np.where((ws["kappa_g"] > 0.0) & (ws["Volume"] > 0.0),
                              (e[ns:] / (4 * ws["kappa_g"] * ws["Volume"] * STEFAN_BOLTZMANN)) ** 0.25, 0.0)
    T = np.nan_to_num(T, nan=0.0)
    # writeResultsToDomainGrey! (writeResultsToDomain3D.jl:122-144)
    for (c, f, w), s in rtm.surface_mapping.items():
        cell = rtm.fine_mesh[c - 1][f - 1]
        i = s - 1
        cell.T_w[w - 1] = float(T[i]); cell.j_w[w - 1] = float(j[i]); cell.g_a_w[w - 1] = float(Abs[i])
        cell.e_w[w - 1] = float(e[i]); cell.r_w[w - 1] = float(r[i]); cell.g_w[w - 1] = float(Abs[i] + r[i])
        cell.q_w[w - 1] = float(e[i] - Abs[i]); cell.i_w[w - 1] = float(j[i] / (math.pi * cell.area[w - 1]))
    if not rtm.surfaces_only:
        for (c, f), v in rtm.volume_mapping.items():
            cell = rtm.fine_mesh[c - 1][f - 1]
            i = ns + v - 1
            cell.T_g = float(T[i]); cell.j_g = float(j[i]); cell.g_a_g = float(Abs[i]); cell.e_g = float(e[i])
            cell.r_g = float(r[i]); cell.g_g = float(Abs[i] + r[i]); cell.q_g = float(e[i] - Abs[i])
            cell.i_g = float(j[i] / (4 * math.pi * cell.volume))
    rtm.energy_error = float(np.sum(j - r - Abs))                   # :207
    say("=== Variable Extinction Steady State Solution Complete ===")
    return None


def _sample_of(F: np.ndarray) -> np.ndarray:
    """A strided sample of a dense matrix: detects in-place edits of the array whose device copy would be reused."""
    flat = F.reshape(-1)
    return flat[:: max(1, flat.size // 4096)].copy()


def _same_sample(F, sample: np.ndarray) -> bool:
    return isinstance(F, np.ndarray) and np.array_equal(_sample_of(F), sample)


def solveEquilibrium(rtm, F_matrices, max_iters: int = 1000, convergence_tol: float = 1e-12, device: int = 0,
                     verbose: bool = True) -> None:
    """solveEquilibrium!(domain, F_matrices; max_iters, convergence_tol) — solveEquilibrium.jl:1-25.  Grey domains only:
    the spectral solvers (equilibriumSpectral2D.jl, 1218 lines of Newton / Woodbury iterations) are outside the hot path
    this repository replaces and stay with the reference."""
    if rtm.spectral_mode == "grey":
        return equilibriumGrey2D(rtm, F_matrices, device=device, verbose=verbose)
    if rtm.spectral_mode in ("spectral_uniform", "spectral_variable"):
        raise NotImplementedError("spectral equilibrium solvers are outside the scope of this drop-in (grey solve only)")
    raise ValueError(f"Unknown spectral mode: {rtm.spectral_mode}")
