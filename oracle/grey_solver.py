"""numpy restatement of the reference's grey GERT equilibrium solve — TEST INFRASTRUCTURE.

The reference pins its `:exchange` tracer only through the temperature field obtained after
`solveEquilibrium!` (test/test_2d_grey.jl:199-216).  This is the consumer needed to check traced F matrices
against those golden vectors.  It follows src/HeatTransfer/equilibrium/equilibriumGrey2D.jl:80-211:
  populateWorkspace!            WorkspaceStructs.jl:68-117   (T_in < 0  =>  flux known)
  computeEmissivePowersVariable! equilibriumGrey2D.jl:4-39
  b, coeff, M = I - diag(coeff)·F', j = M \\ h             :104-158
  receiver-indexed split r = b·g, Abs = (1-b)·g            :168-194
  computeTemperaturesVariable!  :42-77
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

STEFAN_BOLTZMANN = 5.670374419e-8

# Crosbie & Schrenker (1984) centre-line source function, test/test_2d_grey.jl:25-33
RELATIVE_TAU_Z = np.array([0.0, 0.00611, 0.02037, 0.04251, 0.07216, 0.10884, 0.15194, 0.20076, 0.25449, 0.31225,
                           0.37309, 0.43602, 0.50000, 0.56398, 0.62691, 0.68775, 0.74551, 0.79924, 0.84806,
                           0.89116, 0.92784, 0.95749, 0.97963, 0.99390, 1.00000])
SOURCE_FUNC_CENTER = np.array([0.6293, 0.6198, 0.6017, 0.5767, 0.5460, 0.5108, 0.4724, 0.4323, 0.3919, 0.3525,
                               0.3153, 0.2810, 0.2500, 0.2224, 0.1981, 0.1768, 0.1584, 0.1424, 0.1287, 0.1171,
                               0.1073, 0.0992, 0.0930, 0.0885, 0.0863])


def solve_grey(rtm, F, spectral_bin: int = 1):
    """Returns dict(T_w, T_g, j, energy_error); also writes T_g / T_w back into the fine cells."""
    F = F.toarray() if sp.issparse(F) else np.asarray(F, dtype=np.float64)
    ns, nv = rtm.num_surfaces, rtm.num_volumes
    n = ns + nv
    b_ = spectral_bin - 1
    area = np.zeros(ns); epsw = np.zeros(ns); Tw = np.zeros(ns); qw = np.zeros(ns)
    vol = np.zeros(nv); kap = np.zeros(nv); omega = np.zeros(nv); Tg = np.zeros(nv); qg = np.zeros(nv)
    for (c, f, w), s in rtm.surface_mapping.items():
        cell = rtm.fine_mesh[c - 1][f - 1]
        area[s - 1] = cell.area[w - 1]; epsw[s - 1] = cell.eps(w - 1, b_)
        Tw[s - 1] = cell.T_in_w[w - 1]; qw[s - 1] = cell.q_in_w[w - 1]
    for (c, f), v in rtm.volume_mapping.items():
        cell = rtm.fine_mesh[c - 1][f - 1]
        k = cell.kappa_g[b_] if isinstance(cell.kappa_g, list) else cell.kappa_g
        s_ = cell.sigma_s_g[b_] if isinstance(cell.sigma_s_g, list) else cell.sigma_s_g
        vol[v - 1] = cell.volume; kap[v - 1] = k
        omega[v - 1] = s_ / (k + s_) if (k + s_) > 0 else 0.0
        Tg[v - 1] = cell.T_in_g; qg[v - 1] = cell.q_in_g
    Qk = np.concatenate([Tw < 0.0, Tg < 0.0])
    E = np.concatenate([epsw * STEFAN_BOLTZMANN * area * np.abs(Tw) ** 4,
                        4 * kap * STEFAN_BOLTZMANN * vol * np.abs(Tg) ** 4])
    Q = np.concatenate([qw, qg])
    h = np.where(Qk, Q, E)
    b = np.concatenate([1.0 - epsw, omega])
    coeff = np.where(Qk, 1.0, b)
    M = np.eye(n) - coeff[:, None] * F.T
    j = np.linalg.solve(M, h)
    g = F.T @ j
    r = b * g
    Abs = (1.0 - b) * g
    e = np.maximum(j - r, 0.0)
    T = np.zeros(n)
    with np.errstate(divide="ignore", invalid="ignore"):
        T[:ns] = np.where((epsw > 0) & (area > 0), (e[:ns] / (epsw * STEFAN_BOLTZMANN * area)) ** 0.25, 0.0)
        T[ns:] = np.where((kap > 0) & (vol > 0), (e[ns:] / (4 * kap * vol * STEFAN_BOLTZMANN)) ** 0.25, 0.0)
    T = np.nan_to_num(T)
    for (c, f), v in rtm.volume_mapping.items():
        rtm.fine_mesh[c - 1][f - 1].T_g = float(T[ns + v - 1])
    for (c, f, w), s in rtm.surface_mapping.items():
        rtm.fine_mesh[c - 1][f - 1].T_w[w - 1] = float(T[s - 1])
    q = e - Abs                                    # net source per element, writeResultsToDomainGrey!: q_w = e - Abs
    return dict(T_w=T[:ns], T_g=T[ns:], j=j, q_w=q[:ns], q_g=q[ns:], area=area, energy_error=float(np.sum(j - r - Abs)))


def diffusion_S(z, beta, D, eps_w1, eps_w2, E_bw1, E_bw2):
    """Diffusion-limit source function between two plates, test/test_2d_diffusion.jl:19-23."""
    q_z = (E_bw1 - E_bw2) / (3 * beta * D / 4 + 1 / eps_w1 + 1 / eps_w2 - 1)
    E_b1 = E_bw1 + q_z * (1 / 2 - 1 / eps_w1)
    return E_b1 - (3 * beta * z / 4) * q_z


def centerline_source_function(rtm, Ndim: int, T_hot: float) -> np.ndarray:
    """extractCenterlineTemperatures + dimensionlessSourceFunction, test/test_2d_grey.jl:94-118:
    reshape(T_g, Ndim, Ndim)[(Ndim+1)÷2, :] (column-major reshape), then (T/T_hot)^4."""
    T = np.array([cell.T_g for cell in rtm.fine_mesh[0]])
    Tm = T.reshape((Ndim, Ndim), order="F")
    return (Tm[(Ndim + 1) // 2 - 1, :] / T_hot) ** 4


def analytical_centerline(Ndim: int) -> np.ndarray:
    """Linear interpolation of the C&S table at the cell centres, test/test_2d_grey.jl:186-187."""
    tau = np.linspace(1.0 / (2 * Ndim), 1.0 - 1.0 / (2 * Ndim), Ndim)
    return np.interp(tau, RELATIVE_TAU_Z, SOURCE_FUNC_CENTER)
