"""Host-side reciprocity smoothing of F (numpy/scipy restatement of the reference's alternating projection).

The north star keeps smoothing on the host, unchanged; this module exists so that the Python mirror of
`mesh(N; method=:exchange)` returns an `F_smooth` like the reference does.  It follows
src/HeatTransfer/exchangeFactorSmoothing/smoothExchangeFactors.jl:
  smooth_F   :412-459   (sparse/dense switch at density 0.25 :432-434, weight renormalisation :452-456)
  AP         :548-612   X = (WF + (WF)')/2 (`build_X` :474-490); r = X·1, u = w/r (`hunger!` :492-510);
                        X <- X ∘ (u_i+u_j)/2 (`scale!` :512-531); F = X / r (`recover_F` :537-546)
  cross_coupling_chi :215-241, Y_mat :243-259, Xbar_b :261-281, solve_R :15-37, OP :292-297, DkAP :299-318
                        Dykstra rounds in front of AP; the reference's default for a dense F is one round when the
                        surface-gas cross-coupling chi >= 0.4, else none (:441-450)
The stride/floor heuristics of AP's stopping rule are not restated (the iteration runs to the 8 eps target).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _delta_R(X, w, r):
    """Row-sum defect of the iterate recovered from X: max_i |r_i / w_i - 1|."""
    return float(np.max(np.abs(r / w - 1.0)))


def AP(F, w: np.ndarray, num_surfaces: int, max_iters: int = 1000, verbose: bool = False):
    N = len(w)
    if num_surfaces == N and N > 0:            # AP_convergence_check :461-472
        if not np.max(w) < 0.5 * np.sum(w):
            raise ValueError("Smoothing convergence check failed: max surface w >= half of total w")
    sparse = sp.issparse(F)
    target = 8 * np.finfo(np.float64).eps
    if sparse:
        WF = sp.diags(w) @ F.tocsr()
        X = (0.5 * (WF + WF.T)).tocsr()
        rows = np.repeat(np.arange(N), np.diff(X.indptr))
        cols = X.indices
    else:
        WF = w[:, None] * np.asarray(F, dtype=np.float64)
        X = 0.5 * (WF + WF.T)
    r = np.asarray(X.sum(axis=1)).ravel()
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.where(r > 0, w / r, 1.0)
    delta = _delta_R(X, w, np.where(r > 0, r, w))
    k = 0
    while k < max_iters and delta > target:
        if sparse:
            X.data *= 0.5 * (u[rows] + u[cols])
        else:
            X *= 0.5 * (u[:, None] + u[None, :])
        k += 1
        r = np.asarray(X.sum(axis=1)).ravel()
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.where(r > 0, w / r, 1.0)
        delta = _delta_R(X, w, np.where(r > 0, r, w))
    verbose and print(f"AP: {k} iterations, final delta_R = {delta:.3e}")
    rr = np.where(r > 0, r, 1.0)
    if sparse:
        X.data /= rr[rows]
        return X.tocsc()
    return X / rr[:, None]


def cross_coupling_chi(F, n_surf: int) -> float:
    """cross_coupling_chi :215-241 — (sum F[s,g] + sum F[g,s]) / N: how much of F couples surfaces with gas cells."""
    N = F.shape[0]
    if sp.issparse(F):
        Fc = F.tocsr()
        return float(Fc[:n_surf, n_surf:].sum() + Fc[n_surf:, :n_surf].sum()) / N
    F = np.asarray(F)
    return float(F[:n_surf, n_surf:].sum() + F[n_surf:, :n_surf].sum()) / N


def default_k_dykstra(F_raw, num_surfaces: int, smooth_surfaces_only: bool = False) -> int:
    """The k_dykstra === nothing branch of smooth_F :441-450 for the matrix as it will be smoothed (dense iff density
    > 0.25, :432-434): AP only for chi < 0.4 or a sparse matrix, else one Dykstra round."""
    if smooth_surfaces_only:
        return 0                                                   # chi := 0 for a convex enclosure (:425-426)
    n = F_raw.shape[0]
    dense = (not sp.issparse(F_raw)) or F_raw.nnz / float(n * n) > 0.25
    return 1 if (dense and cross_coupling_chi(F_raw, num_surfaces) >= 0.4) else 0


def _solve_R(Y, rowsum, dinv, b, rtol: float = 1e-14, maxiter: int = 200):
    """solve_R :15-37 — Jacobi-preconditioned CG on R = Y + Diagonal(Y 1)."""
    x = np.zeros_like(b)
    bn = np.linalg.norm(b)
    if bn == 0.0:
        return x, 0
    r = b.copy(); z = dinv * r; p = z.copy(); rz = r @ z
    for it in range(1, maxiter + 1):
        Ap = Y @ p + rowsum * p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxiter


def DkAP(F_raw, w: np.ndarray, num_surfaces: int, k_dykstra: int = 0, max_iters: int = 1000, verbose: bool = False):
    """DkAP :299-318 (dense): k_dykstra rounds of OP (:292-297) + clipping with the Dykstra correction, row
    renormalisation, then AP."""
    if k_dykstra <= 0:
        return AP(F_raw, w, num_surfaces, max_iters=max_iters, verbose=verbose)
    F = F_raw.toarray() if sp.issparse(F_raw) else np.array(F_raw, dtype=np.float64)
    w2 = w * w
    Y = (w2[:, None] * w2[None, :]) / (w2[:, None] + w2[None, :])            # Y_mat :243-259
    rowsum = Y.sum(axis=1)
    dinv = 1.0 / (np.diag(Y) + rowsum)
    P = np.zeros_like(F)
    for k in range(1, k_dykstra + 1):
        Z = F / w[:, None]
        Xbar = Y * (Z + Z.T)                                                   # Xbar_b :261-281
        lam, iters = _solve_R(Y, rowsum, dinv, Xbar.sum(axis=1) - w)
        G = (Xbar - Y * (lam[:, None] + lam[None, :])) / w[:, None]          # OP :292-297
        F_new = np.maximum(G + P, 0.0)
        delta = np.inf
        if k % 5 == 0 or k == k_dykstra:                                       # delta_perp(:DYK) :136-142
            bb = w * (F_new.sum(axis=1) - 1.0)
            l2, _ = _solve_R(Y, rowsum, dinv, bb)
            delta = float(np.sqrt(max(bb @ l2, 0.0)))
            verbose and print(f"Dykstra round {k} ({iters} PCG iterations): delta_perp = {delta}")
        P = G + P - F_new
        F = F_new
        if delta < 8 * np.finfo(np.float64).eps:
            break
    F = F / F.sum(axis=1, keepdims=True)
    return AP(F, w, num_surfaces, max_iters=max_iters, verbose=verbose)


def smooth_F(F_raw, w: np.ndarray, num_surfaces: int, max_iters: int = 1000, smooth_surfaces_only: bool = False,
             k_dykstra=None, verbose: bool = False, renorm: bool = True):
    w = np.asarray(w, dtype=np.float64)
    verbose and print(f"Matrix size: {len(w)}x{len(w)}")
    if k_dykstra is None:
        k_dykstra = default_k_dykstra(F_raw, num_surfaces, smooth_surfaces_only)
    if not smooth_surfaces_only and sp.issparse(F_raw):
        nz = F_raw.nnz
        if nz / len(w) ** 2 > 0.25:
            F_raw = F_raw.toarray()
    if smooth_surfaces_only and not sp.issparse(F_raw):
        w = w[:num_surfaces]
        F_raw = F_raw[:num_surfaces, :num_surfaces]
    elif smooth_surfaces_only:
        w = w[: F_raw.shape[0]]
    if renorm:
        w = w / np.min(w)
    return DkAP(F_raw, w, num_surfaces, k_dykstra=int(k_dykstra), max_iters=max_iters, verbose=verbose)
