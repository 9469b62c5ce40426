"""Host-side reciprocity smoothing of F (numpy/scipy restatement of the reference's alternating projection).

The north star keeps smoothing on the host, unchanged; this module exists so that the Python mirror of
`mesh(N; method=:exchange)` returns an `F_smooth` like the reference does.  It follows
src/HeatTransfer/exchangeFactorSmoothing/smoothExchangeFactors.jl:
  smooth_F   :412-459   (sparse/dense switch at density 0.25 :432-434, weight renormalisation :452-456)
  AP         :548-612   X = (WF + (WF)')/2 (`build_X` :474-490); r = X·1, u = w/r (`hunger!` :492-510);
                        X <- X ∘ (u_i+u_j)/2 (`scale!` :512-531); F = X / r (`recover_F` :537-546)
Only the AP iteration is restated; the stride/floor heuristics of the stopping rule and the Dykstra rounds
(`DkAP` :299-318) are not needed for the checks in this repository (k_dykstra is accepted and ignored with a note).
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


def smooth_F(F_raw, w: np.ndarray, num_surfaces: int, max_iters: int = 1000, smooth_surfaces_only: bool = False,
             k_dykstra=None, verbose: bool = False, renorm: bool = True):
    w = np.asarray(w, dtype=np.float64)
    verbose and print(f"Matrix size: {len(w)}x{len(w)}")
    if k_dykstra:
        verbose and print("    note: Dykstra rounds are not restated here; running AP only")
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
    return AP(F_raw, w, num_surfaces, max_iters=max_iters, verbose=verbose)
