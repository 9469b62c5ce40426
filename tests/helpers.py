"""Shared helpers for the parity tests (statistics of SURVEY.md App. D)."""
import numpy as np


def n_differing_rays(a, b):
    """Number of rays whose outcome differs between two tallies of the same ray set (half the L1 distance)."""
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum() // 2)


def z_statistics(cg, co, min_expected=10.0):
    """Two-sample binomial z per entry between count matrices cg, co ([N,N] each, rows = emitters).
    Returns (max|z|, fraction |z|>3, number of tested entries)."""
    cg = cg.astype(np.float64); co = co.astype(np.float64)
    ng = cg.sum(axis=1, keepdims=True); no = co.sum(axis=1, keepdims=True)
    p = (cg + co) / (ng + no)
    var = p * (1 - p) * (1.0 / ng + 1.0 / no)
    mask = ((cg + co) >= 2 * min_expected) & (var > 0)
    z = np.zeros_like(cg)
    z[mask] = (cg / ng - co / no)[mask] / np.sqrt(var[mask])
    M = int(mask.sum())
    return float(np.abs(z[mask]).max()) if M else 0.0, float((np.abs(z[mask]) > 3).mean()) if M else 0.0, M
