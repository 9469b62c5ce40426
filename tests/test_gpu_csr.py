"""Sparse (CSR) read-out of the device-resident counts (rthx_counts_nnz / rthx_counts_csr) and the public call on an
optically thick mesh whose F must stay sparse (test/test_2d_diffusion.jl:57-65)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def test_csr_equals_dense(rthx_mod, cuda_lib):
    rtm = rthx_mod.meshes.cfg4(Ndim=9, n_bins=3)
    flat = rthx_mod.flatten_domain(rtm)
    tr = rthx_mod.DeviceTracer(flat, device=0)
    N = tr.n_elements
    dense = tr.trace(3000, seed=51, bins=[0, 1, 2])
    for b in (2, 0, 1):
        row_ptr, cols, vals, fv = tr.counts_csr(b)
        m = sp.csr_matrix((vals, cols, row_ptr), shape=(N, N))
        assert m.has_sorted_indices or np.all(np.diff(cols[row_ptr[0]:row_ptr[1]]) > 0)
        assert np.array_equal(m.toarray(), dense["counts"][b])
        assert len(vals) == int((dense["counts"][b] != 0).sum()) and vals.min() > 0
        F = rthx_mod.counts_to_F(dense["counts"][b], 3000, verbose_loss=False)
        assert np.allclose(sp.csr_matrix((fv, cols, row_ptr), shape=(N, N)).toarray(), F.toarray(), rtol=0, atol=1e-15)
    only = tr.trace(3000, seed=51, bins=[0, 1, 2], dense=False)          # nothing dense comes back
    assert only["counts"] is None and np.array_equal(only["lost"], dense["lost"])
    assert np.array_equal(tr.counts_csr(1)[2], tr.counts_csr(1)[2])


def test_diffusion_limit_mesh_stays_sparse(rthx_mod, cuda_lib):
    """beta = 25 on a 1000:1 slab, 31x31 (test_2d_diffusion.jl): F_raw / F_smooth must be sparse matrices."""
    rtm = rthx_mod.meshes.square_domain(kappa=25.0, size=(1000.0, 1.0), Ndiv=(31, 31))
    N = rtm.num_elements
    F_smooth = rtm((4 * 31 + 961) * 1000, method="exchange", verbose=False, seed=52)
    assert sp.issparse(rtm.F_raw) and sp.issparse(F_smooth)                 # test_2d_diffusion.jl:64
    assert rtm.F_raw.nnz < 0.25 * N * N
    assert F_smooth.min() >= 0                                              # :65 no negative entries
    assert np.allclose(np.asarray(rtm.F_raw.sum(axis=1)).ravel(), 1.0)
    assert np.allclose(np.asarray(F_smooth.sum(axis=1)).ravel(), 1.0, atol=1e-9)
