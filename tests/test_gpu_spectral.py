"""`:spectral_variable` end to end (exchangeRayTracing.jl:14-48): every traced band smoothed on the device from the resident counts
with that band's weights, cfg4 at its NAMED size (51 x 51, 8 bands in one launch) against the oracle ray for ray, and the 10-band
1 % kappa ramp of test/test_2d_spectral_dense_sparse.jl:62-85 checked at the F level (per-band grey solve vs Crosbie & Schrenker)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import n_differing_rays

pytestmark = pytest.mark.gpu


def test_cfg4_named_size_exact_parity(rthx_mod, oracle_mod, cuda_lib):
    """BASELINE.json configs[3]: 51 x 51, 8 bands batched in the grid — the full named mesh, 1500 rays per emitter and band."""
    rtm = rthx_mod.meshes.cfg4()
    assert rtm.spectral_mode == "spectral_variable" and rtm.num_elements == 2805
    flat = rthx_mod.flatten_domain(rtm)
    tr = rthx_mod.DeviceTracer(flat, device=0)
    bins, rpe = list(range(8)), 1500
    ref = oracle_mod.trace(flat, rpe, seed=44, bins=bins)
    got = tr.trace(rpe, seed=44, bins=bins)
    total = int(ref["counts"].sum())
    assert np.all(got["counts"].sum(axis=2) + got["lost"] == rpe)
    nd = n_differing_rays(got["counts"], ref["counts"])
    assert nd <= max(2, int(2e-6 * total)), f"{nd} of {total} ray outcomes differ from the oracle"
    assert np.array_equal(got["lost"], ref["lost"])


def test_per_band_device_smoothing_matches_host_restatement(rthx_mod, cuda_lib):
    """Each traced band of a :spectral_variable domain is smoothed on the device with ITS weights (get_w(rtm, bin)); result equals the
    numpy restatement of smooth_F on that band's F_raw; grouped bins alias one matrix (parallelRayTracing.jl:38-41)."""
    from rthx import smoothing
    kb = [1.0, 1.0, 1.6, 2.4]                                     # bins 1 and 2 share beta -> one group, one trace
    rtm = rthx_mod.meshes.square_domain(9, kappa=1.0, n_bins=4, kappa_bins=kb)
    assert rtm.spectral_mode == "spectral_variable"
    Fs = rtm((36 + 81) * 6000, method="exchange", verbose=False, seed=91)
    assert len(Fs) == 4 and Fs[0] is Fs[1] and rtm.F_raw[0] is rtm.F_raw[1] and Fs[2] is not Fs[3]
    ns = rtm.num_surfaces
    for b in (1, 3, 4):
        F = Fs[b - 1]
        assert isinstance(F, np.ndarray)                           # dense branch -> device
        w = rthx_mod.get_w(rtm, spectral_bin=b)
        host = smoothing.smooth_F(rtm.F_raw[b - 1], w, ns, max_iters=1000, verbose=False)
        assert np.abs(F - np.asarray(host)).max() < 1e-10
        wn = w / w.min()
        assert np.allclose(F.sum(axis=1), 1.0, atol=1e-12)
        assert np.abs(wn[:, None] * F - (wn[:, None] * F).T).max() < 1e-12      # reciprocity w_i F_ij = w_j F_ji
    assert not np.allclose(Fs[2], Fs[3])


def test_ten_band_ramp_reproduces_crosbie_schrenker_per_band(rthx_mod, cuda_lib):
    """test/test_2d_spectral_dense_sparse.jl:62-85 at the F level: 11 x 11, 10 bins with a 1 % kappa ramp -> :spectral_variable,
    10 separate traces in ONE launch; every band's F_smooth through the grey solve lands on the C&S centre-line (RMS < 0.05, :83),
    no negative entries (:70), and smoothing does not hurt (:82, on the band average)."""
    from oracle import grey_solver as gs
    n_side, n_bins = 11, 10
    kb = [1.0 * (1.0 + 0.01 * k / (n_bins - 1)) for k in range(n_bins)]
    rtm = rthx_mod.meshes.square_domain(n_side, kappa=1.0, n_bins=n_bins, kappa_bins=kb)
    assert rtm.spectral_mode == "spectral_variable"
    Fs = rtm((4 * n_side + n_side ** 2) * 1000, method="exchange", verbose=False, seed=17)
    assert len(Fs) == n_bins and len({id(F) for F in Fs}) == n_bins
    A = gs.analytical_centerline(n_side)
    err_raw, err_smooth = [], []
    for b in range(1, n_bins + 1):
        Fb = Fs[b - 1]
        assert (Fb.min() if isinstance(Fb, np.ndarray) else Fb.data.min()) >= 0.0
        for F, acc in ((rtm.F_raw[b - 1], err_raw), (Fb, err_smooth)):
            rthx_mod.equilibriumGrey2D(rtm, F, spectral_bin=b, verbose=False)
            S = gs.centerline_source_function(rtm, n_side, 1000.0)
            acc.append(float(np.sqrt(np.mean((S - A) ** 2))))
        assert abs(rtm.energy_error) < 1e-4
    assert max(err_smooth) < 0.05
    assert np.mean(err_smooth) <= np.mean(err_raw) * 1.05
