"""GPU reciprocity smoothing (rthx_smooth_F) against the numpy restatement of the reference's alternating projection."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _checks(Fs, w):
    assert np.allclose(Fs.sum(axis=1), 1.0, atol=1e-12)
    WF = w[:, None] * Fs
    assert np.abs(WF - WF.T).max() <= 1e-12 * np.abs(WF).max() + 1e-15
    assert (Fs >= 0).all()


def test_smoothing_matches_numpy_ap(rthx_mod, oracle_mod, cuda_lib):
    rtm = rthx_mod.meshes.square_domain(9, kappa=0.7, sigma_s=0.3)
    flat = rthx_mod.flatten_domain(rtm)
    tr = rthx_mod.DeviceTracer(flat, device=0)
    rpe = 20000
    out = tr.trace(rpe, seed=41)
    w = rthx_mod.get_w(rtm)
    wn = w / w.min()
    F_raw = out["counts"][0] / out["counts"][0].sum(axis=1, keepdims=True)
    ref = rthx_mod.smoothing.AP(F_raw, wn, flat.n_surfaces, max_iters=1000)
    # (a) straight from the counts still resident on the device, (b) from host counts, (c) from a host F matrix
    Fa, sa = tr.smooth(wn)
    Fb, sb = tr.smooth(wn, counts=out["counts"][0])
    Fc, sc = tr.smooth(wn, F=F_raw, measure_pass=True)
    for Fs in (Fa, Fb, Fc):
        _checks(Fs, wn)
        assert np.abs(Fs - ref).max() < 1e-10
    assert np.array_equal(Fa, Fb)
    assert sa["iterations"] > 0 and sa["delta"] <= max(8 * np.finfo(float).eps, 1e-3 * sa["delta_init"])
    assert sc["pass_gbs"] > 0


def test_surfaces_only_crop(rthx_mod, cuda_lib):
    """exchangeRayTracing.jl:9-11: transparent medium -> smooth the leading Ns x Ns block only."""
    rtm = rthx_mod.meshes.square_domain(8, kappa=0.0)
    flat = rthx_mod.flatten_domain(rtm)
    tr = rthx_mod.DeviceTracer(flat, device=0)
    out = tr.trace(40000, seed=42)
    ns = flat.n_surfaces
    w = rthx_mod.get_w(rtm)[:ns]
    Fs, st = tr.smooth(w / w.min(), n=ns)
    _checks(Fs, w / w.min())
    F_raw = out["counts"][0][:ns, :ns] / out["counts"][0][:ns, :ns].sum(axis=1, keepdims=True)
    assert np.abs(Fs - F_raw).max() < 0.01          # smoothing is a small correction of the MC estimate


def test_public_call_uses_gpu_smoothing(rthx_mod, cuda_lib):
    from oracle import grey_solver as gs
    rtm = rthx_mod.meshes.cfg1()
    rtm(1_000_000, method="exchange", verbose=False, seed=43)
    assert getattr(rtm, "last_smooth_stats", None) is not None and rtm.last_smooth_stats["iterations"] > 0
    Fs = rtm.F_smooth if isinstance(rtm.F_smooth, np.ndarray) else rtm.F_smooth.toarray()
    w = rthx_mod.get_w(rtm)
    _checks(Fs, w / w.min())
    gs.solve_grey(rtm, rtm.F_smooth)
    S = gs.centerline_source_function(rtm, 11, 1000.0)
    A = gs.analytical_centerline(11)
    assert np.linalg.norm(S - A) <= 0.05 * max(np.linalg.norm(S), np.linalg.norm(A))


@pytest.mark.parametrize("k_dykstra", [1, 3, 6])
def test_dykstra_rounds_match_numpy_dkap(rthx_mod, cuda_lib, k_dykstra):
    """rthx_smooth_DkAP (OP through the dual PCG solve, clipping with the Dykstra correction, then AP) against the numpy
    restatement of DkAP (smoothExchangeFactors.jl:299-318), from resident counts, host counts and a host F matrix."""
    rtm = rthx_mod.meshes.square_domain(9, kappa=0.7, sigma_s=0.3)
    flat = rthx_mod.flatten_domain(rtm)
    tr = rthx_mod.DeviceTracer(flat, device=0)
    out = tr.trace(3000, seed=41)                      # few rays: a noisy F, so the projection has real work to do
    w = rthx_mod.get_w(rtm)
    wn = w / w.min()
    F_raw = out["counts"][0] / out["counts"][0].sum(axis=1, keepdims=True)
    ref = rthx_mod.smoothing.DkAP(F_raw, wn, flat.n_surfaces, k_dykstra=k_dykstra, max_iters=1000)
    Fa, sa = tr.smooth(wn, k_dykstra=k_dykstra)
    Fb, sb = tr.smooth(wn, counts=out["counts"][0], k_dykstra=k_dykstra)
    Fc, sc = tr.smooth(wn, F=F_raw, k_dykstra=k_dykstra)
    for Fs in (Fa, Fb, Fc):
        _checks(Fs, wn)
        assert np.abs(Fs - ref).max() < 1e-10
    assert np.array_equal(Fa, Fb)
    assert sa["dykstra_rounds"] == k_dykstra and sa["pcg_iterations"] >= k_dykstra and np.isfinite(sa["dykstra_delta"])
    # the orthogonal projection stays closer to the traced matrix than plain AP does
    F_ap, s_ap = tr.smooth(wn, k_dykstra=0)
    assert s_ap["dykstra_rounds"] == 0
    wd = lambda F: np.linalg.norm(F - F_raw)       # OP is orthogonal in the Frobenius norm of F
    assert wd(Fa) <= wd(F_ap) * (1 + 1e-9)
    tr.close()


def test_public_call_default_follows_reference_chi_rule(rthx_mod, cuda_lib):
    """smooth_F :441-450: dense F with chi >= 0.4 -> one Dykstra round by default; k_dykstra=0 forces AP only."""
    rtm = rthx_mod.meshes.cfg1()
    rtm(1_000_000, method="exchange", verbose=False, seed=43)
    chi = rthx_mod.smoothing.cross_coupling_chi(rtm.F_raw, rtm.num_surfaces)
    assert chi >= 0.4 and rtm.last_smooth_stats["dykstra_rounds"] == 1
    F1 = np.array(rtm.F_smooth)
    rtm(1_000_000, method="exchange", verbose=False, seed=43, k_dykstra=0)
    assert rtm.last_smooth_stats["dykstra_rounds"] == 0
    assert 0 < np.abs(F1 - rtm.F_smooth).max() < 5e-2      # both are small corrections of the same traced matrix
    w = rthx_mod.get_w(rtm)
    _checks(F1, w / w.min())
