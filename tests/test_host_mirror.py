"""Host-side logic: mesh conventions of the reference, index maps, spectral bookkeeping, F assembly, smoothing."""
import math

import numpy as np
import pytest
import scipy.sparse as sp


def test_square_numbering_matches_readme(rthx_mod):
    rtm = rthx_mod.meshes.cfg1()
    assert (rtm.num_surfaces, rtm.num_volumes, rtm.num_elements) == (44, 121, 165)
    # readme.md:71-73: bottom_wall_indices = [1; 3:Ndim+1]; cell 1 carries surfaces 1 (bottom) and 2 (left)
    bottom = sorted(s for (c, f, w), s in rtm.surface_mapping.items() if w == 1)
    assert bottom == [1] + list(range(3, 13))
    assert rtm.surface_mapping[(1, 1, 4)] == 2
    # n fastest: fine index f = n + (m-1)*Nx, vertices (n,m),(n+1,m),(n+1,m+1),(n,m+1)
    cell = rtm.fine_mesh[0][1 + 2 * 11]
    assert np.allclose(cell.vertices, [(1 / 11, 2 / 11), (2 / 11, 2 / 11), (2 / 11, 3 / 11), (1 / 11, 3 / 11)])
    assert abs(sum(c.volume for c in rtm.fine_mesh[0]) - 1.0) < 1e-12
    assert rtm.spectral_mode == "grey" and rtm.uniform_across_bin == [1.0] and not rtm.surfaces_only


def test_config_sizes(rthx_mod):
    for cfg, (ns, nv) in (("cfg2", (164, 1681)), ("cfg4", (204, 2601)), ("cfg5", (176, 1056))):
        rtm = getattr(rthx_mod.meshes, cfg)()
        assert (rtm.num_surfaces, rtm.num_volumes) == (ns, nv)


def test_single_division_quirk(rthx_mod):
    """meshQuad.jl:145-161 `if m==1 ... elseif m==Ny`: with Ny == 1 wall 3 is never solid, with Nx == 1 wall 2."""
    rtm = rthx_mod.meshes.square_domain(Ndiv=(3, 1))
    assert [c.solidWalls for c in rtm.fine_mesh[0]] == [[True, False, False, True], [True, False, False, False],
                                                        [True, True, False, False]]
    rtm = rthx_mod.meshes.square_domain(Ndiv=(1, 2))
    assert [c.solidWalls for c in rtm.fine_mesh[0]] == [[True, False, False, True], [False, False, True, True]]


def test_triangle_submesh(rthx_mod):
    rtm = rthx_mod.meshes.circle_domain(16, 11)
    for fine in rtm.fine_mesh:
        assert len(fine) == 66 and sum(len(c.vertices) == 3 for c in fine) == 11
        assert len(fine[0].vertices) == 3                         # first cell is the diagonal triangle at the centre
        assert sum(sum(c.solidWalls) for c in fine) == 11          # only the rim is solid
    area = sum(c.volume for fine in rtm.fine_mesh for c in fine)
    assert abs(area - 16 * 0.5 * math.sin(2 * math.pi / 16)) < 1e-12
    for fine in rtm.fine_mesh:                                     # all cells CCW
        assert all(c.volume > 0 for c in fine)
    # each of the three possible longest edges
    P = rthx_mod.PolyVolume2D
    for verts in ([(0, 0), (2, 0), (0.5, 1)], [(0, 0), (1, 0.2), (-1, 2)], [(0, 0), (0.2, -1), (2, 1.5)]):
        face = P(verts, (True, True, True), 1, 1.0, 0.0)
        dom = rthx_mod.RayTracingDomain2D([face], [(4, 4)])
        assert len(dom.fine_mesh[0]) == 10
        assert abs(sum(c.volume for c in dom.fine_mesh[0]) - face.volume) < 1e-12
        assert dom.num_surfaces == 12                              # 3 edges x 4 divisions


def test_triangle_divisions_must_match(rthx_mod):
    face = rthx_mod.PolyVolume2D([(0, 0), (1, 0), (0, 1)], (True, True, True), 1, 1.0, 0.0)
    with pytest.raises(ValueError):
        rthx_mod.RayTracingDomain2D([face], [(3, 4)])


def test_group_uniform_bins(rthx_mod):
    groups, reps, non = rthx_mod.group_uniform_bins([1.0, -1.0, 2.0, 1.0 + 1e-10, 2.0, -1.0])
    assert groups == [[1, 4], [3, 5]] and reps == [1.0, 2.0] and non == [2, 6]


def test_spectral_modes(rthx_mod):
    assert rthx_mod.meshes.cfg4().spectral_mode == "spectral_variable"
    uni = rthx_mod.meshes.square_domain(5, n_bins=4)
    assert uni.spectral_mode == "spectral_uniform" and uni.uniform_across_bin == [1.0] * 4
    grey_eps = rthx_mod.meshes.square_domain(5, n_bins=4, epsilon=(1, .5, .5, .5))
    assert grey_eps.spectral_mode == "spectral_variable"          # non-black walls: full spectral solver
    assert rthx_mod.meshes.square_domain(3, kappa=0.0).surfaces_only


def test_counts_to_F_matches_reference_composition(rthx_mod, capsys):
    c = np.array([[0, 3, 1], [2, 0, 0], [0, 0, 0]], dtype=np.uint64)
    F = rthx_mod.counts_to_F(c, 5)
    assert sp.issparse(F) and F.format == "csc" and F.nnz == 3     # zero tallies are not stored
    assert np.allclose(F.toarray(), [[0, .75, .25], [1, 0, 0], [0, 0, 0]])
    assert "Maximum ray tracing ray loss per emitter: 5/5" in capsys.readouterr().out   # row 3 lost everything


def test_flatten_roundtrip(rthx_mod):
    rtm = rthx_mod.meshes.cfg5()
    flat = rthx_mod.flatten_domain(rtm)
    assert flat.n_elements == 1232 and flat.fine_off[-1] == 1056
    sid = flat.cell_surf_id[flat.cell_surf_id >= 0]
    assert sorted(sid.tolist()) == list(range(176))
    assert flat.c.n_cells == 1056 and flat.c.n_coarse == 16
    rtm4 = rthx_mod.meshes.cfg4()
    f4 = rthx_mod.flatten_domain(rtm4)
    assert f4.kappa.shape == (8, 2601) and np.allclose(f4.kappa[:, 0], rtm4.coarse_mesh[0].kappa_g)


def test_get_w_and_b(rthx_mod):
    rtm = rthx_mod.meshes.square_domain(4, kappa=0.5, sigma_s=1.5, epsilon=(1, .5, .5, .25))
    w = rthx_mod.get_w(rtm)
    assert np.allclose(w[:16], 0.25) and np.allclose(w[16:], 4 * 2.0 / 16)
    b = rthx_mod.get_b(rtm)
    assert np.allclose(b[16:, 0], 0.75) and sorted(set(np.round(b[:16, 0], 6))) == [0.0, 0.5, 0.75]


def test_smoothing_enforces_reciprocity_and_rowsums(rthx_mod):
    rng = np.random.default_rng(0)
    N = 30
    w = rng.uniform(0.5, 2.0, N)
    S = rng.random((N, N)); S = S + S.T
    F_true = S / w[:, None]
    F_true /= (w[:, None] * F_true).sum(axis=1, keepdims=True) / w[:, None]          # w_i F_ij symmetric, rows sum to 1?
    X = 0.5 * (w[:, None] * F_true + (w[:, None] * F_true).T)
    noisy = np.abs(F_true * (1 + 0.05 * rng.standard_normal((N, N))))
    noisy /= noisy.sum(axis=1, keepdims=True)
    for mat in (noisy, sp.csc_matrix(noisy)):
        Fs = rthx_mod.smoothing.AP(mat, w, N // 2, max_iters=2000)
        Fs = Fs.toarray() if sp.issparse(Fs) else Fs
        assert np.allclose(Fs.sum(axis=1), 1.0, atol=1e-12)
        WF = w[:, None] * Fs
        assert np.allclose(WF, WF.T, atol=1e-12)
        assert (Fs >= 0).all()


def test_unknown_method_raises(rthx_mod):
    rtm = rthx_mod.meshes.square_domain(3)
    with pytest.raises(ValueError, match="Unknown ray tracing method"):
        rtm(1000, method="bogus")


def test_equilibrium_host_vectors_match_the_numpy_restatement(rthx_mod):
    """populateWorkspace! / emissive powers / coeff / h of the package mirror (equilibrium.py) against the independent
    numpy restatement used by the tests (oracle/grey_solver.py), on a mesh with reflecting walls, scattering gas, a flux
    wall and a prescribed-temperature gas cell."""
    from rthx import equilibrium as eq
    rtm = rthx_mod.meshes.square_domain(5, kappa=0.7, sigma_s=0.3, epsilon=(1.0, 0.5, 0.8, 0.5))
    cells = rtm.fine_mesh[0]
    cells[7].T_in_g = 650.0                                     # one gas cell at prescribed temperature
    for cell in cells:
        for w, solid in enumerate(cell.solidWalls):
            if solid and w == 2:
                cell.T_in_w[w] = -1.0; cell.q_in_w[w] = 3.0     # flux-specified top wall
    ws = eq.populateWorkspace(rtm)
    Qk, b, coeff, h = eq._system_vectors(rtm, ws)
    ns, nv = rtm.num_surfaces, rtm.num_volumes
    assert len(h) == ns + nv and Qk[:ns].sum() == 5 and Qk[ns:].sum() == nv - 1
    assert np.allclose(b, rthx_mod.get_b(rtm)[:, 0])
    assert np.array_equal(coeff, np.where(Qk, 1.0, b))
    # surfaces in mapping order
    for (c, f, w), s in rtm.surface_mapping.items():
        cell = rtm.fine_mesh[c - 1][f - 1]
        assert ws["Area"][s - 1] == cell.area[w - 1] and ws["epsw"][s - 1] == cell.eps(w - 1)
        want = cell.q_in_w[w - 1] if cell.T_in_w[w - 1] < 0 else cell.eps(w - 1) * eq.STEFAN_BOLTZMANN * cell.area[w - 1] * cell.T_in_w[w - 1] ** 4
        assert h[s - 1] == pytest.approx(want, rel=1e-15)
    v = rtm.volume_mapping[(1, 8)]
    assert h[ns + v - 1] == pytest.approx(4 * 0.7 * eq.STEFAN_BOLTZMANN * cells[7].volume * 650.0 ** 4, rel=1e-15)
    # buildSystemMatrix.jl: M = I - Diagonal(coeff) F'
    rng = np.random.default_rng(0)
    F = rng.random((ns + nv, ns + nv)); F /= F.sum(axis=1, keepdims=True)
    M = rthx_mod.buildSystemMatrix(rtm, sp.csc_matrix(F))
    assert np.allclose(M, np.eye(ns + nv) - coeff[:, None] * F.T, atol=1e-15)


def test_solve_equilibrium_dispatch(rthx_mod):
    rtm = rthx_mod.meshes.cfg4(Ndim=3, n_bins=2)
    with pytest.raises(NotImplementedError):
        rthx_mod.solveEquilibrium(rtm, [None, None])
    rtm = rthx_mod.meshes.square_domain(3)
    rtm.spectral_mode = "nonsense"
    with pytest.raises(ValueError, match="Unknown spectral mode"):
        rthx_mod.solveEquilibrium(rtm, np.eye(rtm.num_elements))
    rtm.spectral_mode = "grey"
    with pytest.raises(ValueError, match="expected"):
        rthx_mod.solveEquilibrium(rtm, np.eye(3), verbose=False)


def test_dykstra_restatement_properties(rthx_mod):
    """numpy DkAP (smoothExchangeFactors.jl:292-318): one OP round is the exact projection onto {reciprocal, unit row
    sums}; the full DkAP result is in addition non-negative and no farther from F_raw (Frobenius norm) than AP alone."""
    from rthx import smoothing as sm
    rng = np.random.default_rng(0)
    n, ns = 60, 12
    w = np.concatenate([np.full(ns, 0.1), np.full(n - ns, 0.04)])
    w = w / w.min()
    S = rng.random((n, n)); S = S + S.T
    F = S / w[:, None]; F /= F.sum(axis=1, keepdims=True)
    Fn = np.abs(F * (1 + 0.05 * rng.standard_normal((n, n)))); Fn /= Fn.sum(axis=1, keepdims=True)
    w2 = w * w
    Y = (w2[:, None] * w2[None, :]) / (w2[:, None] + w2[None, :])
    rs = Y.sum(axis=1)
    Z = Fn / w[:, None]; Xbar = Y * (Z + Z.T)
    lam, it = sm._solve_R(Y, rs, 1 / (np.diag(Y) + rs), Xbar.sum(axis=1) - w)
    G = (Xbar - Y * (lam[:, None] + lam[None, :])) / w[:, None]
    assert it < 50 and np.abs(w[:, None] * G - (w[:, None] * G).T).max() < 1e-14 and np.abs(G.sum(axis=1) - 1).max() < 1e-13
    for k in (1, 4):
        Fd = sm.DkAP(Fn, w, ns, k_dykstra=k)
        WF = w[:, None] * Fd
        assert np.abs(WF - WF.T).max() < 1e-13 and np.abs(Fd.sum(axis=1) - 1).max() < 1e-13 and Fd.min() >= 0
    Fa = sm.AP(Fn, w, ns)
    wd = lambda A: np.linalg.norm(A - Fn)          # OP is orthogonal in the Frobenius norm of F
    assert wd(sm.DkAP(Fn, w, ns, k_dykstra=1)) <= wd(Fa)
    # the default rule of smooth_F :441-450
    assert sm.default_k_dykstra(Fn, ns) == (1 if sm.cross_coupling_chi(Fn, ns) >= 0.4 else 0)
    assert sm.default_k_dykstra(sp.csc_matrix(np.eye(n)), ns) == 0 and sm.default_k_dykstra(Fn, ns, smooth_surfaces_only=True) == 0
    strong = np.zeros((n, n)); strong[:ns, ns:] = 1.0 / (n - ns); strong[ns:, :ns] = 1.0 / ns
    assert sm.cross_coupling_chi(strong, ns) == pytest.approx(1.0) and sm.default_k_dykstra(strong, ns) == 1


def test_triangle_submesh_against_a_hand_execution_of_the_reference(rthx_mod):
    """Independent pin of the mesh restatement (both sides of every parity test consume domain.py, so a slip in the cell order would
    cancel there): meshTriangle.jl:2-103 executed BY HAND for the triangle (0,0),(1,0),(0,1), Ndim = 2.
      longest edge = edge 2 (first maximum of [1, sqrt 2, 1], :15-18)  ->  diag_ind = 2, mirror_ind = 3, mirrored point (1,1) (:27-42),
      new_points = (v1, v2, mirrored, v3) (:47), tria_ids = [1, 2, 4] (:55); meshQuad 2 x 2 visits (n, m) = (1,1), (2,1), (1,2), (2,2)
      (meshQuad.jl:139,151); cell (1,1) lies on the triangle's side -> kept as a quad (:89-93); (2,1) and (1,2) straddle the diagonal ->
      triangles of their vertices [1, 2, 4] with walls [own 1, the diagonal (parent wall 2), own 4] (:73-86); (2,2) is dropped (:94-96).
    Surfaces are numbered along the walk, wall by wall (createIndexMapping2D.jl:1-20)."""
    face = rthx_mod.PolyVolume2D([(0, 0), (1, 0), (0, 1)], (True, True, True), 1, 1.0, 0.0)
    dom = rthx_mod.RayTracingDomain2D([face], [(2, 2)])
    cells = dom.fine_mesh[0]
    expect = [
        ([(0, 0), (.5, 0), (.5, .5), (0, .5)], (True, False, False, True)),      # quad: bottom on edge 1, left on edge 3
        ([(.5, 0), (1, 0), (.5, .5)], (True, True, False)),                      # triangle: bottom on edge 1, diagonal, inner
        ([(0, .5), (.5, .5), (0, 1)], (False, True, True)),                      # triangle: inner, diagonal, left on edge 3
    ]
    assert len(cells) == 3
    for cell, (verts, solid) in zip(cells, expect):
        assert len(cell.vertices) == len(verts)
        assert all(abs(a - x) < 1e-15 and abs(b - y) < 1e-15 for (a, b), (x, y) in zip(cell.vertices, verts))
        assert tuple(bool(s) for s in cell.solidWalls) == solid
    assert dom.surface_mapping == {(1, 1, 1): 1, (1, 1, 4): 2, (1, 2, 1): 3, (1, 2, 2): 4, (1, 3, 2): 5, (1, 3, 3): 6}
    assert dom.volume_mapping == {(1, 1): 1, (1, 2): 2, (1, 3): 3}
    flat = rthx_mod.flatten_domain(dom)
    assert flat.cell_surf_id.tolist() == [[0, -1, -1, 1], [2, 3, -1, -1], [-1, 4, 5, -1]]
