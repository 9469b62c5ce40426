"""Known-answer tests of the oracle's primitives (the reference has no unit tests at this level)."""
import math

import numpy as np
import pytest


def test_philox_known_answers(oracle_mod):
    # Random123 kat_vectors for philox4x32-10
    assert oracle_mod.philox((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert oracle_mod.philox((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert oracle_mod.philox((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_uniform_conversions(oracle_mod):
    L = oracle_mod.lib()
    assert L.rthx_oracle_u52(0, 0) == 2.0 ** -53
    assert L.rthx_oracle_u52(0xFFFFFFFF, 0xFFFFFFFF) == 1.0 - 2.0 ** -53
    assert L.rthx_oracle_u52(0xFFF, 0) == 2.0 ** -53            # low 12 bits are discarded
    assert L.rthx_oracle_u52(0x1000, 0) == 1.5 * 2.0 ** -52
    assert L.rthx_oracle_u23(0) == 2.0 ** -24
    assert L.rthx_oracle_u23(0xFFFFFFFF) == 1.0 - 2.0 ** -24


def test_dist_to_surface_unit_square(oracle_mod):
    vx, vy = [0, 1, 1, 0], [0, 0, 1, 1]
    u, k = oracle_mod.dist_to_surface(vx, vy, (0.5, 0.5), (1.0, 0.0))
    assert (u, k) == (0.5, 1)                                   # right wall (index 1, 0-based)
    u, k = oracle_mod.dist_to_surface(vx, vy, (0.25, 0.5), (0.0, -0.5))
    assert (u, k) == (1.0, 0)                                   # |d| = 0.5: u is a ray parameter, not a length
    u, k = oracle_mod.dist_to_surface(vx, vy, (0.5, 0.5), (0.5, 0.5))
    assert u == 1.0 and k == 1                                  # corner tie -> first index (right before top)
    u, k = oracle_mod.dist_to_surface(vx, vy, (0.5, 0.5), (1e-11, 0.0))
    assert math.isinf(u) and k == 0                             # |d.n| < 1e-10 everywhere -> (Inf, first)
    u, k = oracle_mod.dist_to_surface([0, 1, 0], [0, 0, 1], (0.25, 0.25), (1.0, 1.0))
    assert k == 1 and abs(u - 0.25) < 1e-15                     # hypotenuse of a triangle


def test_find_face_matches_lattice(oracle_mod, rthx_mod):
    rtm = rthx_mod.meshes.square_domain(7, Ndiv=(7, 5))
    flat = rthx_mod.flatten_domain(rtm)
    rng = np.random.default_rng(0)
    for x, y in rng.random((200, 2)):
        f = oracle_mod.find_face(flat, 0, x, y)
        assert f == int(x * 7) + int(y * 5) * 7                 # n fastest: f = n + m*Nx (meshQuad.jl:139,151)
    assert oracle_mod.find_face(flat, -1, 0.3, 0.3) == 0
    assert oracle_mod.find_face(flat, -1, 1.3, 0.3) == -1
    assert oracle_mod.find_face(flat, 0, -0.1, 0.3) == -1


def test_volume_sampler_moments(oracle_mod, rthx_mod):
    flat = rthx_mod.flatten_domain(rthx_mod.meshes.cfg1())
    e = flat.n_surfaces + 60
    r = oracle_mod.emit(flat, e, 400_000)
    # isotropic 3-D direction projected on the plane: <dx^2> = <dy^2> = 1/3, <|d|> = pi/4
    assert abs((r[:, 2] ** 2).mean() - 1 / 3) < 3e-3
    assert abs((r[:, 3] ** 2).mean() - 1 / 3) < 3e-3
    assert abs(np.hypot(r[:, 2], r[:, 3]).mean() - math.pi / 4) < 3e-3
    # points uniform inside the cell (cell 60 = lattice (5,5) of the 11x11 square)
    x0, y0, h = 5 / 11, 5 / 11, 1 / 11
    assert r[:, 0].min() >= x0 and r[:, 0].max() <= x0 + h and r[:, 1].min() >= y0 and r[:, 1].max() <= y0 + h
    assert abs(r[:, 0].mean() - (x0 + h / 2)) < 2e-4 and abs(r[:, 1].mean() - (y0 + h / 2)) < 2e-4


def test_surface_sampler_moments(oracle_mod, rthx_mod):
    flat = rthx_mod.flatten_domain(rthx_mod.meshes.cfg1())
    r = oracle_mod.emit(flat, 0, 400_000)                       # surface 0 = bottom wall of cell 0, normal +y
    assert abs(r[:, 3].mean() - 2 / 3) < 3e-3                   # Lambert: <d.n> = 2/3
    assert abs(r[:, 2].mean()) < 3e-3
    assert np.all(r[:, 3] > 0)
    assert r[:, 0].min() >= 0 and r[:, 0].max() <= 1 / 11
    assert np.all(r[:, 1] > 0) and np.all(r[:, 1] < 1e-12)      # nudged off the wall towards the cell midpoint


def test_surface_frame_on_rotated_wall(oracle_mod, rthx_mod):
    rtm = rthx_mod.meshes.square_domain(3, rotation_angle=math.pi / 2)   # wall 1 now points along +y, normal -x
    flat = rthx_mod.flatten_domain(rtm)
    r = oracle_mod.emit(flat, 0, 50_000)
    assert abs(r[:, 2].mean() + 2 / 3) < 1e-2 and np.all(r[:, 2] < 0)
