"""Host tables of the generic locator (csrc/rthx_grid.h) — no GPU.

The device's grid is not the reference's (spatialAccelerations.jl:72-89): it is finer, lists only the polygons that meet a bucket and
marks buckets lying wholly inside one face, whose points are located without any vertex test.  What must hold is the reference's
RESULT: findFaceUniformGrid2D (findFace2D.jl:2-27) returns the first face, in ascending order, whose crossing-number test accepts the
point.  tests/native/grid_check.cpp replays the device's locator on the CPU, statement for statement, and compares it with that scan
over all faces of the set — on random points, points next to vertices and points next to bucket corners — and checks the wall lookup
from vertices (traceRay.jl:51) against the form with unit normals."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rthx
from rthx._abi import rthx_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("grid") / "libgrid_check.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "raytraceheattransfer.jl_b200", "csrc"), "-I", cuda_inc, "-o", so,
                           os.path.join(ROOT, "tests", "native", "grid_check.cpp")])
    lib = C.CDLL(so)
    lib.rthx_grid_check.argtypes = [C.POINTER(rthx_mesh), C.c_int, C.c_uint64, C.POINTER(C.c_double)]
    lib.rthx_grid_check.restype = C.c_int

    def run(flat, points_per_set, seed=11):
        st = (C.c_double * 8)()
        assert lib.rthx_grid_check(C.byref(flat.c), points_per_set, seed, st) == 0
        keys = ("points", "mismatches", "sole", "candidate_tests", "found", "buckets", "walls", "wall_mismatches")
        return dict(zip(keys, list(st)))
    return run


@pytest.mark.parametrize("name,pts", [("cfg1", 40000), ("cfg5", 4000), ("cfg3", 60000)])
def test_grid_locator_equals_full_scan(checker, name, pts):
    flat = rthx.flatten_domain(getattr(rthx.meshes, name)())
    st = checker(flat, pts)
    assert st["points"] > 0 and st["mismatches"] == 0, st
    assert st["walls"] > 0 and st["wall_mismatches"] == 0, st
    # the point of the layout: most located points need no vertex test at all on quadrilateral meshes
    if name != "cfg5":
        assert st["sole"] / st["found"] > 0.6, st
    assert st["buckets"] <= 200 * (flat.n_cells + flat.c.n_coarse) + 256 * (1 + flat.c.n_coarse), st


def test_grid_locator_on_triangles_and_flat_cells(checker):
    st = checker(rthx.flatten_domain(rthx.meshes.circle_domain(N_seg=7, Ndim=5)), 4000, seed=3)
    assert st["mismatches"] == 0 and st["wall_mismatches"] == 0, st
    st = checker(rthx.flatten_domain(rthx.meshes.two_quads_domain(skew=0.3)), 20000, seed=4)      # skewed quadrilaterals
    assert st["mismatches"] == 0 and st["wall_mismatches"] == 0, st
    # test_2d_diffusion.jl:19-23 — cells of 1000:1
    st = checker(rthx.flatten_domain(rthx.meshes.diffusion_slab_domain()), 60000, seed=5)
    assert st["mismatches"] == 0 and st["wall_mismatches"] == 0, st
    assert st["candidate_tests"] / st["points"] < 2.0, st     # anisotropic buckets: a 1000:1 cell does not flood its neighbours' buckets


def test_grid_fine_knob(checker, monkeypatch):
    flat = rthx.flatten_domain(rthx.meshes.cfg1())
    monkeypatch.setenv("RTHX_GRID_FINE", "2")
    st2 = checker(flat, 20000)
    monkeypatch.setenv("RTHX_GRID_FINE", "12")
    st12 = checker(flat, 20000)
    assert st2["mismatches"] == 0 and st12["mismatches"] == 0
    assert st12["sole"] > st2["sole"] and st12["buckets"] > st2["buckets"]
