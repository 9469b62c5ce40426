"""ctypes loader for oracle/librthx_oracle.so — TEST INFRASTRUCTURE (see rthx_oracle.c header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import rthx  # noqa: E402  (repo-root shim that loads raytraceheattransfer.jl_b200)
from rthx._abi import rthx_mesh, rthx_trace_args, rthx_rec_out, c_i32p, c_f64p, c_u64p  # noqa: E402

_LIB = None


class oracle_stats(C.Structure):
    _fields_ = [("n_surface_gas", C.c_uint64), ("n_surface_wall", C.c_uint64), ("n_volume_gas", C.c_uint64),
                ("n_volume_wall", C.c_uint64), ("n_crossings", C.c_uint64), ("n_lost", C.c_uint64),
                ("n_threads", C.c_int32), ("pad_", C.c_int32), ("loop_seconds", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "librthx_oracle.so")
    src = os.path.join(_HERE, "rthx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.rthx_oracle_trace.restype = C.c_int
        L.rthx_oracle_trace.argtypes = [C.POINTER(rthx_mesh), C.POINTER(rthx_trace_args), c_u64p, c_u64p,
                                        C.POINTER(rthx_rec_out), C.c_int, C.POINTER(oracle_stats)]
        L.rthx_oracle_shoot.restype = C.c_int
        L.rthx_oracle_shoot.argtypes = [C.POINTER(rthx_mesh), C.POINTER(rthx_trace_args), C.c_int, C.c_int,
                                        C.c_uint64, c_f64p]
        L.rthx_oracle_emit.restype = C.c_int
        L.rthx_oracle_emit.argtypes = [C.POINTER(rthx_mesh), C.POINTER(rthx_trace_args), C.c_int, C.c_int,
                                       C.c_int64, c_f64p]
        L.rthx_oracle_find_face.restype = C.c_int
        L.rthx_oracle_find_face.argtypes = [C.POINTER(rthx_mesh), C.c_int, C.c_double, C.c_double]
        L.rthx_oracle_dist_to_surface.restype = C.c_double
        L.rthx_oracle_dist_to_surface.argtypes = [C.c_int, c_f64p, c_f64p, C.c_double, C.c_double, C.c_double,
                                                  C.c_double, C.POINTER(C.c_int)]
        L.rthx_oracle_philox4x32_10.restype = None
        L.rthx_oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.rthx_oracle_u52.restype = C.c_double
        L.rthx_oracle_u52.argtypes = [C.c_uint32, C.c_uint32]
        L.rthx_oracle_u23.restype = C.c_float
        L.rthx_oracle_u23.argtypes = [C.c_uint32]
        L.rthx_oracle_set_faithful.restype = None
        L.rthx_oracle_set_faithful.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def trace(flat, rays_per_emitter, seed=0x5EED0001, bins=(0,), nudge=None, n_threads=0, rec_ids=None, rec_bin=0,
          ray_id_offset=0, emitter_rank=0, emitter_world=1, mode=0, faithful=False):
    """Run the oracle on a FlatMesh.  Returns dict(counts[nb,N,N] u64, lost[nb,N] u64, stats, origins, endpoints).
    faithful=True: the reference-faithful TIMING mode (per-thread xoshiro256++, Dict-like row tally) — not the Philox contract."""
    lib().rthx_oracle_set_faithful(1 if faithful else 0)
    args, keep = rthx.make_trace_args(rays_per_emitter, seed=seed, bins=bins, nudge=nudge, rec_ids=rec_ids,
                                      rec_bin=rec_bin, ray_id_offset=ray_id_offset, emitter_rank=emitter_rank,
                                      emitter_world=emitter_world, mode=mode)
    N = flat.n_elements
    nb = len(bins)
    counts = np.zeros((nb, N, N), np.uint64)
    lost = np.zeros((nb, N), np.uint64)
    st = oracle_stats()
    rec = None
    origins = endpoints = None
    if rec_ids:
        cap = len(rec_ids) * int(rays_per_emitter)
        origins = np.zeros((max(cap, 1), 2))
        endpoints = np.zeros((max(cap, 1), 2))
        rec = rthx_rec_out(cap, origins.ctypes.data_as(c_f64p), endpoints.ctypes.data_as(c_f64p), 0)
    rc = lib().rthx_oracle_trace(C.byref(flat.c), C.byref(args), counts.ctypes.data_as(c_u64p),
                                 lost.ctypes.data_as(c_u64p), C.byref(rec) if rec else None, int(n_threads),
                                 C.byref(st))
    if rc != 0:
        raise RuntimeError(f"rthx_oracle_trace failed with status {rc}")
    out = dict(counts=counts, lost=lost, stats=st.as_dict())
    if rec is not None:
        out["origins"] = origins[: rec.n_recorded].copy()
        out["endpoints"] = endpoints[: rec.n_recorded].copy()
    return out


def shoot(flat, emitter, ray_id, band=0, seed=0x5EED0001, nudge=None):
    args, keep = rthx.make_trace_args(1, seed=seed, bins=(band,), nudge=nudge)
    out = np.zeros(6)
    ab = lib().rthx_oracle_shoot(C.byref(flat.c), C.byref(args), int(emitter), int(band), int(ray_id),
                                 out.ctypes.data_as(c_f64p))
    return ab, out


def emit(flat, emitter, n, band=0, seed=0x5EED0001, nudge=None):
    args, keep = rthx.make_trace_args(1, seed=seed, bins=(band,), nudge=nudge)
    out = np.zeros((int(n), 4))
    rc = lib().rthx_oracle_emit(C.byref(flat.c), C.byref(args), int(emitter), int(band), int(n),
                                out.ctypes.data_as(c_f64p))
    if rc != 0:
        raise RuntimeError("rthx_oracle_emit failed")
    return out


def find_face(flat, face_set, x, y):
    return lib().rthx_oracle_find_face(C.byref(flat.c), int(face_set), float(x), float(y))


def dist_to_surface(vx, vy, p, d):
    vx = np.ascontiguousarray(vx, dtype=np.float64)
    vy = np.ascontiguousarray(vy, dtype=np.float64)
    idx = C.c_int(0)
    u = lib().rthx_oracle_dist_to_surface(len(vx), vx.ctypes.data_as(c_f64p), vy.ctypes.data_as(c_f64p),
                                          p[0], p[1], d[0], d[1], C.byref(idx))
    return u, idx.value


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().rthx_oracle_philox4x32_10(c, k, o)
    return tuple(int(x) for x in o)
