"""The C-ABI library: builds, loads, exports every symbol of include/rthx.h, struct layouts agree with the header,
and — with no GPU — fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(cuda_lib, rthx_mod):
    from rthx._abi import EXPORTED_SYMBOLS
    header = open(os.path.join(ROOT, "include", "rthx.h")).read()
    declared = set(re.findall(r"\b(rthx_[A-Za-z0-9_]+)\s*\(", header))
    assert declared == set(EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(cuda_lib, name) is not None
    assert cuda_lib.rthx_version() == 3                                 # 0.3: rthx_smooth_stats.converged, rthx_create_multi, rthx_counts_csc, ...


def test_struct_layouts_match_header(rthx_mod):
    from rthx import _abi
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rthx.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(rthx_mesh), sizeof(rthx_trace_args), sizeof(rthx_rec_out), sizeof(rthx_stats), sizeof(rthx_info));
  printf("%zu %zu %zu %zu\n", offsetof(rthx_trace_args, nudge), offsetof(rthx_trace_args, bins), offsetof(rthx_trace_args, rec_ids), offsetof(rthx_trace_args, row_chunks));
  printf("%zu %zu\n", offsetof(rthx_mesh, uniform_beta), offsetof(rthx_stats, n_launches));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(rthx_solve_args), sizeof(rthx_solve_stats), sizeof(rthx_smooth_stats),
         offsetof(rthx_solve_args, F_dense), offsetof(rthx_solve_args, atol), offsetof(rthx_solve_stats, matvec_bytes));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    got = [int(x) for x in out]
    A = _abi.rthx_trace_args
    want = [C.sizeof(_abi.rthx_mesh), C.sizeof(A), C.sizeof(_abi.rthx_rec_out), C.sizeof(_abi.rthx_stats), C.sizeof(_abi.rthx_info),
            A.nudge.offset, A.bins.offset, A.rec_ids.offset, A.row_chunks.offset,
            _abi.rthx_mesh.uniform_beta.offset, _abi.rthx_stats.n_launches.offset,
            C.sizeof(_abi.rthx_solve_args), C.sizeof(_abi.rthx_solve_stats), C.sizeof(_abi.rthx_smooth_stats),
            _abi.rthx_solve_args.F_dense.offset, _abi.rthx_solve_args.atol.offset, _abi.rthx_solve_stats.matvec_bytes.offset]
    assert got == want


def test_no_gpu_fails_loudly(cuda_lib, rthx_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    flat = rthx_mod.flatten_domain(rthx_mod.meshes.square_domain(3))
    with pytest.raises(rthx_mod.RthxError, match="no CUDA device|CPU fallback"):
        rthx_mod.DeviceTracer(flat, device=0)
    rtm = rthx_mod.meshes.square_domain(3)
    with pytest.raises(rthx_mod.RthxError):
        rtm(1000, method="exchange", verbose=False)               # the public call has no CPU fallback either
    import numpy as np
    with pytest.raises(rthx_mod.RthxError):
        rthx_mod.solveEquilibrium(rtm, np.eye(rtm.num_elements), verbose=False)   # nor has the equilibrium solve


def test_bad_mesh_rejected_before_touching_the_device(cuda_lib, rthx_mod):
    from rthx._abi import rthx_mesh
    h = C.c_void_p()
    rc = cuda_lib.rthx_create(C.byref(h), C.byref(rthx_mesh()), 0)
    assert rc == 1 and b"NULL or empty" in cuda_lib.rthx_last_error(None)
    assert cuda_lib.rthx_create(None, None, 0) == 1


def test_product_path_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or link it."""
    pkg = os.path.join(ROOT, "raytraceheattransfer.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the CPU oracle", "").replace("CPU oracle", ""), f"{f} mentions the oracle"
    so = os.path.join(pkg, "csrc", "librthx.so")
    if os.path.exists(so):
        out = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_julia_shim_structs_match_the_abi(rthx_mod):
    """julia/RTHXExchange.jl cannot run here (no Julia in the image): its struct declarations are compared field by field —
    name, order, width and kind — with the ctypes structs, which test_struct_layouts_match_header ties to include/rthx.h."""
    from rthx import _abi
    src = open(os.path.join(ROOT, "julia", "RTHXExchange.jl")).read()
    pairs = {"RthxMesh": _abi.rthx_mesh, "RthxTraceArgs": _abi.rthx_trace_args, "RthxRecOut": _abi.rthx_rec_out,
             "RthxStats": _abi.rthx_stats, "RthxSolveArgs": _abi.rthx_solve_args, "RthxSolveStats": _abi.rthx_solve_stats}
    kinds = {"Int32": ("i", 4), "Int64": ("i", 8), "UInt64": ("u", 8), "UInt8": ("u", 1), "Float64": ("f", 8)}

    def ctypes_kind(t):
        if hasattr(t, "contents") or t in (C.c_void_p, C.c_char_p):
            return ("p", C.sizeof(C.c_void_p))
        code = t._type_
        return ({"d": "f", "f": "f"}.get(code, "u" if code in "BHILQ" else "i"), C.sizeof(t))

    for jl_name, ct in pairs.items():
        m = re.search(r"struct\s+%s\b(.*?)\n\s*(?:%s\(\)|end)" % (jl_name, jl_name), src, re.S)
        assert m, jl_name
        fields = re.findall(r"(\w+)::([A-Za-z0-9]+(?:\{[A-Za-z0-9]+\})?)", m.group(1))
        assert [f for f, _ in fields] == [n for n, _ in ct._fields_], jl_name
        for (fname, jt), (_, t) in zip(fields, ct._fields_):
            want = ("p", C.sizeof(C.c_void_p)) if jt.startswith("Ptr{") else kinds[jt]
            assert ctypes_kind(t) == want, (jl_name, fname, jt)
