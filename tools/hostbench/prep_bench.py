"""Times the host half of rthx_create (mesh preparation) on this machine's CPU; no GPU needed."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rthx
from rthx._abi import rthx_mesh
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "libprep_bench.so")
csrc = os.path.join(ROOT, "raytraceheattransfer.jl_b200", "csrc")
srcs = [os.path.join(here, "prep_bench.cu")] + [os.path.join(csrc, f) for f in ("rthx_kernels.cu", "rthx_smooth.cu", "rthx_solve.cu")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2", "-shared",
                       "-I", os.path.join(ROOT, "include"), "-I", csrc, "-o", so] + srcs)
L = C.CDLL(so)
L.rthx_prep_bench.restype = C.c_double
L.rthx_prep_bench.argtypes = [C.POINTER(rthx_mesh), C.c_int, C.c_int]
for name in sys.argv[1:] or ["cfg1", "cfg3", "cfg5"]:
    flat = rthx.flatten_domain(getattr(rthx.meshes, name)())
    print(name, "N =", flat.n_elements, "prepare_mesh best of 20: %.3f ms" % L.rthx_prep_bench(C.byref(flat.c), 20, 0))
