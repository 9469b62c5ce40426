"""`import rthx` — repo-root shim that loads the package living in `raytraceheattransfer.jl_b200/`.

The directory name (fixed by the project layout) contains a dot and so cannot be imported by name; this shim
registers it under the importable alias `rthx` with its sub-modules (`rthx.domain`, `rthx.tracing`, ...).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "raytraceheattransfer.jl_b200")
_spec = _ilu.spec_from_file_location("rthx", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["rthx"] = _mod
_spec.loader.exec_module(_mod)
