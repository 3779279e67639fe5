"""Import shim: the package directory is named ``itensorsgpu.jl_b200`` (not a valid Python
identifier), so load it explicitly and expose it as ``tn`` / ``itensorsgpu_jl_b200``."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "itensorsgpu.jl_b200")
_NAME = "itensorsgpu_jl_b200"

if _NAME in sys.modules:
    tn = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    tn = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = tn
    _spec.loader.exec_module(tn)
