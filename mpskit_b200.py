"""Import shim: the product package lives in the directory ``mpskit.jl_b200/`` (a name Python cannot import
directly because of the dot); ``import mpskit_b200`` loads it from there."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpskit.jl_b200")
_spec = importlib.util.spec_from_file_location("mpskit_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mpskit_b200"] = _mod
_spec.loader.exec_module(_mod)
