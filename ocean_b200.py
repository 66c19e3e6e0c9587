"""Import shim: the package directory `oceananigans.jl_b200/` is not a valid Python identifier, so it is loaded
here under the module name `ocean_b200` (`import ocean_b200 as ob`)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "oceananigans.jl_b200")
_name = "ocean_b200"
_spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_name] = _mod
_spec.loader.exec_module(_mod)
