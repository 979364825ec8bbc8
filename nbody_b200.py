"""Import shim: the package directory is named ``nbodysimulator.jl_b200`` (a dot cannot appear
in a Python module name), so ``import nbody_b200`` loads that directory as the package
``nbody_b200``; ``import nbody_b200.api`` etc. then resolve inside it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nbodysimulator.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "nbody_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nbody_b200"] = _mod
_spec.loader.exec_module(_mod)
