"""Import shim: the product package lives in ``stormwater-management-model_b200/`` (a directory
name Python cannot import directly); this registers it as the package ``swmm_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stormwater-management-model_b200")
_spec = importlib.util.spec_from_file_location(
    "swmm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["swmm_b200"] = _mod
_spec.loader.exec_module(_mod)
