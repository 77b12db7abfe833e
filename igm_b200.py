"""Import alias: ``import igm_b200`` loads the package in ``image-generation-models_b200/``
(the directory name required by the project layout is not a valid Python identifier)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image-generation-models_b200")
_spec = importlib.util.spec_from_file_location(
    "igm_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["igm_b200"] = _mod
_spec.loader.exec_module(_mod)
