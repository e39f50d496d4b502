"""Import alias: ``import inflexion_b200`` loads the package directory ``alpha-zero-general-inflexion_b200/``.

The directory name is fixed by the build contract but is not a valid Python identifier, so this shim
registers it under an importable name (``inflexion_b200`` and its submodules).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alpha-zero-general-inflexion_b200")
_spec = importlib.util.spec_from_file_location(
    "inflexion_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["inflexion_b200"] = _mod
_spec.loader.exec_module(_mod)
