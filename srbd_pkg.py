"""Loader for the product package.

The package directory is named `srbd-nmpc-solver_b200/` (the name the build contract fixes); a
hyphen is not importable, so this shim registers it in sys.modules as `srbd_nmpc_solver_b200`.
"""
import importlib.util
import os
import sys

_NAME = "srbd_nmpc_solver_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "srbd-nmpc-solver_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
