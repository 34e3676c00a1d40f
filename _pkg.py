"""Loader for the product package.  Its directory is named `sdemodels.jl_b200` (after the reference,
SDEModels.jl); a dot is not importable, so it is registered under the module name `sdemodels_jl_b200`."""
import importlib.util
import os
import sys

NAME = "sdemodels_jl_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "sdemodels.jl_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
