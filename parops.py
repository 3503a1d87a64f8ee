"""Repo-root loader: ``import parops`` gives the package in ``parametricoperators.jl_b200/``
(whose directory name, mirroring the reference package name, is not importable by name)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parametricoperators.jl_b200")
_spec = importlib.util.spec_from_file_location("parops", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["parops"] = _mod
_spec.loader.exec_module(_mod)
