"""B200-native spectral-operator apply path with the ParametricOperators.jl operator API.

``import parops`` (repo-root loader) or load this package by path; the directory name
``parametricoperators.jl_b200`` is not a valid Python identifier on purpose (it names the
reference package it is a drop-in for).
"""
from ._lib import ParException  # noqa: F401
from .operators import *  # noqa: F401,F403
from .operators import jtype, kron_order  # noqa: F401
from .engine import (Engine, apply_operator, apply_right, block_epilogue, block_epilogue_pullback, colmajor, cpu, current_engine, default_engine, get_plan,  # noqa: F401
                     gpu, rrule, use_engine)
from .distributed import *  # noqa: F401,F403
from .fno import *  # noqa: F401,F403
from .serialization import *  # noqa: F401,F403
from . import _lib, engine, operators, distributed, fno, serialization  # noqa: F401

__version__ = "0.1.0"
