"""The Julia shim cannot be executed here (no Julia in the image): keep it honest statically.  Every symbol it `ccall`s must be
declared in include/parops.h with the same number of arguments and be exported by the built library."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "parametricoperators.jl_b200", "julia", "ParametricOperatorsB200.jl")
HEADER = os.path.join(ROOT, "include", "parops.h")


def _header_decls():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)   # drop comments
    decls = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|const char\*|void)\s+(po_\w+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        decls[m.group(1)] = len(args)
    return decls


def _shim_calls():
    j = open(SHIM).read()
    calls = {}
    for m in re.finditer(r"ccall\(sym\(:(\w+)\),\s*\w+,\s*\(([^)]*)\)", j, re.S):
        calls[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    return calls


def test_every_ccall_matches_the_header():
    decls, calls = _header_decls(), _shim_calls()
    assert len(calls) >= 12
    for name, nargs in calls.items():
        assert name in decls, "%s is ccall'ed by the Julia shim but not declared in parops.h" % name
        assert decls[name] == nargs, "%s: header takes %d arguments, the shim passes %d" % (name, decls[name], nargs)


def test_every_ccall_symbol_is_exported():
    lib_path = os.path.join(ROOT, "parametricoperators.jl_b200", "csrc", "libparops.so")
    if not os.path.exists(lib_path):
        pytest.skip("libparops.so not built")
    # symbol table only: loading the library needs libcudart, which the CPU container has
    lib = ctypes.CDLL(lib_path)
    for name in _shim_calls():
        assert hasattr(lib, name), name
