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


# ---- argument TYPES, not only counts -------------------------------------------------------------------------------
def _c_category(arg: str) -> str:
    a = arg.strip()
    if "*" in a or a.startswith(("po_exchange_fn", "po_collective_fn")):
        return "ptr"
    t = a.replace("const", "").split()[0]
    return {"int32_t": "i32", "int64_t": "i64", "uint32_t": "u32", "size_t": "size"}[t]


def _jl_category(arg: str) -> str:
    a = arg.strip()
    if a.startswith(("Ptr{", "CuPtr{", "Ref{")) or a == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Int64": "i64", "UInt32": "u32", "Csize_t": "size"}[a]


def _split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        if ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a for a in out if a.strip()]


def test_every_ccall_argument_type_matches_the_header():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    cdecl = {}
    for m in re.finditer(r"\b(int32_t|int64_t|const char\*|void)\s+(po_\w+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        args = [a for a in _split_top(m.group(3)) if a.strip() != "void"]
        cdecl[m.group(2)] = (m.group(1), [_c_category(a) for a in args])
    j = open(SHIM).read()
    seen = 0
    for m in re.finditer(r"ccall\(sym\(:(\w+)\),\s*(\w+),\s*\(", j):
        name, ret = m.group(1), m.group(2)
        # the argument-type tuple: balanced parentheses from the "(" that ends the match
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(j[i], 0)
            i += 1
        jl_args = [_jl_category(a) for a in _split_top(j[m.end():i - 1])]
        cret, c_args = cdecl[name]
        assert jl_args == c_args, "%s: header %s, shim %s" % (name, c_args, jl_args)
        assert ret == {"int32_t": "Int32", "int64_t": "Int64", "const char*": "Cstring"}[cret], (name, ret, cret)
        seen += 1
    assert seen >= 20


def test_every_module_the_shim_uses_is_imported():
    j = open(SHIM).read()
    code = re.sub(r"#.*", "", j)
    used = set(re.findall(r"\b(MPI|CUDA|Libdl|ChainRulesCore)\.", code))
    for mod in used | {"MPI", "CUDA", "Libdl", "ChainRulesCore"}:
        assert re.search(r"^using %s\b" % mod, code, re.M), "the shim uses %s. but never imports it" % mod
    # names pulled from ParametricOperators must be listed in the `using ParametricOperators:` clause when used unqualified
    clause = re.search(r"using ParametricOperators:(.*?)\nusing", j, re.S).group(1)
    for name in ("ParReduce", "ParRepartition", "ParBroadcasted", "ParTensor", "ParException"):
        assert re.search(r"\b%s\b" % name, clause), name


def test_plan_cache_is_keyed_on_the_lowered_signature_and_plans_are_finalized():
    j = open(SHIM).read()
    assert "IdDict{Any,Dict{Int,Plan}}" not in j          # the round-1 cache keyed plans on the (per-call) operator object
    assert re.search(r"PlanKey\s*=\s*Tuple\{Vector\{PoNode\}", j)
    assert "po_plan_destroy" in j and "finalizer" in j
    # the operator cotangent mirrors the tree (ops / op / params fields), not a flat `params` on a ParCompose
    assert "tangent_tree" in j and "Tangent{typeof(A)}(; ops" in j
