"""No-GPU checks of the boundary: libparops.so loads and exports every symbol include/parops.h
declares; host integer entry points (no device work) agree bit-for-bit with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parops
from oracle import parops_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return parops._lib.Lib(g.LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "parops.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(po_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib.dll, n), "libparops.so does not export %s" % n
        assert n in parops._lib.SIGNATURES, "no ctypes signature for %s" % n


def test_version_and_status_strings(lib):
    assert lib.po_version() == 100
    assert lib.po_status_string(0) == b"ok"
    assert b"mismatch" in lib.po_status_string(-2)


@pytest.mark.parametrize("g,p", [(10, 8), (10, 3), (7, 7), (5, 8), (128, 8), (0, 2)])
def test_local_size(lib, g, p):
    for r in range(p):
        assert lib.po_local_size(g, r, p) == O.local_size(g, r, p)


def _plan(lib, gs, din, dout, rank):
    N = len(gs)
    li, lo = (C.c_int64 * N)(), (C.c_int64 * N)()
    cap, rec = 1024, 1 + 2 * N
    sb, rb = (C.c_int64 * (cap * rec))(), (C.c_int64 * (cap * rec))()
    ns, nr = C.c_int32(), C.c_int32()
    lib.check(lib.po_repartition_plan(N, (C.c_int64 * N)(*gs), (C.c_int32 * N)(*din), (C.c_int32 * N)(*dout), rank, li, lo, sb, cap,
                                      C.byref(ns), rb, cap, C.byref(nr)))

    def un(buf, n):
        return [(int(buf[k * rec]), tuple((int(buf[k * rec + 1 + 2 * d]), int(buf[k * rec + 2 + 2 * d])) for d in range(N))) for k in range(n)]

    return tuple(li), tuple(lo), un(sb, ns.value), un(rb, nr.value)


@pytest.mark.parametrize("gs,din,dout", [
    ((10, 10, 10), (2, 2, 2), (8, 1, 1)),            # test/distributed.jl:45-47 (uneven 10 over 8)
    ((20, 16, 16, 128, 8), (1, 1, 1, 8, 1), (1, 1, 1, 1, 8)),  # config C4 z-slabs -> t-slabs
    ((7, 5), (3, 1), (1, 3)),
    ((4, 4), (2, 2), (2, 2)),
    ((3,), (4,), (4,)),                               # more ranks than elements: empty blocks
])
def test_repartition_plan_bit_exact(lib, gs, din, dout):
    P = int(np.prod(din))
    total_msgs = 0
    for r in range(P):
        li, lo, send, recv = _plan(lib, gs, din, dout, r)
        oli, olo, osend, orecv = O.repartition_plan(list(din), list(dout), gs, r)
        assert li == tuple(oli) and lo == tuple(olo)
        assert send == [(p, tuple(b)) for p, b in osend]
        assert recv == [(p, tuple(b)) for p, b in orecv]
        total_msgs += len(send)
    if gs == (10, 10, 10):
        assert total_msgs == 32  # 32 send boxes in total for the reference's own test case


def test_error_behaviour(lib):
    """Non-zero status + message, never an abort (ParException in the Julia shim)."""
    ns, nr = C.c_int32(), C.c_int32()
    st = lib.po_repartition_plan(2, (C.c_int64 * 2)(4, 4), (C.c_int32 * 2)(2, 2), (C.c_int32 * 2)(3, 1), 0, None, None, None, 0,
                                 C.byref(ns), None, 0, C.byref(nr))
    assert st == -2 and b"prod(dims_in)" in lib.po_last_error()
    with pytest.raises(parops.ParException):
        lib.check(st)
    # block epilogue: argument checks happen before any device work (no context: invalid-argument status, message set)
    st = lib.po_block_epilogue(None, 0, 20, 20, 128, None, None, None, None, 1, None, None)
    assert st < 0 and b"null argument" in lib.po_last_error()


# ---- randomised: the integer bookkeeping must be bit-exact for EVERY grid, not only the listed ones ---------------------------
def _factorisations(P, N, rng):
    """a random way of writing P as a product of N positive factors"""
    dims = [1] * N
    p = P
    f = 2
    primes = []
    while p > 1:
        while p % f == 0:
            primes.append(f)
            p //= f
        f += 1
    for q in primes:
        dims[int(rng.integers(N))] *= q
    return tuple(dims)


def test_repartition_plan_bit_exact_random_grids(lib):
    """300 random (global size, dims_in, dims_out) triples with 1-4 array dims, 1-12 ranks, uneven splits and empty blocks: local sizes, send and
    recv boxes (peer, per-axis 1-based ranges, first axis of the peer product fastest) of every rank against the oracle's restatement of
    src/ParRepartition.jl:14-110 -- bit for bit."""
    rng = np.random.default_rng(1234)
    checked = 0
    for _ in range(300):
        N = int(rng.integers(1, 5))
        P = int(rng.integers(1, 13))
        gs = tuple(int(rng.integers(1, 15)) for _ in range(N))
        din, dout = _factorisations(P, N, rng), _factorisations(P, N, rng)
        for r in range(P):
            li, lo, send, recv = _plan(lib, gs, din, dout, r)
            oli, olo, osend, orecv = O.repartition_plan(list(din), list(dout), gs, r)
            assert li == tuple(oli) and lo == tuple(olo), (gs, din, dout, r)
            assert send == [(p, tuple(b)) for p, b in osend], (gs, din, dout, r)
            assert recv == [(p, tuple(b)) for p, b in orecv], (gs, din, dout, r)
            checked += 1
    assert checked >= 300


def test_local_size_random(lib):
    rng = np.random.default_rng(5)
    for _ in range(500):
        g, p = int(rng.integers(0, 1000)), int(rng.integers(1, 70))
        sizes = [lib.po_local_size(g, r, p) for r in range(p)]
        assert sizes == [O.local_size(g, r, p) for r in range(p)]
        assert sum(sizes) == g and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)  # balanced, remainder on the low ranks
