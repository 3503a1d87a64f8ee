"""Host-side mirror of the operator API (no GPU): traits, constructors, order heuristic,
init, parameterisation, lowering to the po_node wire format, distribute."""
import numpy as np
import pytest

import parops
from parops import (ComplexF32, ComplexF64, Float32, Float64, ParCompose, ParDFT, ParDiagonal, ParException, ParIdentity, ParKron,
                    ParMatrix, ParRestriction, ParTensor, adjoint, compose, init, kron)
from oracle import parops_oracle as O
from tests import parity_cases as pc

TYPES = [Float32, Float64, ComplexF32, ComplexF64]


@pytest.mark.parametrize("T", TYPES)
def test_types_and_sizes(T):
    """test/types_and_sizes.jl:17-33 (traits only; the applies are in the parity suites)."""
    CT = parops.Complex(T)
    cases = [(ParIdentity(T, 4), 4, 4, T, T), (ParDiagonal(T, 4), 4, 4, T, T), (ParMatrix(T, 4, 8), 8, 4, T, T),
             (ParDFT(T, 8), 8, 8 if T.is_complex else 5, T, CT), (ParRestriction(T, 8, [(1, 2), (7, 8)]), 8, 4, T, T)]
    for A, dom, ran, D, R in cases:
        assert (A.Domain, A.Range, A.D, A.R) == (dom, ran, D, R)
        B = adjoint(A)
        assert (B.Domain, B.Range, B.D, B.R) == (ran, dom, R, D)
        assert adjoint(B) is A or isinstance(A, ParIdentity)


def test_defaults_match_reference():
    assert ParDFT(8).D is ComplexF64                      # src/ParDFT.jl:18
    assert ParMatrix(3, 4).D is Float64 and ParDiagonal(5).D is Float64 and ParIdentity(3).D is Float64
    assert ParRestriction(8, [(1, 2)]).D is Float64        # src/ParRestriction.jl:7
    assert ParKron(ParDFT(4)) .n == 4 and isinstance(ParCompose(ParDFT(4)), ParDFT)  # single-op collapse
    assert isinstance(kron(ParIdentity(Float32, 2), ParIdentity(Float32, 3)), ParIdentity)  # src/ParIdentity.jl:35
    assert parops.subset_type(ComplexF64, Float32) is Float32 and parops.subset_type(ComplexF32, ComplexF64) is ComplexF32


def test_compose_asserts():
    with pytest.raises(ParException):
        compose(ParMatrix(Float32, 4, 5), ParMatrix(Float32, 4, 5))          # Domain != Range
    with pytest.raises(ParException):
        compose(ParMatrix(ComplexF32, 4, 5), ParMatrix(Float32, 5, 5))       # DDT != RDT
    C = compose(ParMatrix(ComplexF32, 4, 5), ParDFT(Float32, 8))
    assert (C.D, C.R, C.Domain, C.Range) == (Float32, ComplexF32, 8, 4)
    assert C.P is parops.Parametric and C.L is parops.Linear


def _mirror(op):
    """product tree -> same-structure oracle tree"""
    if isinstance(op, ParKron):
        return O.ParKron(*[_mirror(c) for c in op.ops])
    if isinstance(op, ParCompose):
        return O.ParCompose(*[_mirror(c) for c in op.ops])
    if isinstance(op, ParDFT):
        return O.ParDFT(op.D.np, op.n)
    if isinstance(op, ParRestriction):
        return O.ParRestriction(op.D.np, op.n, op.ranges)
    if isinstance(op, ParMatrix):
        return O.ParMatrix(op.D.np, op.m, op.n)
    if isinstance(op, ParDiagonal):
        return O.ParDiagonal(op.D.np, op.n)
    if isinstance(op, ParIdentity):
        return O.ParIdentity(op.D.np, op.n)
    raise TypeError(op)


def test_kron_order_matches_oracle_on_random_trees():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(2, 6))
        ops, real_used = [], False
        for _ in range(n):
            kind = int(rng.integers(0, 5))
            T = Float32
            CT = ComplexF32
            if kind == 0:
                ops.append(ParMatrix(CT, int(rng.integers(1, 9)), int(rng.integers(1, 9))))
            elif kind == 1:
                ops.append(ParDFT(CT, int(rng.integers(2, 9))))
            elif kind == 2:
                m = int(rng.integers(2, 9))
                ops.append(compose(ParRestriction(CT, m, [(1, int(rng.integers(1, m + 1)))]), ParDFT(CT, m)))
            elif kind == 3 and not real_used:
                ops.append(ParDFT(T, int(rng.integers(2, 9))))
                real_used = True
            else:
                ops.append(ParIdentity(CT, int(rng.integers(1, 5))))
        if real_used:
            ops = [o if o.D is Float32 else o for o in ops]
            # a real r2c factor needs every other factor to be complex->complex: already so
        try:
            K = ParKron(*ops)
        except ParException:
            continue
        assert K.order == O.ParKron(*[_mirror(o) for o in ops]).order


def test_fno_structure_matches_survey():
    S = parops.spectral_convolution(Float32, [64, 64, 64, 32], [[(1, 8), (57, 64)]] * 3 + [[(1, 8)]], 20)
    assert S.Domain == S.Range == 20 * 64 ** 3 * 32 and S.D is Float32 and S.R is Float32
    dft_all = S.ops[-1]
    assert dft_all.order == [2, 1] and dft_all.Range == 20 * 16 ** 3 * 8
    inner = dft_all.ops[0]
    assert inner.ops[1].order == [1, 4, 3, 2]             # rfft on the slowest dim first, then right-to-left
    assert S.ops[1].order == [2, 1] and S.ops[1].ops[0].n == 32768


@pytest.mark.parametrize("T", TYPES)
def test_init_distributions(T):
    M, D, B = ParMatrix(T, 30, 50), ParDiagonal(T, 64), ParMatrix(T, 7, 1)
    W = ParTensor(T, (1, 2, 3, 4, 5), (4, 3, 2, 5, 2), (2, 3, 4, 5), (3, 2, 5, 2), (1, 3, 4, 5), (4, 2, 5, 2))
    th = init(M, D, B, W, rng=0)
    assert th[M].shape == (30, 50) and th[M].dtype == T.np and th[M].flags.f_contiguous
    assert th[D].shape == (64,) and th[B].shape == (7, 1) and not th[B].any()            # bias = zeros (src/ParMatrix.jl:24-27)
    assert th[W].shape == (4, 3, 2, 5, 2)
    if T.is_complex:
        assert (th[M].real >= 0).all() and (th[M].real <= 1 / np.sqrt(1500) + 1e-6).all()   # rand/sqrt(mn)
        assert (th[D].real >= 0).all() and (th[D].imag < 1).all()
    else:
        lim = 0.5 * np.sqrt(24.0 / 80)
        assert (np.abs(th[M]) <= lim + 1e-6).all() and abs(float(th[M].mean())) < 0.05    # (rand-0.5)*sqrt(24/(m+n))
    assert (np.abs(th[W]) <= np.sqrt(2) / np.sqrt(240) + 1e-6).all()


def test_parameterize_semantics():
    A, D = ParMatrix(ComplexF32, 4, 4), ParDiagonal(ComplexF32, 6)
    K = kron(D, A)
    assert K.P is parops.Parametric
    th = init(K, rng=1)
    Kt = K(th)
    assert Kt.P is parops.Parameterized and isinstance(Kt.ops[1], parops.ParParameterized) and Kt.ops[1].params is th[A]
    Ka = adjoint(K)(th)                                    # src/ParAdjoint.jl:19: adjoint leaf keeps the WHOLE dict
    leaf, adj, tensor = Ka.ops[1].leaf_and_tensor()
    assert leaf is A and adj and tensor is th[A]
    assert adjoint(K).order == [1, 2] and Ka.order == [2, 1]  # rebuild re-derives the order (src/ParKron.jl:81)
    with pytest.raises(ParException):
        K.apply(np.zeros(24, dtype=np.complex64))           # unbound parameters
    with pytest.raises(ParException):
        ParMatrix(Float32, 3, 3)(dict())                    # missing key


def test_lowering_wire_format():
    from parops.engine import Lowered, lower
    from parops._lib import PO_NODE_COMPOSE, PO_NODE_DFT, PO_NODE_KRON, PO_NODE_RESTRICTION, PO_NODE_IDENTITY
    A = parops.kron(compose(ParRestriction(ComplexF32, 5, [(1, 2)]), ParDFT(Float32, 8)), ParIdentity(Float32, 3))
    lw = Lowered()
    root = lower(A, lw)
    kinds = [n[0] for n in lw.nodes]
    assert kinds == [PO_NODE_RESTRICTION, PO_NODE_DFT, PO_NODE_COMPOSE, PO_NODE_IDENTITY, PO_NODE_KRON] and root == 4
    kron_node = lw.nodes[root]
    assert lw.aux[kron_node[7]:kron_node[7] + kron_node[8]] == [2, 1]           # application order
    assert lw.child_index[kron_node[5]:kron_node[5] + 2] == [2, 3]
    assert lw.nodes[0][9:] == (5, 2) and lw.aux[lw.nodes[0][7]:lw.nodes[0][7] + 4] == [5, 1, 1, 2]
    # adjoint: child adjoints, reversed order, inverse DFT flag
    lw2 = Lowered()
    lower(adjoint(A), lw2)
    dft = [n for n in lw2.nodes if n[0] == PO_NODE_DFT][0]
    assert dft[3] == 1 and dft[1] == ComplexF32.code and dft[2] == Float32.code and dft[9:] == (5, 8)


def test_distribute_structure_matches_oracle():
    T, CT = Float64, ComplexF64

    def rf(ns, Tin, n, rng):
        Fo = ns.ParDFT(Tin, n)
        return ns.compose(ns.ParRestriction(CT.np if ns is pc.OracleNS else CT, Fo.Range, rng), Fo)

    def build(ns):
        t, ct = (T.np, CT.np) if ns is pc.OracleNS else (T, CT)
        return ns.ParKron(rf(ns, t, 8, [(1, 3)]), rf(ns, ct, 8, [(1, 2), (7, 8)]), rf(ns, ct, 4, [(1, 2)]), ns.ParIdentity(t, 3))

    Ko, Kp = build(pc.OracleNS), build(pc.ProductNS)
    assert Ko.order == Kp.order
    lib = parops._lib.default_lib()
    for dims in ([1, 1, 2, 1], [1, 2, 2, 1], [1, 1, 4, 1]):
        P = int(np.prod(dims))
        stages = O.distribute_kron(Ko, dims)
        for r in range(P):
            parops.set_world(r, P)
            import parops.distributed as D
            orig = D.ParRepartition.__init__
            D.ParRepartition.__init__ = lambda self, *a, **k: orig(self, *a, lib=lib, **k)
            try:
                Cp = parops.distribute(Kp, dims, rank=r)
            finally:
                D.ParRepartition.__init__ = orig
            got = list(reversed(Cp.ops))  # application order
            assert len(got) == len(stages)
            for (kind, obj), g in zip(stages, got):
                if kind == "repartition":
                    assert isinstance(g, parops.ParRepartition)
                    assert (g.dims_in, g.dims_out, g.global_size) == (obj.dims_in, obj.dims_out, obj.global_size)
                    assert g.local_size_in == obj.local_in(r) and g.local_size_out == obj.local_out(r)
                    assert g.send == [(p, tuple(b)) for p, b in obj.plans[r][2]]
                else:
                    assert isinstance(g, ParKron) and (g.Domain, g.Range) == (obj[r].Domain, obj[r].Range)
                    assert g.order == obj[r].order
    parops.set_world(0, 1)


def test_no_cpu_fallback():
    """On a box without a GPU the product path must refuse to run (fail loudly)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ParException, match="no CPU fallback"):
        parops.default_engine()
    with pytest.raises(ParException):
        ParDFT(Float32, 8) * np.zeros(8, dtype=np.float32)
    import torch as t
    with pytest.raises(ParException):
        ParDFT(Float32, 8) * t.zeros(8)


def test_range_axes_and_local_size_match_the_oracle_on_random_grids():
    """src/ParCommon.jl:45-64 (a15, a16: integer, bit-exact): the host mirror against the oracle on 300 random cart grids, plus the
    defining properties -- blocks tile 1..n without gaps, sizes differ by at most one, the remainder sits on the low coordinates."""
    from oracle import parops_oracle as O
    rng = np.random.default_rng(77)
    for _ in range(300):
        N = int(rng.integers(1, 5))
        dims = [int(rng.integers(1, 9)) for _ in range(N)]
        shape = [int(rng.integers(0, 40)) for _ in range(N)]
        got = parops.range_axes(dims, shape)
        assert got == O.range_axes(dims, shape)
        for d in range(N):
            assert [parops.local_size(shape[d], i, dims[d]) for i in range(dims[d])] == [b - a + 1 for a, b in got[d]]
            flat = [i for a, b in got[d] for i in range(a, b + 1)]
            assert flat == list(range(1, shape[d] + 1))
            sizes = [b - a + 1 for a, b in got[d]]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
