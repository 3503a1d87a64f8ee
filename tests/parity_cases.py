"""Parity cases shared by the CPU-emulator suite (test_emul_parity.py) and the GPU suite
(test_gpu_parity.py): every case builds the SAME operator tree twice -- once from the oracle
(oracle/parops_oracle.py, reference-order NumPy restatement) and once from the product host
mirror (parops) -- applies both to the same seeded input and compares.

Tolerances are BASELINE.json's: relative L2 <= 1e-5 (Float32 / ComplexF32), <= 1e-12
(Float64 / ComplexF64), bit-exact for Restriction / Repartition index mapping.
"""
import numpy as np

import parops
from oracle import parops_oracle as O

F32, F64, C64, C128 = O.F32, O.F64, O.C64, O.C128
ALL_TYPES = [F32, F64, C64, C128]
TOL = {F32: 1e-5, C64: 1e-5, F64: 1e-12, C128: 1e-12}


class OracleNS:
    ParIdentity, ParDFT, ParRestriction, ParMatrix = O.ParIdentity, O.ParDFT, O.ParRestriction, O.ParMatrix
    ParDiagonal, ParTensor, ParKron = O.ParDiagonal, O.ParTensor, O.ParKron
    kron = staticmethod(lambda *ops: _fold(O.kron, ops))
    compose = staticmethod(lambda *ops: O.ParCompose(*ops))


class ProductNS:
    ParIdentity, ParDFT, ParRestriction, ParMatrix = parops.ParIdentity, parops.ParDFT, parops.ParRestriction, parops.ParMatrix
    ParDiagonal, ParTensor, ParKron = parops.ParDiagonal, parops.ParTensor, parops.ParKron
    kron = staticmethod(parops.kron)
    compose = staticmethod(parops.compose)


def _fold(f, xs):
    acc = xs[0]
    for v in xs[1:]:
        acc = f(acc, v)
    return acc


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num


def randn(rng, T, *shape):
    T = np.dtype(T)
    if T.kind == "c":
        v = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    else:
        v = rng.standard_normal(shape)
    return np.asfortranarray(v.astype(T))


def leaves_oracle(op, out=None):
    out = [] if out is None else out
    kids = op.children()
    if not kids:
        if op.parametric:
            out.append(op)
    for c in kids:
        leaves_oracle(c, out)
    return out


def leaves_product(op, out=None):
    out = [] if out is None else out
    kids = op.children()
    if not kids:
        if op.P is parops.Parametric:
            out.append(op)
    for c in kids:
        leaves_product(c, out)
    return out


def paired_params(A_o, A_p, eng, seed=2):
    """theta for the oracle tree and the SAME values (on the engine's device) for the product tree."""
    theta_o = O.init(A_o, seed=seed)
    lo, lp = leaves_oracle(A_o), leaves_product(A_p)
    assert len(lo) == len(lp)
    theta_p = {}
    for a, b in zip(lo, lp):
        theta_p[b] = eng.to_device(theta_o[a])
    return theta_o, theta_p


def check_case(eng, build, batch=3, seed=0, tol=None, exact=False, flags=0, vector=False):
    A_o, A_p = build(OracleNS), build(ProductNS)
    assert A_o.Domain == A_p.Domain and A_o.Range == A_p.Range
    assert np.dtype(A_o.D) == A_p.D.np and np.dtype(A_o.R) == A_p.R.np
    theta_o, theta_p = paired_params(A_o, A_p, eng)
    rng = np.random.default_rng(seed)
    x = randn(rng, A_o.D, A_o.Domain) if vector else randn(rng, A_o.D, A_o.Domain, batch)
    y_ref = A_o(x, theta_o)
    with parops.use_engine(eng):
        Ap = A_p(theta_p) if A_p.P is parops.Parametric else A_p
        y = parops.apply_operator(Ap, eng.to_device(x), eng, flags)
        y = parops.cpu(y)
    assert y.shape == y_ref.shape, (y.shape, y_ref.shape)
    assert y.dtype == y_ref.dtype, (y.dtype, y_ref.dtype)
    if exact:
        assert np.array_equal(y, y_ref)
        return 0.0
    err = rel_l2(y, y_ref)
    t = TOL[np.dtype(A_o.R)] if tol is None else tol
    assert err <= t, "relative L2 %.3e > %.1e" % (err, t)
    return err


# ---- case builders ------------------------------------------------------------------------
def dft(T, n, adj=False):
    def b(ns):
        A = ns.ParDFT(T, n)
        return A.adjoint() if adj else A
    return b


def restriction(T, n, ranges, adj=False):
    def b(ns):
        A = ns.ParRestriction(T, n, ranges)
        return A.adjoint() if adj else A
    return b


def matrix(T, m, n, adj=False):
    def b(ns):
        A = ns.ParMatrix(T, m, n)
        return A.adjoint() if adj else A
    return b


def diagonal(T, n, adj=False):
    def b(ns):
        A = ns.ParDiagonal(T, n)
        return A.adjoint() if adj else A
    return b


def kron_of(*builders, adj=False):
    def b(ns):
        A = ns.kron(*[f(ns) for f in builders])
        return A.adjoint() if adj else A
    return b


def compose_of(*builders, adj=False):
    def b(ns):
        A = ns.compose(*[f(ns) for f in builders])
        return A.adjoint() if adj else A
    return b


def identity(T, n):
    return lambda ns: ns.ParIdentity(T, n)


def tensor(T, rank, ic, oc, modes):
    """ParTensor with the reference's weight layouts (src/ParTensor.jl:38-39 vs :63-64)."""
    def b(ns):
        if rank == 4:
            ws = (ic, oc, *modes)
        else:
            ws = (oc, ic, *modes)
        n = len(modes)
        return ns.ParTensor(T, tuple(range(1, 3 + n)), ws, tuple(range(2, 3 + n)), (ic, *modes), (1,) + tuple(range(3, 3 + n)), (oc, *modes))
    return b


def fno_spectral(T, shape, modes, c, adj=False):
    """examples/fno.jl:9-32 through both namespaces."""
    def b(ns):
        CT = O.complex_of(T)
        dft_time = ns.ParDFT(T, shape[-1])
        dft_space = ns.ParKron(*[ns.ParDFT(CT, s) for s in shape[:-1]])
        dft_spacetime = ns.kron(dft_time, dft_space)
        st = dft_spacetime.ops if hasattr(dft_spacetime, "ops") else [dft_spacetime]
        restrict = ns.ParKron(*[ns.ParRestriction(CT, Fo.Range, m) for Fo, m in zip(st, reversed(modes))])
        dft_restrict = ns.compose(restrict, dft_spacetime)
        dft_all = ns.kron(dft_restrict, ns.ParIdentity(T, c))
        mix = ns.ParMatrix(CT, c, c)
        ew = ns.ParDiagonal(CT, dft_restrict.Range)
        fw = ns.kron(ew, mix)
        S = ns.compose(dft_all.adjoint(), fw, dft_all)
        return S.adjoint() if adj else S
    return b


def fno_forward_transform(T, shape, modes, c, adj=False):
    def b(ns):
        CT = O.complex_of(T)
        dft_time = ns.ParDFT(T, shape[-1])
        dft_space = ns.ParKron(*[ns.ParDFT(CT, s) for s in shape[:-1]])
        dft_spacetime = ns.kron(dft_time, dft_space)
        st = dft_spacetime.ops if hasattr(dft_spacetime, "ops") else [dft_spacetime]
        restrict = ns.ParKron(*[ns.ParRestriction(CT, Fo.Range, m) for Fo, m in zip(st, reversed(modes))])
        dft_all = ns.kron(ns.compose(restrict, dft_spacetime), ns.ParIdentity(T, c))
        return dft_all.adjoint() if adj else dft_all
    return b


def check_pullback(eng, build, batch=2, seed=0, tol=None, flags=0):
    """rrule of the product (tape + po_plan_pullback) against the oracle's reverse mode
    (oracle.pullback: true adjoints, validated against finite differences in test_oracle.py)."""
    A_o, A_p = build(OracleNS), build(ProductNS)
    theta_o, theta_p = paired_params(A_o, A_p, eng)
    rng = np.random.default_rng(seed)
    x = randn(rng, A_o.D, A_o.Domain, batch)
    yb = randn(rng, A_o.R, A_o.Range, batch)
    xb_ref, g_ref = O.pullback(A_o, x, yb, theta_o)
    with parops.use_engine(eng):
        Ap = A_p(theta_p) if A_p.P is parops.Parametric else A_p
        y, back = parops.rrule(Ap, eng.to_device(x), eng, flags)
        grads, xb = back(eng.to_device(yb))
    t = TOL[np.dtype(A_o.R)] if tol is None else tol
    assert rel_l2(parops.cpu(y), A_o(x, theta_o)) <= t
    err = rel_l2(parops.cpu(xb), xb_ref)
    assert err <= t, "xbar relative L2 %.3e > %.1e" % (err, t)
    lo, lp = leaves_oracle(A_o), leaves_product(A_p)
    for a, b in zip(lo, lp):
        if a in g_ref:
            g = parops.cpu(grads[b]).reshape(np.asarray(g_ref[a]).shape, order="F")
            e = rel_l2(g, g_ref[a])
            assert e <= 10 * t, "grad of %s: relative L2 %.3e" % (type(a).__name__, e)
    return err
