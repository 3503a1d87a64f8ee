"""Pins for the oracle (oracle/parops_oracle.py).  The reference ships no golden vectors
("parity unpinned"), so the oracle is anchored on analytic known answers, on
the structural pins the reference's own tests hold, and on committed golden fixtures generated
by tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

from oracle import parops_oracle as O

F32, F64, C64, C128 = O.F32, O.F64, O.C64, O.C128
ALL = [F32, F64, C64, C128]
HERE = os.path.dirname(os.path.abspath(__file__))


def randn(rng, T, *shape):
    T = np.dtype(T)
    v = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if T.kind == "c" else 0)
    return np.asfortranarray(v.astype(T))


# ---- analytic DFT identities -------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 8, 12, 64])
def test_delta_and_single_mode(n):
    A = O.ParDFT(C128, n)
    d = np.zeros(n, dtype=C128); d[0] = 1
    assert np.allclose(A(d), np.ones(n))                       # delta -> constant
    k = 3 % n
    e = np.exp(2j * np.pi * k * np.arange(n) / n)
    y = A(e)
    want = np.zeros(n, dtype=C128); want[k] = n                # single exponential -> one-hot * n (unnormalised)
    assert np.allclose(y, want, atol=1e-10)
    assert np.allclose(A.adjoint()(y), e)                      # "adjoint" is the 1/n inverse (src/ParDFT.jl:30)


@pytest.mark.parametrize("n", [7, 8, 64, 253])
def test_real_dft_roundtrip_and_range(n):
    A = O.ParDFT(F64, n)
    assert A.Range == n // 2 + 1 and A.R == C128
    x = np.random.default_rng(0).standard_normal((n, 3))
    assert np.allclose(A.adjoint()(A(x)), x, atol=1e-12)
    # irfft ignores Im(DC) (and Im(Nyquist) for even n)
    y = A(x); y2 = y.copy(); y2[0] += 5j
    if n % 2 == 0:
        y2[-1] += 3j
    assert np.allclose(A.adjoint()(y2), x, atol=1e-12)


def test_dft_dot_test_has_the_1_over_n_factor():
    """test/serial.jl:41 asserts ratio 1, which contradicts src/ParDFT.jl:30 -- the code wins."""
    n = 64
    A = O.ParDFT(C128, n)
    rng = np.random.default_rng(1)
    x, y = randn(rng, C128, n), randn(rng, C128, n)
    u = np.vdot(x, A.adjoint()(y)).real
    v = np.vdot(y, A(x)).real   # note <y, Ax> vs <x, A'y>: A' = A^H / n
    assert np.isclose(u / np.vdot(A(x), y).real, 1.0 / n)
    assert v != 0


# ---- restriction -------------------------------------------------------------------------
@pytest.mark.parametrize("T", ALL)
def test_restriction_projector_and_sizes(T):
    A = O.ParRestriction(T, 8, [(1, 2), (7, 8)])
    assert A.Domain == 8 and A.Range == 4
    x = randn(np.random.default_rng(0), T, 8, 2)
    y = A(x)
    assert np.array_equal(y, x[[0, 1, 6, 7]])
    p = A.adjoint()(y)
    want = np.zeros_like(x); want[[0, 1, 6, 7]] = x[[0, 1, 6, 7]]
    assert np.array_equal(p, want)
    # overlapping ranges: the adjoint overwrites, later ranges win (src/ParRestriction.jl:31-37)
    B = O.ParRestriction(T, 6, [(1, 3), (2, 4)])
    yy = randn(np.random.default_rng(1), T, 6, 1)
    xx = B.adjoint()(yy)
    assert np.array_equal(xx[:, 0], np.array([yy[0, 0], yy[3, 0], yy[4, 0], yy[5, 0], 0, 0], dtype=T))


# ---- types and sizes (test/types_and_sizes.jl:35-57) ----------------------------------------
@pytest.mark.parametrize("T", ALL)
def test_types_and_sizes(T):
    ops = [O.ParIdentity(T, 4), O.ParDiagonal(T, 4), O.ParMatrix(T, 4, 4), O.ParMatrix(T, 4, 8), O.ParDFT(T, 4), O.ParDFT(T, 8),
           O.ParRestriction(T, 8, [(1, 2)]), O.ParRestriction(T, 8, [(1, 2), (7, 8)])]
    ops += [o.adjoint() for o in ops]
    for op in ops:
        th = O.init(op)
        rng = np.random.default_rng(0)
        xv, xm = randn(rng, op.D, op.Domain), randn(rng, op.D, op.Domain, 2)
        yv, ym = op(xv, th), op(xm, th)
        assert yv.shape == (op.Range,) and ym.shape == (op.Range, 2)
        assert yv.dtype == np.dtype(op.R) and ym.dtype == np.dtype(op.R)
    assert O.ParDFT(F32, 8).Range == 5


# ---- matrix / kron (test/serial.jl:35-36, 46-54) ----------------------------------------------
@pytest.mark.parametrize("T", ALL)
def test_matrix_dot_test(T):
    A = O.ParMatrix(T, 4, 5)
    th = O.init(A)
    rng = np.random.default_rng(0)
    x, y = randn(rng, T, 5, 10), randn(rng, T, 4, 10)
    u = np.vdot(x, A.adjoint()(y, th)).real
    v = np.vdot(y, A(x, th)).real
    assert np.isclose(u / v, 1.0, rtol=1e-3)


def test_kron_matches_dense_kronecker_product():
    A, B, Cm = O.ParMatrix(F64, 3, 4), O.ParMatrix(F64, 5, 2), O.ParMatrix(F64, 2, 3)
    K = O.kron(O.kron(A, B), Cm)
    th = O.init(K)
    dense = np.kron(np.kron(th[A], th[B]), th[Cm])
    x = np.random.default_rng(0).standard_normal((K.Domain, 2))
    assert np.allclose(K(x, th), dense @ x)
    assert np.allclose(K.adjoint()(K(x, th), th), dense.T @ (dense @ x))
    assert K.order == [3, 1, 2]  # max(Domain-Range) first, then right-most (src/ParKron.jl:47-55)


def test_kron_of_dfts_is_fftn_on_reversed_axes():
    K = O.ParKron(O.ParDFT(C128, 4), O.ParDFT(C128, 6), O.ParDFT(C128, 5))
    x = randn(np.random.default_rng(0), C128, 120, 2)
    ref = np.fft.fftn(x.reshape(5, 6, 4, 2, order="F"), axes=(0, 1, 2)).reshape(120, 2, order="F")
    assert np.allclose(K(x), ref)


def test_kron_order_heuristic():
    """src/ParKron.jl:33-57 on the cases SURVEY.md a8 lists."""
    Ft, Fx, Fy = O.ParDFT(F32, 32), O.ParDFT(C64, 64), O.ParDFT(C64, 64)
    assert O.ParKron(Ft, Fx, Fy).order == [1, 3, 2]
    assert O.ParKron(Ft, Fx, Fy, O.ParDFT(C64, 64)).order == [1, 4, 3, 2]
    I = O.ParIdentity(F32, 20)
    assert O.kron(O.ParCompose(O.ParRestriction(C64, 17, [(1, 8)]), Ft), I).order == [2, 1]
    assert O.kron(O.ParDiagonal(C64, 100), O.ParMatrix(C64, 20, 20)).order == [2, 1]
    M = [O.ParMatrix(F64, 256, 256) for _ in range(3)]
    assert O.ParKron(*M).order == [3, 2, 1]
    # truncation first: the factor that shrinks most goes first among equals
    R1 = O.ParCompose(O.ParRestriction(C64, 64, [(1, 8)]), O.ParDFT(C64, 64))
    R2 = O.ParCompose(O.ParRestriction(C64, 64, [(1, 16)]), O.ParDFT(C64, 64))
    assert O.ParKron(R2, R1).order == [2, 1] and O.ParKron(R1, R2).order == [1, 2]


# ---- fno.jl ---------------------------------------------------------------------------------
def test_fno_spectral_conv_equals_explicit_fft_pipeline():
    shape, modes, c = [8, 8, 6], [[(1, 2), (7, 8)], [(1, 2), (7, 8)], [(1, 2)]], 3
    S, parts = O.spectral_convolution(F64, shape, modes, c)
    th = O.init(S)
    x = np.random.default_rng(0).standard_normal((S.Domain, 2))
    y = S(x, th)
    # explicit: x is (c, y8, x8, t6, b)
    X = x.reshape(c, 8, 8, 6, 2, order="F")
    Xh = np.fft.rfft(X, axis=3)[:, :, :, :2]
    Xh = np.fft.fft(np.fft.fft(Xh, axis=1), axis=2)
    Xh = Xh[:, [0, 1, 6, 7]][:, :, [0, 1, 6, 7]]
    W, d = th[parts["mix_channels"]], th[parts["elementwise_weighting"]]
    Yh = np.einsum("oi,iabtn->oabtn", W, Xh) * d.reshape(4, 4, 2, order="F")[None, :, :, :, None]
    full = np.zeros((c, 8, 8, 4, 2), dtype=complex)
    full[np.ix_(range(c), [0, 1, 6, 7], [0, 1, 6, 7], [0, 1], [0, 1])] = Yh
    out = np.fft.irfft(np.fft.ifft(np.fft.ifft(full, axis=1), axis=2), n=6, axis=3)
    assert np.allclose(y, out.reshape(-1, 2, order="F"), atol=1e-12)
    assert parts["dft_all"].order == [2, 1]


# ---- repartition (test/distributed.jl:45-57) ----------------------------------------------------
@pytest.mark.parametrize("din,dout,gs", [([2, 2, 2], [8, 1, 1], (10, 10, 10)), ([1, 4], [4, 1], (6, 9)), ([2, 1, 3], [3, 2, 1], (5, 4, 7))])
def test_repartition_roundtrip_and_dot(din, dout, gs):
    R = O.ParRepartition(F32, din, dout, gs)
    n = int(np.prod(gs))
    xg = np.arange(n * 2, dtype=np.float32).reshape(n, 2, order="F")
    ys = R.apply_all(O.scatter_global(xg, din, gs))
    assert np.array_equal(O.gather_global(ys, dout, gs), xg)           # exact redistribution
    back = R.adjoint().apply_all(ys)
    assert np.array_equal(O.gather_global(back, din, gs), xg)          # adjoint swaps the plans
    if gs == (10, 10, 10):
        assert sum(len(p[2]) for p in R.plans) == 32 and [R.local_in(r)[0] for r in range(8)] == [5] * 8
        assert [R.local_out(r)[0] for r in range(8)] == [2, 2, 1, 1, 1, 1, 1, 1]


def test_distributed_kron_equals_serial():
    """distribute(K, dims) applied SPMD == serial K on the gathered array (src/ParKron.jl:244-309)."""
    T, CT = F64, C128
    def rf(Tin, n, rng):
        Fo = O.ParDFT(Tin, n)
        return O.ParCompose(O.ParRestriction(CT, Fo.Range, rng), Fo)
    K = O.ParKron(rf(T, 8, [(1, 3)]), rf(CT, 8, [(1, 2), (7, 8)]), rf(CT, 4, [(1, 2)]), O.ParIdentity(T, 3))
    assert K.order == [4, 1, 2, 3]
    gs_in = (3, 4, 8, 8)
    x = np.random.default_rng(0).standard_normal((K.Domain, 2))
    y = K(x)
    for dims in ([1, 1, 2, 1], [1, 2, 2, 1], [1, 1, 4, 1]):
        stages = O.distribute_kron(K, dims)
        xs = O.scatter_global(x, dims, gs_in)
        ys = O.apply_distributed(stages, xs)
        gs_out = (3, 2, 4, 3)
        assert np.allclose(O.gather_global(ys, dims, gs_out), y, atol=1e-12)
        assert any(k == "repartition" for k, _ in stages)


# ---- pullbacks vs finite differences ----------------------------------------------------------------
def test_pullback_matches_finite_differences():
    S, parts = O.spectral_convolution(F64, [8, 8, 6], [[(1, 2), (7, 8)], [(1, 2), (7, 8)], [(1, 2)]], 3)
    th = O.init(S)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((S.Domain, 2))
    yb = rng.standard_normal((S.Range, 2))
    xb, g = O.pullback(S, x, yb, th)
    loss = lambda xx, tt: np.sum(S(xx, tt) * yb)
    eps = 1e-6
    dx = rng.standard_normal(x.shape)
    assert np.isclose((loss(x + eps * dx, th) - loss(x - eps * dx, th)) / (2 * eps), np.sum(xb * dx), rtol=1e-6)
    for key in ("mix_channels", "elementwise_weighting"):
        P = parts[key]
        dP = rng.standard_normal(th[P].shape) + 1j * rng.standard_normal(th[P].shape)
        tp, tm = dict(th), dict(th)
        tp[P], tm[P] = th[P] + eps * dP, th[P] - eps * dP
        fd = (loss(x, tp) - loss(x, tm)) / (2 * eps)
        assert np.isclose(fd, np.sum((np.conj(g[P]) * dP).real), rtol=1e-6)


# ---- golden fixtures ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fno2d_f32", "fno2d_f64", "kron_c2_f32"])
def test_golden_fixtures(name):
    from tests.golden.make_golden import CASES, run_case
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    y = run_case(CASES[name], g)
    assert np.array_equal(y, g["y"]) or np.allclose(y, g["y"], rtol=1e-6, atol=0)
