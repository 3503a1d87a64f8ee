"""Reverse mode (po_plan_apply_taped + po_plan_pullback) on the CPU kernel emulator against
the oracle's pullback (true DFT adjoints, validated against finite differences)."""
import os

import numpy as np
import pytest

from tests import parity_cases as pc
from tests.parity_cases import C64, C128, F32, F64


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_leaf_pullbacks(emul_engine, T):
    for n in (8, 12):  # radix-2 FFT kernel / dense DFT-matrix kernel
        pc.check_pullback(emul_engine, pc.dft(T, n))
        pc.check_pullback(emul_engine, pc.dft(T, n, adj=True))
    pc.check_pullback(emul_engine, pc.matrix(T, 4, 5))
    pc.check_pullback(emul_engine, pc.matrix(T, 4, 5, adj=True))
    pc.check_pullback(emul_engine, pc.diagonal(T, 9))
    pc.check_pullback(emul_engine, pc.diagonal(T, 9, adj=True))
    pc.check_pullback(emul_engine, pc.restriction(T, 8, [(1, 2), (7, 8)]))
    pc.check_pullback(emul_engine, pc.restriction(T, 8, [(1, 2), (7, 8)], adj=True))


def test_kron_and_truncated_pullbacks(emul_engine):
    A, B = pc.matrix(F64, 3, 4), pc.matrix(F64, 5, 2)
    pc.check_pullback(emul_engine, pc.kron_of(A, B, pc.diagonal(F64, 3)))
    RF = pc.compose_of(pc.restriction(C128, 9, [(1, 4)]), pc.dft(F64, 16))
    pc.check_pullback(emul_engine, pc.kron_of(RF, pc.dft(C128, 8), pc.dft(C128, 8)))
    pc.check_pullback(emul_engine, pc.kron_of(RF, pc.dft(C128, 8), pc.dft(C128, 8), adj=True))
    RF2 = pc.compose_of(pc.restriction(C64, 33, [(1, 20)]), pc.dft(F32, 64))   # FFT kernel with gathering store
    pc.check_pullback(emul_engine, pc.kron_of(RF2, pc.identity(F32, 3)))
    pc.check_pullback(emul_engine, pc.kron_of(RF2, pc.identity(F32, 3), adj=True))


@pytest.mark.parametrize("T", [F32, F64])
@pytest.mark.parametrize("flags", [0, 1])
def test_fno_pullback(emul_engine, T, flags):
    shape, modes, c = [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3
    pc.check_pullback(emul_engine, pc.fno_spectral(T, shape, modes, c), flags=flags)
    pc.check_pullback(emul_engine, pc.fno_spectral(T, shape, modes, c, adj=True), flags=flags)


@pytest.mark.parametrize("shape,c,batch", [([4, 64, 16], 3, 2), ([2, 2, 64, 32], 2, 1), ([32, 64, 16], 2, 1), ([2, 128, 16], 20, 2),
                                           ([8, 64, 16], 4, 1), ([4, 64, 32], 20, 2),  # second-generation pipelined kernels
                                           ([8, 128, 16], 20, 1), ([8, 128, 32], 6, 1)])  # wide rows: CTA-pair forward / single-tile inverse
def test_fast_path_pullback(emul_engine, shape, c, batch):
    """fused (t,x) kernels, pruned c16 mode kernel and the fused mix+diag step in reverse mode"""
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    pc.check_pullback(emul_engine, pc.fno_forward_transform(F32, shape, modes, c), batch=batch)
    pc.check_pullback(emul_engine, pc.fno_forward_transform(F32, shape, modes, c, adj=True), batch=batch)


@pytest.mark.parametrize("shape,c,batch", [([4, 32, 64, 16], 2, 2)])
def test_chunked_inverse_pair_pullback(emul_engine, shape, c, batch):
    """the chunked y-inverse + fused inverse pair of the pullback (cotangents of F_y and of the fused forward)"""
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    os.environ["PO_INV_CHUNKS"] = "2"
    try:
        pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
        pc.check_pullback(emul_engine, pc.fno_forward_transform(F32, shape, modes, c), batch=batch)
    finally:
        os.environ.pop("PO_INV_CHUNKS", None)


@pytest.mark.parametrize("rank,modes", [(4, (3, 4)), (5, (2, 3, 2)), (6, (2, 2, 3, 2))])
def test_par_tensor_pullback(emul_engine, rank, modes):
    pc.check_pullback(emul_engine, pc.tensor(C64, rank, 3, 4, modes))
    pc.check_pullback(emul_engine, pc.tensor(F64, rank, 3, 4, modes))


@pytest.mark.parametrize("shape,c,batch", [([32, 32, 16], 8, 1), ([32, 32, 16], 20, 2)])
def test_zmz_pullback(emul_engine, shape, c, batch):
    """reverse mode of the fused z-mix-z step (x-bar, theta-bar, d-bar from the taped spectral input)"""
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c, adj=True), batch=batch)


@pytest.mark.parametrize("n", [32, 128])
def test_pruned_real16_pullback(emul_engine, n):
    RF = pc.compose_of(pc.restriction(C64, n // 2 + 1, [(1, 16)]), pc.dft(F32, n))
    pc.check_pullback(emul_engine, pc.kron_of(RF, pc.dft(C64, 4), pc.dft(C64, 8)))
    pc.check_pullback(emul_engine, pc.kron_of(RF, pc.dft(C64, 4), pc.dft(C64, 8), adj=True))


def test_full_fft_one_pass_pullback(emul_engine):
    pc.check_pullback(emul_engine, pc.kron_of(pc.dft(C64, 32), pc.dft(C64, 64)))
    pc.check_pullback(emul_engine, pc.kron_of(pc.dft(C64, 32), pc.dft(C64, 64), adj=True))


def _gelu_grad(z):
    c, a = np.sqrt(2.0 / np.pi), 0.044715
    u = c * (z + a * z ** 3)
    t = np.tanh(u)
    return 0.5 * (1 + t) + 0.5 * z * (1 - t * t) * c * (1 + 3 * a * z * z)


@pytest.mark.parametrize("T,c_in,c_out,npts,bias,act", [(F32, 20, 20, 300, True, True), (F64, 5, 7, 130, True, True), (F32, 8, 12, 257, False, True),
                                                        (F64, 6, 6, 64, True, False)])
def test_block_epilogue_pullback(emul_engine, T, c_in, c_out, npts, bias, act):
    """Cotangents of y = act.(s .+ (I ⊗ W) x .+ b) (po_block_epilogue_pullback) against the closed form in NumPy Float64."""
    import parops
    eng = emul_engine
    rng = np.random.default_rng(8)
    W, x, s = pc.randn(rng, T, c_out, c_in), pc.randn(rng, T, c_in * npts, 1), pc.randn(rng, T, c_out * npts, 1)
    b = pc.randn(rng, T, c_out) if bias else None
    yb = pc.randn(rng, T, c_out * npts, 1)
    X, S, Yb = (v.astype(np.float64).reshape(-1, npts, order="F") for v in (x, s, yb))
    z = S + W.astype(np.float64) @ X + (b.astype(np.float64)[:, None] if bias else 0.0)
    g = (_gelu_grad(z) if act else 1.0) * Yb
    with parops.use_engine(eng):
        dev = eng.to_device
        sbar, xbar, Wbar, bbar = parops.block_epilogue_pullback(dev(W), dev(x), dev(yb), s=dev(s), b=dev(b) if bias else None, gelu=act, need_b=bias)
    tol = 2e-5 if np.dtype(T) == np.float32 else 1e-12
    assert pc.rel_l2(parops.cpu(sbar).reshape(c_out, npts, order="F"), g) <= tol
    assert pc.rel_l2(parops.cpu(xbar).reshape(c_in, npts, order="F"), W.astype(np.float64).T @ g) <= tol
    assert pc.rel_l2(parops.cpu(Wbar), g @ X.T) <= 10 * tol
    if bias:
        assert pc.rel_l2(parops.cpu(bbar), g.sum(axis=1)) <= 10 * tol


@pytest.mark.parametrize("T,c,shape", [(F64, 4, [8, 8, 4]), (F32, 20, [8, 8, 4])])
def test_fno_block_rrule_matches_finite_differences(emul_engine, T, c, shape):
    """fno_block_rrule (spectral tape + block-epilogue pullback): <ybar, dy> for a random perturbation of x and of every parameter
    equals the inner product of the perturbation with the returned cotangents."""
    import parops
    eng = emul_engine
    modes = [[(1, 2), (s - 1, s)] for s in reversed(shape[:-1])] + [[(1, 3)]]
    JT = parops.jtype(np.dtype(T))
    rng = np.random.default_rng(9)
    with parops.use_engine(eng):
        S = parops.spectral_convolution(JT, shape, modes, c)
        Wc = parops.channel_mixing(JT, shape, c)
        th = parops.gpu(parops.init(S, Wc, rng=4))
        x = eng.to_device(pc.randn(rng, T, S.Domain, 1))
        yb = eng.to_device(pc.randn(rng, T, S.Range, 1))
        y, back = parops.fno_block_rrule(x, S(th), Wc(th))
        assert pc.rel_l2(parops.cpu(y), parops.cpu(parops.fno_block(x, S(th), Wc(th)))) <= 1e-6
        grads, xbar = back(yb)
        eps = 1e-3 if np.dtype(T) == np.float32 else 1e-6
        f = lambda xx, tt: float((parops.fno_block(xx, S(tt), Wc(tt)).double() * yb.double()).sum())  # noqa: E731
        dx = eng.to_device(pc.randn(rng, T, S.Domain, 1))
        fd = (f(x + eps * dx, th) - f(x - eps * dx, th)) / (2 * eps)
        an = float((xbar.double() * dx.double()).sum())
        rtol = 3e-2 if np.dtype(T) == np.float32 else 1e-6  # Float32: finite-difference round-off; the Float64 case pins the math
        assert abs(fd - an) <= rtol * max(abs(fd), abs(an), 1e-3), ("x", fd, an)
        for leaf, g in grads.items():
            d = eng.to_device(pc.randn(rng, np.dtype(leaf.D.np), *th[leaf].shape))
            tp, tm = dict(th), dict(th)
            tp[leaf], tm[leaf] = th[leaf] + eps * d, th[leaf] - eps * d
            fd = (f(x, tp) - f(x, tm)) / (2 * eps)
            an = float((g.conj() * d).sum().real) if g.is_complex() else float((g.double() * d.double()).sum())
            assert abs(fd - an) <= rtol * max(abs(fd), abs(an), 1e-3), (type(leaf).__name__, fd, an)
