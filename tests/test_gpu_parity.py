"""GPU parity suite (-m gpu): the sm_100a library through the C ABI against the oracle on the
same seeded inputs, plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import parops
from tests import parity_cases as pc
from tests import test_emul_parity as E
from tests.parity_cases import C64, C128, F32, F64

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T", pc.ALL_TYPES)
@pytest.mark.parametrize("n", [8, 12, 1])
def test_dft_forward_inverse(cuda_engine, T, n):
    E.test_dft_forward_inverse(cuda_engine, T, n)


@pytest.mark.parametrize("T", pc.ALL_TYPES)
@pytest.mark.parametrize("n,batch", [(64, 5), (128, 33), (1024, 3), (2048, 2), (253, 4), (640, 2)])
def test_dft_sizes(cuda_engine, T, n, batch):
    """radix-2 shared-memory FFT sizes, the dense fallback (253 = 11*23, 640) of test/serial.jl:56-68"""
    pc.check_case(cuda_engine, pc.dft(T, n), batch=batch)
    pc.check_case(cuda_engine, pc.dft(T, n, adj=True), batch=batch)


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_restriction_bit_exact(cuda_engine, T):
    E.test_restriction_bit_exact(cuda_engine, T)
    pc.check_case(cuda_engine, pc.restriction(T, 64, [(1, 4), (61, 64)]), batch=10, exact=True)
    pc.check_case(cuda_engine, pc.restriction(T, 64, [(1, 4), (61, 64)], adj=True), batch=10, exact=True)


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_matrix_diagonal(cuda_engine, T):
    E.test_matrix_diagonal(cuda_engine, T)
    pc.check_case(cuda_engine, pc.matrix(T, 70, 130), batch=200)
    pc.check_case(cuda_engine, pc.matrix(T, 70, 130, adj=True), batch=200)


def test_kron_of_matrices_both_orders(cuda_engine):
    E.test_kron_of_matrices_both_orders(cuda_engine)
    A, B, Cm = pc.matrix(F64, 10, 20), pc.matrix(F64, 30, 50), pc.matrix(F64, 60, 90)  # test/serial.jl:46-54 sizes
    pc.check_case(cuda_engine, pc.kron_of(Cm, B, A), batch=2)
    pc.check_case(cuda_engine, pc.kron_of(A, B, Cm), batch=2)


@pytest.mark.parametrize("T", [F32, F64])
def test_kron_of_dfts(cuda_engine, T):
    E.test_kron_of_dfts(cuda_engine, T)
    CT = pc.O.complex_of(T)
    pc.check_case(cuda_engine, pc.kron_of(pc.dft(T, 32), pc.dft(CT, 64), pc.dft(CT, 16)), batch=3)


@pytest.mark.parametrize("T", [F32, F64])
def test_truncated_dft_kron(cuda_engine, T):
    E.test_truncated_dft_kron(cuda_engine, T)
    # config C2 at 1/8 linear size per dim... (R[1:16] ∘ rDFT_64) ⊗ DFT_32 ⊗ DFT_32, batch 4
    CT = pc.O.complex_of(T)
    RF = pc.compose_of(pc.restriction(CT, 33, [(1, 16)]), pc.dft(T, 64))
    pc.check_case(cuda_engine, pc.kron_of(RF, pc.dft(CT, 32), pc.dft(CT, 32)), batch=4)
    pc.check_case(cuda_engine, pc.kron_of(RF, pc.dft(CT, 32), pc.dft(CT, 32), adj=True), batch=4)


@pytest.mark.parametrize("T", [F32, F64])
@pytest.mark.parametrize("flags", [0, 1])
def test_fno_spectral_conv_2d(cuda_engine, T, flags):
    E.test_fno_spectral_conv_2d(cuda_engine, T, flags)


def test_fno_c1_config(cuda_engine):
    """BASELINE.json configs[0]: 2-D layer 64x64, 20 channels, 12 modes, Float32."""
    shape, modes, c = [64, 64], [[(1, 12), (53, 64)], [(1, 12)]], 20
    for b in (1, 16):
        pc.check_case(cuda_engine, pc.fno_spectral(F32, shape, modes, c), batch=b)
        pc.check_case(cuda_engine, pc.fno_spectral(F32, shape, modes, c, adj=True), batch=b)


def test_fno_4d_small(cuda_engine):
    """config C3's operator at 16x16x16x8 (the oracle finishes in seconds)."""
    shape = [16, 16, 16, 8]
    modes = [[(1, 4), (13, 16)]] * 3 + [[(1, 4)]]
    pc.check_case(cuda_engine, pc.fno_spectral(F32, shape, modes, 20), batch=1)
    pc.check_case(cuda_engine, pc.fno_spectral(F32, shape, modes, 20, adj=True), batch=2)
    pc.check_case(cuda_engine, pc.fno_spectral(F64, shape, modes, 5), batch=1)


def test_fno_3d_unequal(cuda_engine):
    E.test_fno_spectral_conv_3d(cuda_engine)


@pytest.mark.parametrize("rank,modes", [(4, (3, 4)), (5, (2, 3, 2)), (6, (2, 2, 3, 2))])
def test_par_tensor(cuda_engine, rank, modes):
    E.test_par_tensor(cuda_engine, rank, modes)


def test_fused_plan_is_short(cuda_engine):
    E.test_fused_plan_is_short(cuda_engine)


@pytest.mark.parametrize("shape,c,batch", [([4, 64, 16], 3, 2), ([2, 2, 64, 32], 2, 1), ([2, 128, 16], 1, 1), ([8, 8, 64, 32], 20, 1),
                                           ([4, 4, 128, 64], 20, 1), ([4, 4, 64, 64], 20, 2), ([2, 128, 16], 20, 2), ([16, 16, 128, 64], 20, 1),
                                           # >= 4 x 148 units: the persistent kernels the bench runs (row ring + pipelined inverse; single-tile wide rows;
                                           # generic channel count) against the oracle
                                           ([32, 32, 64, 16], 20, 1), ([32, 32, 128, 16], 20, 1), ([16, 40, 64, 32], 6, 1)])
def test_fused_tx_fast_path(cuda_engine, shape, c, batch):
    E.test_fused_tx_fast_path(cuda_engine, shape, c, batch)


@pytest.mark.parametrize("shape,c,batch", [([32, 32, 16], 8, 1), ([32, 32, 16], 20, 2), ([32, 64, 32, 16], 2, 1), ([64, 32, 128, 16], 20, 1),
                                           ([32, 64, 64, 32], 6, 2)])
def test_zmz_fused_step(cuda_engine, shape, c, batch):
    E.test_zmz_fused_step(cuda_engine, shape, c, batch)


@pytest.mark.parametrize("n", [32, 64, 128])
def test_pruned_real16_kron(cuda_engine, n):
    E.test_pruned_real16_kron(cuda_engine, n)


@pytest.mark.parametrize("n", [32, 64, 128])
def test_c2_fused_pair(cuda_engine, n, capfd):
    E.test_c2_fused_pair(cuda_engine, n, capfd)


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_c2_fused_pair_variants(cuda_engine, variant):
    E.test_c2_fused_pair_compile_time_pitch(cuda_engine, variant)


def test_full_fft_one_pass(cuda_engine):
    E.test_full_fft_one_pass(cuda_engine)


def test_empty_operands(cuda_engine):
    E.test_empty_operands(cuda_engine)


@pytest.mark.parametrize("T,c,shape,batch", [(F32, 20, [8, 8, 4], 2), (F32, 4, [16, 16, 8], 1), (F32, 6, [8, 8, 4], 1), (F64, 5, [8, 8, 4], 2),
                                             (F32, 20, [64, 32, 16], 3)])
def test_fno_block_fused_epilogue(cuda_engine, T, c, shape, batch):
    E.test_fno_block_fused_epilogue(cuda_engine, T, c, shape, batch)


@pytest.mark.parametrize("c_in,c_out,npts,bias,act", [(8, 6, 1000, True, True), (20, 12, 4097, False, True), (32, 32, 640, True, False),
                                                     (4, 20, 129, False, True), (24, 3, 300, True, True)])
def test_block_epilogue_rectangular(cuda_engine, c_in, c_out, npts, bias, act):
    """po_block_epilogue on tensor cores (tcgen05 path: c_in % 4 == 0, <= 32 channels, >= 128 points) with rectangular W,
    ragged last tile, vector and scalar epilogues, against float64 NumPy."""
    import torch
    rng = np.random.default_rng(11)
    W = rng.standard_normal((c_out, c_in)).astype(np.float32)
    x = rng.standard_normal((c_in * npts, 1)).astype(np.float32)
    s = rng.standard_normal((c_out * npts, 1)).astype(np.float32)
    b = rng.standard_normal(c_out).astype(np.float32) if bias else None
    eng = cuda_engine
    with parops.use_engine(eng):
        y = parops.cpu(parops.block_epilogue(eng.to_device(np.asfortranarray(W)), eng.to_device(x), eng.to_device(s),
                                             eng.to_device(b) if bias else None, gelu=act))
    z = s.reshape(npts, c_out).astype(np.float64) + x.reshape(npts, c_in).astype(np.float64) @ W.astype(np.float64).T
    if bias:
        z = z + b.astype(np.float64)
    if act:
        z = 0.5 * z * (1 + np.tanh(np.sqrt(2 / np.pi) * (z + 0.044715 * z ** 3)))
    assert y.shape == (c_out * npts, 1) and y.dtype == np.float32
    assert pc.rel_l2(y.reshape(npts, c_out), z) <= 1e-5


def test_host_buffer_path(cuda_engine):
    """`A * x` with a NumPy x goes through po_plan_apply_host (H2D + plan + D2H)."""
    A_o = pc.fno_spectral(F32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)(pc.OracleNS)
    A_p = pc.fno_spectral(F32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)(pc.ProductNS)
    th_o, th_p = pc.paired_params(A_o, A_p, cuda_engine)
    x = pc.randn(np.random.default_rng(5), F32, A_o.Domain, 2)
    y = A_p(th_p) * x
    assert isinstance(y, np.ndarray)
    assert pc.rel_l2(y, A_o(x, th_o)) <= 1e-5


def test_host_buffer_async_pipeline(cuda_engine):
    """po_plan_apply_host_async: page-locked operands, three batches enqueued back to back on a side stream (upload, apply and
    download of successive batches are stream-ordered; the plan owns the device copies), one wait at the end."""
    import torch
    E.test_host_apply_async_entry(cuda_engine)
    A_o = pc.fno_spectral(F32, [16, 16, 8], [[(1, 4), (13, 16)], [(1, 4), (13, 16)], [(1, 4)]], 4)(pc.OracleNS)
    A_p = pc.fno_spectral(F32, [16, 16, 8], [[(1, 4), (13, 16)], [(1, 4), (13, 16)], [(1, 4)]], 4)(pc.ProductNS)
    th_o, th_p = pc.paired_params(A_o, A_p, cuda_engine)
    At = A_p(th_p)
    rng = np.random.default_rng(6)
    xs, outs, keep = [], [], []
    for _ in range(3):
        xp = torch.empty(A_o.Domain, dtype=torch.float32).pin_memory()
        yp = torch.zeros(A_o.Range, dtype=torch.float32).pin_memory()
        xp.copy_(torch.from_numpy(pc.randn(rng, F32, A_o.Domain, 1).reshape(-1)))
        keep += [xp, yp]
        xs.append(xp.numpy().reshape(-1, 1, order="F"))
        outs.append(yp.numpy().reshape(-1, 1, order="F"))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for x, o in zip(xs, outs):
            assert parops.apply_operator(At, x, out=o, wait=False) is o
    side.synchronize()
    cuda_engine.raise_if_failed()
    for x, o in zip(xs, outs):
        assert pc.rel_l2(o, A_o(x, th_o)) <= 1e-5


# ---- reverse mode ------------------------------------------------------------------------------
from tests import test_emul_pullback as EP  # noqa: E402


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_leaf_pullbacks(cuda_engine, T):
    EP.test_leaf_pullbacks(cuda_engine, T)


def test_kron_and_truncated_pullbacks(cuda_engine):
    EP.test_kron_and_truncated_pullbacks(cuda_engine)


@pytest.mark.parametrize("T", [F32, F64])
@pytest.mark.parametrize("flags", [0, 1])
def test_fno_pullback(cuda_engine, T, flags):
    EP.test_fno_pullback(cuda_engine, T, flags)


@pytest.mark.parametrize("shape,c,batch", [([4, 64, 16], 3, 2), ([2, 2, 64, 32], 2, 1), ([32, 64, 16], 2, 1), ([8, 8, 64, 32], 20, 1),
                                           ([32, 32, 64, 16], 20, 1), ([32, 32, 128, 16], 20, 1)])  # persistent kernels (>= 592 units)
def test_fast_path_pullback(cuda_engine, shape, c, batch):
    EP.test_fast_path_pullback(cuda_engine, shape, c, batch)


@pytest.mark.parametrize("shape,c,batch", [([32, 32, 16], 8, 1), ([32, 32, 16], 20, 2), ([32, 64, 32, 16], 4, 1), ([32, 16, 64, 32], 20, 1)])
def test_zmz_pullback(cuda_engine, shape, c, batch):
    EP.test_zmz_pullback(cuda_engine, shape, c, batch)


@pytest.mark.parametrize("n", [32, 128])
def test_pruned_real16_pullback(cuda_engine, n):
    EP.test_pruned_real16_pullback(cuda_engine, n)


def test_full_fft_one_pass_pullback(cuda_engine):
    EP.test_full_fft_one_pass_pullback(cuda_engine)


@pytest.mark.parametrize("rank,modes", [(4, (3, 4)), (5, (2, 3, 2)), (6, (2, 2, 3, 2))])
def test_par_tensor_pullback(cuda_engine, rank, modes):
    EP.test_par_tensor_pullback(cuda_engine, rank, modes)


def test_pullback_dot_test_full_size(cuda_engine):
    """C3 size: <S x, ybar> == <x, xbar> (true adjoint), and the gradient of the channel-mixing
    matrix equals the directional derivative along a random perturbation."""
    import torch
    shape = [64, 64, 64, 32]
    modes = [[(1, 8), (57, 64)]] * 3 + [[(1, 8)]]
    S = parops.spectral_convolution(parops.Float32, shape, modes, 20)
    theta = parops.gpu(parops.init(S, rng=2))
    St = S(theta)
    x = _device_randn(cuda_engine, S.Domain, 1, 0, torch.float32)
    yb = _device_randn(cuda_engine, S.Range, 1, 1, torch.float32)
    y, back = parops.rrule(St, x)
    grads, xb = back(yb)
    lhs = float((y.double() * yb.double()).sum())
    rhs = float((x.double() * xb.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs))
    M = [k for k in theta if isinstance(k, parops.ParMatrix)][0]
    dW = torch.randn_like(theta[M])
    eps = 1e-2
    tp, tm = dict(theta), dict(theta)
    tp[M], tm[M] = theta[M] + eps * dW, theta[M] - eps * dW
    fd = float((((S(tp) * x).double() - (S(tm) * x).double()) * yb.double()).sum()) / (2 * eps)
    an = float((grads[M].conj().to(torch.complex128) * dW.to(torch.complex128)).sum().real)
    assert abs(fd - an) <= 2e-3 * max(abs(fd), abs(an)), (fd, an)


# ---- full-size properties (BASELINE.json configs[1] and configs[2]) ---------------------------
def _device_randn(eng, n, b, seed, dtype):
    import torch
    g = torch.Generator(device=eng.torch_device).manual_seed(seed)
    return torch.randn((b, n), generator=g, device=eng.torch_device, dtype=dtype).t()


def test_c2_full_size_properties(cuda_engine):
    """C2: (R[1:16]∘rDFT_128) ⊗ DFT_128 ⊗ DFT_128 on 128^3 x batch 16, Float32:
    fused plan == reference-order plan; linearity; forward∘adjoint is the projector."""
    import torch
    T, CT = parops.Float32, parops.ComplexF32
    F1 = parops.ParDFT(T, 128)
    K = parops.ParKron(parops.compose(parops.ParRestriction(CT, 65, [(1, 16)]), F1), parops.ParDFT(CT, 128), parops.ParDFT(CT, 128))
    assert K.order == [1, 3, 2]
    x = _device_randn(cuda_engine, K.Domain, 16, 0, torch.float32)
    y = K * x
    y_ref = parops.apply_operator(K, x, cuda_engine, parops._lib.PO_PLAN_NO_FUSION)
    assert float((y - y_ref).norm() / y_ref.norm()) <= 1e-5
    x2 = _device_randn(cuda_engine, K.Domain, 16, 1, torch.float32)
    lin = K * (x + 2 * x2) - (y + 2 * (K * x2))
    assert float(lin.norm() / y.norm()) <= 1e-5
    # K K' = identity on the kept modes whose spectrum is consistent: use y itself
    y_back = K * (K.adjoint() * y)
    # irfft drops Im of the DC bin of the real dim: compare all other bins
    yv = y.t().reshape(16, 16, 128, 128)
    yb = y_back.t().reshape(16, 16, 128, 128)
    assert float((yv[:, 1:] - yb[:, 1:]).norm() / yv[:, 1:].norm()) <= 2e-5


def test_c3_full_size_properties(cuda_engine):
    """C3: 4-D spectral conv 64x64x64x32, c=20, 8 modes: fused == reference-order plan on a
    delta + noise input; adjoint dot-test with the reference's inverse-as-adjoint 1/n factor."""
    import torch
    shape = [64, 64, 64, 32]
    modes = [[(1, 8), (57, 64)]] * 3 + [[(1, 8)]]
    S = parops.spectral_convolution(parops.Float32, shape, modes, 20)
    theta = parops.gpu(parops.init(S, rng=2))
    St = S(theta)
    x = _device_randn(cuda_engine, S.Domain, 1, 0, torch.float32)
    y = St * x
    assert y.shape == (S.Range, 1) and y.dtype == torch.float32
    y_nf = parops.apply_operator(St, x, cuda_engine, parops._lib.PO_PLAN_NO_FUSION)
    assert float((y - y_nf).norm() / y_nf.norm()) <= 1e-5
    del y_nf
    # forward transform T = dft_all and its reference "adjoint" (inverse): T T' z = z on kept modes
    Tf = St.ops[-1]
    z = Tf * x
    z2 = Tf * (Tf.adjoint() * z)
    zv, z2v = z.reshape(-1), z2.reshape(-1)
    # skip the DC bin of the real (time) dim whose imaginary part irfft discards
    kt = torch.arange(zv.numel(), device=zv.device) // (20 * 16 * 16 * 16)
    keep = kt != 0
    assert float((zv[keep] - z2v[keep]).norm() / zv[keep].norm()) <= 2e-5


@pytest.mark.parametrize("shape", [[64, 64, 64, 64], [16, 128, 128, 64]])
def test_c4_local_shapes_fused_equals_reference_order(cuda_engine, shape):
    """The C4 ladder's per-GPU tensors (1.34 GB): 64^4 (row ring with nt = 64) and the 8-GPU slab 128 x 128 x 16 x 64 (x fastest;
    single-tile kernels for 10 KB rows).  Size-independent property: the fused plan equals the plan that keeps one kernel per
    reference leaf in reference order (PO_PLAN_NO_FUSION), for the layer and for its reference adjoint."""
    import torch
    modes = [[(1, 8), (s - 7, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    S = parops.spectral_convolution(parops.Float32, shape, modes, 20)
    theta = parops.gpu(parops.init(S, rng=2))
    St = S(theta)
    x = _device_randn(cuda_engine, S.Domain, 1, 3, torch.float32)
    for op in (St, St.adjoint()):
        y = op * x
        y_nf = parops.apply_operator(op, x, cuda_engine, parops._lib.PO_PLAN_NO_FUSION)
        assert float((y - y_nf).norm() / y_nf.norm()) <= 1e-5
        del y, y_nf


# ---- full-size BASELINE.json configs against the ORACLE (not against the library's own unfused plan) ------------
# The indexing paths that only trigger at full size (32-bit offset guards, units >= 4 x SMs, nit > 1 pipeline
# wrap of the persistent kernels) are checked here against the independent NumPy/pocketfft answer.  The oracle
# needs ~10-60 s per case on the GPU box's host cores; the relative L2 errors are printed (pytest -s / -rP).
C3_SHAPE = [64, 64, 64, 32]
C3_MODES = [[(1, 8), (57, 64)]] * 3 + [[(1, 8)]]


def test_c3_full_size_vs_oracle_forward(cuda_engine):
    """BASELINE.json configs[2] forward apply S(θ) * x at 64x64x64x32, c = 20, Float32, rel-L2 <= 1e-5."""
    err = pc.check_case(cuda_engine, pc.fno_spectral(F32, C3_SHAPE, C3_MODES, 20), batch=1)
    print("C3 forward vs oracle: rel-L2 %.3e" % err)


def test_c3_full_size_vs_oracle_adjoint(cuda_engine):
    """configs[2] adjoint apply S(θ)' * y (the reference's inverse-as-adjoint convention), rel-L2 <= 1e-5."""
    err = pc.check_case(cuda_engine, pc.fno_spectral(F32, C3_SHAPE, C3_MODES, 20, adj=True), batch=1, seed=1)
    print("C3 adjoint vs oracle: rel-L2 %.3e" % err)


def test_c3_full_size_vs_oracle_pullback(cuda_engine):
    """configs[2] reverse mode: y, x̄ (<= 1e-5) and the parameter gradients (<= 1e-4) against oracle.pullback."""
    err = pc.check_pullback(cuda_engine, pc.fno_spectral(F32, C3_SHAPE, C3_MODES, 20), batch=1)
    print("C3 pullback (xbar) vs oracle: rel-L2 %.3e" % err)


def _c2(adj=False):
    return pc.kron_of(pc.compose_of(pc.restriction(C64, 65, [(1, 16)]), pc.dft(F32, 128)), pc.dft(C64, 128), pc.dft(C64, 128), adj=adj)


def test_c2_full_size_vs_oracle(cuda_engine):
    """BASELINE.json configs[1]: (R[1:16]∘rDFT_128) ⊗ DFT_128 ⊗ DFT_128 on 128^3 x batch 16, forward and adjoint."""
    e1 = pc.check_case(cuda_engine, _c2(), batch=16)
    e2 = pc.check_case(cuda_engine, _c2(adj=True), batch=16, seed=1)
    print("C2 vs oracle: forward rel-L2 %.3e, adjoint %.3e" % (e1, e2))


def _c5(adj=False, diag_first=False):
    K = pc.kron_of(pc.matrix(F64, 256, 256), pc.matrix(F64, 256, 256), pc.matrix(F64, 256, 256))
    D = pc.diagonal(F64, 256 ** 3)
    return pc.compose_of(K, D, adj=adj) if diag_first else pc.compose_of(D, K, adj=adj)


@pytest.mark.parametrize("diag_first", [False, True])
def test_c5_full_size_vs_oracle(cuda_engine, diag_first):
    """BASELINE.json configs[4]: ParDiagonal ∘ (M1 ⊗ M2 ⊗ M3), 256x256 Float64 factors on a 256^3 vector, matvec and adjoint,
    rel-L2 <= 1e-12."""
    e1 = pc.check_case(cuda_engine, _c5(diag_first=diag_first), batch=1)
    e2 = pc.check_case(cuda_engine, _c5(adj=True, diag_first=diag_first), batch=1, seed=1)
    print("C5 vs oracle (diag %s): matvec rel-L2 %.3e, adjoint %.3e" % ("first" if diag_first else "last", e1, e2))


@pytest.mark.parametrize("rank,modes", [(6, (16, 16, 16, 8)), (4, (8, 4096)), (5, (16, 16, 128))])
def test_par_tensor_c3_size_vs_oracle(cuda_engine, rank, modes):
    """per-mode ParTensor weights, ic = oc = 20, M = 32768 modes (104.9 MB ComplexF32) -- the weight-streaming
    kernels (po_tensor.cuh) for all three reference layouts, forward and reverse mode, against the oracle."""
    e1 = pc.check_case(cuda_engine, pc.tensor(C64, rank, 20, 20, modes), batch=1)
    e2 = pc.check_pullback(cuda_engine, pc.tensor(C64, rank, 20, 20, modes), batch=2)
    print("ParTensor rank %d at C3 size vs oracle: apply rel-L2 %.3e, pullback xbar %.3e" % (rank, e1, e2))


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_par_tensor_stream_types(cuda_engine, T):
    pc.check_case(cuda_engine, pc.tensor(T, 5, 12, 10, (32, 32, 16)), batch=5)
    pc.check_pullback(cuda_engine, pc.tensor(T, 4, 6, 8, (128, 256)), batch=3)


@pytest.mark.parametrize("T,c_in,c_out,npts,bias,act", [(F32, 20, 20, 100000, True, True), (F64, 5, 7, 4099, True, True), (F32, 8, 12, 70001, False, True),
                                                        (F32, 20, 20, 4096, True, False)])
def test_block_epilogue_pullback(cuda_engine, T, c_in, c_out, npts, bias, act):
    """reverse mode of the block epilogue on the GPU (tcgen05 path for the recomputed pre-activation and for W' g)"""
    EP.test_block_epilogue_pullback(cuda_engine, T, c_in, c_out, npts, bias, act)


def test_fno_block_rrule(cuda_engine):
    EP.test_fno_block_rrule_matches_finite_differences(cuda_engine, F64, 4, [16, 16, 8])
    EP.test_fno_block_rrule_matches_finite_differences(cuda_engine, F32, 20, [32, 32, 16])
