"""Kernel-logic parity on the CPU emulator build of csrc/ (no GPU needed).

These run the REAL plan compiler and the REAL kernel sources (compiled for CPU threads by
tests/emul) against the oracle; the GPU suite (test_gpu_parity.py) repeats the same cases on
the sm_100a library."""
import numpy as np
import os

import pytest

from tests import parity_cases as pc
from tests.parity_cases import C64, C128, F32, F64


@pytest.mark.parametrize("T", pc.ALL_TYPES)
@pytest.mark.parametrize("n", [8, 12, 1])
def test_dft_forward_inverse(emul_engine, T, n):
    pc.check_case(emul_engine, pc.dft(T, n))
    pc.check_case(emul_engine, pc.dft(T, n, adj=True))
    pc.check_case(emul_engine, pc.dft(T, n), vector=True)


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_restriction_bit_exact(emul_engine, T):
    for ranges in ([(1, 2)], [(1, 2), (7, 8)], [(3, 5), (4, 6)]):
        pc.check_case(emul_engine, pc.restriction(T, 8, ranges), exact=True)
        pc.check_case(emul_engine, pc.restriction(T, 8, ranges, adj=True), exact=True)


@pytest.mark.parametrize("T", pc.ALL_TYPES)
def test_matrix_diagonal(emul_engine, T):
    pc.check_case(emul_engine, pc.matrix(T, 4, 5))
    pc.check_case(emul_engine, pc.matrix(T, 4, 5, adj=True))
    pc.check_case(emul_engine, pc.matrix(T, 20, 20), batch=70)
    pc.check_case(emul_engine, pc.diagonal(T, 42))
    pc.check_case(emul_engine, pc.diagonal(T, 42, adj=True))


def test_kron_of_matrices_both_orders(emul_engine):
    """test/serial.jl:46-54"""
    A, B, Cm = pc.matrix(F64, 3, 4), pc.matrix(F64, 5, 6), pc.matrix(F64, 7, 2)
    pc.check_case(emul_engine, pc.kron_of(Cm, B, A), batch=2)
    pc.check_case(emul_engine, pc.kron_of(A, B, Cm), batch=2)
    pc.check_case(emul_engine, pc.kron_of(A, B, Cm, adj=True), batch=2)


@pytest.mark.parametrize("T", [F32, F64])
def test_kron_of_dfts(emul_engine, T):
    CT = pc.O.complex_of(T)
    K = pc.kron_of(pc.dft(T, 8), pc.dft(CT, 4), pc.dft(CT, 16))
    pc.check_case(emul_engine, K, batch=2)
    pc.check_case(emul_engine, pc.kron_of(pc.dft(T, 8), pc.dft(CT, 4), pc.dft(CT, 16), adj=True), batch=2)
    # non power of two + rectangular real DFT (test/serial.jl:56-68 pattern, small)
    pc.check_case(emul_engine, pc.kron_of(pc.dft(CT, 10), pc.dft(T, 7)), batch=2)


@pytest.mark.parametrize("T", [F32, F64])
def test_truncated_dft_kron(emul_engine, T):
    """config C2's shape family: (R∘rDFT) ⊗ DFT ⊗ DFT"""
    CT = pc.O.complex_of(T)
    RF = pc.compose_of(pc.restriction(CT, 9, [(1, 4)]), pc.dft(T, 16))
    pc.check_case(emul_engine, pc.kron_of(RF, pc.dft(CT, 8), pc.dft(CT, 8)), batch=2)
    pc.check_case(emul_engine, pc.kron_of(RF, pc.dft(CT, 8), pc.dft(CT, 8), adj=True), batch=2)


@pytest.mark.parametrize("T", [F32, F64])
@pytest.mark.parametrize("flags", [0, 1])
def test_fno_spectral_conv_2d(emul_engine, T, flags):
    """examples/fno.jl pattern (config C1 scaled down), fused and reference-order plans."""
    shape, modes, c = [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3
    pc.check_case(emul_engine, pc.fno_spectral(T, shape, modes, c), batch=2, flags=flags)
    pc.check_case(emul_engine, pc.fno_spectral(T, shape, modes, c, adj=True), batch=2, flags=flags)


def test_fno_spectral_conv_3d(emul_engine):
    # NB examples/fno.jl:17 zips the Kron factors with reverse(modes): spatial dim i gets
    # modes[end-i] -- a quirk that only shows with unequal spatial extents; mirrored faithfully.
    shape, modes, c = [8, 4, 8], [[(1, 2), (4, 4)], [(1, 2), (7, 8)], [(1, 3)]], 2
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=1)


@pytest.mark.parametrize("rank,modes", [(4, (3, 4)), (5, (2, 3, 2)), (6, (2, 2, 3, 2))])
def test_par_tensor(emul_engine, rank, modes):
    pc.check_case(emul_engine, pc.tensor(C64, rank, 3, 4, modes), batch=2)
    pc.check_case(emul_engine, pc.tensor(F64, rank, 3, 4, modes), batch=2)


@pytest.mark.parametrize("T", [C64, F64, F32, C128])
@pytest.mark.parametrize("rank,ic,oc,modes,batch", [(4, 4, 6, (20, 20), 1), (5, 6, 4, (10, 8, 5), 2), (6, 20, 20, (4, 3, 2, 2), 1), (4, 6, 6, (25, 16), 5),
                                                    (5, 3, 5, (16, 8, 5), 1)])
def test_par_tensor_stream(emul_engine, T, rank, ic, oc, modes, batch):
    """Sizes above the streaming threshold: the weight-streaming bulk-copy ring kernels (po_tensor.cuh) for both weight
    layouts (rank 4: (ic, oc, ...), rank 5/6: (oc, ic, ...)), ragged last chunk, several batch passes (batch 5 > BMAX = 4),
    odd channel counts (alignment fallbacks stay correct)."""
    pc.check_case(emul_engine, pc.tensor(T, rank, ic, oc, modes), batch=batch)


@pytest.mark.parametrize("rank,ic,oc,modes,batch", [(4, 4, 6, (20, 20), 1), (5, 6, 4, (10, 8, 5), 2), (6, 20, 20, (4, 3, 2, 2), 1), (4, 6, 6, (25, 16), 5)])
def test_par_tensor_stream_pullback(emul_engine, rank, ic, oc, modes, batch):
    pc.check_pullback(emul_engine, pc.tensor(C64, rank, ic, oc, modes), batch=batch)
    pc.check_pullback(emul_engine, pc.tensor(F64, rank, ic, oc, modes), batch=batch)


def test_fused_plan_is_short(emul_engine):
    """The plan compiler must fuse restriction into the DFTs: 2-D FNO forward transform =
    one truncated DFT per grid dim, no gather steps, no permutes."""
    import parops
    A = pc.fno_forward_transform(F32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)(pc.ProductNS)
    plan, _ = parops.get_plan(A, 2, emul_engine)
    text = plan.describe()
    assert plan.num_steps() == 2, text
    assert "gather" not in text, text
    plan_nf, _ = parops.get_plan(A, 2, emul_engine, parops._lib.PO_PLAN_NO_FUSION)
    assert plan_nf.num_steps() == 4


@pytest.mark.parametrize("shape,c,batch", [([4, 64, 16], 3, 2), ([2, 2, 64, 32], 2, 1), ([2, 128, 16], 1, 1), ([2, 64, 64], 5, 1),
                                           ([2, 128, 16], 20, 1), ([2, 128, 16], 20, 2),  # channel groups (IG = 10 of 20)
                                           # second-generation pipelined kernels (po_fused2.cuh): even channel counts, >= 8 units
                                           ([8, 64, 16], 4, 1), ([4, 64, 32], 20, 2), ([3, 3, 64, 64], 2, 1), ([8, 128, 16], 6, 1),
                                           ([8, 128, 16], 20, 1),  # rows of 20 * 128 floats: CTA-pair forward (po_fused4.cuh) / single-tile inverse
                                           ([8, 128, 32], 20, 1), ([4, 128, 64], 6, 2), ([2, 4, 128, 16], 2, 1)])  # CTA-pair: RT = 4, 8; generic channel counts; batch
def test_fused_tx_fast_path(emul_engine, shape, c, batch):
    """The shape-specialised fused (t, x) kernels (po_fused.cuh): forward transform, its
    reference "adjoint" (inverse) and the whole spectral convolution, against the oracle; the
    plan must actually contain the fused step."""
    import parops
    nsp = len(shape) - 1
    # examples/fno.jl zips factors with reverse(modes): emit so that dim i gets ITS extent's ranges
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    fwd = pc.fno_forward_transform(F32, shape, modes, c)
    A = fwd(pc.ProductNS)
    plan, _ = parops.get_plan(A, batch, emul_engine)
    assert "fused-fwd-tx" in plan.describe(), plan.describe()
    pc.check_case(emul_engine, fwd, batch=batch)
    inv = pc.fno_forward_transform(F32, shape, modes, c, adj=True)
    plan_i, _ = parops.get_plan(inv(pc.ProductNS), batch, emul_engine)
    assert "fused-inv-xt" in plan_i.describe(), plan_i.describe()
    pc.check_case(emul_engine, inv, batch=batch)
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)


@pytest.mark.parametrize("shape,c,batch", [([32, 32, 16], 8, 1), ([32, 32, 16], 20, 2), ([32, 64, 32, 16], 2, 1)])
def test_zmz_fused_step(emul_engine, shape, c, batch):
    """po_zmz.cuh: pruned z-DFT -> ParDiagonal (x) ParMatrix -> zero-padded inverse z-DFT in one kernel; the plan must
    contain the fused step and agree with the oracle and with the unfused plan."""
    import parops
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    S = pc.fno_spectral(F32, shape, modes, c)
    A_o, A_p = S(pc.OracleNS), S(pc.ProductNS)
    _, theta_p = pc.paired_params(A_o, A_p, emul_engine)
    with parops.use_engine(emul_engine):
        plan, _ = parops.get_plan(A_p(theta_p), batch, emul_engine)
    assert "z-mix-z" in plan.describe(), plan.describe()
    pc.check_case(emul_engine, S, batch=batch)
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c, adj=True), batch=batch)


@pytest.mark.parametrize("shape,c,batch,chunks", [([4, 32, 64, 16], 2, 1, "2"), ([4, 32, 64, 16], 4, 2, "4"), ([2, 64, 64, 32], 2, 1, "2"),
                                                  ([2, 32, 64, 64], 2, 1, None)])  # None: the default (on from nt = 64)
def test_chunked_inverse_pair(emul_engine, shape, c, batch, chunks, capfd):
    """y-inverse (pruned c16) + fused inverse (x, t) run as z-chunks so that the intermediate stays in L2
    (po_run_inv_pair in po_plan.cuh): same results as the oracle, and the chunked path must be the one that ran."""
    modes = [[(1, 8), (s - 7, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, 8)]]
    os.environ["PO_TRACE"] = "1"
    if chunks:
        os.environ["PO_INV_CHUNKS"] = chunks
    try:
        pc.check_case(emul_engine, pc.fno_forward_transform(F32, shape, modes, c, adj=True), batch=batch)
        err = capfd.readouterr().err
        assert "inverse pair in %s z-chunks" % (chunks or "2") in err, err
        pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
        pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c, adj=True), batch=batch)
    finally:
        os.environ.pop("PO_TRACE", None)
        os.environ.pop("PO_INV_CHUNKS", None)


@pytest.mark.parametrize("n", [32, 64, 128])
def test_pruned_real16_kron(emul_engine, n):
    """config C2's slow-dim factor R[1:16] o rfft_n through the pruned real-16 kernel (po_r16_mode_kernel) and its
    reference "adjoint" (zero-padded irfft)"""
    import parops
    RF = pc.compose_of(pc.restriction(C64, n // 2 + 1, [(1, 16)]), pc.dft(F32, n))
    K = pc.kron_of(RF, pc.dft(C64, 4), pc.dft(C64, 8))
    plan, _ = parops.get_plan(K(pc.ProductNS), 2, emul_engine)
    assert "pruned-r16 r2c" in plan.describe(), plan.describe()
    pc.check_case(emul_engine, K, batch=2)
    Ka = pc.kron_of(RF, pc.dft(C64, 4), pc.dft(C64, 8), adj=True)
    plan_a, _ = parops.get_plan(Ka(pc.ProductNS), 2, emul_engine)
    assert "pruned-r16 c2r" in plan_a.describe(), plan_a.describe()
    pc.check_case(emul_engine, Ka, batch=2)


def test_full_fft_one_pass(emul_engine):
    """po_cfull_mode_kernel: untruncated c2c FFT of length 32 / 64 / 128 along contiguous and strided modes"""
    import parops
    K = pc.kron_of(pc.dft(C64, 128), pc.dft(C64, 32), pc.dft(C64, 64))
    plan, _ = parops.get_plan(K(pc.ProductNS), 1, emul_engine)
    assert plan.describe().count("fft16x") == 3, plan.describe()
    pc.check_case(emul_engine, K, batch=1)
    pc.check_case(emul_engine, pc.kron_of(pc.dft(C64, 128), pc.dft(C64, 32), pc.dft(C64, 64), adj=True), batch=1)
    pc.check_case(emul_engine, pc.dft(C64, 128), batch=5)   # I == 1, ragged last CTA


@pytest.mark.parametrize("T,c,shape,batch", [(F32, 20, [8, 8, 4], 2), (F32, 4, [16, 16, 8], 1), (F32, 6, [8, 8, 4], 1), (F64, 5, [8, 8, 4], 2)])
def test_fno_block_fused_epilogue(emul_engine, T, c, shape, batch):
    """examples/fno.jl:40-44 `gelu.(S x .+ W x)` with the pointwise branch + add + gelu in ONE kernel
    (po_block_epilogue: fast float path for c % 4 == 0, generic path otherwise) against the oracle's fno_block."""
    import parops
    O = pc.O
    modes = [[(1, 2), (s - 1, s)] for s in reversed(shape[:-1])] + [[(1, 3)]]
    eng = emul_engine
    S_o, S_p = pc.fno_spectral(T, shape, modes, c)(pc.OracleNS), pc.fno_spectral(T, shape, modes, c)(pc.ProductNS)
    W_o, W_p = O.channel_mixing(T, shape, c), parops.channel_mixing(parops.jtype(np.dtype(T)), shape, c)
    th_o, th_p = pc.paired_params(S_o, S_p, eng)
    thw_o, thw_p = pc.paired_params(W_o, W_p, eng)
    rng = np.random.default_rng(3)
    x = pc.randn(rng, T, S_o.Domain, batch)
    ref = O.fno_block(x, S_o, W_o, {**th_o, **thw_o})
    with parops.use_engine(eng):
        y = parops.fno_block(eng.to_device(x), S_p(th_p), W_p(thw_p))
        y = parops.cpu(y)
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert pc.rel_l2(y, ref) <= (2e-5 if np.dtype(T) == np.float32 else 1e-12)
    # bias + no activation path of the C entry point
    Wm = list(thw_p.values())[0]
    b = eng.to_device(pc.randn(rng, T, c))
    with parops.use_engine(eng):
        z = parops.cpu(parops.block_epilogue(Wm, eng.to_device(x), None, b, gelu=False))
    Wn = list(thw_o.values())[0]
    zr = (Wn @ x.reshape(c, -1, order="F") + parops.cpu(b).reshape(c, 1)).reshape(x.shape, order="F")
    assert pc.rel_l2(z, zr) <= (1e-5 if np.dtype(T) == np.float32 else 1e-12)


def test_kron_apply_convenience(emul_engine):
    """po_kron_apply (one-call Kron of leaf factors, cached plan + workspace) against the oracle's ParKron."""
    import ctypes as C
    import parops
    from parops._lib import po_node
    eng = emul_engine
    A_o = pc.O.ParKron(pc.O.ParRestriction(C64, 6, [(1, 2), (5, 6)]), pc.O.ParDFT(C64, 8), pc.O.ParIdentity(C64, 3))
    A_p = parops.ParKron(parops.ParRestriction(parops.ComplexF32, 6, [(1, 2), (5, 6)]), parops.ParDFT(parops.ComplexF32, 8),
                         parops.ParIdentity(parops.ComplexF32, 3))
    assert list(A_o.order) == list(A_p.order)
    nodes = (po_node * 3)()
    aux = [6, 2, 1, 2, 5, 6, 8]
    c64 = parops.ComplexF32.code
    for k, (kind, dom, ran, ab, ac) in enumerate([(parops._lib.PO_NODE_RESTRICTION, 6, 4, 0, 6), (parops._lib.PO_NODE_DFT, 8, 8, 6, 1),
                                                  (parops._lib.PO_NODE_IDENTITY, 3, 3, 0, 0)]):
        nodes[k].kind, nodes[k].ddt, nodes[k].rdt, nodes[k].adjoint, nodes[k].param_slot = kind, c64, c64, 0, -1
        nodes[k].child_begin = nodes[k].child_count = 0
        nodes[k].aux_begin, nodes[k].aux_count, nodes[k].domain, nodes[k].range = ab, ac, dom, ran
    rng = np.random.default_rng(5)
    x = pc.randn(rng, C64, A_o.Domain, 3)
    xd = eng.to_device(x)
    y = eng.empty_colmajor(A_o.Range, 3, parops.ComplexF32)
    for _ in range(2):  # second call: cached plan
        eng.lib.check(eng.lib.po_kron_apply(eng.ctx, nodes, 3, (C.c_int64 * len(aux))(*aux), len(aux), (C.c_int32 * 3)(*A_p.order), 3,
                                            C.c_void_p(xd.data_ptr()), C.c_void_p(y.data_ptr()), None, 0, eng.stream_ptr()))
    assert pc.rel_l2(parops.cpu(y), A_o(x)) <= 1e-5


# ---- a20: distribute for ParMatrix / ParDiagonal / ParTensor (src/ParMatrix.jl:21, src/ParDiagonal.jl:62-65, src/ParTensor.jl:153-196)
@pytest.mark.parametrize("n,size", [(10, 4), (7, 3), (16, 8)])
def test_distribute_diagonal_shards_apply_like_the_serial_operator(emul_engine, n, size):
    import parops
    from parops.distributed import CartComm
    eng = emul_engine
    rng = np.random.default_rng(3)
    d = pc.randn(rng, C64, n)
    x = pc.randn(rng, C64, n, 2)
    want = d[:, None] * x
    off = 0
    with parops.use_engine(eng):
        for r in range(size):
            Dl = parops.distribute(parops.ParDiagonal(parops.ComplexF32, n), CartComm([size], r))
            nl = pc.O.local_size(n, r, size)
            assert isinstance(Dl, parops.ParDiagonal) and Dl.n == nl == parops.local_size(n, r, size)
            if nl:
                y = parops.cpu(Dl({Dl: eng.to_device(d[off:off + nl])}) * eng.to_device(np.asfortranarray(x[off:off + nl])))
                assert pc.rel_l2(y, want[off:off + nl]) <= 1e-6
            off += nl
    assert off == n


def test_distribute_matrix_is_broadcasted():
    import parops
    A = parops.ParMatrix(parops.Float32, 3, 4)
    B = parops.distribute(A)
    assert isinstance(B, parops.ParBroadcasted) and B.op is A and B.Domain == 4 and B.Range == 3


@pytest.mark.parametrize("dims,axis", [([1, 2, 1, 1], 1), ([1, 1, 1, 4], 3)])
def test_distribute_tensor_shards_apply_like_the_serial_operator(emul_engine, dims, axis):
    """ParTensor.distribute divides weight / input / target extents along the non-contracted dims (ids get the cart coords);
    each rank's local operator with its slab of the weights reproduces the serial result on its slab of the input."""
    import parops
    from parops.distributed import CartComm
    eng = emul_engine
    ic, oc, nx, ny, nt, b = 3, 4, 4, 2, 4, 2
    # einsum labels: i = 1, x = 2, y = 3, t = 4, o = 5  ->  weights (o, i, x, y, t), input (i, x, y, t), target (o, x, y, t)
    mk = lambda ns, T, shp_w, shp_i, shp_t, *idv: ns.ParTensor(T, (5, 1, 2, 3, 4), shp_w, (1, 2, 3, 4), shp_i, (5, 2, 3, 4), shp_t, *idv)  # noqa: E731
    A_o = mk(pc.OracleNS, C64, (oc, ic, nx, ny, nt), (ic, nx, ny, nt), (oc, nx, ny, nt))
    A_p = mk(pc.ProductNS, parops.ComplexF32, (oc, ic, nx, ny, nt), (ic, nx, ny, nt), (oc, nx, ny, nt), "mix")
    rng = np.random.default_rng(4)
    W = pc.randn(rng, C64, oc, ic, nx, ny, nt)
    x = pc.randn(rng, C64, ic, nx, ny, nt, b)
    want = A_o(np.asfortranarray(x.reshape(-1, b, order="F")), {A_o: W}).reshape((oc, nx, ny, nt, b), order="F")
    P = dims[axis]
    with pytest.raises(parops.ParException):
        parops.distribute(A_p, [2, 1, 1, 1], CartComm([2, 1, 1, 1], 0))       # the contracted dim cannot be distributed
    with parops.use_engine(eng):
        for r in range(P):
            Al = parops.distribute(A_p, dims, CartComm(dims, r))
            coords = CartComm(dims, r).coords()
            assert Al.id == "mix:(%s)" % ",".join(str(c) for c in coords)
            ext = [nx, ny, nt][axis - 1] // P
            sl = [slice(None)] * 5
            sl[axis] = slice(coords[axis] * ext, (coords[axis] + 1) * ext)
            slw = [slice(None)] + sl[:4]                                      # weights carry o in front of (i, x, y, t)
            assert Al.weight_shape == W[tuple(slw)].shape and Al.input_shape == x[tuple(sl)].shape[:4]
            xl = np.asfortranarray(x[tuple(sl)]).reshape(-1, b, order="F")
            y = parops.cpu(Al({Al: eng.to_device(np.asfortranarray(W[tuple(slw)]))}) * eng.to_device(np.asfortranarray(xl)))
            assert pc.rel_l2(y.reshape(want[tuple(sl)].shape, order="F"), want[tuple(sl)]) <= 1e-6


@pytest.mark.parametrize("T", [F32, C64, F64])
def test_matrix_batched_mul_3d_and_divide(emul_engine, T):
    """src/ParMatrix.jl:44 (3-D operand -> NNlib.batched_mul(θ, x), the 2-D weight broadcast over the batch dim) and :60 (A / x)."""
    import parops
    eng = emul_engine
    rng = np.random.default_rng(6)
    W = pc.randn(rng, T, 5, 7)
    x = pc.randn(rng, T, 7, 3, 4)
    want = np.einsum("mn,ncb->mcb", W, x)
    tol = 1e-5 if np.dtype(T).itemsize <= 8 and np.dtype(T) != np.dtype(F64) else 1e-12
    with parops.use_engine(eng):
        A = parops.ParMatrix(parops.jtype(T), 5, 7)
        At = A({A: eng.to_device(W)})
        y = parops.cpu(At * eng.to_device(x))
        assert y.shape == (5, 3, 4) and pc.rel_l2(y, want) <= tol
        yb = parops.cpu(parops.ParBroadcasted(A)({A: eng.to_device(W)}) * eng.to_device(x))   # src/ParBroadcasted.jl:35
        assert pc.rel_l2(yb, want) <= tol
        yh = At * np.asfortranarray(x)                                                         # host operand
        assert pc.rel_l2(yh, want) <= tol
        with pytest.raises(parops.ParException):
            At.adjoint() * eng.to_device(pc.randn(rng, T, 5, 3, 4))                            # no 3-D method for the adjoint
        d = pc.randn(rng, T, 5, 7) + 3
        q = parops.cpu(At / eng.to_device(d))
        assert pc.rel_l2(q, W / d) <= 1e-6


def test_large_fp64_factors_use_the_gemm_arm(emul_engine, monkeypatch=None):
    """BASELINE.json configs[4] at 1/8 size: Kron of three 128 x 128 Float64 ParMatrix factors (+ ParDiagonal) on a 128^3 vector: the
    128 x 128 x 8 prefetching GEMM (po_dgemm.cuh, SIMT arm) in both operand layouts (I == 1 and I > 1), forward and adjoint."""
    K = pc.kron_of(pc.matrix(F64, 128, 128), pc.matrix(F64, 128, 128), pc.matrix(F64, 128, 128))
    A = pc.compose_of(pc.diagonal(F64, 128 ** 3), K)
    pc.check_case(emul_engine, A, batch=1)
    pc.check_case(emul_engine, pc.compose_of(pc.diagonal(F64, 128 ** 3), K, adj=True), batch=1)


def test_host_apply_async_entry(emul_engine):
    """po_plan_apply_host_async (enqueue only; the caller synchronises the stream) returns what the waiting form returns;
    wait=False without a caller-owned result buffer is refused."""
    import parops
    eng = emul_engine
    A_o = pc.fno_spectral(F32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)(pc.OracleNS)
    A_p = pc.fno_spectral(F32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)(pc.ProductNS)
    th_o, th_p = pc.paired_params(A_o, A_p, eng)
    x = np.asfortranarray(pc.randn(np.random.default_rng(5), F32, A_o.Domain, 2))
    with parops.use_engine(eng):
        At = A_p(th_p)
        out = np.zeros((A_o.Range, 2), dtype=np.float32, order="F")
        y = parops.apply_operator(At, x, out=out, wait=False)
        eng.synchronize()
        assert y is out and pc.rel_l2(out, A_o(x, th_o)) <= 1e-5
        assert np.array_equal(out, At * x)
        with pytest.raises(parops.ParException):
            parops.apply_operator(At, x, wait=False)


@pytest.mark.parametrize("n", [32, 128])
def test_c2_fused_pair(emul_engine, n, capfd):
    """po_c2.cuh: config C2's (R[1:16] o rDFT_n) on the slow dim fused with the full FFT128 on the fastest dim, forward
    (r2c -> fft128) and adjoint (fft128^-1 -> c2r), against the oracle; the same plans with PO_C2_FUSE=0 (three separate
    kernels) must agree with the fused ones to rounding."""
    import parops
    RF = pc.compose_of(pc.restriction(C64, n // 2 + 1, [(1, 16)]), pc.dft(F32, n))
    facs = (RF, pc.dft(C64, 4), pc.dft(C64, 128))
    os.environ["PO_TRACE"] = "1"
    try:
        capfd.readouterr()
        pc.check_case(emul_engine, pc.kron_of(*facs), batch=2)
        assert "fused C2 pair r2c -> fft128" in capfd.readouterr().err
        pc.check_case(emul_engine, pc.kron_of(*facs, adj=True), batch=2, seed=3)
        assert "fused C2 pair fft128^-1 -> c2r" in capfd.readouterr().err
        os.environ["PO_C2_FUSE"] = "0"
        pc.check_case(emul_engine, pc.kron_of(*facs), batch=2)
        pc.check_case(emul_engine, pc.kron_of(*facs, adj=True), batch=2, seed=3)
        assert "fused C2 pair" not in capfd.readouterr().err
    finally:
        os.environ.pop("PO_TRACE", None)
        os.environ.pop("PO_C2_FUSE", None)


def test_empty_operands(emul_engine):
    """`0 in size(x)` (src/ParDFT.jl:26,30, src/ParCommon.jl:73: the reference special-cases empty shards): an operand with no columns
    goes through every apply form as a no-op that returns the right (Range, 0) shape and type; the pullback of nothing leaves zero
    parameter cotangents."""
    import parops
    eng = emul_engine
    with parops.use_engine(eng):
        A = parops.ParDFT(parops.Float32, 8)
        y = A * eng.to_device(np.zeros((8, 0), dtype=F32, order="F"))
        assert tuple(y.shape) == (5, 0) and parops.jtype(y.dtype) is parops.ComplexF32
        x = A.adjoint() * eng.to_device(np.zeros((5, 0), dtype=C64, order="F"))
        assert tuple(x.shape) == (8, 0) and parops.jtype(x.dtype) is parops.Float32
        S = parops.spectral_convolution(parops.Float32, [8, 8], [[(1, 2), (7, 8)], [(1, 3)]], 3)
        th = parops.gpu(parops.init(S, rng=2))
        x0 = eng.to_device(np.zeros((S.Domain, 0), dtype=F32, order="F"))
        assert tuple((S(th) * x0).shape) == (S.Range, 0)
        yh = S(th) * np.zeros((S.Domain, 0), dtype=F32, order="F")          # host operand (po_plan_apply_host)
        assert isinstance(yh, np.ndarray) and yh.shape == (S.Range, 0)
        y0, back = parops.rrule(S(th), x0)
        grads, xbar = back(eng.to_device(np.zeros((S.Range, 0), dtype=F32, order="F")))
        assert tuple(y0.shape) == (S.Range, 0) and tuple(xbar.shape) == (S.Domain, 0)
        assert all(float(parops.cpu(g).__abs__().max()) == 0.0 for g in grads.values())


@pytest.mark.parametrize("variant", [1, 3])
def test_c2_fused_pair_compile_time_pitch(emul_engine, variant):
    """The compile-time row-pitch instantiations of the fused C2 kernels (I = 128 x 32 = 4096: every row address an immediate,
    inverse unrolled over its residues) and the 5-CTA register cap, forward and adjoint, against the oracle."""
    RF = pc.compose_of(pc.restriction(C64, 65, [(1, 16)]), pc.dft(F32, 128))
    facs = (RF, pc.dft(C64, 32), pc.dft(C64, 128))
    os.environ["PO_C2_VARIANT"] = str(variant)
    try:
        pc.check_case(emul_engine, pc.kron_of(*facs), batch=1)
        pc.check_case(emul_engine, pc.kron_of(*facs, adj=True), batch=1, seed=3)
    finally:
        os.environ.pop("PO_C2_VARIANT", None)


@pytest.mark.parametrize("shape,c,kx,kt,batch", [([4, 64, 16], 4, 4, 5, 2), ([2, 2, 64, 32], 2, 2, 3, 1), ([32, 64, 16], 4, 6, 8, 1), ([32, 64, 32], 6, 8, 4, 1)])
def test_fewer_kept_modes_reach_the_fused_kernels(emul_engine, shape, c, kx, kt, batch):
    """Widening pass (po_widen): a layer that keeps FEWER than 8 (+ 8) modes per dim -- here kx low + kx high spatial modes and kt
    time modes -- is compiled to the 8 / 8 + 8 kernels followed by restrictions of the few-MB spectral tensor (and the mirror image on
    the way back): the plan contains the fused (t, x) steps, and forward transform, its reference adjoint, the whole spectral convolution
    and the reverse mode agree with the oracle.  PO_WIDEN=0 (generic truncated-DFT kernels) must give the same numbers."""
    import parops
    modes = [[(1, kx), (s - kx + 1, s)] if s >= 16 else [(1, s)] for s in reversed(shape[:-1])] + [[(1, kt)]]
    fwd = pc.fno_forward_transform(F32, shape, modes, c)
    plan, _ = parops.get_plan(fwd(pc.ProductNS), batch, emul_engine)
    assert "fused-fwd-tx" in plan.describe(), plan.describe()
    pc.check_case(emul_engine, fwd, batch=batch)
    inv = pc.fno_forward_transform(F32, shape, modes, c, adj=True)
    plan_i, _ = parops.get_plan(inv(pc.ProductNS), batch, emul_engine)
    assert "fused-inv-xt" in plan_i.describe(), plan_i.describe()
    pc.check_case(emul_engine, inv, batch=batch)
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c, adj=True), batch=batch, seed=2)
    pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    os.environ["PO_WIDEN"] = "0"
    try:
        plan_g, _ = parops.get_plan(fwd(pc.ProductNS), batch, emul_engine, parops._lib.PO_PLAN_NO_PIPELINE)  # a distinct cache key
        assert "fused-" not in plan_g.describe(), plan_g.describe()
        pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch, flags=parops._lib.PO_PLAN_NO_PIPELINE)
    finally:
        os.environ.pop("PO_WIDEN", None)


def _transforms_jl_expressions():
    """The expression set of the reference's test/transforms.jl:22-47 (its `transforms(...)` rewrites belong to the out-of-scope
    ASTOptimization.jl; the expressions themselves are apply-path trees: Compose chains with identities, Kron of Composes,
    inverse-of-forward DFT chains, and the mixed real/complex chain of :45-47)."""
    ex = []
    for T in (F32, F64, C64, C128):
        i4, i8 = pc.identity(T, 4), pc.identity(T, 8)
        m4x4, m4x8, m8x4, m8x8 = pc.matrix(T, 4, 4), pc.matrix(T, 4, 8), pc.matrix(T, 8, 4), pc.matrix(T, 8, 8)
        dft4 = pc.dft(T, 4)
        idft4 = pc.dft(T, 4, adj=True)
        ex += [(T, pc.compose_of(m8x4, m4x4, m4x8)), (T, pc.compose_of(m4x8, m8x8, m8x4)),
               (T, pc.compose_of(m8x4, i4, m4x4, m4x8)), (T, pc.compose_of(m4x8, i8, m8x8, m8x4)),
               # Julia parses  i8 * (m8x8 * m8x8) ⊗ (m8x4 * m4x8)  left to right: (i8 * (m8x8 * m8x8)) ⊗ (m8x4 * m4x8)
               (T, pc.kron_of(pc.compose_of(i8, pc.compose_of(m8x8, m8x8)), pc.compose_of(m8x4, m4x8))),
               (T, pc.kron_of(pc.compose_of(i8, pc.compose_of(m8x8, m8x8)), pc.compose_of(m4x8, m8x4))),
               (T, pc.compose_of(idft4, dft4, m4x4))]
    for RT, CT in ((F32, C64), (F64, C128)):
        ex.append((RT, pc.compose_of(pc.matrix(CT, 4, 4), pc.matrix(CT, 4, 5), pc.dft(RT, 8), pc.matrix(RT, 8, 4), pc.diagonal(RT, 4))))
    return ex


def test_reference_transforms_jl_expressions(emul_engine):
    """Every expression the reference's test/transforms.jl builds, applied (vector and matrix operands) and differentiated, against
    the oracle."""
    for T, build in _transforms_jl_expressions():
        pc.check_case(emul_engine, build, batch=3)
        pc.check_case(emul_engine, build, vector=True)
    for T, build in _transforms_jl_expressions()[:7] + _transforms_jl_expressions()[-2:]:
        pc.check_pullback(emul_engine, build, batch=2)


@pytest.mark.parametrize("seed", range(12))
def test_random_truncated_spectral_layers(emul_engine, seed):
    """Randomised plan-level check (lower / schedule / fuse / widen / fast-path matchers): FNO-style layers with random extents (incl. the
    ones the specialised kernels take), random ASYMMETRIC kept sets (k_lo low + k_hi high bins per spatial dim, k_t low time bins), random
    channel counts and batch -- forward transform, its reference adjoint, the spectral convolution and the reverse mode against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    nd = int(rng.integers(1, 4))                                   # spatial dims
    sp = [int(rng.choice([8, 12, 16, 32, 64] if i == nd - 1 else [4, 8, 16, 32])) for i in range(nd)]   # last entry = fastest spatial dim
    nt = int(rng.choice([8, 16, 32]))
    shape = sp + [nt]
    c = int(rng.integers(1, 7))
    batch = int(rng.integers(1, 3))
    modes = []
    for s in reversed(sp):                                         # examples/fno.jl:17 pairs dim i with modes[end - i]
        klo, khi = int(rng.integers(1, min(8, s // 2) + 1)), int(rng.integers(0, min(8, s // 2) + 1))
        modes.append([(1, klo)] + ([(s - khi + 1, s)] if khi else []))
    modes.append([(1, int(rng.integers(1, min(8, nt // 2 + 1) + 1)))])
    pc.check_case(emul_engine, pc.fno_forward_transform(F32, shape, modes, c), batch=batch)
    pc.check_case(emul_engine, pc.fno_forward_transform(F32, shape, modes, c, adj=True), batch=batch, seed=1)
    pc.check_case(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    pc.check_pullback(emul_engine, pc.fno_spectral(F32, shape, modes, c), batch=batch)
    if seed % 4 == 0:
        pc.check_case(emul_engine, pc.fno_spectral(F64, shape, modes, c), batch=batch)


def _random_kron(rng, T):
    """2-4 random factors: ParMatrix (rectangular), ParDiagonal, ParIdentity, ParDFT, ParRestriction o ParDFT -- all on element type T (complex
    whenever a DFT takes part), optionally composed with a ParDiagonal on the Range."""
    cplx = np.dtype(T).kind == "c"
    nf = int(rng.integers(2, 5))
    facs = []
    for _ in range(nf):
        kind = int(rng.integers(0, 5 if cplx else 3))
        n = int(rng.choice([2, 3, 4, 5, 8, 12, 16]))
        if kind == 0:
            facs.append(pc.matrix(T, int(rng.integers(1, 7)), n))
        elif kind == 1:
            facs.append(pc.diagonal(T, n))
        elif kind == 2:
            facs.append(pc.identity(T, n))
        elif kind == 3:
            facs.append(pc.dft(T, n))
        else:
            k = int(rng.integers(1, n + 1))
            facs.append(pc.compose_of(pc.restriction(T, n, [(1, k)]), pc.dft(T, n)))
    return facs


@pytest.mark.parametrize("seed", range(16))
def test_random_kron_trees(emul_engine, seed):
    """Plan-level fuzzing of ParKron / ParCompose lowering, the order heuristic and the commuting / fusing passes: random Krons of random leaves,
    forward, adjoint and reverse mode against the oracle (vector and matrix operands).  A tree the reference would reject (its asserts) must be
    rejected by the mirror too."""
    rng = np.random.default_rng(4000 + seed)
    T = [C64, F32, C128, F64][seed % 4]
    facs = _random_kron(rng, T)
    batch = int(rng.integers(1, 4))
    try:
        pc.kron_of(*facs)(pc.OracleNS)
    except (AssertionError, Exception) as e:  # noqa: BLE001
        with pytest.raises(Exception):
            pc.kron_of(*facs)(pc.ProductNS)
        return
    pc.check_case(emul_engine, pc.kron_of(*facs), batch=batch)
    pc.check_case(emul_engine, pc.kron_of(*facs, adj=True), batch=batch, seed=1)
    pc.check_case(emul_engine, pc.kron_of(*facs), vector=True)
    pc.check_pullback(emul_engine, pc.kron_of(*facs), batch=batch)


def test_random_leaves(emul_engine):
    """Randomised leaf shapes: ParTensor (ranks 4-6, channel counts and mode extents on both sides of the streaming threshold), ParRestriction
    with overlapping / unordered / empty ranges (forward AND overwrite-on-overlap adjoint bit-exact, src/ParRestriction.jl:13-39) and ParDFT of
    arbitrary length (primes included) in all four element types.  314 such cases were run offline; 45 stay in the suite."""
    rng = np.random.default_rng(99)
    for i in range(15):
        rank, ic, oc, T = int(rng.choice([4, 5, 6])), int(rng.integers(1, 24)), int(rng.integers(1, 24)), [C64, F32, F64, C128][i % 4]
        modes = tuple(int(rng.integers(1, 12)) for _ in range(rank - 2))
        pc.check_case(emul_engine, pc.tensor(T, rank, ic, oc, modes), batch=int(rng.integers(1, 7)))
    for i in range(15):
        nn, nr, T = int(rng.integers(1, 40)), int(rng.integers(1, 5)), [F32, C64, F64, C128][i % 4]
        ranges = []
        for _ in range(nr):
            a = int(rng.integers(1, nn + 1))
            ranges.append((a, int(rng.integers(a - 1, nn + 1))))      # (a, a - 1) is an empty range
        if sum(max(0, e - s + 1) for s, e in ranges) == 0:
            continue
        b = int(rng.integers(1, 5))
        pc.check_case(emul_engine, pc.restriction(T, nn, ranges), batch=b, exact=True)
        pc.check_case(emul_engine, pc.restriction(T, nn, ranges, adj=True), batch=b, exact=True)
    for i in range(15):
        nn, T, b = int(rng.integers(1, 200)), [F32, C64, F64, C128][i % 4], int(rng.integers(1, 5))
        pc.check_case(emul_engine, pc.dft(T, nn), batch=b)
        pc.check_case(emul_engine, pc.dft(T, nn, adj=True), batch=b)
