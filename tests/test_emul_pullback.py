"""Reverse mode (po_plan_apply_taped + po_plan_pullback) on the CPU kernel emulator against
the oracle's pullback (true DFT adjoints, validated against finite differences)."""
import os

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
