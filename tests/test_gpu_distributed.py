"""Multi-GPU (-m gpu, needs >= 2 GPUs; skipped otherwise): the same workers as the gloo CPU
tests, but on the sm_100a library with its own NCCL communicator and the NVLink peer-store
repartition kernel.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu`."""
import os

import pytest
import torch

from tests import test_distributed_gloo as G

pytestmark = pytest.mark.gpu


def _need(n):
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs" % n)


@pytest.fixture(autouse=True)
def _nccl_backend(monkeypatch):
    monkeypatch.setenv("PO_TEST_BACKEND", "nccl")
    monkeypatch.setenv("PO_PEER_TIMEOUT_MS", "30000")  # a protocol bug must fail the test, not hang the box


@pytest.mark.parametrize("peer", [1, 0])
@pytest.mark.parametrize("world,din,dout,gs", [(2, [2, 1], [1, 2], (6, 5)), (2, [1, 2], [1, 2], (3, 4)), (2, [1, 1, 2], [2, 1, 1], (20, 16, 9)),
                                                  # long innermost runs: the warp-per-run path of po_peer_copy (ext[0] >= 64)
                                                  (2, [1, 1, 2], [1, 2, 1], (160, 6, 8)), (2, [1, 2, 1], [1, 1, 2], (96, 5, 6))])
def test_repartition_nccl(monkeypatch, peer, world, din, dout, gs):
    _need(world)
    monkeypatch.setenv("PO_NO_PEER", "0" if peer else "1")
    G._spawn(G._worker_repartition, world, din, dout, gs)


# the reference's own distributed test (test/distributed.jl:45-57: 8 ranks, uneven 10^3 split, [2,2,2] -> [8,1,1]), a one-directional
# pattern on 6 ranks ([2,3] -> [3,2]: rank 0 sends to rank 2 and never hears from it) and an uneven 4-rank pencil; 5 calls of the same
# plan each, so the epoch-parity staging regions are reused under the acknowledgement protocol
@pytest.mark.parametrize("peer", [1, 0])
@pytest.mark.parametrize("world,din,dout,gs", [(8, [2, 2, 2], [8, 1, 1], (10, 10, 10)), (6, [2, 3], [3, 2], (7, 11)), (4, [1, 4, 1], [2, 1, 2], (9, 10, 5)),
                                                  (8, [1, 1, 8], [1, 8, 1], (40, 16, 16))])
def test_repartition_nccl_many_ranks(monkeypatch, peer, world, din, dout, gs):
    _need(world)
    monkeypatch.setenv("PO_NO_PEER", "0" if peer else "1")
    G._spawn(G._worker_repartition, world, din, dout, gs, 5)


@pytest.mark.parametrize("world,din,dout,gs", [(6, [2, 3], [3, 2], (64, 66)), (2, [2, 1], [1, 2], (64, 64)), (8, [2, 2, 2], [8, 1, 1], (10, 10, 10))])
def test_repartition_peer_backpressure(monkeypatch, world, din, dout, gs):
    _need(world)
    monkeypatch.setenv("PO_NO_PEER", "0")
    G._spawn(G._worker_repartition_skew, world, din, dout, gs)


def test_repartition_peer_timeout_is_loud(monkeypatch):
    _need(2)
    monkeypatch.setenv("PO_NO_PEER", "0")
    monkeypatch.setenv("PO_PEER_TIMEOUT_MS", "400")
    G._spawn(G._worker_repartition_timeout, 2)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_fno_nccl(world):
    _need(world)
    G._spawn(G._worker_fno, world)


def test_reduce_nccl():
    _need(2)
    G._spawn(G._worker_reduce, 2)


def test_broadcasted_nccl():
    _need(2)
    G._spawn(G._worker_broadcast, 2)


def test_comm_init_all_single_process():
    """po_comm_init_all: one process, one context per device, ParRepartition through the NVLink peer-store kernel (peer access
    instead of IPC), bit-exact against the oracle -- the harness a single Julia session with several CuDevices would use."""
    _need(2)
    import numpy as np
    import parops
    from oracle import parops_oracle as O
    engs = [parops.Engine(d) for d in range(2)]
    try:
        parops.init_comm_all(engs)
        gs, din, dout, b = (12, 10, 6), [1, 2, 1], [2, 1, 1], 2
        xg = np.arange(int(np.prod(gs)) * b, dtype=np.float32).reshape(-1, b, order="F")
        xs = O.scatter_global(xg, din, gs)
        want = O.ParRepartition(O.F32, din, dout, gs).apply_all(xs)
        for rep in range(4):
            ys = []
            for r, eng in enumerate(engs):
                torch.cuda.set_device(r)
                with parops.use_engine(eng):
                    R = parops.ParRepartition(parops.Float32, parops.CartComm(din, r), parops.CartComm(dout, r), gs)
                    ys.append(R * eng.to_device(xs[r] + rep))
            for r in range(2):
                assert np.array_equal(parops.cpu(ys[r]), want[r] + rep)
    finally:
        for e in engs:
            e.destroy()
