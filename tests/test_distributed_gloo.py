"""World-size-2 (and 4) multi-process tests of the N>1 path on CPU: torch.distributed `gloo`
carries the messages (registered through po_comm_init_external), while the plan, the box
tables and the pack/unpack kernels are the library's own (CPU kernel-emulator build).
On the GPU box the same code path runs over NCCL (bench.py --gpus N)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tensor_at(addr, nbytes):
    return torch.frombuffer((ctypes.c_char * nbytes).from_address(addr), dtype=torch.uint8)


BACKEND = "gloo"  # test_gpu_distributed.py switches the workers to "nccl" (real library, one process per GPU)


def _setup(rank, world, port):
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import parops
    backend = os.environ.get("PO_TEST_BACKEND", "gloo")
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port,
                                device_id=torch.device("cuda", rank))
        eng = parops.default_engine()
        parops.init_comm(eng)
        return parops, eng
    from tests.emul.build_emul import build
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port)
    eng = parops.Engine(device="cpu", lib=parops._lib.bind(build()))

    def exchange(sends, recvs):
        reqs = [dist.irecv(_tensor_at(a, n), src=p) for p, a, n in recvs]
        reqs += [dist.isend(_tensor_at(a, n), dst=p) for p, a, n in sends]
        for r in reqs:
            r.wait()

    def collective(op, send, recv, count, dtype, root):
        if op == 0:
            t = _tensor_at(recv, count)
            dist.broadcast(t, root)
            return
        tdt = {0: torch.float32, 1: torch.float64, 2: torch.float32, 3: torch.float64}[dtype]
        n = count * (2 if dtype >= 2 else 1)
        nbytes = n * (4 if tdt == torch.float32 else 8)
        src = torch.frombuffer((ctypes.c_char * nbytes).from_address(send), dtype=tdt).clone()
        if op == 1:
            dist.reduce(src, root)
        else:
            dist.all_reduce(src)
        if op == 2 or dist.get_rank() == root:
            torch.frombuffer((ctypes.c_char * nbytes).from_address(recv), dtype=tdt).copy_(src)

    parops.init_comm_external(eng, rank, world, exchange, collective)
    return parops, eng


def _gather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _worker_repartition(rank, world, port, din, dout, gs, reps=1):
    """ParRepartition apply + adjoint against the oracle, bit-exact; `reps` > 2 calls of the SAME plan with fresh
    data exercise the epoch-parity staging and the consumed-epoch acknowledgements of the NVLink peer-store kernel."""
    parops, eng = _setup(rank, world, port)
    from oracle import parops_oracle as O
    n = int(np.prod(gs))
    b = 3
    with parops.use_engine(eng):
        R = parops.ParRepartition(parops.Float32, parops.CartComm(din), parops.CartComm(dout), gs)
        Ro = O.ParRepartition(O.F32, din, dout, gs)
        for rep in range(reps):
            xg = (np.arange(n * b, dtype=np.float32) + 1000.0 * rep).reshape(n, b, order="F")
            xs = O.scatter_global(xg, din, gs)
            y = parops.cpu(R * eng.to_device(xs[rank]))
            back = parops.cpu(R.adjoint() * eng.to_device(y))
            ys = Ro.apply_all(xs)
            assert np.array_equal(y, ys[rank]), "rank %d rep %d: repartition differs from the oracle" % (rank, rep)   # bit-exact
            assert np.array_equal(back, xs[rank])
        allys = _gather(y)
        if rank == 0:
            assert np.array_equal(O.gather_global(allys, dout, gs), xg)
        # the one-call C entry (po_repartition: cached one-node plan) moves the same boxes
        N = len(gs)
        xd = eng.to_device(xs[rank])
        y2 = eng.empty_colmajor(R.Range, b, parops.Float32)
        eng.lib.check(eng.lib.po_repartition(eng.ctx, parops.Float32.code, N, (ctypes.c_int64 * N)(*gs), (ctypes.c_int32 * N)(*din),
                                             (ctypes.c_int32 * N)(*dout), rank, 0, b, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(y2.data_ptr()),
                                             eng.stream_ptr()))
        assert np.array_equal(parops.cpu(y2), y)
        assert eng.device_status() == 0
    dist.destroy_process_group()


def _worker_repartition_skew(rank, world, port, din, dout, gs):
    """Back-to-back applies with one rank far behind: without back-pressure a fast sender overwrites a staging region
    its receiver has not unpacked yet (one-directional patterns such as [2,3] -> [3,2])."""
    import time
    parops, eng = _setup(rank, world, port)
    from oracle import parops_oracle as O
    n = int(np.prod(gs))
    with parops.use_engine(eng):
        R = parops.ParRepartition(parops.Float32, parops.CartComm(din), parops.CartComm(dout), gs)
        Ro = O.ParRepartition(O.F32, din, dout, gs)
        xgs = [(np.arange(n * 2, dtype=np.float32) + 7777.0 * rep).reshape(n, 2, order="F") for rep in range(6)]
        xd = [eng.to_device(O.scatter_global(xg, din, gs)[rank]) for xg in xgs]
        R * xd[0]  # plan creation (collective) outside the skewed region
        torch.cuda.synchronize() if BACKEND_IS_NCCL() else None
        dist.barrier()
        if rank == world - 1:
            time.sleep(0.5)  # the others enqueue all six exchanges while this rank has not started
        ys = [R * x for x in xd]  # asynchronous launches: fast ranks run epochs ahead as far as the protocol lets them
        for rep, y in enumerate(ys):
            want = Ro.apply_all(O.scatter_global(xgs[rep], din, gs))[rank]
            assert np.array_equal(parops.cpu(y), want), "rank %d rep %d" % (rank, rep)
        assert eng.device_status() == 0
    dist.destroy_process_group()


def BACKEND_IS_NCCL():
    return os.environ.get("PO_TEST_BACKEND", "gloo") == "nccl"


def _worker_repartition_timeout(rank, world, port):
    """A peer that does not show up within PO_PEER_TIMEOUT_MS must be LOUD: NaN-poisoned boxes, ParException at the
    next host synchronisation point and on every later apply until the status word has been read."""
    import time
    parops, eng = _setup(rank, world, port)
    gs, din, dout = (8, 6), [2, 1], [1, 2]
    n_loc = 4 * 6
    with parops.use_engine(eng):
        R = parops.ParRepartition(parops.Float32, parops.CartComm(din), parops.CartComm(dout), gs)
        x = eng.to_device(np.ones((n_loc, 1), dtype=np.float32))
        parops.cpu(R * x)  # healthy first exchange (plan creation is collective)
        dist.barrier()
        if rank == 1:
            time.sleep(3.0)  # >> PO_PEER_TIMEOUT_MS
            y = parops.cpu(R * x)  # rank 0's data of this epoch arrived long ago: completes normally
            assert np.all(y == 1.0)
        else:
            y = R * x
            with pytest.raises(parops.ParException, match="timed out"):
                parops.cpu(y)
            yh = y.cpu().numpy()
            assert np.isnan(yh).any() and not np.isnan(yh).all()  # the peer's box is poisoned, the self overlap is not
            assert eng.device_status() == 0  # the raise above read and cleared the word
        dist.barrier()
    dist.destroy_process_group()


def _worker_fno(rank, world, port):
    """Distributed forward transform Kron(RF_t, RF_z, RF_y, RF_x, I_c) over z-slabs, local
    frequency weights, adjoint pipeline back -- bench.py's C4 operator at toy size -- against
    the serial oracle on the gathered array."""
    parops, eng = _setup(rank, world, port)
    from oracle import parops_oracle as O
    import bench
    bench.CHANNELS, bench.MODES = 3, (2 if world <= 2 else (4 if world <= 4 else 8))
    M = bench.MODES
    shape = [2 * M, 8, 8, 8] if world <= 4 else [16, 16, 16, 16]  # nx, ny, nz, nt
    with parops.use_engine(eng):
        St, Sadj, dom, ran = bench.build_distributed(parops, shape, world, rank)
        # serial oracle operator with the same parameters
        nx, ny, nz, nt = shape
        T, CT = O.F32, O.C64

        def rf(Tin, n, rng):
            Fo = O.ParDFT(Tin, n)
            return O.ParCompose(O.ParRestriction(CT, Fo.Range, rng), Fo)
        Ko = O.ParKron(rf(T, nt, bench.time_modes()), rf(CT, nz, bench.space_modes(nz)), rf(CT, ny, bench.space_modes(ny)),
                       rf(CT, nx, bench.space_modes(nx)), O.ParIdentity(T, 3))
        gs_in, gs_mid = (3, nx, ny, nz, nt), (3, 2 * M, 2 * M, 2 * M, M)
        dims, dims_mid = [1, 1, 1, world, 1], [1, 1, 1, 1, world]
        rng = np.random.default_rng(0)
        xg = rng.standard_normal((Ko.Domain, 2)).astype(np.float32)
        x_loc = O.scatter_global(xg, dims, gs_in)[rank]
        # forward transform only (first stage of S)
        Tdist = St.ops[-1] if not hasattr(St.ops[-1], "ops") else St
        # St = compose(Tdist', W, Tdist): its flattened ops end with Tdist's stages
        z_ref = Ko(xg)
        n_t_stages = len([o for o in St.ops]) // 2  # symmetric pipeline around W
        stages = St.ops[len(St.ops) - n_t_stages:]
        Tloc = parops.compose(*stages)
        z = parops.cpu(Tloc * eng.to_device(x_loc))
        zs = _gather(z)
        if rank == 0:
            zg = O.gather_global(zs, dims_mid, gs_mid)
            err = np.linalg.norm(zg - z_ref) / np.linalg.norm(z_ref)
            assert err <= 1e-5, err
        # inverse pipeline: T' T x == projector applied to x; compare with the oracle's Ko'
        back = parops.cpu(parops.compose(*St.ops[:n_t_stages]) * eng.to_device(z))
        backs = _gather(back)
        if rank == 0:
            bg = O.gather_global(backs, dims, gs_in)
            ref = Ko.adjoint()(z_ref)
            err = np.linalg.norm(bg - ref) / np.linalg.norm(ref)
            assert err <= 1e-5, err
    dist.destroy_process_group()


def _worker_bench_check(rank, world, port):
    """bench.py's one-off distributed-vs-serial-oracle check (printed as `dist_check` beside every N > 1 number)."""
    parops, eng = _setup(rank, world, port)
    import bench
    bench.CHANNELS, bench.MODES = 3, 2
    with parops.use_engine(eng):
        res = bench.distributed_check(parops, eng, world, rank, shape=[8, 16, 4, 8])  # unequal extents: dim-order mistakes show
    if rank == 0:
        assert res["rel_l2_fwd"] <= 1e-5 and res["rel_l2_adj"] <= 1e-5, res
    else:
        assert res is None
    dist.destroy_process_group()


def _worker_bench_check_uneven(rank, world, port, shape, modes):
    """The same check on grids that do NOT divide by the rank count: uneven z-slabs, and (modes < ranks) ranks whose slab of kept t-modes
    is EMPTY -- their local steps are skipped, their repartitions still take part (src/ParCommon.jl:73 `0 in size(x)`)."""
    parops, eng = _setup(rank, world, port)
    import bench
    bench.CHANNELS, bench.MODES = 3, modes
    with parops.use_engine(eng):
        res = bench.distributed_check(parops, eng, world, rank, shape=shape)
    if rank == 0:
        assert res["rel_l2_fwd"] <= 1e-5 and res["rel_l2_adj"] <= 1e-5, res
    dist.destroy_process_group()


def _worker_broadcast(rank, world, port):
    """ParBroadcasted: root-only init, A(θ) broadcasts θ, gradients sum-reduce to root."""
    parops, eng = _setup(rank, world, port)
    with parops.use_engine(eng):
        A = parops.ParBroadcasted(parops.ParMatrix(parops.ComplexF32, 4, 5))
        theta = parops.init(A, rng=7 + rank)           # only root gets an entry (src/ParBroadcasted.jl:17-21)
        assert (len(theta) == 1) == (rank == 0)
        At = A(theta)
        W = parops.cpu(At.op.params)
        Ws = _gather(W)
        assert all(np.array_equal(Ws[0], w) for w in Ws)
        # a ROW-major (torch default) θ on root must arrive as the same matrix, not transposed/garbled
        want = np.arange(20, dtype=np.float32).reshape(4, 5).astype(np.complex64)
        th2 = {A.op: torch.from_numpy(want.copy()).to(eng.torch_device)} if rank == 0 else {}
        W2s = _gather(parops.cpu(A(th2).op.params))
        assert all(np.array_equal(want, w) for w in W2s)
        g = {A.op: eng.to_device(np.full((4, 5), rank + 1, dtype=np.complex64))}
        red = parops.reduce_gradients(g, 0, eng)
        if rank == 0:
            assert np.allclose(parops.cpu(red[A.op]), sum(range(1, world + 1)))
        else:
            assert red == {}
    dist.destroy_process_group()


def _worker_reduce(rank, world, port):
    """ParReduce (src/ParReduce.jl): allreduce-sum on every rank; the rrule passes the cotangent through."""
    parops, eng = _setup(rank, world, port)
    with parops.use_engine(eng):
        R = parops.ParReduce(parops.Float32)
        x = eng.to_device(np.arange(12, dtype=np.float32).reshape(4, 3, order="F") + 100 * rank)
        y = parops.cpu(R * x)
        want = sum(np.arange(12, dtype=np.float32).reshape(4, 3, order="F") + 100 * r for r in range(world))
        assert y.shape == (4, 3) and np.array_equal(y, want)
        y2, back = R.rrule(x)
        assert np.array_equal(parops.cpu(y2), want)
        yb = eng.to_device(np.ones((4, 3), dtype=np.float32))
        assert back(yb) is yb
        assert R.adjoint() is R
        Rc = parops.ParReduce(parops.ComplexF64)
        z = eng.to_device(np.full(5, (rank + 1) * (1 + 2j), dtype=np.complex128))
        assert np.allclose(parops.cpu(Rc(z)), sum(range(1, world + 1)) * (1 + 2j))
    dist.destroy_process_group()


def _spawn(fn, world, *args):
    # the rendezvous port is picked by bind(0) and released before the workers listen on it: another socket of
    # the box (NCCL bootstrap of the previous test, TIME_WAIT) can take it in between -- retry on EADDRINUSE only
    for attempt in range(4):
        try:
            mp.spawn(fn, args=(world, _free_port()) + args, nprocs=world, join=True)
            return
        except Exception as e:  # noqa: BLE001
            if attempt == 3 or ("EADDRINUSE" not in str(e) and "address already in use" not in str(e).lower()):
                raise


@pytest.mark.parametrize("world,din,dout,gs", [(2, [2, 1], [1, 2], (6, 5)), (4, [2, 2, 1], [4, 1, 1], (5, 4, 3)), (2, [1, 2], [1, 2], (3, 4)),
                                                  (2, [1, 1, 2], [1, 2, 1], (160, 6, 8))])
def test_repartition_gloo(world, din, dout, gs):
    _spawn(_worker_repartition, world, din, dout, gs)


def _worker_repartition_random(rank, world, port, seed, ncases):
    """Several random (global size, dims_in, dims_out) triples on ONE process group: uneven blocks, empty blocks, asymmetric send / recv
    peer sets -- apply and adjoint bit-exact against the oracle's SPMD emulation of src/ParRepartition.jl:140-192."""
    parops, eng = _setup(rank, world, port)
    from oracle import parops_oracle as O
    from tests.test_abi import _factorisations
    rng = np.random.default_rng(seed)          # same stream on every rank
    with parops.use_engine(eng):
        for case in range(ncases):
            N = int(rng.integers(1, 4))
            gs = tuple(int(rng.integers(1, 9)) for _ in range(N))
            din, dout = list(_factorisations(world, N, rng)), list(_factorisations(world, N, rng))
            n, b = int(np.prod(gs)), int(rng.integers(1, 4))
            R = parops.ParRepartition(parops.Float64, parops.CartComm(din), parops.CartComm(dout), gs)
            Ro = O.ParRepartition(O.F64, din, dout, gs)
            xg = (np.arange(n * b, dtype=np.float64) + 7.0 * case).reshape(n, b, order="F")
            xs = O.scatter_global(xg, din, gs)
            y = parops.cpu(R * eng.to_device(xs[rank]))
            ys = Ro.apply_all(xs)
            assert np.array_equal(y, ys[rank]), "case %d %r %r -> %r rank %d: repartition differs from the oracle" % (case, gs, din, dout, rank)
            back = parops.cpu(R.adjoint() * eng.to_device(y))
            assert np.array_equal(back, xs[rank]), "case %d %r: adjoint does not undo the repartition" % (case, gs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed", [(6, 11), (4, 12), (3, 13)])
def test_repartition_gloo_random_grids(world, seed):
    _spawn(_worker_repartition_random, world, seed, 8)


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_fno_gloo(world):
    _spawn(_worker_fno, world)


def test_bench_distributed_check_gloo():
    _spawn(_worker_bench_check, 2)


@pytest.mark.parametrize("world,shape,modes", [(3, [8, 16, 7, 8], 2), (3, [8, 8, 10, 8], 4), (4, [8, 8, 9, 8], 2)])
def test_bench_distributed_check_uneven_and_empty_shards_gloo(world, shape, modes):
    _spawn(_worker_bench_check_uneven, world, shape, modes)


def test_reduce_gloo():
    _spawn(_worker_reduce, 2)


def test_broadcasted_gloo():
    _spawn(_worker_broadcast, 2)
