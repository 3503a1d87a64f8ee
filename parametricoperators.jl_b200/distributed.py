"""Distribution layer: ParDistributed, ParBroadcasted, ParRepartition, ``distribute``
(src/ParDistributed.jl, src/ParBroadcasted.jl, src/ParRepartition.jl, src/ParKron.jl:244-309,
src/ParTensor.jl:153-196, src/ParDiagonal.jl:62-65, src/ParMatrix.jl:21).

The reference does its cartesian bookkeeping with MPI communicators; here a communicator is
plain integers (``CartComm``: dims + this process' rank; MPI_Cart_create semantics with
reorder=false, i.e. row-major rank <-> coords, last dim fastest) and the data plane is NCCL
inside ``libparops.so`` (one process per GPU, rendezvous through ``torch.distributed``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ParException
from .operators import (Internal, External, Linear, ParDiagonal, ParIdentity, ParKron, ParMatrix, ParOperator,
                        ParParameterized, ParTensor, Parametric, ParCompose, Float64, jtype)

__all__ = ["init_comm_all", "local_size", "range_axes", "CartComm", "world", "set_world", "ParDistributed", "ParBroadcasted", "bcasted",
           "ParRepartition", "ParReduce", "distribute", "init_comm", "init_comm_external", "reduce_gradients"]


def local_size(global_size: int, rank: int, num_ranks: int) -> int:
    """src/ParCommon.jl:57-64 / src/ParDistributed.jl:7-14"""
    s = global_size // num_ranks
    if rank < global_size % num_ranks:
        s += 1
    return s


def range_axes(dims: Sequence[int], shape: Sequence[int]):
    """src/ParCommon.jl:45-55 -- per-axis lists of 1-based inclusive (start, stop) block ranges."""
    if len(dims) != len(shape):
        raise ParException("range_axes: length(dims) != length(shape)")
    axes = []
    for d in range(len(dims)):
        sizes = [local_size(shape[d], i, dims[d]) for i in range(dims[d])]
        off, ax = 0, []
        for s in sizes:
            ax.append((off + 1, off + s))
            off += s
        axes.append(ax)
    return axes


class _World:
    rank, size = 0, 1


world = _World()


def set_world(rank: int, size: int):
    """Rank / size of MPI.COMM_WORLD's stand-in (set by ``init_comm`` from torch.distributed)."""
    world.rank, world.size = int(rank), int(size)


class CartComm:
    """``MPI.Cart_create(parent, dims)``: dims + this process' rank; coords are row-major."""

    def __init__(self, dims: Sequence[int], rank: Optional[int] = None):
        self.dims = [int(d) for d in dims]
        self.size = int(np.prod(self.dims))
        self.rank = world.rank if rank is None else int(rank)
        if self.rank >= self.size:
            raise ParException("rank %d is not part of a %s cartesian communicator (MPI.COMM_NULL)" % (self.rank, self.dims))

    def coords(self, rank: Optional[int] = None) -> List[int]:
        r = self.rank if rank is None else rank
        c = []
        for d in reversed(self.dims):
            c.append(r % d)
            r //= d
        return c[::-1]

    def cart_rank(self, coords: Sequence[int]) -> int:
        r = 0
        for d, c in zip(self.dims, coords):
            r = r * d + c
        return r


class ParDistributed(ParOperator):
    """1/size-th local shard marker (src/ParDistributed.jl:16-29)."""
    T = Internal

    def __init__(self, op, rank, size):
        self.op, self.rank, self.size = op, int(rank), int(size)
        self.D, self.R, self.L, self.P = op.D, op.R, op.L, op.P

    Domain = property(lambda s: local_size(s.op.Domain, s.rank, s.size))
    Range = property(lambda s: local_size(s.op.Range, s.rank, s.size))

    def children(self):
        return [self.op]

    def rebuild(self, cs):
        return ParDistributed(cs[0], self.rank, self.size)

    def adjoint(self):
        return ParDistributed(self.op.adjoint(), self.rank, self.size)

    def init_(self, d, rng=None):
        if self.P is Parametric:
            raise ParException("`init()` must be implemented on a case-by-case basis for distributed operators")
        return None


class ParBroadcasted(ParOperator):
    """Replicated-parameter wrapper (src/ParBroadcasted.jl:3-57): init on root only,
    ``A(θ)`` broadcasts θ from root, apply delegates, parameter gradients are sum-reduced to
    root (``reduce_gradients``) -- on device through NCCL instead of through the host."""
    T = Internal

    def __init__(self, op, comm=None, root: int = 0):
        self.op, self.comm, self.root = op, comm, int(root)
        self.D, self.R, self.L, self.P = op.D, op.R, op.L, op.P

    Domain = property(lambda s: s.op.Domain)
    Range = property(lambda s: s.op.Range)

    def children(self):
        return [self.op]

    def rebuild(self, cs):
        return ParBroadcasted(cs[0], self.comm, self.root)

    def adjoint(self):
        return ParBroadcasted(self.op.adjoint(), self.comm, self.root)

    def _rank(self):
        return self.comm.rank if self.comm is not None else world.rank

    def init_(self, d, rng=None):
        if self._rank() == self.root:  # src/ParBroadcasted.jl:17-21
            self.op.init_(d, rng)
        return d

    def parameterize(self, theta):
        if self.P is not Parametric:
            return self
        if self.op.T is not External:
            raise ParException("ParBroadcasted must wrap an external (leaf) operator")
        from .engine import current_engine
        eng = current_engine()
        leaf = self.op
        shape = {ParMatrix: lambda a: (a.m, a.n), ParDiagonal: lambda a: (a.n,), ParTensor: lambda a: a.weight_shape}[type(leaf)](leaf)
        if self._rank() == self.root:
            t = theta[leaf]
            # to_device hands back column-major (Julia order) storage for NumPy arrays AND torch tensors: the raw bytes
            # that po_bcast ships must be laid out like the column-major buffers the other ranks allocate
            t = eng.to_device(t)
            if tuple(t.shape) != tuple(shape) or t.dtype != leaf.D.torch:
                raise ParException("parameter of %r has shape/type %s/%s, expected %s/%s" % (leaf, tuple(t.shape), t.dtype, tuple(shape), leaf.D))
        else:
            import torch
            rev = tuple(reversed(shape))
            t = torch.empty(rev, dtype=leaf.D.torch, device=eng.torch_device)
            t = t.permute(*reversed(range(len(shape)))) if len(shape) > 1 else t
        if world.size > 1:
            nbytes = int(np.prod(shape)) * leaf.D.itemsize
            eng.lib.check(eng.lib.po_bcast(eng.ctx, C.c_void_p(t.data_ptr()), nbytes, self.root, eng.stream_ptr()))
        return ParBroadcasted(leaf.parameterize({leaf: t}), self.comm, self.root)


def bcasted(A, comm=None, root=0):
    return ParBroadcasted(A, comm, root)


def reduce_gradients(grads: dict, root: int = 0, engine=None) -> dict:
    """Sum parameter cotangents over ranks onto ``root`` -- the ``MPI.Reduce(SUM)`` of the
    ParBroadcasted rrule (src/ParBroadcasted.jl:41-57), in place on the device."""
    from .engine import current_engine
    eng = engine or current_engine()
    out = {}
    for k, g in grads.items():
        eng.lib.check(eng.lib.po_reduce_sum(eng.ctx, C.c_void_p(g.data_ptr()), C.c_void_p(g.data_ptr()), g.numel(),
                                            jtype(g.dtype).code, root, eng.stream_ptr()))
        if world.rank == root:
            out[k] = g
    return out


class ParReduce(ParOperator):
    """``ParReduce{T}`` (src/ParReduce.jl:6-40): sum of the operand over the ranks of the communicator, returned on
    every rank (``MPI.Allreduce(x, SUM)``; the reference round-trips device arrays through the host, here it is
    ``po_allreduce_sum`` on the device).  Linear, non-parametric, external; shape-agnostic like the reference
    (``Domain``/``Range`` are not defined there either), so it is applied directly rather than compiled into a plan.
    Its rrule passes the cotangent through unchanged (src/ParReduce.jl:42-63)."""
    T = External

    def __init__(self, T=None, comm=None):
        self.D = self.R = jtype(Float64 if T is None else T)
        self.comm = comm

    Domain = property(lambda s: None)
    Range = property(lambda s: None)

    def children(self):
        return []

    def adjoint(self):
        return self  # a sum over ranks is self-adjoint

    def apply(self, x, engine=None):
        from .engine import current_engine
        eng = engine or current_engine()
        if jtype(x.dtype) is not self.D:
            raise ParException("operand element type %s does not match ParReduce{%s}" % (jtype(x.dtype), self.D))
        src = x if x.is_contiguous() else x.contiguous()
        out = src.clone()
        eng.lib.check(eng.lib.po_allreduce_sum(eng.ctx, C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), out.numel(),
                                               self.D.code, eng.stream_ptr()))
        return out.reshape(x.shape)

    def __call__(self, arg):
        return self.apply(arg)

    def __mul__(self, other):
        if isinstance(other, ParOperator):
            return super().__mul__(other)
        return self.apply(other)

    def rrule(self, x, engine=None):
        """``y, pullback``; ``pullback(ybar) -> ybar`` (src/ParReduce.jl:42-63)."""
        return self.apply(x, engine), (lambda ybar: ybar)


class ParRepartition(ParOperator):
    """N-d block redistribution between two cartesian grids (src/ParRepartition.jl:3-205).
    The send/recv box plan comes from ``po_repartition_plan`` -- the same host code the apply
    uses -- so ``send``/``recv`` here are exactly what the kernels move."""

    def __init__(self, T, comm_in: CartComm, comm_out: CartComm, global_size: Sequence[int], lib=None):
        self.D = self.R = jtype(T)
        if comm_in is None or comm_out is None:
            raise ParException("ParRepartition: comm_in / comm_out must not be MPI.COMM_NULL")
        self.comm_in, self.comm_out = comm_in, comm_out
        self.dims_in, self.dims_out = list(comm_in.dims), list(comm_out.dims)
        self.global_size = tuple(int(g) for g in global_size)
        self.N = len(self.global_size)
        if not (self.N == len(self.dims_in) == len(self.dims_out)):
            raise ParException("ParRepartition: dimensionality mismatch")
        self.rank = comm_in.rank
        if lib is None:
            from .engine import current_engine
            try:
                lib = current_engine().lib
            except ParException:
                lib = _lib.default_lib()
        N = self.N
        li, lo = (C.c_int64 * N)(), (C.c_int64 * N)()
        cap = int(np.prod(self.dims_in)) + int(np.prod(self.dims_out)) + 1
        rec = 1 + 2 * N
        sbuf, rbuf = (C.c_int64 * (cap * rec))(), (C.c_int64 * (cap * rec))()
        ns, nr = C.c_int32(), C.c_int32()
        lib.check(lib.po_repartition_plan(N, (C.c_int64 * N)(*self.global_size), (C.c_int32 * N)(*self.dims_in),
                                          (C.c_int32 * N)(*self.dims_out), self.rank, li, lo, sbuf, cap, C.byref(ns), rbuf,
                                          cap, C.byref(nr)))
        self.local_size_in, self.local_size_out = tuple(li), tuple(lo)

        def unpack(buf, n):
            out = []
            for k in range(n):
                r = buf[k * rec:(k + 1) * rec]
                out.append((int(r[0]), tuple((int(r[1 + 2 * d]), int(r[2 + 2 * d])) for d in range(N))))
            return out

        self.send, self.recv = unpack(sbuf, ns.value), unpack(rbuf, nr.value)

    Domain = property(lambda s: int(np.prod(s.local_size_in)))
    Range = property(lambda s: int(np.prod(s.local_size_out)))

    def adjoint(self):
        return ParRepartition(self.D, self.comm_out, self.comm_in, self.global_size)  # src/ParRepartition.jl:120-131


def distribute(A, *args, **kw):
    """``distribute`` for ParKron / ParMatrix / ParDiagonal / ParTensor."""
    if isinstance(A, ParKron):
        return _distribute_kron(A, *args, **kw)
    if isinstance(A, ParMatrix):
        return ParBroadcasted(A)  # src/ParMatrix.jl:21
    if isinstance(A, ParDiagonal):  # src/ParDiagonal.jl:62-65
        rank, size = (args[0].rank, args[0].size) if args else (world.rank, world.size)
        return ParDiagonal(A.D, local_size(A.n, rank, size))
    if isinstance(A, ParTensor):
        return _distribute_tensor(A, *args, **kw)
    raise ParException("no distribute method for %s" % type(A).__name__)


def _distribute_kron(A: ParKron, dims_in, dims_out=None, rank: Optional[int] = None):
    """src/ParKron.jl:244-309"""
    dims = [int(d) for d in dims_in]
    dims_out = [int(d) for d in (dims_out if dims_out is not None else dims_in)]
    N = len(dims)
    if len(A.ops) != N:
        raise ParException("distribute: length(A.ops) != length(dims)")
    comm_prev = CartComm(dims, rank)
    dims_prev = list(dims)
    size_curr = [o.Domain for o in reversed(A.ops)]
    ops = []
    for i in range(N):
        o = A.order[i]
        d = N - o + 1
        Ai = A.ops[o - 1]
        if isinstance(Ai, ParIdentity):
            continue
        dims_i = list(dims)
        dims_i[d - 1] = 1
        dims_i[d % N] *= dims[d - 1]  # mod1(d+1, N)
        comm_i = CartComm(dims_i, rank)
        coords_i = comm_i.coords()
        if dims_prev != dims_i:
            ops.insert(0, ParRepartition(Ai.D, comm_prev, comm_i, tuple(size_curr)))
        lower_ = [ParDistributed(ParIdentity(Ai.D, size_curr[j]), coords_i[j], dims_i[j]) for j in range(d, N)]
        upper_ = [ParDistributed(ParIdentity(Ai.D, size_curr[j]), coords_i[j], dims_i[j]) for j in range(0, d - 1)]
        bc = ParBroadcasted(Ai, comm_i).rebuild([Ai])
        ops.insert(0, ParKron(*lower_[::-1], bc, *upper_[::-1]))
        size_curr[d - 1] = Ai.Range
        comm_prev, dims_prev = comm_i, dims_i
    if dims_prev != dims_out:
        ops.insert(0, ParRepartition(A.ops[A.order[-1] - 1].R, comm_prev, CartComm(dims_out, rank), tuple(size_curr)))
    if not ops:
        raise ParException("distribute: Kronecker product of identities yields an empty composition")
    return ParCompose(*ops)


def _distribute_tensor(A: ParTensor, dims_in, comm: Optional[CartComm] = None):
    """src/ParTensor.jl:153-196"""
    if len(dims_in) != len(A.input_shape):
        raise ParException("distribute(ParTensor): length(dims_in) != length(input_shape)")
    cart = CartComm(dims_in, comm.rank if comm is not None else None)
    coords = cart.coords()
    combined = tuple(A.input_order) + tuple(A.weight_order) + tuple(A.target_order)
    ni, nt, nw = list(A.input_shape), list(A.target_shape), list(A.weight_shape)
    for i, dim in enumerate(A.input_order):
        across = dims_in[dim - 1]
        if sum(1 for e in combined if e == dim) == 2 and across != 1:
            raise ParException("cannot distribute across a contracted dimension")
        if A.input_shape[i] % across != 0:
            raise ParException("only perfect distribution is supported")
        ni[i] = A.input_shape[i] // across
        for j, dj in enumerate(A.weight_order):
            if dj == dim:
                nw[j] = A.weight_shape[j] // across
                break
        for j, dj in enumerate(A.target_order):
            if dj == dim:
                nt[j] = A.target_shape[j] // across
                break
    return ParTensor(A.D, A.weight_order, tuple(nw), A.input_order, tuple(ni), A.target_order, tuple(nt),
                     "%s:(%s)" % (A.id, ",".join(str(c) for c in coords)))


# ---------------------------------------------------------------------------------------
# communicator bring-up
# ---------------------------------------------------------------------------------------
def init_comm(engine=None):
    """Create the library's NCCL communicator for this process (one process per GPU):
    the 128-byte unique id is made on rank 0 and shipped with ``torch.distributed``."""
    import torch
    import torch.distributed as dist
    from .engine import current_engine
    eng = engine or current_engine()
    if not dist.is_initialized():
        set_world(0, 1)
        return eng
    rank, size = dist.get_rank(), dist.get_world_size()
    set_world(rank, size)
    if size == 1:
        return eng
    idbuf = (C.c_uint8 * 128)()
    if rank == 0:
        eng.lib.check(eng.lib.po_comm_unique_id(idbuf))
    dev = eng.torch_device if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    idbytes = bytes(t.cpu().tolist())
    buf = C.create_string_buffer(idbytes, 128)
    eng.lib.check(eng.lib.po_comm_init_rank(eng.ctx, size, rank, buf))
    return eng


def init_comm_all(engines):
    """``po_comm_init_all``: ONE process driving several devices, one engine per device; rank = position in the list.
    Carries ParRepartition (NVLink peer-store kernel); build the operators with explicit ``CartComm(dims, rank)``."""
    arr = (C.c_void_p * len(engines))(*[e.ctx for e in engines])
    engines[0].lib.check(engines[0].lib.po_comm_init_all(len(engines), arr))
    return engines


def init_comm_external(engine, rank: int, size: int, exchange, collective=None):
    """Register a host-supplied transport (see ``po_comm_init_external`` in parops.h).
    ``exchange(sends, recvs)`` gets lists of ``(peer, address, nbytes)``."""
    set_world(rank, size)

    def _ex(user, ns, sp, sb, sby, nr, rp, rb, rby, stream):
        try:
            sends = [(sp[i], sb[i], sby[i]) for i in range(ns)]
            recvs = [(rp[i], rb[i], rby[i]) for i in range(nr)]
            exchange(sends, recvs)
            return 0
        except Exception as e:  # noqa: BLE001 -- must not propagate through the C frame
            print("exchange callback failed:", repr(e))
            return 1

    def _co(user, op, send, recv, count, dtype, root, stream):
        try:
            if collective is None:
                return 2
            collective(op, send, recv, count, dtype, root)
            return 0
        except Exception as e:  # noqa: BLE001
            print("collective callback failed:", repr(e))
            return 1

    ex, co = _lib.EXCHANGE_FN(_ex), _lib.COLLECTIVE_FN(_co)
    engine._keep.extend([ex, co])
    engine.lib.check(engine.lib.po_comm_init_external(engine.ctx, size, rank, ex, co, None))
    return engine
