"""CPU oracle for the ParametricOperators.jl spectral-operator apply path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``parametricoperators.jl_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (Julia) cannot be executed in the build
container (no julia, no libfftw3, no MPI) and it ships no golden vectors, no
known-answer tests and only unseeded-``rand`` tests.  The
arithmetic itself lives in un-vendored, un-pinned third-party packages (FFTW.jl,
LinearAlgebra/OpenBLAS, NNlib ``batched_mul``, MPI.jl; ``Project.toml:6-21`` has
no ``[compat]`` block).  This file therefore restates the *published* algorithms
of those packages (FFTW conventions: unnormalised forward DFT, 1/n backward,
``rfft`` keeps n/2+1 bins, ``irfft(x, n)`` drops Im(DC)/Im(Nyquist) -- exactly the
conventions of ``scipy.fft``/pocketfft which is used here) arranged in the
reference's exact operation sequence, and is anchored on analytic known answers
and on the structural pins the reference's tests do hold
(``test/types_and_sizes.jl:35-57``, ``test/serial.jl:35-36,46-54``,
``test/distributed.jl:45-57``).

Array convention: Julia arrays are column-major.  Every array here is a NumPy
array handled with ``order='F'`` reshapes, so ``x`` has shape ``(Domain, batch)``
exactly like the reference's ``AbstractMatrix`` argument and
``x.reshape(s, order='F')`` is Julia's ``reshape(x, s...)``.

Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft as sfft

F32, F64, C64, C128 = np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.complex64), np.dtype(np.complex128)
_PREC = {F32: 0, C64: 0, F64: 1, C128: 1}


class ParException(Exception):
    """src/ParCommon.jl:4-6"""


def is_complex(T) -> bool:
    return np.dtype(T).kind == "c"


def real_of(T):
    return {F32: F32, C64: F32, F64: F64, C128: F64}[np.dtype(T)]


def complex_of(T):
    return {F32: C64, C64: C64, F64: C128, C128: C128}[np.dtype(T)]


def subset_type(a, b):
    """src/ParCommon.jl:23-37 -- real beats complex, lower precision beats higher.

    The reference only defines the method for (lower, higher) argument order; the
    total symmetric function below agrees with it wherever it is defined.
    """
    a, b = np.dtype(a), np.dtype(b)
    prec = min(_PREC[a], _PREC[b])
    cplx = is_complex(a) and is_complex(b)
    return [[F32, C64], [F64, C128]][prec][int(cplx)]


def F(a: np.ndarray) -> np.ndarray:
    return np.asfortranarray(a)


def jreshape(x: np.ndarray, *shape) -> np.ndarray:
    """Julia ``reshape`` (column-major)."""
    return np.reshape(x, shape, order="F")


# ---------------------------------------------------------------------------
# src/ParCommon.jl
# ---------------------------------------------------------------------------
def local_size(global_size: int, rank: int, num_ranks: int) -> int:
    """src/ParCommon.jl:57-64, src/ParDistributed.jl:7-14"""
    r = global_size % num_ranks
    s = global_size // num_ranks
    if rank < r:
        s += 1
    return s


def range_axes(dims: Sequence[int], shape: Sequence[int]) -> List[List[Tuple[int, int]]]:
    """src/ParCommon.jl:45-55.  Ranges are 1-based inclusive ``(start, stop)``."""
    assert len(dims) == len(shape)
    axes = []
    for d in range(len(dims)):
        sizes = [local_size(shape[d], i, dims[d]) for i in range(dims[d])]
        offs = [0] + list(np.cumsum(sizes[:-1]))
        axes.append([(int(o) + 1, int(o) + s) for o, s in zip(offs, sizes)])
    return axes


def rotate_dims_batched(x: np.ndarray, rot: int) -> np.ndarray:
    """src/ParCommon.jl:66-77 -- a *materialised* permutedims (the bottleneck)."""
    n = x.ndim
    perm = list(np.roll(np.arange(n - 1), rot)) + [n - 1]
    return F(np.transpose(x, perm))


def as_matrix(x: np.ndarray) -> np.ndarray:
    """src/ParCommon.jl:79-82"""
    s = x.shape
    return jreshape(x, s[0], int(np.prod(s[1:], dtype=np.int64)))


# ---------------------------------------------------------------------------
# operator hierarchy (src/ParOperator.jl:63-124).  ``apply(x, theta)`` is the
# callable-struct apply on an AbstractMatrix; vectors go through reshape
# (e.g. src/ParDFT.jl:28, src/ParKron.jl:219-220).
# ---------------------------------------------------------------------------
class ParOperator:
    D = None  # domain dtype
    R = None  # range dtype
    linear = True
    parametric = False

    @property
    def Domain(self) -> int:
        raise ParException("Unimplemented")

    @property
    def Range(self) -> int:
        raise ParException("Unimplemented")

    def children(self):
        return []

    def adjoint(self):
        return ParAdjoint(self)

    def apply(self, x: np.ndarray, theta: Optional[dict] = None) -> np.ndarray:
        raise ParException("Unimplemented")

    def __call__(self, x, theta=None):
        """``A*x`` for vector or matrix ``x`` (src/ParOperator.jl:221-223)."""
        x = np.asarray(x)
        if x.ndim == 1:
            return self.apply(jreshape(x, x.shape[0], 1), theta).reshape(-1, order="F")
        return self.apply(F(x), theta)

    def init(self, theta: dict, rng: np.random.Generator):
        """src/ParOperator.jl:182-189"""
        for c in self.children():
            c.init(theta, rng)
        return theta


def init(*ops, seed=2) -> dict:
    """src/ParOperator.jl:194-200 (host RNG replaced by a seeded PCG64)."""
    rng = np.random.default_rng(seed)
    d: dict = {}
    for A in ops:
        A.init(d, rng)
    return d


def _rand(rng, T, *shape):
    """Julia ``rand(T, shape...)``: U[0,1) (re and im independently for complex)."""
    T = np.dtype(T)
    if is_complex(T):
        r = rng.random(shape) + 1j * rng.random(shape)
    else:
        r = rng.random(shape)
    return F(r.astype(T))


class ParAdjoint(ParOperator):
    """src/ParAdjoint.jl:6-19"""

    def __init__(self, op):
        self.op = op
        self.D, self.R = op.R, op.D
        self.parametric = op.parametric

    @property
    def Domain(self):
        return self.op.Range

    @property
    def Range(self):
        return self.op.Domain

    def children(self):
        return [self.op]

    def adjoint(self):
        return self.op

    def apply(self, x, theta=None):
        return self.op.apply_adjoint(x, theta)


class ParIdentity(ParOperator):
    """src/ParIdentity.jl:6-22"""

    def __init__(self, T, n):
        self.D = self.R = np.dtype(T)
        self.n = int(n)

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.n)

    def adjoint(self):
        return self

    def apply(self, x, theta=None):
        return x


class ParDFT(ParOperator):
    """src/ParDFT.jl:6-32.  NB the "adjoint" is the *inverse* transform."""

    def __init__(self, T, n):
        T = np.dtype(T)
        self.n = int(n)
        if not is_complex(T):
            self.D, self.R, self.m = T, complex_of(T), n // 2 + 1
        else:
            self.D, self.R, self.m = T, T, n

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.m)

    def apply(self, x, theta=None):
        if is_complex(self.D):
            if 0 in x.shape:
                return x
            return F(sfft.fft(x, axis=0))  # src/ParDFT.jl:26
        return F(sfft.rfft(x, axis=0))  # src/ParDFT.jl:27

    def apply_adjoint(self, x, theta=None):
        if is_complex(self.D):
            if 0 in x.shape:
                return x
            return F(sfft.ifft(x, axis=0))  # src/ParDFT.jl:30
        return F(sfft.irfft(x, n=self.n, axis=0))  # src/ParDFT.jl:31


class ParRestriction(ParOperator):
    """src/ParRestriction.jl:3-43.  ``ranges`` are 1-based inclusive (start, stop)."""

    def __init__(self, T, n, ranges):
        self.D = self.R = np.dtype(T)
        self.n = int(n)
        self.ranges = [(int(a), int(b)) for a, b in ranges]

    Domain = property(lambda s: s.n)
    Range = property(lambda s: sum(b - a + 1 for a, b in s.ranges))

    def apply(self, x, theta=None):
        b = x.shape[1]
        y = np.zeros((self.Range, b), dtype=x.dtype, order="F")
        off = 0
        for a, e in self.ranges:
            l = e - a + 1
            y[off:off + l, :] = x[a - 1:e, :]
            off += l
        return y

    def apply_adjoint(self, y, theta=None):
        b = y.shape[1]
        x = np.zeros((self.n, b), dtype=y.dtype, order="F")
        off = 0
        for a, e in self.ranges:
            l = e - a + 1
            x[a - 1:e, :] = y[off:off + l, :]  # overwrite on overlap
            off += l
        return x


class ParMatrix(ParOperator):
    """src/ParMatrix.jl:6-60"""
    parametric = True

    def __init__(self, T, m, n):
        self.D = self.R = np.dtype(T)
        self.m, self.n = int(m), int(n)

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.m)

    def init(self, theta, rng):
        T = self.D
        if self.n == 1:  # bias
            theta[self] = np.zeros((self.m, self.n), dtype=T, order="F")
        elif not is_complex(T):  # src/ParMatrix.jl:23-32
            scale = np.float32(math.sqrt(np.float32(24.0) / (self.m + self.n)))
            w = (_rand(rng, T, self.n, self.m) - T.type(0.5)) * T.type(scale)
            theta[self] = F(w.T.astype(T))
        else:  # src/ParMatrix.jl:33-40
            w = _rand(rng, T, self.n, self.m) / real_of(T).type(math.sqrt(self.m * self.n))
            theta[self] = F(w.T.astype(T))  # permutedims, not adjoint: no conj
        return theta

    def apply(self, x, theta=None):
        return F(theta[self] @ x)  # src/ParMatrix.jl:42-43

    def apply_adjoint(self, x, theta=None):
        return F(theta[self].conj().T @ x)  # src/ParMatrix.jl:45-46

    def bias_add(self, x, theta):
        """``x + A`` (src/ParMatrix.jl:53-55): broadcast add of the m x 1 parameter."""
        return x + theta[self]


class ParDiagonal(ParOperator):
    """src/ParDiagonal.jl:6-31"""
    parametric = True

    def __init__(self, T, n):
        self.D = self.R = np.dtype(T)
        self.n = int(n)

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.n)

    def init(self, theta, rng):
        theta[self] = _rand(rng, self.D, self.n)  # src/ParDiagonal.jl:20-22
        return theta

    def apply(self, x, theta=None):
        return F(theta[self][:, None] * x)

    def apply_adjoint(self, x, theta=None):
        return F(np.conj(theta[self])[:, None] * x)


class ParTensor(ParOperator):
    """src/ParTensor.jl:6-113 (ranks 4, 5, 6 only, as in the reference)."""
    parametric = True

    def __init__(self, T, weight_order, weight_shape, input_order, input_shape, target_order, target_shape, id=None):
        self.D = self.R = np.dtype(T)
        self.weight_order, self.weight_shape = tuple(weight_order), tuple(int(v) for v in weight_shape)
        self.input_order, self.input_shape = tuple(input_order), tuple(int(v) for v in input_shape)
        self.target_order, self.target_shape = tuple(target_order), tuple(int(v) for v in target_shape)
        self.id = id

    Domain = property(lambda s: int(np.prod(s.input_shape)))
    Range = property(lambda s: int(np.prod(s.target_shape)))

    def init(self, theta, rng):
        ws = self.weight_shape
        theta[self] = F(_rand(rng, self.D, *ws) / real_of(self.D).type(math.sqrt(np.prod(ws))))
        return theta

    def apply(self, x, theta=None):
        W = theta[self]
        b = x.shape[1]
        ws = self.weight_shape
        N = len(ws)
        if N == 4:  # src/ParTensor.jl:35-58   weights (ic, oc, nt, nxy)
            ic, oc = ws[0], ws[1]
            xi = jreshape(x, *self.input_shape, b)
            xi = F(np.transpose(xi, (3, 0, 1, 2)))
            xi = jreshape(xi, b, ic, -1)
            P = jreshape(W, ic, oc, -1)
            rest = ws[2:]
        elif N in (5, 6):  # src/ParTensor.jl:60-113   weights (oc, ic, n...)
            oc, ic = ws[0], ws[1]
            xi = jreshape(x, *self.input_shape, b)
            xi = F(np.transpose(xi, (N - 1,) + tuple(range(N - 1))))
            xi = jreshape(xi, b, ic, -1)
            P = F(np.transpose(W, (1, 0) + tuple(range(2, N))))
            P = jreshape(P, ic, oc, -1)
            rest = ws[2:]
        else:
            raise ParException("ParTensor apply is only defined for ranks 4, 5 and 6")
        # NNlib.batched_mul over the trailing (mode) dim: out[b,o,m] = sum_i x[b,i,m] P[i,o,m]
        out = F(np.einsum("bim,iom->bom", xi, P).astype(x.dtype))
        out = jreshape(out, b, oc, *rest)
        out = F(np.transpose(out, tuple(range(1, out.ndim)) + (0,)))
        return jreshape(out, -1, b)


class ParCompose(ParOperator):
    """src/ParCompose.jl:6-56"""

    def __new__(cls, *ops):
        if len(ops) == 1:
            return ops[0]
        return super().__new__(cls)

    def __init__(self, *ops):
        if len(ops) == 1:
            return
        flat = []
        for o in ops:
            flat.extend(o.ops if isinstance(o, ParCompose) else [o])  # src/ParCompose.jl:31-34
        self.ops = flat
        for i in range(len(flat) - 1):
            assert flat[i].Domain == flat[i + 1].Range, "ParCompose: Domain/Range mismatch"
            assert flat[i].D == flat[i + 1].R, "ParCompose: DDT/RDT mismatch"
        self.D, self.R = flat[-1].D, flat[0].R
        self.linear = all(o.linear for o in flat)
        self.parametric = any(o.parametric for o in flat)

    Domain = property(lambda s: s.ops[-1].Domain)
    Range = property(lambda s: s.ops[0].Range)

    def children(self):
        return self.ops

    def adjoint(self):
        return ParCompose(*[o.adjoint() for o in reversed(self.ops)])  # src/ParCompose.jl:42

    def apply(self, x, theta=None):
        for op in reversed(self.ops):  # src/ParCompose.jl:44-56
            x = op.apply(x, theta)
        return x


def kron_order(ops) -> List[int]:
    """Application-order heuristic of the ParKron constructor, src/ParKron.jl:28-57.

    Returns 1-based operator indices.
    """
    N = len(ops)
    DDTs = [o.D for o in ops]
    RDTs = [o.R for o in ops]
    T = DDTs[0]
    for t in DDTs[1:]:
        T = subset_type(T, t)
    order: List[int] = []
    for _ in range(N):
        cand = [j for j in range(N) if DDTs[j] == T and (j + 1) not in order]
        if not cand:
            raise ParException("ParKron: no applicable operator for domain type %s" % T)
        Rt = RDTs[cand[0]]
        for j in cand[1:]:
            Rt = subset_type(Rt, RDTs[j])
        cand = [j for j in cand if RDTs[j] == Rt]
        diffs = [ops[j].Domain - ops[j].Range for j in cand]
        md = max(diffs)
        cand = [j for j, d in zip(cand, diffs) if d == md]
        order.append(cand[-1] + 1)
        T = RDTs[cand[-1]]
    return order


class ParKron(ParOperator):
    """src/ParKron.jl:14-97, 190-220"""

    def __new__(cls, *ops, order=None):
        if len(ops) == 1:
            return ops[0]
        return super().__new__(cls)

    def __init__(self, *ops, order=None):
        if len(ops) == 1:
            return
        self.ops = flat = list(ops)  # flattening happens in kron/⊗ (src/ParKron.jl:71-75); see ``kron``
        self.order = list(order) if order is not None else kron_order(flat)
        self.D = flat[self.order[0] - 1].D
        self.R = flat[self.order[-1] - 1].R
        self.parametric = any(o.parametric for o in flat)

    Domain = property(lambda s: int(np.prod([o.Domain for o in s.ops], dtype=np.int64)))
    Range = property(lambda s: int(np.prod([o.Range for o in s.ops], dtype=np.int64)))

    def children(self):
        return self.ops

    def adjoint(self):
        # src/ParKron.jl:82 -- child adjoints, reversed order
        return ParKron(*[o.adjoint() for o in self.ops], order=list(reversed(self.order)))

    def apply(self, x, theta=None):
        """src/ParKron.jl:190-217"""
        N = len(self.ops)
        b = x.shape[1]
        s = [o.Domain for o in self.ops][::-1]
        x = jreshape(x, *s, b)
        x = rotate_dims_batched(x, -(N - self.order[0]))
        for i in range(N):
            o = self.order[i]
            s = x.shape
            x = as_matrix(x)
            Ai = self.ops[o - 1]
            x = Ai.apply(x, theta)
            x = jreshape(x, Ai.Range, *s[1:])
            if i < N - 1:
                x = rotate_dims_batched(x, N - o - (N - self.order[i + 1]))
            else:
                x = rotate_dims_batched(x, -o)
        return jreshape(x, x.size // b, b)


def kron(A, B):
    """``A ⊗ B`` with flattening, src/ParKron.jl:71-75; ``I ⊗ I`` merge, src/ParIdentity.jl:35."""
    if isinstance(A, ParIdentity) and isinstance(B, ParIdentity) and A.D == B.D:
        return ParIdentity(A.D, A.n * B.n)
    a = A.ops if isinstance(A, ParKron) else [A]
    b = B.ops if isinstance(B, ParKron) else [B]
    return ParKron(*a, *b)


# ---------------------------------------------------------------------------
# distribution (src/ParDistributed.jl, src/ParBroadcasted.jl, src/ParRepartition.jl)
# MPI is emulated: a cartesian communicator over P ranks is (dims, rank) with
# row-major rank<->coords (MPI_Cart_create with reorder=false; last dim fastest).
# ---------------------------------------------------------------------------
def cart_coords(dims: Sequence[int], rank: int) -> List[int]:
    c = []
    for d in reversed(dims):
        c.append(rank % d)
        rank //= d
    return c[::-1]


def cart_rank(dims: Sequence[int], coords: Sequence[int]) -> int:
    r = 0
    for d, c in zip(dims, coords):
        r = r * d + c
    return r


class ParDistributed(ParOperator):
    """src/ParDistributed.jl:16-29; only ever applied for identity (src/ParIdentity.jl:20-22)."""

    def __init__(self, op, rank, size):
        self.op, self.rank, self.size = op, int(rank), int(size)
        self.D, self.R = op.D, op.R
        self.parametric = op.parametric

    Domain = property(lambda s: local_size(s.op.Domain, s.rank, s.size))
    Range = property(lambda s: local_size(s.op.Range, s.rank, s.size))

    def children(self):
        return [self.op]

    def adjoint(self):
        return ParDistributed(self.op.adjoint(), self.rank, self.size)

    def apply(self, x, theta=None):
        if not isinstance(self.op, ParIdentity):
            raise ParException("ParDistributed apply is only defined for ParIdentity")
        return x


class ParBroadcasted(ParOperator):
    """src/ParBroadcasted.jl:3-57.  In the single-process emulation every rank sees the
    same theta dict, which is what ``MPI.bcast`` from root produces (:26-31)."""

    def __init__(self, op, root=0):
        self.op, self.root = op, root
        self.D, self.R = op.D, op.R
        self.parametric = op.parametric

    Domain = property(lambda s: s.op.Domain)
    Range = property(lambda s: s.op.Range)

    def children(self):
        return [self.op]

    def adjoint(self):
        return ParBroadcasted(self.op.adjoint(), self.root)

    def apply(self, x, theta=None):
        return self.op.apply(x, theta)


def repartition_plan(dims_in, dims_out, global_size, rank):
    """Send/recv box plan of the ParRepartition constructor, src/ParRepartition.jl:14-110.

    Returns ``(local_size_in, local_size_out, send, recv)`` where ``send``/``recv`` are
    ordered lists of ``(peer_rank, ((start, stop), ...))`` with 1-based inclusive local
    index ranges, in the reference's iteration order (``Iterators.product``: first axis
    fastest).
    """
    N = len(dims_in)
    assert N == len(dims_out) == len(global_size)
    coords_in, coords_out = cart_coords(dims_in, rank), cart_coords(dims_out, rank)
    ax_in, ax_out = range_axes(dims_in, global_size), range_axes(dims_out, global_size)
    mine_in = [ax_in[i][coords_in[i]] for i in range(N)]
    mine_out = [ax_out[i][coords_out[i]] for i in range(N)]
    lsi = tuple(e - s + 1 for s, e in mine_in)
    lso = tuple(e - s + 1 for s, e in mine_out)

    def isect(q, r):
        s, e = max(q[0], r[0]), min(q[1], r[1])
        return (s, e) if e >= s else None

    def overlaps(mine, axes):
        out, ok = [], True
        for i in range(N):
            hits = [j for j, r in enumerate(axes[i]) if isect(mine[i], r) is not None]
            if hits:
                out.append(range(hits[0], hits[-1] + 1))  # findfirst:findlast
            else:
                ok = False
        return out, ok

    def build(mine, axes_other, dims_other):
        ov, ok = overlaps(mine, axes_other)
        res = []
        if ok:
            # Iterators.product: first axis varies fastest
            for idx in itertools.product(*[list(r) for r in reversed(ov)]):
                idx = idx[::-1]
                peer = cart_rank(dims_other, idx)
                other = [axes_other[i][idx[i]] for i in range(N)]
                box = []
                for q, r in zip(mine, other):
                    s, e = max(q[0], r[0]), min(q[1], r[1])  # intersect(q, r) (possibly empty)
                    box.append((s - q[0] + 1, e - q[0] + 1))
                res.append((peer, tuple(box)))
        return res

    send = build(mine_in, ax_out, dims_out)
    recv = build(mine_out, ax_in, dims_in)
    return lsi, lso, send, recv


class ParRepartition(ParOperator):
    """src/ParRepartition.jl:3-197, emulated over P in-process ranks.

    ``apply_all(xs)`` takes the list of every rank's local ``(prod(local_in), b)`` matrix
    and returns every rank's ``(prod(local_out), b)`` matrix, moving exactly the boxes the
    reference's Isend/Irecv! plan moves (pack = ``vec(view(x, box..., :))``).
    """

    def __init__(self, T, dims_in, dims_out, global_size, adjoint_of=None):
        self.D = self.R = np.dtype(T)
        self.dims_in, self.dims_out = list(dims_in), list(dims_out)
        self.global_size = tuple(int(g) for g in global_size)
        self.P = int(np.prod(dims_in))
        assert self.P == int(np.prod(dims_out))
        self.plans = [repartition_plan(self.dims_in, self.dims_out, self.global_size, r) for r in range(self.P)]

    def local_in(self, rank):
        return self.plans[rank][0]

    def local_out(self, rank):
        return self.plans[rank][1]

    def adjoint(self):
        return ParRepartition(self.D, self.dims_out, self.dims_in, self.global_size)  # :120-131 swaps the plans

    def apply_all(self, xs: List[np.ndarray]) -> List[np.ndarray]:
        b = xs[0].shape[1]
        ys = []
        for r in range(self.P):
            ys.append(np.zeros(self.local_out(r) + (b,), dtype=xs[r].dtype, order="F"))
        n_msgs = 0
        for src in range(self.P):
            lsi, _, send, _ = self.plans[src]
            x = jreshape(xs[src], *lsi, b)
            for dst, box in send:
                sl = tuple(slice(s - 1, e) for s, e in box) + (slice(None),)
                buf = x[sl].reshape(-1, order="F")  # pack, column-major, batch slowest (:174)
                # find matching recv box on dst
                rbox = dict(self.plans[dst][3])[src]
                rsl = tuple(slice(s - 1, e) for s, e in rbox) + (slice(None),)
                tgt = ys[dst][rsl]
                ys[dst][rsl] = buf.reshape(tgt.shape, order="F")
                n_msgs += int(dst != src)
        self.last_message_count = n_msgs
        return [jreshape(y, -1, b) for y in ys]


def distribute_kron(A: ParKron, dims_in, dims_out=None):
    """``distribute(A::ParKron, dims_in, dims_out)``, src/ParKron.jl:244-309, for every rank.

    Returns a list (one entry per pipeline stage, in *application* order) of either
    ``("local", [op_rank0, op_rank1, ...])`` or ``("repartition", ParRepartition)``.
    """
    dims = list(dims_in)
    dims_out = list(dims_out) if dims_out is not None else list(dims_in)
    N = len(dims)
    assert len(A.ops) == N
    P = int(np.prod(dims))
    size_curr = [o.Domain for o in reversed(A.ops)]
    dims_prev = list(dims)
    stages = []
    for i in range(N):
        o = A.order[i]
        d = N - o + 1  # 1-based array dim
        Ai = A.ops[o - 1]
        if isinstance(Ai, ParIdentity):
            continue
        dims_i = list(dims)
        dims_i[d - 1] = 1
        dims_i[(d % N)] *= dims[d - 1]  # mod1(d+1, N) -> 0-based index d % N
        if dims_prev != dims_i:
            stages.append(("repartition", ParRepartition(Ai.D, dims_prev, dims_i, tuple(size_curr))))
        per_rank = []
        for r in range(P):
            coords = cart_coords(dims_i, r)
            lower = [ParDistributed(ParIdentity(Ai.D, size_curr[j]), coords[j], dims_i[j]) for j in range(d, N)]
            upper = [ParDistributed(ParIdentity(Ai.D, size_curr[j]), coords[j], dims_i[j]) for j in range(0, d - 1)]
            # pushfirst! reverses: slowest dim first
            per_rank.append(ParKron(*lower[::-1], ParBroadcasted(Ai), *upper[::-1]))
        stages.append(("local", per_rank))
        size_curr[d - 1] = Ai.Range
        dims_prev = dims_i
    if dims_prev != dims_out:
        stages.append(("repartition", ParRepartition(A.ops[A.order[-1] - 1].R, dims_prev, dims_out, tuple(size_curr))))
    return stages


def apply_distributed(stages, xs: List[np.ndarray], theta=None) -> List[np.ndarray]:
    """SPMD lock-step emulation of the ``ParCompose`` returned by ``distribute`` (src/ParKron.jl:308)."""
    for kind, obj in stages:
        if kind == "repartition":
            xs = obj.apply_all(xs)
        else:
            xs = [op.apply(x, theta) for op, x in zip(obj, xs)]
    return xs


def scatter_global(x_global: np.ndarray, dims, global_size) -> List[np.ndarray]:
    """Split a global ``(prod(global_size), b)`` matrix into per-rank local matrices (test helper)."""
    b = x_global.shape[1]
    xg = jreshape(x_global, *global_size, b)
    axes = range_axes(dims, global_size)
    out = []
    for r in range(int(np.prod(dims))):
        c = cart_coords(dims, r)
        sl = tuple(slice(axes[i][c[i]][0] - 1, axes[i][c[i]][1]) for i in range(len(dims))) + (slice(None),)
        out.append(jreshape(F(xg[sl]), -1, b))
    return out


def gather_global(xs: List[np.ndarray], dims, global_size) -> np.ndarray:
    b = xs[0].shape[1]
    xg = np.zeros(tuple(global_size) + (b,), dtype=xs[0].dtype, order="F")
    axes = range_axes(dims, global_size)
    for r in range(int(np.prod(dims))):
        c = cart_coords(dims, r)
        sl = tuple(slice(axes[i][c[i]][0] - 1, axes[i][c[i]][1]) for i in range(len(dims))) + (slice(None),)
        xg[sl] = jreshape(xs[r], *xg[sl].shape)
    return jreshape(xg, -1, b)


# ---------------------------------------------------------------------------
# examples/fno.jl
# ---------------------------------------------------------------------------
def spectral_convolution(T, shape_spacetime, modes_spacetime, lifted_dim):
    """examples/fno.jl:9-32.  ``modes_spacetime`` is a list (per dim) of lists of 1-based
    inclusive ``(start, stop)`` ranges.  Returns ``(S, parts)``."""
    T = np.dtype(T)
    CT = complex_of(T)
    dft_time = ParDFT(T, shape_spacetime[-1])
    dft_space = ParKron(*[ParDFT(CT, s) for s in shape_spacetime[:-1]])
    dft_spacetime = kron(dft_time, dft_space)
    st_ops = dft_spacetime.ops if isinstance(dft_spacetime, ParKron) else [dft_spacetime]
    restrict = ParKron(*[ParRestriction(CT, Fop.Range, m) for Fop, m in zip(st_ops, reversed(modes_spacetime))])
    identity_channels = ParIdentity(T, lifted_dim)
    dft_restrict = ParCompose(restrict, dft_spacetime)
    dft_all = kron(dft_restrict, identity_channels)
    mix_channels = ParMatrix(CT, lifted_dim, lifted_dim)
    elementwise_weighting = ParDiagonal(CT, dft_restrict.Range)
    frequency_weight = kron(elementwise_weighting, mix_channels)
    S = ParCompose(dft_all.adjoint(), frequency_weight, dft_all)
    return S, dict(dft_all=dft_all, frequency_weight=frequency_weight, mix_channels=mix_channels,
                   elementwise_weighting=elementwise_weighting, dft_restrict=dft_restrict)


def channel_mixing(T, shape_spacetime, lifted_dim):
    """examples/fno.jl:34-38"""
    mix = ParMatrix(T, lifted_dim, lifted_dim)
    ident = ParIdentity(T, int(np.prod(shape_spacetime)))
    return kron(ident, mix)


def gelu(x):
    """NNlib.gelu (tanh approximation): 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))."""
    x = np.asarray(x)
    c = x.dtype.type(math.sqrt(2.0 / math.pi))
    return (x.dtype.type(0.5) * x * (1 + np.tanh(c * (x + x.dtype.type(0.044715) * x ** 3)))).astype(x.dtype)


def fno_block(x, S, W, theta):
    """examples/fno.jl:40-44"""
    return gelu(S.apply(x, theta) + W.apply(x, theta))


# ---------------------------------------------------------------------------
# true (mathematical) adjoints used by the Zygote/ChainRules pullbacks.
# ``pullback(op, x, ybar, theta)`` returns (xbar, {param_op: grad}).
# ---------------------------------------------------------------------------
def _vjp_leaf(op, x, ybar, theta, grads):
    """Cotangent of x for y = op(x); accumulates parameter cotangents into ``grads``."""
    if isinstance(op, ParAdjoint):
        inner = op.op
        if isinstance(inner, ParDFT):
            n = inner.n
            if is_complex(inner.D):  # y = ifft(x): xbar = fft(ybar)/n  (AbstractFFTs rule)
                return F(sfft.fft(ybar, axis=0) / n).astype(x.dtype)
            # y = irfft(x, n): xbar_k = s_k/n * rfft(ybar)_k, s = 1 at DC/Nyquist else 2
            s = np.full(inner.m, 2.0)
            s[0] = 1.0
            if n % 2 == 0:
                s[-1] = 1.0
            return F(sfft.rfft(ybar, axis=0) * (s[:, None] / n)).astype(x.dtype)
        if isinstance(inner, ParRestriction):
            return inner.apply(ybar)  # src/ParRestriction.jl:64-71
        if isinstance(inner, ParMatrix):  # y = W' x : xbar = W ybar ; Wbar = x ybar'
            W = theta[inner]
            grads[inner] = grads.get(inner, 0) + F(x @ ybar.conj().T)
            return F(W @ ybar)
        if isinstance(inner, ParDiagonal):  # y = conj(d) .* x
            d = theta[inner]
            g = np.sum(x * np.conj(ybar), axis=1)
            grads[inner] = grads.get(inner, 0) + (g if is_complex(inner.D) else g.real).astype(inner.D)
            return F(d[:, None] * ybar)
        raise ParException("no pullback for adjoint of %s" % type(inner).__name__)
    if isinstance(op, (ParIdentity, ParDistributed)):
        return ybar
    if isinstance(op, ParBroadcasted):
        return _vjp_leaf(op.op, x, ybar, theta, grads)
    if isinstance(op, ParDFT):
        n = op.n
        if is_complex(op.D):  # y = fft(x): xbar = bfft(ybar)
            return F(sfft.ifft(ybar, axis=0) * n).astype(x.dtype)
        # y = rfft(x), x real: xbar = Re(W_r^H ybar) = sum_k Re(conj(w_kj) ybar_k)
        k = np.arange(op.m)[:, None]
        j = np.arange(n)[None, :]
        Wr = np.exp(-2j * np.pi * k * j / n)
        return F((Wr.conj().T @ ybar.astype(np.complex128)).real).astype(x.dtype)
    if isinstance(op, ParRestriction):
        return op.apply_adjoint(ybar)  # src/ParRestriction.jl:57-63
    if isinstance(op, ParMatrix):  # y = W x : xbar = W' ybar ; Wbar = ybar x'
        W = theta[op]
        grads[op] = grads.get(op, 0) + F(ybar @ x.conj().T)
        return F(W.conj().T @ ybar)
    if isinstance(op, ParDiagonal):  # y = d .* x : xbar = conj(d) .* ybar ; dbar = sum_b conj(x) .* ybar
        d = theta[op]
        g = np.sum(np.conj(x) * ybar, axis=1)
        grads[op] = grads.get(op, 0) + (g if is_complex(op.D) else g.real).astype(op.D)
        return F(np.conj(d)[:, None] * ybar)
    if isinstance(op, ParTensor):
        W = theta[op]
        b = x.shape[1]
        ws = op.weight_shape
        N = len(ws)
        if N == 4:
            ic, oc = ws[0], ws[1]
            Wm = jreshape(W, ic, oc, -1)  # [i, o, m]
            sub = "iom"
        else:
            oc, ic = ws[0], ws[1]
            Wm = jreshape(W, oc, ic, -1)  # [o, i, m]
            sub = "oim"
        xm = jreshape(x, ic, -1, b)  # [i, m, b]
        ym = jreshape(ybar, oc, -1, b)  # [o, m, b]
        xbar = np.einsum(sub + ",omb->imb", np.conj(Wm), ym)
        gW = np.einsum("imb,omb->" + sub, np.conj(xm), ym)
        grads[op] = grads.get(op, 0) + jreshape(F(gW.astype(op.D)), *ws)
        return jreshape(F(xbar.astype(x.dtype)), -1, b)
    raise ParException("no pullback for %s" % type(op).__name__)


def _forward_tape(op, x, theta, tape):
    """Forward pass that records (leaf, input) pairs; structure-only nodes recurse."""
    if isinstance(op, ParCompose):
        for c in reversed(op.ops):
            x = _forward_tape(c, x, theta, tape)
        return x
    if isinstance(op, ParKron):
        N = len(op.ops)
        b = x.shape[1]
        sizes = [o.Domain for o in op.ops][::-1]  # fastest first
        cur = jreshape(x, *sizes, b)
        for o in op.order:
            d = N - o  # 0-based array dim
            Ai = op.ops[o - 1]
            moved = F(np.moveaxis(cur, d, 0))
            shp = moved.shape
            sub_tape = []
            ym = _forward_tape(Ai, as_matrix(moved), theta, sub_tape)
            tape.append(("kron_mode", d, shp, Ai, sub_tape))
            cur = F(np.moveaxis(jreshape(ym, Ai.Range, *shp[1:]), 0, d))
        return jreshape(cur, -1, b)
    tape.append(("leaf", op, x))
    return op.apply(x, theta)


def _backward_tape(tape, ybar, theta, grads):
    for entry in reversed(tape):
        if entry[0] == "leaf":
            _, op, x = entry
            ybar = _vjp_leaf(op, x, ybar, theta, grads)
        else:
            _, d, shp, Ai, sub_tape = entry
            b = ybar.shape[1]
            # current natural layout has Range(Ai) at position d
            nat = list(shp[1:])
            nat.insert(d, Ai.Range)
            cur = jreshape(ybar, *nat)
            moved = as_matrix(F(np.moveaxis(cur, d, 0)))
            xb = _backward_tape(sub_tape, moved, theta, grads)
            xb = jreshape(xb, *shp)
            ybar = jreshape(F(np.moveaxis(xb, 0, d)), -1, b)
    return ybar


def pullback(op, x, ybar, theta=None):
    """Reverse-mode cotangents of ``y = op(theta) * x`` w.r.t. x and every parameter
    (what Zygote produces for the reference): true adjoints for the DFTs,
    ``W̄ = ȳ x'`` for dense factors, ``d̄ = Σ_b conj(x)⊙ȳ`` for diagonals."""
    x = F(np.asarray(x))
    ybar = F(np.asarray(ybar))
    tape: list = []
    _forward_tape(op, x, theta, tape)
    grads: dict = {}
    xbar = _backward_tape(tape, ybar, theta, grads)
    return xbar, grads
