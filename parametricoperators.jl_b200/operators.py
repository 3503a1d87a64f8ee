"""Host-side mirror of the ParametricOperators.jl operator API for the spectral apply path.

Julia is not available in the build environment, so the host layer the reference keeps in
Julia (type lattice, constructors, ``init``, ``A(θ)``, ``'``, ``⊗``, ``∘``) is mirrored here
in Python with the same names, argument meaning and error behaviour, and the same C ABI
underneath that ``julia/ParametricOperatorsB200.jl`` binds with ``ccall``.  Everything here
is bookkeeping: no arithmetic on tensor data happens on the host -- ``A * x`` lowers the
parameterised tree (``lowering.py``) and hands raw device pointers to ``libparops.so``.

Julia -> Python spelling
    A'            adjoint(A)  or  A.H
    A ⊗ B         kron(A, B)  or  A ^ B
    A ∘ B, A * B  compose(A, B)  or  A * B      (operator * operator)
    A * x, A(x)   A * x  or  A(x)               (operator * array)
    A(θ)          A(θ)                          (θ is the dict returned by init)
    init!(A, θ)   init_(A, θ)
"""
from __future__ import annotations

import math
import uuid
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import ParException

__all__ = [
    "ParException", "Float32", "Float64", "ComplexF32", "ComplexF64", "Complex", "real",
    "subset_type", "superset_type", "Linear", "NonLinear", "Parametric", "NonParametric", "Parameterized", "Internal",
    "External", "ParOperator", "DDT", "RDT", "Domain", "Range", "linearity", "parametricity", "ast_location", "children",
    "rebuild", "params", "adjoint", "init", "init_", "ParIdentity", "ParDFT", "ParRestriction", "ParMatrix",
    "ParDiagonal", "ParTensor", "ParAdjoint", "ParParameterized", "ParCompose", "ParKron", "compose", "kron", "order",
    "reorder", "ParBias", "ParGelu", "update_", "scale_", "cumranges",
]


# ---------------------------------------------------------------------------------------
# element types and the type lattice (src/ParCommon.jl:23-37)
# ---------------------------------------------------------------------------------------
class JType:
    def __init__(self, name, code, np_dtype, prec, is_complex):
        self.name, self.code, self.np = name, code, np.dtype(np_dtype)
        self.prec, self.is_complex = prec, is_complex

    def __repr__(self):
        return self.name

    @property
    def torch(self):
        t = self.__dict__.get("_torch")
        if t is None:
            import torch
            t = self.__dict__["_torch"] = {0: torch.float32, 1: torch.float64, 2: torch.complex64, 3: torch.complex128}[self.code]
        return t

    @property
    def itemsize(self):
        return self.np.itemsize


Float32 = JType("Float32", 0, np.float32, 0, False)
Float64 = JType("Float64", 1, np.float64, 1, False)
ComplexF32 = JType("ComplexF32", 2, np.complex64, 0, True)
ComplexF64 = JType("ComplexF64", 3, np.complex128, 1, True)
_TYPES = [Float32, Float64, ComplexF32, ComplexF64]


_JTYPE_OF = {}  # numpy / torch dtype object -> JType (the lookup below costs microseconds; applies of small operators are host-bound)


def jtype(t) -> JType:
    """Accept a JType, a numpy dtype or a torch dtype."""
    if isinstance(t, JType):
        return t
    try:
        hit = _JTYPE_OF.get(t)
    except TypeError:  # unhashable spec
        hit = None
    if hit is not None:
        return hit
    T = _jtype_slow(t)
    try:
        _JTYPE_OF[t] = T
    except TypeError:
        pass
    return T


def _jtype_slow(t) -> JType:
    try:
        d = np.dtype(t)
        for T in _TYPES:
            if T.np == d:
                return T
    except TypeError:
        pass
    for T in _TYPES:
        if str(t) in ("torch." + T.np.name,):
            return T
    raise ParException("Invalid element type %r" % (t,))


def Complex(T) -> JType:
    T = jtype(T)
    return [ComplexF32, ComplexF64][T.prec]


def real(T) -> JType:
    T = jtype(T)
    return [Float32, Float64][T.prec]


def subset_type(a, b) -> JType:
    """Real beats complex, lower precision beats higher (src/ParCommon.jl:27-30)."""
    a, b = jtype(a), jtype(b)
    prec = min(a.prec, b.prec)
    return [[Float32, ComplexF32], [Float64, ComplexF64]][prec][int(a.is_complex and b.is_complex)]


def superset_type(a, b) -> JType:
    """src/ParCommon.jl:31-34 (as written there the complex results keep the LOWER precision)."""
    a, b = jtype(a), jtype(b)
    if not a.is_complex and not b.is_complex:
        return [Float32, Float64][max(a.prec, b.prec)]
    return [ComplexF32, ComplexF64][min(a.prec, b.prec)]


def cumranges(x: Sequence[int]) -> List[Tuple[int, int]]:
    """src/ParCommon.jl:16-21 -- 1-based inclusive (start, stop) ranges between cumsums."""
    out, off = [], 0
    for v in x:
        out.append((off + 1, off + int(v)))
        off += int(v)
    return out


# ---------------------------------------------------------------------------------------
# type flags (src/ParOperator.jl:8-58)
# ---------------------------------------------------------------------------------------
class _Flag:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


Linear, NonLinear = _Flag("Linear"), _Flag("NonLinear")
Parametric, NonParametric, Parameterized = _Flag("Parametric"), _Flag("NonParametric"), _Flag("Parameterized")
Internal, External = _Flag("Internal"), _Flag("External")


def promote_linearity(a, b):
    return Linear if (a is Linear and b is Linear) else NonLinear


def promote_parametricity(a, b):
    """src/ParOperator.jl:38-42"""
    if a is NonParametric and b is NonParametric:
        return NonParametric
    if a in (NonParametric, Parameterized) and b in (NonParametric, Parameterized):
        return Parameterized
    return Parametric


def _fold(f, xs):
    acc = xs[0]
    for v in xs[1:]:
        acc = f(acc, v)
    return acc


def _as_range(r) -> Tuple[int, int]:
    """Accept (start, stop) 1-based inclusive, or a Python ``range`` over 1-based indices."""
    if isinstance(r, range):
        if r.step != 1:
            raise ParException("ParRestriction ranges must have unit step")
        return (r.start, r.stop - 1)
    a, b = r
    return (int(a), int(b))


# ---------------------------------------------------------------------------------------
# base operator (src/ParOperator.jl:63-228)
# ---------------------------------------------------------------------------------------
class ParOperator:
    """``ParOperator{D,R,L,P,T}``: attributes ``D, R`` (element types), ``L`` (linearity),
    ``P`` (parametricity), ``T`` (AST location)."""
    D: JType
    R: JType
    L = Linear
    P = NonParametric
    T = External

    # -- traits ---------------------------------------------------------------------
    @property
    def Domain(self) -> int:
        raise ParException("Unimplemented")

    @property
    def Range(self) -> int:
        raise ParException("Unimplemented")

    def children(self) -> list:
        if self.T is External:
            return []
        raise ParException("Unimplemented")

    def rebuild(self, cs):
        if self.T is External:
            return self
        raise ParException("Unimplemented")

    def adjoint(self):
        if self.L is not Linear:
            raise ParException("adjoint is only defined for linear operators")
        return ParAdjoint(self)

    @property
    def H(self):
        return self.adjoint()

    # -- init (src/ParOperator.jl:182-200) ---------------------------------------------
    def init_(self, d: dict, rng=None):
        if self.P is not Parametric:
            return None
        if self.T is Internal:
            for c in self.children():
                c.init_(d, rng)
            return d
        raise ParException("init! not implemented for %s" % type(self).__name__)

    # -- parameterise / apply -------------------------------------------------------------
    def parameterize(self, theta: dict):
        """``A(θ)`` (src/ParOperator.jl:205-206, src/ParParameterized.jl:22)."""
        if self.P is not Parametric:
            return self
        if self.T is External:
            if self not in theta:
                raise ParException("parameter dict has no entry for %r" % (self,))
            return ParParameterized(self, theta[self])
        return self.rebuild([c.parameterize(theta) if c.P is Parametric else c for c in self.children()])

    def __call__(self, arg):
        if isinstance(arg, dict):
            return self.parameterize(arg)
        return self.apply(arg)

    def apply(self, x):
        """``A*x`` / ``A(x)`` for a vector ``(Domain,)`` or column-major matrix ``(Domain, batch)``."""
        if self.P is Parametric:
            raise ParException("operator has unbound parameters: call A(θ) first (MethodError in the reference)")
        from .engine import apply_operator
        return apply_operator(self, x)

    def __mul__(self, other):
        if isinstance(other, ParOperator):
            return compose(self, other)  # src/ParCompose.jl:35
        return self.apply(other)

    def __rmul__(self, x):
        """``x * A`` = ``(A' * x')'`` (src/ParOperator.jl:228)."""
        if isinstance(x, ParOperator):
            return NotImplemented
        from .engine import apply_right
        return apply_right(self, x)

    def __truediv__(self, x):
        """``A / x`` = ``A.params ./ x`` for a parameterised ParMatrix (src/ParMatrix.jl:60)."""
        from .engine import divide
        return divide(self, x)

    def __xor__(self, other):
        return kron(self, other)

    # operators hash by identity (the uuid4 id of src/ParMatrix.jl:12 plays that role)
    __hash__ = object.__hash__

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return "%s{%s,%s}(%s->%s)" % (type(self).__name__, self.D, self.R, self.Domain, self.Range)


def DDT(A): return A.D
def RDT(A): return A.R
def Domain(A): return A.Domain
def Range(A): return A.Range
def linearity(A): return A.L
def parametricity(A): return A.P
def ast_location(A): return A.T
def children(A): return A.children()
def rebuild(A, cs): return A.rebuild(cs)
def adjoint(A): return A.adjoint()
def params(A): return A.params


def init_(A: ParOperator, d: dict, rng=None):
    """``init!(A, d)``"""
    return A.init_(d, _rng(rng))


def init(*ops: ParOperator, rng=None) -> dict:
    """``init(As...)`` (src/ParOperator.jl:194-200).  Host-side (NumPy) arrays, like the
    reference's CPU ``rand``; ``gpu(θ)`` moves them.  ``rng``: seed or numpy Generator."""
    d: dict = {}
    g = _rng(rng)
    for A in ops:
        A.init_(d, g)
    return d


_global_rng = np.random.default_rng()


def _rng(rng):
    if rng is None:
        return _global_rng
    if isinstance(rng, np.random.Generator):
        return rng
    return np.random.default_rng(rng)


def _rand(rng, T: JType, *shape):
    """Julia ``rand(T, shape...)``: U[0,1), independently for re and im; column-major."""
    if T.is_complex:
        r = rng.random(shape) + 1j * rng.random(shape)
    else:
        r = rng.random(shape)
    return np.asfortranarray(r.astype(T.np))


def update_(params_: dict, grads: dict):
    """``update!(params, grads)`` = ``mergewith!(-)`` (src/ParOperator.jl:158)."""
    for k, g in grads.items():
        if k in params_:
            params_[k] = params_[k] - g
        else:
            params_[k] = g
    return params_


def scale_(a, x):
    """``scale!(a, params)`` (src/ParOperator.jl:163-177)."""
    if isinstance(x, dict):
        for k in x:
            x[k] = scale_(a, x[k])
        return x
    if isinstance(x, (list, tuple)):
        return [scale_(a, v) for v in x]
    x *= a
    return x


# ---------------------------------------------------------------------------------------
# wrappers
# ---------------------------------------------------------------------------------------
class ParAdjoint(ParOperator):
    """Lazy adjoint (src/ParAdjoint.jl:6-19): ``ParLinearOperator{R,D,P,Internal}``."""
    T = Internal

    def __init__(self, op):
        self.op = op
        self.D, self.R = op.R, op.D
        self.P = op.P
        self.L = Linear

    Domain = property(lambda s: s.op.Range)
    Range = property(lambda s: s.op.Domain)

    def children(self):
        return [self.op]

    def rebuild(self, cs):
        return ParAdjoint(cs[0])

    def adjoint(self):
        return self.op

    def parameterize(self, theta):
        if self.P is not Parametric:
            return self
        # src/ParAdjoint.jl:19 -- keeps the WHOLE dict; the leaf is looked up at apply time
        return ParParameterized(self.op.adjoint(), theta)


class ParParameterized(ParOperator):
    """Lazy (op, params) pair (src/ParParameterized.jl:6-22)."""
    T = Internal
    P = Parameterized

    def __init__(self, op, params_):
        self.op = op
        self.params = params_
        self.D, self.R, self.L = op.D, op.R, op.L

    Domain = property(lambda s: s.op.Domain)
    Range = property(lambda s: s.op.Range)

    def children(self):
        return [self.op]

    def adjoint(self):
        return ParParameterized(self.op.adjoint(), self.params)  # src/ParParameterized.jl:16

    def leaf_and_tensor(self):
        """(leaf operator, adjoint flag, parameter array)."""
        if isinstance(self.op, ParAdjoint):
            leaf = self.op.op
            theta = self.params[leaf] if isinstance(self.params, dict) else self.params
            return leaf, True, theta
        return self.op, False, self.params


# ---------------------------------------------------------------------------------------
# leaves
# ---------------------------------------------------------------------------------------
class ParIdentity(ParOperator):
    """src/ParIdentity.jl:6-22"""

    def __init__(self, T=None, n=None):
        if n is None:
            T, n = Float64, T
        self.D = self.R = jtype(T)
        self.n = int(n)

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.n)

    def adjoint(self):
        return self


class ParDFT(ParOperator):
    """Fourier transform along dim 1; real types dispatch the real FFT (src/ParDFT.jl:6-32).
    NB: ``A'`` is the INVERSE transform (``ifft``/``irfft``), as in the reference."""

    def __init__(self, T=None, n=None):
        if n is None:
            T, n = ComplexF64, T  # ParDFT(n) = ParDFT(ComplexF64, n)
        T = jtype(T)
        self.n = int(n)
        if not T.is_complex:
            self.D, self.R, self.m = T, Complex(T), self.n // 2 + 1
        else:
            self.D, self.R, self.m = T, T, self.n

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.m)


class ParRestriction(ParOperator):
    """Row-range gather (src/ParRestriction.jl:3-43); ``ranges``: 1-based inclusive
    ``(start, stop)`` pairs (or Python ranges over 1-based indices)."""

    def __init__(self, T, n=None, ranges=None):
        if ranges is None:
            T, n, ranges = Float64, T, n
        self.D = self.R = jtype(T)
        self.n = int(n)
        self.ranges = [_as_range(r) for r in ranges]

    Domain = property(lambda s: s.n)
    Range = property(lambda s: sum(max(0, b - a + 1) for a, b in s.ranges))


class ParMatrix(ParOperator):
    """Dense m x n parametric operator (src/ParMatrix.jl:6-60)."""
    P = Parametric

    def __init__(self, *args):
        args = list(args)
        T = Float64
        if args and not isinstance(args[0], (int, np.integer)):
            T = args.pop(0)
        self.D = self.R = jtype(T)
        self.m, self.n = int(args[0]), int(args[1])
        self.id = args[2] if len(args) > 2 else uuid.uuid4()

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.m)

    def init_(self, d, rng=None):
        rng = _rng(rng)
        T = self.D
        if self.n == 1:  # bias: zeros (src/ParMatrix.jl:24-27)
            d[self] = np.zeros((self.m, self.n), dtype=T.np, order="F")
        elif not T.is_complex:  # Glorot-like, src/ParMatrix.jl:28-31
            scale = np.float32(math.sqrt(np.float32(24.0) / (self.m + self.n)))
            w = (_rand(rng, T, self.n, self.m) - T.np.type(0.5)) * T.np.type(scale)
            d[self] = np.asfortranarray(w.T.astype(T.np))
        else:  # src/ParMatrix.jl:33-40 (permutedims: transpose WITHOUT conj)
            w = _rand(rng, T, self.n, self.m) / real(T).np.type(math.sqrt(self.m * self.n))
            d[self] = np.asfortranarray(w.T.astype(T.np))
        return d

    def __getitem__(self, idx):
        """Sub-block operator (src/ParMatrix.jl:92-102): a NEW ParMatrix of the block's size."""
        rows, cols = idx

        def ln(sel, full):
            if isinstance(sel, slice):
                return len(range(*sel.indices(full)))
            if isinstance(sel, (int, np.integer)):
                return 1
            return len(sel)

        return ParMatrix(self.D, ln(rows, self.m), ln(cols, self.n))


class ParDiagonal(ParOperator):
    """Elementwise scaling (src/ParDiagonal.jl:6-31)."""
    P = Parametric

    def __init__(self, *args):
        args = list(args)
        T = Float64
        if args and not isinstance(args[0], (int, np.integer)):
            T = args.pop(0)
        self.D = self.R = jtype(T)
        self.n = int(args[0])
        self.id = args[1] if len(args) > 1 else uuid.uuid4()

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.n)

    def init_(self, d, rng=None):
        d[self] = _rand(_rng(rng), self.D, self.n)  # src/ParDiagonal.jl:20-22
        return d


class ParTensor(ParOperator):
    """Per-mode channel-mixing weights (src/ParTensor.jl:6-113).  Apply is defined for
    weight ranks 4 ``(ic, oc, nt, nxy)``, 5 ``(oc, ic, nx, ny, nt)`` and 6
    ``(oc, ic, nx, ny, nz, nt)`` exactly as in the reference."""
    P = Parametric

    def __init__(self, *args):
        args = list(args)
        T = Float64
        if args and isinstance(args[0], JType) or (args and not isinstance(args[0], (tuple, list))):
            T = args.pop(0)
        wo, ws, io, is_, to, ts = args[:6]
        self.D = self.R = jtype(T)
        self.weight_order, self.weight_shape = tuple(wo), tuple(int(v) for v in ws)
        self.input_order, self.input_shape = tuple(io), tuple(int(v) for v in is_)
        self.target_order, self.target_shape = tuple(to), tuple(int(v) for v in ts)
        self.id = args[6] if len(args) > 6 else uuid.uuid4()

    Domain = property(lambda s: int(np.prod(s.input_shape, dtype=np.int64)))
    Range = property(lambda s: int(np.prod(s.target_shape, dtype=np.int64)))

    def init_(self, d, rng=None):
        ws = self.weight_shape
        w = _rand(_rng(rng), self.D, *ws) / real(self.D).np.type(math.sqrt(float(np.prod(ws))))
        d[self] = np.asfortranarray(w.astype(self.D.np))  # src/ParTensor.jl:26-32
        return d


class ParBias(ParOperator):
    """``x + A`` with an ``m x 1`` ParMatrix (src/ParMatrix.jl:53-55) as an operator node so
    that it can be fused into a plan.  Non-linear (affine)."""
    L = NonLinear
    P = Parametric

    def __init__(self, matrix: ParMatrix):
        if matrix.n != 1:
            raise ParException("ParBias needs an m x 1 ParMatrix")
        self.matrix = matrix
        self.D = self.R = matrix.D

    Domain = property(lambda s: s.matrix.m)
    Range = property(lambda s: s.matrix.m)

    def init_(self, d, rng=None):
        return self.matrix.init_(d, rng)

    def parameterize(self, theta):
        return ParParameterized(self, theta[self.matrix])


class ParGelu(ParOperator):
    """``gelu.(x)`` of examples/fno.jl:43 as an operator node (non-linear, non-parametric)."""
    L = NonLinear

    def __init__(self, T, n):
        self.D = self.R = jtype(T)
        if self.D.is_complex:
            raise ParException("gelu is defined for real element types")
        self.n = int(n)

    Domain = property(lambda s: s.n)
    Range = property(lambda s: s.n)


# ---------------------------------------------------------------------------------------
# combinators
# ---------------------------------------------------------------------------------------
class ParCompose(ParOperator):
    """Composition ``ops[1] ∘ ... ∘ ops[N]``; ``ops[N]`` is applied first (src/ParCompose.jl:6-56)."""
    T = Internal

    def __new__(cls, *ops):
        if len(ops) == 1:
            return ops[0]  # src/ParCompose.jl:12-14
        return super().__new__(cls)

    def __init__(self, *ops):
        if len(ops) == 1:
            return
        ops = list(ops)
        for i in range(len(ops) - 1):  # the @assert pair of src/ParCompose.jl:16-19
            if ops[i].Domain != ops[i + 1].Range:
                raise ParException("ParCompose: Domain(ops[%d]) = %d != Range(ops[%d]) = %d"
                                   % (i + 1, ops[i].Domain, i + 2, ops[i + 1].Range))
            if ops[i].D is not ops[i + 1].R:
                raise ParException("ParCompose: DDT(ops[%d]) = %s != RDT(ops[%d]) = %s"
                                   % (i + 1, ops[i].D, i + 2, ops[i + 1].R))
        self.ops = ops
        self.D, self.R = ops[-1].D, ops[0].R
        self.L = _fold(promote_linearity, [o.L for o in ops])
        self.P = _fold(promote_parametricity, [o.P for o in ops])

    Domain = property(lambda s: s.ops[-1].Domain)
    Range = property(lambda s: s.ops[0].Range)

    def children(self):
        return self.ops

    def rebuild(self, cs):
        return ParCompose(*cs)

    def adjoint(self):
        if self.L is not Linear:
            raise ParException("adjoint is only defined for linear operators")
        return ParCompose(*[o.adjoint() for o in reversed(self.ops)])  # src/ParCompose.jl:42


def compose(*ops):
    """``∘`` with flattening (src/ParCompose.jl:31-35)."""
    flat = []
    for o in ops:
        flat.extend(o.ops if isinstance(o, ParCompose) else [o])
    return ParCompose(*flat)


def kron_order(ops) -> List[int]:
    """Application-order heuristic of the ParKron constructor (src/ParKron.jl:28-57);
    returns 1-based operator indices."""
    N = len(ops)
    DDTs, RDTs = [o.D for o in ops], [o.R for o in ops]
    T = _fold(subset_type, DDTs)
    out: List[int] = []
    for _ in range(N):
        cand = [j for j in range(N) if DDTs[j] is T and (j + 1) not in out]
        if not cand:
            raise ParException("ParKron: no factor accepts element type %s at this point of the application order" % T)
        Rt = _fold(subset_type, [RDTs[j] for j in cand])
        cand = [j for j in cand if RDTs[j] is Rt]
        diffs = [ops[j].Domain - ops[j].Range for j in cand]
        md = max(diffs)
        cand = [j for j, d in zip(cand, diffs) if d == md]
        out.append(cand[-1] + 1)  # right-most
        T = RDTs[cand[-1]]
    return out


def _cached_property(slot, fn):
    """Domain / Range of an internal node: operators are immutable once built, so the product over the children is taken once (it used to be
    re-derived recursively on every access -- a third of the host time of a small apply)."""
    def get(s):
        v = s.__dict__.get(slot)
        if v is None:
            v = s.__dict__[slot] = fn(s)
        return v
    return property(get)


class ParKron(ParOperator):
    """Kronecker product; ``ops[N]`` acts on the fastest array dim (src/ParKron.jl:14-97,190-220)."""
    T = Internal

    def __new__(cls, *ops, order=None, D=None, R=None, P=None):
        if len(ops) == 1 and order is None:
            return ops[0]  # src/ParKron.jl:24-26
        return super().__new__(cls)

    def __init__(self, *ops, order=None, D=None, R=None, P=None):
        if len(ops) == 1 and order is None:
            return
        self.ops = list(ops)
        N = len(self.ops)
        if order is None:
            self.order = kron_order(self.ops)
            self.D = self.ops[self.order[0] - 1].D
            self.R = self.ops[self.order[N - 1] - 1].R
            self.P = _fold(promote_parametricity, [o.P for o in self.ops])
        else:  # ParKron(D, R, P, ops, order) (src/ParKron.jl:66-68)
            self.order = [int(o) for o in order]
            self.D = D if D is not None else self.ops[self.order[0] - 1].D
            self.R = R if R is not None else self.ops[self.order[N - 1] - 1].R
            self.P = P if P is not None else _fold(promote_parametricity, [o.P for o in self.ops])
        self.L = Linear

    Domain = _cached_property("_domain", lambda s: int(np.prod([o.Domain for o in s.ops], dtype=object)))
    Range = _cached_property("_range", lambda s: int(np.prod([o.Range for o in s.ops], dtype=object)))

    def children(self):
        return self.ops

    def rebuild(self, cs):
        return ParKron(*cs)  # order re-derived by the heuristic (src/ParKron.jl:81)

    def adjoint(self):
        # src/ParKron.jl:82
        return ParKron(*[o.adjoint() for o in self.ops], order=list(reversed(self.order)), D=self.R, R=self.D, P=self.P)


def order(A: ParKron) -> List[int]:
    return A.order


def reorder(A: ParKron, ord_: Sequence[int]) -> ParKron:
    """src/ParKron.jl:90-97 (including its square-operator type check)."""
    for i in range(len(A.ops) - 1):
        if A.ops[ord_[i] - 1].R is not A.ops[ord_[i] - 1].D:
            raise ParException("Invalid order %s for Kronecker product. Types do not agree" % (list(ord_),))
    return ParKron(*A.ops, order=list(ord_), D=A.D, R=A.R, P=A.P)


def kron(*ops):
    """``⊗`` / ``kron`` with flattening (src/ParKron.jl:71-75) and ``I ⊗ I`` merging
    (src/ParIdentity.jl:35)."""
    if len(ops) > 2:
        return kron(kron(*ops[:2]), *ops[2:])
    A, B = ops
    if isinstance(A, ParIdentity) and isinstance(B, ParIdentity) and A.D is B.D:
        return ParIdentity(A.D, A.n * B.n)
    a = A.ops if isinstance(A, ParKron) else [A]
    b = B.ops if isinstance(B, ParKron) else [B]
    return ParKron(*a, *b)
