"""Lowering of operator trees to the flat ``po_node`` wire format and the apply entry points
(`A * x`, ``x * A``) on top of ``libparops.so``.

This is the Python twin of what ``julia/ParametricOperatorsB200.jl`` does with ``ccall``:
walk the parameterised tree once, emit nodes / children / aux arrays + the list of parameter
device pointers, fetch (or build) the cached plan for ``(tree structure, batch, flags)`` and
call ``po_plan_apply`` on the current CUDA stream with raw device pointers.  torch is used
for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import ParException, po_node
from . import operators as ops_mod
from .operators import (ParAdjoint, ParBias, ParCompose, ParDFT, ParDiagonal, ParGelu, ParIdentity, ParKron, ParMatrix,
                        ParOperator, ParParameterized, ParRestriction, ParTensor, Parametric, jtype)


def _torch():
    import torch
    return torch


class Engine:
    """One device context of libparops + its plan cache.

    ``device``: CUDA device index.  (``device='cpu'`` together with an explicit ``lib`` is
    accepted ONLY for the kernel emulator used by tests/emul -- the default engine is always
    the CUDA library and refuses CPU tensors.)
    """

    def __init__(self, device: Optional[int] = None, lib: Optional[_lib.Lib] = None):
        torch = _torch()
        self.lib = lib if lib is not None else _lib.default_lib()
        self.emulated = lib is not None and device == "cpu"
        if self.emulated:
            self.device_index, self.torch_device = 0, torch.device("cpu")
        else:
            if not torch.cuda.is_available():
                raise ParException("CUDA is not available: libparops is a GPU-only engine (no CPU fallback)")
            self.device_index = torch.cuda.current_device() if device is None else int(device)
            self.torch_device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        self.lib.check(self.lib.po_ctx_create(self.device_index, C.byref(h)))
        self.ctx = h
        self.plans = {}
        self.lock = threading.Lock()
        self.launches = 0  # kernels launched through this engine (bench.py's gpu_launches)
        self._keep = []    # ctypes callbacks that must outlive the communicator

    # -- memory helpers -------------------------------------------------------------------
    def stream_ptr(self):
        if self.emulated:
            return None
        return C.c_void_p(_torch().cuda.current_stream(self.device_index).cuda_stream)

    def empty_colmajor(self, rows: int, cols: int, T):
        torch = _torch()
        return torch.empty((cols, rows), dtype=jtype(T).torch, device=self.torch_device).t()

    def to_device(self, a):
        """numpy (any order) / torch tensor -> device tensor with the SAME shape and
        column-major (Julia) memory order."""
        torch = _torch()
        if isinstance(a, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(a.transpose()))
            t = t.to(self.torch_device)
            return t.permute(*reversed(range(t.dim()))) if t.dim() > 1 else t
        if isinstance(a, torch.Tensor):
            return colmajor(a.to(self.torch_device))
        raise ParException("cannot move %r to the device" % type(a))

    def device_status(self) -> int:
        """0 unless a kernel's bounded wait expired (see po_ctx_device_status); synchronises, reads AND clears."""
        v = C.c_int32()
        self.lib.check(self.lib.po_ctx_device_status(self.ctx, C.byref(v)))
        return v.value

    def raise_if_failed(self):
        """Host synchronisation points (``cpu(x)``, explicit ``synchronize``) call this: a device-side wait that
        timed out (a ParRepartition peer that never showed up, a stuck bulk-copy barrier) poisoned its outputs
        with NaNs and must surface as a ParException, not as silently wrong numbers."""
        if self.ctx is None:
            return
        st = self.device_status()
        if st:
            raise ParException("a device-side wait timed out (status 0x%x): results computed since the last check are invalid "
                               "(0x100/0x200: a ParRepartition peer did not acknowledge / deliver in PO_PEER_TIMEOUT_MS)" % st)

    def synchronize(self):
        if not self.emulated:
            _torch().cuda.synchronize(self.device_index)
        self.raise_if_failed()

    def destroy(self):
        for p in self.plans.values():
            self.lib.po_plan_destroy(p.handle)
        self.plans.clear()
        if self.ctx:
            self.lib.po_ctx_destroy(self.ctx)
            self.ctx = None


_default_engines = {}


def default_engine() -> Engine:
    torch = _torch()
    if not torch.cuda.is_available():
        raise ParException("CUDA is not available: libparops is a GPU-only engine (no CPU fallback)")
    d = torch.cuda.current_device()
    if d not in _default_engines:
        _default_engines[d] = Engine(d)
    return _default_engines[d]


_engine_override = threading.local()


class use_engine:
    """Context manager selecting the engine used by ``A * x`` (tests point it at the emulator)."""

    def __init__(self, eng):
        self.eng = eng

    def __enter__(self):
        self.prev = getattr(_engine_override, "eng", None)
        _engine_override.eng = self.eng
        return self.eng

    def __exit__(self, *a):
        _engine_override.eng = self.prev


def current_engine() -> Engine:
    e = getattr(_engine_override, "eng", None)
    return e if e is not None else default_engine()


def colmajor(t):
    """Return a tensor with the same shape/values whose memory is column-major."""
    if t.dim() <= 1:
        return t.contiguous()
    rev = list(reversed(range(t.dim())))
    return t.permute(*rev).contiguous().permute(*rev)


def is_colmajor(t) -> bool:
    if t.dim() <= 1:
        return t.is_contiguous()
    if t.dim() == 2:
        return t.t().is_contiguous()
    rev = list(reversed(range(t.dim())))
    return t.permute(*rev).is_contiguous()


def gpu(x, engine: Optional[Engine] = None):
    """``gpu(x)`` / ``gpu(θ)`` (src/ParOperator.jl:143-146)."""
    eng = engine or current_engine()
    if isinstance(x, dict):
        return {k: gpu(v, eng) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [gpu(v, eng) for v in x]
    return eng.to_device(x)


def cpu(x):
    """``cpu(x)`` / ``cpu(θ)`` (src/ParOperator.jl:134-137): back to column-major NumPy."""
    torch = _torch()
    if isinstance(x, dict):
        return {k: cpu(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [cpu(v) for v in x]
    if isinstance(x, torch.Tensor):
        h = np.asfortranarray(x.detach().cpu().numpy())  # synchronises the producing stream
        if x.is_cuda:
            for eng in list(_default_engines.values()) + [getattr(_engine_override, "eng", None)]:
                if eng is not None and not eng.emulated and eng.device_index == x.device.index:
                    eng.raise_if_failed()
        return h
    return x


# ---------------------------------------------------------------------------------------
# lowering
# ---------------------------------------------------------------------------------------
class Lowered:
    def __init__(self):
        self.nodes: List[tuple] = []
        self.child_index: List[int] = []
        self.aux: List[int] = []
        self.params: list = []        # slot -> parameter array (device tensor / numpy)
        self.param_ops: list = []     # slot -> leaf operator (key of the θ dict)
        self._slot_of = {}

    def slot(self, leaf, tensor) -> int:
        key = (id(leaf), id(tensor))
        if key not in self._slot_of:
            self._slot_of[key] = len(self.params)
            self.params.append(tensor)
            self.param_ops.append(leaf)
        return self._slot_of[key]

    def add(self, kind, ddt, rdt, adjoint, slot, domain, rng, aux=(), kids=()):
        ab, cb = len(self.aux), len(self.child_index)
        self.aux.extend(int(v) for v in aux)
        self.child_index.extend(int(k) for k in kids)
        self.nodes.append((kind, ddt.code, rdt.code, int(adjoint), int(slot), cb, len(kids), ab, len(aux), int(domain), int(rng)))
        return len(self.nodes) - 1

    def signature(self):
        return (tuple(self.nodes), tuple(self.child_index), tuple(self.aux))


def _restriction_aux(op: ParRestriction):
    aux = [op.n, len(op.ranges)]
    for a, b in op.ranges:
        aux.extend([a, b])
    return aux


def lower(op: ParOperator, lw: Lowered, adj: bool = False) -> int:
    from .distributed import ParBroadcasted, ParDistributed, ParRepartition  # late: avoids an import cycle

    if isinstance(op, ParParameterized):
        leaf, is_adj, tensor = op.leaf_and_tensor()
        is_adj = bool(is_adj) ^ adj
        if isinstance(leaf, ParMatrix):
            s = lw.slot(leaf, tensor)
            d, r = (leaf.m, leaf.n) if is_adj else (leaf.n, leaf.m)
            return lw.add(_lib.PO_NODE_MATRIX, leaf.D, leaf.D, is_adj, s, d, r, aux=[leaf.m, leaf.n])
        if isinstance(leaf, ParDiagonal):
            s = lw.slot(leaf, tensor)
            return lw.add(_lib.PO_NODE_DIAGONAL, leaf.D, leaf.D, is_adj, s, leaf.n, leaf.n, aux=[leaf.n])
        if isinstance(leaf, ParTensor):
            s = lw.slot(leaf, tensor)
            d, r = (leaf.Range, leaf.Domain) if is_adj else (leaf.Domain, leaf.Range)
            return lw.add(_lib.PO_NODE_TENSOR, leaf.D, leaf.D, is_adj, s, d, r, aux=[len(leaf.weight_shape), *leaf.weight_shape])
        if isinstance(leaf, ParBias):
            if is_adj:
                raise ParException("bias add has no adjoint")
            s = lw.slot(leaf.matrix, tensor)
            return lw.add(_lib.PO_NODE_BIAS, leaf.D, leaf.D, 0, s, leaf.Domain, leaf.Domain, aux=[leaf.Domain])
        if isinstance(leaf, (ParBroadcasted,)):
            return lower(leaf, lw, is_adj)
        raise ParException("cannot lower parameterised %s" % type(leaf).__name__)
    if isinstance(op, ParAdjoint):
        return lower(op.op, lw, not adj)
    if isinstance(op, ParBroadcasted):
        return lower(op.op, lw, adj)  # apply delegates (src/ParBroadcasted.jl:33-35)
    if isinstance(op, ParDistributed):
        if not isinstance(op.op, ParIdentity):
            raise ParException("ParDistributed apply is only defined for ParIdentity (src/ParIdentity.jl:20-22)")
        return lw.add(_lib.PO_NODE_IDENTITY, op.D, op.R, 0, -1, op.Domain, op.Range)
    if isinstance(op, ParIdentity):
        return lw.add(_lib.PO_NODE_IDENTITY, op.D, op.R, 0, -1, op.n, op.n)
    if isinstance(op, ParDFT):
        if adj:
            return lw.add(_lib.PO_NODE_DFT, op.R, op.D, 1, -1, op.m, op.n, aux=[op.n])
        return lw.add(_lib.PO_NODE_DFT, op.D, op.R, 0, -1, op.n, op.m, aux=[op.n])
    if isinstance(op, ParRestriction):
        d, r = (op.Range, op.Domain) if adj else (op.Domain, op.Range)
        return lw.add(_lib.PO_NODE_RESTRICTION, op.D, op.D, adj, -1, d, r, aux=_restriction_aux(op))
    if isinstance(op, ParGelu):
        if adj:
            raise ParException("gelu has no adjoint")
        return lw.add(_lib.PO_NODE_GELU, op.D, op.D, 0, -1, op.n, op.n)
    if isinstance(op, ParRepartition):
        o = op.adjoint() if adj else op
        aux = [o.N, *o.global_size, *o.dims_in, *o.dims_out, o.rank]
        return lw.add(_lib.PO_NODE_REPARTITION, o.D, o.D, 0, -1, o.Domain, o.Range, aux=aux)
    if isinstance(op, ParCompose):
        kids_ops = [c.adjoint() for c in reversed(op.ops)] if adj else op.ops
        kids = [lower(c, lw, False) for c in kids_ops]
        d, r = (op.Range, op.Domain) if adj else (op.Domain, op.Range)
        dd, rr = (op.R, op.D) if adj else (op.D, op.R)
        return lw.add(_lib.PO_NODE_COMPOSE, dd, rr, 0, -1, d, r, kids=kids)
    if isinstance(op, ParKron):
        o = op.adjoint() if adj else op
        kids = [lower(c, lw, False) for c in o.ops]
        return lw.add(_lib.PO_NODE_KRON, o.D, o.R, 0, -1, o.Domain, o.Range, aux=o.order, kids=kids)
    if op.P is Parametric:
        raise ParException("operator %r has unbound parameters: call A(θ) first" % (op,))
    raise ParException("cannot lower operator of type %s" % type(op).__name__)


class Plan:
    def __init__(self, eng: Engine, lw: Lowered, root: int, batch: int, flags: int):
        self.eng = eng
        n = len(lw.nodes)
        arr = (po_node * n)()
        for i, t in enumerate(lw.nodes):
            (arr[i].kind, arr[i].ddt, arr[i].rdt, arr[i].adjoint, arr[i].param_slot, arr[i].child_begin, arr[i].child_count,
             arr[i].aux_begin, arr[i].aux_count, arr[i].domain, arr[i].range) = t
        ci = (C.c_int32 * max(1, len(lw.child_index)))(*lw.child_index)
        aux = (C.c_int64 * max(1, len(lw.aux)))(*lw.aux)
        h = C.c_void_p()
        eng.lib.check(eng.lib.po_plan_create(eng.ctx, arr, n, ci, len(lw.child_index), aux, len(lw.aux), root, batch, flags, C.byref(h)))
        self.handle = h
        sz = C.c_size_t()
        eng.lib.check(eng.lib.po_plan_workspace_bytes(h, C.byref(sz)))
        self.ws_bytes = sz.value
        self.ws = None
        eng.lib.check(eng.lib.po_plan_tape_bytes(h, C.byref(sz)))
        self.tape_bytes = sz.value
        self.batch = batch

    def workspace(self):
        if self.ws is None and self.ws_bytes:
            torch = _torch()
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.eng.torch_device)
        return self.ws

    def describe(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        self.eng.lib.check(self.eng.lib.po_plan_describe(self.handle, buf, len(buf)))
        return buf.value.decode()

    def num_steps(self) -> int:
        n = C.c_int32()
        self.eng.lib.check(self.eng.lib.po_plan_num_steps(self.handle, C.byref(n)))
        return n.value

    def algorithmic_bytes(self) -> int:
        n = C.c_int64()
        self.eng.lib.check(self.eng.lib.po_plan_algorithmic_bytes(self.handle, C.byref(n)))
        return n.value

    def set_profiling(self, enable: bool):
        self.eng.lib.check(self.eng.lib.po_plan_set_profiling(self.handle, int(bool(enable))))

    def step_profile(self):
        """[(note, ms_sum, calls, algorithmic_bytes)] per step since profiling was enabled."""
        n = self.num_steps()
        ms, calls, byts = (C.c_double * max(1, n))(), (C.c_int64 * max(1, n))(), (C.c_int64 * max(1, n))()
        nn = C.c_int32()
        self.eng.lib.check(self.eng.lib.po_plan_step_profile(self.handle, ms, calls, byts, n, C.byref(nn)))
        notes = [ln.split(": ", 1)[1] if ": " in ln else ln for ln in self.describe().splitlines()]
        return [(notes[i] if i < len(notes) else "", ms[i], calls[i], byts[i]) for i in range(n)]

    def pullback_profile(self):
        """[(note, ms_sum, calls, algorithmic_bytes)] of the reverse-mode kernels per step."""
        n = self.num_steps()
        ms, calls = (C.c_double * max(1, n))(), (C.c_int64 * max(1, n))()
        nn = C.c_int32()
        self.eng.lib.check(self.eng.lib.po_plan_pullback_step_profile(self.handle, ms, calls, n, C.byref(nn)))
        fwd = self.step_profile()
        return [("pullback of " + fwd[i][0], ms[i], calls[i], fwd[i][3]) for i in range(n)]

    def launch_count(self) -> int:
        n = C.c_int64()
        self.eng.lib.check(self.eng.lib.po_plan_launch_count(self.handle, C.byref(n)))
        return n.value


def get_plan(op: ParOperator, batch: int, eng: Engine, flags: int = 0):
    memo = op.__dict__.setdefault("_plans", {})  # operators are immutable once built: lower each (op, batch) once
    hit = memo.get((id(eng), batch, flags))
    if hit is not None:
        return hit
    hit = _get_plan(op, batch, eng, flags)
    memo[(id(eng), batch, flags)] = hit
    return hit


def _get_plan(op: ParOperator, batch: int, eng: Engine, flags: int = 0):
    lw = Lowered()
    root = lower(op, lw)
    key = (lw.signature(), root, batch, flags)
    with eng.lock:
        plan = eng.plans.get(key)
        if plan is None:
            plan = Plan(eng, lw, root, batch, flags)
            eng.plans[key] = plan
    return plan, lw


def _param_ptrs(eng: Engine, lw: Lowered):
    """Device pointers of the parameters (host arrays are uploaded for this call)."""
    torch = _torch()
    keep, ptrs = [], []
    for leaf, p in zip(lw.param_ops, lw.params):
        if isinstance(p, np.ndarray):
            p = eng.to_device(np.asarray(p, dtype=leaf.D.np))
        if not isinstance(p, torch.Tensor):
            raise ParException("parameter for %r is not an array" % (leaf,))
        if p.device != eng.torch_device:
            raise ParException("parameter for %r lives on %s but the operator is applied on %s" % (leaf, p.device, eng.torch_device))
        if p.dtype != leaf.D.torch:
            raise ParException("parameter for %r has dtype %s, expected %s" % (leaf, p.dtype, leaf.D))
        if not is_colmajor(p):
            p = colmajor(p)
        keep.append(p)
        ptrs.append(p.data_ptr())
    arr = (C.c_void_p * max(1, len(ptrs)))(*ptrs)
    return arr, len(ptrs), keep


def apply_operator(op: ParOperator, x, engine: Optional[Engine] = None, flags: int = 0, out=None, wait: bool = True):
    """``A * x``: vector ``(Domain,)`` or column-major matrix ``(Domain, batch)``.

    * device tensor in -> device tensor out (the drop-in path: raw pointers, current stream);
    * NumPy array in  -> NumPy array out through ``po_plan_apply_host`` (H2D + apply + D2H),
      the reference-facing call with HOST buffers that bench.py's ``e2e`` times.  ``out`` may name the
      host result buffer (Fortran-ordered, right shape and type; e.g. a view of page-locked memory, which
      the D2H copy then reaches at full PCIe rate instead of through the driver's pageable staging).
      ``wait=False`` only enqueues the three operations on the current stream (``po_plan_apply_host_async``): ``x`` and
      ``out`` must then be page-locked and stay alive, and the result is valid once the caller has synchronised the stream.
    """
    torch = _torch()
    eng = engine or current_engine()
    host = isinstance(x, np.ndarray)
    if not host and not isinstance(x, torch.Tensor):
        raise ParException("operand must be a torch tensor (device) or a NumPy array (host)")
    if x.ndim == 3:
        return _apply_batched3(op, x, eng, flags)
    if x.ndim not in (1, 2):
        raise ParException("operand must be a vector or a matrix")
    is_vec = x.ndim == 1
    n_in = x.shape[0]
    batch = 1 if is_vec else x.shape[1]
    if n_in != op.Domain:
        raise ParException("operand has %d rows but Domain(A) = %d" % (n_in, op.Domain))
    xt = jtype(x.dtype)
    if xt is not op.D:
        raise ParException("operand element type %s does not match DDT(A) = %s (MethodError in the reference)" % (xt, op.D))
    plan, lw = get_plan(op, batch, eng, flags)
    params, n_params, keep = _param_ptrs(eng, lw)
    if host:
        xh = np.asfortranarray(x)
        yshape = (op.Range, batch) if not is_vec else (op.Range,)
        if out is not None:
            if not isinstance(out, np.ndarray) or out.shape != yshape or out.dtype != op.R.np or not out.flags.f_contiguous:
                raise ParException("out must be a Fortran-contiguous NumPy array of shape %s and type %s" % (yshape, op.R.np))
            yh = out
        else:
            yh = np.empty(yshape, dtype=op.R.np, order="F")
        if not wait and (out is None or xh is not x):
            raise ParException("wait=False needs a caller-owned Fortran-contiguous x and an out= buffer (both page-locked) that outlive the call")
        entry = eng.lib.po_plan_apply_host if wait else eng.lib.po_plan_apply_host_async
        eng.lib.check(entry(plan.handle, C.c_void_p(xh.ctypes.data), C.c_void_p(yh.ctypes.data), params, n_params, eng.stream_ptr()))
        eng.launches += plan.launch_count()
        del keep   # device tensors: the caching allocator recycles them in stream order, like the device path below
        return yh
    if x.device != eng.torch_device:
        raise ParException("operand lives on %s, engine on %s" % (x.device, eng.torch_device))
    xd = x if is_colmajor(x) else colmajor(x)
    if op.Range == op.Domain and plan.num_steps() == 0:
        return x  # identity: the reference returns its argument (src/ParIdentity.jl:16-17)
    y = eng.empty_colmajor(op.Range, batch, op.R)
    ws = plan.workspace()
    eng.lib.check(eng.lib.po_plan_apply(plan.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(y.data_ptr()), params, n_params,
                                        C.c_void_p(ws.data_ptr()) if ws is not None else None, plan.ws_bytes, eng.stream_ptr()))
    eng.launches += plan.launch_count()
    del keep
    return y.reshape(-1) if is_vec else y


def _apply_batched3(op: ParOperator, x, eng: Engine, flags: int):
    """``A * x`` for a 3-D operand ``(n, C, B)``: defined by the reference for a parameterised ParMatrix only
    (``batched_mul(A.params, x)``, src/ParMatrix.jl:44 -- NNlib broadcasts the 2-D weight over the batch dim) and for a
    ParBroadcasted around one (src/ParBroadcasted.jl:35).  Column-major, so it is the 2-D apply on ``(n, C*B)``."""
    from .distributed import ParBroadcasted
    inner = op.op if isinstance(op, ParBroadcasted) else op
    if not (isinstance(inner, ParParameterized) and isinstance(inner.op, ParMatrix)):
        raise ParException("3-d operands are only defined for a parameterised ParMatrix (MethodError in the reference)")
    n, c, b = x.shape
    if isinstance(x, np.ndarray):
        y = apply_operator(op, np.asfortranarray(x).reshape((n, c * b), order="F"), eng, flags)
        return y.reshape((op.Range, c, b), order="F")
    xd = x if is_colmajor(x) else colmajor(x)
    x2 = xd.permute(2, 1, 0).reshape(b * c, n).t()          # the same column-major storage viewed (n, C*B)
    y2 = apply_operator(op, x2, eng, flags)
    return y2.t().reshape(b, c, op.Range).permute(2, 1, 0)


def divide(A: ParOperator, x):
    """``A / x = A.params ./ x`` (src/ParMatrix.jl:60; ParBroadcasted delegates, src/ParBroadcasted.jl:39): elementwise, host-level."""
    from .distributed import ParBroadcasted
    inner = A.op if isinstance(A, ParBroadcasted) else A
    if not (isinstance(inner, ParParameterized) and isinstance(inner.op, ParMatrix)):
        raise ParException("`/` is only defined for a parameterised ParMatrix (MethodError in the reference)")
    return inner.params / x


def rrule(op: ParOperator, x, engine: Optional[Engine] = None, flags: int = 0):
    """``y, pullback = rrule(A, x)`` -- the ChainRulesCore.rrule every ccall-backed apply needs
    (SURVEY.md 3.5).  ``pullback(ybar)`` returns ``(grads, xbar)`` where ``grads`` maps each
    parametric leaf (the key of the θ dict, as Zygote's Dict adjoint does) to the cotangent of
    its parameter and ``xbar = (dy/dx)' * ybar``.  The forward pass leaves the input of every
    parametric step in a tape; the DFT cotangents use the TRUE adjoints (AbstractFFTs rules),
    not the inverse-as-adjoint of ``A'``."""
    torch = _torch()
    eng = engine or current_engine()
    if not isinstance(x, torch.Tensor):
        raise ParException("rrule needs a device tensor")
    is_vec = x.ndim == 1
    batch = 1 if is_vec else x.shape[1]
    if x.shape[0] != op.Domain or jtype(x.dtype) is not op.D:
        raise ParException("operand does not match Domain/DDT of the operator")
    plan, lw = get_plan(op, batch, eng, flags)
    params, n_params, keep = _param_ptrs(eng, lw)
    xd = x if is_colmajor(x) else colmajor(x)
    y = eng.empty_colmajor(op.Range, batch, op.R)
    tape = torch.empty(max(1, plan.tape_bytes), dtype=torch.uint8, device=eng.torch_device)
    ws = plan.workspace()
    wsp = C.c_void_p(ws.data_ptr()) if ws is not None else None
    if plan.num_steps() == 0:
        y = xd.clone()
    else:
        eng.lib.check(eng.lib.po_plan_apply_taped(plan.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(y.data_ptr()), params, n_params,
                                                  C.c_void_p(tape.data_ptr()), plan.tape_bytes, wsp, plan.ws_bytes, eng.stream_ptr()))
        eng.launches += plan.launch_count()

    def pullback(ybar):
        yb = ybar.reshape(op.Range, batch) if ybar.ndim == 1 else ybar
        if yb.shape != (op.Range, batch) or jtype(yb.dtype) is not op.R:
            raise ParException("cotangent does not match Range/RDT of the operator")
        yb = yb if is_colmajor(yb) else colmajor(yb)
        xbar = eng.empty_colmajor(op.Domain, batch, op.D)
        gts = [torch.zeros_like(colmajor(k)) for k in keep]
        garr = (C.c_void_p * max(1, len(gts)))(*[g.data_ptr() for g in gts])
        eng.lib.check(eng.lib.po_plan_pullback(plan.handle, C.c_void_p(tape.data_ptr()), C.c_void_p(yb.data_ptr()), C.c_void_p(xbar.data_ptr()),
                                               params, garr, n_params, wsp, plan.ws_bytes, eng.stream_ptr()))
        eng.launches += plan.launch_count()
        grads = {}
        for leaf, g in zip(lw.param_ops, gts):
            grads[leaf] = grads[leaf] + g if leaf in grads else g
        return grads, (xbar.reshape(-1) if is_vec else xbar)

    return (y.reshape(-1) if is_vec else y), pullback


def apply_right(op: ParOperator, x, engine: Optional[Engine] = None):
    """``x * A = (A' * x')'`` (src/ParOperator.jl:228)."""
    torch = _torch()
    if isinstance(x, np.ndarray):
        y = apply_operator(op.adjoint(), np.asfortranarray(np.conj(x.T)), engine)
        return np.conj(y.T)
    xh = x.conj().t() if x.dim() == 2 else x.conj()
    y = apply_operator(op.adjoint(), colmajor(xh.resolve_conj()), engine)
    return (y.conj().t() if y.dim() == 2 else y.conj()).resolve_conj()


def fused_add_gelu(y1, y2=None, engine: Optional[Engine] = None):
    """``gelu.(y1 .+ y2)`` (examples/fno.jl:43) as one kernel."""
    eng = engine or current_engine()
    T = jtype(y1.dtype)
    if y2 is not None and (y2.shape != y1.shape or y2.dtype != y1.dtype):
        raise ParException("gelu.(y1 .+ y2): operands differ in shape or type")
    a = y1 if is_colmajor(y1) else colmajor(y1)
    b = None if y2 is None else (y2 if is_colmajor(y2) else colmajor(y2))
    out = _torch().empty_like(a)
    eng.lib.check(eng.lib.po_add_gelu(eng.ctx, T.code, a.numel(), C.c_void_p(a.data_ptr()),
                                      C.c_void_p(b.data_ptr()) if b is not None else None, C.c_void_p(out.data_ptr()),
                                      eng.stream_ptr()))
    eng.launches += 1
    return out


def block_epilogue(W, x, s=None, b=None, gelu=True, engine: Optional[Engine] = None):
    """``act.(s .+ (I_grid ⊗ W) x .+ b)`` in one pass (examples/fno.jl:34-44; po_block_epilogue).

    ``W``: (c_out, c_in) column-major parameter of the pointwise ParMatrix; ``x``: (c_in*npts, batch) column-major
    with channels fastest; ``s``: same layout with c_out channels (the spectral-conv output) or None."""
    eng = engine or current_engine()
    T = jtype(x.dtype)
    c_out, c_in = int(W.shape[0]), int(W.shape[1])
    if x.shape[0] % c_in:
        raise ParException("block epilogue: x has %d rows, not a multiple of c_in=%d" % (x.shape[0], c_in))
    batch = 1 if x.dim() == 1 else int(x.shape[1])
    npts = (int(x.shape[0]) // c_in) * batch
    xa = x if is_colmajor(x) else colmajor(x)
    Wa = W if is_colmajor(W) else colmajor(W)
    sa = None
    if s is not None:
        if s.shape[0] * (1 if s.dim() == 1 else s.shape[1]) != c_out * npts or s.dtype != x.dtype:
            raise ParException("block epilogue: s does not match (c_out*npts, batch)")
        sa = s if is_colmajor(s) else colmajor(s)
    if W.dtype != x.dtype or (b is not None and b.dtype != x.dtype):
        raise ParException("block epilogue: operands differ in element type")
    out = eng.empty_colmajor((int(x.shape[0]) // c_in) * c_out, batch, T) if x.dim() == 2 else _torch().empty(
        (int(x.shape[0]) // c_in) * c_out, dtype=x.dtype, device=x.device)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    eng.lib.check(eng.lib.po_block_epilogue(eng.ctx, T.code, c_in, c_out, npts, ptr(Wa), ptr(xa), ptr(sa),
                                            ptr(b.contiguous() if b is not None else None), 1 if gelu else 0, ptr(out), eng.stream_ptr()))
    eng.launches += 1
    return out


def block_epilogue_pullback(W, x, ybar, s=None, b=None, gelu=True, need_x=True, need_W=True, need_b=None, engine: Optional[Engine] = None):
    """Cotangents of ``y = act.(s .+ (I_grid ⊗ W) x .+ b)`` (po_block_epilogue_pullback): returns ``(sbar, xbar, Wbar, bbar)``;
    the pre-activation is recomputed, nothing is taped.  Entries that were not asked for are None."""
    torch = _torch()
    eng = engine or current_engine()
    T = jtype(x.dtype)
    c_out, c_in = int(W.shape[0]), int(W.shape[1])
    batch = 1 if x.dim() == 1 else int(x.shape[1])
    npts = (int(x.shape[0]) // c_in) * batch
    need_b = (b is not None) if need_b is None else need_b
    cm = lambda t: None if t is None else (t if is_colmajor(t) else colmajor(t))  # noqa: E731
    Wa, xa, sa, ya = cm(W), cm(x), cm(s), cm(ybar)
    if ya.shape[0] * (1 if ya.dim() == 1 else ya.shape[1]) != c_out * npts or ya.dtype != x.dtype:
        raise ParException("block epilogue pullback: ybar does not match (c_out*npts, batch)")
    sbar = torch.empty_like(ya)
    xbar = torch.empty_like(xa) if need_x else None
    Wbar = torch.empty((c_in, c_out), dtype=x.dtype, device=x.device).t() if need_W else None
    bbar = torch.empty(c_out, dtype=x.dtype, device=x.device) if need_b else None
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    eng.lib.check(eng.lib.po_block_epilogue_pullback(eng.ctx, T.code, c_in, c_out, npts, ptr(Wa), ptr(xa), ptr(sa), ptr(b.contiguous() if b is not None else None),
                                                     1 if gelu else 0, ptr(ya), ptr(sbar), ptr(xbar), ptr(Wbar), ptr(bbar), eng.stream_ptr()))
    eng.launches += 2 + int(need_x) * 2 + int(need_W) + int(need_b)
    return sbar, xbar, Wbar, bbar


ops_mod.apply_operator = apply_operator
