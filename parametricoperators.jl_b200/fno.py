"""The FNO building blocks of examples/fno.jl, written against the mirrored operator API."""
from __future__ import annotations

import numpy as np

from .operators import (Complex, ParCompose, ParDFT, ParDiagonal, ParIdentity, ParKron, ParMatrix, ParRestriction,
                        ParTensor, compose, jtype, kron)

__all__ = ["spectral_convolution", "spectral_convolution_tensor", "channel_mixing", "fno_block", "fno_block_rrule"]


def _dft_restrict(T, shape_spacetime, modes_spacetime):
    T = jtype(T)
    CT = Complex(T)
    dft_time = ParDFT(T, shape_spacetime[-1])
    dft_space = ParKron(*[ParDFT(CT, s) for s in shape_spacetime[:-1]])
    dft_spacetime = kron(dft_time, dft_space)
    st_ops = dft_spacetime.ops if isinstance(dft_spacetime, ParKron) else [dft_spacetime]
    restrict = ParKron(*[ParRestriction(CT, F.Range, m) for F, m in zip(st_ops, reversed(list(modes_spacetime)))])
    return compose(restrict, dft_spacetime)  # restrict*dft_spacetime


def spectral_convolution(T, shape_spacetime, modes_spacetime, lifted_dim):
    """examples/fno.jl:9-32: ``S = dft_all' * (ParDiagonal ⊗ ParMatrix) * dft_all``.

    ``shape_spacetime`` lists the grid extents slowest-Kron-factor-last exactly as in the
    example (``[nx, ny, nt]``: the LAST entry is the real-transformed "time" dim, which is
    the slowest array dim); ``modes_spacetime[i]`` is the list of kept 1-based ranges of dim i.
    """
    T = jtype(T)
    CT = Complex(T)
    dft_restrict = _dft_restrict(T, shape_spacetime, modes_spacetime)
    identity_channels = ParIdentity(T, lifted_dim)
    dft_all = kron(dft_restrict, identity_channels)
    mix_channels = ParMatrix(CT, lifted_dim, lifted_dim)
    elementwise_weighting = ParDiagonal(CT, dft_restrict.Range)
    frequency_weight = kron(elementwise_weighting, mix_channels)
    return compose(dft_all.adjoint(), frequency_weight, dft_all)


def spectral_convolution_tensor(T, shape_spacetime, modes_spacetime, lifted_dim, out_dim=None):
    """Per-mode-weights variant: the frequency weighting is a ``ParTensor`` (rank = 2 + number
    of grid dims, weights ``(oc, ic, modes...)``; src/ParTensor.jl:60-113) instead of
    ``ParDiagonal ⊗ ParMatrix``."""
    T = jtype(T)
    CT = Complex(T)
    oc = lifted_dim if out_dim is None else out_dim
    dft_restrict = _dft_restrict(T, shape_spacetime, modes_spacetime)
    dft_all = kron(dft_restrict, ParIdentity(T, lifted_dim))
    # kept modes per ARRAY dim, fastest first: shape_spacetime[-2], ..., shape_spacetime[0], time
    st_ops = dft_restrict.ops[0].ops if isinstance(dft_restrict.ops[0], ParKron) else [dft_restrict.ops[0]]
    kept = [r.Range for r in reversed(st_ops)]
    nd = len(kept)
    if nd == 2:  # rank 4: (ic, oc, nt, nxy) -- the example's "it(xy)" layout
        raise NotImplementedError("rank-4 ParTensor uses a (t, xy) mode layout; build it explicitly")
    ws = (oc, lifted_dim, *kept)
    order_w = tuple(range(1, 3 + nd))
    W = ParTensor(CT, order_w, ws, tuple(range(2, 3 + nd)), (lifted_dim, *kept), (1,) + tuple(range(3, 3 + nd)), (oc, *kept))
    if oc != lifted_dim:
        raise NotImplementedError("dft_all' assumes out_dim == lifted_dim")
    return compose(dft_all.adjoint(), W, dft_all)


def channel_mixing(T, shape_spacetime, lifted_dim):
    """examples/fno.jl:34-38: ``I_grid ⊗ ParMatrix(c, c)``."""
    T = jtype(T)
    mix_channels = ParMatrix(T, lifted_dim, lifted_dim)
    identity_spacetime = ParIdentity(T, int(np.prod(shape_spacetime)))
    return kron(identity_spacetime, mix_channels)


def _pointwise_matrix(channel_mix):
    """(W parameter tensor) if ``channel_mix`` is a parameterised ``ParIdentity ⊗ ParMatrix`` with a real type, else None."""
    from .operators import ParParameterized
    op = channel_mix
    if not isinstance(op, ParKron) or len(op.ops) != 2 or not isinstance(op.ops[0], ParIdentity):
        return None
    leaf = op.ops[1]
    if not isinstance(leaf, ParParameterized) or not isinstance(leaf.op, ParMatrix):
        return None
    return leaf.params


def fno_block(x, spectral_conv, channel_mix):
    """examples/fno.jl:40-44: ``gelu.(S x .+ W x)``.  On device tensors the pointwise branch ``W x`` (``I_grid ⊗
    ParMatrix``), the add and the gelu run as ONE kernel that reads ``S x`` and ``x`` once and writes the result once
    (po_block_epilogue); other operand kinds fall back to ``W x`` + the fused add+gelu kernel."""
    from .engine import block_epilogue, fused_add_gelu
    y1 = spectral_conv(x)
    W = _pointwise_matrix(channel_mix)
    if W is not None and hasattr(x, "data_ptr") and not x.is_complex():
        return block_epilogue(W, x, y1)
    y2 = channel_mix(x)
    return fused_add_gelu(y1, y2)


def fno_block_rrule(x, spectral_conv, channel_mix):
    """``y, pullback = fno_block_rrule(x, S(θ), W(θ))``: reverse mode of examples/fno.jl:40-44 as Zygote would produce it.
    ``pullback(ybar)`` returns ``(grads, xbar)`` with ``grads`` keyed by the parametric leaves of both branches.  The spectral
    branch uses the plan's taped apply + po_plan_pullback, the pointwise branch + activation po_block_epilogue_pullback
    (pre-activation recomputed: only the spectral tape is kept)."""
    from .engine import block_epilogue, block_epilogue_pullback, rrule
    from .operators import ParParameterized
    W = _pointwise_matrix(channel_mix)
    if W is None or not hasattr(x, "data_ptr") or x.is_complex():
        raise ParException("fno_block_rrule needs a device tensor and a parameterised ParIdentity ⊗ ParMatrix pointwise branch")
    leafW = channel_mix.ops[1].op
    s, back_s = rrule(spectral_conv, x)
    y = block_epilogue(W, x, s)

    def pullback(ybar):
        sbar, xbar_w, Wbar, _ = block_epilogue_pullback(W, x, ybar, s=s)
        grads, xbar_s = back_s(sbar)
        grads = dict(grads)
        grads[leafW] = grads[leafW] + Wbar if leafW in grads else Wbar
        return grads, xbar_s + xbar_w

    return y, pullback
