"""JSON wire format of operator trees (src/ASTSerialization.jl + the per-type ``to_Dict`` / ``from_Dict`` methods).

Same keys and value conventions as the reference, so a tree written by either side is read by the other:

    ParDFT          {"type", "T", "n", "m"}                      src/ParDFT.jl:34-49 ("m" informational, checked when present)
    ParIdentity     {"type", "T", "n"}                           src/ParIdentity.jl:24-33
    ParRestriction  {"type", "T", "n", "ranges": [[start, stop], ...]}   src/ParRestriction.jl:45-55 (1-based, inclusive)
    ParMatrix       {"type", "T", "m", "n", "id"}                src/ParMatrix.jl:62-90 (uuid ids as "UUID:<uuid>")
    ParDiagonal     {"type", "T", "n", "id"}                     src/ParDiagonal.jl:33-60
    ParTensor       {"type", "T", "ws", "wo", "is", "io", "to", "ts", "id"}   src/ParTensor.jl:119-151
    ParAdjoint      {"type", "of": <op>}                         src/ParAdjoint.jl:23-28
    ParCompose      {"type", "of": [<op>, ...]}                  src/ParCompose.jl:100-105
    ParKron         {"type", "of": [<op>, ...], "order": [...]}  src/ParKron.jl:311-322 (1-based application order kept as written)

    ParRepartition  {"type", "T", "global_size", "dims_in", "dims_out", "rank"}      (extension)
    ParBroadcasted  {"type", "of": <op>, "root"}                                      (extension)
    ParDistributed  {"type", "of": <op>, "rank", "size"}                              (extension)

The reference's type registry (src/ASTSerialization.jl:16-27) leaves ParTensor and the distributed operators out (commented lines 17, 21,
25); ParTensor is registered here because its to_Dict / from_Dict pair exists (src/ParTensor.jl:119-151), and the three distributed node
types get the dictionaries above so that a distributed pipeline can be persisted / handed to the C side too (``JsonPlan`` =
po_plan_create_json: the library parses the same text and needs no host mirror of the tree).
Unknown "type" / "T" values raise ``ParException`` with the reference's messages (src/ASTSerialization.jl:38-50).
"""
from __future__ import annotations

import json
import uuid

from ._lib import ParException
from .operators import (ComplexF32, ComplexF64, Float32, Float64, ParAdjoint, ParCompose, ParDFT, ParDiagonal, ParIdentity,
                        ParKron, ParMatrix, ParOperator, ParRestriction, ParTensor)

__all__ = ["to_Dict", "from_Dict", "to_json", "from_json", "JsonPlan"]

# Float16 / ComplexF16 are in the reference's table but no kernel exists for them: reading one is an error at apply time there too
_DATA_TYPES = {"Float32": Float32, "Float64": Float64, "ComplexF32": ComplexF32, "ComplexF64": ComplexF64}


def _dtype(d):
    ts = d["T"]
    if ts not in _DATA_TYPES:
        raise ParException("unknown data type `%s`" % ts)
    return _DATA_TYPES[ts]


def _enc_id(i):
    if isinstance(i, str):
        return i
    if isinstance(i, uuid.UUID):
        return "UUID:%s" % i
    raise ParException("I don't know how to encode id of type %s" % type(i).__name__)


def _dec_id(s):
    return uuid.UUID(s[5:]) if s.startswith("UUID:") else s


def to_Dict(A: ParOperator) -> dict:
    if isinstance(A, ParAdjoint):
        return {"type": "ParAdjoint", "of": to_Dict(A.op)}
    if isinstance(A, ParCompose):
        return {"type": "ParCompose", "of": [to_Dict(o) for o in A.ops]}
    if isinstance(A, ParKron):
        return {"type": "ParKron", "of": [to_Dict(o) for o in A.ops], "order": [int(o) for o in A.order]}
    if isinstance(A, ParDFT):
        return {"type": "ParDFT", "T": A.D.name, "n": A.n, "m": A.m}
    if isinstance(A, ParIdentity):
        return {"type": "ParIdentity", "T": A.D.name, "n": A.n}
    if isinstance(A, ParRestriction):
        return {"type": "ParRestriction", "T": A.D.name, "n": A.n, "ranges": [[int(a), int(b)] for a, b in A.ranges]}
    if isinstance(A, ParMatrix):
        return {"type": "ParMatrix", "T": A.D.name, "m": A.m, "n": A.n, "id": _enc_id(A.id)}
    if isinstance(A, ParDiagonal):
        return {"type": "ParDiagonal", "T": A.D.name, "n": A.n, "id": _enc_id(A.id)}
    if isinstance(A, ParTensor):
        return {"type": "ParTensor", "T": A.D.name, "ws": list(A.weight_shape), "wo": list(A.weight_order), "is": list(A.input_shape),
                "io": list(A.input_order), "to": list(A.target_order), "ts": list(A.target_shape), "id": _enc_id(A.id)}
    from .distributed import ParBroadcasted, ParDistributed, ParRepartition
    if isinstance(A, ParRepartition):
        return {"type": "ParRepartition", "T": A.D.name, "global_size": [int(g) for g in A.global_size], "dims_in": [int(d) for d in A.dims_in],
                "dims_out": [int(d) for d in A.dims_out], "rank": int(A.rank)}
    if isinstance(A, ParBroadcasted):
        return {"type": "ParBroadcasted", "of": to_Dict(A.op), "root": int(A.root)}
    if isinstance(A, ParDistributed):
        return {"type": "ParDistributed", "of": to_Dict(A.op), "rank": int(A.rank), "size": int(A.size)}
    raise ParException("Unimplemented")  # src/ParOperator.jl:233


def _from_kron(d):
    ops = [from_Dict(o) for o in d["of"]]
    return ParKron(*ops, order=[int(o) for o in d["order"]])  # D, R, P re-derived from the children like src/ParKron.jl:317-321


_FROM = {
    "ParAdjoint": lambda d: ParAdjoint(from_Dict(d["of"])),
    "ParCompose": lambda d: ParCompose(*[from_Dict(o) for o in d["of"]]),
    "ParKron": _from_kron,
    "ParIdentity": lambda d: ParIdentity(_dtype(d), d["n"]),
    "ParRestriction": lambda d: ParRestriction(_dtype(d), d["n"], [(int(a), int(b)) for a, b in d["ranges"]]),
    "ParMatrix": lambda d: ParMatrix(_dtype(d), d["m"], d["n"], _dec_id(d["id"])),
    "ParDiagonal": lambda d: ParDiagonal(_dtype(d), d["n"], _dec_id(d["id"])),
    "ParTensor": lambda d: ParTensor(_dtype(d), d["wo"], d["ws"], d["io"], d["is"], d["to"], d["ts"], _dec_id(d["id"])),
}


def _from_dft(d):
    A = ParDFT(_dtype(d), d["n"])
    if "m" in d and A.m != d["m"]:  # the "m" is informational, but if it's present, make sure it matches
        raise ParException("range mismatch: data says %s, but parDFT calculated %s" % (d["m"], A.m))
    return A


_FROM["ParDFT"] = _from_dft


def _from_repartition(d):
    from .distributed import CartComm, ParRepartition
    r = d.get("rank")
    return ParRepartition(_dtype(d), CartComm(d["dims_in"], r), CartComm(d["dims_out"], r), d["global_size"])


def _from_broadcasted(d):
    from .distributed import ParBroadcasted
    return ParBroadcasted(from_Dict(d["of"]), None, int(d.get("root", 0)))


def _from_distributed(d):
    from .distributed import ParDistributed
    return ParDistributed(from_Dict(d["of"]), d["rank"], d["size"])


_FROM["ParRepartition"] = _from_repartition
_FROM["ParBroadcasted"] = _from_broadcasted
_FROM["ParDistributed"] = _from_distributed


def from_Dict(d: dict) -> ParOperator:
    if "type" not in d:
        raise ParException("dict has no 'type' key")
    ts = d["type"]
    if ts not in _FROM:
        raise ParException("cannot convert Dict to unknown type `%s`" % ts)
    return _FROM[ts](d)


def to_json(A: ParOperator) -> str:
    return json.dumps(to_Dict(A))


def from_json(s: str) -> ParOperator:
    return from_Dict(json.loads(s))


class JsonPlan:
    """A plan compiled by the LIBRARY from the JSON text (po_plan_create_json): no host-side lowering.  ``param_ids`` lists, per
    parameter slot, the "id" of the leaf it belongs to; ``apply(x, params)`` takes ``{id: device array}`` (uuid ids as "UUID:<uuid>")."""

    def __init__(self, text: str, batch: int, engine=None, flags: int = 0):
        import ctypes as C
        from .engine import current_engine
        self.eng = engine or current_engine()
        lib = self.eng.lib
        h = C.c_void_p()
        lib.check(lib.po_plan_create_json(self.eng.ctx, text.encode(), int(batch), flags, C.byref(h)))
        self.handle, self.batch = h, int(batch)
        n = C.c_int32()
        lib.check(lib.po_plan_num_params(h, C.byref(n)))
        self.param_ids = []
        for s_ in range(n.value):
            buf = C.create_string_buffer(256)
            lib.check(lib.po_plan_param_id(h, s_, buf, len(buf)))
            self.param_ids.append(buf.value.decode())
        sz = C.c_size_t()
        lib.check(lib.po_plan_workspace_bytes(h, C.byref(sz)))
        self.ws_bytes, self.ws = sz.value, None
        dn, dt, rn, rt = C.c_int64(), C.c_int32(), C.c_int64(), C.c_int32()
        lib.check(lib.po_plan_domain(h, C.byref(dn), C.byref(dt)))
        lib.check(lib.po_plan_range(h, C.byref(rn), C.byref(rt)))
        self.Domain, self.Range, self.ddt, self.rdt = dn.value, rn.value, dt.value, rt.value

    def describe(self) -> str:
        import ctypes as C
        buf = C.create_string_buffer(1 << 16)
        self.eng.lib.check(self.eng.lib.po_plan_describe(self.handle, buf, len(buf)))
        return buf.value.decode()

    def apply(self, x, params: dict):
        import ctypes as C
        import torch
        from .engine import colmajor, is_colmajor
        from .operators import jtype
        eng = self.eng
        T = [Float32, Float64, ComplexF32, ComplexF64]
        if x.shape[0] != self.Domain or (1 if x.dim() == 1 else x.shape[1]) != self.batch or jtype(x.dtype) is not T[self.ddt]:
            raise ParException("operand does not match the plan's Domain / batch / DDT")
        keep = []
        for pid in self.param_ids:
            if pid not in params:
                raise ParException("no parameter for id %r" % pid)
            t = params[pid]
            keep.append(t if is_colmajor(t) else colmajor(t))
        arr = (C.c_void_p * max(1, len(keep)))(*[t.data_ptr() for t in keep])
        xd = x if is_colmajor(x) else colmajor(x)
        y = eng.empty_colmajor(self.Range, self.batch, T[self.rdt])
        if self.ws is None and self.ws_bytes:
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=eng.torch_device)
        eng.lib.check(eng.lib.po_plan_apply(self.handle, C.c_void_p(xd.data_ptr()), C.c_void_p(y.data_ptr()), arr, len(keep),
                                            C.c_void_p(self.ws.data_ptr()) if self.ws is not None else None, self.ws_bytes, eng.stream_ptr()))
        return y.reshape(-1) if x.dim() == 1 else y

    def destroy(self):
        if self.handle:
            self.eng.lib.po_plan_destroy(self.handle)
            self.handle = None
