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

The reference's type registry (src/ASTSerialization.jl:16-27) leaves ParTensor and the distributed operators out; ParTensor
is registered here because its to_Dict / from_Dict pair exists (src/ParTensor.jl:119-151), the distributed ones stay out.
Unknown "type" / "T" values raise ``ParException`` with the reference's messages (src/ASTSerialization.jl:38-50).
"""
from __future__ import annotations

import json
import uuid

from ._lib import ParException
from .operators import (ComplexF32, ComplexF64, Float32, Float64, ParAdjoint, ParCompose, ParDFT, ParDiagonal, ParIdentity,
                        ParKron, ParMatrix, ParOperator, ParRestriction, ParTensor)

__all__ = ["to_Dict", "from_Dict", "to_json", "from_json"]

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
