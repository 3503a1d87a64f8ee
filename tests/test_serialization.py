"""JSON wire format (src/ASTSerialization.jl + per-type to_Dict / from_Dict): key names, value conventions, round trips and
the reference's error messages."""
import json
import uuid

import numpy as np
import pytest

import parops
from parops import ComplexF32, ComplexF64, Float32, Float64, ParException


def _fno():
    return parops.spectral_convolution(Float32, [16, 16, 8], [[(1, 4), (13, 16)], [(1, 4), (13, 16)], [(1, 4)]], 4)


def test_leaf_dicts_match_the_reference_keys():
    assert parops.to_Dict(parops.ParDFT(Float32, 16)) == {"type": "ParDFT", "T": "Float32", "n": 16, "m": 9}        # src/ParDFT.jl:34
    assert parops.to_Dict(parops.ParDFT(ComplexF64, 8)) == {"type": "ParDFT", "T": "ComplexF64", "n": 8, "m": 8}
    assert parops.to_Dict(parops.ParIdentity(Float64, 5)) == {"type": "ParIdentity", "T": "Float64", "n": 5}        # src/ParIdentity.jl:24
    assert parops.to_Dict(parops.ParRestriction(ComplexF32, 9, [(1, 3), (7, 9)])) == {
        "type": "ParRestriction", "T": "ComplexF32", "n": 9, "ranges": [[1, 3], [7, 9]]}                              # src/ParRestriction.jl:45
    u = uuid.uuid4()
    assert parops.to_Dict(parops.ParMatrix(Float32, 3, 4, u)) == {"type": "ParMatrix", "T": "Float32", "m": 3, "n": 4, "id": "UUID:%s" % u}
    assert parops.to_Dict(parops.ParMatrix(Float32, 3, 4, "W1"))["id"] == "W1"                                     # src/ParMatrix.jl:69-72
    assert parops.to_Dict(parops.ParDiagonal(ComplexF64, 6, "d"))  == {"type": "ParDiagonal", "T": "ComplexF64", "n": 6, "id": "d"}


def test_round_trip_preserves_structure_order_and_ids():
    S = _fno()
    d = parops.to_Dict(S)
    assert d["type"] == "ParCompose" and len(d["of"]) == 3                                                           # src/ParCompose.jl:100
    s = parops.to_json(S)
    S2 = parops.from_json(s)
    assert parops.to_Dict(S2) == json.loads(s)
    assert S2.Domain == S.Domain and S2.Range == S.Range and S2.D is S.D and S2.R is S.R

    def krons(op, out):
        if isinstance(op, parops.ParKron):
            out.append(op.order)
        for c in (op.children() if hasattr(op, "children") else []):
            krons(c, out)
        return out
    assert krons(S2, []) == krons(S, [])                                                                             # "order" kept as written, src/ParKron.jl:315
    # parameter identity survives: same ids => the same θ dict keys can be rebuilt
    m1 = [o for o in json.dumps(d).split('"id": "')[1:]]
    m2 = [o for o in parops.to_json(S2).split('"id": "')[1:]]
    assert [x.split('"')[0] for x in m1] == [x.split('"')[0] for x in m2]


def test_adjoint_tensor_and_errors():
    A = parops.ParAdjoint(parops.ParDFT(Float64, 12))
    assert parops.to_Dict(A) == {"type": "ParAdjoint", "of": {"type": "ParDFT", "T": "Float64", "n": 12, "m": 7}}  # src/ParAdjoint.jl:23
    B = parops.from_Dict(parops.to_Dict(A))
    assert isinstance(B, parops.ParAdjoint) and B.op.n == 12
    T = parops.ParTensor(ComplexF32, (1, 2, 3, 4, 5), (4, 3, 2, 3, 2), (2, 3, 4, 5), (3, 2, 3, 2), (1, 3, 4, 5), (4, 2, 3, 2), "W")
    dt = parops.to_Dict(T)
    assert set(dt) == {"type", "T", "ws", "wo", "is", "io", "to", "ts", "id"}                                        # src/ParTensor.jl:119-134
    T2 = parops.from_Dict(json.loads(json.dumps(dt)))
    assert T2.weight_shape == T.weight_shape and T2.input_shape == T.input_shape and T2.target_order == T.target_order and T2.id == "W"
    with pytest.raises(ParException, match="no 'type' key"):
        parops.from_Dict({"T": "Float32"})                                                                            # src/ASTSerialization.jl:39-41
    with pytest.raises(ParException, match="unknown type `ParRepartition`"):
        parops.from_Dict({"type": "ParRepartition"})                                                                  # :44-46 (registry leaves it out)
    with pytest.raises(ParException, match="unknown data type `Int32`"):
        parops.from_Dict({"type": "ParIdentity", "T": "Int32", "n": 3})
    with pytest.raises(ParException, match="range mismatch"):
        parops.from_Dict({"type": "ParDFT", "T": "Float32", "n": 16, "m": 16})                                       # src/ParDFT.jl:44-47


def test_deserialised_tree_applies_like_the_original(emul_engine):
    """from_json(to_json(S)) computes exactly what S computes (CPU kernel emulator; nothing structural is lost on the wire)."""
    from tests import parity_cases as pc
    S = parops.ParKron(parops.compose(parops.ParRestriction(ComplexF64, 5, [(1, 3)]), parops.ParDFT(Float64, 8)), parops.ParDFT(ComplexF64, 4))
    S2 = parops.from_json(parops.to_json(S))
    assert parops.to_json(S2) == parops.to_json(S)
    assert S2.order == S.order and S2.Domain == S.Domain and S2.Range == S.Range
    x = pc.randn(np.random.default_rng(0), np.float64, S.Domain, 2)
    with parops.use_engine(emul_engine):
        y1 = parops.cpu(S * emul_engine.to_device(x))
        y2 = parops.cpu(S2 * emul_engine.to_device(x))
    assert np.array_equal(y1, y2)
