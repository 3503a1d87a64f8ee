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
    with pytest.raises(ParException, match="unknown type `ParReduce`"):
        parops.from_Dict({"type": "ParReduce"})                                                                       # :44-46 (not in the registry)
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


# ---- the wire format handed to the C side (po_plan_create_json): the library parses the reference's dictionaries itself ---------
def _ids(theta):
    return {parops.to_Dict(k)["id"]: v for k, v in theta.items()}


@pytest.mark.parametrize("adj", [False, True])
def test_json_plan_equals_host_lowered_plan(emul_engine, adj):
    from tests import parity_cases as pc
    eng = emul_engine
    with parops.use_engine(eng):
        S = _fno()
        theta = parops.gpu(parops.init(S, rng=5))
        A = S.adjoint() if adj else S
        text = parops.to_json(A)                       # ParAdjoint(ParCompose(...)) for adj: the C side pushes the adjoint down itself
        plan = parops.JsonPlan(text, 2, eng)
        assert plan.Domain == A.Domain and plan.Range == A.Range
        assert sorted(plan.param_ids) == sorted(_ids(theta))
        x = eng.to_device(pc.randn(np.random.default_rng(1), np.float32, A.Domain, 2))
        y_json = parops.cpu(plan.apply(x, _ids(theta)))
        y_host = parops.cpu((S(theta).adjoint() if adj else S(theta)) * x)
        assert pc.rel_l2(y_json, y_host) <= 1e-6
        # same fused plan: the JSON route reaches the same kernels as the host-lowered tree
        host_plan, _ = parops.get_plan(S(theta).adjoint() if adj else S(theta), 2, eng)
        assert plan.describe() == host_plan.describe()
        plan.destroy()


def test_json_adjoint_kron_rederives_the_application_order(emul_engine):
    """src/ParKron.jl:82: adjoint(K) = ParKron(R, D, P, adjoint.(ops), reverse(order)); the C-side reader pushes ParAdjoint down the same way
    (forward order [1, 3, 2] -> adjoint order [2, 3, 1]).  A ParKron dictionary WITHOUT "order" gets the constructor's heuristic order."""
    from tests import parity_cases as pc
    eng = emul_engine
    with parops.use_engine(eng):
        K = parops.ParKron(parops.compose(parops.ParRestriction(ComplexF32, 9, [(1, 4)]), parops.ParDFT(Float32, 16)), parops.ParDFT(ComplexF32, 8),
                           parops.ParDFT(ComplexF32, 4))
        Ka = K.adjoint()
        text = json.dumps({"type": "ParAdjoint", "of": parops.to_Dict(K)})
        plan = parops.JsonPlan(text, 3, eng)
        hp, _ = parops.get_plan(Ka, 3, eng)
        assert plan.describe() == hp.describe()
        y = eng.to_device(pc.randn(np.random.default_rng(2), np.complex64, Ka.Domain, 3))
        assert pc.rel_l2(parops.cpu(plan.apply(y, {})), parops.cpu(Ka * y)) <= 1e-6
        dk = parops.to_Dict(K)
        del dk["order"]
        p2 = parops.JsonPlan(json.dumps(dk), 3, eng)
        assert p2.describe() == parops.get_plan(K, 3, eng)[0].describe()


def test_json_distributed_node_types_round_trip_and_lower(emul_engine):
    """the three node types the reference's registry comments out (src/ASTSerialization.jl:17,21,25)"""
    from parops.distributed import CartComm
    eng = emul_engine
    with parops.use_engine(eng):
        R = parops.ParRepartition(ComplexF32, CartComm([1, 1], 0), CartComm([1, 1], 0), (6, 5))
        d = parops.to_Dict(R)
        assert d == {"type": "ParRepartition", "T": "ComplexF32", "global_size": [6, 5], "dims_in": [1, 1], "dims_out": [1, 1], "rank": 0}
        R2 = parops.from_Dict(d)
        assert R2.global_size == (6, 5) and R2.dims_in == [1, 1] and R2.local_size_in == R.local_size_in
        B = parops.ParBroadcasted(parops.ParMatrix(Float32, 3, 4, "W"))
        assert parops.to_Dict(B) == {"type": "ParBroadcasted", "of": {"type": "ParMatrix", "T": "Float32", "m": 3, "n": 4, "id": "W"}, "root": 0}
        assert isinstance(parops.from_Dict(parops.to_Dict(B)), parops.ParBroadcasted)
        Dn = parops.ParDistributed(parops.ParIdentity(Float32, 10), 1, 4)
        assert parops.to_Dict(Dn) == {"type": "ParDistributed", "of": {"type": "ParIdentity", "T": "Float32", "n": 10}, "rank": 1, "size": 4}
        assert parops.from_Dict(parops.to_Dict(Dn)).Domain == 3
        # one rank: the repartition is the self-overlap copy; ParBroadcasted delegates; ParDistributed(ParIdentity) is an identity of local size
        tree = {"type": "ParCompose", "of": [parops.to_Dict(parops.ParBroadcasted(parops.ParMatrix(ComplexF32, 7, 30, "M"))), d]}
        plan = parops.JsonPlan(json.dumps(tree), 2, eng)
        assert plan.param_ids == ["M"] and plan.Domain == 30 and plan.Range == 7
        rng = np.random.default_rng(3)
        M = (rng.standard_normal((7, 30)) + 1j * rng.standard_normal((7, 30))).astype(np.complex64)
        x = (rng.standard_normal((30, 2)) + 1j * rng.standard_normal((30, 2))).astype(np.complex64)
        y = parops.cpu(plan.apply(eng.to_device(x), {"M": eng.to_device(M)}))
        assert np.linalg.norm(y - M @ x) <= 1e-5 * np.linalg.norm(M @ x)
        Dc = parops.ParDistributed(parops.ParIdentity(ComplexF32, 10), 1, 4)   # applied AFTER the real DFT: complex identity
        kd = {"type": "ParKron", "of": [parops.to_Dict(Dc), {"type": "ParDFT", "T": "Float32", "n": 8, "m": 5}], "order": [2, 1]}
        pk = parops.JsonPlan(json.dumps(kd), 1, eng)
        assert pk.Domain == 3 * 8 and pk.Range == 3 * 5


@pytest.mark.parametrize("text,msg", [("{", "malformed JSON"), ('{"T": "Float32"}', "no 'type' key"), ('{"type": "ParFoo", "T": "Float32"}', "unknown type"),
                                      ('{"type": "ParIdentity", "T": "Float16", "n": 3}', "unknown data type"),
                                      ('{"type": "ParDFT", "T": "Float32", "n": 16, "m": 16}', "range mismatch"),
                                      ('{"type": "ParKron", "of": [{"type": "ParIdentity", "T": "Float32", "n": 3}], "order": [1, 2]}', "order")])
def test_json_plan_errors_carry_the_reference_messages(emul_engine, text, msg):
    with pytest.raises(ParException, match=msg):
        parops.JsonPlan(text, 1, emul_engine)


@pytest.mark.parametrize("seed", range(10))
def test_json_route_on_random_trees(emul_engine, seed):
    """The C-side reader (po_plan_create_json: parser, adjoint push-down, Kron order, lowering) against the host-side lowering on random Kron trees
    (tests/test_emul_parity.py::_random_kron): same step list, same numbers, for the tree and for ParAdjoint(tree); and the Python round trip
    from_json(to_json(A)) rebuilds an operator that lowers to the same plan."""
    from tests import parity_cases as pc
    from tests.test_emul_parity import _random_kron
    rng = np.random.default_rng(7000 + seed)
    T = [np.complex64, np.float32, np.complex128, np.float64][seed % 4]
    build = pc.kron_of(*_random_kron(rng, T))
    try:
        K = build(pc.ProductNS)
    except Exception:  # noqa: BLE001 -- a tree the constructor rejects (covered by test_random_kron_trees)
        return
    eng = emul_engine
    with parops.use_engine(eng):
        theta = parops.gpu(parops.init(K, rng=seed)) if K.P is parops.Parametric else {}
        for adj in (False, True):
            A = K.adjoint() if adj else K
            At = (K(theta) if theta else K)
            At = At.adjoint() if adj else At
            plan = parops.JsonPlan(parops.to_json(A), 2, eng)
            host_plan, _ = parops.get_plan(At, 2, eng)
            assert plan.describe() == host_plan.describe()
            x = eng.to_device(pc.randn(np.random.default_rng(seed), At.D.np, At.Domain, 2))
            y_json, y_host = parops.cpu(plan.apply(x, _ids(theta))), parops.cpu(At * x)
            assert np.array_equal(y_json, y_host)       # the same plan on the same data: bit-identical
            plan.destroy()
        K2 = parops.from_json(parops.to_json(K))
        assert parops.to_Dict(K2) == parops.to_Dict(K)
