"""ctypes binding of ``libparops.so`` (the C ABI declared in ``include/parops.h``).

The product path has exactly one backend: the sm_100a CUDA library built in-tree at
``csrc/libparops.so``.  If it is missing or cannot be loaded the import fails loudly --
there is no CPU fallback.  (``bind(path)`` exists so that ``tests/emul`` can point the very
same host code at the CPU *kernel emulator* build of the same sources; nothing in this
package ever does that on its own.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "csrc", "libparops.so")

PO_F32, PO_F64, PO_C64, PO_C128 = 0, 1, 2, 3

(PO_NODE_IDENTITY, PO_NODE_DFT, PO_NODE_RESTRICTION, PO_NODE_MATRIX, PO_NODE_DIAGONAL, PO_NODE_TENSOR, PO_NODE_KRON,
 PO_NODE_COMPOSE, PO_NODE_REPARTITION, PO_NODE_BIAS, PO_NODE_GELU) = range(11)

PO_PLAN_DEFAULT, PO_PLAN_NO_FUSION, PO_PLAN_NO_FASTPATH, PO_PLAN_NO_PIPELINE, PO_PLAN_NO_PEER = 0, 1, 2, 4, 8

PO_ERR_UNSUPPORTED = -6


class ParException(Exception):
    """Package exception type (src/ParCommon.jl:4-6); raised for every non-zero status."""

    def __init__(self, msg, status=None):
        super().__init__(msg)
        self.msg = msg
        self.status = status


class po_node(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("ddt", C.c_int32), ("rdt", C.c_int32), ("adjoint", C.c_int32), ("param_slot", C.c_int32),
        ("child_begin", C.c_int32), ("child_count", C.c_int32), ("aux_begin", C.c_int32), ("aux_count", C.c_int32),
        ("domain", C.c_int64), ("range", C.c_int64),
    ]


EXCHANGE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                          C.POINTER(C.c_size_t), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                          C.POINTER(C.c_size_t), C.c_void_p)
COLLECTIVE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                            C.c_void_p)

# name -> (restype, argtypes); every symbol include/parops.h declares
SIGNATURES = {
    "po_version": (C.c_int32, []),
    "po_last_error": (C.c_char_p, []),
    "po_status_string": (C.c_char_p, [C.c_int32]),
    "po_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "po_ctx_destroy": (C.c_int32, [C.c_void_p]),
    "po_ctx_device": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "po_ctx_device_status": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "po_plan_create": (C.c_int32, [C.c_void_p, C.POINTER(po_node), C.c_int32, C.POINTER(C.c_int32), C.c_int32,
                                   C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int64, C.c_uint32,
                                   C.POINTER(C.c_void_p)]),
    "po_plan_create_json": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int64, C.c_uint32, C.POINTER(C.c_void_p)]),
    "po_plan_num_params": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "po_plan_param_id": (C.c_int32, [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]),
    "po_plan_destroy": (C.c_int32, [C.c_void_p]),
    "po_plan_domain": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "po_plan_range": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "po_plan_workspace_bytes": (C.c_int32, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "po_plan_num_steps": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "po_plan_describe": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "po_plan_launch_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    "po_plan_algorithmic_bytes": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    "po_plan_set_profiling": (C.c_int32, [C.c_void_p, C.c_int32]),
    "po_plan_step_profile": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                         C.c_int32, C.POINTER(C.c_int32)]),
    "po_plan_pullback_step_profile": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32)]),
    "po_plan_apply": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "po_plan_apply_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "po_plan_apply_host_async": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "po_plan_tape_bytes": (C.c_int32, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "po_plan_apply_taped": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                        C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "po_plan_pullback": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "po_dft": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_restrict": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_repartition": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32,
                                   C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_kron_apply": (C.c_int32, [C.c_void_p, C.POINTER(po_node), C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32), C.c_int64,
                                  C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "po_add_gelu": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_block_epilogue": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_void_p]),
    "po_block_epilogue_pullback": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "po_local_size": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    "po_repartition_plan": (C.c_int32, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int32,
                                        C.POINTER(C.c_int32)]),
    "po_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "po_comm_init_rank": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "po_comm_init_all": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "po_comm_init_external": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, EXCHANGE_FN, COLLECTIVE_FN, C.c_void_p]),
    "po_comm_destroy": (C.c_int32, [C.c_void_p]),
    "po_comm_rank": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "po_bcast": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "po_reduce_sum": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "po_allreduce_sum": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
}


class Lib:
    """A loaded libparops with typed entry points and status checking."""

    def __init__(self, path: str):
        self.path = path
        if not os.path.exists(path):
            raise ParException(
                "libparops.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % path)
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, status: int):
        if status != 0:
            msg = self.po_last_error()
            msg = msg.decode() if msg else ""
            raise ParException("%s (%s)" % (msg, self.po_status_string(status).decode()), status)


_default = None


def bind(path: str) -> Lib:
    return Lib(path)


def default_lib() -> Lib:
    global _default
    if _default is None:
        _default = Lib(DEFAULT_LIB)
    return _default
