// Reverse mode for a compiled plan (the Zygote/ChainRules surface of examples/fno.jl,
// SURVEY.md 3.5).  Filled in below po_plan_run_pullback.
#pragma once
#include "po_comm.cuh"

static int po_plan_run_pullback(po_plan* pl, const void* tape, const void* ybar, void* xbar, const void* const* params,
                                void* const* param_grads, int32_t n_params, void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)pl; (void)tape; (void)ybar; (void)xbar; (void)params; (void)param_grads; (void)n_params; (void)ws; (void)ws_bytes; (void)st;
  return po_fail(PO_ERR_UNSUPPORTED, "po_plan_pullback: not built yet");
}
