// extern "C" surface of libparops.so (see include/parops.h for the contract).
#include "po_comm.cuh"
#include "po_pullback.cuh"
#include "po_block_tc.cuh"
#include "po_json.cuh"

#define PO_API extern "C" __attribute__((visibility("default")))

// NVTX ranges around every C-ABI entry point that enqueues work (the reference has no tracing; profilers
// see the library's calls by name).  nvtx3 is header-only and resolves its injection library lazily: free without a tool.
#ifndef PO_EMUL
#include <nvtx3/nvToolsExt.h>
struct po_nvtx_scope {
  explicit po_nvtx_scope(const char* name) { nvtxRangePushA(name); }
  ~po_nvtx_scope() { nvtxRangePop(); }
};
#define PO_NVTX(name) po_nvtx_scope _po_nvtx_scope_(name)
#else
#define PO_NVTX(name)
#endif

PO_API int32_t po_version(void) { return PO_VERSION; }
PO_API const char* po_last_error(void) { return g_po_error.c_str(); }
PO_API const char* po_status_string(int32_t status) {
  switch (status) {
    case PO_OK: return "ok";
    case PO_ERR_INVALID: return "invalid argument";
    case PO_ERR_SHAPE: return "Domain/Range mismatch";
    case PO_ERR_DTYPE: return "DDT/RDT mismatch";
    case PO_ERR_CUDA: return "CUDA error";
    case PO_ERR_NCCL: return "communicator error";
    case PO_ERR_UNSUPPORTED: return "unsupported";
    case PO_ERR_WORKSPACE: return "workspace too small";
  }
  return "unknown status";
}

PO_API int32_t po_ctx_create(int32_t device, po_ctx** out) {
  if (!out) return po_fail(PO_ERR_INVALID, "po_ctx_create: out is null");
  int ndev = 0;
  PO_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return po_fail(PO_ERR_INVALID, "po_ctx_create: device %d not in [0, %d)", device, ndev);
  PO_CUDA_CHECK(cudaSetDevice(device));
  po_ctx* c = new po_ctx();
  c->device = device;
#ifdef PO_EMUL
  c->num_sms = 2;
#else
  PO_CUDA_CHECK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
#endif
#ifdef PO_EMUL
  c->h_err = (volatile int*)calloc(1, sizeof(int));
  c->d_err = (int*)c->h_err;
#else
  void* herr = nullptr;
  PO_CUDA_CHECK(cudaHostAlloc(&herr, sizeof(int), cudaHostAllocMapped));
  *(int*)herr = 0;
  c->h_err = (volatile int*)herr;
  PO_CUDA_CHECK(cudaHostGetDevicePointer((void**)&c->d_err, herr, 0));
#endif
  *out = c;
  return PO_OK;
}

PO_API int32_t po_comm_destroy(po_ctx* ctx);

PO_API int32_t po_ctx_destroy(po_ctx* ctx) {
  if (!ctx) return PO_OK;
  for (auto& kv : ctx->plan_cache) { po_pullback_release(kv.second); delete kv.second; }
  ctx->plan_cache.clear();
  po_comm_destroy(ctx);
#ifdef PO_EMUL
  free((void*)ctx->h_err);
#else
  if (ctx->h_err) cudaFreeHost((void*)ctx->h_err);
#endif
  if (ctx->scratch) cudaFree(ctx->scratch);
  delete ctx;
  return PO_OK;
}

PO_API int32_t po_ctx_device_status(po_ctx* ctx, int32_t* status) {
  if (!ctx || !status) return po_fail(PO_ERR_INVALID, "null argument");
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  PO_CUDA_CHECK(cudaDeviceSynchronize());  // everything enqueued so far has had its chance to raise the word
  *status = ctx->h_err ? *ctx->h_err : 0;
  if (ctx->h_err) *ctx->h_err = 0;         // read-and-clear
  return PO_OK;
}

PO_API int32_t po_ctx_device(const po_ctx* ctx, int32_t* device) {
  if (!ctx || !device) return po_fail(PO_ERR_INVALID, "po_ctx_device: null argument");
  *device = ctx->device;
  return PO_OK;
}

// ---- plans --------------------------------------------------------------------------------
PO_API int32_t po_plan_create(po_ctx* ctx, const po_node* nodes, int32_t n_nodes, const int32_t* child_index,
                              int32_t n_child_index, const int64_t* aux, int32_t n_aux, int32_t root, int64_t batch,
                              uint32_t flags, po_plan** out) {
  PO_NVTX("po_plan_create");
  return po_plan_compile(ctx, nodes, n_nodes, child_index, n_child_index, aux, n_aux, root, batch, flags, out);
}

// The reference's JSON wire format (src/ASTSerialization.jl + per-type to_Dict) straight into a plan: see po_json.cuh
PO_API int32_t po_plan_create_json(po_ctx* ctx, const char* json, int64_t batch, uint32_t flags, po_plan** out) {
  PO_NVTX("po_plan_create_json");
  if (!ctx || !json || !out) return po_fail(PO_ERR_INVALID, "po_plan_create_json: null argument");
  po_jparser jp;
  jp.p = json; jp.end = json + strlen(json);
  po_jval root;
  if (!jp.parse(root)) return po_fail(PO_ERR_INVALID, "po_plan_create_json: malformed JSON (%s) at byte %lld", jp.err.c_str(), (long long)(jp.p - json));
  jp.ws();
  if (jp.p != jp.end) return po_fail(PO_ERR_INVALID, "po_plan_create_json: trailing characters after the document");
  std::unique_ptr<po_jop> tree;
  int rc = po_j_build(root, tree, 0);
  if (rc) return rc;
  po_jlower lw;
  lw.comm_rank = ctx->comm ? ctx->comm->rank : 0;
  int ridx = -1;
  if ((rc = lw.lower(*tree, false, ridx))) return rc;
  rc = po_plan_compile(ctx, lw.nodes.data(), (int32_t)lw.nodes.size(), lw.child.data(), (int32_t)lw.child.size(), lw.aux.data(), (int32_t)lw.aux.size(), ridx,
                       batch, flags, out);
  if (rc) return rc;
  (*out)->param_ids = lw.param_ids;
  return PO_OK;
}
PO_API int32_t po_plan_num_params(const po_plan* plan, int32_t* n_params) {
  if (!plan || !n_params) return po_fail(PO_ERR_INVALID, "null argument");
  *n_params = plan->n_params;
  return PO_OK;
}
PO_API int32_t po_plan_param_id(const po_plan* plan, int32_t slot, char* buf, size_t buf_bytes) {
  if (!plan || !buf || !buf_bytes) return po_fail(PO_ERR_INVALID, "null argument");
  if (slot < 0 || (size_t)slot >= plan->param_ids.size()) return po_fail(PO_ERR_INVALID, "parameter slot %d has no id (the plan was not created from JSON, or the slot is out of range)", slot);
  snprintf(buf, buf_bytes, "%s", plan->param_ids[(size_t)slot].c_str());
  return PO_OK;
}

PO_API int32_t po_plan_destroy(po_plan* plan) {
  if (!plan) return PO_OK;
  po_pullback_release(plan);
  for (auto& s : plan->steps) delete s.repart;
  delete plan;
  return PO_OK;
}

PO_API int32_t po_plan_domain(const po_plan* plan, int64_t* n, int32_t* dtype) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  if (n) *n = plan->domain;
  if (dtype) *dtype = plan->ddt;
  return PO_OK;
}
PO_API int32_t po_plan_range(const po_plan* plan, int64_t* n, int32_t* dtype) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  if (n) *n = plan->range;
  if (dtype) *dtype = plan->rdt;
  return PO_OK;
}
PO_API int32_t po_plan_workspace_bytes(const po_plan* plan, size_t* bytes) {
  if (!plan || !bytes) return po_fail(PO_ERR_INVALID, "null argument");
  *bytes = po_plan_ws_bytes(plan);
  return PO_OK;
}
PO_API int32_t po_plan_num_steps(const po_plan* plan, int32_t* n_steps) {
  if (!plan || !n_steps) return po_fail(PO_ERR_INVALID, "null argument");
  *n_steps = (int32_t)plan->steps.size();
  return PO_OK;
}
PO_API int32_t po_plan_describe(const po_plan* plan, char* buf, size_t buf_bytes) {
  if (!plan || !buf || !buf_bytes) return po_fail(PO_ERR_INVALID, "null argument");
  std::string out;
  char line[512];
  for (size_t i = 0; i < plan->steps.size(); ++i) {
    const po_step& s = plan->steps[i];
    snprintf(line, sizeof(line), "%zu: %s [%s->%s] I=%lld n=%lld m=%lld O=%lld\n", i, s.note.c_str(), po_dtype_name(s.in_dt),
             po_dtype_name(s.out_dt), (long long)s.I, (long long)s.n, (long long)s.m, (long long)s.O);
    out += line;
  }
  snprintf(buf, buf_bytes, "%s", out.c_str());
  return PO_OK;
}
PO_API int32_t po_plan_launch_count(const po_plan* plan, int64_t* launches) {
  if (!plan || !launches) return po_fail(PO_ERR_INVALID, "null argument");
  *launches = plan->launches;
  return PO_OK;
}
PO_API int32_t po_plan_algorithmic_bytes(const po_plan* plan, int64_t* bytes) {
  if (!plan || !bytes) return po_fail(PO_ERR_INVALID, "null argument");
  int64_t b = plan->domain * plan->batch * (int64_t)po_dtype_size(plan->ddt) + plan->range * plan->batch * (int64_t)po_dtype_size(plan->rdt);
  for (const po_step& s : plan->steps) {
    if (s.kind == ST_MATRIX || s.kind == ST_MIXDIAG) b += s.prm_rows * s.prm_cols * (int64_t)po_dtype_size(s.in_dt);
    if (s.kind == ST_MIXDIAG) b += s.nd * (int64_t)po_dtype_size(s.in_dt);
    if (s.kind == ST_DIAG || s.kind == ST_BIAS) b += s.n * (int64_t)po_dtype_size(s.in_dt);
    if (s.kind == ST_TENSOR) b += (int64_t)s.ic * s.oc * s.M * (int64_t)po_dtype_size(s.in_dt);
  }
  *bytes = b;
  return PO_OK;
}

PO_API int32_t po_plan_set_profiling(po_plan* plan, int32_t enable) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  int rc = po_plan_collect_profile(plan);
  if (rc) return rc;
  plan->profiling = enable != 0;
  plan->step_ms.assign(plan->steps.size(), 0.0);
  plan->step_calls.assign(plan->steps.size(), 0);
  plan->pb_step_ms.assign(plan->steps.size(), 0.0);
  plan->pb_step_calls.assign(plan->steps.size(), 0);
  return PO_OK;
}
PO_API int32_t po_plan_pullback_step_profile(po_plan* plan, double* ms_sum, int64_t* calls, int32_t capacity, int32_t* n_steps) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  int rc = po_plan_collect_profile(plan);
  if (rc) return rc;
  const int32_t n = (int32_t)plan->steps.size();
  if (n_steps) *n_steps = n;
  for (int32_t i = 0; i < n && i < capacity; ++i) {
    if (ms_sum) ms_sum[i] = plan->pb_step_ms[(size_t)i];
    if (calls) calls[i] = plan->pb_step_calls[(size_t)i];
  }
  return PO_OK;
}
PO_API int32_t po_plan_step_profile(po_plan* plan, double* ms_sum, int64_t* calls, int64_t* bytes, int32_t capacity, int32_t* n_steps) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  int rc = po_plan_collect_profile(plan);
  if (rc) return rc;
  const int32_t n = (int32_t)plan->steps.size();
  if (n_steps) *n_steps = n;
  for (int32_t i = 0; i < n && i < capacity; ++i) {
    const po_step& s = plan->steps[(size_t)i];
    if (ms_sum) ms_sum[i] = plan->step_ms[(size_t)i];
    if (calls) calls[i] = plan->step_calls[(size_t)i];
    if (bytes) {
      int64_t b = s.I * s.n * s.O * (int64_t)po_dtype_size(s.in_dt) + po_step_out_elems(s) * (int64_t)po_dtype_size(s.out_dt);
      if (s.kind == ST_MATRIX || s.kind == ST_MIXDIAG) b += s.prm_rows * s.prm_cols * (int64_t)po_dtype_size(s.in_dt);
      if (s.kind == ST_MIXDIAG) b += s.nd * (int64_t)po_dtype_size(s.in_dt);
      if (s.kind == ST_DIAG || s.kind == ST_BIAS) b += s.n * (int64_t)po_dtype_size(s.in_dt);
      if (s.kind == ST_TENSOR) b += (int64_t)s.ic * s.oc * s.M * (int64_t)po_dtype_size(s.in_dt);
      bytes[i] = b;
    }
  }
  return PO_OK;
}

PO_API int32_t po_plan_apply(po_plan* plan, const void* x, void* y, const void* const* params, int32_t n_params,
                             void* workspace, size_t workspace_bytes, void* stream) {
  PO_NVTX("po_plan_apply");
  return po_plan_run(plan, x, y, params, n_params, workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}

// H2D of x, the apply and the D2H of y, all stream-ordered on `st`; the plan owns the device-side x / y / workspace (one set per
// plan: successive host applies of one plan must use one stream, which orders their reuse)
static int po_plan_enqueue_host(po_plan* plan, const void* x_host, void* y_host, const void* const* params, int32_t n_params, cudaStream_t st) {
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  if (plan->batch == 0) return PO_OK;  // (Domain, 0) -> (Range, 0): nothing to move or compute
  if (!x_host || !y_host) return po_fail(PO_ERR_INVALID, "null argument");
  PO_CUDA_CHECK(cudaSetDevice(plan->ctx->device));
  const size_t xb = (size_t)(plan->domain * plan->batch) * po_dtype_size(plan->ddt);
  const size_t yb = (size_t)(plan->range * plan->batch) * po_dtype_size(plan->rdt);
  const size_t wb = po_plan_ws_bytes(plan);
  if (!plan->h_x && xb) PO_CUDA_CHECK(cudaMalloc(&plan->h_x, xb));
  if (!plan->h_y && yb) PO_CUDA_CHECK(cudaMalloc(&plan->h_y, yb));
  if (!plan->h_ws && wb) PO_CUDA_CHECK(cudaMalloc(&plan->h_ws, wb));
  if (xb) PO_CUDA_CHECK(cudaMemcpyAsync(plan->h_x, x_host, xb, cudaMemcpyHostToDevice, st));
  int rc = po_plan_run(plan, plan->h_x, plan->h_y, params, n_params, plan->h_ws, wb, st, nullptr);
  if (rc) return rc;
  if (yb) PO_CUDA_CHECK(cudaMemcpyAsync(y_host, plan->h_y, yb, cudaMemcpyDeviceToHost, st));
  return PO_OK;
}

PO_API int32_t po_plan_apply_host(po_plan* plan, const void* x_host, void* y_host, const void* const* params,
                                  int32_t n_params, void* stream) {
  PO_NVTX("po_plan_apply_host");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = po_plan_enqueue_host(plan, x_host, y_host, params, n_params, st);
  if (rc) return rc;
  PO_CUDA_CHECK(cudaStreamSynchronize(st));
  return po_ctx_check(plan->ctx);  // the stream has drained: a timed-out wait of THIS apply is reported by this call
}

PO_API int32_t po_plan_apply_host_async(po_plan* plan, const void* x_host, void* y_host, const void* const* params,
                                        int32_t n_params, void* stream) {
  PO_NVTX("po_plan_apply_host_async");
  return po_plan_enqueue_host(plan, x_host, y_host, params, n_params, (cudaStream_t)stream);
}

// ---- reverse mode ---------------------------------------------------------------------------
PO_API int32_t po_plan_tape_bytes(const po_plan* plan, size_t* bytes) {
  if (!plan || !bytes) return po_fail(PO_ERR_INVALID, "null argument");
  *bytes = po_tape_bytes(plan);
  return PO_OK;
}
PO_API int32_t po_plan_apply_taped(po_plan* plan, const void* x, void* y, const void* const* params, int32_t n_params,
                                   void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  PO_NVTX("po_plan_apply_taped");
  if (!plan) return po_fail(PO_ERR_INVALID, "null plan");
  if (tape_bytes < po_tape_bytes(plan) || (!tape && po_tape_bytes(plan))) return po_fail(PO_ERR_WORKSPACE, "tape too small: need %zu bytes", po_tape_bytes(plan));
  return po_plan_run(plan, x, y, params, n_params, workspace, workspace_bytes, (cudaStream_t)stream, tape);
}
PO_API int32_t po_plan_pullback(po_plan* plan, const void* tape, const void* ybar, void* xbar, const void* const* params,
                                void* const* param_grads, int32_t n_params, void* workspace, size_t workspace_bytes,
                                void* stream) {
  PO_NVTX("po_plan_pullback");
  return po_plan_run_pullback(plan, tape, ybar, xbar, params, param_grads, n_params, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- single-operator conveniences -------------------------------------------------------------
static int po_cached_plan(po_ctx* ctx, const std::string& key, const po_node& nd, const std::vector<int64_t>& aux, int64_t batch, po_plan** out) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->plan_cache.find(key);
  if (it != ctx->plan_cache.end()) { *out = it->second; return PO_OK; }
  po_plan* pl = nullptr;
  int rc = po_plan_compile(ctx, &nd, 1, nullptr, 0, aux.data(), (int32_t)aux.size(), 0, batch, PO_PLAN_DEFAULT, &pl);
  if (rc) return rc;
  ctx->plan_cache[key] = pl;
  *out = pl;
  return PO_OK;
}

PO_API int32_t po_dft(po_ctx* ctx, int32_t in_dtype, int64_t n, int32_t adjoint, int64_t batch, const void* x, void* y, void* stream) {
  PO_NVTX("po_dft");
  if (!ctx) return po_fail(PO_ERR_INVALID, "null ctx");
  if (in_dtype < 0 || in_dtype > 3) return po_fail(PO_ERR_DTYPE, "po_dft: bad dtype");
  po_node nd;
  memset(&nd, 0, sizeof(nd));
  nd.kind = PO_NODE_DFT; nd.adjoint = adjoint ? 1 : 0; nd.param_slot = -1; nd.aux_begin = 0; nd.aux_count = 1;
  // in_dtype is the dtype of x.  forward: real -> r2c, complex -> c2c.  adjoint: x is the
  // spectrum (complex); n is always the transform length; a real-DFT adjoint is requested by
  // passing the REAL dtype of the original operator's domain with adjoint=1 and complex x.
  if (!adjoint) {
    nd.ddt = in_dtype;
    nd.rdt = po_dtype_complex_of(in_dtype);
    nd.domain = n;
    nd.range = po_dtype_is_complex(in_dtype) ? n : n / 2 + 1;
  } else {
    nd.rdt = in_dtype;  // dtype of the ORIGINAL domain = output of the adjoint
    nd.ddt = po_dtype_complex_of(in_dtype);
    nd.range = n;
    nd.domain = po_dtype_is_complex(in_dtype) ? n : n / 2 + 1;
  }
  std::vector<int64_t> aux(1, n);
  char key[128];
  snprintf(key, sizeof(key), "dft:%d:%lld:%d:%lld", in_dtype, (long long)n, adjoint, (long long)batch);
  po_plan* pl = nullptr;
  int rc = po_cached_plan(ctx, key, nd, aux, batch, &pl);
  if (rc) return rc;
  return po_plan_run(pl, x, y, nullptr, 0, nullptr, 0, (cudaStream_t)stream, nullptr);
}

PO_API int32_t po_restrict(po_ctx* ctx, int32_t dtype, int64_t n, const int64_t* ranges, int32_t n_ranges, int32_t adjoint,
                           int64_t batch, const void* x, void* y, void* stream) {
  PO_NVTX("po_restrict");
  if (!ctx || (!ranges && n_ranges)) return po_fail(PO_ERR_INVALID, "null argument");
  po_node nd;
  memset(&nd, 0, sizeof(nd));
  nd.kind = PO_NODE_RESTRICTION; nd.adjoint = adjoint ? 1 : 0; nd.param_slot = -1; nd.ddt = nd.rdt = dtype;
  std::vector<int64_t> aux;
  aux.push_back(n);
  aux.push_back(n_ranges);
  int64_t total = 0;
  std::string key = "restrict:";
  for (int r = 0; r < n_ranges; ++r) {
    aux.push_back(ranges[2 * r]);
    aux.push_back(ranges[2 * r + 1]);
    if (ranges[2 * r + 1] >= ranges[2 * r]) total += ranges[2 * r + 1] - ranges[2 * r] + 1;
    key += std::to_string(ranges[2 * r]) + "-" + std::to_string(ranges[2 * r + 1]) + ",";
  }
  nd.aux_begin = 0; nd.aux_count = (int32_t)aux.size();
  nd.domain = adjoint ? total : n;
  nd.range = adjoint ? n : total;
  key += ":" + std::to_string(dtype) + ":" + std::to_string(n) + ":" + std::to_string(adjoint) + ":" + std::to_string(batch);
  po_plan* pl = nullptr;
  int rc = po_cached_plan(ctx, key, nd, aux, batch, &pl);
  if (rc) return rc;
  return po_plan_run(pl, x, y, nullptr, 0, nullptr, 0, (cudaStream_t)stream, nullptr);
}

// ParRepartition / its adjoint (= the same exchange with the grids swapped, src/ParRepartition.jl:200-201) as a cached one-node plan
PO_API int32_t po_repartition(po_ctx* ctx, int32_t dtype, int32_t ndim, const int64_t* global_size, const int32_t* dims_in, const int32_t* dims_out,
                              int32_t rank, int32_t adjoint, int64_t batch, const void* x, void* y, void* stream) {
  PO_NVTX("po_repartition");
  if (!ctx || !global_size || !dims_in || !dims_out) return po_fail(PO_ERR_INVALID, "null argument");
  if (dtype < 0 || dtype > 3) return po_fail(PO_ERR_DTYPE, "po_repartition: bad dtype");
  if (ndim < 1 || ndim > 8) return po_fail(PO_ERR_INVALID, "po_repartition: bad ndim");
  const int32_t* di = adjoint ? dims_out : dims_in;
  const int32_t* dq = adjoint ? dims_in : dims_out;
  po_repart_info info;
  int rc = po_repart_build(ndim, global_size, di, dq, rank, info);
  if (rc) return rc;
  po_node nd;
  memset(&nd, 0, sizeof(nd));
  nd.kind = PO_NODE_REPARTITION; nd.param_slot = -1; nd.ddt = nd.rdt = dtype;
  nd.domain = nd.range = 1;
  for (int d = 0; d < ndim; ++d) { nd.domain *= info.local_in[(size_t)d]; nd.range *= info.local_out[(size_t)d]; }
  std::vector<int64_t> aux;
  aux.push_back(ndim);
  for (int d = 0; d < ndim; ++d) aux.push_back(global_size[d]);
  for (int d = 0; d < ndim; ++d) aux.push_back(di[d]);
  for (int d = 0; d < ndim; ++d) aux.push_back(dq[d]);
  aux.push_back(rank);
  nd.aux_begin = 0; nd.aux_count = (int32_t)aux.size();
  std::string key = "repart:" + std::to_string(dtype) + ":" + std::to_string(batch);
  for (int64_t v : aux) key += ":" + std::to_string(v);
  po_plan* pl = nullptr;
  rc = po_cached_plan(ctx, key, nd, aux, batch, &pl);
  if (rc) return rc;
  return po_plan_run(pl, x, y, nullptr, 0, nullptr, 0, (cudaStream_t)stream, nullptr);
}

// (A_1 (x) ... (x) A_N) * x for leaf factors, src/ParKron.jl:190-220: `factors` are leaf nodes exactly as po_plan_create takes them
// (their aux_begin index `aux`), `order` the 1-based application order (src/ParKron.jl:33-57; the host computes it).  The plan and its
// workspace are cached on the context.
PO_API int32_t po_kron_apply(po_ctx* ctx, const po_node* factors, int32_t n_factors, const int64_t* aux, int32_t n_aux, const int32_t* order,
                             int64_t batch, const void* x, void* y, const void* const* params, int32_t n_params, void* stream) {
  PO_NVTX("po_kron_apply");
  if (!ctx || !factors || !order || n_factors < 1 || (n_aux && !aux)) return po_fail(PO_ERR_INVALID, "null argument");
  std::vector<po_node> nodes(factors, factors + n_factors);
  std::vector<int64_t> a(aux, aux + n_aux);
  std::vector<int32_t> kids((size_t)n_factors);
  po_node root;
  memset(&root, 0, sizeof(root));
  root.kind = PO_NODE_KRON; root.param_slot = -1; root.child_begin = 0; root.child_count = n_factors;
  root.aux_begin = (int32_t)a.size(); root.aux_count = n_factors;
  root.domain = root.range = 1;
  std::string key = "kron:" + std::to_string(batch);
  for (int32_t k = 0; k < n_factors; ++k) {
    kids[(size_t)k] = k;
    if (factors[k].kind == PO_NODE_KRON || factors[k].kind == PO_NODE_COMPOSE) return po_fail(PO_ERR_UNSUPPORTED, "po_kron_apply: factors must be leaves (use po_plan_create for trees)");
    root.domain *= factors[k].domain; root.range *= factors[k].range;
    a.push_back(order[k]);
    const po_node& f = factors[k];
    key += "|" + std::to_string(f.kind) + "," + std::to_string(f.ddt) + "," + std::to_string(f.rdt) + "," + std::to_string(f.adjoint) + "," + std::to_string(f.param_slot) +
           "," + std::to_string(f.domain) + "," + std::to_string(f.range) + "," + std::to_string(order[k]);
    if (f.aux_begin < 0 || f.aux_count < 0 || f.aux_begin + f.aux_count > n_aux) return po_fail(PO_ERR_INVALID, "po_kron_apply: factor %d: aux slice out of range", k);
    for (int32_t j = 0; j < f.aux_count; ++j) key += ":" + std::to_string(aux[f.aux_begin + j]);
  }
  if (order[0] < 1 || order[0] > n_factors || order[n_factors - 1] < 1 || order[n_factors - 1] > n_factors) return po_fail(PO_ERR_INVALID, "po_kron_apply: order is not a permutation of 1..N");
  root.ddt = factors[order[0] - 1].ddt;               // D = DDT(first applied), R = RDT(last applied)  (src/ParKron.jl:59-60)
  root.rdt = factors[order[n_factors - 1] - 1].rdt;
  nodes.push_back(root);
  po_plan* pl = nullptr;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->plan_cache.find(key);
    if (it != ctx->plan_cache.end()) pl = it->second;
    else {
      int rc = po_plan_compile(ctx, nodes.data(), (int32_t)nodes.size(), kids.data(), n_factors, a.data(), (int32_t)a.size(), n_factors, batch, PO_PLAN_DEFAULT, &pl);
      if (rc) return rc;
      ctx->plan_cache[key] = pl;
    }
  }
  const size_t wb = po_plan_ws_bytes(pl);
  if (wb && !pl->h_ws) PO_CUDA_CHECK(cudaMalloc(&pl->h_ws, wb));
  return po_plan_run(pl, x, y, params, n_params, pl->h_ws, wb, (cudaStream_t)stream, nullptr);
}

PO_API int32_t po_add_gelu(po_ctx* ctx, int32_t dtype, int64_t n, const void* x1, const void* x2, void* y, void* stream) {
  PO_NVTX("po_add_gelu");
  if (!ctx || !x1 || !y) return po_fail(PO_ERR_INVALID, "null argument");
  if (dtype != PO_F32 && dtype != PO_F64) return po_fail(PO_ERR_DTYPE, "gelu: real element types only");
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaError_t e = po_launch_gelu(dtype, x1, x2, y, n, (cudaStream_t)stream);
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "gelu launch failed: %s", cudaGetErrorString(e));
  return PO_OK;
}

PO_API int32_t po_block_epilogue(po_ctx* ctx, int32_t dtype, int64_t c_in, int64_t c_out, int64_t npts, const void* W, const void* x,
                                 const void* s, const void* b, int32_t apply_gelu, void* y, void* stream) {
  PO_NVTX("po_block_epilogue");
  if (!ctx || !W || !x || !y) return po_fail(PO_ERR_INVALID, "null argument");
  if (dtype != PO_F32 && dtype != PO_F64) return po_fail(PO_ERR_DTYPE, "block epilogue: real element types only");
  if (c_in < 1 || c_out < 1 || npts < 0 || c_in > 0x7fffffffLL || c_out > 0x7fffffffLL) return po_fail(PO_ERR_SHAPE, "block epilogue: bad extents");
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaError_t e = cudaErrorNotSupported;
  if (dtype == PO_F32) e = po_launch_block_epilogue_tc(W, x, s, b, y, c_in, c_out, npts, apply_gelu, ctx->num_sms, ctx->d_err, (cudaStream_t)stream);
  if (e == cudaErrorNotSupported) e = po_launch_block_epilogue(dtype, W, x, s, b, y, c_in, c_out, npts, apply_gelu, ctx->num_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "block epilogue launch failed: %s", cudaGetErrorString(e));
  return PO_OK;
}

// small transposing copy of a parameter matrix (rows x cols column-major -> cols x rows column-major)
template <class T>
__global__ void po_transpose_small_kernel(const T* __restrict__ A, T* __restrict__ At, int rows, int cols) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols; e += gridDim.x * blockDim.x) {
    const int r = e % rows, c = e / rows;
    At[c + cols * r] = A[e];
  }
}

// Reverse mode of the block epilogue (examples/fno.jl:40-44 under Zygote): y = act.(s .+ (I (x) W) x .+ b).  Nothing is taped: the
// pre-activation is recomputed by the forward kernel (act off) into `sbar`, which then becomes g = act'(z) .* ybar = the cotangent of s.
PO_API int32_t po_block_epilogue_pullback(po_ctx* ctx, int32_t dtype, int64_t c_in, int64_t c_out, int64_t npts, const void* W, const void* x,
                                          const void* s, const void* b, int32_t apply_gelu, const void* ybar, void* sbar, void* xbar, void* Wbar,
                                          void* bbar, void* stream) {
  PO_NVTX("po_block_epilogue_pullback");
  if (!ctx || !W || !x || !ybar || !sbar) return po_fail(PO_ERR_INVALID, "null argument");
  if (dtype != PO_F32 && dtype != PO_F64) return po_fail(PO_ERR_DTYPE, "block epilogue: real element types only");
  if (c_in < 1 || c_out < 1 || npts < 0 || c_in > 4096 || c_out > 4096) return po_fail(PO_ERR_SHAPE, "block epilogue: bad extents");
  if (int rc0 = po_ctx_check(ctx)) return rc0;
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = po_dtype_size(dtype);
  const i64 tot = c_out * npts;
  cudaError_t e = cudaSuccess;
  if (apply_gelu) {
    int rc = po_block_epilogue(ctx, dtype, c_in, c_out, npts, W, x, s, b, 0, sbar, stream);  // z = s + W x + b
    if (rc) return rc;
    if (tot) {
      if (dtype == PO_F32) { PO_LAUNCH((po_gelu_grad_kernel<float>), dim3(po_grid_1d(tot, 256)), dim3(256), 0, st, (const float*)sbar, (const float*)ybar, (float*)sbar, tot); }
      else { PO_LAUNCH((po_gelu_grad_kernel<double>), dim3(po_grid_1d(tot, 256)), dim3(256), 0, st, (const double*)sbar, (const double*)ybar, (double*)sbar, tot); }
      e = cudaGetLastError();
    }
  } else if (tot && sbar != ybar) {
    e = cudaMemcpyAsync(sbar, ybar, (size_t)tot * esz, cudaMemcpyDeviceToDevice, st);
  }
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "block epilogue pullback (activation) failed: %s", cudaGetErrorString(e));
  if (xbar) {  // xbar = (I (x) W') g: the same one-pass kernel with the transposed weight
    if (!ctx->scratch) PO_CUDA_CHECK(cudaMalloc(&ctx->scratch, (size_t)1 << 20));
    if ((size_t)(c_in * c_out) * esz > ((size_t)1 << 20)) return po_fail(PO_ERR_UNSUPPORTED, "block epilogue pullback: weight larger than the 1 MB scratch");
    if (dtype == PO_F32) { PO_LAUNCH((po_transpose_small_kernel<float>), dim3(8), dim3(256), 0, st, (const float*)W, (float*)ctx->scratch, (int)c_out, (int)c_in); }
    else { PO_LAUNCH((po_transpose_small_kernel<double>), dim3(8), dim3(256), 0, st, (const double*)W, (double*)ctx->scratch, (int)c_out, (int)c_in); }
    int rc = po_block_epilogue(ctx, dtype, c_out, c_in, npts, ctx->scratch, sbar, nullptr, nullptr, 0, xbar, stream);
    if (rc) return rc;
  }
  if (Wbar) {  // Wbar = sum over points of g x'
    e = po_launch_gram(dtype, sbar, x, Wbar, 1, c_out, c_in, npts, ctx->num_sms, st);
    if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "block epilogue pullback (weight cotangent) failed: %s", cudaGetErrorString(e));
  }
  if (bbar) {  // bbar = sum over points of g
    e = po_launch_modesum(dtype, nullptr, sbar, bbar, 1, c_out, npts, st);
    if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "block epilogue pullback (bias cotangent) failed: %s", cudaGetErrorString(e));
  }
  return PO_OK;
}

// ---- host integer bookkeeping --------------------------------------------------------------------
PO_API int64_t po_local_size(int64_t global_size, int64_t rank, int64_t num_ranks) {
  if (num_ranks <= 0) return -1;
  return po_local_size_impl(global_size, rank, num_ranks);
}

PO_API int32_t po_repartition_plan(int32_t ndim, const int64_t* global_size, const int32_t* dims_in, const int32_t* dims_out,
                                   int32_t rank, int64_t* local_in, int64_t* local_out, int64_t* send_out,
                                   int32_t send_capacity, int32_t* n_send, int64_t* recv_out, int32_t recv_capacity,
                                   int32_t* n_recv) {
  if (!global_size || !dims_in || !dims_out || !n_send || !n_recv) return po_fail(PO_ERR_INVALID, "null argument");
  po_repart_info info;
  int rc = po_repart_build(ndim, global_size, dims_in, dims_out, rank, info);
  if (rc) return rc;
  for (int d = 0; d < ndim; ++d) {
    if (local_in) local_in[d] = info.local_in[(size_t)d];
    if (local_out) local_out[d] = info.local_out[(size_t)d];
  }
  *n_send = (int32_t)info.send.size();
  *n_recv = (int32_t)info.recv.size();
  const int rec = 1 + 2 * ndim;
  auto emit = [&](const std::vector<po_box_rec>& v, int64_t* out, int32_t cap) {
    if (!out) return;
    for (size_t k = 0; k < v.size() && (int32_t)k < cap; ++k) {
      out[k * rec] = v[k].peer;
      for (int d = 0; d < ndim; ++d) { out[k * rec + 1 + 2 * d] = v[k].lo[(size_t)d]; out[k * rec + 2 + 2 * d] = v[k].hi[(size_t)d]; }
    }
  };
  emit(info.send, send_out, send_capacity);
  emit(info.recv, recv_out, recv_capacity);
  return PO_OK;
}

// ---- communicator -------------------------------------------------------------------------------
PO_API int32_t po_comm_unique_id(void* id_128_bytes) {
  if (!id_128_bytes) return po_fail(PO_ERR_INVALID, "null id");
  int rc = po_nccl_load();
  if (rc) return rc;
#ifndef PO_EMUL
  PO_NCCL_CHECK(g_nccl.GetUniqueId(id_128_bytes));
#endif
  return PO_OK;
}

PO_API int32_t po_comm_init_rank(po_ctx* ctx, int32_t n_ranks, int32_t rank, const void* id_128_bytes) {
  if (!ctx || !id_128_bytes) return po_fail(PO_ERR_INVALID, "null argument");
  if (ctx->comm) return po_fail(PO_ERR_INVALID, "context already has a communicator");
  int rc = po_nccl_load();
  if (rc) return rc;
#ifndef PO_EMUL
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  po_nccl_id id;
  memcpy(&id, id_128_bytes, sizeof(id));
  void* comm = nullptr;
  PO_NCCL_CHECK(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
  po_comm_state* cs = new po_comm_state();
  cs->n_ranks = n_ranks; cs->rank = rank; cs->nccl_comm = comm;
  ctx->comm = cs;
  rc = po_comm_setup_peer(ctx);  // NVLink peer-store path for ParRepartition (collective; falls back to NCCL on any failure)
  if (rc) return rc;
#endif
  return PO_OK;
}

// Single-process variant: one context per device of THIS process (a Julia session driving several GPUs with CUDA.device!).  The
// symmetric buffers are mapped by peer access instead of IPC; ParRepartition runs through the NVLink peer-store kernel only (a
// grouped NCCL exchange issued rank after rank from one thread would deadlock), parameter collectives need one thread per device.
PO_API int32_t po_comm_init_all(int32_t n_ctx, po_ctx* const* ctxs) {
  if (!ctxs || n_ctx < 1) return po_fail(PO_ERR_INVALID, "null argument");
  for (int r = 0; r < n_ctx; ++r) {
    if (!ctxs[r]) return po_fail(PO_ERR_INVALID, "context %d is null", r);
    if (ctxs[r]->comm) return po_fail(PO_ERR_INVALID, "context %d already has a communicator", r);
    for (int q = 0; q < r; ++q) if (ctxs[q]->device == ctxs[r]->device) return po_fail(PO_ERR_INVALID, "contexts %d and %d share device %d", q, r, ctxs[r]->device);
  }
  if (n_ctx > 64) return po_fail(PO_ERR_UNSUPPORTED, "at most 64 ranks");
#ifdef PO_EMUL
  return po_fail(PO_ERR_UNSUPPORTED, "po_comm_init_all needs GPUs (use po_comm_init_external in the CPU emulator)");
#else
  const char* mb = getenv("PO_SYM_MB");
  const size_t S = (size_t)(mb ? atoi(mb) : 256) << 20;
  std::vector<char*> sym((size_t)n_ctx, nullptr);
  auto cleanup = [&]() {
    for (int r = 0; r < n_ctx; ++r) {
      if (sym[(size_t)r]) { cudaSetDevice(ctxs[r]->device); cudaFree(sym[(size_t)r]); }
      if (ctxs[r]->comm) { if (ctxs[r]->comm->d_peer_sym) cudaFree(ctxs[r]->comm->d_peer_sym); delete ctxs[r]->comm; ctxs[r]->comm = nullptr; }
    }
  };
  for (int r = 0; r < n_ctx; ++r) {
    if (cudaSetDevice(ctxs[r]->device) != cudaSuccess || cudaMalloc((void**)&sym[(size_t)r], S) != cudaSuccess || cudaMemset(sym[(size_t)r], 0, S) != cudaSuccess) {
      cleanup();
      return po_fail(PO_ERR_CUDA, "po_comm_init_all: cannot allocate the symmetric buffer on device %d", ctxs[r]->device);
    }
    for (int q = 0; q < n_ctx; ++q) {
      if (q == r) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device);
      cudaError_t e = can ? cudaDeviceEnablePeerAccess(ctxs[q]->device, 0) : cudaErrorPeerAccessUnsupported;
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      if (e != cudaSuccess) { cudaGetLastError(); cleanup(); return po_fail(PO_ERR_UNSUPPORTED, "po_comm_init_all: no peer access from device %d to device %d", ctxs[r]->device, ctxs[q]->device); }
    }
  }
  for (int r = 0; r < n_ctx; ++r) {
    po_comm_state* cs = new po_comm_state();
    cs->n_ranks = n_ctx; cs->rank = r; cs->single_process = true;
    cs->sym = sym[(size_t)r]; cs->sym_bytes = S; cs->sym_used = PO_SYM_FLAG_BYTES;
    cs->peer_sym = sym;
    cs->peer_is_ipc = false;
    if (const char* t = getenv("PO_PEER_TIMEOUT_MS")) cs->timeout_ns = (long long)atoll(t) * 1000000LL;
    ctxs[r]->comm = cs;
    cudaSetDevice(ctxs[r]->device);
    if (cudaMalloc((void**)&cs->d_peer_sym, (size_t)n_ctx * sizeof(char*)) != cudaSuccess ||
        cudaMemcpy(cs->d_peer_sym, sym.data(), (size_t)n_ctx * sizeof(char*), cudaMemcpyHostToDevice) != cudaSuccess) {
      for (int q = 0; q <= r; ++q) sym[(size_t)q] = ctxs[q]->comm->sym;
      cleanup();
      return po_fail(PO_ERR_CUDA, "po_comm_init_all: device table allocation failed");
    }
  }
  return PO_OK;
#endif
}

PO_API int32_t po_comm_init_external(po_ctx* ctx, int32_t n_ranks, int32_t rank, po_exchange_fn exchange,
                                     po_collective_fn collective, void* user) {
  if (!ctx || !exchange) return po_fail(PO_ERR_INVALID, "null argument");
  if (ctx->comm) return po_fail(PO_ERR_INVALID, "context already has a communicator");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return po_fail(PO_ERR_INVALID, "bad rank %d of %d", rank, n_ranks);
  po_comm_state* cs = new po_comm_state();
  cs->n_ranks = n_ranks; cs->rank = rank; cs->ext_exchange = exchange; cs->ext_collective = collective; cs->ext_user = user;
  ctx->comm = cs;
  return PO_OK;
}

PO_API int32_t po_comm_destroy(po_ctx* ctx) {
  if (!ctx || !ctx->comm) return PO_OK;
#ifndef PO_EMUL
  po_comm_teardown_peer(ctx->comm);
  if (ctx->comm->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->nccl_comm);
#endif
  delete ctx->comm;
  ctx->comm = nullptr;
  return PO_OK;
}

PO_API int32_t po_comm_rank(const po_ctx* ctx, int32_t* rank, int32_t* n_ranks) {
  if (!ctx) return po_fail(PO_ERR_INVALID, "null ctx");
  if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
  if (n_ranks) *n_ranks = ctx->comm ? ctx->comm->n_ranks : 1;
  return PO_OK;
}

static int po_collective(po_ctx* ctx, int op, const void* send, void* recv, int64_t count, int32_t dtype, int32_t root, void* stream) {
  if (!ctx) return po_fail(PO_ERR_INVALID, "null ctx");
  po_comm_state* cm = ctx->comm;
  const size_t esz = op == 0 ? 1 : po_dtype_size(dtype);
  if (!cm || cm->n_ranks == 1) {  // single rank: bcast/reduce/allreduce are copies
    if (send != recv && count) PO_CUDA_CHECK(cudaMemcpyAsync(recv, send, (size_t)count * esz, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PO_OK;
  }
  if (cm->ext_collective || cm->ext_exchange) {
    if (!cm->ext_collective) return po_fail(PO_ERR_UNSUPPORTED, "external communicator has no collective callback");
    int rc = cm->ext_collective(cm->ext_user, op, send, recv, count, dtype, root, stream);
    if (rc) return po_fail(PO_ERR_NCCL, "external collective callback failed with %d", rc);
    return PO_OK;
  }
#ifndef PO_EMUL
  if (!cm->nccl_comm) return po_fail(PO_ERR_UNSUPPORTED, "this communicator (po_comm_init_all) carries ParRepartition only; parameter collectives need po_comm_init_rank");
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (op == 0) {
    PO_NCCL_CHECK(g_nccl.Broadcast(send, recv, (size_t)count, PO_NCCL_CHAR, root, cm->nccl_comm, st));
    return PO_OK;
  }
  const int nt = po_dtype_is_double(dtype) ? PO_NCCL_F64 : PO_NCCL_F32;
  const size_t n = (size_t)count * (po_dtype_is_complex(dtype) ? 2 : 1);
  if (op == 1) PO_NCCL_CHECK(g_nccl.Reduce(send, recv, n, nt, PO_NCCL_SUM, root, cm->nccl_comm, st));
  else PO_NCCL_CHECK(g_nccl.AllReduce(send, recv, n, nt, PO_NCCL_SUM, cm->nccl_comm, st));
#endif
  return PO_OK;
}

PO_API int32_t po_bcast(po_ctx* ctx, void* buf, size_t bytes, int32_t root, void* stream) {
  return po_collective(ctx, 0, buf, buf, (int64_t)bytes, PO_F32, root, stream);
}
PO_API int32_t po_reduce_sum(po_ctx* ctx, const void* send, void* recv, int64_t count, int32_t dtype, int32_t root, void* stream) {
  if (dtype < 0 || dtype > 3) return po_fail(PO_ERR_DTYPE, "bad dtype");
  return po_collective(ctx, 1, send, recv, count, dtype, root, stream);
}
PO_API int32_t po_allreduce_sum(po_ctx* ctx, const void* send, void* recv, int64_t count, int32_t dtype, void* stream) {
  if (dtype < 0 || dtype > 3) return po_fail(PO_ERR_DTYPE, "bad dtype");
  return po_collective(ctx, 2, send, recv, count, dtype, 0, stream);
}
