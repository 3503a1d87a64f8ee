// Plan compiler: flat operator tree -> short list of fused mode-product kernels.
//
// The reference interprets the tree node by node (src/ParCompose.jl:44-56,
// src/ParKron.jl:190-217) with a materialised permutedims around every Kronecker factor.
// Here the (parameterised) tree is compiled once per (tree, batch):
//   1. lower   : every leaf becomes a step acting on a strided mode (I, n -> m, O) of the
//                resident column-major tensor -- Kron factor o of N acts on array dim N-o+1
//                (src/ParKron.jl:193-195), factors are visited in the reference's
//                application order, Compose children right-to-left;
//   2. schedule: restriction gathers commute with steps on disjoint modes, so shrinking
//                gathers are moved as early as possible and zero-fill scatters as late as
//                possible (this is the mixed-product rewrite Kron(R..)*Kron(F..) ->
//                Kron(R*F ..) of src/ParCompose.jl:144-183 done deterministically);
//   3. fuse    : DFT followed by a gather on the same mode becomes ONE truncated DFT
//                (dense DFT-matrix rows when few modes are kept, FFT with a gathering store
//                otherwise); scatter followed by an inverse DFT becomes ONE zero-padded
//                inverse;
//   4. select  : per-step kernel + device tables (DFT matrices / twiddles in FP64 on the
//                host, rounded once), ping-pong workspace assignment.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "po_kernels.cuh"
#include "po_fused.cuh"
#include "po_fused2.cuh"
#include "po_fused3.cuh"
#include "po_fused4.cuh"
#include "po_zmz.cuh"
#include "po_tensor.cuh"
#include "po_c2.cuh"

// ---- errors ---------------------------------------------------------------------------
static thread_local std::string g_po_error;

static int po_fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_po_error = buf;
  return status;
}
#define PO_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return po_fail(PO_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

struct po_comm_state;  // po_comm.cuh

struct po_ctx {
  int device = 0;
  int num_sms = 0;  // persistent kernels launch one CTA per SM
  // Error word raised by a kernel whose bounded wait expired (mbarrier of the row ring, peer flag of the NVLink
  // repartition).  It lives in page-locked HOST memory mapped into the device (d_err is the device alias), so the
  // host can poll it without synchronising: every apply / pullback entry point checks it and fails loudly.
  volatile int* h_err = nullptr;
  int* d_err = nullptr;
  po_comm_state* comm = nullptr;
  void* scratch = nullptr;  // 1 MB of device scratch for the small transposed weights of the block-epilogue pullback
  std::mutex mu;
  std::map<std::string, po_plan*> plan_cache;  // single-operator conveniences
};

// Sticky: once a device wait timed out, results on this context cannot be trusted (the kernel poisoned what it could
// not compute); every entry point fails until po_ctx_device_status has been read (which clears the word).
static int po_ctx_check(const po_ctx* ctx) {
  if (ctx && ctx->h_err && *ctx->h_err)
    return po_fail(PO_ERR_CUDA, "a device-side wait timed out on this context (status 0x%x: 0x1-0xff bulk-copy barrier of a streaming kernel, "
                   "0x100 peer never acknowledged a ParRepartition staging region, 0x200 peer's ParRepartition data never arrived); "
                   "results since then are invalid -- read po_ctx_device_status to clear", *ctx->h_err);
  return PO_OK;
}

// ---- steps ----------------------------------------------------------------------------
enum po_step_kind {
  ST_DFT = 0,      // (possibly truncated / zero-padded) DFT; realised as ST_FFT or ST_DENSE
  ST_GATHER = 1,   // restriction / its zero-fill adjoint
  ST_MATRIX = 2,   // parameter matrix (dense kernel on the caller's theta)
  ST_DIAG = 3,
  ST_TENSOR = 4,
  ST_BIAS = 5,
  ST_GELU = 6,
  ST_REPART = 7,
  ST_FUSED_TX = 8,  // fused (t, x) truncated transform pair of the FNO spectral stack (po_fused.cuh)
  ST_MIXDIAG = 9,   // ParDiagonal ⊗ ParMatrix frequency weighting in one pass (po_fused.cuh)
  ST_ZMZ = 10       // pruned z-DFT -> mix+diag -> zero-padded inverse z-DFT in one kernel (po_zmz.cuh)
};
enum po_impl_kind { IM_NONE = 0, IM_FFT, IM_DENSE_TABLE, IM_DENSE_PARAM, IM_GATHER, IM_DIAG, IM_TENSOR, IM_BIAS, IM_GELU, IM_REPART, IM_FUSED_TX, IM_C16, IM_MIXDIAG, IM_ZMZ, IM_R16, IM_CFULL };

struct po_repart_info;  // po_comm.cuh

struct po_step {
  int kind = 0;
  i64 I = 1, n = 0, m = 0, O = 1;  // input viewed (I, n, O) -> output (I, m, O)
  int in_dt = 0, out_dt = 0;
  int param_slot = -1;
  int node_idx = -1;
  int adjoint = 0;
  int param_slot2 = -1;  // ST_MIXDIAG: the diagonal's slot / adjoint flag / length
  int adjoint2 = 0;
  i64 nd = 0;
  // DFT family
  i64 dft_n = 0;              // transform length
  int dft_inverse = 0;        // 1: 1/n-normalised backward transform (src/ParDFT.jl:30-31)
  int dft_real = 0;           // 1: r2c (forward) / c2r (inverse)
  std::vector<int> in_map;    // spectrum bin -> source row (-1: zero), empty = identity   (zero-fill scatter fused)
  std::vector<int> out_map;   // output row -> spectrum bin, empty = identity               (gather fused)
  // gather
  std::vector<int> map;       // output row -> source row (-1: zero)
  std::vector<int> map_T;     // the other direction of the same ParRestriction (pullback, src/ParRestriction.jl:57-71)
  // matrix
  i64 prm_rows = 0, prm_cols = 0;
  // tensor
  int ic = 0, oc = 0, w_oi = 0;
  i64 M = 0;
  // repartition
  po_repart_info* repart = nullptr;
  // fused (t, x) pair
  po_fused_tw* ftw = nullptr;
  po_fused2_tw* f2tw = nullptr;  // second-generation pipelined kernels (po_fused2.cuh)
  po_c16_tw* ctw = nullptr;
  po_c16_tw* ctw2 = nullptr;  // ST_ZMZ: inverse-side table (ctw is the forward side)
  i64 z_Iout = 0, z_Mi = 0;   // ST_ZMZ: output inner extent (m_ch * Mi), modes per slab
  int f_nt = 0, f_nx = 0, f_I0 = 0;
  i64 f_mid = 0, f_B = 0;
  // ---- realised
  int impl = IM_NONE;
  void* d_table = nullptr;    // DFT matrix (dense) or twiddles (fft)
  int table_dt = 0;
  i64 a_rs = 0, a_cs = 0;
  int conj_a = 0;
  int* d_in_map = nullptr;
  int* d_out_map = nullptr;
  int* d_map = nullptr;
  int fft_mode = 0;
  std::string note;
};

struct po_plan {
  po_ctx* ctx = nullptr;
  int64_t batch = 1;
  uint32_t flags = 0;
  int64_t domain = 0, range = 0;
  int ddt = 0, rdt = 0;
  int n_params = 0;
  std::vector<po_step> steps;
  std::vector<size_t> step_out_bytes;
  size_t ws_half = 0;  // bytes of one ping-pong half
  int64_t launches = 0;
  int64_t param_bytes = 0;
  void* cur_tape_slot = nullptr;  // set around a step that fills its own tape slot (ST_ZMZ)
  std::vector<void*> owned;  // device allocations
  std::vector<std::string> param_ids;  // po_plan_create_json: "id" of the parametric leaf bound to each parameter slot
  // buffers for po_plan_apply_host
  void* h_x = nullptr; void* h_y = nullptr; void* h_ws = nullptr;
  // per-step CUDA-event timing (po_plan_set_profiling)
  bool profiling = false;
  std::vector<std::vector<cudaEvent_t> > pending_events;  // one (nsteps+1) set per profiled apply
  std::vector<double> step_ms;
  std::vector<int64_t> step_calls;
  std::vector<std::vector<cudaEvent_t> > pending_pb_events;  // pullback launches (reverse step order)
  std::vector<double> pb_step_ms;
  std::vector<int64_t> pb_step_calls;
  ~po_plan();
};

static void po_step_free(po_step& s);

po_plan::~po_plan() {
  for (void* p : owned) cudaFree(p);
  if (h_x) cudaFree(h_x);
  if (h_y) cudaFree(h_y);
  if (h_ws) cudaFree(h_ws);
  for (auto& s : steps) po_step_free(s);
#ifndef PO_EMUL
  for (auto& v : pending_events) for (cudaEvent_t e : v) cudaEventDestroy(e);
  for (auto& v : pending_pb_events) for (cudaEvent_t e : v) cudaEventDestroy(e);
#endif
}

// ---- lowering -------------------------------------------------------------------------
struct po_lowerer {
  const po_node* nodes; int n_nodes;
  const int32_t* child; int n_child;
  const int64_t* aux; int n_aux;
  std::vector<po_step> steps;
  int cur_dt = 0;
  int max_param = -1;
  int rank = -1;  // communicator rank (repartition)

  int check_node(int idx) const {
    if (idx < 0 || idx >= n_nodes) return po_fail(PO_ERR_INVALID, "node index %d out of range", idx);
    const po_node& nd = nodes[idx];
    if (nd.aux_count < 0 || nd.aux_begin < 0 || nd.aux_begin + nd.aux_count > n_aux)
      return po_fail(PO_ERR_INVALID, "node %d: aux slice out of range", idx);
    if (nd.child_count < 0 || nd.child_begin < 0 || nd.child_begin + nd.child_count > n_child)
      return po_fail(PO_ERR_INVALID, "node %d: child slice out of range", idx);
    if (nd.ddt < 0 || nd.ddt > 3 || nd.rdt < 0 || nd.rdt > 3) return po_fail(PO_ERR_DTYPE, "node %d: bad dtype", idx);
    if (nd.domain < 0 || nd.range < 0) return po_fail(PO_ERR_SHAPE, "node %d: negative size", idx);
    return PO_OK;
  }

  int enter_leaf(int idx, const po_node& nd, const char* name) {
    if (cur_dt != nd.ddt)
      return po_fail(PO_ERR_DTYPE, "%s (node %d): operand is %s but the operator's domain type is %s", name, idx,
                     po_dtype_name(cur_dt), po_dtype_name(nd.ddt));
    return PO_OK;
  }

  int lower(int idx, i64 I, i64 O, int depth);
};

static int po_build_restriction_maps(const int64_t* a, int cnt, int adjoint, i64& n, i64& total, std::vector<int>& map,
                                     std::vector<int>* map_other = nullptr) {
  if (map_other) {
    int rc2 = po_build_restriction_maps(a, cnt, !adjoint, n, total, *map_other, nullptr);
    if (rc2) return rc2;
  }
  if (cnt < 2) return po_fail(PO_ERR_INVALID, "ParRestriction: aux too short");
  n = a[0];
  const i64 nr = a[1];
  if (nr < 0 || cnt != 2 + 2 * nr) return po_fail(PO_ERR_INVALID, "ParRestriction: aux length does not match range count");
  total = 0;
  for (i64 r = 0; r < nr; ++r) {
    const i64 s = a[2 + 2 * r], e = a[3 + 2 * r];
    if (e >= s) {
      if (s < 1 || e > n) return po_fail(PO_ERR_SHAPE, "ParRestriction: range %lld:%lld outside 1:%lld", (long long)s, (long long)e, (long long)n);
      total += e - s + 1;
    }
  }
  if (n > 0x7fffffffLL || total > 0x7fffffffLL) return po_fail(PO_ERR_UNSUPPORTED, "ParRestriction: extent exceeds int32");
  if (!adjoint) {  // y[off + t] = x[s + t]
    map.assign((size_t)total, -1);
    i64 off = 0;
    for (i64 r = 0; r < nr; ++r) {
      const i64 s = a[2 + 2 * r], e = a[3 + 2 * r];
      for (i64 t = s; t <= e; ++t) map[(size_t)(off++)] = (int)(t - 1);
    }
  } else {  // x = 0; x[s + t] = y[off + t], later ranges overwrite (src/ParRestriction.jl:31-37)
    map.assign((size_t)n, -1);
    i64 off = 0;
    for (i64 r = 0; r < nr; ++r) {
      const i64 s = a[2 + 2 * r], e = a[3 + 2 * r];
      for (i64 t = s; t <= e; ++t) map[(size_t)(t - 1)] = (int)(off++);
    }
  }
  return PO_OK;
}

int po_lowerer::lower(int idx, i64 I, i64 O, int depth) {
  int rc = check_node(idx);
  if (rc) return rc;
  if (depth > 64) return po_fail(PO_ERR_INVALID, "operator tree too deep (cycle?)");
  const po_node& nd = nodes[idx];
  const int64_t* a = aux + nd.aux_begin;
  if (nd.param_slot > max_param) max_param = nd.param_slot;
  switch (nd.kind) {
    case PO_NODE_IDENTITY: {
      if ((rc = enter_leaf(idx, nd, "ParIdentity"))) return rc;
      if (nd.domain != nd.range || nd.ddt != nd.rdt) return po_fail(PO_ERR_SHAPE, "ParIdentity: Domain != Range");
      return PO_OK;
    }
    case PO_NODE_DFT: {
      if ((rc = enter_leaf(idx, nd, "ParDFT"))) return rc;
      if (nd.aux_count != 1) return po_fail(PO_ERR_INVALID, "ParDFT: aux must be [n]");
      const i64 n = a[0];
      po_step s;
      s.kind = ST_DFT; s.I = I; s.O = O; s.in_dt = nd.ddt; s.out_dt = nd.rdt; s.dft_n = n;
      s.dft_inverse = nd.adjoint;
      const int real_side = nd.adjoint ? nd.rdt : nd.ddt;
      const int cplx_side = nd.adjoint ? nd.ddt : nd.rdt;
      s.dft_real = !po_dtype_is_complex(real_side);
      if (!po_dtype_is_complex(cplx_side) || po_dtype_is_double(real_side) != po_dtype_is_double(cplx_side))
        return po_fail(PO_ERR_DTYPE, "ParDFT: invalid type pair %s -> %s", po_dtype_name(nd.ddt), po_dtype_name(nd.rdt));
      const i64 nspec = s.dft_real ? n / 2 + 1 : n;
      s.n = nd.adjoint ? nspec : n;
      s.m = nd.adjoint ? n : nspec;
      if (nd.domain != s.n || nd.range != s.m)
        return po_fail(PO_ERR_SHAPE, "ParDFT(n=%lld): Domain/Range %lld/%lld do not match %lld/%lld", (long long)n,
                       (long long)nd.domain, (long long)nd.range, (long long)s.n, (long long)s.m);
      if (n > 0x3fffffffLL) return po_fail(PO_ERR_UNSUPPORTED, "ParDFT: n too large");
      cur_dt = nd.rdt;
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_RESTRICTION: {
      if ((rc = enter_leaf(idx, nd, "ParRestriction"))) return rc;
      if (nd.ddt != nd.rdt) return po_fail(PO_ERR_DTYPE, "ParRestriction: DDT != RDT");
      po_step s;
      s.kind = ST_GATHER; s.I = I; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.adjoint = nd.adjoint;
      i64 n, total;
      if ((rc = po_build_restriction_maps(a, nd.aux_count, nd.adjoint, n, total, s.map, &s.map_T))) return rc;
      s.n = nd.adjoint ? total : n;
      s.m = nd.adjoint ? n : total;
      if (nd.domain != s.n || nd.range != s.m)
        return po_fail(PO_ERR_SHAPE, "ParRestriction: Domain/Range %lld/%lld do not match %lld/%lld", (long long)nd.domain,
                       (long long)nd.range, (long long)s.n, (long long)s.m);
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_MATRIX: {
      if ((rc = enter_leaf(idx, nd, "ParMatrix"))) return rc;
      if (nd.aux_count != 2) return po_fail(PO_ERR_INVALID, "ParMatrix: aux must be [m, n]");
      if (nd.ddt != nd.rdt) return po_fail(PO_ERR_DTYPE, "ParMatrix: DDT != RDT");
      if (nd.param_slot < 0) return po_fail(PO_ERR_INVALID, "ParMatrix: not parameterised (no param slot)");
      po_step s;
      s.kind = ST_MATRIX; s.I = I; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.adjoint = nd.adjoint;
      s.param_slot = nd.param_slot; s.prm_rows = a[0]; s.prm_cols = a[1];
      s.n = nd.adjoint ? a[0] : a[1];
      s.m = nd.adjoint ? a[1] : a[0];
      if (nd.domain != s.n || nd.range != s.m) return po_fail(PO_ERR_SHAPE, "ParMatrix: Domain/Range mismatch");
      if (s.n > 0x7fffffffLL || s.m > 0x7fffffffLL) return po_fail(PO_ERR_UNSUPPORTED, "ParMatrix: extent exceeds int32");
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_DIAGONAL: {
      if ((rc = enter_leaf(idx, nd, "ParDiagonal"))) return rc;
      if (nd.aux_count != 1) return po_fail(PO_ERR_INVALID, "ParDiagonal: aux must be [n]");
      if (nd.ddt != nd.rdt) return po_fail(PO_ERR_DTYPE, "ParDiagonal: DDT != RDT");
      if (nd.param_slot < 0) return po_fail(PO_ERR_INVALID, "ParDiagonal: not parameterised (no param slot)");
      if (nd.domain != a[0] || nd.range != a[0]) return po_fail(PO_ERR_SHAPE, "ParDiagonal: Domain/Range mismatch");
      po_step s;
      s.kind = ST_DIAG; s.I = I; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.adjoint = nd.adjoint;
      s.param_slot = nd.param_slot; s.n = s.m = a[0];
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_TENSOR: {
      if ((rc = enter_leaf(idx, nd, "ParTensor"))) return rc;
      if (nd.adjoint) return po_fail(PO_ERR_UNSUPPORTED, "ParTensor: the reference defines no adjoint apply (src/ParTensor.jl)");
      if (nd.ddt != nd.rdt) return po_fail(PO_ERR_DTYPE, "ParTensor: DDT != RDT");
      if (nd.param_slot < 0) return po_fail(PO_ERR_INVALID, "ParTensor: not parameterised (no param slot)");
      if (I != 1) return po_fail(PO_ERR_UNSUPPORTED, "ParTensor must act on the leading (fastest) dims");
      if (nd.aux_count < 1 || nd.aux_count != 1 + a[0]) return po_fail(PO_ERR_INVALID, "ParTensor: aux must be [rank, weight_shape...]");
      const int rank_w = (int)a[0];
      if (rank_w < 4 || rank_w > 6) return po_fail(PO_ERR_UNSUPPORTED, "ParTensor: only ranks 4, 5, 6 are defined (src/ParTensor.jl:35-113)");
      po_step s;
      s.kind = ST_TENSOR; s.I = 1; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.param_slot = nd.param_slot;
      i64 M = 1;
      for (int d = 3; d <= rank_w; ++d) M *= a[d];
      if (rank_w == 4) { s.ic = (int)a[1]; s.oc = (int)a[2]; s.w_oi = 0; }
      else { s.oc = (int)a[1]; s.ic = (int)a[2]; s.w_oi = 1; }
      s.M = M; s.n = (i64)s.ic * M; s.m = (i64)s.oc * M;
      if (nd.domain != s.n || nd.range != s.m) return po_fail(PO_ERR_SHAPE, "ParTensor: Domain/Range mismatch");
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_BIAS: {
      if ((rc = enter_leaf(idx, nd, "bias"))) return rc;
      if (nd.aux_count != 1 || nd.param_slot < 0) return po_fail(PO_ERR_INVALID, "bias: aux must be [m] with a param slot");
      if (nd.domain != a[0] || nd.range != a[0] || nd.ddt != nd.rdt) return po_fail(PO_ERR_SHAPE, "bias: Domain/Range mismatch");
      po_step s;
      s.kind = ST_BIAS; s.I = I; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.param_slot = nd.param_slot; s.n = s.m = a[0];
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_GELU: {
      if ((rc = enter_leaf(idx, nd, "gelu"))) return rc;
      if (po_dtype_is_complex(nd.ddt) || nd.ddt != nd.rdt || nd.domain != nd.range)
        return po_fail(PO_ERR_DTYPE, "gelu: real T -> T only");
      po_step s;
      s.kind = ST_GELU; s.I = I; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.n = s.m = nd.domain;
      steps.push_back(s);
      return PO_OK;
    }
    case PO_NODE_COMPOSE: {
      if (nd.child_count < 1) return po_fail(PO_ERR_INVALID, "ParCompose: no children");
      // src/ParCompose.jl:16-19 asserts
      for (int c = 0; c + 1 < nd.child_count; ++c) {
        const int l = child[nd.child_begin + c], r = child[nd.child_begin + c + 1];
        if ((rc = check_node(l)) || (rc = check_node(r))) return rc;
        if (nodes[l].domain != nodes[r].range)
          return po_fail(PO_ERR_SHAPE, "ParCompose: Domain(ops[%d])=%lld != Range(ops[%d])=%lld", c + 1, (long long)nodes[l].domain, c + 2, (long long)nodes[r].range);
        if (nodes[l].ddt != nodes[r].rdt)
          return po_fail(PO_ERR_DTYPE, "ParCompose: DDT(ops[%d])=%s != RDT(ops[%d])=%s", c + 1, po_dtype_name(nodes[l].ddt), c + 2, po_dtype_name(nodes[r].rdt));
      }
      for (int c = nd.child_count - 1; c >= 0; --c)  // src/ParCompose.jl:51-54: ops[N] first
        if ((rc = lower(child[nd.child_begin + c], I, O, depth + 1))) return rc;
      return PO_OK;
    }
    case PO_NODE_KRON: {
      const int N = nd.child_count;
      if (N < 1 || nd.aux_count != N) return po_fail(PO_ERR_INVALID, "ParKron: order must list every factor");
      std::vector<i64> cur((size_t)N);  // cur[d] = current extent of array dim d (0-based, fastest first)
      i64 dom = 1, ran = 1;
      for (int o = 1; o <= N; ++o) {
        const int ci = child[nd.child_begin + o - 1];
        if ((rc = check_node(ci))) return rc;
        cur[(size_t)(N - o)] = nodes[ci].domain;
        dom *= nodes[ci].domain;
        ran *= nodes[ci].range;
      }
      if (dom != nd.domain || ran != nd.range) return po_fail(PO_ERR_SHAPE, "ParKron: Domain/Range is not the product of the factors'");
      std::vector<char> seen((size_t)N, 0);
      for (int t = 0; t < N; ++t) {
        const i64 o = a[t];
        if (o < 1 || o > N || seen[(size_t)(o - 1)]) return po_fail(PO_ERR_INVALID, "ParKron: order is not a permutation of 1..N");
        seen[(size_t)(o - 1)] = 1;
        const int d = (int)(N - o);
        i64 inner = I, outer = O;
        for (int e = 0; e < d; ++e) inner *= cur[(size_t)e];
        for (int e = d + 1; e < N; ++e) outer *= cur[(size_t)e];
        const int ci = child[nd.child_begin + o - 1];
        if ((rc = lower(ci, inner, outer, depth + 1))) return rc;
        cur[(size_t)d] = nodes[ci].range;
      }
      return PO_OK;
    }
    case PO_NODE_REPARTITION: {
      if ((rc = enter_leaf(idx, nd, "ParRepartition"))) return rc;
      if (I != 1) return po_fail(PO_ERR_UNSUPPORTED, "ParRepartition must be applied to the whole local tensor");
      po_step s;
      s.kind = ST_REPART; s.I = 1; s.O = O; s.in_dt = s.out_dt = nd.ddt; s.n = nd.domain; s.m = nd.range;
      s.node_idx = idx;  // aux slice of this node is parsed in po_comm.cuh
      steps.push_back(s);
      return PO_OK;
    }
    default:
      return po_fail(PO_ERR_INVALID, "node %d: unknown kind %d", idx, nd.kind);
  }
}

// ---- scheduling: commute steps acting on disjoint modes ---------------------------------
static bool po_is_mode_op(const po_step& s) {
  return s.kind == ST_DFT || s.kind == ST_GATHER || s.kind == ST_MATRIX || s.kind == ST_DIAG;
}

// A is applied first, then B.  On success both are rewritten so that B' then A' is equivalent.
static bool po_try_swap(po_step& A, po_step& B) {
  if (!po_is_mode_op(A) || !po_is_mode_op(B)) return false;
  // only steps that keep the element type (and agree on it) are exchanged; the r2c / c2r
  // steps are always first / last in a valid chain, so nothing ever needs to cross them
  if (A.in_dt != A.out_dt || B.in_dt != B.out_dt || A.in_dt != B.in_dt) return false;
  if (A.n <= 0 || A.m <= 0 || B.n <= 0 || B.m <= 0 || A.I <= 0 || B.I <= 0) return false;
  const i64 a_span = A.I * A.m;  // extent of A's mode after A
  const i64 b_span = B.I * B.n;  // extent of B's mode before B
  if (B.I >= a_span && B.I % a_span == 0) {
    // B's mode is slower than A's: tensor is (A.I, A.n, c, B.n, O')
    if (A.O % B.n != 0) return false;
    const i64 c = B.I / a_span;
    po_step B2 = B, A2 = A;
    B2.I = c * A.I * A.n;
    A2.O = A.O / B.n * B.m;
    A = B2; B = A2;
    return true;
  }
  if (A.I >= b_span && A.I % b_span == 0) {
    // A's mode is slower than B's: tensor is (B.I, B.n, c, A.n, O')
    if (B.O % A.m != 0) return false;
    po_step B2 = B, A2 = A;
    B2.O = B.O / A.m * A.n;
    A2.I = A.I / B.n * B.m;
    A = B2; B = A2;
    return true;
  }
  return false;
}

static void po_schedule(std::vector<po_step>& st) {
  // shrinking gathers bubble towards the front
  for (size_t p = 1; p < st.size(); ++p) {
    if (st[p].kind != ST_GATHER || !(st[p].m < st[p].n)) continue;
    size_t q = p;
    while (q > 0) {
      const po_step& prev = st[q - 1];
      // stop behind the DFT of our own mode (fusion candidate) or any non-commuting step
      if (prev.I == st[q].I && prev.O == st[q].O && prev.m == st[q].n) break;
      if (!po_try_swap(st[q - 1], st[q])) break;
      --q;
    }
  }
  // expanding (zero-fill) gathers sink towards the back
  for (size_t pp = st.size(); pp-- > 0;) {
    if (st[pp].kind != ST_GATHER || !(st[pp].m > st[pp].n)) continue;
    size_t q = pp;
    while (q + 1 < st.size()) {
      const po_step& next = st[q + 1];
      if (next.I == st[q].I && next.O == st[q].O && next.n == st[q].m) break;
      if (!po_try_swap(st[q], st[q + 1])) break;
      ++q;
    }
  }
}

static void po_fuse(std::vector<po_step>& st) {
  std::vector<po_step> out;
  for (size_t p = 0; p < st.size(); ++p) {
    po_step& s = st[p];
    // forward DFT followed by a gather of its output mode -> truncated DFT
    if (s.kind == ST_GATHER && !out.empty()) {
      po_step& prev = out.back();
      if (prev.kind == ST_DFT && !prev.dft_inverse && prev.I == s.I && prev.O == s.O && prev.m == s.n && prev.out_dt == s.in_dt) {
        bool ok = true;
        for (int v : s.map) if (v < 0) ok = false;
        if (ok) {
          std::vector<int> bins(s.map.size());
          for (size_t k = 0; k < s.map.size(); ++k) bins[k] = prev.out_map.empty() ? s.map[k] : prev.out_map[(size_t)s.map[k]];
          prev.out_map = bins;
          prev.m = s.m;
          continue;
        }
      }
    }
    // gather (zero-fill scatter) followed by an inverse DFT on the same mode -> zero-padded inverse
    if (s.kind == ST_DFT && s.dft_inverse && !out.empty()) {
      po_step& prev = out.back();
      if (prev.kind == ST_GATHER && prev.I == s.I && prev.O == s.O && prev.m == s.n && prev.out_dt == s.in_dt && s.in_map.empty()) {
        po_step f = s;
        f.in_map = prev.map;  // spectrum bin -> source row
        f.n = prev.n;
        out.back() = f;
        continue;
      }
    }
    out.push_back(s);
  }
  st.swap(out);
}

// ---- shape-specialised fast paths -------------------------------------------------------
static inline void po_cis(i64 num, i64 den, double sign, double& c, double& s);
static bool po_is_low8(const std::vector<int>& m) {
  if (m.size() != 8) return false;
  for (int k = 0; k < 8; ++k) if (m[(size_t)k] != k) return false;
  return true;
}
static bool po_is_low16(const std::vector<int>& m) {
  if (m.size() != 16) return false;
  for (int k = 0; k < 16; ++k) if (m[(size_t)k] != k) return false;
  return true;
}
static bool po_is_low16_scatter(const std::vector<int>& m) {  // bin -> src: {0..15 -> 0..15}, everything else empty
  if (m.size() < 17) return false;
  for (size_t b = 0; b < m.size(); ++b) if (m[b] != (b < 16 ? (int)b : -1)) return false;
  return true;
}
static bool po_is_lowhigh16(const std::vector<int>& m, i64 n) {
  if (m.size() != 16) return false;
  for (int r = 0; r < 16; ++r) if (m[(size_t)r] != (r < 8 ? r : (int)n - 16 + r)) return false;
  return true;
}
// in_map of a zero-padded inverse: bin -> src; accept exactly {0..7 -> 0..7} (+ {n-8..n-1 -> 8..15})
static bool po_is_low8_scatter(const std::vector<int>& m) {
  if (m.size() < 9) return false;
  for (size_t b = 0; b < m.size(); ++b) if (m[b] != (b < 8 ? (int)b : -1)) return false;
  return true;
}
static bool po_is_lowhigh16_scatter(const std::vector<int>& m, i64 n) {
  if ((i64)m.size() != n || n < 32) return false;
  for (i64 b = 0; b < n; ++b) {
    const int want = b < 8 ? (int)b : (b >= n - 8 ? (int)(b - (n - 16)) : -1);
    if (m[(size_t)b] != want) return false;
  }
  return true;
}

static void po_fill_fused_tw(po_fused_tw& tw, int nt, int nx, bool inverse) {
  const int RT = nt / 8, RX = nx / 16;
  const double sgn = inverse ? +1.0 : -1.0;
  for (int j2 = 0; j2 < 8; ++j2)
    for (int k = 0; k < 8; ++k) {
      double c = 0, s = 0;
      if (j2 < RT) po_cis((i64)k * j2, nt, sgn, c, s);
      double w = 1.0;
      if (inverse) w = ((k == 0) ? 1.0 : 2.0) / ((double)nt * (double)nx);
      const float wx = (float)(w * c), wy = (float)(w * s);
      float2* e = tw.tw_t + (j2 * 8 + k) * 4;
      e[0] = make_float2(wx, wx);
      e[1] = make_float2(wy, wy);
      e[2] = make_float2(-wx, -wx);
      e[3] = make_float2(-wy, -wy);
    }
  for (int j2 = 0; j2 < 8; ++j2)
    for (int r = 0; r < 16; ++r) {
      double c = 0, s = 0;
      const i64 bin = r < 8 ? r : (i64)nx - 16 + r;
      if (j2 < RX) po_cis(bin * j2, nx, sgn, c, s);
      tw.tw_x[j2 * 16 + r] = make_float2((float)c, (float)s);
    }
  for (int k = 0; k < 8; ++k) tw.kscale[k] = 1.f;
}

// per-t-mode weights of the backward pair: 1/(nt nx) and the Hermitian doubling of the c2r transform
static void po_herm_kscale(int nt, int nx, double* ks) {
  for (int k = 0; k < 8; ++k) ks[k] = ((k == 0) ? 1.0 : 2.0) / ((double)nt * (double)nx);
}
static int po_fused_gen_env() {
  static const int env = getenv("PO_FUSED_GEN") ? atoi(getenv("PO_FUSED_GEN")) : 3;
  return env;
}
static bool po_use_gen2(const po_plan* pl) { return po_fused_gen_env() >= 2 && !(pl->flags & PO_PLAN_NO_PIPELINE); }
// single-tile second-generation kernels: only for rows too wide for the double-buffered tile
static bool po_use_tile2(bool inverse, int I0, int nx) {
  static const int env = getenv("PO_TILE2") ? atoi(getenv("PO_TILE2")) : 1;
  return env && po_fused2_group(inverse, I0, nx, 216 * 1024) != I0;
}
// staged-input inverse (po_inv_xt_stage_kernel); PO_INV_STAGE=0 keeps the generation-2 inverse
static bool po_use_inv_stage() {
  static const int env = getenv("PO_INV_STAGE") ? atoi(getenv("PO_INV_STAGE")) : 1;
  return env != 0;
}
// CTA-pair kernels (po_fused4.cuh) for rows too wide for the single-CTA ring; PO_PAIR=0 keeps the single-tile generation-2 kernels
static bool po_use_pair() {
  static const int env = getenv("PO_PAIR") ? atoi(getenv("PO_PAIR")) : 1;
  return env != 0;
}
static bool po_use_gen3(const po_plan* pl) { return po_fused_gen_env() >= 3 && !(pl->flags & PO_PLAN_NO_PIPELINE) && pl->ctx->d_err; }

// ---- widening: kept-mode sets SMALLER than the ones the shape-specialised kernels compute -----------------------------------
// The fused kernels compute exactly 8 low bins of the real (time) transform and 8 low + 8 high bins of the complex ones.  A layer that
// keeps FEWER modes of a dim (any subset of those bins, e.g. modes = 4: [1:4, n-3:n]) is rewritten as
//     truncated DFT to the wide set  ->  restriction of the wide set to the kept bins            (forward)
//     zero-fill of the kept bins into the wide set  ->  zero-padded inverse DFT of the wide set   (inverse)
// and the inserted restrictions / zero-fills are moved behind all forward DFTs / in front of all inverse DFTs (they act on other modes, so
// they commute exactly), where the tensor is already <= 8 x 16^d entries per channel.  The big kernels then see the shapes they are
// written for -- their cost is the input / output stream, which does not depend on the kept count -- and the extra passes touch a few MB.
// (More than 8 / 16 kept modes per dim stay on the generic truncated-DFT kernels.)
static std::vector<int> po_wide_bins(const po_step& s) {
  std::vector<int> w;
  const i64 n = s.dft_n;
  if (s.dft_real) {
    if (n == 16 || n == 32 || n == 64) for (int k = 0; k < 8; ++k) w.push_back(k);
  } else if (n == 32 || n == 64 || n == 128) {
    for (int r = 0; r < 16; ++r) w.push_back(r < 8 ? r : (int)n - 16 + r);
  }
  return w;
}
// every kept bin of a (forward: out_map, inverse: in_map) truncated DFT lies in the wide set of its kernel family
static bool po_kept_in_wide(const po_step& s) {
  const std::vector<int> W = po_wide_bins(s);
  if (W.empty()) return false;
  auto in_w = [&](int b) { for (int w : W) if (w == b) return true; return false; };
  if (!s.dft_inverse) {
    if (s.out_map.empty() || !s.in_map.empty()) return false;
    for (int b : s.out_map) if (!in_w(b)) return false;
  } else {
    if (s.in_map.empty() || !s.out_map.empty()) return false;
    for (size_t b = 0; b < s.in_map.size(); ++b) if (s.in_map[b] >= 0 && !in_w((int)b)) return false;
  }
  return true;
}
// t = truncated real DFT, x = truncated complex DFT on the next-faster transformed dim, same direction: would the pair reach the fused
// (t, x) kernels once both kept sets are widened?
static bool po_tx_pair_widens(const po_step& t, const po_step& x) {
  if (t.kind != ST_DFT || x.kind != ST_DFT || !t.dft_real || x.dft_real || (x.dft_inverse != 0) != (t.dft_inverse != 0) || x.in_dt != PO_C64) return false;
  if (!po_kept_in_wide(t) || !po_kept_in_wide(x) || !po_fused_tx_shape_ok(t.dft_n, x.dft_n, x.I)) return false;
  return t.I % (x.I * x.dft_n) == 0;
}
// Is there a shape-specialised kernel that takes this step once it is widened?  Complex dims: the pruned-c16 kernels (>= 8 inner columns) or,
// as the x-step, the fused (t, x) kernels; the real (time) dim only has the fused kernels, so its neighbour must be the x-step they fuse with
// (forward: t then x; inverse: x then t).
static bool po_widen_pays(const std::vector<po_step>& st, size_t p) {
  const po_step& s = st[p];
  const bool fwd = !s.dft_inverse;
  if (s.dft_real) {
    if (fwd ? p + 1 >= st.size() : p == 0) return false;
    return po_tx_pair_widens(s, st[fwd ? p + 1 : p - 1]);
  }
  if (s.I >= 8) return true;
  if (fwd ? p == 0 : p + 1 >= st.size()) return false;
  return po_tx_pair_widens(st[fwd ? p - 1 : p + 1], s);
}
static void po_widen(std::vector<po_step>& st) {
  const char* env = getenv("PO_WIDEN");  // read per plan (host side): tests switch it
  if (env && !atoi(env)) return;
  std::vector<po_step> out;
  std::vector<char> ins;  // 1: inserted restriction, 2: inserted zero-fill
  for (size_t p = 0; p < st.size(); ++p) {
    const po_step& s = st[p];
    const bool f32 = (s.in_dt == PO_F32 || s.in_dt == PO_C64) && (s.out_dt == PO_F32 || s.out_dt == PO_C64) && s.kind == ST_DFT && po_widen_pays(st, p);
    if (s.kind == ST_DFT && f32 && !s.dft_inverse && s.in_map.empty() && !s.out_map.empty() && s.n == s.dft_n) {
      const std::vector<int> W = po_wide_bins(s);
      if (!W.empty() && s.out_map.size() < W.size()) {
        std::vector<int> map(s.out_map.size(), -1), map_T(W.size(), -1);
        bool ok = true;
        for (size_t k = 0; k < s.out_map.size() && ok; ++k) {
          int at = -1;
          for (size_t w = 0; w < W.size(); ++w) if (W[w] == s.out_map[k]) at = (int)w;
          if (at < 0 || map_T[(size_t)at] >= 0) ok = false;  // not in the wide set, or a bin kept twice
          else { map[k] = at; map_T[(size_t)at] = (int)k; }
        }
        if (ok) {
          po_step d = s;
          d.out_map = W; d.m = (i64)W.size();
          po_step g;
          g.kind = ST_GATHER; g.I = s.I; g.O = s.O; g.n = (i64)W.size(); g.m = s.m; g.in_dt = g.out_dt = s.out_dt;
          g.map = map; g.map_T = map_T;
          out.push_back(d); ins.push_back(0);
          out.push_back(g); ins.push_back(1);
          continue;
        }
      }
    }
    if (s.kind == ST_DFT && f32 && s.dft_inverse && s.out_map.empty() && !s.in_map.empty() && s.m == s.dft_n) {
      const std::vector<int> W = po_wide_bins(s);
      if (!W.empty() && s.n < (i64)W.size()) {
        std::vector<int> in_wide(s.in_map.size(), -1), map(W.size(), -1), map_T((size_t)s.n, -1);
        bool ok = true;
        for (size_t w = 0; w < W.size(); ++w) {
          if ((size_t)W[w] >= s.in_map.size()) { ok = false; break; }
          in_wide[(size_t)W[w]] = (int)w;
          const int src = s.in_map[(size_t)W[w]];
          map[w] = src;
          if (src >= 0) { if (src >= s.n || map_T[(size_t)src] >= 0) ok = false; else map_T[(size_t)src] = (int)w; }
        }
        for (size_t b = 0; b < s.in_map.size() && ok; ++b) if (s.in_map[b] >= 0 && in_wide[b] < 0) ok = false;  // a source bin outside the wide set
        for (i64 r = 0; r < s.n && ok; ++r) if (map_T[(size_t)r] < 0) ok = false;                                 // a source row nobody reads
        if (ok) {
          po_step g;
          g.kind = ST_GATHER; g.I = s.I; g.O = s.O; g.n = s.n; g.m = (i64)W.size(); g.in_dt = g.out_dt = s.in_dt;
          g.map = map; g.map_T = map_T;
          po_step d = s;
          d.in_map = in_wide; d.n = (i64)W.size();
          out.push_back(g); ins.push_back(2);
          out.push_back(d); ins.push_back(0);
          continue;
        }
      }
    }
    out.push_back(s); ins.push_back(0);
  }
  // inserted restrictions sink behind the forward DFTs (and the other inserted restrictions) that follow them
  for (size_t pp = out.size(); pp-- > 0;) {
    if (ins[pp] != 1) continue;
    size_t q = pp;
    while (q + 1 < out.size() && ((out[q + 1].kind == ST_DFT && !out[q + 1].dft_inverse) || ins[q + 1] == 1)) {
      if (!po_try_swap(out[q], out[q + 1])) break;
      std::swap(ins[q], ins[q + 1]);
      ++q;
    }
  }
  // inserted zero-fills rise in front of the inverse DFTs (and the other inserted zero-fills) that precede them
  for (size_t p = 0; p < out.size(); ++p) {
    if (ins[p] != 2) continue;
    size_t q = p;
    while (q > 0 && ((out[q - 1].kind == ST_DFT && out[q - 1].dft_inverse) || ins[q - 1] == 2)) {
      if (!po_try_swap(out[q - 1], out[q])) break;
      std::swap(ins[q - 1], ins[q]);
      --q;
    }
  }
  st.swap(out);
}

static void po_fastpath(std::vector<po_step>& st) {
  std::vector<po_step> out;
  for (size_t p = 0; p < st.size(); ++p) {
    if (p + 1 < st.size() && st[p].kind == ST_DFT && st[p + 1].kind == ST_DFT) {
      const po_step& a = st[p];
      const po_step& b = st[p + 1];
      // forward: a = truncated r2c on t, b = truncated c2c on x
      if (!a.dft_inverse && !b.dft_inverse && a.dft_real && !b.dft_real && a.in_dt == PO_F32 && b.in_dt == PO_C64 &&
          a.in_map.empty() && b.in_map.empty() && po_is_low8(a.out_map) && po_is_lowhigh16(b.out_map, b.dft_n) &&
          po_fused_tx_shape_ok(a.dft_n, b.dft_n, b.I) && a.I % (b.I * b.dft_n) == 0 && a.n == a.dft_n && b.n == b.dft_n) {
        const i64 mid = a.I / (b.I * b.dft_n);
        if (b.O == mid * 8 * a.O && mid * a.O < 0x7fffffffLL) {
          po_step f;
          f.kind = ST_FUSED_TX; f.in_dt = PO_F32; f.out_dt = PO_C64; f.dft_inverse = 0;
          f.I = 1; f.O = 1; f.n = a.I * a.n * a.O; f.m = b.I * b.m * b.O;
          f.f_nt = (int)a.dft_n; f.f_nx = (int)b.dft_n; f.f_I0 = (int)b.I; f.f_mid = mid; f.f_B = a.O;
          out.push_back(f);
          ++p;
          continue;
        }
      }
      // inverse: a = zero-padded c2c inverse on x, b = zero-padded c2r inverse on t
      if (a.dft_inverse && b.dft_inverse && !a.dft_real && b.dft_real && a.in_dt == PO_C64 && b.out_dt == PO_F32 &&
          a.out_map.empty() && b.out_map.empty() && po_is_lowhigh16_scatter(a.in_map, a.dft_n) && po_is_low8_scatter(b.in_map) &&
          (i64)b.in_map.size() == b.dft_n / 2 + 1 && po_fused_tx_shape_ok(b.dft_n, a.dft_n, a.I) && a.n == 16 && b.n == 8 &&
          b.I % (a.I * a.dft_n) == 0) {
        const i64 mid = b.I / (a.I * a.dft_n);
        if (a.O == mid * 8 * b.O && mid * b.O < 0x7fffffffLL) {
          po_step f;
          f.kind = ST_FUSED_TX; f.in_dt = PO_C64; f.out_dt = PO_F32; f.dft_inverse = 1;
          f.I = 1; f.O = 1; f.n = a.I * a.n * a.O; f.m = b.I * b.m * b.O;
          f.f_nt = (int)b.dft_n; f.f_nx = (int)a.dft_n; f.f_I0 = (int)a.I; f.f_mid = mid; f.f_B = b.O;
          out.push_back(f);
          ++p;
          continue;
        }
      }
    }
    // frequency weighting: ParMatrix on the channel dim (I == 1) and ParDiagonal over the modes, either order
    if (p + 1 < st.size()) {
      const po_step* mat = nullptr;
      const po_step* dg = nullptr;
      if (st[p].kind == ST_MATRIX && st[p + 1].kind == ST_DIAG) { mat = &st[p]; dg = &st[p + 1]; }
      if (st[p].kind == ST_DIAG && st[p + 1].kind == ST_MATRIX) { dg = &st[p]; mat = &st[p + 1]; }
      if (mat && dg && mat->I == 1 && mat->in_dt == dg->in_dt && mat->n <= 64 && mat->m <= 64 &&
          dg->I == (mat == &st[p] ? mat->m : mat->n) && dg->n * dg->O == mat->O) {
        po_step f = *mat;
        f.kind = ST_MIXDIAG;
        f.param_slot2 = dg->param_slot; f.adjoint2 = dg->adjoint; f.nd = dg->n;
        out.push_back(f);
        ++p;
        continue;
      }
    }
    out.push_back(st[p]);
  }
  st.swap(out);
}

// c16 eligibility of a (fused) DFT step, shared by po_realise_step and the z-mix-z matcher
static bool po_c16_fwd_ok(const po_step& s) {
  return s.kind == ST_DFT && !s.dft_real && s.in_dt == PO_C64 && (s.dft_n == 32 || s.dft_n == 64 || s.dft_n == 128) && s.I >= 8 && !s.dft_inverse &&
         s.in_map.empty() && po_is_lowhigh16(s.out_map, s.dft_n);
}
static bool po_c16_inv_ok(const po_step& s) {
  return s.kind == ST_DFT && !s.dft_real && s.in_dt == PO_C64 && (s.dft_n == 32 || s.dft_n == 64 || s.dft_n == 128) && s.I >= 8 && s.dft_inverse &&
         s.out_map.empty() && po_is_lowhigh16_scatter(s.in_map, s.dft_n) && s.n == 16;
}

// [pruned fwd DFT on the slowest transformed dim][mix+diag][zero-padded inverse DFT on the same dim] -> ST_ZMZ
static void po_fastpath_zmz(std::vector<po_step>& st) {
  static const int env = getenv("PO_ZMZ") ? atoi(getenv("PO_ZMZ")) : 1;
  if (!env) return;
  std::vector<po_step> out;
  for (size_t p = 0; p < st.size(); ++p) {
    if (p + 2 < st.size() && po_c16_fwd_ok(st[p]) && st[p + 1].kind == ST_MIXDIAG && po_c16_inv_ok(st[p + 2])) {
      const po_step& a = st[p];
      const po_step& mx = st[p + 1];
      const po_step& c = st[p + 2];
      const i64 nin = mx.n, nout = mx.m;
      if (mx.in_dt == PO_C64 && a.dft_n == c.dft_n && a.I % nin == 0 && c.I == nout * (a.I / nin) && a.O == c.O && mx.O == (a.I / nin) * 16 * a.O &&
          po_zmz_ok(a.dft_n, nin, nout, a.I / nin) && mx.nd > 0) {
        po_step f = mx;  // keeps the parameter slots, adjoint flags, prm_rows/cols, nd
        f.kind = ST_ZMZ;
        f.in_dt = f.out_dt = PO_C64;
        f.I = a.I; f.n = a.dft_n; f.m = a.dft_n; f.O = a.O;
        f.z_Iout = c.I; f.z_Mi = a.I / nin;
        f.dft_n = a.dft_n;
        f.ic = (int)nin; f.oc = (int)nout;
        out.push_back(f);
        p += 2;
        continue;
      }
    }
    out.push_back(st[p]);
  }
  st.swap(out);
}

static i64 po_step_out_elems(const po_step& s) { return (s.kind == ST_ZMZ ? s.z_Iout : s.I) * s.m * s.O; }

// ---- realisation ----------------------------------------------------------------------
static void po_step_free(po_step& s) { delete s.ftw; s.ftw = nullptr; delete s.ctw; s.ctw = nullptr; delete s.f2tw; s.f2tw = nullptr; delete s.ctw2; s.ctw2 = nullptr; }

template <class T>
static int po_upload(po_plan* pl, const std::vector<T>& host, void** dev) {
  *dev = nullptr;
  if (host.empty()) return PO_OK;
  void* p = nullptr;
  PO_CUDA_CHECK(cudaMalloc(&p, host.size() * sizeof(T)));
  pl->owned.push_back(p);
  PO_CUDA_CHECK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = p;
  return PO_OK;
}

static inline void po_cis(i64 num, i64 den, double sign, double& c, double& s) {
  // exp(sign * 2*pi*i * num/den) with exact argument reduction
  const i64 r = ((num % den) + den) % den;
  // use symmetry to keep the argument in [0, pi/4] territory for accuracy
  const double ang = 2.0 * M_PI * (double)r / (double)den;
  c = cos(ang);
  s = sign * sin(ang);
  if (4 * r == den) { c = 0.0; s = sign; }
  else if (2 * r == den) { c = -1.0; s = 0.0; }
  else if (4 * r == 3 * den) { c = 0.0; s = -sign; }
  else if (r == 0) { c = 1.0; s = 0.0; }
}

// Dense DFT matrix, column-major (m rows x ncol), complex in the step's precision.
static int po_build_dft_matrix(po_plan* pl, po_step& s) {
  const i64 n = s.dft_n;
  const bool dbl = po_dtype_is_double(s.in_dt);
  const i64 rows = s.m, cols = s.n;
  std::vector<double> re((size_t)(rows * cols), 0.0), im((size_t)(rows * cols), 0.0);
  if (!s.dft_inverse) {
    // Y[k] = sum_j exp(-2 pi i b_k j / n) X[j]
    for (i64 k = 0; k < rows; ++k) {
      const i64 b = s.out_map.empty() ? k : s.out_map[(size_t)k];
      for (i64 j = 0; j < cols; ++j) po_cis(b * j, n, -1.0, re[(size_t)(k + rows * j)], im[(size_t)(k + rows * j)]);
    }
  } else {
    // X[j] = (1/n) sum_b c_b exp(+2 pi i b j / n) Y[src(b)]   (c_b: Hermitian doubling for c2r)
    const i64 nspec = s.dft_real ? n / 2 + 1 : n;
    for (i64 b = 0; b < nspec; ++b) {
      const i64 src = s.in_map.empty() ? b : s.in_map[(size_t)b];
      if (src < 0) continue;
      double w = 1.0 / (double)n;
      if (s.dft_real && !(b == 0 || 2 * b == n)) w *= 2.0;
      for (i64 j = 0; j < rows; ++j) {
        double c, sn;
        po_cis(b * j, n, +1.0, c, sn);
        re[(size_t)(j + rows * src)] = w * c;
        im[(size_t)(j + rows * src)] = w * sn;
      }
    }
  }
  s.table_dt = dbl ? PO_C128 : PO_C64;
  s.a_rs = 1; s.a_cs = rows; s.conj_a = 0;
  if (dbl) {
    std::vector<double2> h((size_t)(rows * cols));
    for (size_t e = 0; e < h.size(); ++e) h[e] = make_double2(re[e], im[e]);
    return po_upload(pl, h, &s.d_table);
  }
  std::vector<float2> h((size_t)(rows * cols));
  for (size_t e = 0; e < h.size(); ++e) h[e] = make_float2((float)re[e], (float)im[e]);
  return po_upload(pl, h, &s.d_table);
}

static int po_build_twiddles(po_plan* pl, po_step& s) {
  const i64 n = s.dft_n;
  const bool dbl = po_dtype_is_double(s.in_dt);
  const i64 half = n / 2;
  if (dbl) {
    std::vector<double2> h((size_t)half);
    for (i64 k = 0; k < half; ++k) { double c, sn; po_cis(k, n, -1.0, c, sn); h[(size_t)k] = make_double2(c, sn); }
    return po_upload(pl, h, &s.d_table);
  }
  std::vector<float2> h((size_t)half);
  for (i64 k = 0; k < half; ++k) { double c, sn; po_cis(k, n, -1.0, c, sn); h[(size_t)k] = make_float2((float)c, (float)sn); }
  return po_upload(pl, h, &s.d_table);
}

static int po_realise_step(po_plan* pl, po_step& s) {
  char note[256];
  switch (s.kind) {
    case ST_DFT: {
      const bool dbl = po_dtype_is_double(s.in_dt);
      const i64 kept = s.dft_inverse ? s.n : s.m;  // rows actually read (inverse) / written (forward)
      const bool truncated = !s.in_map.empty() || !s.out_map.empty();
      bool use_fft = po_fft_supported(s.dft_n, dbl);
      if (use_fft && truncated && kept <= 32 && 2 * kept <= s.dft_n) use_fft = false;
      if (s.dft_n == 1) use_fft = false;
      const bool c16_shape = !dbl && !s.dft_real && s.in_dt == PO_C64 && (s.dft_n == 32 || s.dft_n == 64 || s.dft_n == 128) && s.I >= 8 &&
                             !(pl->flags & PO_PLAN_NO_FASTPATH);
      const bool c16_fwd = c16_shape && !s.dft_inverse && s.in_map.empty() && po_is_lowhigh16(s.out_map, s.dft_n);
      const bool c16_inv = c16_shape && s.dft_inverse && s.out_map.empty() && po_is_lowhigh16_scatter(s.in_map, s.dft_n) && s.n == 16;
      // real transform, 16 low bins kept (C2's restricted rfft on the slow dim)
      const bool r16_shape = !dbl && s.dft_real && (s.dft_n == 32 || s.dft_n == 64 || s.dft_n == 128) && s.I >= 8 && !(pl->flags & PO_PLAN_NO_FASTPATH);
      const bool r16_fwd = r16_shape && !s.dft_inverse && s.in_dt == PO_F32 && s.in_map.empty() && po_is_low16(s.out_map);
      const bool r16_inv = r16_shape && s.dft_inverse && s.out_dt == PO_F32 && s.out_map.empty() && po_is_low16_scatter(s.in_map) &&
                           (i64)s.in_map.size() == s.dft_n / 2 + 1 && s.n == 16;
      if (r16_fwd || r16_inv) {
        s.impl = IM_R16;
        s.ctw = new po_c16_tw();
        const int R = (int)s.dft_n / 16;
        for (int j2 = 0; j2 < 8; ++j2)
          for (int r = 0; r < 16; ++r) {
            double c = 0, sn = 0;
            if (j2 < R) po_cis((i64)r * j2, s.dft_n, r16_inv ? +1.0 : -1.0, c, sn);
            const double w = r16_inv ? ((r == 0) ? 1.0 : 2.0) / (double)s.dft_n : 1.0;
            s.ctw->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(w * sn));
          }
        snprintf(note, sizeof(note), "pruned-r16 %s n=%lld kept=16", s.dft_inverse ? "c2r inv" : "r2c", (long long)s.dft_n);
      } else if (!dbl && !s.dft_real && s.in_dt == PO_C64 && !truncated && (s.dft_n == 32 || s.dft_n == 64 || s.dft_n == 128) && (s.I >= 8 || s.I == 1) &&
                 !(pl->flags & PO_PLAN_NO_FASTPATH)) {
        s.impl = IM_CFULL;
        s.ctw = new po_c16_tw();
        const int R2 = (int)s.dft_n / 16;
        for (int j2 = 0; j2 < 8; ++j2)
          for (int k1 = 0; k1 < 16; ++k1) {
            double c = 0, sn = 0;
            if (j2 < R2) po_cis((i64)j2 * k1, s.dft_n, s.dft_inverse ? +1.0 : -1.0, c, sn);
            s.ctw->tw[j2 * 16 + k1] = make_float2((float)c, (float)sn);
          }
        snprintf(note, sizeof(note), "fft16x%d c2c%s n=%lld", R2, s.dft_inverse ? " inv" : "", (long long)s.dft_n);
      } else if (c16_fwd || c16_inv) {
        s.impl = IM_C16;
        s.ctw = new po_c16_tw();
        const int R = (int)s.dft_n / 16;
        for (int j2 = 0; j2 < 8; ++j2)
          for (int r = 0; r < 16; ++r) {
            double c = 0, sn = 0;
            const i64 bin = r < 8 ? r : s.dft_n - 16 + r;
            if (j2 < R) po_cis(bin * j2, s.dft_n, c16_inv ? +1.0 : -1.0, c, sn);
            const double w = c16_inv ? 1.0 / (double)s.dft_n : 1.0;
            s.ctw->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(w * sn));
          }
        snprintf(note, sizeof(note), "pruned-c16 c2c%s n=%lld kept=16", s.dft_inverse ? " inv" : "", (long long)s.dft_n);
      } else if (use_fft) {
        s.impl = IM_FFT;
        s.fft_mode = s.dft_real ? (s.dft_inverse ? 2 : 1) : 0;
        int rc = po_build_twiddles(pl, s);
        if (rc) return rc;
        if ((rc = po_upload(pl, s.in_map, (void**)&s.d_in_map))) return rc;
        if ((rc = po_upload(pl, s.out_map, (void**)&s.d_out_map))) return rc;
        snprintf(note, sizeof(note), "fft %s%s n=%lld%s", s.dft_real ? (s.dft_inverse ? "c2r" : "r2c") : "c2c",
                 s.dft_inverse ? " inv" : "", (long long)s.dft_n, truncated ? " +restrict" : "");
      } else {
        if ((double)s.n * (double)s.m > 16e6)
          return po_fail(PO_ERR_UNSUPPORTED, "ParDFT(n=%lld): only n = 2^p <= %d or dense-matrix sizes are supported", (long long)s.dft_n, dbl ? 2048 : 4096);
        s.impl = IM_DENSE_TABLE;
        int rc = po_build_dft_matrix(pl, s);
        if (rc) return rc;
        snprintf(note, sizeof(note), "dense-dft %s%s n=%lld kept=%lld", s.dft_real ? (s.dft_inverse ? "c2r" : "r2c") : "c2c",
                 s.dft_inverse ? " inv" : "", (long long)s.dft_n, (long long)kept);
      }
      break;
    }
    case ST_GATHER: {
      s.impl = IM_GATHER;
      int rc = po_upload(pl, s.map, (void**)&s.d_map);
      if (rc) return rc;
      snprintf(note, sizeof(note), "gather%s", s.m > s.n ? " zero-fill" : "");
      break;
    }
    case ST_MATRIX:
      s.impl = IM_DENSE_PARAM;
      s.table_dt = s.in_dt;
      if (!s.adjoint) { s.a_rs = 1; s.a_cs = s.prm_rows; s.conj_a = 0; }
      else { s.a_rs = s.prm_rows; s.a_cs = 1; s.conj_a = 1; }
      snprintf(note, sizeof(note), "matrix%s %lldx%lld", s.adjoint ? "'" : "", (long long)s.prm_rows, (long long)s.prm_cols);
      break;
    case ST_DIAG: s.impl = IM_DIAG; snprintf(note, sizeof(note), "diag%s", s.adjoint ? "'" : ""); break;
    case ST_TENSOR: s.impl = IM_TENSOR; snprintf(note, sizeof(note), "tensor-mix ic=%d oc=%d M=%lld", s.ic, s.oc, (long long)s.M); break;
    case ST_BIAS: s.impl = IM_BIAS; snprintf(note, sizeof(note), "bias"); break;
    case ST_GELU: s.impl = IM_GELU; snprintf(note, sizeof(note), "gelu"); break;
    case ST_REPART: s.impl = IM_REPART; snprintf(note, sizeof(note), "repartition"); break;
    case ST_MIXDIAG:
      s.impl = IM_MIXDIAG;
      s.table_dt = s.in_dt;
      if (!s.adjoint) { s.a_rs = 1; s.a_cs = s.prm_rows; s.conj_a = 0; }
      else { s.a_rs = s.prm_rows; s.a_cs = 1; s.conj_a = 1; }
      snprintf(note, sizeof(note), "mix+diag matrix%s %lldx%lld diag%s %lld", s.adjoint ? "'" : "", (long long)s.prm_rows, (long long)s.prm_cols,
               s.adjoint2 ? "'" : "", (long long)s.nd);
      break;
    case ST_ZMZ: {
      s.impl = IM_ZMZ;
      s.table_dt = s.in_dt;
      if (!s.adjoint) { s.a_rs = 1; s.a_cs = s.prm_rows; s.conj_a = 0; }
      else { s.a_rs = s.prm_rows; s.a_cs = 1; s.conj_a = 1; }
      s.ctw = new po_c16_tw(); s.ctw2 = new po_c16_tw();
      const int R = (int)s.dft_n / 16;
      for (int j2 = 0; j2 < 8; ++j2)
        for (int r = 0; r < 16; ++r) {
          double c = 0, sn = 0;
          const i64 bin = r < 8 ? r : s.dft_n - 16 + r;
          if (j2 < R) po_cis(bin * j2, s.dft_n, -1.0, c, sn);
          s.ctw->tw[j2 * 16 + r] = make_float2((float)c, (float)sn);
          const double w = 1.0 / (double)s.dft_n;
          s.ctw2->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(-w * sn));
        }
      snprintf(note, sizeof(note), "z-mix-z: pruned-c16 n=%lld -> matrix%s %lldx%lld diag%s %lld -> pruned-c16 inv", (long long)s.dft_n, s.adjoint ? "'" : "",
               (long long)s.prm_rows, (long long)s.prm_cols, s.adjoint2 ? "'" : "", (long long)s.nd);
      break;
    }
    case ST_FUSED_TX:
      s.impl = IM_FUSED_TX;
      s.ftw = new po_fused_tw();
      po_fill_fused_tw(*s.ftw, s.f_nt, s.f_nx, s.dft_inverse != 0);
      {
        double ks[8];
        po_herm_kscale(s.f_nt, s.f_nx, ks);
        s.f2tw = new po_fused2_tw();
        po_fill_fused2_tw(*s.f2tw, s.f_nx, s.dft_inverse != 0, s.dft_inverse ? ks : nullptr);
      }
      snprintf(note, sizeof(note), "fused-%s nt=%d nx=%d kept=8x16 I0=%d mid=%lld B=%lld", s.dft_inverse ? "inv-xt(c2r)" : "fwd-tx(r2c)",
               s.f_nt, s.f_nx, s.f_I0, (long long)s.f_mid, (long long)s.f_B);
      break;
    default: return po_fail(PO_ERR_INVALID, "unknown step kind");
  }
  s.note = note;
  return PO_OK;
}

static int po_repart_prepare(po_plan* pl, po_step& s, const po_node* nodes, const int64_t* aux);  // po_comm.cuh
static int po_repart_run(po_plan* pl, po_step& s, const void* src, void* dst, cudaStream_t st);  // po_comm.cuh

static int po_run_step(po_plan* pl, po_step& s, const void* src, void* dst, const void* const* params, int n_params,
                       cudaStream_t st) {
  // a rank whose shard is empty has nothing to compute in a local step (the reference returns x for `0 in size(x)`); the repartition still
  // runs: an empty sender may receive, an empty receiver may send
  if (s.kind != ST_REPART && (s.I * s.n * s.O == 0 || po_step_out_elems(s) == 0)) return PO_OK;
  const void* prm = nullptr;
  if (s.param_slot >= 0) {
    if (s.param_slot >= n_params || !params || !params[s.param_slot])
      return po_fail(PO_ERR_INVALID, "parameter slot %d missing (got %d params)", s.param_slot, n_params);
    prm = params[s.param_slot];
  }
  cudaError_t e = cudaSuccess;
  switch (s.impl) {
    case IM_FFT:
      e = po_launch_fft(s.fft_mode, po_dtype_is_double(s.in_dt), src, dst, s.d_table, s.I, s.dft_n, s.O, s.dft_inverse,
                        s.dft_inverse ? 1.0 / (double)s.dft_n : 1.0, 1.0, s.d_in_map, (int)s.n, s.d_out_map, (int)s.m, st);
      break;
    case IM_DENSE_TABLE:
      e = po_launch_dense(s.table_dt, s.in_dt, s.out_dt, s.d_table, s.a_rs, s.a_cs, s.conj_a, src, dst, s.I, s.n, s.m, s.O, st);
      break;
    case IM_DENSE_PARAM:
      e = po_launch_dense(s.table_dt, s.in_dt, s.out_dt, prm, s.a_rs, s.a_cs, s.conj_a, src, dst, s.I, s.n, s.m, s.O, st);
      break;
    case IM_GATHER: e = po_launch_gather(s.in_dt, src, dst, s.d_map, s.I, s.n, s.m, s.O, st); break;
    case IM_DIAG: e = po_launch_diag(s.in_dt, prm, s.adjoint, src, dst, s.I, s.n, s.O, st); break;
    case IM_TENSOR:
      e = po_launch_tensor_stream(s.in_dt, false, prm, src, dst, s.ic, s.oc, s.M, s.O, s.w_oi, pl->ctx->num_sms, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported) e = po_launch_tensor_mix(s.in_dt, prm, src, dst, s.ic, s.oc, s.M, s.O, s.w_oi, st);
      break;
    case IM_BIAS: e = po_launch_bias(s.in_dt, prm, src, dst, s.I, s.n, s.O, st); break;
    case IM_GELU: e = po_launch_gelu(s.in_dt, src, nullptr, dst, s.I * s.n * s.O, st); break;
    case IM_REPART: return po_repart_run(pl, s, src, dst, st);
    case IM_C16: e = po_launch_c16(s.dft_inverse != 0, (int)s.dft_n, src, dst, *s.ctw, s.I, s.O, st); break;
    case IM_CFULL: e = po_launch_cfull(s.dft_inverse != 0, (int)s.dft_n, src, dst, *s.ctw, s.I, s.O, s.dft_inverse ? 1.f / (float)s.dft_n : 1.f, st); break;
    case IM_R16: e = po_launch_r16(s.dft_inverse != 0, (int)s.dft_n, src, dst, *s.ctw, s.I, s.O, true, st); break;
    case IM_MIXDIAG: {
      if (s.param_slot2 < 0 || s.param_slot2 >= n_params || !params[s.param_slot2])
        return po_fail(PO_ERR_INVALID, "parameter slot %d missing (got %d params)", s.param_slot2, n_params);
      e = po_launch_mix_diag(s.in_dt, prm, s.a_rs, s.a_cs, s.conj_a, params[s.param_slot2], s.adjoint2, s.nd, src, dst, s.n, s.m, s.O, st);
      break;
    }
    case IM_ZMZ: {
      if (s.param_slot2 < 0 || s.param_slot2 >= n_params || !params[s.param_slot2])
        return po_fail(PO_ERR_INVALID, "parameter slot %d missing (got %d params)", s.param_slot2, n_params);
      po_zmz_params P;
      memset(&P, 0, sizeof(P));
      P.X = (const float2*)src; P.Y = (float2*)dst; P.A = (const float2*)prm;
      P.b_rs = s.a_rs; P.b_cs = s.a_cs; P.conj_b = s.conj_a;
      P.d = (const float2*)params[s.param_slot2]; P.conj_d = s.adjoint2; P.diag_first = 0; P.nd = s.nd;
      P.nin = s.ic; P.nout = s.oc; P.Mi = s.z_Mi; P.O = s.O;
      P.tape_out = (float2*)pl->cur_tape_slot;
      e = po_launch_zmz(false, (int)s.dft_n, P, *s.ctw, *s.ctw2, st);
      break;
    }
    case IM_FUSED_TX:
      e = cudaErrorNotSupported;
      if (s.f2tw && !s.dft_inverse && po_use_gen3(pl))
        e = po_launch_fwd_ring(s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && s.f2tw && !s.dft_inverse && po_use_gen3(pl) && po_use_pair())
        e = po_launch_fwd_pair(s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, pl->ctx->d_err, st);
      if (s.f2tw && s.dft_inverse && po_use_gen3(pl) && po_use_inv_stage())
        e = po_launch_inv_stage(s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && s.f2tw && s.dft_inverse && po_use_gen3(pl) && po_use_inv_stage() && po_use_pair())
        e = po_launch_inv_pair(s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && s.f2tw && po_use_gen2(pl) && po_use_tile2(s.dft_inverse != 0, s.f_I0, s.f_nx))
        e = po_launch_tile2(s.dft_inverse != 0, s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, st);
      if (e == cudaErrorNotSupported && s.f2tw && po_use_gen2(pl))
        e = po_launch_fused2(s.dft_inverse != 0, s.f_nt, s.f_nx, src, dst, *s.f2tw, s.f_I0, s.f_mid, s.f_B, pl->ctx->num_sms, st);
      if (e != cudaErrorNotSupported) break;
      e = po_launch_fused_tx(s.dft_inverse != 0, s.f_nt, s.f_nx, src, dst, *s.ftw, s.f_I0, s.f_mid, s.f_B,
                             (pl->flags & PO_PLAN_NO_PIPELINE) ? 0 : pl->ctx->num_sms, st);
      break;
    default: return po_fail(PO_ERR_INVALID, "step not realised");
  }
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "kernel launch failed (%s): %s", s.note.c_str(), cudaGetErrorString(e));
  pl->launches += 1;
  return PO_OK;
}

// ---- L2-resident hand-over between the y-inverse and the fused inverse (x, t) kernel ---------------------------
// The inverse half of the spectral stack ends with  pruned-c16 inverse along y (21 MB -> 84 MB)  followed by the fused
// inverse (x, t) kernel (84 MB -> 671 MB and more).  ncu (profiles/r1u_*): the second kernel still fetches 48 of its
// 84 MB from DRAM -- the intermediate does not survive in L2 as a whole.  Running the PAIR in z-chunks (y-inverse of
// chunk k, fused inverse of chunk k, ...) makes every chunk's intermediate L2-resident when it is consumed, at the
// price of one more pipeline fill/drain of the persistent kernel (~5 us) per extra launch.
// Measured (profiles/r1_inverse_pair_chunks_ab.txt): at nt = 32 (C3) the pair LOSES 10 us with 2 chunks (0.155 ->
// 0.166 ms; so DRAM-read interference is not what holds that kernel below its store-only rate), at nt = 64 (the
// C4 ladder, kernel twice as long) it gains 2 % (0.2856 -> 0.2796 ms).  Default: 2 chunks from nt = 64 on, else off;
// PO_INV_CHUNKS overrides (1 = off).  The same pair appears in the pullback (cotangent of F_y, then of the fused
// forward), where it runs with the pullback's tables.
static int po_inv_pair_chunks(int nt) {
  const char* env = getenv("PO_INV_CHUNKS");  // read per call (host side, once per apply): tests switch it
  return env ? atoi(env) : (nt >= 64 ? 2 : 1);
}
// c: the c16 step (16 -> c.dft_n along y), f: the fused step; both run in their inverse-type direction
static bool po_inv_pair_ok(const po_plan* pl, const po_step& c, const po_step& f, const po_c16_tw* ctw, const po_fused2_tw* f2tw, const void* mid,
                           const void* dst) {
  const int chunks = po_inv_pair_chunks(f.f_nt);
  if (chunks < 2 || !ctw || !f2tw || !po_use_gen2(pl) || po_use_tile2(true, f.f_I0, f.f_nx)) return false;
  if (c.dft_n <= 0 || c.I != (i64)f.f_I0 * 16 || f.f_mid % c.dft_n) return false;
  const i64 period = f.f_mid / c.dft_n;  // extent of the slowest spatial dim (z)
  if (period % chunks || c.O != period * 8 * f.f_B) return false;
  return po_fused2_inv_ok(f.f_nt, f.f_nx, f.f_I0, f.f_mid, f.f_mid / chunks, f.f_B, pl->ctx->num_sms, mid, dst);
}
static int po_run_inv_pair(po_plan* pl, const po_step& c, const po_step& f, const po_c16_tw& ctw, const po_fused2_tw& f2tw, const void* src, void* mid,
                           void* dst, cudaStream_t st) {
  const int chunks = po_inv_pair_chunks(f.f_nt);
  const i64 period = f.f_mid / c.dft_n, zc = period / chunks;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] inverse pair in %d z-chunks of %lld (y-inverse n=%lld, fused inverse mid=%lld)\n", chunks, (long long)zc, (long long)c.dft_n, (long long)f.f_mid);
  for (int k = 0; k < chunks; ++k) {
    cudaError_t e = po_launch_c16(true, (int)c.dft_n, src, mid, ctw, c.I, c.O, st, k * zc, zc, period);
    if (e == cudaSuccess) {
      e = cudaErrorNotSupported;
      if (po_use_gen3(pl) && po_use_inv_stage())
        e = po_launch_inv_stage(f.f_nt, f.f_nx, mid, dst, f2tw, f.f_I0, f.f_mid, f.f_B, pl->ctx->num_sms, pl->ctx->d_err, st, k * zc * c.dft_n, zc * c.dft_n);
      if (e == cudaErrorNotSupported)
        e = po_launch_fused2(true, f.f_nt, f.f_nx, mid, dst, f2tw, f.f_I0, f.f_mid, f.f_B, pl->ctx->num_sms, st, k * zc * c.dft_n, zc * c.dft_n);
    }
    if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "chunked inverse pair failed (%s | %s): %s", c.note.c_str(), f.note.c_str(), cudaGetErrorString(e));
    pl->launches += 2;
  }
  return PO_OK;
}

// ---- C2's factor pair: pruned real DFT on the slow dim + full FFT128 on the fastest dim in one kernel (po_c2.cuh) ------------
// forward: a = the r16 step, b = the fft16x8 step (I == 1) that follows it; inverse: a = the fft16x8 step, b = the c2r step
static bool po_c2_pair_ok(const po_step& a, const po_step& b) {
  const char* env = getenv("PO_C2_FUSE");  // read per call (host side): tests and scripts/bench_configs.py switch it
  if (env && !atoi(env)) return false;
  const po_step& r = a.impl == IM_R16 ? a : b;
  const po_step& c = a.impl == IM_R16 ? b : a;
  if (r.impl != IM_R16 || c.impl != IM_CFULL || !r.ctw || !c.ctw) return false;
  const bool inv = r.dft_inverse != 0;
  if ((c.dft_inverse != 0) != inv || (&r == &a) == inv) return false;  // forward: r16 first; inverse: r16 last
  if (c.I != 1 || c.dft_n != 128 || r.I % 128) return false;
  return c.O == (r.I / 128) * 16 * r.O;
}
static int po_run_c2_pair(po_plan* pl, const po_step& a, const po_step& b, const void* src, void* dst, cudaStream_t st) {
  const po_step& r = a.impl == IM_R16 ? a : b;
  const po_step& c = a.impl == IM_R16 ? b : a;
  cudaError_t e = po_launch_c2(r.dft_inverse != 0, (int)r.dft_n, src, dst, *r.ctw, *c.ctw, r.I, r.O, st);
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "fused C2 pair failed (%s | %s): %s", a.note.c_str(), b.note.c_str(), cudaGetErrorString(e));
  pl->launches += 1;
  return PO_OK;
}

static int po_plan_compile(po_ctx* ctx, const po_node* nodes, int32_t n_nodes, const int32_t* child_index,
                           int32_t n_child_index, const int64_t* aux, int32_t n_aux, int32_t root, int64_t batch,
                           uint32_t flags, po_plan** out) {
  if (!ctx || !nodes || !out || n_nodes <= 0) return po_fail(PO_ERR_INVALID, "po_plan_create: null argument");
  if (batch < 0) return po_fail(PO_ERR_INVALID, "po_plan_create: negative batch");
  if (root < 0 || root >= n_nodes) return po_fail(PO_ERR_INVALID, "po_plan_create: root out of range");
  po_lowerer lw;
  lw.nodes = nodes; lw.n_nodes = n_nodes; lw.child = child_index; lw.n_child = n_child_index; lw.aux = aux; lw.n_aux = n_aux;
  lw.cur_dt = nodes[root].ddt;
  int rc = lw.lower(root, 1, batch, 0);
  if (rc) return rc;
  if (lw.cur_dt != nodes[root].rdt && !lw.steps.empty())
    return po_fail(PO_ERR_DTYPE, "operator produces %s but its range type says %s", po_dtype_name(lw.cur_dt), po_dtype_name(nodes[root].rdt));

  if (!(flags & PO_PLAN_NO_FUSION)) {
    po_schedule(lw.steps);
    po_fuse(lw.steps);
    if (!(flags & PO_PLAN_NO_FASTPATH)) { po_widen(lw.steps); po_fastpath(lw.steps); po_fastpath_zmz(lw.steps); }
  }

  po_plan* pl = new po_plan();
  pl->ctx = ctx; pl->batch = batch; pl->flags = flags;
  pl->domain = nodes[root].domain; pl->range = nodes[root].range;
  pl->ddt = nodes[root].ddt; pl->rdt = nodes[root].rdt;
  pl->n_params = lw.max_param + 1;
  pl->steps.swap(lw.steps);
  PO_CUDA_CHECK(cudaSetDevice(ctx->device));
  size_t half = 0;
  for (size_t i = 0; i < pl->steps.size(); ++i) {
    po_step& s = pl->steps[i];
    if (s.kind == ST_REPART) {
      if ((rc = po_repart_prepare(pl, s, nodes, aux))) { delete pl; return rc; }
    }
    if ((rc = po_realise_step(pl, s))) { delete pl; return rc; }
    const size_t ob = (size_t)po_step_out_elems(s) * po_dtype_size(s.out_dt);
    pl->step_out_bytes.push_back(ob);
    if (i + 1 < pl->steps.size() && ob > half) half = ob;
  }
  pl->ws_half = (half + 255) & ~(size_t)255;
  *out = pl;
  return PO_OK;
}

static size_t po_plan_ws_bytes(const po_plan* pl) { return pl->steps.size() > 1 ? 2 * pl->ws_half : 0; }

static bool po_step_needs_input(const po_step& s) {
  return s.kind == ST_MATRIX || s.kind == ST_DIAG || s.kind == ST_TENSOR || s.kind == ST_GELU || s.kind == ST_MIXDIAG || s.kind == ST_ZMZ;
}
// ST_ZMZ tapes the spectral tensor it forms internally (the input of its mixing), not its own input
static bool po_step_tapes_own(const po_step& s) { return s.kind == ST_ZMZ; }

// tape layout: the INPUT of every parametric / non-linear step, in step order
static size_t po_tape_layout(const po_plan* pl, std::vector<long long>* offs) {
  size_t off = 0;
  if (offs) offs->assign(pl->steps.size(), -1);
  for (size_t i = 0; i < pl->steps.size(); ++i) {
    const po_step& s = pl->steps[i];
    if (!po_step_needs_input(s)) continue;
    if (offs) (*offs)[i] = (long long)off;
    const size_t b = (s.kind == ST_ZMZ ? (size_t)(s.ic * s.z_Mi * 16 * s.O) : (size_t)(s.I * s.n * s.O)) * po_dtype_size(s.in_dt);
    off += (b + 255) & ~(size_t)255;
  }
  return off;
}
static size_t po_tape_bytes(const po_plan* pl) { return po_tape_layout(pl, nullptr); }

// y = A x.  With `tape` the input of every parametric step is left in the tape (the step
// that produces it writes straight into the tape slot -- no extra copy).
static int po_plan_run(po_plan* pl, const void* x, void* y, const void* const* params, int32_t n_params, void* ws,
                       size_t ws_bytes, cudaStream_t st, void* tape) {
  if (!pl) return po_fail(PO_ERR_INVALID, "null plan");
  if (int rc0 = po_ctx_check(pl->ctx)) return rc0;
  const size_t need = po_plan_ws_bytes(pl);
  if (need > ws_bytes || (need && !ws)) return po_fail(PO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  PO_CUDA_CHECK(cudaSetDevice(pl->ctx->device));
  pl->launches = 0;
  if (pl->batch == 0) return PO_OK;  // no columns: y is (Range, 0), nothing to compute (`0 in size(x)` of src/ParDFT.jl:26,30); x / y may be NULL
  const size_t nsteps = pl->steps.size();
  if (nsteps == 0) {  // identity: the reference returns x itself; across the ABI that is a copy
    const size_t bytes = (size_t)(pl->domain * pl->batch) * po_dtype_size(pl->ddt);
    if (x != y && bytes) PO_CUDA_CHECK(cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice, st));
    return PO_OK;
  }
  // an EMPTY side (a rank that owns no rows of a distributed operand: `0 in size(x)`, src/ParCommon.jl:73) may be NULL
  if ((!x && pl->domain * pl->batch > 0) || (!y && pl->range * pl->batch > 0)) return po_fail(PO_ERR_INVALID, "null x or y");
  std::vector<long long> toff;
  if (tape) po_tape_layout(pl, &toff);
  const void* src = x;
#ifndef PO_EMUL
  std::vector<cudaEvent_t> evs;
  if (pl->profiling) {
    evs.resize(nsteps + 1);
    for (auto& e : evs) PO_CUDA_CHECK(cudaEventCreate(&e));
    PO_CUDA_CHECK(cudaEventRecord(evs[0], st));
  }
#endif
  if (tape && toff[0] >= 0 && !po_step_tapes_own(pl->steps[0])) {
    const size_t b = (size_t)(pl->steps[0].I * pl->steps[0].n * pl->steps[0].O) * po_dtype_size(pl->steps[0].in_dt);
    if (b) PO_CUDA_CHECK(cudaMemcpyAsync((char*)tape + toff[0], x, b, cudaMemcpyDeviceToDevice, st));
  }
  auto dst_of = [&](size_t i) -> void* {
    if (i + 1 == nsteps) return y;
    if (tape && toff[i + 1] >= 0 && !po_step_tapes_own(pl->steps[i + 1])) return (char*)tape + toff[i + 1];
    return (char*)ws + (i % 2) * pl->ws_half;
  };
  for (size_t i = 0; i < nsteps; ++i) {
    void* dst = dst_of(i);
    if (i + 1 < nsteps && pl->steps[i].impl == IM_C16 && pl->steps[i].dft_inverse && pl->steps[i + 1].impl == IM_FUSED_TX &&
        pl->steps[i + 1].dft_inverse && po_inv_pair_ok(pl, pl->steps[i], pl->steps[i + 1], pl->steps[i].ctw, pl->steps[i + 1].f2tw, dst, dst_of(i + 1))) {
      void* out = dst_of(i + 1);
#ifndef PO_EMUL
      if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[i + 1], st));  // the pair's time is booked on the fused step
#endif
      int rc = po_run_inv_pair(pl, pl->steps[i], pl->steps[i + 1], *pl->steps[i].ctw, *pl->steps[i + 1].f2tw, src, dst, out, st);
      if (rc) return rc;
#ifndef PO_EMUL
      if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[i + 2], st));
#endif
      src = out;
      ++i;
      continue;
    }
    if (i + 1 < nsteps && (pl->steps[i].impl == IM_R16 || pl->steps[i].impl == IM_CFULL) && po_c2_pair_ok(pl->steps[i], pl->steps[i + 1])) {
      void* out = dst_of(i + 1);
#ifndef PO_EMUL
      if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[i + 1], st));  // the pair's time is booked on its second step
#endif
      int rc = po_run_c2_pair(pl, pl->steps[i], pl->steps[i + 1], src, out, st);
      if (rc) return rc;
#ifndef PO_EMUL
      if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[i + 2], st));
#endif
      src = out;
      ++i;
      continue;
    }
    pl->cur_tape_slot = (tape && toff[i] >= 0 && po_step_tapes_own(pl->steps[i])) ? (void*)((char*)tape + toff[i]) : nullptr;
    int rc = po_run_step(pl, pl->steps[i], src, dst, params, n_params, st);
    pl->cur_tape_slot = nullptr;
    if (rc) return rc;
#ifndef PO_EMUL
    if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[i + 1], st));
#endif
    src = dst;
  }
#ifndef PO_EMUL
  if (pl->profiling) pl->pending_events.push_back(evs);
#endif
  return PO_OK;
}

// fold finished event sets into step_ms / step_calls (waits for the recorded work)
static int po_plan_collect_profile(po_plan* pl) {
  if (pl->step_ms.size() != pl->steps.size()) { pl->step_ms.assign(pl->steps.size(), 0.0); pl->step_calls.assign(pl->steps.size(), 0); }
  if (pl->pb_step_ms.size() != pl->steps.size()) { pl->pb_step_ms.assign(pl->steps.size(), 0.0); pl->pb_step_calls.assign(pl->steps.size(), 0); }
#ifndef PO_EMUL
  for (auto& evs : pl->pending_pb_events) {  // evs[j] .. evs[j+1] brackets step nsteps-1-j
    PO_CUDA_CHECK(cudaEventSynchronize(evs.back()));
    const size_t ns = evs.size() - 1;
    for (size_t j = 0; j < ns; ++j) {
      float ms = 0.f;
      PO_CUDA_CHECK(cudaEventElapsedTime(&ms, evs[j], evs[j + 1]));
      pl->pb_step_ms[ns - 1 - j] += ms;
      pl->pb_step_calls[ns - 1 - j] += 1;
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
  }
  pl->pending_pb_events.clear();
#endif
#ifndef PO_EMUL
  for (auto& evs : pl->pending_events) {
    PO_CUDA_CHECK(cudaEventSynchronize(evs.back()));
    for (size_t i = 0; i + 1 < evs.size(); ++i) {
      float ms = 0.f;
      PO_CUDA_CHECK(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
      pl->step_ms[i] += ms;
      pl->step_calls[i] += 1;
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
  }
  pl->pending_events.clear();
#endif
  return PO_OK;
}
