// Distribution: ParRepartition (src/ParRepartition.jl), ParBroadcasted parameter traffic
// (src/ParBroadcasted.jl:26-57) and ParReduce (src/ParReduce.jl) over NCCL / NVLink.
//
// B200-first: the reference stages every message through host-side per-peer buffers and
// blocking MPI point-to-point calls (and round-trips GPU parameters through the host).  Here
//  - the cartesian bookkeeping (MPI_Cart_*) is pure host integer arithmetic, bit-exact with
//    the reference's row-major rank <-> coords convention;
//  - one pack kernel per peer writes into a single staging buffer, the exchange is ONE
//    grouped ncclSend/ncclRecv (all-to-all-v over NVSwitch) on the caller's stream, one
//    unpack kernel per peer scatters into y; the self-overlap is a direct strided copy;
//  - NCCL is resolved with dlopen at communicator creation (the library has no link-time
//    NCCL dependency, so a host process that already loaded NCCL -- torch, NCCL.jl -- shares it).
//  - po_comm_init_external lets the host supply the exchange (e.g. CUDA-aware MPI.jl, or
//    gloo in the CPU tests) behind the same pack/unpack kernels.
#pragma once
#include "po_plan.cuh"

#ifndef PO_EMUL
#include <dlfcn.h>
#endif

// ---- cartesian helpers (MPI_Cart_create with reorder = false: last dim fastest) ---------
static void po_cart_coords(const std::vector<int>& dims, int rank, std::vector<int>& coords) {
  coords.assign(dims.size(), 0);
  for (size_t d = dims.size(); d-- > 0;) {
    coords[d] = rank % dims[d];
    rank /= dims[d];
  }
}
static int po_cart_rank(const std::vector<int>& dims, const std::vector<int>& coords) {
  int r = 0;
  for (size_t d = 0; d < dims.size(); ++d) r = r * dims[d] + coords[d];
  return r;
}
static i64 po_local_size_impl(i64 global_size, i64 rank, i64 num_ranks) {  // src/ParCommon.jl:57-64
  i64 s = global_size / num_ranks;
  if (rank < global_size % num_ranks) s += 1;
  return s;
}
// 1-based inclusive [start, stop] of block `c` of `p` over `g` (src/ParCommon.jl:45-55)
static void po_block_range(i64 g, int c, int p, i64& start, i64& stop) {
  i64 off = 0;
  for (int i = 0; i < c; ++i) off += po_local_size_impl(g, i, p);
  start = off + 1;
  stop = off + po_local_size_impl(g, c, p);
}

struct po_box_rec {
  int peer;
  std::vector<i64> lo, hi;  // 1-based inclusive local ranges (hi < lo: empty)
  i64 count() const {
    i64 c = 1;
    for (size_t d = 0; d < lo.size(); ++d) c *= (hi[d] >= lo[d]) ? (hi[d] - lo[d] + 1) : 0;
    return c;
  }
};

// One side of the ParRepartition constructor (src/ParRepartition.jl:44-97): boxes of `mine`
// (this rank's block in grid dims_mine) against every overlapping block of grid dims_other.
static bool po_repart_side(int N, const std::vector<i64>& g, const std::vector<int>& dims_mine,
                           const std::vector<int>& dims_other, int rank, std::vector<i64>& local,
                           std::vector<po_box_rec>& out) {
  std::vector<int> cm;
  po_cart_coords(dims_mine, rank, cm);
  std::vector<i64> ms((size_t)N), me((size_t)N);
  local.resize((size_t)N);
  for (int d = 0; d < N; ++d) {
    po_block_range(g[(size_t)d], cm[(size_t)d], dims_mine[(size_t)d], ms[(size_t)d], me[(size_t)d]);
    local[(size_t)d] = me[(size_t)d] - ms[(size_t)d] + 1;
  }
  // per axis: findfirst:findlast of peer coords whose block intersects mine
  std::vector<int> first((size_t)N, -1), last((size_t)N, -1);
  bool ok = true;
  for (int d = 0; d < N; ++d) {
    for (int c = 0; c < dims_other[(size_t)d]; ++c) {
      i64 s, e;
      po_block_range(g[(size_t)d], c, dims_other[(size_t)d], s, e);
      const i64 is = s > ms[(size_t)d] ? s : ms[(size_t)d], ie = e < me[(size_t)d] ? e : me[(size_t)d];
      if (ie >= is) {
        if (first[(size_t)d] < 0) first[(size_t)d] = c;
        last[(size_t)d] = c;
      }
    }
    if (first[(size_t)d] < 0) ok = false;
  }
  out.clear();
  if (!ok) return true;
  // Iterators.product: first axis fastest
  std::vector<int> idx(first);
  while (true) {
    po_box_rec b;
    b.peer = po_cart_rank(dims_other, idx);
    b.lo.resize((size_t)N);
    b.hi.resize((size_t)N);
    for (int d = 0; d < N; ++d) {
      i64 s, e;
      po_block_range(g[(size_t)d], idx[(size_t)d], dims_other[(size_t)d], s, e);
      const i64 is = s > ms[(size_t)d] ? s : ms[(size_t)d], ie = e < me[(size_t)d] ? e : me[(size_t)d];
      b.lo[(size_t)d] = is - ms[(size_t)d] + 1;
      b.hi[(size_t)d] = ie - ms[(size_t)d] + 1;
    }
    out.push_back(b);
    int d = 0;
    for (; d < N; ++d) {
      if (++idx[(size_t)d] <= last[(size_t)d]) break;
      idx[(size_t)d] = first[(size_t)d];
    }
    if (d == N) break;
  }
  return true;
}

struct po_repart_info {
  int N = 0, rank = 0, P = 1;
  std::vector<i64> global, local_in, local_out;
  std::vector<int> dims_in, dims_out;
  std::vector<po_box_rec> send, recv;
  std::vector<i64> send_off, recv_off;  // element offsets (x batch) into the staging buffers
  i64 send_total = 0, recv_total = 0;   // elements per batch column
  void* d_send = nullptr;
  void* d_recv = nullptr;
};

static int po_repart_build(int N, const int64_t* global_size, const int32_t* dims_in, const int32_t* dims_out, int rank,
                           po_repart_info& info) {
  if (N < 1 || N > PO_BOX_MAXD - 1) return po_fail(PO_ERR_UNSUPPORTED, "ParRepartition: 1 <= ndim <= %d", PO_BOX_MAXD - 1);
  info.N = N; info.rank = rank;
  info.global.assign(global_size, global_size + N);
  info.dims_in.assign(dims_in, dims_in + N);
  info.dims_out.assign(dims_out, dims_out + N);
  i64 pin = 1, pout = 1;
  for (int d = 0; d < N; ++d) {
    if (dims_in[d] < 1 || dims_out[d] < 1 || global_size[d] < 0) return po_fail(PO_ERR_INVALID, "ParRepartition: bad dims");
    pin *= dims_in[d];
    pout *= dims_out[d];
  }
  if (pin != pout) return po_fail(PO_ERR_SHAPE, "ParRepartition: prod(dims_in)=%lld != prod(dims_out)=%lld (comm_in/comm_out must cover the same ranks)", (long long)pin, (long long)pout);
  if (rank < 0 || rank >= pin) return po_fail(PO_ERR_INVALID, "ParRepartition: rank %d outside communicator of size %lld", rank, (long long)pin);
  info.P = (int)pin;
  po_repart_side(N, info.global, info.dims_in, info.dims_out, rank, info.local_in, info.send);
  po_repart_side(N, info.global, info.dims_out, info.dims_in, rank, info.local_out, info.recv);
  return PO_OK;
}

// ---- transports -----------------------------------------------------------------------
struct po_nccl_id { char internal[128]; };  // ncclUniqueId
struct po_nccl_api {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ po_nccl_id, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
enum { PO_NCCL_CHAR = 0, PO_NCCL_F32 = 7, PO_NCCL_F64 = 8, PO_NCCL_SUM = 0 };

struct po_comm_state {
  int n_ranks = 1, rank = 0;
  void* nccl_comm = nullptr;
  po_exchange_fn ext_exchange = nullptr;
  po_collective_fn ext_collective = nullptr;
  void* ext_user = nullptr;
};

static po_nccl_api g_nccl;
static std::mutex g_nccl_mu;

static int po_nccl_load() {
#ifdef PO_EMUL
  return po_fail(PO_ERR_NCCL, "NCCL is not available in the CPU kernel emulator; use po_comm_init_external");
#else
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return PO_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return po_fail(PO_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define PO_SYM(field, name)                                                               \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                              \
  if (!g_nccl.field) return po_fail(PO_ERR_NCCL, "libnccl: missing symbol %s", name);
  PO_SYM(GetUniqueId, "ncclGetUniqueId")
  PO_SYM(CommInitRank, "ncclCommInitRank")
  PO_SYM(CommDestroy, "ncclCommDestroy")
  PO_SYM(Send, "ncclSend")
  PO_SYM(Recv, "ncclRecv")
  PO_SYM(GroupStart, "ncclGroupStart")
  PO_SYM(GroupEnd, "ncclGroupEnd")
  PO_SYM(Broadcast, "ncclBroadcast")
  PO_SYM(Reduce, "ncclReduce")
  PO_SYM(AllReduce, "ncclAllReduce")
  PO_SYM(GetErrorString, "ncclGetErrorString")
#undef PO_SYM
  g_nccl.lib = h;
  return PO_OK;
#endif
}
#define PO_NCCL_CHECK(expr)                                                                              \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != 0) return po_fail(PO_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
  } while (0)

// ---- ParRepartition apply (src/ParRepartition.jl:140-192) -------------------------------
static int po_repart_prepare(po_plan* pl, po_step& s, const po_node* nodes, const int64_t* aux) {
  const po_node& nd = nodes[s.node_idx];
  const int64_t* a = aux + nd.aux_begin;
  if (nd.aux_count < 1) return po_fail(PO_ERR_INVALID, "ParRepartition: aux too short");
  const int N = (int)a[0];
  if (nd.aux_count != 2 + 3 * N) return po_fail(PO_ERR_INVALID, "ParRepartition: aux must be [N, global[N], dims_in[N], dims_out[N], rank]");
  std::vector<int32_t> din((size_t)N), dout((size_t)N);
  for (int d = 0; d < N; ++d) { din[(size_t)d] = (int32_t)a[1 + N + d]; dout[(size_t)d] = (int32_t)a[1 + 2 * N + d]; }
  po_repart_info* info = new po_repart_info();
  int rc = po_repart_build(N, a + 1, din.data(), dout.data(), (int)a[1 + 3 * N], *info);
  if (rc) { delete info; return rc; }
  i64 pin = 1, pout = 1;
  for (int d = 0; d < N; ++d) { pin *= info->local_in[(size_t)d]; pout *= info->local_out[(size_t)d]; }
  if (pin != s.n || pout != s.m) {
    delete info;
    return po_fail(PO_ERR_SHAPE, "ParRepartition: Domain/Range %lld/%lld do not match the local blocks %lld/%lld", (long long)s.n, (long long)s.m, (long long)pin, (long long)pout);
  }
  const size_t esz = po_dtype_size(s.in_dt);
  const i64 batch = s.O;
  info->send_off.clear(); info->recv_off.clear();
  i64 so = 0, ro = 0;
  for (auto& b : info->send) { info->send_off.push_back(so); if (b.peer != info->rank) so += b.count() * batch; }
  for (auto& b : info->recv) { info->recv_off.push_back(ro); if (b.peer != info->rank) ro += b.count() * batch; }
  info->send_total = so; info->recv_total = ro;
  if (so) { PO_CUDA_CHECK(cudaMalloc(&info->d_send, (size_t)so * esz)); pl->owned.push_back(info->d_send); }
  if (ro) { PO_CUDA_CHECK(cudaMalloc(&info->d_recv, (size_t)ro * esz)); pl->owned.push_back(info->d_recv); }
  s.repart = info;  // leaked with the plan's lifetime; freed in po_plan_destroy via owned list for device memory
  return PO_OK;
}

static void po_make_box(const po_repart_info& info, const std::vector<i64>& local, const po_box_rec& b, i64 batch,
                        bool local_is_src, po_box& box, i64& base_off) {
  // dims 0..N-1 are the tensor dims (fastest first), dim N is the batch (slowest)
  const int N = info.N;
  box.nd = N + 1;
  i64 lstride = 1, pstride = 1;
  base_off = 0;
  for (int d = 0; d < N; ++d) {
    const i64 ext = (b.hi[(size_t)d] >= b.lo[(size_t)d]) ? (b.hi[(size_t)d] - b.lo[(size_t)d] + 1) : 0;
    box.ext[d] = ext;
    base_off += (b.lo[(size_t)d] - 1) * lstride;
    if (local_is_src) { box.sstride[d] = lstride; box.dstride[d] = pstride; }
    else { box.sstride[d] = pstride; box.dstride[d] = lstride; }
    lstride *= local[(size_t)d];
    pstride *= ext;
  }
  box.ext[N] = batch;
  if (local_is_src) { box.sstride[N] = lstride; box.dstride[N] = pstride; }
  else { box.sstride[N] = pstride; box.dstride[N] = lstride; }
}

static int po_repart_run(po_plan* pl, po_step& s, const void* src, void* dst, cudaStream_t st) {
  po_repart_info& info = *s.repart;
  po_comm_state* cm = pl->ctx->comm;
  const size_t esz = po_dtype_size(s.in_dt);
  const i64 batch = s.O;
  bool remote = false;
  for (auto& b : info.send) if (b.peer != info.rank && b.count() > 0) remote = true;
  for (auto& b : info.recv) if (b.peer != info.rank && b.count() > 0) remote = true;
  if (remote && (!cm || cm->n_ranks != info.P || cm->rank != info.rank))
    return po_fail(PO_ERR_NCCL, "ParRepartition over %d ranks needs a communicator of that size on this context (po_comm_init_rank)", info.P);

  // pack / self copy
  for (size_t k = 0; k < info.send.size(); ++k) {
    const po_box_rec& b = info.send[k];
    if (b.count() == 0) continue;
    po_box box;
    i64 base;
    if (b.peer != info.rank) {
      po_make_box(info, info.local_in, b, batch, true, box, base);
      cudaError_t e = po_launch_box_copy(esz, (const char*)src + (size_t)base * esz, (char*)info.d_send + (size_t)info.send_off[k] * esz, box, st);
      if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "repartition pack failed: %s", cudaGetErrorString(e));
      pl->launches += 1;
    } else {
      // self overlap: strided -> strided (src/ParRepartition.jl:176-178)
      const po_box_rec* rb = nullptr;
      for (auto& r : info.recv) if (r.peer == info.rank) rb = &r;
      if (!rb) continue;
      po_box sb, db;
      i64 sbase, dbase;
      po_make_box(info, info.local_in, b, batch, true, sb, sbase);
      po_make_box(info, info.local_out, *rb, batch, false, db, dbase);
      for (int d = 0; d <= info.N; ++d) sb.dstride[d] = db.dstride[d];
      cudaError_t e = po_launch_box_copy(esz, (const char*)src + (size_t)sbase * esz, (char*)dst + (size_t)dbase * esz, sb, st);
      if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "repartition self copy failed: %s", cudaGetErrorString(e));
      pl->launches += 1;
    }
  }
  // exchange
  if (remote) {
    std::vector<int32_t> sp, rp;
    std::vector<const void*> sbuf;
    std::vector<void*> rbuf;
    std::vector<size_t> sby, rby;
    for (size_t k = 0; k < info.send.size(); ++k) {
      const po_box_rec& b = info.send[k];
      if (b.peer == info.rank || b.count() == 0) continue;
      sp.push_back(b.peer); sbuf.push_back((char*)info.d_send + (size_t)info.send_off[k] * esz); sby.push_back((size_t)(b.count() * batch) * esz);
    }
    for (size_t k = 0; k < info.recv.size(); ++k) {
      const po_box_rec& b = info.recv[k];
      if (b.peer == info.rank || b.count() == 0) continue;
      rp.push_back(b.peer); rbuf.push_back((char*)info.d_recv + (size_t)info.recv_off[k] * esz); rby.push_back((size_t)(b.count() * batch) * esz);
    }
    if (cm->ext_exchange) {
      int rc = cm->ext_exchange(cm->ext_user, (int32_t)sp.size(), sp.data(), sbuf.data(), sby.data(), (int32_t)rp.size(), rp.data(), rbuf.data(), rby.data(), (void*)st);
      if (rc) return po_fail(PO_ERR_NCCL, "external exchange callback failed with %d", rc);
    } else {
      PO_NCCL_CHECK(g_nccl.GroupStart());
      for (size_t k = 0; k < rp.size(); ++k) PO_NCCL_CHECK(g_nccl.Recv(rbuf[k], rby[k], PO_NCCL_CHAR, rp[k], cm->nccl_comm, st));
      for (size_t k = 0; k < sp.size(); ++k) PO_NCCL_CHECK(g_nccl.Send(sbuf[k], sby[k], PO_NCCL_CHAR, sp[k], cm->nccl_comm, st));
      PO_NCCL_CHECK(g_nccl.GroupEnd());
    }
  }
  // unpack
  for (size_t k = 0; k < info.recv.size(); ++k) {
    const po_box_rec& b = info.recv[k];
    if (b.peer == info.rank || b.count() == 0) continue;
    po_box box;
    i64 base;
    po_make_box(info, info.local_out, b, batch, false, box, base);
    cudaError_t e = po_launch_box_copy(esz, (const char*)info.d_recv + (size_t)info.recv_off[k] * esz, (char*)dst + (size_t)base * esz, box, st);
    if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "repartition unpack failed: %s", cudaGetErrorString(e));
    pl->launches += 1;
  }
  return PO_OK;
}
