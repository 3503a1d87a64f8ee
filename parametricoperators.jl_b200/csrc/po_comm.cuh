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
  // peer-store path (po_repart_peer_kernel)
  bool peer = false;
  int slot = 0;
  size_t region_off = 0, region_bytes = 0;  // per-parity staging region inside the symmetric buffer
  long long epoch = 0;
  void* d_desc = nullptr;                   // po_peer_desc[] (send boxes then recv boxes)
  int n_send_desc = 0, n_recv_desc = 0;
  i64 peer_work = 0;                        // elements one exchange copies (grid sizing)
  long long* h_dbg = nullptr;               // PO_PEER_TRACE stamps (host-mapped; leaked on purpose: debug only)
  long long* d_dbg = nullptr;
  int* d_done = nullptr;
  std::vector<i64> send_peer_off;           // element offset of my box inside the RECEIVER's staging region
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
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
enum { PO_NCCL_CHAR = 0, PO_NCCL_F32 = 7, PO_NCCL_F64 = 8, PO_NCCL_SUM = 0 };

#define PO_SYM_SLOTS 256            // repartition steps that can own a flag row
#define PO_SYM_ACK_OFF (PO_SYM_SLOTS * 64 * 8)      // data flags [slot][source], then ack flags [slot][receiver]
#define PO_SYM_FLAG_BYTES (2 * PO_SYM_SLOTS * 64 * 8)

struct po_comm_state {
  int n_ranks = 1, rank = 0;
  void* nccl_comm = nullptr;
  po_exchange_fn ext_exchange = nullptr;
  po_collective_fn ext_collective = nullptr;
  void* ext_user = nullptr;
  // NVLink peer path: one symmetric buffer per rank, IPC-mapped into every peer.  Layout:
  // [ flags: PO_SYM_SLOTS rows x 64 ranks x u64 | bump-allocated staging regions ... ]
  // Offsets are identical on all ranks (plans are created collectively, in the same order).
  char* sym = nullptr;
  size_t sym_bytes = 0, sym_used = 0;
  int next_slot = 0;
  std::vector<char*> peer_sym;  // peer_sym[r] = rank r's buffer in MY address space (peer_sym[rank] == sym)
  char** d_peer_sym = nullptr;  // the same table in device memory (kernel argument of the peer-store exchange)
  long long timeout_ns = 600LL * 1000000000LL;  // PO_PEER_TIMEOUT_MS; <= 0: wait forever
  bool single_process = false;  // po_comm_init_all: every rank is a context of THIS process
  bool peer_is_ipc = true;      // peer_sym[r != rank] came from cudaIpcOpenMemHandle (else: plain peer access)
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
  PO_SYM(AllGather, "ncclAllGather")
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

struct po_repart_info;
static void po_make_box(const po_repart_info& info, const std::vector<i64>& local, const po_box_rec& b, i64 batch,
                        bool local_is_src, po_box& box, i64& base_off);

// ---- NVLink peer-store repartition -------------------------------------------------------
// One kernel per repartition instead of P pack launches + a grouped NCCL send/recv + P unpack
// launches: every rank packs its send boxes STRAIGHT INTO the receivers' staging regions
// (symmetric buffer, IPC-mapped peer memory, NVLink stores), publishes a per-(step, source)
// epoch flag with a system-scope release once all of its blocks have fenced, then waits with
// system-scope acquires for its own sources and unpacks.  The message bytes are tiny after
// truncation (C4: 5 MB / GPU), so the exchange is latency bound and the launch count is what
// matters.  All CTAs of the launch are co-resident (grid <= #SMs) and the two phases of a rank
// never wait on its own later work.
//
// Back-pressure.  The staging region is double-buffered by epoch parity, so the pack of epoch E
// overwrites what the receiver unpacked at epoch E-2.  Every receiver therefore ACKNOWLEDGES each
// source once all of its blocks are done with that source's box (ack flag in the SENDER's symmetric
// buffer = the consumed epoch), and a sender waits for ack >= E-2 before it packs epoch E.  This
// holds for one-directional patterns too (dims_in=[2,3] -> dims_out=[3,2]: rank 0 sends to rank 2
// but never receives from it), where "the receiver also sends to me" gave no ordering.
//
// Waits are bounded by TIME (PO_PEER_TIMEOUT_MS, default 10 minutes like torch's NCCL watchdog;
// <= 0: wait forever like NCCL).  Expiry is LOUD: the kernel poisons the boxes it could not receive
// with NaN bit patterns, does not publish or acknowledge anything, raises the context's host-mapped
// error word -- and every later po_plan_apply / po_plan_pullback on the context fails with PO_ERR_CUDA
// until the host reads the word with po_ctx_device_status (which clears it).
#define PO_PEER_MAXD 6
#define PO_PEER_MAX_BOXES 4096
struct po_peer_box {
  int nd;
  int kind;  // 0: pack into a remote staging region, 1: self overlap (x -> y), 2: unpack from my staging region
  int peer;  // kind 0: receiver, kind 2: source
  int esz;   // bytes per copied element: the tensor's element size, or a wider power of two (<= 16) when every run, stride and offset allows
  int ext[PO_PEER_MAXD];
  i64 ss[PO_PEER_MAXD], ds[PO_PEER_MAXD];  // element strides
  i64 count;
  i64 src_off, dst_off;  // BYTE offsets: kind 0: x / receiver's staging region, 1: x / y, 2: my staging region / y
};
struct po_peer_args {
  const po_peer_box* boxes;  // n_send (incl. the self copy) then n_recv, device memory owned by the plan
  int n_send, n_recv;
  unsigned long long epoch;
  int* sync;                 // [0] pack-phase block counter, [1] unpack-phase block counter, [2] sticky abort
  int* err;                  // host-mapped error word of the context
  const char* x;
  char* y;
  char* const* peer_sym;     // device table: rank r's symmetric buffer in my address space
  size_t stage_off;          // staging region of this step and epoch parity inside every symmetric buffer
  int slot, me;
  long long timeout_ns;
  long long* dbg;            // PO_PEER_TRACE: 8 globaltimer stamps of block 0 (host-mapped), else null
};

PO_D void po_st_release_sys(unsigned long long* p, unsigned long long v) {
#ifdef PO_EMUL
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
PO_D unsigned long long po_ld_acquire_sys(const unsigned long long* p) {
#ifdef PO_EMUL
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
PO_D long long po_now_ns() {
#ifdef PO_EMUL
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
#else
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}
// true: *f reached `want`; false: timeout_ns (> 0) expired first
PO_D bool po_flag_wait(const unsigned long long* f, unsigned long long want, long long timeout_ns) {
  if (po_ld_acquire_sys(f) >= want) return true;
  const long long t0 = po_now_ns();
  while (po_ld_acquire_sys(f) < want) {
#ifdef PO_EMUL
    std::this_thread::yield();
#else
    __nanosleep(128);
#endif
    if (timeout_ns > 0 && po_now_ns() - t0 > timeout_ns) return false;
  }
  return true;
}
PO_D unsigned long long* po_peer_flag(char* sym, size_t base, int slot, int who) {
  return (unsigned long long*)(sym + base) + (size_t)slot * 64 + (size_t)who;
}

// merge dims that are contiguous on both sides, drop unit dims (host): after truncation a repartition
// box is usually a 1-d or 2-d copy, so the per-element index arithmetic disappears
static void po_peer_collapse(po_peer_box& b) {
  int nd = 0;
  int ext[PO_PEER_MAXD];
  i64 ss[PO_PEER_MAXD], ds[PO_PEER_MAXD];
  for (int d = 0; d < b.nd; ++d) {
    if (b.ext[d] == 1) continue;
    if (nd > 0 && b.ss[d] == ss[nd - 1] * ext[nd - 1] && b.ds[d] == ds[nd - 1] * ext[nd - 1] &&
        (i64)ext[nd - 1] * b.ext[d] < 0x7fffffffLL) {
      ext[nd - 1] *= b.ext[d];
    } else {
      ext[nd] = b.ext[d]; ss[nd] = b.ss[d]; ds[nd] = b.ds[d]; ++nd;
    }
  }
  if (nd == 0) { ext[0] = 1; ss[0] = 1; ds[0] = 1; nd = 1; }
  b.nd = nd;
  for (int d = 0; d < nd; ++d) { b.ext[d] = ext[d]; b.ss[d] = ss[d]; b.ds[d] = ds[d]; }
}

// Copy with the widest element (<= 16 bytes) the box allows: innermost runs contiguous on both sides and every extent / stride /
// offset a multiple of the wider element (host).  The tensor bases are checked for alignment at launch time.
static void po_peer_widen(po_peer_box& b, int esz) {
  b.esz = esz;
  if (b.nd < 1 || b.ss[0] != 1 || b.ds[0] != 1) return;
  for (int w = 16 / esz; w > 1; w >>= 1) {
    bool ok = (b.ext[0] % w) == 0 && (b.src_off % ((i64)w * esz)) == 0 && (b.dst_off % ((i64)w * esz)) == 0;
    for (int d = 1; d < b.nd && ok; ++d) ok = (b.ss[d] % w) == 0 && (b.ds[d] % w) == 0;
    if (!ok) continue;
    b.ext[0] /= w;
    for (int d = 1; d < b.nd; ++d) { b.ss[d] /= w; b.ds[d] /= w; }
    b.count /= w;
    b.esz = esz * w;
    return;
  }
}

// POISON: write all-ones (a NaN for every floating type) instead of the source.  The box is shared by `nparts` CTAs; this one is `part`.
template <class E, bool POISON>
PO_D void po_peer_copy(const po_peer_box& b, const E* __restrict__ src, E* __restrict__ dst, unsigned part, unsigned nparts) {
  E ones;
  memset(&ones, 0xff, sizeof(E));
  if (b.count < 0x3fffffffLL) {  // 32-bit index arithmetic (64-bit div/mod costs ~100 instructions each)
    // four independent elements per trip: the loads are issued back to back (one element per trip measured 0.4 us PER ELEMENT --
    // every trip waited out an L2 / NVLink round trip -- and made the pack phase 15 of the exchange's 35-48 us)
    const unsigned n = (unsigned)b.count, stride = nparts * blockDim.x;
    for (unsigned idx0 = part * blockDim.x + threadIdx.x; idx0 < n; idx0 += 4 * stride) {
      E v[4];
      i64 dof[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned idx = idx0 + u * stride;
        dof[u] = -1;
        if (idx < n) {
          unsigned r = idx;
          i64 so = 0, d_ = 0;
          for (int d = 0; d < b.nd - 1; ++d) {
            const unsigned q = r / (unsigned)b.ext[d], c = r - q * (unsigned)b.ext[d];
            r = q;
            so += c * b.ss[d];
            d_ += c * b.ds[d];
          }
          so += r * b.ss[b.nd - 1];
          d_ += r * b.ds[b.nd - 1];
          dof[u] = d_;
          if (!POISON) v[u] = src[so];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dof[u] >= 0) dst[dof[u]] = POISON ? ones : v[u];
    }
    return;
  }
  for (i64 idx = (i64)part * blockDim.x + threadIdx.x; idx < b.count; idx += (i64)nparts * blockDim.x) {
    i64 r = idx, so = 0, dof = 0;
    for (int d = 0; d < b.nd; ++d) {
      const i64 c = r % b.ext[d];
      r /= b.ext[d];
      so += c * b.ss[d];
      dof += c * b.ds[d];
    }
    dst[dof] = POISON ? ones : src[so];
  }
}
template <bool POISON>
PO_D void po_peer_copy_any(const po_peer_box& b, const char* src, char* dst, unsigned part, unsigned nparts) {
  switch (b.esz) {
    case 16: po_peer_copy<po_u128, POISON>(b, (const po_u128*)src, (po_u128*)dst, part, nparts); break;
    case 8: po_peer_copy<unsigned long long, POISON>(b, (const unsigned long long*)src, (unsigned long long*)dst, part, nparts); break;
    default: po_peer_copy<unsigned, POISON>(b, (const unsigned*)src, (unsigned*)dst, part, nparts); break;
  }
}

// The exchange is latency-bound (C4 at 8 GPUs: 7 + 7 boxes of 0.65 MB per rank), so the boxes are NOT walked one after the
// other by the whole grid (that costs one flag round trip + two block barriers per box on every CTA: measured 47 us per
// exchange): the CTAs are dealt out over the boxes (CTA c -> box c mod n, slice c div n), each CTA waits for exactly ONE
// acknowledgement before it packs and for exactly ONE source flag before it unpacks.
__global__ void __launch_bounds__(256)
po_repart_peer_kernel(const PO_GRID_CONSTANT po_peer_args a) {
  __shared__ int s_last, s_abort, s_ok;
  __shared__ po_peer_box sb;
  char* const mine = a.peer_sym[a.me];
  po_pdl_trigger();
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  auto raise = [&](int code) {  // loud: host-mapped error word + sticky plan-level abort
    s_abort = 1;
    atomicExch(a.sync + 2, 1);
    *(volatile int*)a.err = code;
    __threadfence_system();
  };
  auto load_box = [&](int k) {
    __syncthreads();  // previous box is no longer in use
    if (threadIdx.x < sizeof(po_peer_box) / sizeof(int)) ((int*)&sb)[threadIdx.x] = ((const int*)(a.boxes + k))[threadIdx.x];
    __syncthreads();
  };
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const bool dbg = a.dbg && c == 0 && threadIdx.x == 0;
  if (dbg) a.dbg[0] = po_now_ns();
  po_pdl_wait();  // x may still be written by the previous kernel of the chain
  if (dbg) a.dbg[1] = po_now_ns();
  // phase 1: pack my send boxes into the receivers' staging regions (and the self overlap into y)
  for (int k = c % (a.n_send > 0 ? a.n_send : 1); k < a.n_send; k += G) {
    const unsigned part = G >= a.n_send ? (unsigned)(c / a.n_send) : 0u, nparts = G >= a.n_send ? (unsigned)((G - k + a.n_send - 1) / a.n_send) : 1u;
    load_box(k);
    if (sb.kind == 0 && a.epoch > 2) {
      // the parity region epoch E packs into was last used at E - 2: the receiver must have consumed that
      if (threadIdx.x == 0 && !po_flag_wait(po_peer_flag(mine, PO_SYM_ACK_OFF, a.slot, sb.peer), a.epoch - 2, a.timeout_ns)) raise(0x100);
      __syncthreads();
    }
    if (s_abort) break;
    po_peer_copy_any<false>(sb, a.x + sb.src_off, (sb.kind == 0 ? a.peer_sym[sb.peer] + a.stage_off : a.y) + sb.dst_off, part, nparts);
    if (G >= a.n_send) break;  // one box per CTA
  }
  if (dbg) a.dbg[2] = po_now_ns();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative over the block barrier: every store of this CTA is performed system-wide before the count
    if (dbg) a.dbg[3] = po_now_ns();
    s_last = (atomicAdd(a.sync, 1) == G - 1);
  }
  __syncthreads();
  if (s_last) {  // every block of this rank has fenced its stores: publish (never after an abort: stale data must not look fresh)
    __threadfence_system();
    const bool ok = *(volatile int*)(a.sync + 2) == 0;
    if (ok)
      for (int k = threadIdx.x; k < a.n_send; k += blockDim.x)
        if (a.boxes[k].kind == 0) po_st_release_sys(po_peer_flag(a.peer_sym[a.boxes[k].peer], 0, a.slot, a.me), a.epoch);
    if (threadIdx.x == 0) a.sync[0] = 0;
  }
  // phase 2: wait for my box's source, unpack (or poison the box when the source never arrived)
  for (int k = c % (a.n_recv > 0 ? a.n_recv : 1); k < a.n_recv; k += G) {
    const unsigned part = G >= a.n_recv ? (unsigned)(c / a.n_recv) : 0u, nparts = G >= a.n_recv ? (unsigned)((G - k + a.n_recv - 1) / a.n_recv) : 1u;
    load_box(a.n_send + k);
    if (threadIdx.x == 0) {
      s_ok = !s_abort && po_flag_wait(po_peer_flag(mine, 0, a.slot, sb.peer), a.epoch, a.timeout_ns);
      if (!s_ok && !s_abort) raise(0x200);
      if (dbg) a.dbg[4] = po_now_ns();
    }
    __syncthreads();
    if (s_ok) po_peer_copy_any<false>(sb, mine + a.stage_off + sb.src_off, a.y + sb.dst_off, part, nparts);
    else po_peer_copy_any<true>(sb, mine + a.stage_off + sb.src_off, a.y + sb.dst_off, part, nparts);
    if (G >= a.n_recv) break;
  }
  // phase 3: all blocks are done reading the staging region -> acknowledge every source
  if (dbg) a.dbg[5] = po_now_ns();
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.sync + 1, 1) == G - 1);
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    const bool ok = *(volatile int*)(a.sync + 2) == 0;
    if (ok)
      for (int k = a.n_send + threadIdx.x; k < a.n_send + a.n_recv; k += blockDim.x)
        po_st_release_sys(po_peer_flag(a.peer_sym[a.boxes[k].peer], PO_SYM_ACK_OFF, a.slot, a.me), a.epoch);
    if (threadIdx.x == 0) a.sync[1] = 0;
  }
  if (dbg) a.dbg[6] = po_now_ns();
}

static void po_comm_teardown_peer(po_comm_state* cs) {
#ifndef PO_EMUL
  for (size_t r = 0; r < cs->peer_sym.size(); ++r)
    if (cs->peer_is_ipc && (int)r != cs->rank && cs->peer_sym[r]) cudaIpcCloseMemHandle(cs->peer_sym[r]);
  if (cs->sym) cudaFree(cs->sym);
  if (cs->d_peer_sym) cudaFree(cs->d_peer_sym);
#endif
  cs->peer_sym.clear();
  cs->sym = nullptr;
  cs->d_peer_sym = nullptr;
}

// Collective: allocate the symmetric buffer, exchange IPC handles with ncclAllGather, map the peers.
// Any failure on any rank disables the peer path on ALL ranks (agreed with an all-reduce).
static int po_comm_setup_peer(po_ctx* ctx) {
#ifdef PO_EMUL
  (void)ctx;
  return PO_OK;
#else
  po_comm_state* cs = ctx->comm;
  if (!cs || !cs->nccl_comm || cs->n_ranks < 2 || cs->n_ranks > 64) return PO_OK;
  const char* off = getenv("PO_NO_PEER");
  if (off && atoi(off)) return PO_OK;
  const char* mb = getenv("PO_SYM_MB");
  const size_t S = (size_t)(mb ? atoi(mb) : 256) << 20;
  const int P = cs->n_ranks;
  float ok = 1.f;
  char* sym = nullptr;
  if (cudaMalloc((void**)&sym, S) != cudaSuccess) { ok = 0.f; sym = nullptr; cudaGetLastError(); }
  if (sym && cudaMemset(sym, 0, S) != cudaSuccess) ok = 0.f;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (sym && cudaIpcGetMemHandle(&h, sym) != cudaSuccess) { ok = 0.f; cudaGetLastError(); }
  char* dbuf = nullptr;
  PO_CUDA_CHECK(cudaMalloc((void**)&dbuf, (size_t)P * 64 + 16));
  PO_CUDA_CHECK(cudaMemcpy(dbuf + (size_t)cs->rank * 64, &h, 64, cudaMemcpyHostToDevice));
  PO_NCCL_CHECK(g_nccl.AllGather(dbuf + (size_t)cs->rank * 64, dbuf, 64, PO_NCCL_CHAR, cs->nccl_comm, (cudaStream_t)0));
  PO_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)0));
  std::vector<cudaIpcMemHandle_t> hs((size_t)P);
  PO_CUDA_CHECK(cudaMemcpy(hs.data(), dbuf, (size_t)P * 64, cudaMemcpyDeviceToHost));
  std::vector<char*> peers((size_t)P, nullptr);
  if (ok > 0.f) {
    peers[(size_t)cs->rank] = sym;
    for (int r = 0; r < P && ok > 0.f; ++r) {
      if (r == cs->rank) continue;
      void* pp = nullptr;
      if (cudaIpcOpenMemHandle(&pp, hs[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.f; cudaGetLastError(); }
      peers[(size_t)r] = (char*)pp;
    }
  }
  float* dok = (float*)(dbuf + (size_t)P * 64);
  PO_CUDA_CHECK(cudaMemcpy(dok, &ok, 4, cudaMemcpyHostToDevice));
  PO_NCCL_CHECK(g_nccl.AllReduce(dok, dok, 1, PO_NCCL_F32, 3 /* ncclMin */, cs->nccl_comm, (cudaStream_t)0));
  PO_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)0));
  PO_CUDA_CHECK(cudaMemcpy(&ok, dok, 4, cudaMemcpyDeviceToHost));
  cudaFree(dbuf);
  cs->sym = sym;
  cs->peer_sym = peers;
  if (ok <= 0.f) { po_comm_teardown_peer(cs); return PO_OK; }
  cs->sym_bytes = S;
  cs->sym_used = PO_SYM_FLAG_BYTES;
  PO_CUDA_CHECK(cudaMalloc((void**)&cs->d_peer_sym, (size_t)P * sizeof(char*)));
  PO_CUDA_CHECK(cudaMemcpy(cs->d_peer_sym, peers.data(), (size_t)P * sizeof(char*), cudaMemcpyHostToDevice));
  if (const char* t = getenv("PO_PEER_TIMEOUT_MS")) cs->timeout_ns = (long long)atoll(t) * 1000000LL;
  return PO_OK;
#endif
}

// reserve a symmetric staging region + flag row for one repartition step (same result on every rank)
static void po_repart_try_peer(po_plan* pl, po_repart_info& info, size_t esz, i64 batch) {
  po_comm_state* cm = pl->ctx->comm;
  info.peer = false;
  if (!cm || !cm->sym || cm->n_ranks != info.P || info.N + 1 > PO_PEER_MAXD) return;
  if (pl->flags & PO_PLAN_NO_PEER) return;
  // the same decision on every rank: sizes of ALL ranks' plans
  size_t max_recv = 0;
  size_t max_boxes = 0;
  for (int q = 0; q < info.P; ++q) {
    po_repart_info other;
    std::vector<int32_t> din(info.dims_in.begin(), info.dims_in.end()), dout(info.dims_out.begin(), info.dims_out.end());
    std::vector<int64_t> g(info.global.begin(), info.global.end());
    if (po_repart_build(info.N, g.data(), din.data(), dout.data(), q, other)) return;
    i64 tot = 0;
    for (auto& b : other.recv) if (b.peer != q) tot += b.count() * batch;
    if ((size_t)tot * esz > max_recv) max_recv = (size_t)tot * esz;
    if (other.send.size() + other.recv.size() > max_boxes) max_boxes = other.send.size() + other.recv.size();
    if (q == info.rank) continue;
    // where does MY box land inside rank q's staging region?
    i64 off = 0;
    for (auto& b : other.recv) {
      if (b.peer == info.rank) break;
      if (b.peer != q) off += b.count() * batch;
    }
    for (size_t k = 0; k < info.send.size(); ++k)
      if (info.send[k].peer == q) { if (info.send_peer_off.size() != info.send.size()) info.send_peer_off.assign(info.send.size(), 0); info.send_peer_off[k] = off; }
  }
  if (info.send_peer_off.size() != info.send.size()) info.send_peer_off.assign(info.send.size(), 0);
  if (max_boxes > PO_PEER_MAX_BOXES) return;
  const size_t region = (max_recv + 255) & ~(size_t)255;
  if (cm->next_slot >= PO_SYM_SLOTS || cm->sym_used + 2 * region > cm->sym_bytes) return;
  info.slot = cm->next_slot++;
  info.region_off = cm->sym_used;
  info.region_bytes = region;
  cm->sym_used += 2 * region;
  // box descriptors (byte offsets relative to x / y / the staging region: the same table serves every call)
  std::vector<po_peer_box> boxes;
  auto fill = [&](po_peer_box& pb, const std::vector<i64>& local, const po_box_rec& b, bool local_is_src, const std::vector<i64>* local2, const po_box_rec* b2) {
    po_box box;
    i64 base;
    po_make_box(info, local, b, batch, local_is_src, box, base);
    memset(&pb, 0, sizeof(pb));
    pb.nd = box.nd;
    pb.count = 1;
    for (int d = 0; d < box.nd; ++d) { pb.ext[d] = (int)box.ext[d]; pb.ss[d] = box.sstride[d]; pb.ds[d] = box.dstride[d]; pb.count *= box.ext[d]; }
    i64 base2 = 0;
    if (local2) {
      po_box box2;
      po_make_box(info, *local2, *b2, batch, false, box2, base2);
      for (int d = 0; d < box.nd; ++d) pb.ds[d] = box2.dstride[d];
    }
    po_peer_collapse(pb);
    pb.esz = (int)esz;
    return std::make_pair(base, base2);
  };
  for (size_t k = 0; k < info.send.size(); ++k) {
    const po_box_rec& b = info.send[k];
    if (b.count() == 0) continue;
    po_peer_box pb;
    if (b.peer != info.rank) {
      auto off = fill(pb, info.local_in, b, true, nullptr, nullptr);
      pb.kind = 0; pb.peer = b.peer;
      pb.src_off = off.first * (i64)esz;
      pb.dst_off = info.send_peer_off[k] * (i64)esz;
    } else {
      const po_box_rec* rb = nullptr;
      for (auto& r : info.recv) if (r.peer == info.rank) rb = &r;
      if (!rb) continue;
      auto off = fill(pb, info.local_in, b, true, &info.local_out, rb);
      pb.kind = 1; pb.peer = info.rank;
      pb.src_off = off.first * (i64)esz;
      pb.dst_off = off.second * (i64)esz;
    }
    boxes.push_back(pb);
  }
  info.n_send_desc = (int)boxes.size();
  for (size_t k = 0; k < info.recv.size(); ++k) {
    const po_box_rec& b = info.recv[k];
    if (b.peer == info.rank || b.count() == 0) continue;
    po_peer_box pb;
    auto off = fill(pb, info.local_out, b, false, nullptr, nullptr);
    pb.kind = 2; pb.peer = b.peer;
    pb.src_off = info.recv_off[k] * (i64)esz;
    pb.dst_off = off.first * (i64)esz;
    boxes.push_back(pb);
  }
  info.n_recv_desc = (int)boxes.size() - info.n_send_desc;
  info.peer_work = 0;
  for (auto& pb : boxes) info.peer_work += pb.count;
  // second table: the same boxes with the widest element each allows (used when x and y are 16-byte aligned)
  const size_t nb = boxes.size();
  for (size_t k = 0; k < nb; ++k) { po_peer_box w = boxes[k]; po_peer_widen(w, (int)esz); boxes.push_back(w); }
#ifndef PO_EMUL
  if (cudaMalloc((void**)&info.d_done, 4 * sizeof(int)) != cudaSuccess) return;
  cudaMemset(info.d_done, 0, 4 * sizeof(int));
  pl->owned.push_back(info.d_done);
  if (!boxes.empty()) {
    if (cudaMalloc(&info.d_desc, boxes.size() * sizeof(po_peer_box)) != cudaSuccess) return;
    pl->owned.push_back(info.d_desc);
    if (cudaMemcpy(info.d_desc, boxes.data(), boxes.size() * sizeof(po_peer_box), cudaMemcpyHostToDevice) != cudaSuccess) return;
  }
#endif
  info.peer = true;
}

static int po_repart_run_peer(po_plan* pl, po_step& s, po_repart_info& info, const void* src, void* dst, cudaStream_t st) {
  po_comm_state* cm = pl->ctx->comm;
  const size_t esz = po_dtype_size(s.in_dt);
  info.epoch += 1;
  po_peer_args a;
  memset(&a, 0, sizeof(a));
  const bool aligned = (((size_t)src | (size_t)dst) % 16) == 0;
  a.boxes = (const po_peer_box*)info.d_desc + (aligned ? (size_t)(info.n_send_desc + info.n_recv_desc) : 0);
  a.n_send = info.n_send_desc; a.n_recv = info.n_recv_desc;
  a.epoch = (unsigned long long)info.epoch;
  a.sync = info.d_done;
  a.err = pl->ctx->d_err;
  a.x = (const char*)src; a.y = (char*)dst;
  a.peer_sym = cm->d_peer_sym;
  a.stage_off = info.region_off + (size_t)(info.epoch & 1) * info.region_bytes;
  a.slot = info.slot; a.me = info.rank;
  a.timeout_ns = cm->timeout_ns;
  if (a.n_send + a.n_recv == 0) return PO_OK;
#ifndef PO_EMUL
  static const bool peer_trace = getenv("PO_PEER_TRACE") != nullptr;
  if (peer_trace) {
    if (!info.h_dbg) {
      cudaHostAlloc((void**)&info.h_dbg, 8 * sizeof(long long), cudaHostAllocMapped);
      memset(info.h_dbg, 0, 8 * sizeof(long long));
      cudaHostGetDevicePointer((void**)&info.d_dbg, info.h_dbg, 0);
    } else if (info.rank == 0 && info.h_dbg[6] && (info.epoch % 16) == 0) {  // stamps of an EARLIER launch (whatever has landed)
      const long long* t = info.h_dbg;
      fprintf(stderr, "[parops] peer exchange slot %d (block 0, us): pdl-wait %.1f | pack %.1f | fence %.1f | flag seen +%.1f | unpack %.1f | ack %.1f | total %.1f\n", info.slot,
              (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - t[3]) / 1e3, (t[5] - t[4]) / 1e3, (t[6] - t[5]) / 1e3, (t[6] - t[0]) / 1e3);
    }
    a.dbg = info.d_dbg;
  }
#endif
  int grid = (int)((info.peer_work + 256 * 8 - 1) / (256 * 8));
  // all CTAs must be co-resident (they wait on one another through the block counters): 256 threads, no shared memory -> two per SM
  // is far inside the residency limit, and halves the trips of the latency-bound copy loops
  const int cap = pl->ctx->num_sms > 0 ? 2 * pl->ctx->num_sms : 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (esz != 4 && esz != 8 && esz != 16) return po_fail(PO_ERR_DTYPE, "repartition: bad element size");
  PO_LAUNCH_PDL(po_repart_peer_kernel, dim3(grid), dim3(256), 0, st, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "peer repartition launch failed: %s", cudaGetErrorString(e));
  pl->launches += 1;
  return PO_OK;
}

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
  po_repart_try_peer(pl, *info, esz, batch);
  s.repart = info;  // freed in po_plan_destroy; device memory through the plan's owned list
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
  if (info.peer) return po_repart_run_peer(pl, s, info, src, dst, st);
  if (remote && cm->single_process)
    return po_fail(PO_ERR_UNSUPPORTED, "ParRepartition on a po_comm_init_all communicator needs the NVLink peer-store path (symmetric buffer too small or PO_PLAN_NO_PEER set): "
                   "a grouped NCCL exchange issued rank after rank from one thread would deadlock");

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
