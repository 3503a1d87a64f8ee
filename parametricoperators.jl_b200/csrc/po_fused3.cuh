// Third-generation FORWARD fused (t, x) kernel: asynchronous row ring.
//
// Measured on B200 (scripts/microbench/mb_stream.cu, profiles/r1_microbench_stream.txt): the forward kernel's access pattern
// (per unit 8*RT rows of I0*nx floats, 21 MB apart) streams at 6.5 TB/s from register-staged LDG bursts when
// nothing else happens, and at 7.0 TB/s through a 3-stage shared-memory ring filled by cp.async.bulk -- but the
// generation-1/2 kernels reach only 4.7-5.0 TB/s because every streaming warp alternates "32 loads in flight"
// with "0 loads in flight while it computes" (ncu: 62 % long-scoreboard stalls, issue slots 31 % busy).  Here the
// loads are decoupled from the arithmetic:
//   * one producer thread keeps D = 3 stages x 8 rows (120 KB) of bulk asynchronous copies (TMA engine,
//     completion on mbarriers) in flight at ALL times;
//   * 10 transform warps wait for a stage, run one step of the pruned t-decimation on it straight from shared
//     memory (LDS.64 with immediate offsets, compile-time twiddles), accumulate the 8 kept t-modes of their two
//     element-pair columns in registers and release the stage;
//   * after the RT-th stage of a unit they hand the 8x-reduced tile to 5 x-phase warps through a single
//     shared-memory tile guarded by a full/free mbarrier pair; the packed x-phase of po_fused2.cuh finishes the
//     unit while the transform warps are already consuming the next unit's stages.
// No __syncthreads() after start-up.  Needs whole rows with I0*nx <= 1280 (ring + tile <= 227 KB) and an even I0.
#pragma once
#include "po_fused2.cuh"

#define PO_RING_STAGES 3
#define PO_RING_TWARPS 10
#define PO_RING_XWARPS 5
#define PO_RING_THREADS 512  // 10 transform + 5 x-phase + 1 producer warp

// ---- mbarrier primitives (GPU: PTX; emulator: two counters and a spin) ---------------------------------------
#ifdef PO_EMUL
#include <thread>
struct po_mbar_t { int pending; unsigned phase; int count; int pad; };
PO_D void po_mb_init(po_mbar_t* b, int count) { b->pending = count; b->phase = 0; b->count = count; }
PO_D void po_mb_fence_init() {}
PO_D void po_mb_arrive(po_mbar_t* b) {
  if (__atomic_sub_fetch(&b->pending, 1, __ATOMIC_ACQ_REL) == 0) {
    __atomic_store_n(&b->pending, b->count, __ATOMIC_RELAXED);
    __atomic_fetch_add(&b->phase, 1u, __ATOMIC_RELEASE);
  }
}
PO_D bool po_mb_test(po_mbar_t* b, unsigned parity) { return (__atomic_load_n(&b->phase, __ATOMIC_ACQUIRE) & 1u) != parity; }
PO_D bool po_mb_test_suspend(po_mbar_t* b, unsigned parity) { return po_mb_test(b, parity); }
PO_D void po_mb_backoff() { std::this_thread::yield(); }
// fill: copy the rows, then arrive (the emulator has no transaction counts)
PO_D void po_mb_fill_begin(po_mbar_t*, unsigned) {}
PO_D void po_mb_fill_row(void* dst, const void* src, unsigned bytes, po_mbar_t*, unsigned long long) { memcpy(dst, src, bytes); }
PO_D void po_mb_fill_end(po_mbar_t* b) { po_mb_arrive(b); }
#else
typedef unsigned long long po_mbar_t;
PO_D void po_mb_init(po_mbar_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(po_smem_u32(b)), "r"(count) : "memory"); }
PO_D void po_mb_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PO_D void po_mb_arrive(po_mbar_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(po_smem_u32(b)) : "memory"); }
PO_D bool po_mb_test(po_mbar_t* b, unsigned parity) { return po_mbar_try_wait(b, parity); }
// blocking flavour: the hardware suspends the thread until the phase completes or the hint (ns) expires, so a
// waiting warp stops competing for issue slots and shared-memory bandwidth (ncu, r1y: 31 % of the ring kernel's
// executed instructions were wait-loop iterations of the transform warps, next to the x-phase warps that bound it)
PO_D bool po_mb_test_suspend(po_mbar_t* b, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(po_smem_u32(b)), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
PO_D void po_mb_backoff() {}
PO_D void po_mb_fill_begin(po_mbar_t* b, unsigned bytes) { po_mbar_expect_tx(b, bytes); }
PO_D void po_mb_fill_row(void* dst, const void* src, unsigned bytes, po_mbar_t* b, unsigned long long pol) {
  if (pol)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(po_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(po_smem_u32(b)), "l"(pol)
                 : "memory");
  else
    po_bulk_load(dst, src, bytes, b);
}
PO_D void po_mb_fill_end(po_mbar_t*) {}
#endif

// bounded wait: a protocol bug must not hang the GPU box.  On expiry the CTA-wide abort flag is raised, every
// later wait returns at once and the context's error word is set (po_ctx_check reports it).
PO_D void po_mb_wait(po_mbar_t* b, unsigned parity, volatile int* aborted, int* err) {
  if (po_mb_test(b, parity)) return;  // fast path: no bookkeeping
  int spins = 0;
  while (!po_mb_test_suspend(b, parity)) {
    po_mb_backoff();
    if ((++spins & 63) == 0) {
      if (*aborted) return;
#ifdef PO_EMUL
      if (spins > (1 << 28)) {  // emulator: iterations are sched_yield()s of an oversubscribed host, allow minutes
#else
      if (spins > (1 << 20)) {  // each iteration may suspend for up to 20 us: this is seconds
#endif
        *aborted = 1; *err = 3;
        return;
      }
    }
  }
}

template <int RT, int J2>
PO_D void po_ring_tstep(const float2* __restrict__ st, int L2, int v, float2 (&ar)[8], float2 (&ai)[8]) {
  float2 x[8];
#pragma unroll
  for (int j1 = 0; j1 < 8; ++j1) x[j1] = st[j1 * L2 + v];
  po_fwd_tstep<RT, J2>(x, ar, ai);
}

struct po_ring_state { int s; unsigned ph; };
PO_D void po_ring_next(po_ring_state& r) { if (++r.s == PO_RING_STAGES) { r.s = 0; r.ph ^= 1u; } }

// one stage = step J2 of the current unit for the thread's (up to) two pair-columns
template <int RT, int J2>
PO_D void po_ring_consume(const float* __restrict__ ring, po_mbar_t* full, po_mbar_t* empty, po_ring_state& rs, int L, int v0, int v1, bool has0,
                          bool has1, float2 (&ar0)[8], float2 (&ai0)[8], float2 (&ar1)[8], float2 (&ai1)[8], volatile int* aborted, int* err) {
  po_mb_wait(&full[rs.s], rs.ph, aborted, err);
  const float2* st = (const float2*)(ring + (size_t)rs.s * 8 * L);
  if (has0) po_ring_tstep<RT, J2>(st, L >> 1, v0, ar0, ai0);
  if (has1) po_ring_tstep<RT, J2>(st, L >> 1, v1, ar1, ai1);
  PO_SYNCWARP();
  if ((threadIdx.x & 31) == 0) po_mb_arrive(&empty[rs.s]);
  po_ring_next(rs);
}

template <int RT, int RX, int HP>
__global__ void __launch_bounds__(PO_RING_THREADS, 1)
po_fwd_tx_ring_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units,
                      int* __restrict__ err, int l2hint, int xflags) {
  const int share_x = xflags & 1, xshfl = xflags & 2;  // bit 0: transform warps help with the x-phase; bit 1: shuffle combine
  constexpr int D = PO_RING_STAGES, NT = PO_RING_TWARPS * 32, NX = PO_RING_XWARPS * 32;
  PO_DYN_SMEM(smem_raw);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX;
  float* ring = (float*)smem_raw;                                 // [D][8][L]
  float* T = ring + (size_t)D * 8 * L;                            // [8][2L], pair-planar
  float4* tw4 = (float4*)(T + 16 * (size_t)L);                    // [RX][17]
  float* ksc = (float*)(tw4 + RX * PO_TW4_LD);                    // [8]
  po_mbar_t* full = (po_mbar_t*)(ksc + 8);                        // [D]
  po_mbar_t* empty = full + D;                                    // [D]
  po_mbar_t* tfull = empty + D;
  po_mbar_t* tfree = tfull + 1;
  volatile int* aborted = (volatile int*)(tfree + 1);
  const int tid = threadIdx.x;
  po_pdl_trigger();
  for (int e = tid; e < RX * 16; e += PO_RING_THREADS) {
    const float2 w = tw.tw_x[e];
    tw4[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x, w.x, w.y, w.y);
  }
  if (tid < 8) ksc[tid] = tw.kscale[tid];
  if (tid == 0) {
    for (int s = 0; s < D; ++s) { po_mb_init(&full[s], 1); po_mb_init(&empty[s], PO_RING_TWARPS); }
    po_mb_init(tfull, PO_RING_TWARPS);
    po_mb_init(tfree, PO_RING_XWARPS + (((I0 >> 1) * 8 * RX >= 64 && share_x) ? PO_RING_TWARPS / 2 : 0));
    *aborted = 0;
    po_mb_fence_init();
  }
  __syncthreads();
  po_pdl_wait();  // tables and barriers are set up: from here on the kernel touches tensors
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;  // units of this CTA
  const i64 rstride = (i64)L * Mid;                                  // floats between consecutive t rows
  // OPTIONAL (PO_RING_SHARE_X=1, off by default: it measured slower) x-phase work split: the 5 x-phase warps take the first half of the items; half of the
  // transform warps -- alternating between units -- take the second half right after they have filled the tile.
  // (ncu r1y: the transform warps spent 12-17 % of their time waiting for the x-phase warps to free the tile.)
  const int x_items = (I0 >> 1) * 8 * RX;
  const int x_split = (share_x && x_items >= 64) ? ((x_items / 2 + 31) & ~31) : x_items;  // x_items is a multiple of 32

  if (tid >= NT + NX) {
    // ---- producer: one thread issues every bulk copy of this CTA ----
    if (tid == NT + NX) {
      const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;  // input rows are read once: do not keep them in L2
      po_ring_state rs = {0, 0};
      for (i64 it = 0; it < nit; ++it) {
        const i64 unit = blockIdx.x + it * gridDim.x;
        const i64 m = unit % Mid, bb = unit / Mid;
        const float* ub = X + (i64)L * m + rstride * (i64)(8 * RT) * bb;
        for (int j2 = 0; j2 < RT; ++j2) {
          if (it * RT + j2 >= D) po_mb_wait(&empty[rs.s], rs.ph ^ 1u, aborted, err);  // previous use of the stage released
          float* dst = ring + (size_t)rs.s * 8 * L;
          po_mb_fill_begin(&full[rs.s], 8u * (unsigned)L * 4u);
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) po_mb_fill_row(dst + (size_t)j1 * L, ub + rstride * (i64)(RT * j1 + j2), (unsigned)L * 4u, &full[rs.s], pol);
          po_mb_fill_end(&full[rs.s]);
          po_ring_next(rs);
        }
      }
    }
  } else if (tid < NT) {
    // ---- transform warps: t-decimation out of the ring, accumulators in registers ----
    const int v0 = tid, v1 = tid + NT;
    const bool has0 = v0 < (L >> 1), has1 = v1 < (L >> 1);
    const unsigned long long zpol_t = l2hint ? po_policy_evict_last() : 0ull;
    po_ring_state rs = {0, 0};
    for (i64 it = 0; it < nit; ++it) {
      float2 ar0[8], ai0[8], ar1[8], ai1[8];
#define PO_RING_C(J2) po_ring_consume<RT, J2>(ring, full, empty, rs, L, v0, v1, has0, has1, ar0, ai0, ar1, ai1, aborted, err);
      PO_RING_C(0) PO_RING_C(1)
      if constexpr (RT > 2) { PO_RING_C(2) PO_RING_C(3) }
      if constexpr (RT > 4) { PO_RING_C(4) PO_RING_C(5) PO_RING_C(6) PO_RING_C(7) }
#undef PO_RING_C
      if (it > 0) po_mb_wait(tfree, (unsigned)((it - 1) & 1), aborted, err);  // x-phase of the previous unit is done with the tile
      if (has0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) *(float4*)(T + k * 2 * L + 4 * v0) = make_float4(ar0[k].x, ar0[k].y, ai0[k].x, ai0[k].y);
      }
      if (has1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) *(float4*)(T + k * 2 * L + 4 * v1) = make_float4(ar1[k].x, ar1[k].y, ai1[k].x, ai1[k].y);
      }
      PO_SYNCWARP();
      if ((tid & 31) == 0) po_mb_arrive(tfull);
      if (share_x && x_split < x_items) {
        const int half = (int)(it & 1);                      // which half of the transform warps helps this unit
        const int w = tid >> 5;
        if ((w >= PO_RING_TWARPS / 2) == (half != 0)) {
          const i64 unit = blockIdx.x + it * gridDim.x;
          const i64 m = unit % Mid, bb = unit / Mid;
          po_mb_wait(tfull, (unsigned)(it & 1), aborted, err);  // every transform warp has written its columns
          po_fwd_xphase2<RX>(T, tw4, ksc, Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, I0, L, tid - half * (NT / 2), NT / 2, I0, 0, zpol_t,
                             x_split, x_items);
          PO_SYNCWARP();
          if ((tid & 31) == 0) po_mb_arrive(tfree);
        }
      }
    }
  } else {
    // ---- x-phase warps ----
    const int tx = tid - NT;
    const unsigned long long zpol = l2hint ? po_policy_evict_last() : 0ull;  // the 8x-reduced output is the next kernel's input
    for (i64 it = 0; it < nit; ++it) {
      const i64 unit = blockIdx.x + it * gridDim.x;
      const i64 m = unit % Mid, bb = unit / Mid;
      po_mb_wait(tfull, (unsigned)(it & 1), aborted, err);
      if (xshfl) po_fwd_xphase2<RX, true>(T, tw4, ksc, Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, I0, L, tx, NX, I0, 0, zpol, 0, x_split);
      else po_fwd_xphase2<RX, false>(T, tw4, ksc, Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, I0, L, tx, NX, I0, 0, zpol, 0, x_split);
      PO_SYNCWARP();
      if ((tid & 31) == 0) po_mb_arrive(tfree);
    }
  }
}

static inline size_t po_ring_smem(i64 L, int RX) {
  return (size_t)(PO_RING_STAGES * 8 + 16) * L * sizeof(float) + (size_t)RX * PO_TW4_LD * sizeof(float4) + 8 * sizeof(float) +
         (2 * PO_RING_STAGES + 2) * sizeof(po_mbar_t) + 64;
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_RING)
// cudaErrorNotSupported: shape not handled here (caller falls back to generation 2 / 1)
template <int RT, int RX>
static cudaError_t po_launch_fwd_ring_t(const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                                        cudaStream_t st) {
  const i64 L = (i64)I0 * 16 * RX;
  if ((I0 & 1) || num_sms <= 0 || !err) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16)) return cudaErrorNotSupported;  // bulk copies and 128-bit stores need 16-byte aligned tensors
  if (L / 2 > 2 * PO_RING_TWARPS * 32) return cudaErrorNotSupported;  // two pair-columns per transform thread
  const size_t smem = po_ring_smem(L, RX);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  const i64 units = Mid * B;
  if (units < 4 * (i64)num_sms) return cudaErrorNotSupported;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const int share_x = (getenv("PO_RING_SHARE_X") ? (atoi(getenv("PO_RING_SHARE_X")) & 1) : 0)   // measured: 0.131 -> 0.147 ms at C3 when on (gpurun r1 share-x A/B)
                             | ((getenv("PO_RING_XSHFL") ? atoi(getenv("PO_RING_XSHFL")) : 1) ? 2 : 0);
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-3 ring fwd RT=%d RX=%d I0=%d units=%lld smem=%zu\n", RT, RX, I0, (long long)units, smem);
#ifndef PO_EMUL
  cudaError_t e;
#define PO_RING_GO(K)                                                                                   \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
  if (e != cudaSuccess) return e;                                                                       \
  PO_LAUNCH_PDL(K, dim3((unsigned)num_sms), dim3(PO_RING_THREADS), smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, err, l2hint, share_x);
#else
#define PO_RING_GO(K) PO_LAUNCH(K, dim3((unsigned)num_sms), dim3(PO_RING_THREADS), smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, err, l2hint, share_x);
#endif
  if (I0 == 20) { PO_RING_GO((po_fwd_tx_ring_kernel<RT, RX, 10>)) }
  else { PO_RING_GO((po_fwd_tx_ring_kernel<RT, RX, 0>)) }
#undef PO_RING_GO
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_fwd_ring(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                                      int* err, cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_RING_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_fwd_ring_t<rt, rx>(src, dst, tw, I0, Mid, B, num_sms, err, st);
  PO_RING_CASE(2, 4) PO_RING_CASE(4, 4) PO_RING_CASE(8, 4)
  PO_RING_CASE(2, 8) PO_RING_CASE(4, 8) PO_RING_CASE(8, 8)
#undef PO_RING_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_fwd_ring(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                               int* err, cudaStream_t st);
#endif

// =======================================================================================
// Third-generation INVERSE fused (x, t) kernel: the pipelined generation-2 kernel (po_inv_xt_pipe2_kernel: x-phase
// warps fill tile i+1 while the streaming warps run the t-phase of tile i and store) with the x-phase's INPUT staged
// through shared memory by bulk asynchronous copies.
//
// Why: the phase-skip experiments (profiles/r1_inverse_phase_skip_experiments.txt and the later L2-hint runs) put
// the store stream alone at 0.109 ms and the x-phase alone at 0.066-0.082 ms, yet together they take 0.135 ms: the
// x-phase warps' 16 register-staged LDG.128 per item queue behind the store flood of their own SM, the x-phase
// becomes the longer leg of the two-stage pipeline and the streaming warps idle at the unit barrier.  (An L2 prefetch
// one unit ahead already bought 14 %; making the intermediate fully L2-resident bought nothing more,
// profiles/r1_inverse_pair_chunks_ab.txt.)  Here one thread issues the 8 t-mode rows of unit i+1 (8 x I0*16 complex,
// 20 KB at c = 20) as cp.async.bulk copies a whole unit ahead; the x-phase reads them with LDS and never waits on
// the memory system.  Shared memory: 2 tiles (160 KB) + tables (8.5 KB) + 2 stages (40 KB) at c = 20, nx = 64.
template <int RT, int RX, int HP, int NTHR>
__global__ void __launch_bounds__(NTHR, 1)
po_inv_xt_stage_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units,
                       int nstream, int l2hint, int* __restrict__ err, i64 m_lo, i64 m_cnt) {
  PO_DYN_SMEM(smem_raw);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX, ZR = I0 * 16;  // floats per tile row / complex numbers per staged t-mode row
  float* T0 = (float*)smem_raw;                            // two [8][2L] tiles
  float4* tw4k = (float4*)(T0 + 32 * (size_t)L);           // [8][RX][17]
  float2* zs = (float2*)(tw4k + 8 * RX * PO_TW4_LD);       // [2][8][ZR] staged input rows
  po_mbar_t* full = (po_mbar_t*)(zs + 2 * 8 * (size_t)ZR); // [2]
  volatile int* aborted = (volatile int*)(full + 2);
  const int tid = threadIdx.x;
  po_pdl_trigger();
  const i64 rs2 = ((i64)L * Mid) >> 1;
  for (int e = tid; e < 8 * RX * 16; e += NTHR) {
    const int k = e / (RX * 16);
    const float2 w = tw.tw_x[e % (RX * 16)];
    const float s = tw.kscale[k];
    tw4k[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x * s, w.x * s, w.y * s, w.y * s);
  }
  if (tid == 0) {
    po_mb_init(&full[0], 1);
    po_mb_init(&full[1], 1);
    *aborted = 0;
    po_mb_fence_init();
  }
  __syncthreads();
  po_pdl_wait();
  const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;  // output rows: stream through L2
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const i64 kstride = (i64)ZR * Mid;
  // stage `it & 1` <- the 8 t-mode rows of this CTA's unit `it` (one thread; completion on full[it & 1])
  auto issue = [&](i64 it) {
    const po_unit u = po_decode_unit(blockIdx.x + it * gridDim.x, 1, I0, m_lo, m_cnt);
    const float2* zb = Z + (i64)ZR * (u.m + Mid * 8 * u.bb);
    float2* dst = zs + (size_t)(it & 1) * 8 * ZR;
    po_mbar_t* b = &full[it & 1];
    po_mb_fill_begin(b, 8u * (unsigned)ZR * 8u);
#pragma unroll
    for (int k = 0; k < 8; ++k) po_mb_fill_row(dst + (size_t)k * ZR, zb + kstride * k, (unsigned)ZR * 8u, b, 0ull);
    po_mb_fill_end(b);
  };
  if (tid == nstream && nit > 0) issue(0);
  for (i64 it = 0; it <= nit; ++it) {
    if (tid >= nstream) {
      // stage (it + 1) & 1 was consumed by the x-phase of iteration it - 1, which the barrier below has closed
      if (tid == nstream && it + 1 < nit) issue(it + 1);
      if (it < nit) {
        po_mb_wait(&full[it & 1], (unsigned)((it >> 1) & 1), aborted, err);
        po_inv_xphase2<RX, true>(zs + (size_t)(it & 1) * 8 * ZR, (i64)ZR, T0 + (it & 1) * 16 * (size_t)L, tw4k, I0, L, tid - nstream,
                                 NTHR - nstream, I0, 0);
      }
    } else if (it > 0) {
      const po_unit u = po_decode_unit(blockIdx.x + (it - 1) * gridDim.x, 1, I0, m_lo, m_cnt);
      const unsigned ubase = (unsigned)(((i64)L * u.m + (i64)L * Mid * (i64)(8 * RT) * u.bb) >> 1);
      po_inv_tphase2<RT, true>(T0 + ((it - 1) & 1) * 16 * (size_t)L, (float2*)Y, ubase, (unsigned)rs2, L, tid, nstream, I0, 0, I0, pol);
    }
    __syncthreads();
  }
}

static inline size_t po_stage_inv_smem(i64 L, int I0, int RX) {
  return (size_t)32 * L * sizeof(float) + (size_t)8 * RX * PO_TW4_LD * sizeof(float4) + (size_t)2 * 8 * I0 * 16 * sizeof(float2) + 2 * sizeof(po_mbar_t) + 64;
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_STAGE)
// cudaErrorNotSupported: shape not handled here (caller falls back to the generation-2 inverse)
template <int RT, int RX>
static cudaError_t po_launch_inv_stage_t(const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                                         cudaStream_t st, i64 m_lo, i64 m_cnt) {
  const i64 nx = 16 * RX, L = (i64)I0 * nx;
  if ((I0 & 1) || num_sms <= 0 || !err) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16)) return cudaErrorNotSupported;
  if (m_cnt >= 0 && (m_lo < 0 || m_lo + m_cnt > Mid)) return cudaErrorInvalidValue;
  if (m_cnt < 0) { m_lo = 0; m_cnt = Mid; }
  const size_t smem = po_stage_inv_smem(L, I0, RX);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  const i64 units = m_cnt * B;
  if (units < 4 * (i64)num_sms) return cudaErrorNotSupported;
  if ((i64)I0 * nx * Mid * (8 * RT) * B / 2 >= (1LL << 32)) return cudaErrorNotSupported;  // 32-bit float2 row offsets
  // warp split: with the input staged the x-phase is pure arithmetic, so it gets 4 warps and the streaming t-phase the rest
  // (measured at 512 threads, profiles/r1_staged_inverse_ab.txt: 320 streaming threads 0.139 ms, 352/384 0.134 ms, 416 0.147 ms at C3)
  static const int ns_env = getenv("PO_PIPE2_STREAM_INV") ? atoi(getenv("PO_PIPE2_STREAM_INV")) : 0;
  // CTA size: the kernel is latency-bound at 16 warps / SM, and with the staged input it needs only 94 registers, so a 640-thread
  // CTA (16 streaming + 4 x-phase warps) fits the register file.  Measured (profiles/r1_staged_inverse_ab.txt): C3 (RT = 4)
  // 0.1347 -> 0.1297 ms with 640 threads, 0.140 with 768 (80 registers, spills); at RT = 8 (96 registers) 512 stays best.
  const int nthr_env = getenv("PO_INV_STAGE_THREADS") ? atoi(getenv("PO_INV_STAGE_THREADS")) : 0;  // per launch: tests switch it
  const int nthr = (nthr_env == 512 || nthr_env == 640 || nthr_env == 768) ? nthr_env : (RT <= 4 ? 640 : 512);
  const int ns = ns_env ? ns_env : nthr - 128;
  if (ns <= 0 || ns >= nthr || (ns & 31)) return cudaErrorInvalidValue;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-3 staged inv RT=%d RX=%d I0=%d units=%lld nstream=%d smem=%zu\n", RT, RX, I0, (long long)units, ns, smem);
#ifndef PO_EMUL
  cudaError_t e;
#define PO_STG_GO(K)                                                                                    \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
  if (e != cudaSuccess) return e;                                                                       \
  PO_LAUNCH_PDL(K, dim3((unsigned)num_sms), dim3(nthr), smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, ns, l2hint, err, m_lo, m_cnt);
#else
#define PO_STG_GO(K) PO_LAUNCH_PDL(K, dim3((unsigned)num_sms), dim3(nthr), smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, ns, l2hint, err, m_lo, m_cnt);
#endif
#define PO_STG_NT(NT)                                                        \
  if (nthr == NT) {                                                          \
    if (I0 == 20) { PO_STG_GO((po_inv_xt_stage_kernel<RT, RX, 10, NT>)) }    \
    else { PO_STG_GO((po_inv_xt_stage_kernel<RT, RX, 0, NT>)) }              \
  }
  PO_STG_NT(512) PO_STG_NT(640) PO_STG_NT(768)
#undef PO_STG_NT
#undef PO_STG_GO
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_inv_stage(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                                       int* err, cudaStream_t st, i64 m_lo = 0, i64 m_cnt = -1) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_STG_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_inv_stage_t<rt, rx>(src, dst, tw, I0, Mid, B, num_sms, err, st, m_lo, m_cnt);
  PO_STG_CASE(2, 4) PO_STG_CASE(4, 4) PO_STG_CASE(8, 4)
  PO_STG_CASE(2, 8) PO_STG_CASE(4, 8) PO_STG_CASE(8, 8)
#undef PO_STG_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_inv_stage(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                                int* err, cudaStream_t st, i64 m_lo = 0, i64 m_cnt = -1);
#endif
