// Fourth generation: CTA-PAIR (thread-block cluster of 2) fused (t, x) kernels for rows too wide for one SM.
//
// The C4 ladder's local shapes from 2 GPUs on have rows of c * nx = 20 * 128 floats = 10 KB.  The asynchronous row ring of
// po_fused3.cuh needs 3 stages x 8 rows + the 8x-reduced tile = 40 row lengths of shared memory (400 KB at this width), so these
// shapes fell back to the single-tile generation-2 kernels, where the HBM pipe idles during every x-phase (forward 4.75 TB/s
// = 0.73 of the measured peak, the weak-scaling ladder's largest loss).
//
// Here TWO CTAs on the two SMs of a cluster share one unit (one (y, z, batch) column of 8*RT rows):
//   * CTA h streams the x-HALF h of every row (a contiguous 5 KB run) through ITS OWN 3-stage bulk-copy ring and reduces it
//     over t into ITS OWN half tile -- per SM exactly the byte stream, ring and tile of the 64-wide single-CTA kernel;
//   * the x-DFT of a (channel pair, t-mode) needs all of x: the x-phase work is split by t-MODE (CTA h owns modes 4h..4h+3)
//     and every lane PULLS 8 of its 16 inputs from the partner's tile through distributed shared memory (mapa + generic ld).
//     (Measured alternative, r2d: the transform warps PUSH half of their results into the partner's tile with remote stores so
//     that the x-phase reads local memory only -- slower, 0.280 vs 0.268 ms at the 8-GPU local shape: the stores sit on the
//     transform warps, which hold the ring stages, while the x-phase has slack.)
//   * tile hand-over: the transform warps fill their own tile and arrive on a CTA-local "full" barrier exactly as in the ring
//     kernel; ONE x-phase thread per CTA then publishes the tile to the partner (fence.acq_rel.cluster + a release.cluster arrive
//     on the partner's "peer full" barrier), so the cluster-scope fence is paid once per unit and off the transform warps
//     (first version: every transform warp arrived with release.cluster on both CTAs -- ncu r2e: 13 % of all stall samples sat
//     on that MEMBAR, another 11 % on the x-phase warps' release, which also waited for their global stores).  "Free" arrivals
//     (done reading) go to both CTAs' barriers with default semantics.
// No global-memory exchange, no extra pass; per unit 40 KB cross the SM-to-SM network against 80 KB x RT from HBM.
#pragma once
#include "po_fused3.cuh"

// ---- cluster primitives (GPU: PTX; emulator: a pair is ONE block of 2 x N threads over a 2 x shared-memory buffer) ----------
struct po_pair_ctx { int rank; int tid; i64 pair; i64 npairs; };
#ifdef PO_EMUL
PO_D po_pair_ctx po_pair_init(int nthreads) {
  po_pair_ctx c;
  c.rank = (int)threadIdx.x / nthreads; c.tid = (int)threadIdx.x % nthreads; c.pair = blockIdx.x; c.npairs = gridDim.x;
  return c;
}
PO_D unsigned char* po_pair_smem(unsigned char* base, int rank, size_t cta_bytes) { return base + (size_t)rank * cta_bytes; }
PO_D void po_pair_sync() { __syncthreads(); }
template <class T> PO_D T* po_pair_map(T* p, int my_rank, int to_rank, size_t cta_bytes) {
  return (T*)((unsigned char*)p + ((ptrdiff_t)to_rank - (ptrdiff_t)my_rank) * (ptrdiff_t)cta_bytes);
}
PO_D void po_mb_arrive_pair(po_mbar_t* mine, int my_rank, size_t cta_bytes) {
  po_mb_arrive(mine);
  po_mb_arrive(po_pair_map(mine, my_rank, my_rank ^ 1, cta_bytes));
}
PO_D void po_mb_publish_to_partner(po_mbar_t* mine, int my_rank, size_t cta_bytes) { po_mb_arrive(po_pair_map(mine, my_rank, my_rank ^ 1, cta_bytes)); }
PO_D bool po_mb_test_cl(po_mbar_t* b, unsigned parity) { return po_mb_test(b, parity); }
#else
PO_D po_pair_ctx po_pair_init(int) {
  po_pair_ctx c;
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  c.rank = (int)r; c.tid = (int)threadIdx.x; c.pair = blockIdx.x >> 1; c.npairs = gridDim.x >> 1;
  return c;
}
PO_D unsigned char* po_pair_smem(unsigned char* base, int, size_t) { return base; }
PO_D void po_pair_sync() {
  asm volatile("barrier.cluster.arrive.release;\n barrier.cluster.wait.acquire;" ::: "memory");
}
// generic address of the same shared-memory object in CTA `to_rank` of the cluster
template <class T> PO_D T* po_pair_map(T* p, int, int to_rank, size_t) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(to_rank));
  return (T*)out;
}
// "I am done READING": arrive on my own and on the partner's copy of the barrier.  Default semantics (release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive does for its consumer->producer barriers: the reads completed when their values were consumed,
// and a cluster-scope release here would also wait for the warp's outstanding GLOBAL stores (ncu r2e: 11 % of all stall samples).
PO_D void po_mb_arrive_pair(po_mbar_t* mine, int my_rank, size_t) {
  const unsigned local = po_smem_u32(mine);
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(my_rank ^ 1));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(local) : "memory");
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// "my CTA's shared-memory WRITES are complete": ONE thread, after it has acquired them at CTA scope, publishes them to the partner with
// a cluster-scope release (cumulative) on the partner's copy of the barrier
PO_D void po_mb_publish_to_partner(po_mbar_t* mine, int my_rank, size_t) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(po_smem_u32(mine)), "r"(my_rank ^ 1));
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// cluster-scope acquire: the partner CTA's shared-memory writes made before its arrive are visible after this succeeds
PO_D bool po_mb_test_cl(po_mbar_t* b, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(po_smem_u32(b)), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
#endif

PO_D void po_mb_wait_cl(po_mbar_t* b, unsigned parity, volatile int* aborted, int* err) {
  int spins = 0;
  while (!po_mb_test_cl(b, parity)) {
    po_mb_backoff();
    if ((++spins & 63) == 0) {
      if (*aborted) return;
#ifdef PO_EMUL
      if (spins > (1 << 28)) {
#else
      if (spins > (1 << 20)) {
#endif
        *aborted = 1; *err = 4;
        return;
      }
    }
  }
}

// ---- forward x-phase over a tile split in two x-halves (tlo: x < nx/2, thi: x >= nx/2; one of them is the partner's) ----------
// Same work-item numbering as po_fwd_xphase2 (item = (k * hp + p) * RX + j2); shuffle combine only (the tile is read-only here:
// the partner reads it at the same time).  Lh = floats per half-tile row (= I0 * 8 * RX).
template <int RX>
PO_D void po_fwd_xphase2_split(const float* __restrict__ tlo, const float* __restrict__ thi, const float4* __restrict__ tw4, const float* __restrict__ ksc,
                               float2* __restrict__ zu, i64 kstride, int I0, int Lh, int t, int nthreads, unsigned long long zpol, int item_begin, int item_end) {
  constexpr int NB = 16 / RX;
  const int hp = I0 >> 1;
  for (int base = item_begin; base < item_end; base += nthreads) {
    if (base + (t & ~31) >= item_end) break;  // warp-uniform (ranges are multiples of 32)
    const int item = base + t;
    const int j2 = item % RX, pk = item / RX;
    const int p = pk % hp, k = pk / hp;
    const float4* lo = (const float4*)(tlo + (size_t)k * 2 * Lh) + p;  // float4 slot of (pair p, x'): lo[hp * x']
    const float4* hi = (const float4*)(thi + (size_t)k * 2 * Lh) + p;
    float2 vr[16], vi[16];
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) {
      const float4 u = lo[hp * (RX * j1 + j2)];
      vr[j1] = make_float2(u.x, u.y);
      vi[j1] = make_float2(u.z, u.w);
    }
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) {
      const float4 u = hi[hp * (RX * j1 + j2)];
      vr[8 + j1] = make_float2(u.x, u.y);
      vi[8 + j1] = make_float2(u.z, u.w);
    }
    po_pk_fft16<-1>(vr, vi);
    const float4* twj = tw4 + j2 * PO_TW4_LD;
    float2 gr[16], gi[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 w = twj[r];
      const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
      const float2 fr = vr[po_brev4(r)], fi = vi[po_brev4(r)];
      gr[r] = __ffma2_rn(fi, po_n2(wi), __fmul2_rn(fr, wr));
      gi[r] = __ffma2_rn(fi, wr, __fmul2_rn(fr, wi));
    }
    po_lane_reduce_scatter<RX>(gr, gi, j2);
    const float ks = ksc[k];
    float2* zb = zu + 2 * p + kstride * k;
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) {
      const int r = qb * RX + j2;
      const float2 sr = gr[qb * RX], si = gi[qb * RX];
      const float4 zv = make_float4(sr.x * ks, si.x * ks, sr.y * ks, si.y * ks);
      if (zpol) po_st_hint((float4*)(zb + (i64)I0 * r), zv, zpol); else *(float4*)(zb + (i64)I0 * r) = zv;
    }
  }
}

static inline size_t po_pair_fwd_smem(i64 Lh, int RX) {
  return (size_t)(PO_RING_STAGES * 8 + 16) * Lh * sizeof(float) + (size_t)RX * PO_TW4_LD * sizeof(float4) + 8 * sizeof(float) +
         (2 * PO_RING_STAGES + 3) * sizeof(po_mbar_t) + 64;
}

template <int RT, int RX, int HP>
__global__ void __launch_bounds__(PO_RING_THREADS, 1)
po_fwd_tx_pair_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units,
                      int* __restrict__ err, int l2hint, unsigned cta_smem) {
  constexpr int D = PO_RING_STAGES, NT = PO_RING_TWARPS * 32, NX = PO_RING_XWARPS * 32;
  PO_DYN_SMEM(smem_all);
  const po_pair_ctx pc = po_pair_init(PO_RING_THREADS);
  unsigned char* smem_raw = po_pair_smem(smem_all, pc.rank, cta_smem);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX, Lh = L >> 1;                         // full row / half row (floats)
  float* ring = (float*)smem_raw;                                  // [D][8][Lh]
  float* T = ring + (size_t)D * 8 * Lh;                            // [8][2 Lh], pair-planar: my x-half of the reduced tile
  float4* tw4 = (float4*)(T + 16 * (size_t)Lh);                    // [RX][17]
  float* ksc = (float*)(tw4 + RX * PO_TW4_LD);                     // [8]
  po_mbar_t* full = (po_mbar_t*)(ksc + 8);                         // [D]
  po_mbar_t* empty = full + D;                                     // [D]
  po_mbar_t* tfull = empty + D;
  po_mbar_t* tfree = tfull + 1;
  po_mbar_t* tpeer = tfree + 1;                                    // the PARTNER's tile is complete and visible
  volatile int* aborted = (volatile int*)(tpeer + 1);
  const int tid = pc.tid, h = pc.rank;
  po_pdl_trigger();
  for (int e = tid; e < RX * 16; e += PO_RING_THREADS) {
    const float2 w = tw.tw_x[e];
    tw4[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x, w.x, w.y, w.y);
  }
  if (tid < 8) ksc[tid] = tw.kscale[tid];
  if (tid == 0) {
    for (int s = 0; s < D; ++s) { po_mb_init(&full[s], 1); po_mb_init(&empty[s], PO_RING_TWARPS); }
    po_mb_init(tfull, PO_RING_TWARPS);       // my transform warps (CTA scope)
    po_mb_init(tpeer, 1);                    // the partner's publishing thread (cluster scope)
    po_mb_init(tfree, 2 * PO_RING_XWARPS);   // both CTAs' x-phase warps
    *aborted = 0;
    po_mb_fence_init();
  }
  po_pair_sync();  // barriers of BOTH CTAs are initialised before anyone arrives remotely
  po_pdl_wait();
  const i64 nit = (units - pc.pair + pc.npairs - 1) / pc.npairs;   // units of this pair (both CTAs walk the same list)
  const i64 rstride = (i64)L * Mid;                                // floats between consecutive t rows
  const float* Tpartner = po_pair_map(T, h, h ^ 1, cta_smem);
  const float* tlo = h == 0 ? T : Tpartner;
  const float* thi = h == 0 ? Tpartner : T;
  const int x_items = (I0 >> 1) * 8 * RX, x_half = x_items >> 1;   // t-modes 4h .. 4h+3 are mine

  if (tid >= NT + NX) {
    if (tid == NT + NX) {  // producer: my x-half of every row
      const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;
      po_ring_state rs = {0, 0};
      for (i64 it = 0; it < nit; ++it) {
        const i64 unit = pc.pair + it * pc.npairs;
        const i64 m = unit % Mid, bb = unit / Mid;
        const float* ub = X + (i64)L * m + rstride * (i64)(8 * RT) * bb + (i64)h * Lh;
        for (int j2 = 0; j2 < RT; ++j2) {
          if (it * RT + j2 >= D) po_mb_wait(&empty[rs.s], rs.ph ^ 1u, aborted, err);
          float* dst = ring + (size_t)rs.s * 8 * Lh;
          po_mb_fill_begin(&full[rs.s], 8u * (unsigned)Lh * 4u);
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) po_mb_fill_row(dst + (size_t)j1 * Lh, ub + rstride * (i64)(RT * j1 + j2), (unsigned)Lh * 4u, &full[rs.s], pol);
          po_mb_fill_end(&full[rs.s]);
          po_ring_next(rs);
        }
      }
    }
  } else if (tid < NT) {
    // transform warps: t-decimation of my half rows, accumulators in registers
    const int v0 = tid, v1 = tid + NT;
    const bool has0 = v0 < (Lh >> 1), has1 = v1 < (Lh >> 1);
    po_ring_state rs = {0, 0};
    for (i64 it = 0; it < nit; ++it) {
      float2 ar0[8], ai0[8], ar1[8], ai1[8];
#define PO_PAIR_C(J2) po_ring_consume<RT, J2>(ring, full, empty, rs, Lh, v0, v1, has0, has1, ar0, ai0, ar1, ai1, aborted, err);
      PO_PAIR_C(0) PO_PAIR_C(1)
      if constexpr (RT > 2) { PO_PAIR_C(2) PO_PAIR_C(3) }
      if constexpr (RT > 4) { PO_PAIR_C(4) PO_PAIR_C(5) PO_PAIR_C(6) PO_PAIR_C(7) }
#undef PO_PAIR_C
      if (it > 0) po_mb_wait_cl(tfree, (unsigned)((it - 1) & 1), aborted, err);  // both CTAs' x-phases are done with my tile
      if (has0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) *(float4*)(T + k * 2 * Lh + 4 * v0) = make_float4(ar0[k].x, ar0[k].y, ai0[k].x, ai0[k].y);
      }
      if (has1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) *(float4*)(T + k * 2 * Lh + 4 * v1) = make_float4(ar1[k].x, ar1[k].y, ai1[k].x, ai1[k].y);
      }
      PO_SYNCWARP();
      if ((tid & 31) == 0) po_mb_arrive(tfull);
    }
  } else {
    // x-phase warps: my 4 t-modes, inputs from both half tiles
    const int tx = tid - NT;
    const unsigned long long zpol = l2hint ? po_policy_evict_last() : 0ull;
    for (i64 it = 0; it < nit; ++it) {
      const i64 unit = pc.pair + it * pc.npairs;
      const i64 m = unit % Mid, bb = unit / Mid;
      po_mb_wait(tfull, (unsigned)(it & 1), aborted, err);            // my CTA's half tile is complete
      if (tx == 0) po_mb_publish_to_partner(tpeer, h, cta_smem);      // ... tell the partner (one cluster-scope release per unit)
      po_mb_wait_cl(tpeer, (unsigned)(it & 1), aborted, err);         // the partner's half tile is complete and visible
      po_fwd_xphase2_split<RX>(tlo, thi, tw4, ksc, Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, I0, Lh, tx, NX, zpol, h * x_half,
                               (h + 1) * x_half);
      PO_SYNCWARP();
      if ((tid & 31) == 0) po_mb_arrive_pair(tfree, h, cta_smem);
    }
  }
  po_pair_sync();  // the partner may still be reading my tile / arriving on my barriers: leave together
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_PAIR)
// cudaErrorNotSupported: shape not handled here (caller falls back to the single-tile generation-2 kernel)
template <int RT, int RX>
static cudaError_t po_launch_fwd_pair_t(const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err, cudaStream_t st) {
  const i64 L = (i64)I0 * 16 * RX, Lh = L / 2;
  if ((I0 & 1) || num_sms < 2 || !err) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16) || ((Lh * 4) % 16)) return cudaErrorNotSupported;
  if (Lh / 2 > 2 * PO_RING_TWARPS * 32) return cudaErrorNotSupported;  // two pair-columns per transform thread
  const size_t smem = po_pair_fwd_smem(Lh, RX);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  const i64 units = Mid * B;
  if (units < 4 * (i64)(num_sms / 2)) return cudaErrorNotSupported;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-4 CTA-pair fwd RT=%d RX=%d I0=%d units=%lld smem/CTA=%zu\n", RT, RX, I0, (long long)units, smem);
#ifndef PO_EMUL
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms & ~1)); cfg.blockDim = dim3(PO_RING_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = po_pdl_enabled() ? 2 : 1;
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg);
    if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return cudaErrorNotSupported; }
    if (2 * nclusters < (int)cfg.gridDim.x) cfg.gridDim = dim3((unsigned)(2 * nclusters));  // persistent: one resident wave of pairs
    return cudaLaunchKernelEx(&cfg, kernel, (const float*)src, (float2*)dst, tw, I0, Mid, units, err, l2hint, (unsigned)smem);
  };
  cudaError_t e = (I0 == 20) ? go(po_fwd_tx_pair_kernel<RT, RX, 10>) : go(po_fwd_tx_pair_kernel<RT, RX, 0>);
  if (e != cudaSuccess) return e;
#else
  const unsigned npairs = (unsigned)(num_sms > 1 ? num_sms / 2 : 1);
  if (I0 == 20) { PO_LAUNCH((po_fwd_tx_pair_kernel<RT, RX, 10>), dim3(npairs), dim3(2 * PO_RING_THREADS), 2 * smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, err, l2hint, (unsigned)smem); }
  else { PO_LAUNCH((po_fwd_tx_pair_kernel<RT, RX, 0>), dim3(npairs), dim3(2 * PO_RING_THREADS), 2 * smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, err, l2hint, (unsigned)smem); }
#endif
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_fwd_pair(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                                      cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_PAIR_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_fwd_pair_t<rt, rx>(src, dst, tw, I0, Mid, B, num_sms, err, st);
  PO_PAIR_CASE(2, 8) PO_PAIR_CASE(4, 8) PO_PAIR_CASE(8, 8)
#undef PO_PAIR_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_fwd_pair(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                               cudaStream_t st);
#endif

// =======================================================================================
// CTA-pair INVERSE fused (x, t) kernel: the staged-input pipelined inverse of po_fused3.cuh for 10 KB rows.
//   * x-phase split by t-MODE: CTA h transforms its four modes (its 4 staged input rows, 10 KB per unit) and scatters the 16 outputs
//     of every packed FFT16 into the two half tiles -- x < nx/2 into CTA 0's, x >= nx/2 into CTA 1's (posted DSMEM stores from
//     the x-phase warps, which have slack; the streaming warps never touch remote memory);
//   * t-phase and store stream: CTA h owns the x-half h of every output row (contiguous 5 KB runs), all 8 modes from its local tile;
//   * the per-unit hand-over of the double-buffered tiles is one cluster barrier: the x-phase warps arrive with release (their
//     remote stores become visible), the streaming warps with a relaxed arrive (a release would wait for their global store stream).
template <int RX>
PO_D void po_inv_xphase2_split(const float2* __restrict__ zs, int ZR, float* __restrict__ Tlo, float* __restrict__ Thi, const float4* __restrict__ tw4k,
                               int I0, int Lh, int t, int nthreads, int kbase) {
  const int hp = I0 >> 1;
  const int nitems = hp * 4 * RX;
  for (int item = t; item < nitems; item += nthreads) {
    const int j2 = item % RX, pk = item / RX;
    const int p = pk % hp, kk = pk / hp;
    const float2* zb = zs + 2 * p + (size_t)ZR * kk;
    const float4* tw = tw4k + (kk * RX + j2) * PO_TW4_LD;
    float2 vr[16], vi[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 z = *(const float4*)(zb + I0 * r);  // (re_a, im_a, re_b, im_b)
      const float4 w = tw[r];
      const float2 zr = make_float2(z.x, z.z), zi = make_float2(z.y, z.w);
      const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
      vr[r] = __ffma2_rn(zi, po_n2(wi), __fmul2_rn(zr, wr));
      vi[r] = __ffma2_rn(zi, wr, __fmul2_rn(zr, wi));
    }
    po_pk_fft16<1>(vr, vi);
    float4* lo = (float4*)(Tlo + (size_t)(kbase + kk) * 2 * Lh) + p;
    float4* hi = (float4*)(Thi + (size_t)(kbase + kk) * 2 * Lh) + p;
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) {
      const float2 a = vr[po_brev4(j1)], b = vi[po_brev4(j1)];
      lo[hp * (RX * j1 + j2)] = make_float4(a.x, a.y, b.x, b.y);
    }
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) {
      const float2 a = vr[po_brev4(8 + j1)], b = vi[po_brev4(8 + j1)];
      hi[hp * (RX * j1 + j2)] = make_float4(a.x, a.y, b.x, b.y);
    }
  }
}

#ifdef PO_EMUL
PO_D void po_pair_sync_relaxed() { __syncthreads(); }
#else
PO_D void po_pair_sync_relaxed() { asm volatile("barrier.cluster.arrive.relaxed;\n barrier.cluster.wait.acquire;" ::: "memory"); }
#endif

static inline size_t po_pair_inv_smem(i64 Lh, int I0, int RX) {
  return (size_t)32 * Lh * sizeof(float) + (size_t)4 * RX * PO_TW4_LD * sizeof(float4) + (size_t)2 * 4 * I0 * 16 * sizeof(float2) + 2 * sizeof(po_mbar_t) + 64;
}

template <int RT, int RX, int HP, int NTHR>
__global__ void __launch_bounds__(NTHR, 1)
po_inv_xt_pair_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units, int nstream,
                      int l2hint, int* __restrict__ err, unsigned cta_smem) {
  PO_DYN_SMEM(smem_all);
  const po_pair_ctx pc = po_pair_init(NTHR);
  unsigned char* smem_raw = po_pair_smem(smem_all, pc.rank, cta_smem);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX, Lh = L >> 1, ZR = I0 * 16;   // full / half output row (floats), complex numbers per staged t-mode row
  float* T0 = (float*)smem_raw;                              // two [8][2 Lh] tiles: my x-half, all 8 t-modes
  float4* tw4k = (float4*)(T0 + 32 * (size_t)Lh);            // [4][RX][17]: my four t-modes
  float2* zs = (float2*)(tw4k + 4 * RX * PO_TW4_LD);         // [2][4][ZR] staged input rows of my four t-modes
  po_mbar_t* full = (po_mbar_t*)(zs + 2 * 4 * (size_t)ZR);   // [2]
  volatile int* aborted = (volatile int*)(full + 2);
  const int tid = pc.tid, h = pc.rank, kbase = 4 * h;
  po_pdl_trigger();
  const i64 rs2 = ((i64)L * Mid) >> 1;
  for (int e = tid; e < 4 * RX * 16; e += NTHR) {
    const int kk = e / (RX * 16);
    const float2 w = tw.tw_x[e % (RX * 16)];
    const float s = tw.kscale[kbase + kk];
    tw4k[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x * s, w.x * s, w.y * s, w.y * s);
  }
  if (tid == 0) {
    po_mb_init(&full[0], 1);
    po_mb_init(&full[1], 1);
    *aborted = 0;
    po_mb_fence_init();
  }
  po_pair_sync();
  po_pdl_wait();
  const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;
  const i64 nit = (units - pc.pair + pc.npairs - 1) / pc.npairs;
  const i64 kstride = (i64)ZR * Mid;
  float* T0p = po_pair_map(T0, h, h ^ 1, cta_smem);           // the partner's tiles
  auto issue = [&](i64 it) {
    const i64 unit = pc.pair + it * pc.npairs;
    const i64 m = unit % Mid, bb = unit / Mid;
    const float2* zb = Z + (i64)ZR * (m + Mid * 8 * bb) + kstride * kbase;
    float2* dst = zs + (size_t)(it & 1) * 4 * ZR;
    po_mbar_t* b = &full[it & 1];
    po_mb_fill_begin(b, 4u * (unsigned)ZR * 8u);
#pragma unroll
    for (int k = 0; k < 4; ++k) po_mb_fill_row(dst + (size_t)k * ZR, zb + kstride * k, (unsigned)ZR * 8u, b, 0ull);
    po_mb_fill_end(b);
  };
  if (tid == nstream && nit > 0) issue(0);
  for (i64 it = 0; it <= nit; ++it) {
    if (tid >= nstream) {
      if (tid == nstream && it + 1 < nit) issue(it + 1);
      if (it < nit) {
        po_mb_wait(&full[it & 1], (unsigned)((it >> 1) & 1), aborted, err);
        float* mine = T0 + (it & 1) * 16 * (size_t)Lh;
        float* theirs = T0p + (it & 1) * 16 * (size_t)Lh;
        po_inv_xphase2_split<RX>(zs + (size_t)(it & 1) * 4 * ZR, ZR, h == 0 ? mine : theirs, h == 0 ? theirs : mine, tw4k, I0, Lh, tid - nstream, NTHR - nstream,
                                 kbase);
      }
      po_pair_sync();          // release: my remote tile stores are visible to the partner's streaming warps after their wait
    } else {
      if (it > 0) {
        const i64 unit = pc.pair + (it - 1) * pc.npairs;
        const i64 m = unit % Mid, bb = unit / Mid;
        const unsigned ubase = (unsigned)((((i64)L * m + (i64)L * Mid * (i64)(8 * RT) * bb) + (i64)h * Lh) >> 1);
        po_inv_tphase2<RT, true>(T0 + ((it - 1) & 1) * 16 * (size_t)Lh, (float2*)Y, ubase, (unsigned)rs2, Lh, tid, nstream, I0, 0, I0, pol);
      }
      po_pair_sync_relaxed();  // done READING tile (it-1)&1 (the loads were consumed); no wait for the global store stream
    }
  }
  po_pair_sync();
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_PAIR)
template <int RT, int RX>
static cudaError_t po_launch_inv_pair_t(const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err, cudaStream_t st) {
  const i64 nx = 16 * RX, L = (i64)I0 * nx, Lh = L / 2;
  if ((I0 & 1) || num_sms < 2 || !err) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16) || ((Lh * 4) % 16)) return cudaErrorNotSupported;
  const size_t smem = po_pair_inv_smem(Lh, I0, RX);
  if (smem > 227 * 1024) return cudaErrorNotSupported;
  const i64 units = Mid * B;
  if (units < 4 * (i64)(num_sms / 2)) return cudaErrorNotSupported;
  if ((i64)I0 * nx * Mid * (8 * RT) * B / 2 >= (1LL << 32)) return cudaErrorNotSupported;  // 32-bit float2 row offsets
  constexpr int NTHR = RT <= 4 ? 640 : 512;
  const int ns = NTHR - 128;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-4 CTA-pair inv RT=%d RX=%d I0=%d units=%lld nstream=%d smem/CTA=%zu\n", RT, RX, I0, (long long)units, ns, smem);
#ifndef PO_EMUL
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(num_sms & ~1)); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = po_pdl_enabled() ? 2 : 1;
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg);
    if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return cudaErrorNotSupported; }
    if (2 * nclusters < (int)cfg.gridDim.x) cfg.gridDim = dim3((unsigned)(2 * nclusters));
    return cudaLaunchKernelEx(&cfg, kernel, (const float2*)src, (float*)dst, tw, I0, Mid, units, ns, l2hint, err, (unsigned)smem);
  };
  cudaError_t e = (I0 == 20) ? go(po_inv_xt_pair_kernel<RT, RX, 10, NTHR>) : go(po_inv_xt_pair_kernel<RT, RX, 0, NTHR>);
  if (e != cudaSuccess) return e;
#else
  const unsigned npairs = (unsigned)(num_sms > 1 ? num_sms / 2 : 1);
  if (I0 == 20) { PO_LAUNCH((po_inv_xt_pair_kernel<RT, RX, 10, NTHR>), dim3(npairs), dim3(2 * NTHR), 2 * smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, ns, l2hint, err, (unsigned)smem); }
  else { PO_LAUNCH((po_inv_xt_pair_kernel<RT, RX, 0, NTHR>), dim3(npairs), dim3(2 * NTHR), 2 * smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, ns, l2hint, err, (unsigned)smem); }
#endif
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_inv_pair(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                                      cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_PAIRI_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_inv_pair_t<rt, rx>(src, dst, tw, I0, Mid, B, num_sms, err, st);
  PO_PAIRI_CASE(2, 8) PO_PAIRI_CASE(4, 8) PO_PAIRI_CASE(8, 8)
#undef PO_PAIRI_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_inv_pair(int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms, int* err,
                               cudaStream_t st);
#endif
