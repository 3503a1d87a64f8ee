// Weight-streaming kernels for ParTensor (src/ParTensor.jl:35-113): per-mode channel mixing
//     y[o, m, b] = sum_i W(i, o, m) x[i, m, b]
// with rank-4 weights stored (ic, oc, modes...) and rank-5/6 weights stored (oc, ic, modes...).  The reference runs
// 3 permutedims + NNlib.batched_mul; at FNO sizes (C3(ii): ic = oc = 20, M = 32768 modes, ComplexF32) the 105 MB of
// weights are read ONCE per apply and everything else (x-hat, y-hat: 5 MB per sample) is noise, so the op is a pure
// HBM stream of W with ~0.8 real FMA per byte -- far below the FP32 pipe, let alone tensor cores: a tensor-core
// batched GEMM (b x ic)(ic x oc) has M-extent b = 1..16 and would idle on the same stream.  What matters is
//   * W is read exactly once, in whole contiguous chunks of G modes, by cp.async.bulk (TMA engine) into a
//     multi-stage shared-memory ring that one producer thread keeps full (bytes in flight ~ 150 KB / SM);
//   * both weight layouts and both contractions (forward: reduce over i; cotangent x-bar: reduce over o with conj(W))
//     are the same kernel: a mode's block is an (nf x ns) column-major matrix (nf = the fastest weight dim) and a
//     consumer thread owns one OUTPUT index of one mode, reducing either over the fast index (it walks a contiguous
//     row of its own, 128-bit LDS) or over the slow index (lanes walk adjacent columns: conflict-free);
//   * x-hat of the chunk rides along in the same stage (bulk copy), outputs are written coalesced.
// The parameter cotangent W-bar(i, o, m) = sum_b conj(x[i, m, b]) y-bar[o, m, b] is a pure 105 MB WRITE stream
// (po_tensor_wbar_stream_kernel: operands staged per chunk, 128-bit coalesced stores).
#pragma once
#include "po_fused3.cuh"  // mbarrier / bulk-copy primitives (GPU: PTX, emulator: counters + memcpy)

#define PO_TS_CONSUMERS 256
#define PO_TS_THREADS 288   // 8 consumer warps + 1 producer warp
#define PO_TS_BMAX 4        // batch columns per pass (x-hat of the chunk for BMAX columns sits in the stage)

struct po_ts_ring { int s; unsigned ph; };

template <class T> struct po_ts_vec { static constexpr bool ok = false; };
template <> struct po_ts_vec<float2> { static constexpr bool ok = true; };

// y[out] = sum_in W * x over one mode's (nf x ns) block, blk = nf * ns.
// RED_FAST: out = slow index s, in = fast index f (thread walks the contiguous row W[. + nf*s]);
// else    : out = fast index f, in = slow index s (thread walks the strided column W[f + nf*.]).
// NCT > 0: the contracted extent is the compile-time constant NCT (FNO widths: fully unrolled, loads issued back to back, two
// independent accumulator chains); NCT == 0: run-time extent.
template <class T, bool RED_FAST, bool CONJ, int NCT>
PO_D T po_ts_dot(const T* __restrict__ Wm, const T* __restrict__ xv, int nf, int ns, int out) {
  T acc0 = po_zero<T>(), acc1 = po_zero<T>();
  if constexpr (RED_FAST) {
    const T* row = Wm + (size_t)nf * out;
    const int n = NCT ? NCT : nf;
    if constexpr (po_ts_vec<T>::ok) {
      if ((n & 1) == 0) {  // rows of an even number of complex floats start 16-byte aligned (stage base is)
        const float4* r4 = (const float4*)row;
        const float4* x4 = (const float4*)xv;
#pragma unroll
        for (int f = 0; f < (NCT ? NCT / 2 : (n >> 1)); ++f) {
          const float4 w = r4[f], x = x4[f];
          po_mac(acc0, CONJ ? make_float2(w.x, -w.y) : make_float2(w.x, w.y), make_float2(x.x, x.y));
          po_mac(acc1, CONJ ? make_float2(w.z, -w.w) : make_float2(w.z, w.w), make_float2(x.z, x.w));
        }
        return po_add(acc0, acc1);
      }
    }
#pragma unroll
    for (int f = 0; f < (NCT ? NCT : n); ++f) po_mac(acc0, CONJ ? po_conj(row[f]) : row[f], xv[f]);
    return acc0;
  } else {
    const T* col = Wm + out;
    const int n = NCT ? NCT : ns;
    int s = 0;
#pragma unroll
    for (; s + 1 < (NCT ? NCT : n); s += 2) {
      po_mac(acc0, CONJ ? po_conj(col[(size_t)nf * s]) : col[(size_t)nf * s], xv[s]);
      po_mac(acc1, CONJ ? po_conj(col[(size_t)nf * (s + 1)]) : col[(size_t)nf * (s + 1)], xv[s + 1]);
    }
    if (s < n) po_mac(acc0, CONJ ? po_conj(col[(size_t)nf * s]) : col[(size_t)nf * s], xv[s]);
    return po_add(acc0, acc1);
  }
}

// Every consumer WARP owns whole chunks (chunk `it` of the CTA goes to warp it mod 8): eight chunks are being reduced at the same
// time, each behind its own full / empty barrier pair, so no warp ever waits for another one's arithmetic.  (First version: all
// 8 warps shared every chunk -- 180 work items for 256 threads and a CTA-wide hand-over per 29 KB: 3.6 TB/s on the C3(ii) weights.)
template <class T, bool RED_FAST, bool CONJ, int NCT>
__global__ void __launch_bounds__(PO_TS_THREADS, 1)
po_tensor_stream_kernel(const T* __restrict__ W, const T* __restrict__ X, T* __restrict__ Y, int nf, int ns, i64 M, int B0, int BB, i64 Bstride_in,
                        i64 Bstride_out, int G, int nstages, i64 nchunks, int* __restrict__ err, int l2hint) {
  PO_DYN_SMEM(smem_raw);
  constexpr int NW = PO_TS_CONSUMERS / 32;
  const int blk = nf * ns;
  const int nin = RED_FAST ? nf : ns, nout = RED_FAST ? ns : nf;
  const size_t stage_elems = (size_t)G * blk + (size_t)BB * G * nin;  // W chunk, then x chunk [bb][ml][in]
  const size_t stage_bytes = (stage_elems * sizeof(T) + 15) & ~(size_t)15;
  unsigned char* ring = smem_raw;
  po_mbar_t* full = (po_mbar_t*)(ring + (size_t)nstages * stage_bytes);
  po_mbar_t* empty = full + nstages;
  volatile int* aborted = (volatile int*)(empty + nstages);
  const int tid = threadIdx.x;
  po_pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { po_mb_init(&full[s], 1); po_mb_init(&empty[s], 1); }
    *aborted = 0;
    po_mb_fence_init();
  }
  __syncthreads();
  po_pdl_wait();
  const i64 nit = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (tid >= PO_TS_CONSUMERS) {
    if (tid == PO_TS_CONSUMERS) {  // producer
      const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;  // weights are read once
      po_ts_ring rs = {0, 0};
      for (i64 it = 0; it < nit; ++it) {
        const i64 c = blockIdx.x + it * gridDim.x;
        const i64 m0 = c * G;
        const int g = (int)((M - m0) < G ? (M - m0) : G);
        if (it >= nstages) po_mb_wait(&empty[rs.s], rs.ph ^ 1u, aborted, err);
        T* st = (T*)(ring + (size_t)rs.s * stage_bytes);
        const unsigned wbytes = (unsigned)((size_t)g * blk * sizeof(T)), xbytes = (unsigned)((size_t)g * nin * sizeof(T));
        po_mb_fill_begin(&full[rs.s], wbytes + (unsigned)BB * xbytes);
        po_mb_fill_row(st, W + (size_t)m0 * blk, wbytes, &full[rs.s], pol);
        for (int bb = 0; bb < BB; ++bb)
          po_mb_fill_row(st + (size_t)G * blk + (size_t)bb * G * nin, X + (size_t)nin * m0 + Bstride_in * (i64)(B0 + bb), xbytes, &full[rs.s], 0ull);
        po_mb_fill_end(&full[rs.s]);
        if (++rs.s == nstages) { rs.s = 0; rs.ph ^= 1u; }
      }
    }
    return;
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (i64 it = warp; it < nit; it += NW) {
    const int stg = (int)(it % nstages);
    const unsigned ph = (unsigned)((it / nstages) & 1);
    const i64 c = blockIdx.x + it * gridDim.x;
    const i64 m0 = c * G;
    const int g = (int)((M - m0) < G ? (M - m0) : G);
    po_mb_wait(&full[stg], ph, aborted, err);
    const T* st = (const T*)(ring + (size_t)stg * stage_bytes);
    const T* xs = st + (size_t)G * blk;
    const int items = g * nout;
    for (int item = lane; item < items; item += 32) {
      const int ml = item / nout, out = item - ml * nout;
      for (int bb = 0; bb < BB; ++bb) {
        const T acc = po_ts_dot<T, RED_FAST, CONJ, NCT>(st + (size_t)ml * blk, xs + ((size_t)bb * G + ml) * nin, nf, ns, out);
        Y[(size_t)out + (size_t)nout * (size_t)(m0 + ml) + (size_t)Bstride_out * (size_t)(B0 + bb)] = acc;
      }
    }
    PO_SYNCWARP();
    if (lane == 0) po_mb_arrive(&empty[stg]);
  }
}

// W-bar(f, s, m) = sum_b conj(x[i]) * ybar[o]   with (i, o) = w_oi ? (s, f) : (f, s); written in weight order
template <class T>
__global__ void __launch_bounds__(256)
po_tensor_wbar_stream_kernel(const T* __restrict__ X, const T* __restrict__ Yb, T* __restrict__ Wb, int ic, int oc, i64 M, i64 B, int w_oi, int G,
                             i64 nchunks) {
  PO_DYN_SMEM(smem_raw);
  const int blk = ic * oc, nf = w_oi ? oc : ic;
  T* xs = (T*)smem_raw;                 // [b][ml][ic]
  T* ys = xs + (size_t)B * G * ic;      // [b][ml][oc]
  po_pdl_trigger();
  po_pdl_wait();
  for (i64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const i64 m0 = c * G;
    const int g = (int)((M - m0) < G ? (M - m0) : G);
    __syncthreads();  // previous chunk's operands are no longer read
    for (i64 e = threadIdx.x; e < (i64)B * g * ic; e += blockDim.x) {
      const int b = (int)(e / ((i64)g * ic)), r = (int)(e - (i64)b * g * ic);
      xs[(size_t)b * G * ic + r] = X[(size_t)ic * m0 + r + (size_t)ic * M * b];
    }
    for (i64 e = threadIdx.x; e < (i64)B * g * oc; e += blockDim.x) {
      const int b = (int)(e / ((i64)g * oc)), r = (int)(e - (i64)b * g * oc);
      ys[(size_t)b * G * oc + r] = Yb[(size_t)oc * m0 + r + (size_t)oc * M * b];
    }
    __syncthreads();
    T* wout = Wb + (size_t)m0 * blk;
    // (i, o) of a weight slot depend on the slot only, not on the mode: one division per slot, none per element
    for (int r = threadIdx.x; r < blk; r += blockDim.x) {
      const int sidx = r / nf, f = r - sidx * nf;
      const int i = w_oi ? sidx : f, o = w_oi ? f : sidx;
      for (int ml = 0; ml < g; ++ml) {
        T acc = po_zero<T>();
        for (int b = 0; b < (int)B; ++b) po_mac(acc, po_conj(xs[((size_t)b * G + ml) * ic + i]), ys[((size_t)b * G + ml) * oc + o]);
        wout[(size_t)ml * blk + r] = acc;
      }
    }
  }
}

// ---- launchers --------------------------------------------------------------------------------------------------
struct po_ts_cfg { int G, nstages; size_t smem; };

// chunk size / stage count for a block of `blk` elements and `nin` inputs per mode; G = 0: shape not handled here
static inline po_ts_cfg po_ts_config(size_t esz, int blk, int nin, int BB) {
  po_ts_cfg c = {0, 0, 0};
  // bulk copies need 16-byte aligned sources and sizes for every chunk start: G * blk * esz and G * nin * esz multiples of 16
  int gstep = 1;
  while (gstep <= 16 && (((size_t)gstep * blk * esz) % 16 || ((size_t)gstep * nin * esz) % 16)) ++gstep;
  if (gstep > 16) return c;
  const size_t per_mode = ((size_t)blk + (size_t)BB * nin) * esz;
  static const int chunk_kb = getenv("PO_TS_CHUNK_KB") ? atoi(getenv("PO_TS_CHUNK_KB")) : 18;  // ncu (profiles/r2h_*): 7 KB 37.5 us, 14 KB 22.5 us, 18 KB 20.3 us, 24 KB 20.1 us on the C3(ii) weights
  int G = (int)(((size_t)chunk_kb * 1024) / per_mode);   // ~18 KB chunks: one warp reduces a chunk; the ring must hold >= 8 stages (one per consumer warp)
  G -= G % gstep;
  if (G < gstep) G = gstep;
  const size_t stage = (((size_t)G * per_mode) + 15) & ~(size_t)15;
  if (stage > 96 * 1024) return c;
  int ns = (int)((200 * 1024) / stage);
  if (ns > 16) ns = 16;
  // a consumer warp's next chunk is 8 chunks ahead: with fewer than 8 stages it would wait on a stage a whole ring cycle early, where
  // the phase-parity test of the PREVIOUS phase answers "complete" at once -- the ring must be at least as deep as the warp count
  if (ns < PO_TS_CONSUMERS / 32) return c;
  c.G = G; c.nstages = ns; c.smem = (size_t)ns * stage + 2 * (size_t)ns * sizeof(po_mbar_t) + 64;
  return c;
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_TENSOR)
template <class T, bool RED_FAST, bool CONJ, int NCT>
static cudaError_t po_launch_tensor_stream_k(const void* W, const void* X, void* Y, int nf, int ns, i64 M, i64 B, int num_sms, int* err, cudaStream_t st) {
  const int blk = nf * ns, nin = RED_FAST ? nf : ns, nout = RED_FAST ? ns : nf;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  for (i64 b0 = 0; b0 < B; b0 += PO_TS_BMAX) {
    const int BB = (int)((B - b0) < PO_TS_BMAX ? (B - b0) : PO_TS_BMAX);
    const po_ts_cfg c = po_ts_config(sizeof(T), blk, nin, BB);
    if (!c.G) return cudaErrorNotSupported;
    const i64 tail = M % c.G;  // the last chunk's copies must be 16-byte multiples too
    if (tail && ((((size_t)tail * blk * sizeof(T)) % 16) || (((size_t)tail * nin * sizeof(T)) % 16))) return cudaErrorNotSupported;
    const i64 nchunks = (M + c.G - 1) / c.G;
    int grid = (int)(nchunks < num_sms ? nchunks : num_sms);
    if (grid < 1) grid = 1;
    if (trace) fprintf(stderr, "[parops] tensor stream %s%s nf=%d ns=%d M=%lld B=%d G=%d stages=%d smem=%zu\n", RED_FAST ? "red-fast" : "red-slow", CONJ ? " conj" : "", nf, ns, (long long)M, BB, c.G, c.nstages, c.smem);
#ifndef PO_EMUL
    cudaError_t e = cudaFuncSetAttribute(po_tensor_stream_kernel<T, RED_FAST, CONJ, NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
#endif
    PO_LAUNCH_PDL((po_tensor_stream_kernel<T, RED_FAST, CONJ, NCT>), dim3((unsigned)grid), dim3(PO_TS_THREADS), c.smem, st, (const T*)W, (const T*)X, (T*)Y, nf, ns, M,
                  (int)b0, BB, (i64)nin * M, (i64)nout * M, c.G, c.nstages, nchunks, err, l2hint);
  }
  return cudaGetLastError();
}

template <class T>
static cudaError_t po_launch_tensor_stream_t(bool xbar, const void* W, const void* X, void* Y, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms, int* err,
                                             cudaStream_t st) {
  // forward reduces over i, the cotangent over o (with conj(W)); w_oi = 1: o is the fast weight dim
  const int nf = w_oi ? oc : ic, ns = w_oi ? ic : oc;
  const bool red_fast = xbar ? (w_oi != 0) : (w_oi == 0);
  const int nred = red_fast ? nf : ns;  // contracted extent
#define PO_TS_GO(RF, CJ)                                                                                                               \
  do {                                                                                                                                 \
    if (sizeof(T) == 8 && po_ts_vec<T>::ok && nred == 20) return po_launch_tensor_stream_k<T, RF, CJ, 20>(W, X, Y, nf, ns, M, B, num_sms, err, st); \
    if (sizeof(T) == 8 && po_ts_vec<T>::ok && nred == 32) return po_launch_tensor_stream_k<T, RF, CJ, 32>(W, X, Y, nf, ns, M, B, num_sms, err, st); \
    return po_launch_tensor_stream_k<T, RF, CJ, 0>(W, X, Y, nf, ns, M, B, num_sms, err, st);                                            \
  } while (0)
  if (!xbar) { if (red_fast) PO_TS_GO(true, false); else PO_TS_GO(false, false); }
  else { if (red_fast) PO_TS_GO(true, true); else PO_TS_GO(false, true); }
#undef PO_TS_GO
}

// cudaErrorNotSupported: shape / alignment not handled here (caller falls back to the generic kernels)
PO_XTU cudaError_t po_launch_tensor_stream(int dt, bool xbar, const void* W, const void* X, void* Y, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms,
                                           int* err, cudaStream_t st) {
  if (!err || num_sms <= 0 || M <= 0 || B <= 0 || ic <= 0 || oc <= 0 || (i64)ic * oc > (1 << 20)) return cudaErrorNotSupported;
  if (((size_t)W % 16) || ((size_t)X % 16) || ((size_t)Y % 16)) return cudaErrorNotSupported;
  const size_t esz = po_dtype_size(dt);
  // every batch column of X must start 16-byte aligned for the bulk copies
  if ((((size_t)(xbar ? oc : ic) * (size_t)M * esz) % 16) != 0 && B > 1) return cudaErrorNotSupported;
  if ((i64)ic * oc * M < (i64)num_sms * 4096) return cudaErrorNotSupported;  // small problems: the generic kernel is launch-bound anyway
  switch (dt) {
    case PO_F32: return po_launch_tensor_stream_t<float>(xbar, W, X, Y, ic, oc, M, B, w_oi, num_sms, err, st);
    case PO_F64: return po_launch_tensor_stream_t<double>(xbar, W, X, Y, ic, oc, M, B, w_oi, num_sms, err, st);
    case PO_C64: return po_launch_tensor_stream_t<float2>(xbar, W, X, Y, ic, oc, M, B, w_oi, num_sms, err, st);
    case PO_C128: return po_launch_tensor_stream_t<double2>(xbar, W, X, Y, ic, oc, M, B, w_oi, num_sms, err, st);
  }
  return cudaErrorInvalidValue;
}

template <class T>
static cudaError_t po_launch_tensor_wbar_t(const void* X, const void* Yb, void* Wb, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms, cudaStream_t st) {
  const size_t per_mode = (size_t)B * (ic + oc) * sizeof(T);
  int G = (int)((48 * 1024) / per_mode);
  if (G < 1) return cudaErrorNotSupported;
  if (G > 16) G = 16;
  const size_t smem = (size_t)G * per_mode + 16;
  const i64 nchunks = (M + G - 1) / G;
  i64 grid = nchunks < (i64)num_sms * 8 ? nchunks : (i64)num_sms * 8;
  if (grid < 1) grid = 1;
#ifndef PO_EMUL
  cudaError_t e = cudaFuncSetAttribute(po_tensor_wbar_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
#endif
  PO_LAUNCH_PDL((po_tensor_wbar_stream_kernel<T>), dim3((unsigned)grid), dim3(256), smem, st, (const T*)X, (const T*)Yb, (T*)Wb, ic, oc, M, B, w_oi, G, nchunks);
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_tensor_wbar(int dt, const void* X, const void* Yb, void* Wb, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms, cudaStream_t st) {
  if (num_sms <= 0 || M <= 0 || B <= 0 || (i64)ic * oc * M < (i64)num_sms * 4096) return cudaErrorNotSupported;
  switch (dt) {
    case PO_F32: return po_launch_tensor_wbar_t<float>(X, Yb, Wb, ic, oc, M, B, w_oi, num_sms, st);
    case PO_F64: return po_launch_tensor_wbar_t<double>(X, Yb, Wb, ic, oc, M, B, w_oi, num_sms, st);
    case PO_C64: return po_launch_tensor_wbar_t<float2>(X, Yb, Wb, ic, oc, M, B, w_oi, num_sms, st);
    case PO_C128: return po_launch_tensor_wbar_t<double2>(X, Yb, Wb, ic, oc, M, B, w_oi, num_sms, st);
  }
  return cudaErrorInvalidValue;
}
#else
cudaError_t po_launch_tensor_stream(int dt, bool xbar, const void* W, const void* X, void* Y, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms,
                                    int* err, cudaStream_t st);
cudaError_t po_launch_tensor_wbar(int dt, const void* X, const void* Yb, void* Wb, int ic, int oc, i64 M, i64 B, int w_oi, int num_sms, cudaStream_t st);
#endif
