// FP64 dense mode products on large factors (BASELINE.json configs[4]: ParKron of three 256 x 256 ParMatrix factors on a 256^3
// Float64 vector, src/ParMatrix.jl:42-46 inside src/ParKron.jl:190-217): compute-bound on the FP64 pipe (25.8 GFLOP against 0.27 GB).
//
// Y(I, m, O) = A(m x n) X(I, n, O) is a GEMM in both layouts the Kron produces:
//   I == 1 : Y[m x ncols] = A X,          X columns contiguous in the contracted index   (P = A,   Q = X,   one problem)
//   I  > 1 : Y_o[I x m]   = X_o A^T,      X rows contiguous in the free index             (P = X_o, Q = A^T, one problem per o)
// Two arms, both 128 x 128 x 8 tiles, 256 threads, operands staged through registers into a double-buffered shared tile (the next
// tile's global loads are in flight while the current one is multiplied; ONE block barrier per k-step):
//   * SIMT: 8 x 8 accumulators per thread (2-wide chunks -> 128-bit LDS), 64 DFMA per 8 LDS.128;
//   * DMMA: mma.sync.aligned.m8n8k4.f64 warp tiles of 32 x 64 (the legacy tensor path: tcgen05 has no FP64 kind).
// Measured FP64 peaks on this part (scripts/microbench/mb_fp64.cu): DFMA 34.0 TFLOP/s, DMMA 37.1 TFLOP/s.
// The generic po_dense_mode_kernel (64 x 64 tile, 4 x 4 accumulators, no prefetch) reached 13.5 TFLOP/s = 0.40 here.
#pragma once
#include "po_common.cuh"
#include <stdio.h>
#include <stdlib.h>

#define PO_DG_BM 128
#define PO_DG_BN 128
#define PO_DG_BK 8
#define PO_DG_THREADS 256

// 0: off (generic kernel), 1: SIMT arm (128 x 128 tiles), 2: DMMA arm, 3: SIMT arm with 128 x 64 tiles / two CTAs per SM.  PO_DGEMM overrides.
static inline int po_dgemm_arm() {
  const char* env = getenv("PO_DGEMM");  // read per call (host side): scripts/bench_configs.py times both arms in one process
  return env ? atoi(env) : 2;  // measured (profiles/r2g_*): DMMA 21.3, SIMT 128 x 128 20.1, SIMT 128 x 64 19.4, generic 13.5 TFLOP/s on the C5 matvec
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_DGEMM)
struct po_dg_args {
  const double* A; i64 a_rs, a_cs;   // parameter matrix element (r, c) = A[r * a_rs + c * a_cs], m x n as applied
  const double* X; double* Y;
  i64 I; int n, m; i64 ncols;        // ncols = I * O
  int x_is_p;                        // 1: I > 1 layout (P = X_o), 0: I == 1 layout (Q = X)
};

// global -> registers: one k-step of both tiles.  pv: 4 doubles of the P tile [BK][BM], qv: 4 (BN = 128) or 2 (BN = 64) doubles of the Q tile [BK][BN]
template <int BN>
PO_D void po_dg_load(const po_dg_args& a, i64 pbase, i64 qbase, int k0, int tid, double (&pv)[4], double (&qv)[4]) {
  constexpr int QE = BN / 32;  // Q elements per thread
  if (a.x_is_p) {
    // P = X_o: rows (free index i) contiguous.  thread -> k = tid / 32, rows 4 * (tid % 32) .. + 3
    const int k = tid >> 5, r = (tid & 31) << 2;
    const double* p = a.X + pbase + (i64)r + a.I * (i64)(k0 + k);
    const double2 u0 = *(const double2*)p, u1 = *(const double2*)(p + 2);
    pv[0] = u0.x; pv[1] = u0.y; pv[2] = u1.x; pv[3] = u1.y;
    // Q = A^T: Q[k][j] = A(j, k).  thread -> j = tid % BN, k = QE * (tid / BN) .. + QE - 1
    if (a.a_cs == 1 && a.a_rs != 1) {  // adjoint parameter: k is the contiguous index -> threads along k (8 doubles = 64 B per row)
      const int k = tid & 7, j = tid >> 3;
#pragma unroll
      for (int e = 0; e < QE; ++e) qv[e] = __ldg(a.A + (qbase + j + 32 * e) * a.a_rs + (k0 + k));
    } else {
      const int j = tid % BN, kq = (tid / BN) * QE;
      const double* q = a.A + (qbase + j) * a.a_rs + (i64)(k0 + kq) * a.a_cs;
#pragma unroll
      for (int e = 0; e < QE; ++e) qv[e] = __ldg(q + (i64)e * a.a_cs);
    }
  } else {
    // P = A: P[k][r] = A(r, k).  thread -> r = tid % 128, k = 4 * (tid / 128) .. + 3
    if (a.a_cs == 1 && a.a_rs != 1) {
      const int k = tid & 7, r = tid >> 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) pv[e] = __ldg(a.A + (pbase + r + 32 * e) * a.a_rs + (k0 + k));
    } else {
      const int r = tid & 127, kp = (tid >> 7) << 2;
      const double* p = a.A + (pbase + r) * a.a_rs + (i64)(k0 + kp) * a.a_cs;
#pragma unroll
      for (int e = 0; e < 4; ++e) pv[e] = __ldg(p + (i64)e * a.a_cs);
    }
    // Q = X: columns contiguous in k.  thread -> column tid % BN, k = QE * (tid / BN) .. (16 or 32 contiguous bytes)
    const int c = tid % BN, kq = (tid / BN) * QE;
    const double* q = a.X + (qbase + c) * (i64)a.n + (k0 + kq);
    const double2 u0 = *(const double2*)q;
    qv[0] = u0.x; qv[1] = u0.y;
    if constexpr (QE == 4) { const double2 u1 = *(const double2*)(q + 2); qv[2] = u1.x; qv[3] = u1.y; }
  }
}
// registers -> shared tiles Ps[BK][BM], Qs[BK][BN]
template <int BN>
PO_D void po_dg_store(const po_dg_args& a, double* Ps, double* Qs, int tid, const double (&pv)[4], const double (&qv)[4]) {
  constexpr int QE = BN / 32;
  if (a.x_is_p) {
    const int k = tid >> 5, r = (tid & 31) << 2;
    *(double2*)(Ps + k * PO_DG_BM + r) = make_double2(pv[0], pv[1]);
    *(double2*)(Ps + k * PO_DG_BM + r + 2) = make_double2(pv[2], pv[3]);
    if (a.a_cs == 1 && a.a_rs != 1) {
      const int k = tid & 7, j = tid >> 3;
#pragma unroll
      for (int e = 0; e < QE; ++e) Qs[k * BN + j + 32 * e] = qv[e];
    } else {
      const int j = tid % BN, kq = (tid / BN) * QE;
#pragma unroll
      for (int e = 0; e < QE; ++e) Qs[(kq + e) * BN + j] = qv[e];
    }
  } else {
    if (a.a_cs == 1 && a.a_rs != 1) {
      const int k = tid & 7, r = tid >> 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) Ps[k * PO_DG_BM + r + 32 * e] = pv[e];
    } else {
      const int r = tid & 127, kp = (tid >> 7) << 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) Ps[(kp + e) * PO_DG_BM + r] = pv[e];
    }
    const int c = tid % BN, kq = (tid / BN) * QE;
#pragma unroll
    for (int e = 0; e < QE; ++e) Qs[(kq + e) * BN + c] = qv[e];
  }
}

// tile origin of this CTA: returns the global offsets (elements) of P / Q / C tiles and the leading dimension of C
template <int BN>
PO_D void po_dg_tile(const po_dg_args& a, i64& pbase, i64& qbase, double*& C, i64& ldc) {
  if (a.x_is_p) {  // grid.x: row tiles of I, grid.y: column tiles of m, grid.z: o
    const i64 i0 = (i64)blockIdx.x * PO_DG_BM, j0 = (i64)blockIdx.y * BN, o = blockIdx.z;
    pbase = i0 + a.I * (i64)a.n * o;
    qbase = j0;
    C = a.Y + i0 + a.I * (j0 + (i64)a.m * o);
    ldc = a.I;
  } else {         // grid.x: column tiles of ncols, grid.y: row tiles of m
    const i64 c0 = (i64)blockIdx.x * BN, r0 = (i64)blockIdx.y * PO_DG_BM;
    pbase = r0;
    qbase = c0;
    C = a.Y + r0 + (i64)a.m * c0;
    ldc = a.m;
  }
}

// BN = 128: 8 x 8 accumulators, one CTA per SM;  BN = 64: 8 x 4 accumulators, two CTAs per SM (one CTA's tile prologue / epilogue
// overlaps the other's multiply)
template <int BN>
__global__ void __launch_bounds__(PO_DG_THREADS, BN == 128 ? 1 : 2)
po_dgemm_simt_kernel(const po_dg_args a) {
  constexpr int TN = BN / 16, C2 = TN / 2;
  __align__(16) __shared__ double Ps[2][PO_DG_BK * PO_DG_BM];
  __align__(16) __shared__ double Qs[2][PO_DG_BK * BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  i64 pbase, qbase, ldc;
  double* C;
  po_dg_tile<BN>(a, pbase, qbase, C, ldc);
  double acc[8][TN];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.0;
  double pv[4], qv[4];
  po_dg_load<BN>(a, pbase, qbase, 0, tid, pv, qv);
  po_dg_store<BN>(a, Ps[0], Qs[0], tid, pv, qv);
  __syncthreads();
  const int nk = a.n / PO_DG_BK;
  for (int ks = 0; ks < nk; ++ks) {
    const int cur = ks & 1;
    if (ks + 1 < nk) po_dg_load<BN>(a, pbase, qbase, (ks + 1) * PO_DG_BK, tid, pv, qv);  // in flight during the multiply
    const double* ps = Ps[cur];
    const double* qs = Qs[cur];
#pragma unroll
    for (int kk = 0; kk < PO_DG_BK; ++kk) {
      double av[8], bv[TN];
#pragma unroll
      for (int r2 = 0; r2 < 4; ++r2) {  // rows (r2 * 16 + ty) * 2 + {0, 1}
        const double2 u = *(const double2*)(ps + kk * PO_DG_BM + ((r2 * 16 + ty) << 1));
        av[2 * r2] = u.x; av[2 * r2 + 1] = u.y;
      }
#pragma unroll
      for (int c2 = 0; c2 < C2; ++c2) {
        const double2 u = *(const double2*)(qs + kk * BN + ((c2 * 16 + tx) << 1));
        bv[2 * c2] = u.x; bv[2 * c2 + 1] = u.y;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
    }
    if (ks + 1 < nk) {
      po_dg_store<BN>(a, Ps[cur ^ 1], Qs[cur ^ 1], tid, pv, qv);  // the other buffer was last read before the previous barrier
      __syncthreads();
    }
  }
#pragma unroll
  for (int c2 = 0; c2 < C2; ++c2)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = ((c2 * 16 + tx) << 1) + e;
#pragma unroll
      for (int r2 = 0; r2 < 4; ++r2) {
        const int row = (r2 * 16 + ty) << 1;
        *(double2*)(C + row + ldc * (i64)col) = make_double2(acc[2 * r2][2 * c2 + e], acc[2 * r2 + 1][2 * c2 + e]);
      }
    }
}

#ifndef PO_EMUL
PO_D void po_dmma_884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
#endif

// DMMA arm: 8 warps in a 4 (rows) x 2 (columns) arrangement, warp tile 32 x 64 = 4 x 8 mma tiles of 8 x 8
__global__ void __launch_bounds__(PO_DG_THREADS, 1)
po_dgemm_dmma_kernel(const po_dg_args a) {
#ifndef PO_EMUL
  __align__(16) __shared__ double Ps[2][PO_DG_BK * PO_DG_BM];
  __align__(16) __shared__ double Qs[2][PO_DG_BK * PO_DG_BN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = (warp & 3) * 32, wc = (warp >> 2) * 64;
  const int g = lane >> 2, t4 = lane & 3;  // fragment coordinates: A(row g, k t4), B(k t4, col g), C(row g, cols 2 t4, 2 t4 + 1)
  i64 pbase, qbase, ldc;
  double* C;
  po_dg_tile<PO_DG_BN>(a, pbase, qbase, C, ldc);
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double pv[4], qv[4];
  po_dg_load<PO_DG_BN>(a, pbase, qbase, 0, tid, pv, qv);
  po_dg_store<PO_DG_BN>(a, Ps[0], Qs[0], tid, pv, qv);
  __syncthreads();
  const int nk = a.n / PO_DG_BK;
  for (int ks = 0; ks < nk; ++ks) {
    const int cur = ks & 1;
    if (ks + 1 < nk) po_dg_load<PO_DG_BN>(a, pbase, qbase, (ks + 1) * PO_DG_BK, tid, pv, qv);
    const double* ps = Ps[cur];
    const double* qs = Qs[cur];
#pragma unroll
    for (int k4 = 0; k4 < PO_DG_BK; k4 += 4) {
      double af[4], bf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = ps[(k4 + t4) * PO_DG_BM + wr + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 8; ++j) bf[j] = qs[(k4 + t4) * PO_DG_BN + wc + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) po_dmma_884(acc[i][j], af[i], bf[j]);
    }
    if (ks + 1 < nk) {
      po_dg_store<PO_DG_BN>(a, Ps[cur ^ 1], Qs[cur ^ 1], tid, pv, qv);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = wr + 8 * i + g, col = wc + 8 * j + 2 * t4;
      C[row + ldc * (i64)col] = acc[i][j][0];
      C[row + ldc * (i64)(col + 1)] = acc[i][j][1];
    }
#else
  (void)a;
#endif
}

// cudaErrorNotSupported: shape / alignment not handled here (the caller keeps the generic dense kernel)
PO_XTU cudaError_t po_launch_dgemm(const void* A, i64 a_rs, i64 a_cs, const void* X, void* Y, i64 I, i64 n, i64 m, i64 O, cudaStream_t st) {
  int arm = po_dgemm_arm();
  if (!arm) return cudaErrorNotSupported;
#ifdef PO_EMUL
  if (arm == 2) arm = 1;  // mma.sync has no emulation: the SIMT arm (same tiles, loaders and epilogue layout rules) is what the CPU tests run
#endif
  if (n < PO_DG_BK || n % PO_DG_BK || m % PO_DG_BM || n > 0x7fffffff || m > 0x7fffffff) return cudaErrorNotSupported;
  if (((size_t)X % 16) || ((size_t)Y % 16)) return cudaErrorNotSupported;
  po_dg_args a;
  a.A = (const double*)A; a.a_rs = a_rs; a.a_cs = a_cs; a.X = (const double*)X; a.Y = (double*)Y;
  a.I = I; a.n = (int)n; a.m = (int)m; a.ncols = I * O;
  dim3 grid;
  const i64 BN = arm == 3 ? 64 : PO_DG_BN;  // arm 3: SIMT with 128 x 64 tiles, two CTAs per SM
  if (I == 1) {
    if (O % PO_DG_BN || (n % 2) || (m % 2)) return cudaErrorNotSupported;
    if (O / BN > 0x7fffffffLL) return cudaErrorNotSupported;
    a.x_is_p = 0;
    grid = dim3((unsigned)(O / BN), (unsigned)(m / PO_DG_BM), 1);
  } else {
    if (I % PO_DG_BM || m % PO_DG_BN || O > 65535) return cudaErrorNotSupported;
    a.x_is_p = 1;
    grid = dim3((unsigned)(I / PO_DG_BM), (unsigned)(m / BN), (unsigned)O);
  }
  if ((i64)grid.x * grid.y * grid.z < 64) return cudaErrorNotSupported;  // too few tiles to fill the GPU: the generic kernel's small tiles do better
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] dgemm %s arm: I=%lld n=%lld m=%lld O=%lld grid=%u x %u x %u\n", arm == 2 ? "DMMA" : "SIMT", (long long)I, (long long)n, (long long)m, (long long)O, grid.x, grid.y, grid.z);
  if (arm == 2) { PO_LAUNCH(po_dgemm_dmma_kernel, grid, dim3(PO_DG_THREADS), 0, st, a); }
  else if (arm == 3) { PO_LAUNCH(po_dgemm_simt_kernel<64>, grid, dim3(PO_DG_THREADS), 0, st, a); }
  else { PO_LAUNCH(po_dgemm_simt_kernel<128>, grid, dim3(PO_DG_THREADS), 0, st, a); }
  return cudaGetLastError();
}
#else
cudaError_t po_launch_dgemm(const void* A, i64 a_rs, i64 a_cs, const void* X, void* Y, i64 I, i64 n, i64 m, i64 O, cudaStream_t st);
#endif
