// Shape-specialised fused kernels for the FNO spectral transform (the hot path's two big
// passes):
//
//   po_fwd_tx_kernel : real X[I0, nx, Mid, nt, B]  ->  complex Z[I0, 16, Mid, 8, B]
//       = (R[1:8] ∘ rfft_nt) on the slowest dim  fused with  (R[1:8, nx-7:nx] ∘ fft_nx) on the
//       second-fastest dim: ONE read of the full-size real tensor, an 8x smaller write.
//   po_inv_xt_kernel : the mirror image (zero-fill + ifft_nx, zero-fill + irfft_nt): an 8x
//       smaller read, ONE write of the full-size real tensor.
//
// The reference does each of these with 2 library FFTs over the untruncated tensor, 2
// restriction copies and ~6 materialised permutedims (src/ParKron.jl:199-213).
//
// Arithmetic: pruned decimation.  For an N = R*K point DFT of which only K contiguous (mod N)
// bins are kept, write j = R*j1 + j2:
//     X[b] = sum_{j2<R} W_N^{b*j2} * ( K-point FFT over j1 of x[R*j1 + j2] )[b mod K]
// i.e. R small in-register FFTs (K = 8 real / 16 complex codelets) + an N-term twiddle combine,
// ~2.5x (t) and ~4x (x) fewer flops than the dense DFT-matrix contraction, in plain FP32 FMA
// (no TF32 splitting needed for the 1e-5 parity bar).  HBM-bound by design: every input byte is
// read once with 128 B-coalesced loads straight into registers, the t-contraction never touches
// shared memory, and only the 8x-reduced intermediate is transposed through shared memory.
#pragma once
#include "po_common.cuh"

#define PO_S2 0.70710678118654752440f
#define PO_C8 0.92387953251128673848f
#define PO_S8 0.38268343236508978178f

struct po_fused_tw {
  float2 tw_t[8 * 8];   // [j2][k]: t-dim combine / pre-twiddle, j2 < RT <= 8, k < 8
  float2 tw_x[8 * 16];  // [j2][r]: x-dim combine / pre-twiddle, j2 < RX <= 8, r < 16 (residue of the kept bin mod 16)
};

#ifdef PO_EMUL
#define PO_SHFL_XOR(v, m) po_emul::shfl_xor((v), (m))
#else
#define PO_SHFL_XOR(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#endif

// ---- codelets ---------------------------------------------------------------------------
// real 8-point forward DFT, Y[k] = sum_j v[j] exp(-2 pi i j k / 8)
PO_D void po_rfft8(const float (&v)[8], float2 (&Y)[8]) {
  const float e0 = v[0] + v[4], e1 = v[0] - v[4], e2 = v[2] + v[6], e3 = v[2] - v[6];
  const float o0 = v[1] + v[5], o1 = v[1] - v[5], o2 = v[3] + v[7], o3 = v[3] - v[7];
  const float E0 = e0 + e2, E2 = e0 - e2, O0 = o0 + o2, O2 = o0 - o2;
  const float p = (o1 - o3) * PO_S2, q = (o1 + o3) * PO_S2;
  Y[0] = make_float2(E0 + O0, 0.f);
  Y[4] = make_float2(E0 - O0, 0.f);
  Y[2] = make_float2(E2, -O2);
  Y[6] = make_float2(E2, O2);
  Y[1] = make_float2(e1 + p, -e3 - q);
  Y[7] = make_float2(e1 + p, e3 + q);
  Y[3] = make_float2(e1 - p, e3 - q);
  Y[5] = make_float2(e1 - p, q - e3);
}

// real part of the 8-point backward DFT: out[j] = Re sum_k V[k] exp(+2 pi i j k / 8)
PO_D void po_irfft8_re(const float2 (&V)[8], float (&out)[8]) {
  const float H0 = V[0].x, H4 = V[4].x;
  const float r1 = V[1].x + V[7].x, i1 = V[1].y - V[7].y;
  const float r2 = V[2].x + V[6].x, i2 = V[2].y - V[6].y;
  const float r3 = V[3].x + V[5].x, i3 = V[3].y - V[5].y;
  const float a = H0 + H4, b = H0 - H4;
  const float p1 = (r1 - i1) * PO_S2, q1 = (r1 + i1) * PO_S2, p3 = (r3 - i3) * PO_S2, q3 = (r3 + i3) * PO_S2;
  out[0] = a + r2 + r1 + r3;
  out[1] = b - i2 + p1 - q3;
  out[2] = a - r2 - i1 + i3;
  out[3] = b + i2 - q1 + p3;
  out[4] = a + r2 - r1 - r3;
  out[5] = b - i2 - p1 + q3;
  out[6] = a - r2 + i1 - i3;
  out[7] = b + i2 + q1 - p3;
}

// a * exp(SIGN * 2 pi i k / 16), k in 0..7 known at compile time after unrolling
template <int SIGN>
PO_D float2 po_mul_w16(float2 a, const int k) {
  if (k == 0) return a;
  if (k == 4) return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
  const float wr = (k == 1) ? PO_C8 : (k == 2) ? PO_S2 : (k == 3) ? PO_S8 : (k == 5) ? -PO_S8 : (k == 6) ? -PO_S2 : -PO_C8;
  const float ws = (k == 1) ? PO_S8 : (k == 2) ? PO_S2 : (k == 3) ? PO_C8 : (k == 5) ? PO_C8 : (k == 6) ? PO_S2 : PO_S8;
  const float wi = (SIGN < 0) ? -ws : ws;
  return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}

template <int SIGN, int LEN>
PO_D void po_fft16_stage(float2 (&v)[16]) {
  constexpr int HALF = LEN / 2, STEP = 16 / LEN;
#pragma unroll
  for (int base = 0; base < 16; base += LEN) {
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
      const float2 a = v[base + j], b = v[base + j + HALF];
      v[base + j] = make_float2(a.x + b.x, a.y + b.y);
      v[base + j + HALF] = po_mul_w16<SIGN>(make_float2(a.x - b.x, a.y - b.y), j * STEP);
    }
  }
}

// 16-point complex DFT (decimation in frequency): natural-order input, v[pos] = X[bitrev4(pos)] on return
template <int SIGN>
PO_D void po_fft16(float2 (&v)[16]) {
  po_fft16_stage<SIGN, 16>(v);
  po_fft16_stage<SIGN, 8>(v);
  po_fft16_stage<SIGN, 4>(v);
  po_fft16_stage<SIGN, 2>(v);
}

PO_HD constexpr int po_brev4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r & 8) >> 3); }

// ---- forward: t (real, nt = 8*RT -> bins 0..7) then x (nx = 16*RX -> bins 0..7, nx-8..nx-1) --------
template <int RT, int RX>
__global__ void __launch_bounds__(256, 2)
po_fwd_tx_kernel(const float* __restrict__ X, float2* __restrict__ Z, const po_fused_tw tw, int I0, i64 Mid) {
  PO_DYN_SMEM(smem_raw);
  const int L = I0 * 16 * RX;
  float2* T = (float2*)smem_raw;  // [8][L]
  float2* twx = T + 8 * (size_t)L;  // [RX][16]
  const int tid = threadIdx.x;
  const i64 unit = blockIdx.x;
  const i64 m = unit % Mid, bb = unit / Mid;
  const i64 row_stride = (i64)L * Mid;
  const float* xb = X + (i64)L * m + row_stride * (i64)(8 * RT) * bb;
  for (int e = tid; e < RX * 16; e += 256) twx[e] = tw.tw_x[e];

  // phase 1: pruned real DFT along t, global -> registers -> shared
  for (int e = tid; e < L; e += 256) {
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll(RT <= 4 ? RT : 2)
    for (int j2 = 0; j2 < RT; ++j2) {
      float v[8];
#pragma unroll
      for (int j1 = 0; j1 < 8; ++j1) v[j1] = __ldg(xb + e + row_stride * (i64)(RT * j1 + j2));
      float2 Y[8];
      po_rfft8(v, Y);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (j2 == 0) {
          acc[k].x += Y[k].x;
          acc[k].y += Y[k].y;
        } else {
          const float2 w = tw.tw_t[j2 * 8 + k];
          acc[k].x = fmaf(w.x, Y[k].x, acc[k].x);
          acc[k].x = fmaf(-w.y, Y[k].y, acc[k].x);
          acc[k].y = fmaf(w.x, Y[k].y, acc[k].y);
          acc[k].y = fmaf(w.y, Y[k].x, acc[k].y);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) T[k * L + e] = acc[k];
  }
  __syncthreads();

  // phase 2: pruned complex DFT along x; RX adjacent lanes share one (i, k) pair
  const int nitems = I0 * 8 * RX;  // multiple of 32 (RX >= 4)
  for (int base = 0; base < nitems; base += 256) {
    if (base + (tid & ~31) >= nitems) break;  // warp-uniform
    const int item = base + tid;
    const int p = item / RX, j2 = item % RX;
    const int i = p % I0, k = p / I0;
    float2 v[16];
    const float2* tp = T + k * L + i + I0 * j2;
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) v[j1] = tp[I0 * RX * j1];
    po_fft16<-1>(v);
    float2 o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) o[r] = po_mul(v[po_brev4(r)], twx[j2 * 16 + r]);
#pragma unroll
    for (int s = 1; s < RX; s <<= 1) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        o[r].x += PO_SHFL_XOR(o[r].x, s);
        o[r].y += PO_SHFL_XOR(o[r].y, s);
      }
    }
    float2* zb = Z + i + (i64)I0 * 16 * (m + Mid * ((i64)k + 8 * bb));
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r / (16 / RX) == j2) zb[I0 * r] = o[r];
  }
}

// ---- inverse: x (bins 0..7, nx-8..nx-1 -> nx = 16*RX) then t (bins 0..7 -> real nt = 8*RT) -----------
// All scaling (1/nx, 1/nt, Hermitian doubling c_k) is folded into tw.tw_t.
template <int RT, int RX>
__global__ void __launch_bounds__(256, 2)
po_inv_xt_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const po_fused_tw tw, int I0, i64 Mid) {
  PO_DYN_SMEM(smem_raw);
  const int L = I0 * 16 * RX;
  float2* T = (float2*)smem_raw;  // [8][L]
  float2* twx = T + 8 * (size_t)L;
  const int tid = threadIdx.x;
  const i64 unit = blockIdx.x;
  const i64 m = unit % Mid, bb = unit / Mid;
  const i64 row_stride = (i64)L * Mid;
  float* yb = Y + (i64)L * m + row_stride * (i64)(8 * RT) * bb;
  for (int e = tid; e < RX * 16; e += 256) twx[e] = tw.tw_x[e];
  __syncthreads();

  // phase A: zero-padded backward DFT along x -> shared
  const int nitems = I0 * 8 * RX;
  for (int base = 0; base < nitems; base += 256) {
    const int item = base + tid;
    if (item >= nitems) break;
    const int p = item / RX, j2 = item % RX;
    const int i = p % I0, k = p / I0;
    const float2* zb = Z + i + (i64)I0 * 16 * (m + Mid * ((i64)k + 8 * bb));
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = po_mul(zb[I0 * r], twx[j2 * 16 + r]);
    po_fft16<1>(v);
    float2* tp = T + k * L + i + I0 * j2;
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) tp[I0 * RX * j1] = v[po_brev4(j1)];
  }
  __syncthreads();

  // phase B: zero-padded c2r backward DFT along t, shared -> registers -> global
  for (int e = tid; e < L; e += 256) {
    float2 U[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) U[k] = T[k * L + e];
#pragma unroll(RT <= 4 ? RT : 2)
    for (int j2 = 0; j2 < RT; ++j2) {
      float2 V[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) V[k] = po_mul(U[k], tw.tw_t[j2 * 8 + k]);
      float out[8];
      po_irfft8_re(V, out);
#pragma unroll
      for (int j1 = 0; j1 < 8; ++j1) yb[e + row_stride * (i64)(RT * j1 + j2)] = out[j1];
    }
  }
}

static inline bool po_fused_tx_shape_ok(i64 nt, i64 nx, i64 I0) {
  const bool t_ok = (nt == 16 || nt == 32 || nt == 64);
  const bool x_ok = (nx == 64 || nx == 128);
  const i64 L = I0 * nx;
  const size_t smem = (size_t)(8 * L + 16 * 8) * sizeof(float2);
  return t_ok && x_ok && I0 >= 1 && smem <= 200 * 1024;
}

template <int RT, int RX>
static cudaError_t po_launch_fused_tx_t(bool inverse, const void* src, void* dst, const po_fused_tw& tw, int I0, i64 Mid, i64 B,
                                        cudaStream_t st) {
  const i64 L = (i64)I0 * 16 * RX;
  const size_t smem = (size_t)(8 * L + 16 * RX) * sizeof(float2);
  const i64 units = Mid * B;
  if (units == 0) return cudaSuccess;
#ifndef PO_EMUL
  cudaError_t e;
  if (!inverse) e = cudaFuncSetAttribute(po_fwd_tx_kernel<RT, RX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  else e = cudaFuncSetAttribute(po_inv_xt_kernel<RT, RX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
#endif
  if (!inverse) {
    PO_LAUNCH((po_fwd_tx_kernel<RT, RX>), dim3((unsigned)units), dim3(256), smem, st, (const float*)src, (float2*)dst, tw, I0, Mid);
  } else {
    PO_LAUNCH((po_inv_xt_kernel<RT, RX>), dim3((unsigned)units), dim3(256), smem, st, (const float2*)src, (float*)dst, tw, I0, Mid);
  }
  return cudaGetLastError();
}

static cudaError_t po_launch_fused_tx(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused_tw& tw, int I0,
                                      i64 Mid, i64 B, cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_FUSED_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_fused_tx_t<rt, rx>(inverse, src, dst, tw, I0, Mid, B, st);
  PO_FUSED_CASE(2, 4) PO_FUSED_CASE(4, 4) PO_FUSED_CASE(8, 4)
  PO_FUSED_CASE(2, 8) PO_FUSED_CASE(4, 8) PO_FUSED_CASE(8, 8)
#undef PO_FUSED_CASE
  return cudaErrorInvalidValue;
}
