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
#include <stdlib.h>

#define PO_S2 0.70710678118654752440f
#define PO_C8 0.92387953251128673848f
#define PO_S8 0.38268343236508978178f

// Twiddles travel as a kernel parameter (constant bank, uniform loads).  The t-dim entries are
// pre-broadcast 2-wide with both signs so that the packed FP32x2 instructions (FFMA2/FADD2,
// sm_100) need no operand negation or register shuffling.
struct po_fused_tw {
  float2 tw_t[8 * 8 * 4];  // [(j2*8 + k)*4 + {0: (wx,wx), 1: (wy,wy), 2: (-wx,-wx), 3: (-wy,-wy)}], j2 < RT <= 8, k < 8
  float2 tw_x[8 * 16];     // [j2][r]: x-dim combine / pre-twiddle, j2 < RX <= 8, r < 16 (residue of the kept bin mod 16)
  float kscale[8];         // forward kernel only: per-t-mode output scale (1 for the transform; c_k/(nt*nx) for the
                           // true adjoint of the inverse pair used by the pullback)
};

// packed FP32x2 arithmetic: one FADD2 / FFMA2 / FMUL2 instruction per pair on sm_100
#ifdef PO_EMUL
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
PO_D float2 po_a2(float2 a, float2 b) { return __fadd2_rn(a, b); }
PO_D float2 po_s2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }  // a - b
PO_D float2 po_m2(float2 a, float2 b) { return __fmul2_rn(a, b); }
PO_D float2 po_f2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }  // a*b + c

#ifdef PO_EMUL
#define PO_SHFL_XOR(v, m) po_emul::shfl_xor((v), (m))
#else
#define PO_SHFL_XOR(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#endif

// ---- L2 residency hints (see po_fused2.cuh: po_policy_evict_first) ---------------------------------------
// Intermediates that the NEXT kernel of the plan re-reads (<= 84 MB, the L2 holds 126 MB) are written with an
// evict-last policy so that the 671 MB streams around them cannot push them out to DRAM.
PO_D unsigned long long po_policy_evict_last() {
  unsigned long long pol = 0;
#ifndef PO_EMUL
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
PO_D void po_st_hint(float2* p, float2 v, unsigned long long pol) {
#ifndef PO_EMUL
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#else
  (void)pol; *p = v;
#endif
}
PO_D void po_st_hint(float4* p, float4 v, unsigned long long pol) {
#ifndef PO_EMUL
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
#else
  (void)pol; *p = v;
#endif
}

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
      v[base + j] = po_a2(a, b);                                      // one FADD2 per complex add / sub (sm_100 packed FP32x2;
      v[base + j + HALF] = po_mul_w16<SIGN>(po_s2(a, b), j * STEP);   // a - b as fma(b, -1, a): the same rounding as a subtraction)
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

// Shared-memory layout of the 8x-reduced intermediate T[k][e] (k < 8 t-modes, e < L = I0*nx):
// "pair-planar" -- element pair (e, e+1), e even, is stored as {re_e, re_e+1, im_e, im_e+1}, so
// the packed t-phase moves whole register pairs with one 128-bit access and no shuffling.
PO_D int po_pp_re(int e) { return ((e >> 1) << 2) | (e & 1); }  // float index of Re; Im is +2

// ---- phase functions (shared by the simple and the pipelined kernels) --------------------------------
// forward t-phase: pruned real DFT along t on element PAIRS, global -> registers (packed FP32x2) -> shared
// Channel groups: when the whole (I0 x nx) row does not fit the shared-memory budget, a unit is
// one GROUP of IG (even) channels of a row: local element e = i + IG*x  <->  global i0 + i + I0*x.
// The other groups' bytes of the same 128 B lines are picked up from L2 by the sibling units
// that run at the same time, so DRAM traffic stays 1x.
PO_D int po_gpair(int v, int h2, int i0h, int I0h) { return i0h + (v % h2) + I0h * (v / h2); }  // float2 index in the row

template <int RT>
PO_D void po_fwd_tphase(const float2* __restrict__ xb, i64 rs2, float* __restrict__ T, int L, const po_fused_tw& tw, int t,
                        int nthreads, int IG, int i0, int I0) {
  for (int v = t; v < (L >> 1); v += nthreads) {
    const int gp = (IG == I0) ? v : po_gpair(v, IG >> 1, i0 >> 1, I0 >> 1);
    float2 ar[8], ai[8];  // accumulators: (elem0, elem1) of Re / Im of t-mode k
#pragma unroll(RT <= 4 ? RT : 4)
    for (int j2 = 0; j2 < RT; ++j2) {
      float2 x[8];
      const float2* q = xb + gp + rs2 * j2;
#pragma unroll
      for (int j1 = 0; j1 < 8; ++j1) { x[j1] = __ldg(q); q += rs2 * RT; }
      // real 8-point DFT, packed over the element pair (see po_rfft8)
      const float2 e0 = po_a2(x[0], x[4]), e1 = po_s2(x[0], x[4]), e2 = po_a2(x[2], x[6]), e3 = po_s2(x[2], x[6]);
      const float2 o0 = po_a2(x[1], x[5]), o1 = po_s2(x[1], x[5]), o2 = po_a2(x[3], x[7]), o3 = po_s2(x[3], x[7]);
      const float2 E0 = po_a2(e0, e2), E2 = po_s2(e0, e2), O0 = po_a2(o0, o2), O2 = po_s2(o0, o2);
      const float2 s2 = make_float2(PO_S2, PO_S2);
      const float2 pp = po_m2(po_s2(o1, o3), s2), qq = po_m2(po_a2(o1, o3), s2);
      // Y_k = (P_k, sg_k * Q_k):  k: 0 (E0+O0, 0)  4 (E0-O0, 0)  2 (E2,-O2)  6 (E2,+O2)
      //                           1 (e1+p, -(e3+q))  7 (e1+p, +(e3+q))  3 (e1-p, +(e3-q))  5 (e1-p, -(e3-q))
      const float2 P0 = po_a2(E0, O0), P4 = po_s2(E0, O0);
      const float2 A1 = po_a2(e1, pp), B1 = po_a2(e3, qq), A3 = po_s2(e1, pp), B3 = po_s2(e3, qq);
      if (j2 == 0) {
        const float2 zero = make_float2(0.f, 0.f), neg = make_float2(-1.f, -1.f);
        ar[0] = P0; ai[0] = zero; ar[4] = P4; ai[4] = zero;
        ar[2] = E2; ai[2] = po_m2(O2, neg); ar[6] = E2; ai[6] = O2;
        ar[1] = A1; ai[1] = po_m2(B1, neg); ar[7] = A1; ai[7] = B1;
        ar[3] = A3; ai[3] = B3; ar[5] = A3; ai[5] = po_m2(B3, neg);
      } else {
        const float2* w = tw.tw_t + j2 * 32;  // [k][4]
        // acc_k += w_k * Y_k ; re += wx*P - sg*wy*Q ; im += sg*wx*Q + wy*P
#define PO_ACC_POS(k, P, Q) \
  ar[k] = po_f2(P, w[(k)*4 + 0], ar[k]); ar[k] = po_f2(Q, w[(k)*4 + 3], ar[k]); \
  ai[k] = po_f2(Q, w[(k)*4 + 0], ai[k]); ai[k] = po_f2(P, w[(k)*4 + 1], ai[k]);
#define PO_ACC_NEG(k, P, Q) \
  ar[k] = po_f2(P, w[(k)*4 + 0], ar[k]); ar[k] = po_f2(Q, w[(k)*4 + 1], ar[k]); \
  ai[k] = po_f2(Q, w[(k)*4 + 2], ai[k]); ai[k] = po_f2(P, w[(k)*4 + 1], ai[k]);
        ar[0] = po_f2(P0, w[0], ar[0]); ai[0] = po_f2(P0, w[1], ai[0]);
        ar[4] = po_f2(P4, w[16], ar[4]); ai[4] = po_f2(P4, w[17], ai[4]);
        PO_ACC_NEG(2, E2, O2) PO_ACC_POS(6, E2, O2)
        PO_ACC_NEG(1, A1, B1) PO_ACC_POS(7, A1, B1)
        PO_ACC_POS(3, A3, B3) PO_ACC_NEG(5, A3, B3)
#undef PO_ACC_POS
#undef PO_ACC_NEG
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      *(float4*)(T + k * 2 * L + 4 * v) = make_float4(ar[k].x, ar[k].y, ai[k].x, ai[k].y);
  }
}

// forward x-phase: pruned complex DFT along x from shared; two adjacent lanes share one (i, k)
// pair, each runs RX/2 of the RX sub-FFTs and accumulates its twiddled outputs; one shuffle
// round combines.  `zu` = Z + I0*16*m + I0*16*Mid*8*b (the unit's output base), kstride = I0*16*Mid.
template <int RX>
PO_D void po_fwd_xphase(const float* __restrict__ T, const float2* __restrict__ twx, float2* __restrict__ zu, i64 kstride, int I0, int L,
                        int t, int nthreads, int IG, int i0) {
  const float* ksc = (const float*)(twx + 8 * 16);  // [8] per-t-mode output scale, staged behind the x twiddles
  const int nitems = IG * 8 * 2;
  for (int base = 0; base < nitems; base += nthreads) {
    if (base + (t & ~31) >= nitems) break;  // warp-uniform
    const bool active = base + t < nitems;
    const int item = active ? base + t : 0;
    const int p = item >> 1, half = item & 1;
    const int i = p % IG, k = p / IG;
    float2 o[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) o[r] = make_float2(0.f, 0.f);
    const float* tk = T + k * 2 * L;
#pragma unroll 1
    for (int s = 0; s < RX / 2; ++s) {
      const int j2 = half * (RX / 2) + s;
      float2 v[16];
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) {
        const int f = po_pp_re(i + IG * (RX * j1 + j2));
        v[j1] = make_float2(tk[f], tk[f + 2]);
      }
      po_fft16<-1>(v);
#pragma unroll
      for (int r = 0; r < 16; ++r) po_mac(o[r], v[po_brev4(r)], twx[j2 * 16 + r]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      o[r].x += PO_SHFL_XOR(o[r].x, 1);
      o[r].y += PO_SHFL_XOR(o[r].y, 1);
    }
    if (active) {
      float2* zb = zu + (i0 + i) + kstride * k;
      const float ks = ksc[k];
#pragma unroll
      for (int r = 0; r < 16; ++r)
        if ((r >> 3) == half) zb[I0 * r] = make_float2(o[r].x * ks, o[r].y * ks);
    }
  }
}

// inverse x-phase: zero-padded backward DFT along x, global -> shared; two lanes per (i, k) pair
template <int RX>
PO_D void po_inv_xphase(const float2* __restrict__ zu, i64 kstride, float* __restrict__ T, const float2* __restrict__ twx, int I0, int L,
                        int t, int nthreads, int IG, int i0) {
  const int nitems = IG * 8 * 2;
  for (int item = t; item < nitems; item += nthreads) {
    const int p = item >> 1, half = item & 1;
    const int i = p % IG, k = p / IG;
    const float2* zb = zu + (i0 + i) + kstride * k;
    float2 z[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) z[r] = __ldg(zb + I0 * r);
    float* tk = T + k * 2 * L;
#pragma unroll 1
    for (int s = 0; s < RX / 2; ++s) {
      const int j2 = half * (RX / 2) + s;
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = po_mul(z[r], twx[j2 * 16 + r]);
      po_fft16<1>(v);
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) {
        const int f = po_pp_re(i + IG * (RX * j1 + j2));
        tk[f] = v[po_brev4(j1)].x;
        tk[f + 2] = v[po_brev4(j1)].y;
      }
    }
  }
}

// inverse t-phase: zero-padded c2r backward DFT along t on element PAIRS, shared -> registers (packed) -> global
template <int RT>
PO_D void po_inv_tphase(const float* __restrict__ T, float2* __restrict__ yb, i64 rs2, int L, const po_fused_tw& tw, int t, int nthreads,
                        int IG, int i0, int I0) {
  for (int v = t; v < (L >> 1); v += nthreads) {
    const int gp = (IG == I0) ? v : po_gpair(v, IG >> 1, i0 >> 1, I0 >> 1);
    float2 ur[8], ui[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 u = *(const float4*)(T + k * 2 * L + 4 * v);
      ur[k] = make_float2(u.x, u.y);
      ui[k] = make_float2(u.z, u.w);
    }
#pragma unroll 2
    for (int j2 = 0; j2 < RT; ++j2) {
      const float2* w = tw.tw_t + j2 * 32;
      // V_k = w_k * U_k (only the parts the real-output codelet needs, see po_irfft8_re)
      const float2 H0 = po_m2(ur[0], w[0]);
      const float2 H4 = po_f2(ui[4], w[16 + 3], po_m2(ur[4], w[16 + 0]));
#define PO_VR(k) po_f2(ui[k], w[(k)*4 + 3], po_m2(ur[k], w[(k)*4 + 0]))
#define PO_VI(k) po_f2(ur[k], w[(k)*4 + 1], po_m2(ui[k], w[(k)*4 + 0]))
      const float2 r1 = po_a2(PO_VR(1), PO_VR(7)), i1 = po_s2(PO_VI(1), PO_VI(7));
      const float2 r2 = po_a2(PO_VR(2), PO_VR(6)), i2 = po_s2(PO_VI(2), PO_VI(6));
      const float2 r3 = po_a2(PO_VR(3), PO_VR(5)), i3 = po_s2(PO_VI(3), PO_VI(5));
#undef PO_VR
#undef PO_VI
      const float2 s2 = make_float2(PO_S2, PO_S2);
      const float2 a = po_a2(H0, H4), b = po_s2(H0, H4);
      const float2 p1 = po_m2(po_s2(r1, i1), s2), q1 = po_m2(po_a2(r1, i1), s2);
      const float2 p3 = po_m2(po_s2(r3, i3), s2), q3 = po_m2(po_a2(r3, i3), s2);
      const float2 ap = po_a2(a, r2), am = po_s2(a, r2), bp = po_a2(b, i2), bm = po_s2(b, i2);
      const float2 r13 = po_a2(r1, r3), i13 = po_s2(i1, i3), pq = po_s2(p1, q3), qp = po_s2(q1, p3);
      float2 out[8];
      out[0] = po_a2(ap, r13);
      out[4] = po_s2(ap, r13);
      out[2] = po_s2(am, i13);
      out[6] = po_a2(am, i13);
      out[1] = po_a2(bm, pq);
      out[5] = po_s2(bm, pq);
      out[3] = po_s2(bp, qp);
      out[7] = po_a2(bp, qp);
      float2* q = yb + gp + rs2 * j2;
#pragma unroll
      for (int j1 = 0; j1 < 8; ++j1) { *q = out[j1]; q += rs2 * RT; }
    }
  }
}

#ifdef PO_EMUL
#define PO_GRID_CONSTANT
#else
#define PO_GRID_CONSTANT __grid_constant__
#endif

// A unit = (channel group g, middle index m, batch b), g fastest.  Row = I0*nx floats, group row = IG*nx.
struct po_unit {
  i64 m, bb;
  int i0;
};
PO_D po_unit po_decode_unit(i64 unit, int G, int IG, i64 Mid) {
  po_unit u;
  u.i0 = (int)(unit % G) * IG;
  const i64 r = unit / G;
  u.m = r % Mid;
  u.bb = r / Mid;
  return u;
}

// the units of a launch restricted to the mid range [m_lo, m_lo + m_cnt)
PO_D po_unit po_decode_unit(i64 unit, int G, int IG, i64 m_lo, i64 m_cnt) {
  po_unit u = po_decode_unit(unit, G, IG, m_cnt);
  u.m += m_lo;
  return u;
}

// ---- simple kernels: one unit per CTA, 2 CTAs / SM (small problems) ------------------------------------
template <int RT, int RX, int NT>
__global__ void __launch_bounds__(NT, (NT == 256) ? 2 : 1)
po_fwd_tx_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused_tw tw, int I0, int IG, i64 Mid) {
  PO_DYN_SMEM(smem_raw);
  const int L = IG * 16 * RX, LR = I0 * 16 * RX;  // group / full row length (even)
  float* T = (float*)smem_raw;                 // [8][2L] floats, pair-planar
  float2* twx = (float2*)(T + 16 * (size_t)L);  // [8][16] + kscale
  const int tid = threadIdx.x;
  const po_unit u = po_decode_unit(blockIdx.x, I0 / IG, IG, Mid);
  const i64 rs2 = ((i64)LR * Mid) >> 1;  // row stride in float2
  const float2* xb = (const float2*)(X + (i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb);
  for (int e = tid; e < 8 * 16; e += NT) twx[e] = tw.tw_x[e];
  if (tid < 8) ((float*)(twx + 8 * 16))[tid] = tw.kscale[tid];
  po_fwd_tphase<RT>(xb, rs2, T, L, tw, tid, NT, IG, u.i0, I0);
  __syncthreads();
  po_fwd_xphase<RX>(T, twx, Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, I0, L, tid, NT, IG, u.i0);
}

template <int RT, int RX, int NT>
__global__ void __launch_bounds__(NT, (NT == 256) ? 2 : 1)
po_inv_xt_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused_tw tw, int I0, int IG, i64 Mid) {
  PO_DYN_SMEM(smem_raw);
  const int L = IG * 16 * RX, LR = I0 * 16 * RX;
  float* T = (float*)smem_raw;
  float2* twx = (float2*)(T + 16 * (size_t)L);
  const int tid = threadIdx.x;
  const po_unit u = po_decode_unit(blockIdx.x, I0 / IG, IG, Mid);
  const i64 rs2 = ((i64)LR * Mid) >> 1;
  float2* yb = (float2*)(Y + (i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb);
  for (int e = tid; e < 8 * 16; e += NT) twx[e] = tw.tw_x[e];
  __syncthreads();
  po_inv_xphase<RX>(Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, T, twx, I0, L, tid, NT, IG, u.i0);
  __syncthreads();
  po_inv_tphase<RT>(T, yb, rs2, L, tw, tid, NT, IG, u.i0, I0);
}

// ---- pipelined persistent kernels: one CTA per SM, 16 warps, warp-specialised ---------------------------
// The big streaming phase (t) and the small on-chip phase (x) of CONSECUTIVE units run
// concurrently on different warps with the 8x-reduced intermediate double-buffered in shared
// memory, so the HBM stream never pauses for the x-phase (the simple kernel leaves HBM idle
// whenever both co-resident CTAs are in their x-phase).
#define PO_PIPE_THREADS 512
#define PO_PIPE_STREAM 320  // warps 0..9 stream (t-phase), warps 10..15 do the x-phase

template <int RT, int RX>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_fwd_tx_pipe_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused_tw tw, int I0, int IG, i64 Mid,
                      i64 units, int nstream) {
  PO_DYN_SMEM(smem_raw);
  const int L = IG * 16 * RX, LR = I0 * 16 * RX, G = I0 / IG;
  float* T0 = (float*)smem_raw;  // two [8][2L] buffers
  float2* twx = (float2*)(T0 + 32 * (size_t)L);
  const int tid = threadIdx.x;
  const i64 rs2 = ((i64)LR * Mid) >> 1;
  for (int e = tid; e < 8 * 16; e += PO_PIPE_THREADS) twx[e] = tw.tw_x[e];
  if (tid < 8) ((float*)(twx + 8 * 16))[tid] = tw.kscale[tid];
  __syncthreads();
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;  // units of this CTA
  for (i64 it = 0; it <= nit; ++it) {
    if (tid < nstream) {
      if (it < nit) {
        const po_unit u = po_decode_unit(blockIdx.x + it * gridDim.x, G, IG, Mid);
        const float2* xb = (const float2*)(X + (i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb);
        po_fwd_tphase<RT>(xb, rs2, T0 + (it & 1) * 16 * (size_t)L, L, tw, tid, nstream, IG, u.i0, I0);
      }
    } else if (it > 0) {
      const po_unit u = po_decode_unit(blockIdx.x + (it - 1) * gridDim.x, G, IG, Mid);
      po_fwd_xphase<RX>(T0 + ((it - 1) & 1) * 16 * (size_t)L, twx, Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, I0, L,
                        tid - nstream, PO_PIPE_THREADS - nstream, IG, u.i0);
    }
    __syncthreads();
  }
}

template <int RT, int RX>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_inv_xt_pipe_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused_tw tw, int I0, int IG, i64 Mid,
                      i64 units, int nstream) {
  PO_DYN_SMEM(smem_raw);
  const int L = IG * 16 * RX, LR = I0 * 16 * RX, G = I0 / IG;
  float* T0 = (float*)smem_raw;
  float2* twx = (float2*)(T0 + 32 * (size_t)L);
  const int tid = threadIdx.x;
  const i64 rs2 = ((i64)LR * Mid) >> 1;
  for (int e = tid; e < 8 * 16; e += PO_PIPE_THREADS) twx[e] = tw.tw_x[e];
  __syncthreads();
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (i64 it = 0; it <= nit; ++it) {
    if (tid >= nstream) {  // x-phase of unit `it` fills the buffer ...
      if (it < nit) {
        const po_unit u = po_decode_unit(blockIdx.x + it * gridDim.x, G, IG, Mid);
        po_inv_xphase<RX>(Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, T0 + (it & 1) * 16 * (size_t)L, twx, I0, L,
                          tid - nstream, PO_PIPE_THREADS - nstream, IG, u.i0);
      }
    } else if (it > 0) {  // ... while the streaming warps drain the previous unit's buffer
      const po_unit u = po_decode_unit(blockIdx.x + (it - 1) * gridDim.x, G, IG, Mid);
      float2* yb = (float2*)(Y + (i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb);
      po_inv_tphase<RT>(T0 + ((it - 1) & 1) * 16 * (size_t)L, yb, rs2, L, tw, tid, nstream, IG, u.i0, I0);
    }
    __syncthreads();
  }
}

// channel-group size: the largest even divisor of I0 whose group row fits the shared-memory budget
static inline int po_fused_group(i64 I0, i64 nx, size_t budget_bytes, int bufs) {
  for (i64 ig = I0; ig >= 1; --ig) {
    if (I0 % ig) continue;
    if (ig != I0 && (ig & 1)) continue;  // float2 pairs must not straddle groups
    if ((I0 & 1) && ig != I0) continue;
    const size_t smem = (size_t)(8 * bufs * ig * nx + 16 * 8 + 4) * sizeof(float2);
    if (smem <= budget_bytes) return (int)ig;
  }
  return 0;
}

static inline bool po_fused_tx_shape_ok(i64 nt, i64 nx, i64 I0) {
  const bool t_ok = (nt == 16 || nt == 32 || nt == 64);
  const bool x_ok = (nx == 64 || nx == 128);
  return t_ok && x_ok && I0 >= 1 && po_fused_group(I0, nx, 100 * 1024, 1) > 0;
}

template <int RT, int RX>
static cudaError_t po_launch_fused_tx_t(bool inverse, const void* src, void* dst, const po_fused_tw& tw, int I0, i64 Mid, i64 B,
                                        int num_sms, cudaStream_t st) {
  const i64 nx = 16 * RX;
  int IGp = po_fused_group(I0, nx, 200 * 1024, 2);  // pipelined: two buffers, one CTA / SM
  int IGs = po_fused_group(I0, nx, 100 * 1024, 1);  // simple: one buffer, two CTAs / SM
  static const int fwd_whole = getenv("PO_FWD_WHOLE_ROWS") ? atoi(getenv("PO_FWD_WHOLE_ROWS")) : 0;  // A/B switch
  if ((inverse || fwd_whole) && IGp != I0) {
    // Channel groups make the streaming STORES partial-sector (40 B pieces of 80 B periods measured
    // 1.8 TB/s vs 4.4 TB/s for whole rows at c=20, nx=128): the inverse keeps whole rows, one CTA / SM.
    IGp = 0;
    IGs = po_fused_group(I0, nx, 200 * 1024, 1) == I0 ? I0 : IGs;
  }
  const bool pipe = num_sms > 0 && IGp > 0 && Mid * B * (I0 / IGp) >= 4 * (i64)num_sms;
  const int IG = pipe ? IGp : IGs;
  if (IG <= 0) return cudaErrorInvalidValue;
  const i64 L = (i64)IG * nx;
  const size_t smem = (size_t)(8 * (pipe ? 2 : 1) * L + 16 * 8 + 4) * sizeof(float2);
  const bool wide = !pipe && smem > 100 * 1024;  // one CTA / SM: run it with 16 warps instead of 8
  const i64 units = Mid * B * (I0 / IG);
  if (units == 0) return cudaSuccess;
#ifndef PO_EMUL
  cudaError_t e;
  if (pipe) {
    if (!inverse) e = cudaFuncSetAttribute(po_fwd_tx_pipe_kernel<RT, RX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else e = cudaFuncSetAttribute(po_inv_xt_pipe_kernel<RT, RX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  } else {
    if (!inverse && !wide) e = cudaFuncSetAttribute(po_fwd_tx_kernel<RT, RX, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else if (!inverse) e = cudaFuncSetAttribute(po_fwd_tx_kernel<RT, RX, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else if (!wide) e = cudaFuncSetAttribute(po_inv_xt_kernel<RT, RX, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else e = cudaFuncSetAttribute(po_inv_xt_kernel<RT, RX, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (e != cudaSuccess) return e;
#endif
  // warps that stream HBM (t-phase) vs warps that run the on-chip x-phase; multiples of 32 (tunable for A/B runs)
  static const int ns_fwd_env = getenv("PO_PIPE_STREAM_FWD") ? atoi(getenv("PO_PIPE_STREAM_FWD")) : 0;
  static const int ns_inv_env = getenv("PO_PIPE_STREAM_INV") ? atoi(getenv("PO_PIPE_STREAM_INV")) : 0;
  const int ns_fwd = ns_fwd_env ? ns_fwd_env : PO_PIPE_STREAM;
  const int ns_inv = ns_inv_env ? ns_inv_env : PO_PIPE_STREAM;
  if (pipe) {
    if (!inverse) {
      PO_LAUNCH((po_fwd_tx_pipe_kernel<RT, RX>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float*)src, (float2*)dst, tw, I0, IG, Mid, units, ns_fwd);
    } else {
      PO_LAUNCH((po_inv_xt_pipe_kernel<RT, RX>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float2*)src, (float*)dst, tw, I0, IG, Mid, units, ns_inv);
    }
  } else if (!inverse) {
    if (wide) { PO_LAUNCH((po_fwd_tx_kernel<RT, RX, 512>), dim3((unsigned)units), dim3(512), smem, st, (const float*)src, (float2*)dst, tw, I0, IG, Mid); }
    else { PO_LAUNCH((po_fwd_tx_kernel<RT, RX, 256>), dim3((unsigned)units), dim3(256), smem, st, (const float*)src, (float2*)dst, tw, I0, IG, Mid); }
  } else {
    if (wide) { PO_LAUNCH((po_inv_xt_kernel<RT, RX, 512>), dim3((unsigned)units), dim3(512), smem, st, (const float2*)src, (float*)dst, tw, I0, IG, Mid); }
    else { PO_LAUNCH((po_inv_xt_kernel<RT, RX, 256>), dim3((unsigned)units), dim3(256), smem, st, (const float2*)src, (float*)dst, tw, I0, IG, Mid); }
  }
  return cudaGetLastError();
}

static cudaError_t po_launch_fused_tx(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused_tw& tw, int I0,
                                      i64 Mid, i64 B, int num_sms, cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_FUSED_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_fused_tx_t<rt, rx>(inverse, src, dst, tw, I0, Mid, B, num_sms, st);
  PO_FUSED_CASE(2, 4) PO_FUSED_CASE(4, 4) PO_FUSED_CASE(8, 4)
  PO_FUSED_CASE(2, 8) PO_FUSED_CASE(4, 8) PO_FUSED_CASE(8, 8)
#undef PO_FUSED_CASE
  return cudaErrorInvalidValue;
}

// =======================================================================================
// Pruned complex DFT along a mode, 16 kept bins (0..7, n-8..n-1), n = 16*R, FP32:
//   forward  X (I, n, O) -> Y (I, 16, O)      R in-register FFT16 + twiddle-accumulate
//   inverse  X (I, 16, O) -> Y (I, n, O)      pre-twiddle + R in-register FFT16 (1/n folded in)
// One thread per column (i, o); consecutive threads = consecutive i, so every one of the n
// (or 16) row accesses of a warp is a contiguous 256 B segment.  Used for the y / z dims of
// the FNO stack (src/ParDFT.jl:26,30 fused with src/ParRestriction.jl:13-39).
// =======================================================================================
struct po_c16_tw {
  float2 tw[8 * 16];  // [j2][r]
};

template <int R, int INVERSE>
__global__ void __launch_bounds__(128)
po_c16_mode_kernel(const float2* __restrict__ X, float2* __restrict__ Y, const po_c16_tw tw, i64 I, i64 ncols, int keep, i64 o_lo, i64 o_cnt,
                   i64 o_period) {
  po_pdl_trigger();
  const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  po_pdl_wait();
  const unsigned long long pol = keep ? po_policy_evict_last() : 0ull;  // output is the next kernel's input: keep it in L2
  i64 o = col / I;
  const i64 i = col - o * I;
  if (o_period) {  // chunked launch: outer index = o_lo + zz + o_period * rep, zz < o_cnt (see po_run_inv_pair)
    const i64 rep = o / o_cnt;
    o = o_lo + (o - rep * o_cnt) + o_period * rep;
  }
  constexpr int N = 16 * R;
  if (!INVERSE) {
    const float2* xp = X + i + I * (i64)N * o;
    float2 acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int j2 = 0; j2 < R; ++j2) {
      float2 v[16];
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) v[j1] = __ldg(xp + I * (i64)(R * j1 + j2));
      po_fft16<-1>(v);
#pragma unroll
      for (int r = 0; r < 16; ++r) po_mac(acc[r], v[po_brev4(r)], tw.tw[j2 * 16 + r]);
    }
    float2* yp = Y + i + I * 16 * o;
#pragma unroll
    for (int r = 0; r < 16; ++r) { if (keep) po_st_hint(yp + I * r, acc[r], pol); else yp[I * r] = acc[r]; }
  } else {
    const float2* xp = X + i + I * 16 * o;
    float2 z[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) z[r] = __ldg(xp + I * r);
    float2* yp = Y + i + I * (i64)N * o;
#pragma unroll 1
    for (int j2 = 0; j2 < R; ++j2) {
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = po_mul(z[r], tw.tw[j2 * 16 + r]);
      po_fft16<1>(v);
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) { if (keep) po_st_hint(yp + I * (i64)(R * j1 + j2), v[po_brev4(j1)], pol); else yp[I * (i64)(R * j1 + j2)] = v[po_brev4(j1)]; }
    }
  }
}

// o_period > 0 restricts the launch to the outer indices {o_lo + zz + o_period * rep : zz < o_cnt, rep < O / o_period}
static cudaError_t po_launch_c16(bool inverse, int n, const void* src, void* dst, const po_c16_tw& tw, i64 I, i64 O, cudaStream_t st, i64 o_lo = 0,
                                 i64 o_cnt = 0, i64 o_period = 0) {
  static const int keep = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  if (o_period > 0 && (o_cnt <= 0 || o_lo < 0 || o_lo + o_cnt > o_period || O % o_period)) return cudaErrorInvalidValue;
  const i64 ncols = o_period > 0 ? I * o_cnt * (O / o_period) : I * O;
  if (ncols == 0) return cudaSuccess;
  const int R = n / 16;
  dim3 grid((unsigned)((ncols + 127) / 128)), block(128);
#define PO_C16_CASE(r)                                                                                                    \
  if (R == r) {                                                                                                           \
    if (!inverse) { PO_LAUNCH_PDL((po_c16_mode_kernel<r, 0>), grid, block, 0, st, (const float2*)src, (float2*)dst, tw, I, ncols, keep, o_lo, o_cnt, o_period); } \
    else { PO_LAUNCH_PDL((po_c16_mode_kernel<r, 1>), grid, block, 0, st, (const float2*)src, (float2*)dst, tw, I, ncols, keep, o_lo, o_cnt, o_period); } \
    return cudaGetLastError();                                                                                            \
  }
  PO_C16_CASE(2) PO_C16_CASE(4) PO_C16_CASE(8)
#undef PO_C16_CASE
  return cudaErrorInvalidValue;
}

// =======================================================================================
// Pruned REAL DFT along a mode, 16 kept low bins (0..15), n = 16*R, FP32 (config C2's (R o rfft) factor):
//   forward  real X (I, n, O) -> complex Y (I, 16, O): two residues share one complex FFT16 (Hermitian untangle)
//   inverse  complex X (I, 16, O) -> real Y (I, n, O): zero-padded c2r (weights c_b / n in the table), real part kept
// Same thread mapping as po_c16_mode_kernel (one thread per column, 128 B / 256 B row segments per warp).
// =======================================================================================
// acc += conj(a) * b
PO_D void po_mac_conj(float2& c, float2 a, float2 b) {
  c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(a.y, b.y, c.x);
  c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(-a.y, b.x, c.y);
}

// pruned real DFT of one column: acc[r] = sum_j x[j] exp(-2 pi i j r / n), r < 16, n = 16 R (the arithmetic of po_r16_mode_kernel and of
// the fused C2 kernel, po_c2.cuh).  x[R j1 + j2] = residue sequence j2; X[r] = sum_j2 W_n^(j2 r) F_j2[r] with F_j2 the 16-point DFT of
// residue j2.  Two REAL residues a, b share one complex FFT16 of a + i b; their spectra are Hermitian, so only r = 0..8 are untangled
// (A = Z[r] + conj Z[16 - r] = 2 F_a[r], B = -i (Z[r] - conj Z[16 - r]) = 2 F_b[r]) and r = 9..15 reuse the conjugates; the factor 1/2 is
// applied once at the end.  The rows of residue pair jp + 1 are loaded BEFORE pair jp is transformed, so 32 loads per thread are always
// in flight (the one-shot version alternated "32 outstanding" with "none while it computes": 3.5 TB/s at C2).
// IC > 0: the row pitch I is the compile-time constant IC, so all n row addresses are immediates off one base pointer (the generic form
// spends 500 of its 2400 instructions per column on 64-bit address arithmetic); IC == 0: run-time pitch.
template <int R, int IC = 0>
PO_D void po_r16_fwd_column(const float* __restrict__ xp, i64 I_rt, const po_c16_tw& tw, float2 (&acc)[16]) {
  static_assert(R % 2 == 0, "residues are processed in pairs");
  const i64 I = IC ? (i64)IC : I_rt;
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = make_float2(0.f, 0.f);
  float2 nxt[16];
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) nxt[j1] = make_float2(__ldg(xp + I * (i64)(R * j1)), __ldg(xp + I * (i64)(R * j1 + 1)));
#pragma unroll
  for (int jp = 0; jp < R / 2; ++jp) {
    const int j2a = 2 * jp, j2b = j2a + 1;
    float2 v[16];
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) v[j1] = nxt[j1];
    if (jp + 1 < R / 2) {
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) nxt[j1] = make_float2(__ldg(xp + I * (i64)(R * j1 + j2a + 2)), __ldg(xp + I * (i64)(R * j1 + j2b + 2)));
    }
    po_fft16<-1>(v);
#pragma unroll
    for (int r = 0; r <= 8; ++r) {
      const float2 zb = v[po_brev4(r)], zn = v[po_brev4((16 - r) & 15)];
      const float2 A = make_float2(zb.x + zn.x, zb.y - zn.y);
      const float2 B = make_float2(zb.y + zn.y, zn.x - zb.x);
      po_mac(acc[r], A, tw.tw[j2a * 16 + r]);
      po_mac(acc[r], B, tw.tw[j2b * 16 + r]);
      if (r >= 1 && r <= 7) {
        po_mac_conj(acc[16 - r], A, tw.tw[j2a * 16 + 16 - r]);
        po_mac_conj(acc[16 - r], B, tw.tw[j2b * 16 + 16 - r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = po_scale(acc[r], 0.5f);
}

template <int R, int INVERSE>
__global__ void __launch_bounds__(128)
po_r16_mode_kernel(const void* __restrict__ Xv, void* __restrict__ Yv, const po_c16_tw tw, i64 I, i64 ncols, int keep) {
  const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const i64 o = col / I, i = col - o * I;
  constexpr int N = 16 * R;
  const unsigned long long pol = keep ? po_policy_evict_last() : 0ull;
  if (!INVERSE) {
    float2 acc[16];
    po_r16_fwd_column<R>((const float*)Xv + i + I * (i64)N * o, I, tw, acc);
    float2* yp = (float2*)Yv + i + I * 16 * o;
#pragma unroll
    for (int r = 0; r < 16; ++r) { if (keep) po_st_hint(yp + I * r, acc[r], pol); else yp[I * r] = acc[r]; }
  } else {
    const float2* xp = (const float2*)Xv + i + I * 16 * o;
    float2 z[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) z[r] = __ldg(xp + I * r);
    float* yp = (float*)Yv + i + I * (i64)N * o;
#pragma unroll 1
    for (int j2 = 0; j2 < R; ++j2) {
      float2 v[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = po_mul(z[r], tw.tw[j2 * 16 + r]);
      po_fft16<1>(v);
#pragma unroll
      for (int j1 = 0; j1 < 16; ++j1) yp[I * (i64)(R * j1 + j2)] = v[po_brev4(j1)].x;
    }
  }
}

static cudaError_t po_launch_r16(bool inverse, int n, const void* src, void* dst, const po_c16_tw& tw, i64 I, i64 O, bool keep, cudaStream_t st) {
  static const int hints = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  const int kp = (keep && hints) ? 1 : 0;
  const i64 ncols = I * O;
  if (ncols == 0) return cudaSuccess;
  const int R = n / 16;
  dim3 grid((unsigned)((ncols + 127) / 128)), block(128);
#define PO_R16_CASE(r)                                                                                          \
  if (R == r) {                                                                                                 \
    if (!inverse) { PO_LAUNCH((po_r16_mode_kernel<r, 0>), grid, block, 0, st, src, dst, tw, I, ncols, kp); }    \
    else { PO_LAUNCH((po_r16_mode_kernel<r, 1>), grid, block, 0, st, src, dst, tw, I, ncols, 0); }              \
    return cudaGetLastError();                                                                                  \
  }
  PO_R16_CASE(2) PO_R16_CASE(4) PO_R16_CASE(8)
#undef PO_R16_CASE
  return cudaErrorInvalidValue;
}

template <int SIGN> PO_D void po_fft8(float2 (&v)[8]);
PO_HD constexpr int po_brev3(int r);
// =======================================================================================
// Full complex FFT along a mode, n = 16*R2 (32 / 64 / 128), FP32, ONE pass through shared memory:
//   thread (c, j2), j2 < R2: residue-j2 FFT16 in registers (16 strided loads straight from global), twiddle
//   W_n^{j2 k1}, write S[c][k1][j2];  sync;  thread (c, t): for k1 in {t, t + R2, ...}: DFT_R2 over j2 -> X[k1 + 16 k2].
// The generic radix-2 Stockham kernel (po_fft_mode_kernel) makes log2(n) trips through shared memory and reaches
// 1.1 TB/s at C2; this one makes one.  c2c only, no restriction maps (those go through the pruned kernels).
// =======================================================================================
template <int R2>
PO_D void po_small_dft(float2 (&v)[R2], bool inverse) {
  if (R2 == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = po_add(a, b); v[1] = po_sub(a, b);
  } else if (R2 == 4) {
    const float2 a = po_add(v[0], v[2]), b = po_sub(v[0], v[2]), c = po_add(v[1], v[3]), d = po_sub(v[1], v[3]);
    const float2 jd = inverse ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // (+/-) i * d  ->  forward uses -i
    v[0] = po_add(a, c); v[2] = po_sub(a, c); v[1] = po_add(b, jd); v[3] = po_sub(b, jd);
  }
}
template <int R2, int INVERSE>
__global__ void __launch_bounds__(256)
po_cfull_mode_kernel(const float2* __restrict__ X, float2* __restrict__ Y, const po_c16_tw tw, i64 I, i64 ncols, float scale) {
  constexpr int N = 16 * R2, TC = 256 / R2;  // columns per CTA
  __shared__ float2 S[TC * (N + 1)];
  const int tid = threadIdx.x;
  const int c = (I > 1) ? tid % TC : tid / R2, j2 = (I > 1) ? tid / TC : tid % R2;
  const i64 col = (i64)blockIdx.x * TC + c;
  const bool ok = col < ncols;
  const i64 o = ok ? col / I : 0, i = ok ? col - o * I : 0;
  if (ok) {
    const float2* xp = X + i + I * (i64)N * o;
    float2 v[16];
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) v[j1] = __ldg(xp + I * (i64)(R2 * j1 + j2));
    if (INVERSE) po_fft16<1>(v); else po_fft16<-1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) S[c * (N + 1) + k1 * R2 + j2] = po_mul(v[po_brev4(k1)], tw.tw[j2 * 16 + k1]);
  }
  __syncthreads();
  if (ok) {
    float2* yp = Y + i + I * (i64)N * o;
#pragma unroll
    for (int q = 0; q < 16 / R2; ++q) {
      const int k1 = j2 + R2 * q;
      if (R2 == 8) {
        float2 u[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) u[t] = S[c * (N + 1) + k1 * 8 + t];
        if (INVERSE) po_fft8<1>(u); else po_fft8<-1>(u);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) yp[I * (i64)(k1 + 16 * k2)] = po_scale(u[po_brev3(k2)], scale);
      } else {
        float2 u[R2];
#pragma unroll
        for (int t = 0; t < R2; ++t) u[t] = S[c * (N + 1) + k1 * R2 + t];
        po_small_dft<R2>(u, INVERSE != 0);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) yp[I * (i64)(k1 + 16 * k2)] = po_scale(u[k2], scale);
      }
    }
  }
}

static cudaError_t po_launch_cfull(bool inverse, int n, const void* src, void* dst, const po_c16_tw& tw, i64 I, i64 O, float scale, cudaStream_t st) {
  const i64 ncols = I * O;
  if (ncols == 0) return cudaSuccess;
  const int R2 = n / 16;
  const int TC = 256 / R2;
  dim3 grid((unsigned)((ncols + TC - 1) / TC)), block(256);
#define PO_CFULL_CASE(r)                                                                                                   \
  if (R2 == r) {                                                                                                           \
    if (!inverse) { PO_LAUNCH((po_cfull_mode_kernel<r, 0>), grid, block, 0, st, (const float2*)src, (float2*)dst, tw, I, ncols, scale); } \
    else { PO_LAUNCH((po_cfull_mode_kernel<r, 1>), grid, block, 0, st, (const float2*)src, (float2*)dst, tw, I, ncols, scale); }        \
    return cudaGetLastError();                                                                                             \
  }
  PO_CFULL_CASE(2) PO_CFULL_CASE(4) PO_CFULL_CASE(8)
#undef PO_CFULL_CASE
  return cudaErrorInvalidValue;
}

// =======================================================================================
// Channel mixing fused with the per-mode diagonal: the frequency weighting
// `ParDiagonal ⊗ ParMatrix` of examples/fno.jl:25-27 (and its adjoint) as ONE pass:
//     Y[k, col] = d(col % nd) * sum_j A(k, j) X[j, col]         (I = 1: columns are contiguous)
// A(k, j) = A[k*a_rs + j*a_cs] (conj optional: theta'), d conj optional.  Small m, n (channels).
// =======================================================================================
template <class T>
__global__ void __launch_bounds__(256)
po_mix_diag_kernel(const T* __restrict__ A, i64 a_rs, i64 a_cs, int conj_a, const T* __restrict__ d, int conj_d, i64 nd,
                   const T* __restrict__ X, T* __restrict__ Y, int n, int m, i64 ncols, int tcol) {
  PO_DYN_SMEM(smem_raw);
  T* As = (T*)smem_raw;              // [n][m]: consecutive k (lanes) hit consecutive banks
  T* xs = As + (size_t)m * n;        // [tcol][n]
  const int tid = threadIdx.x;
  for (int e = tid; e < m * n; e += 256) {
    const int k = e / n, j = e - k * n;
    T v = A[(i64)k * a_rs + (i64)j * a_cs];
    As[j * m + k] = conj_a ? po_conj(v) : v;
  }
  const i64 col0 = (i64)blockIdx.x * tcol;
  const int nc = (int)((ncols - col0 < tcol) ? (ncols - col0) : tcol);
  const T* xp = X + col0 * n;
  for (int e = tid; e < nc * n; e += 256) xs[e] = xp[e];
  __syncthreads();
  T* yp = Y + col0 * m;
  for (int e = tid; e < nc * m; e += 256) {
    const int c = e / m, k = e - c * m;
    T acc = po_zero<T>();
    const T* a = As + k;
    const T* x = xs + c * n;
    for (int j = 0; j < n; ++j) po_mac(acc, a[j * m], x[j]);
    if (d) {
      T w = d[(col0 + c) % nd];
      if (conj_d) w = po_conj(w);
      acc = po_mul(w, acc);
    }
    yp[e] = acc;
  }
}

static cudaError_t po_launch_mix_diag(int dt, const void* A, i64 a_rs, i64 a_cs, int conj_a, const void* d, int conj_d, i64 nd,
                                      const void* X, void* Y, i64 n, i64 m, i64 ncols, cudaStream_t st) {
  if (ncols == 0) return cudaSuccess;
  const int tcol = 64;
  const size_t esz = po_dtype_size(dt);
  const size_t smem = (size_t)(m * n + tcol * n) * esz;
  dim3 grid((unsigned)((ncols + tcol - 1) / tcol)), block(256);
  switch (dt) {
    case PO_F32: PO_LAUNCH((po_mix_diag_kernel<float>), grid, block, smem, st, (const float*)A, a_rs, a_cs, conj_a, (const float*)d, conj_d, nd, (const float*)X, (float*)Y, (int)n, (int)m, ncols, tcol); break;
    case PO_F64: PO_LAUNCH((po_mix_diag_kernel<double>), grid, block, smem, st, (const double*)A, a_rs, a_cs, conj_a, (const double*)d, conj_d, nd, (const double*)X, (double*)Y, (int)n, (int)m, ncols, tcol); break;
    case PO_C64: PO_LAUNCH((po_mix_diag_kernel<float2>), grid, block, smem, st, (const float2*)A, a_rs, a_cs, conj_a, (const float2*)d, conj_d, nd, (const float2*)X, (float2*)Y, (int)n, (int)m, ncols, tcol); break;
    case PO_C128: PO_LAUNCH((po_mix_diag_kernel<double2>), grid, block, smem, st, (const double2*)A, a_rs, a_cs, conj_a, (const double2*)d, conj_d, nd, (const double2*)X, (double2*)Y, (int)n, (int)m, ncols, tcol); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- bulk-copy / mbarrier primitives (used by the asynchronous kernels of po_fused3.cuh, po_fused4.cuh, po_tensor.cuh) and small codelets ----
#ifndef PO_EMUL
PO_D unsigned po_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PO_D void po_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(po_smem_u32(bar)), "r"(count) : "memory");
}
PO_D void po_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(po_smem_u32(bar)), "r"(bytes) : "memory");
}
PO_D void po_bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(po_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(po_smem_u32(bar))
               : "memory");
}
PO_D bool po_mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(po_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
#endif

template <int SIGN>
PO_D void po_fft8(float2 (&v)[8]) {  // DIF, natural input, v[pos] = X[bitrev3(pos)]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = v[j], b = v[j + 4];
    v[j] = make_float2(a.x + b.x, a.y + b.y);
    v[j + 4] = po_mul_w16<SIGN>(make_float2(a.x - b.x, a.y - b.y), 2 * j);
  }
#pragma unroll
  for (int base = 0; base < 8; base += 4)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 a = v[base + j], b = v[base + j + 2];
      v[base + j] = make_float2(a.x + b.x, a.y + b.y);
      v[base + j + 2] = po_mul_w16<SIGN>(make_float2(a.x - b.x, a.y - b.y), 4 * j);
    }
#pragma unroll
  for (int base = 0; base < 8; base += 2) {
    const float2 a = v[base], b = v[base + 1];
    v[base] = make_float2(a.x + b.x, a.y + b.y);
    v[base + 1] = make_float2(a.x - b.x, a.y - b.y);
  }
}
PO_HD constexpr int po_brev3(int r) { return ((r & 1) << 2) | (r & 2) | ((r & 4) >> 2); }

