// Second-generation phases + pipelined kernels for the fused (t, x) truncated transforms
// (same contract as po_fwd_tx_pipe_kernel / po_inv_xt_pipe_kernel in po_fused.cuh).
//
// The first-generation kernels are issue-bound, not HBM-bound: ncu (profiles/r1f_*) shows 14.5
// thread-instructions per input float, 34-39 % of them integer/address work, the x-phase in
// scalar FP32 and the t-phase twiddles re-loaded from the constant bank.  This file removes
// instructions instead of adding occupancy:
//   * t-phase twiddles are compile-time immediates of FFMA2 (nt is a template parameter), the
//     j2 loop is fully unrolled and trivial twiddles (1, -i) are elided at compile time;
//   * the x-phase runs PACKED: one thread transforms two adjacent channels at once (FP32x2
//     FADD2/FFMA2 on the pair-planar shared-memory layout, 128-bit LDS/STS), one lane per
//     residue j2; the RX partial spectra are combined through the thread's own (already
//     consumed) slots of the shared-memory tile instead of warp shuffles;
//   * the per-t-mode scaling (1/(nt nx), Hermitian doubling, pullback variants) lives in a small
//     shared-memory twiddle table of the x-phase, so the t-phase needs no scale at all;
//   * whole-row units (the common case) are a template flag: no divisions in the inner loops.
// Requires an even channel-group size IG (pairs are two channels at the same x); odd channel
// counts keep the first-generation kernels.
#pragma once
#include <stdio.h>
#include "po_fused.cuh"

struct po_fused2_tw {
  float2 tw_x[8 * 16];  // [j2][r]: exp(-/+ 2 pi i bin(r) j2 / nx), j2 < RX
  float kscale[8];      // per-t-mode scale: forward applies it at the store, inverse at the load
};

#define PO_TW4_LD 17
#ifdef PO_EMUL
#define PO_SYNCWARP() po_emul::syncwarp()
#else
#define PO_SYNCWARP() __syncwarp()
#endif

// One instruction asks the memory system to bring `bytes` contiguous bytes into L2 (no destination, no
// completion to wait for).  Used to pull the NEXT unit's rows out of DRAM while the current unit is being
// transformed: the register-staged loads then see L2 latency, which the ~80 KB a CTA can keep in flight covers.
PO_D void po_prefetch_l2(const void* g, unsigned bytes) {
#ifndef PO_EMUL
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
#else
  (void)g; (void)bytes;
#endif
}

// L2 residency control.  The spectral stack hands an 84 MB intermediate from kernel to kernel; it fits the 126 MB
// L2 only if the 671 MB streams next to it (the forward kernel's input rows, the inverse kernel's output rows)
// do not evict it.  Streams therefore carry an evict-first cache policy, everything else keeps the default.
// Measured (profiles/raw/gpurun_out_g3b.txt): the inverse kernel alone writes at 6.9 TB/s, but only at 5.3 TB/s when its
// 84 MB of input have to be re-read from DRAM in between the writes.
PO_D unsigned long long po_policy_evict_first() {
  unsigned long long pol = 0;
#ifndef PO_EMUL
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
// explicit global-space 64-bit store (a pointer that went through an opaque asm would otherwise become a generic ST)
template <bool HINT>
PO_D void po_st_global(float2* p, float2 v, unsigned long long pol) {
#ifndef PO_EMUL
  if (HINT) asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
  else asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#else
  (void)pol;
  *p = v;
#endif
}

PO_D float2 po_n2(float2 a) { return make_float2(-a.x, -a.y); }
PO_D float2 po_sub2(float2 a, float2 b) { return __fadd2_rn(a, po_n2(b)); }  // FADD2 a, -b
PO_D float2 po_bc(float w) { return make_float2(w, w); }

// cos / sin of 2 pi m / 64 as compile-time constants (float-rounded exact values)
PO_HD constexpr float po_cos64(int m) {
  constexpr float q[17] = {1.000000000e+00f, 9.951847196e-01f, 9.807852507e-01f, 9.569403529e-01f, 9.238795042e-01f, 8.819212914e-01f,
                           8.314695954e-01f, 7.730104327e-01f, 7.071067691e-01f, 6.343932748e-01f, 5.555702448e-01f, 4.713967443e-01f,
                           3.826834261e-01f, 2.902846634e-01f, 1.950903237e-01f, 9.801714122e-02f, 0.f};
  m &= 63;
  return m <= 16 ? q[m] : m <= 32 ? -q[32 - m] : m <= 48 ? -q[m - 32] : q[64 - m];
}
PO_HD constexpr float po_sin64(int m) { return po_cos64(m - 16); }

// ---- packed 16-point complex FFT on two sequences at once ------------------------------------
// v = (vr, vi): vr[j] = (Re a_j, Re b_j), vi[j] = (Im a_j, Im b_j).  DIF, natural order in,
// bit-reversed out (v[pos] = X[brev4(pos)]), same conventions as po_fft16.
template <int SIGN, int K>
PO_D void po_pk_mul_w16(float2& xr, float2& xi) {
  if constexpr (K == 0) {
    return;
  } else if constexpr (K == 4) {
    const float2 r = xr;
    if constexpr (SIGN < 0) { xr = xi; xi = po_n2(r); } else { xr = po_n2(xi); xi = r; }
  } else {
    constexpr float wr = (K == 1) ? PO_C8 : (K == 2) ? PO_S2 : (K == 3) ? PO_S8 : (K == 5) ? -PO_S8 : (K == 6) ? -PO_S2 : -PO_C8;
    constexpr float ws = (K == 1) ? PO_S8 : (K == 2) ? PO_S2 : (K == 3) ? PO_C8 : (K == 5) ? PO_C8 : (K == 6) ? PO_S2 : PO_S8;
    constexpr float wi = (SIGN < 0) ? -ws : ws;
    const float2 r = __ffma2_rn(xi, po_bc(-wi), __fmul2_rn(xr, po_bc(wr)));
    const float2 i = __ffma2_rn(xi, po_bc(wr), __fmul2_rn(xr, po_bc(wi)));
    xr = r; xi = i;
  }
}

template <int SIGN, int LEN, int BASE, int J>
PO_D void po_pk_bfly(float2 (&vr)[16], float2 (&vi)[16]) {
  constexpr int HALF = LEN / 2, STEP = 16 / LEN;
  const float2 ar = vr[BASE + J], ai = vi[BASE + J], br = vr[BASE + J + HALF], bi = vi[BASE + J + HALF];
  vr[BASE + J] = __fadd2_rn(ar, br);
  vi[BASE + J] = __fadd2_rn(ai, bi);
  float2 dr = po_sub2(ar, br), di = po_sub2(ai, bi);
  po_pk_mul_w16<SIGN, (J * STEP) & 7>(dr, di);
  vr[BASE + J + HALF] = dr;
  vi[BASE + J + HALF] = di;
}

template <int SIGN>
PO_D void po_pk_fft16(float2 (&vr)[16], float2 (&vi)[16]) {
#define PO_PK_B(LEN, BASE, J) po_pk_bfly<SIGN, LEN, BASE, J>(vr, vi);
  PO_PK_B(16, 0, 0) PO_PK_B(16, 0, 1) PO_PK_B(16, 0, 2) PO_PK_B(16, 0, 3) PO_PK_B(16, 0, 4) PO_PK_B(16, 0, 5) PO_PK_B(16, 0, 6) PO_PK_B(16, 0, 7)
  PO_PK_B(8, 0, 0) PO_PK_B(8, 0, 1) PO_PK_B(8, 0, 2) PO_PK_B(8, 0, 3) PO_PK_B(8, 8, 0) PO_PK_B(8, 8, 1) PO_PK_B(8, 8, 2) PO_PK_B(8, 8, 3)
  PO_PK_B(4, 0, 0) PO_PK_B(4, 0, 1) PO_PK_B(4, 4, 0) PO_PK_B(4, 4, 1) PO_PK_B(4, 8, 0) PO_PK_B(4, 8, 1) PO_PK_B(4, 12, 0) PO_PK_B(4, 12, 1)
  PO_PK_B(2, 0, 0) PO_PK_B(2, 2, 0) PO_PK_B(2, 4, 0) PO_PK_B(2, 6, 0) PO_PK_B(2, 8, 0) PO_PK_B(2, 10, 0) PO_PK_B(2, 12, 0) PO_PK_B(2, 14, 0)
#undef PO_PK_B
}

// ---- forward t-phase, compile-time twiddles -----------------------------------------------------
// One step j2 = J2 of the pruned decimation: 8 rows (t = RT*j1 + J2) -> packed real FFT8 -> accumulate
// acc_k += exp(-2 pi i k J2 / nt) * Y_k with the twiddle as FFMA2 immediates.
template <int RT, int J2>
PO_D void po_fwd_tstep(const float2 (&x)[8], float2 (&ar)[8], float2 (&ai)[8]) {
  const float2 e0 = __fadd2_rn(x[0], x[4]), e1 = po_sub2(x[0], x[4]), e2 = __fadd2_rn(x[2], x[6]), e3 = po_sub2(x[2], x[6]);
  const float2 o0 = __fadd2_rn(x[1], x[5]), o1 = po_sub2(x[1], x[5]), o2 = __fadd2_rn(x[3], x[7]), o3 = po_sub2(x[3], x[7]);
  const float2 E0 = __fadd2_rn(e0, e2), E2 = po_sub2(e0, e2), O0 = __fadd2_rn(o0, o2), O2 = po_sub2(o0, o2);
  const float2 pp = __fmul2_rn(po_sub2(o1, o3), po_bc(PO_S2)), qq = __fmul2_rn(__fadd2_rn(o1, o3), po_bc(PO_S2));
  // Y_k = (P_k, sg_k * Q_k):  0 (E0+O0, 0)  4 (E0-O0, 0)  2 (E2,-O2)  6 (E2,+O2)
  //                           1 (e1+p, -(e3+q))  7 (e1+p, +(e3+q))  3 (e1-p, +(e3-q))  5 (e1-p, -(e3-q))
  const float2 P0 = __fadd2_rn(E0, O0), P4 = po_sub2(E0, O0);
  const float2 A1 = __fadd2_rn(e1, pp), B1 = __fadd2_rn(e3, qq), A3 = po_sub2(e1, pp), B3 = po_sub2(e3, qq);
  if constexpr (J2 == 0) {
    ar[0] = P0; ai[0] = make_float2(0.f, 0.f); ar[4] = P4; ai[4] = make_float2(0.f, 0.f);
    ar[2] = E2; ai[2] = po_n2(O2); ar[6] = E2; ai[6] = O2;
    ar[1] = A1; ai[1] = po_n2(B1); ar[7] = A1; ai[7] = B1;
    ar[3] = A3; ai[3] = B3; ar[5] = A3; ai[5] = po_n2(B3);
  } else {
    constexpr int S = 8 / RT;  // index step into the 64-point circle: angle index = k * J2 * S
    // (P + i sg Q)(wx + i wy): re += wx P - sg wy Q ; im += sg wx Q + wy P,   w = exp(-2 pi i idx / 64)
#define PO_ACC(k, P, Q, SG)                                                       \
  {                                                                               \
    constexpr int idx = ((k) * J2 * S) & 63;                                      \
    constexpr float wx = po_cos64(idx), wy = -po_sin64(idx);                      \
    if constexpr ((idx & 15) == 0 && (idx & 31) == 0) { /* w = +-1 */             \
      ar[k] = __ffma2_rn(P, po_bc(wx), ar[k]);                                    \
      ai[k] = __ffma2_rn(Q, po_bc((SG) * wx), ai[k]);                             \
    } else if constexpr ((idx & 15) == 0) { /* w = +-i */                         \
      ar[k] = __ffma2_rn(Q, po_bc(-(SG) * wy), ar[k]);                            \
      ai[k] = __ffma2_rn(P, po_bc(wy), ai[k]);                                    \
    } else {                                                                      \
      ar[k] = __ffma2_rn(P, po_bc(wx), ar[k]);                                    \
      ar[k] = __ffma2_rn(Q, po_bc(-(SG) * wy), ar[k]);                            \
      ai[k] = __ffma2_rn(Q, po_bc((SG) * wx), ai[k]);                             \
      ai[k] = __ffma2_rn(P, po_bc(wy), ai[k]);                                    \
    }                                                                             \
  }
    ar[0] = __fadd2_rn(ar[0], P0);  // k = 0: w = 1, Y real
    {
      constexpr int idx = (4 * J2 * S) & 63;
      constexpr float wx = po_cos64(idx), wy = -po_sin64(idx);
      if constexpr ((idx & 15) == 0 && (idx & 31) == 0) ar[4] = __ffma2_rn(P4, po_bc(wx), ar[4]);
      else if constexpr ((idx & 15) == 0) ai[4] = __ffma2_rn(P4, po_bc(wy), ai[4]);
      else { ar[4] = __ffma2_rn(P4, po_bc(wx), ar[4]); ai[4] = __ffma2_rn(P4, po_bc(wy), ai[4]); }
    }
    PO_ACC(2, E2, O2, -1.f) PO_ACC(6, E2, O2, 1.f)
    PO_ACC(1, A1, B1, -1.f) PO_ACC(7, A1, B1, 1.f)
    PO_ACC(3, A3, B3, 1.f) PO_ACC(5, A3, B3, -1.f)
#undef PO_ACC
  }
}

// NJ consecutive steps j2 = J0 .. J0+NJ-1: all 8*NJ row loads are issued first, then the steps are evaluated.
// Rows are addressed with 32-bit float2 offsets from the tensor base (host checks the range): one IADD +
// one IMAD.WIDE per load.  `rs` is made opaque per call so the compiler does not hoist 32 row offsets out
// of the caller's loop (they would cost 64 registers).
#ifdef PO_EMUL
#define PO_OPAQUE_U32(x)
#define PO_OPAQUE_PTR(x)
#else
#define PO_OPAQUE_U32(x) asm volatile("" : "+r"(x))
#define PO_OPAQUE_PTR(x) asm volatile("" : "+l"(x))
#endif
template <int RT, int J0, int NJ>
PO_D void po_fwd_tgroup(const float2* __restrict__ xb, unsigned off0, unsigned rs, float2 (&ar)[8], float2 (&ai)[8]) {
  float2 x[NJ][8];
  PO_OPAQUE_U32(rs);
  unsigned off = off0 + rs * J0;
#pragma unroll
  for (int j1 = 0; j1 < 8; ++j1) {
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { x[jj][j1] = __ldg(xb + off); off += rs; }
    if constexpr (RT > NJ) off += rs * (RT - NJ);
  }
  po_fwd_tstep<RT, J0>(x[0], ar, ai);
  po_fwd_tstep<RT, J0 + 1>(x[1], ar, ai);
  if constexpr (NJ > 2) { po_fwd_tstep<RT, J0 + 2>(x[2], ar, ai); po_fwd_tstep<RT, J0 + 3>(x[3], ar, ai); }
}

// xb: tensor base (float2); ubase: float2 offset of the unit's first row; rs: row stride in float2
template <int RT, bool WHOLE>
PO_D void po_fwd_tphase2(const float2* __restrict__ xb, unsigned ubase, unsigned rs, float* __restrict__ T, int L, int t, int nthreads, int IG,
                         int i0, int I0) {
  for (int v = t; v < (L >> 1); v += nthreads) {
    const int gp = WHOLE ? v : po_gpair(v, IG >> 1, i0 >> 1, I0 >> 1);
    float2 ar[8], ai[8];
    po_fwd_tgroup<RT, 0, (RT < 4 ? RT : 4)>(xb, ubase + gp, rs, ar, ai);
    if constexpr (RT > 4) po_fwd_tgroup<RT, 4, 4>(xb, ubase + gp, rs, ar, ai);
#pragma unroll
    for (int k = 0; k < 8; ++k) *(float4*)(T + k * 2 * L + 4 * v) = make_float4(ar[k].x, ar[k].y, ai[k].x, ai[k].y);
  }
}

// ---- forward x-phase, packed ----------------------------------------------------------------------
// work item = (channel pair p < IG/2, t-mode k < 8, residue j2 < RX), j2 fastest: the RX lanes of one (p, k)
// are adjacent.  Each lane: 16 x LDS.128 -> packed FFT16 -> twiddle by exp(-2 pi i bin j2 / nx) -> write the
// 16 partial bins back over its own (consumed) slots -> warp sync -> sum its 16/RX bins over the RX lanes
// (rotated read order: conflict-free) -> scale -> 128-bit stores of (channel a, channel b) complex pairs.
// tw4: shared [RX][16] float4 {wr, wr, wi, wi}; ksc: shared [8].
// SHFL: the RX partial spectra are combined by a reduce-scatter over the RX adjacent lanes (recursive halving with
// shuffles: 3 x 16 SHFL for RX = 4) instead of a round trip through the tile.  Per warp and item that trades 128
// shared-memory wavefronts (16 STS.128 + 16 LDS.128) for 48 shuffles and ~100 selects; the ring kernel, whose
// bound is the shared-memory pipe (ncu r1final: 60 % LSU wavefronts + the bulk-copy fill), uses it.  Needs
// lane % RX == item % RX (t congruent to the lane index mod 32, nthreads and item_begin multiples of 32).
template <int RX>
PO_D void po_lane_reduce_scatter(float2 (&gr)[16], float2 (&gi)[16], int j2) {
#pragma unroll
  for (int bit = RX >> 1; bit >= 1; bit >>= 1) {
    const bool up = (j2 & bit) != 0;  // this lane keeps the bins whose bit `bit` is set
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if ((r & bit) == 0 && (r & (RX - 1) & ~(2 * bit - 1)) == 0) {  // live slots: the bits already exchanged are folded into slot bits = 0
        const int r1 = r | bit;
        const float2 sr = up ? gr[r] : gr[r1], si = up ? gi[r] : gi[r1];
        const float2 kr = up ? gr[r1] : gr[r], ki = up ? gi[r1] : gi[r];
        float2 tr, ti;
        tr.x = PO_SHFL_XOR(sr.x, bit); tr.y = PO_SHFL_XOR(sr.y, bit);
        ti.x = PO_SHFL_XOR(si.x, bit); ti.y = PO_SHFL_XOR(si.y, bit);
        gr[r] = __fadd2_rn(kr, tr);
        gi[r] = __fadd2_rn(ki, ti);
      }
    }
  }
}

template <int RX, bool SHFL = false>
PO_D void po_fwd_xphase2(float* __restrict__ T, const float4* __restrict__ tw4, const float* __restrict__ ksc, float2* __restrict__ zu,
                         i64 kstride, int I0, int L, int t, int nthreads, int IG, int i0, unsigned long long zpol = 0ull, int item_begin = 0,
                         int item_end = -1) {
  constexpr int NB = 16 / RX;  // bins owned by a lane
  const int hp = IG >> 1;
  const int nitems = (item_end >= 0) ? item_end : hp * 8 * RX;  // [item_begin, nitems): multiples of 32
  for (int base = item_begin; base < nitems; base += nthreads) {
    if (base + (t & ~31) >= nitems) break;  // warp-uniform (nitems % 32 == 0)
    const int item = base + t;
    const int j2 = item % RX, pk = item / RX;
    const int p = pk % hp, k = pk / hp;
    float4* tk = (float4*)(T + (size_t)k * 2 * L) + p;  // float4 slot of (pair p, x): tk[hp * x]
    float2 vr[16], vi[16];
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) {
      const float4 u = tk[hp * (RX * j1 + j2)];
      vr[j1] = make_float2(u.x, u.y);
      vi[j1] = make_float2(u.z, u.w);
    }
    po_pk_fft16<-1>(vr, vi);
    const float4* twj = tw4 + j2 * PO_TW4_LD;  // rows padded to 17 float4: the RX lanes hit distinct banks
    if constexpr (SHFL) {
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
      float2* zb = zu + (i0 + 2 * p) + kstride * k;
#pragma unroll
      for (int qb = 0; qb < NB; ++qb) {
        const int r = qb * RX + j2;  // this lane's bin: its sum sits in slot qb * RX
        const float2 sr = gr[qb * RX], si = gi[qb * RX];
        const float4 zv = make_float4(sr.x * ks, si.x * ks, sr.y * ks, si.y * ks);
        if (zpol) po_st_hint((float4*)(zb + (i64)I0 * r), zv, zpol); else *(float4*)(zb + (i64)I0 * r) = zv;
      }
    } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 w = twj[r];
      const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
      const float2 fr = vr[po_brev4(r)], fi = vi[po_brev4(r)];
      const float2 gr = __ffma2_rn(fi, po_n2(wi), __fmul2_rn(fr, wr));
      const float2 gi = __ffma2_rn(fi, wr, __fmul2_rn(fr, wi));
      tk[hp * (RX * r + j2)] = make_float4(gr.x, gr.y, gi.x, gi.y);
    }
    PO_SYNCWARP();
    const float ks = ksc[k];
    float2* zb = zu + (i0 + 2 * p) + kstride * k;
#pragma unroll
    for (int qb = 0; qb < NB; ++qb) {
      const int r = qb * RX + j2;  // this lane's bin
      float2 sr = make_float2(0.f, 0.f), si = make_float2(0.f, 0.f);
#pragma unroll
      for (int s = 0; s < RX; ++s) {
        const float4 u = tk[hp * (RX * r + ((j2 + s) & (RX - 1)))];
        sr = __fadd2_rn(sr, make_float2(u.x, u.y));
        si = __fadd2_rn(si, make_float2(u.z, u.w));
      }
      const float4 zv = make_float4(sr.x * ks, si.x * ks, sr.y * ks, si.y * ks);
      if (zpol) po_st_hint((float4*)(zb + (i64)I0 * r), zv, zpol); else *(float4*)(zb + (i64)I0 * r) = zv;
    }
    }  // !SHFL
  }
}

// ---- inverse x-phase, packed ----------------------------------------------------------------------
// tw4k: shared [8][RX][16] float4 {wr, wr, wi, wi} = kscale[k] * exp(+2 pi i bin(r) j2 / nx)
// SMEM_SRC: zu points at rows staged in shared memory (po_inv_xt_stage_kernel) instead of at the global tensor
template <int RX, bool SMEM_SRC = false>
PO_D void po_inv_xphase2(const float2* __restrict__ zu, i64 kstride, float* __restrict__ T, const float4* __restrict__ tw4k, int I0, int L,
                         int t, int nthreads, int IG, int i0) {
  const int hp = IG >> 1;
  const int nitems = hp * 8 * RX;
  for (int item = t; item < nitems; item += nthreads) {
    const int j2 = item % RX, pk = item / RX;
    const int p = pk % hp, k = pk / hp;
    const float2* zb = zu + (i0 + 2 * p) + kstride * k;
    const float4* tw = tw4k + (k * RX + j2) * PO_TW4_LD;
    float2 vr[16], vi[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 z = SMEM_SRC ? *(const float4*)(zb + I0 * r) : __ldg((const float4*)(zb + (i64)I0 * r));  // (re_a, im_a, re_b, im_b)
      const float4 w = tw[r];
      const float2 zr = make_float2(z.x, z.z), zi = make_float2(z.y, z.w);
      const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
      vr[r] = __ffma2_rn(zi, po_n2(wi), __fmul2_rn(zr, wr));
      vi[r] = __ffma2_rn(zi, wr, __fmul2_rn(zr, wi));
    }
    po_pk_fft16<1>(vr, vi);
    float4* tk = (float4*)(T + (size_t)k * 2 * L) + p;
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) {
      const float2 a = vr[po_brev4(j1)], b = vi[po_brev4(j1)];
      tk[hp * (RX * j1 + j2)] = make_float4(a.x, a.y, b.x, b.y);
    }
  }
}

// ---- inverse t-phase, compile-time twiddles ----------------------------------------------------------
template <int RT, int J2>
PO_D void po_inv_tstep(const float2 (&ur)[8], const float2 (&ui)[8], float2* __restrict__ yb, unsigned off0, unsigned rs, unsigned long long pol) {
  constexpr int S = 8 / RT;
  // The 8 store addresses are formed FIRST and pinned in 8 distinct register pairs: ncu showed the stores of
  // generation 2 serialised on ONE re-used address pair (the next IMAD.WIDE waits on the scoreboard until the
  // LSU has consumed the previous STG's address), capping the kernel at 4.6 TB/s.
  float2* p[8];
  {
    PO_OPAQUE_U32(rs);
    unsigned off = off0 + rs * J2;
#pragma unroll
    for (int j1 = 0; j1 < 8; ++j1) { p[j1] = yb + off; PO_OPAQUE_PTR(p[j1]); off += rs * RT; }
  }
  // V_k = exp(+2 pi i k J2 / nt) U_k; only the parts po_irfft8_re needs
  float2 vr[8], vi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { vr[k] = ur[k]; vi[k] = ui[k]; }
#define PO_TWK(k)                                                                              \
  {                                                                                            \
    constexpr int idx = ((k) * J2 * S) & 63;                                                   \
    constexpr float wx = po_cos64(idx), wy = po_sin64(idx);                                    \
    if constexpr (idx == 0) {                                                                  \
    } else if constexpr (idx == 32) { vr[k] = po_n2(ur[k]); vi[k] = po_n2(ui[k]);              \
    } else if constexpr (idx == 16) { vr[k] = po_n2(ui[k]); vi[k] = ur[k];                     \
    } else if constexpr (idx == 48) { vr[k] = ui[k]; vi[k] = po_n2(ur[k]);                     \
    } else {                                                                                   \
      vr[k] = __ffma2_rn(ui[k], po_bc(-wy), __fmul2_rn(ur[k], po_bc(wx)));                     \
      vi[k] = __ffma2_rn(ui[k], po_bc(wx), __fmul2_rn(ur[k], po_bc(wy)));                      \
    }                                                                                          \
  }
  if constexpr (J2 != 0) { PO_TWK(1) PO_TWK(2) PO_TWK(3) PO_TWK(4) PO_TWK(5) PO_TWK(6) PO_TWK(7) }
#undef PO_TWK
  const float2 H0 = vr[0], H4 = vr[4];
  const float2 r1 = __fadd2_rn(vr[1], vr[7]), i1 = po_sub2(vi[1], vi[7]);
  const float2 r2 = __fadd2_rn(vr[2], vr[6]), i2 = po_sub2(vi[2], vi[6]);
  const float2 r3 = __fadd2_rn(vr[3], vr[5]), i3 = po_sub2(vi[3], vi[5]);
  const float2 a = __fadd2_rn(H0, H4), b = po_sub2(H0, H4);
  const float2 p1 = __fmul2_rn(po_sub2(r1, i1), po_bc(PO_S2)), q1 = __fmul2_rn(__fadd2_rn(r1, i1), po_bc(PO_S2));
  const float2 p3 = __fmul2_rn(po_sub2(r3, i3), po_bc(PO_S2)), q3 = __fmul2_rn(__fadd2_rn(r3, i3), po_bc(PO_S2));
  const float2 ap = __fadd2_rn(a, r2), am = po_sub2(a, r2), bp = __fadd2_rn(b, i2), bm = po_sub2(b, i2);
  const float2 r13 = __fadd2_rn(r1, r3), i13 = po_sub2(i1, i3), pq = po_sub2(p1, q3), qp = po_sub2(q1, p3);
  float2 out[8];
  out[0] = __fadd2_rn(ap, r13);
  out[4] = po_sub2(ap, r13);
  out[2] = po_sub2(am, i13);
  out[6] = __fadd2_rn(am, i13);
  out[1] = __fadd2_rn(bm, pq);
  out[5] = po_sub2(bm, pq);
  out[3] = po_sub2(bp, qp);
  out[7] = __fadd2_rn(bp, qp);
#pragma unroll
  for (int j1 = 0; j1 < 8; ++j1) { if (pol) po_st_global<true>(p[j1], out[j1], pol); else po_st_global<false>(p[j1], out[j1], pol); }
}

template <int RT, bool WHOLE>
PO_D void po_inv_tphase2(const float* __restrict__ T, float2* __restrict__ yb, unsigned ubase, unsigned rs, int L, int t, int nthreads, int IG,
                         int i0, int I0, unsigned long long pol) {
  for (int v = t; v < (L >> 1); v += nthreads) {
    const int gp = WHOLE ? v : po_gpair(v, IG >> 1, i0 >> 1, I0 >> 1);
    float2 ur[8], ui[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 u = *(const float4*)(T + k * 2 * L + 4 * v);
      ur[k] = make_float2(u.x, u.y);
      ui[k] = make_float2(u.z, u.w);
    }
    const unsigned o = ubase + gp;
    po_inv_tstep<RT, 0>(ur, ui, yb, o, rs, pol);
    po_inv_tstep<RT, 1>(ur, ui, yb, o, rs, pol);
    if constexpr (RT > 2) { po_inv_tstep<RT, 2>(ur, ui, yb, o, rs, pol); po_inv_tstep<RT, 3>(ur, ui, yb, o, rs, pol); }
    if constexpr (RT > 4) {
      po_inv_tstep<RT, 4>(ur, ui, yb, o, rs, pol); po_inv_tstep<RT, 5>(ur, ui, yb, o, rs, pol);
      po_inv_tstep<RT, 6>(ur, ui, yb, o, rs, pol); po_inv_tstep<RT, 7>(ur, ui, yb, o, rs, pol);
    }
  }
}

// ---- pipelined persistent kernels (one CTA / SM, warp-specialised, double-buffered tile) ----------------
// HP > 0: the channel-group size is the compile-time constant IG = 2*HP (shared-memory addresses become
// immediates); HP == 0: any even IG at run time.
template <int RT, int RX, bool WHOLE, int HP>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_fwd_tx_pipe2_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused2_tw tw, int I0, int IG_rt, i64 Mid,
                       i64 units, int nstream, int prefetch) {
  PO_DYN_SMEM(smem_raw);
  const int IG = HP ? 2 * HP : IG_rt;
  const int L = IG * 16 * RX, LR = I0 * 16 * RX, G = I0 / IG;
  float* T0 = (float*)smem_raw;                      // two [8][2L] tiles
  float4* tw4 = (float4*)(T0 + 32 * (size_t)L);      // [RX][17]
  float* ksc = (float*)(tw4 + RX * PO_TW4_LD);       // [8]
  const int tid = threadIdx.x;
  po_pdl_trigger();
  const i64 rs2 = ((i64)LR * Mid) >> 1;
  for (int e = tid; e < RX * 16; e += PO_PIPE_THREADS) {
    const float2 w = tw.tw_x[e];
    tw4[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x, w.x, w.y, w.y);
  }
  if (tid < 8) ksc[tid] = tw.kscale[tid];
  __syncthreads();
  po_pdl_wait();
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (i64 it = 0; it <= nit; ++it) {
    if (tid < nstream) {
      if (prefetch && tid < 8 * RT && it + prefetch < nit) {  // one row of unit it+prefetch per thread
        const po_unit un = po_decode_unit(blockIdx.x + (it + prefetch) * gridDim.x, G, IG, Mid);
        po_prefetch_l2(X + (i64)LR * un.m + (i64)LR * Mid * (i64)(8 * RT) * un.bb + (i64)LR * Mid * tid, (unsigned)LR * 4u);
      }
      if (it < nit) {
        const po_unit u = po_decode_unit(blockIdx.x + it * gridDim.x, G, IG, Mid);
        const unsigned ubase = (unsigned)(((i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb) >> 1);
        po_fwd_tphase2<RT, WHOLE>((const float2*)X, ubase, (unsigned)rs2, T0 + (it & 1) * 16 * (size_t)L, L, tid, nstream, IG, u.i0, I0);
      }
    } else if (it > 0) {
      const po_unit u = po_decode_unit(blockIdx.x + (it - 1) * gridDim.x, G, IG, Mid);
      po_fwd_xphase2<RX>(T0 + ((it - 1) & 1) * 16 * (size_t)L, tw4, ksc, Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, I0, L,
                         tid - nstream, PO_PIPE_THREADS - nstream, IG, u.i0);
    }
    __syncthreads();
  }
}

template <int RT, int RX, bool WHOLE, int HP>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_inv_xt_pipe2_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused2_tw tw, int I0, int IG_rt, i64 Mid,
                       i64 units, int nstream, int prefetch, i64 m_lo, i64 m_cnt) {
  PO_DYN_SMEM(smem_raw);
  const int IG = HP ? 2 * HP : IG_rt;
  const int L = IG * 16 * RX, LR = I0 * 16 * RX, G = I0 / IG;
  float* T0 = (float*)smem_raw;
  float4* tw4k = (float4*)(T0 + 32 * (size_t)L);  // [8][RX][16]
  const int tid = threadIdx.x;
  po_pdl_trigger();
  const i64 rs2 = ((i64)LR * Mid) >> 1;
  for (int e = tid; e < 8 * RX * 16; e += PO_PIPE_THREADS) {
    const int k = e / (RX * 16);
    const float2 w = tw.tw_x[e % (RX * 16)];
    const float s = tw.kscale[k];
    tw4k[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x * s, w.x * s, w.y * s, w.y * s);
  }
  __syncthreads();
  po_pdl_wait();
  const unsigned long long pol = (prefetch & 0x400) ? po_policy_evict_first() : 0ull;  // output rows: stream through L2
  const i64 nit = (units - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (i64 it = 0; it <= nit; ++it) {
    if (tid >= nstream) {
      if ((prefetch & 0xff) && tid - nstream < 8 && it + (prefetch & 0xff) < nit) {  // one t-mode row of unit it+prefetch per thread
        const po_unit un = po_decode_unit(blockIdx.x + (it + (prefetch & 0xff)) * gridDim.x, G, IG, m_lo, m_cnt);
        po_prefetch_l2(Z + (i64)I0 * 16 * (un.m + Mid * 8 * un.bb) + (i64)I0 * 16 * Mid * (tid - nstream), (unsigned)I0 * 16u * 8u);
      }
      if (it < nit && !(prefetch & 0x100)) {
        const po_unit u = po_decode_unit(blockIdx.x + it * gridDim.x, G, IG, m_lo, m_cnt);
        po_inv_xphase2<RX>(Z + (i64)I0 * 16 * (u.m + Mid * 8 * u.bb), (i64)I0 * 16 * Mid, T0 + (it & 1) * 16 * (size_t)L, tw4k, I0, L,
                           tid - nstream, PO_PIPE_THREADS - nstream, IG, u.i0);
      }
    } else if (it > 0 && !(prefetch & 0x200)) {
      const po_unit u = po_decode_unit(blockIdx.x + (it - 1) * gridDim.x, G, IG, m_lo, m_cnt);
      const unsigned ubase = (unsigned)(((i64)LR * u.m + (i64)LR * Mid * (i64)(8 * RT) * u.bb) >> 1);
      po_inv_tphase2<RT, WHOLE>(T0 + ((it - 1) & 1) * 16 * (size_t)L, (float2*)Y, ubase, (unsigned)rs2, L, tid, nstream, IG, u.i0, I0, pol);
    }
    __syncthreads();
  }
}

// ---- single-tile persistent kernels for rows that do not fit a double-buffered tile (c * nx = 20 * 128: 160 KB) -------
// Same packed phases, but one tile: all 512 threads stream (t-phase), then all run the x-phase.  Replaces the
// channel-group variant of the pipelined forward (whose sibling groups fetch every sector twice from L2) and the
// first-generation whole-row inverse for these shapes (the local shapes of the C4 ladder from 2 GPUs on).
template <int RT, int RX, int HP>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_fwd_tx_tile2_kernel(const float* __restrict__ X, float2* __restrict__ Z, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units,
                       int l2hint) {
  PO_DYN_SMEM(smem_raw);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX;
  float* T = (float*)smem_raw;                     // [8][2L]
  float4* tw4 = (float4*)(T + 16 * (size_t)L);     // [RX][17]
  float* ksc = (float*)(tw4 + RX * PO_TW4_LD);     // [8]
  const int tid = threadIdx.x;
  po_pdl_trigger();
  const i64 rs2 = ((i64)L * Mid) >> 1;
  for (int e = tid; e < RX * 16; e += PO_PIPE_THREADS) {
    const float2 w = tw.tw_x[e];
    tw4[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x, w.x, w.y, w.y);
  }
  if (tid < 8) ksc[tid] = tw.kscale[tid];
  const unsigned long long zpol = l2hint ? po_policy_evict_last() : 0ull;
  __syncthreads();
  po_pdl_wait();
  for (i64 unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const i64 m = unit % Mid, bb = unit / Mid;
    const unsigned ubase = (unsigned)(((i64)L * m + (i64)L * Mid * (i64)(8 * RT) * bb) >> 1);
    po_fwd_tphase2<RT, true>((const float2*)X, ubase, (unsigned)rs2, T, L, tid, PO_PIPE_THREADS, I0, 0, I0);
    __syncthreads();
    po_fwd_xphase2<RX>(T, tw4, ksc, Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, I0, L, tid, PO_PIPE_THREADS, I0, 0, zpol);
    __syncthreads();
  }
}

template <int RT, int RX, int HP>
__global__ void __launch_bounds__(PO_PIPE_THREADS, 1)
po_inv_xt_tile2_kernel(const float2* __restrict__ Z, float* __restrict__ Y, const PO_GRID_CONSTANT po_fused2_tw tw, int I0_rt, i64 Mid, i64 units,
                       int l2hint, int prefetch) {
  PO_DYN_SMEM(smem_raw);
  const int I0 = HP ? 2 * HP : I0_rt;
  const int L = I0 * 16 * RX;
  float* T = (float*)smem_raw;
  float4* tw4k = (float4*)(T + 16 * (size_t)L);    // [8][RX][17]
  const int tid = threadIdx.x;
  po_pdl_trigger();
  const i64 rs2 = ((i64)L * Mid) >> 1;
  for (int e = tid; e < 8 * RX * 16; e += PO_PIPE_THREADS) {
    const int k = e / (RX * 16);
    const float2 w = tw.tw_x[e % (RX * 16)];
    const float s = tw.kscale[k];
    tw4k[(e >> 4) * PO_TW4_LD + (e & 15)] = make_float4(w.x * s, w.x * s, w.y * s, w.y * s);
  }
  const unsigned long long pol = l2hint ? po_policy_evict_first() : 0ull;
  __syncthreads();
  po_pdl_wait();
  for (i64 unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const i64 m = unit % Mid, bb = unit / Mid;
    if (prefetch && tid < 8 && unit + gridDim.x < units) {  // next unit's 8 t-mode rows -> L2
      const i64 un = unit + gridDim.x, mn = un % Mid, bn = un / Mid;
      po_prefetch_l2(Z + (i64)I0 * 16 * (mn + Mid * 8 * bn) + (i64)I0 * 16 * Mid * tid, (unsigned)I0 * 16u * 8u);
    }
    po_inv_xphase2<RX>(Z + (i64)I0 * 16 * (m + Mid * 8 * bb), (i64)I0 * 16 * Mid, T, tw4k, I0, L, tid, PO_PIPE_THREADS, I0, 0);
    __syncthreads();
    const unsigned ubase = (unsigned)(((i64)L * m + (i64)L * Mid * (i64)(8 * RT) * bb) >> 1);
    po_inv_tphase2<RT, true>(T, (float2*)Y, ubase, (unsigned)rs2, L, tid, PO_PIPE_THREADS, I0, 0, I0, pol);
    __syncthreads();
  }
}

static inline size_t po_tile2_smem(bool inverse, i64 L, int RX) {
  return (size_t)16 * L * sizeof(float) + (size_t)(inverse ? 8 : 1) * RX * PO_TW4_LD * sizeof(float4) + 64;
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_FUSED2)
template <int RT, int RX>
static cudaError_t po_launch_tile2_t(bool inverse, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                                     cudaStream_t st) {
  const i64 L = (i64)I0 * 16 * RX;
  if ((I0 & 1) || num_sms <= 0) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16)) return cudaErrorNotSupported;
  const size_t smem = po_tile2_smem(inverse, L, RX);
  if (smem > 220 * 1024) return cudaErrorNotSupported;
  const i64 units = Mid * B;
  if (units < 4 * (i64)num_sms) return cudaErrorNotSupported;
  if ((i64)I0 * 16 * RX * Mid * (8 * RT) * B / 2 >= (1LL << 32)) return cudaErrorNotSupported;
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  static const int pf_inv = getenv("PO_PIPE2_PREFETCH_INV") ? atoi(getenv("PO_PIPE2_PREFETCH_INV")) : 1;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-2 single-tile %s RT=%d RX=%d I0=%d units=%lld smem=%zu\n", inverse ? "inv" : "fwd", RT, RX, I0, (long long)units, smem);
#ifndef PO_EMUL
  cudaError_t e;
#define PO_T2_ATTR(K) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
#else
#define PO_T2_ATTR(K)
#endif
  if (!inverse) {
    if (I0 == 20) { PO_T2_ATTR((po_fwd_tx_tile2_kernel<RT, RX, 10>)) PO_LAUNCH_PDL((po_fwd_tx_tile2_kernel<RT, RX, 10>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, l2hint); }
    else { PO_T2_ATTR((po_fwd_tx_tile2_kernel<RT, RX, 0>)) PO_LAUNCH_PDL((po_fwd_tx_tile2_kernel<RT, RX, 0>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float*)src, (float2*)dst, tw, I0, Mid, units, l2hint); }
  } else {
    if (I0 == 20) { PO_T2_ATTR((po_inv_xt_tile2_kernel<RT, RX, 10>)) PO_LAUNCH_PDL((po_inv_xt_tile2_kernel<RT, RX, 10>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, l2hint, pf_inv); }
    else { PO_T2_ATTR((po_inv_xt_tile2_kernel<RT, RX, 0>)) PO_LAUNCH_PDL((po_inv_xt_tile2_kernel<RT, RX, 0>), dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (const float2*)src, (float*)dst, tw, I0, Mid, units, l2hint, pf_inv); }
  }
#undef PO_T2_ATTR
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_tile2(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B,
                                   int num_sms, cudaStream_t st) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_TILE2_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_tile2_t<rt, rx>(inverse, src, dst, tw, I0, Mid, B, num_sms, st);
  PO_TILE2_CASE(2, 4) PO_TILE2_CASE(4, 4) PO_TILE2_CASE(8, 4)
  PO_TILE2_CASE(2, 8) PO_TILE2_CASE(4, 8) PO_TILE2_CASE(8, 8)
#undef PO_TILE2_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_tile2(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B,
                            int num_sms, cudaStream_t st);
#endif

static inline size_t po_fused2_smem(bool inverse, i64 L, int RX) {
  return (size_t)32 * L * sizeof(float) + (size_t)(inverse ? 8 : 1) * RX * PO_TW4_LD * sizeof(float4) + 64;
}
// largest even channel group whose double-buffered tile + tables fit
static inline int po_fused2_group(bool inverse, i64 I0, i64 nx, size_t budget) {
  for (i64 ig = I0; ig >= 2; --ig) {
    if (I0 % ig || (ig & 1)) continue;
    if (po_fused2_smem(inverse, ig * nx, (int)(nx / 16)) <= budget) return (int)ig;
  }
  return 0;
}

// returns cudaErrorNotSupported when the shape is not handled here (caller falls back to generation 1)
// shape test of the INVERSE launch below (m_cnt = mid range of one launch); no side effects
static bool po_fused2_inv_ok(int nt, int nx, int I0, i64 Mid, i64 m_cnt, i64 B, int num_sms, const void* src, const void* dst) {
  if (!(nt == 16 || nt == 32 || nt == 64) || !(nx == 64 || nx == 128)) return false;
  if (num_sms <= 0 || po_fused2_group(true, I0, nx, 216 * 1024) != I0) return false;
  if (((size_t)src % 16) || ((size_t)dst % 16)) return false;
  if (m_cnt <= 0 || m_cnt > Mid || m_cnt * B < 4 * (i64)num_sms) return false;
  return (i64)I0 * nx * Mid * nt * B / 2 < (1LL << 32);
}

#if !defined(PO_MULTI_TU) || defined(PO_TU_FUSED2)
// m_cnt >= 0 (inverse only): process the mid range [m_lo, m_lo + m_cnt) of every batch entry instead of all of Mid
template <int RT, int RX>
static cudaError_t po_launch_fused2_t(bool inverse, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B, int num_sms,
                                      cudaStream_t st, i64 m_lo, i64 m_cnt) {
  const i64 nx = 16 * RX;
  int IG = po_fused2_group(inverse, I0, nx, 216 * 1024);
  if (IG <= 0 || num_sms <= 0) return cudaErrorNotSupported;
  if (((size_t)src % 16) || ((size_t)dst % 16)) return cudaErrorNotSupported;  // 128-bit accesses of the spectral tensor
  // channel groups make the inverse's streaming stores partial-sector (see po_launch_fused_tx_t): whole rows only
  if (inverse && IG != I0) return cudaErrorNotSupported;
  if (m_cnt >= 0 && (!inverse || m_lo < 0 || m_lo + m_cnt > Mid)) return cudaErrorInvalidValue;
  if (m_cnt < 0) { m_lo = 0; m_cnt = Mid; }
  const i64 units = m_cnt * B * (I0 / IG);
  if (units < 4 * (i64)num_sms) return cudaErrorNotSupported;
  if ((i64)I0 * nx * Mid * (8 * RT) * B / 2 >= (1LL << 32)) return cudaErrorNotSupported;  // 32-bit float2 row offsets
  const size_t smem = po_fused2_smem(inverse, (i64)IG * nx, RX);
  static const int ns_fwd_env = getenv("PO_PIPE2_STREAM_FWD") ? atoi(getenv("PO_PIPE2_STREAM_FWD")) : 0;
  static const int ns_inv_env = getenv("PO_PIPE2_STREAM_INV") ? atoi(getenv("PO_PIPE2_STREAM_INV")) : 0;
  const int ns = inverse ? (ns_inv_env ? ns_inv_env : PO_PIPE_STREAM) : (ns_fwd_env ? ns_fwd_env : PO_PIPE_STREAM);
  const bool whole = IG == I0;
  // L2 prefetch distance in units (0: off).  Measured on B200 (profiles/raw/gpurun_out_g2b.txt): the inverse gains 14 % at C3
  // (its x-phase loads stop paying DRAM latency); the forward LOSES 6-40 % -- its 8*RT rows per unit sit at a
  // large power-of-two stride and evict each other from L2 before they are used -- so it stays off there.
  static const int pf_fwd = getenv("PO_PIPE2_PREFETCH_FWD") ? atoi(getenv("PO_PIPE2_PREFETCH_FWD")) : 0;
  static const int pf_inv = getenv("PO_PIPE2_PREFETCH_INV") ? atoi(getenv("PO_PIPE2_PREFETCH_INV")) : 1;
  static const int dbg_skip = getenv("PO_DEBUG_INV_SKIP") ? atoi(getenv("PO_DEBUG_INV_SKIP")) : 0;  // experiments: 1 no x-phase, 2 no t-phase
  static const int l2hint = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  const int pf = inverse ? (pf_inv | (dbg_skip << 8) | (l2hint ? 0x400 : 0)) : pf_fwd;
  static const bool trace = getenv("PO_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[parops] fused gen-2 %s RT=%d RX=%d I0=%d IG=%d units=%lld nstream=%d smem=%zu\n", inverse ? "inv" : "fwd", RT, RX, I0, IG, (long long)units, ns, smem);
#ifndef PO_EMUL
  cudaError_t e;
#define PO_F2_GO(K, SRC_T, DST_T, ...)                                                                         \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
  if (e != cudaSuccess) return e;                                                                              \
  PO_LAUNCH_PDL(K, dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (SRC_T)src, (DST_T)dst, tw, I0, IG, Mid, units, ns, pf, ##__VA_ARGS__);
#else
#define PO_F2_GO(K, SRC_T, DST_T, ...) \
  PO_LAUNCH_PDL(K, dim3((unsigned)num_sms), dim3(PO_PIPE_THREADS), smem, st, (SRC_T)src, (DST_T)dst, tw, I0, IG, Mid, units, ns, pf, ##__VA_ARGS__);
#endif
  if (!inverse) {
    if (whole && IG == 20) { PO_F2_GO((po_fwd_tx_pipe2_kernel<RT, RX, true, 10>), const float*, float2*) }
    else if (whole) { PO_F2_GO((po_fwd_tx_pipe2_kernel<RT, RX, true, 0>), const float*, float2*) }
    else { PO_F2_GO((po_fwd_tx_pipe2_kernel<RT, RX, false, 0>), const float*, float2*) }
  } else {
    if (IG == 20) { PO_F2_GO((po_inv_xt_pipe2_kernel<RT, RX, true, 10>), const float2*, float*, m_lo, m_cnt) }
    else { PO_F2_GO((po_inv_xt_pipe2_kernel<RT, RX, true, 0>), const float2*, float*, m_lo, m_cnt) }
  }
#undef PO_F2_GO
  return cudaGetLastError();
}

PO_XTU cudaError_t po_launch_fused2(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B,
                                    int num_sms, cudaStream_t st, i64 m_lo = 0, i64 m_cnt = -1) {
  const int RT = nt / 8, RX = nx / 16;
#define PO_FUSED2_CASE(rt, rx) \
  if (RT == rt && RX == rx) return po_launch_fused2_t<rt, rx>(inverse, src, dst, tw, I0, Mid, B, num_sms, st, m_lo, m_cnt);
  PO_FUSED2_CASE(2, 4) PO_FUSED2_CASE(4, 4) PO_FUSED2_CASE(8, 4)
  PO_FUSED2_CASE(2, 8) PO_FUSED2_CASE(4, 8) PO_FUSED2_CASE(8, 8)
#undef PO_FUSED2_CASE
  return cudaErrorNotSupported;
}
#else
cudaError_t po_launch_fused2(bool inverse, int nt, int nx, const void* src, void* dst, const po_fused2_tw& tw, int I0, i64 Mid, i64 B,
                             int num_sms, cudaStream_t st, i64 m_lo = 0, i64 m_cnt = -1);
#endif

// tables: direction `inverse` of the transform pair; kscale per t-mode
static void po_fill_fused2_tw(po_fused2_tw& tw, int nx, bool inverse, const double* kscale) {
  const int RX = nx / 16;
  const double sgn = inverse ? +1.0 : -1.0;
  for (int j2 = 0; j2 < 8; ++j2)
    for (int r = 0; r < 16; ++r) {
      const long long bin = r < 8 ? r : (long long)nx - 16 + r;
      const long long num = ((bin * j2) % nx + nx) % nx;
      double c = 0, s = 0;
      if (j2 < RX) {
        const double ang = 2.0 * M_PI * (double)num / (double)nx;
        c = cos(ang); s = sgn * sin(ang);
        if (num == 0) { c = 1; s = 0; }
        else if (4 * num == nx) { c = 0; s = sgn; }
        else if (2 * num == nx) { c = -1; s = 0; }
        else if (4 * num == 3 * nx) { c = 0; s = -sgn; }
      }
      tw.tw_x[j2 * 16 + r] = make_float2((float)c, (float)s);
    }
  for (int k = 0; k < 8; ++k) tw.kscale[k] = (float)(kscale ? kscale[k] : 1.0);
}
