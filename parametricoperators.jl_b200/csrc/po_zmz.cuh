// Fused "z - mix - z" step of the FNO spectral stack:
//     (R o fft_z)  ->  ParDiagonal (x) ParMatrix frequency weighting  ->  (ifft_z o R')
// i.e. the three plan steps  pruned-c16 forward on the slowest spatial dim, mix+diag, pruned-c16 inverse
// (src/ParDFT.jl:26,30 + src/ParRestriction.jl:13-39 + src/ParMatrix.jl:42-46 + src/ParDiagonal.jl:24-31 as
// composed by examples/fno.jl:9-32) in ONE kernel.  The 21 MB input is read once, the 5 MB spectral tensor
// never leaves the SM (it is only written out when the caller asked for a tape), the 21 MB output is written
// once; unfused this is three latency-bound launches (8 + 17 + 8 us at C3) and in reverse mode seven
// (mix/diag cotangents, Gram matrix, mode sums: 93 us).
//
// Work decomposition: a CTA owns G consecutive "modes" (kx, ky) of one (kt, batch) slab and ALL channels, and
// every z-column is shared by R = nz/16 threads so that the kernel -- which is latency-bound, not bandwidth-bound:
// 26 MB of traffic -- has R x more threads in flight and R x shorter dependent chains:
//   1  thread (c, g, j2): residue-j2 sub-FFT16 of the column + twiddle            -> partial spectra P (shared)
//   1b thread (c, g, q):  sums the R partials of its 16/R kept bins               -> x-hat S1 (shared) [-> tape]
//   2  thread (c', g, q): channel mixing B(c', :) . x-hat and the diagonal        -> y-hat Z2 (shared)
//   3  thread (c', g, j2): pre-twiddle + FFT16 of residue j2                      -> 16 output rows
// For fixed z the CTA touches G*C consecutive complex numbers: every global access is a full-sector segment.
//
// The same kernel runs the reverse mode (PULLBACK): the cotangent goes through the adjoint z-transform, the
// diagonal is applied BEFORE the (conjugate-transposed) mixing, and the parameter cotangents
//     theta_bar = u_bar x^H (or x u_bar^H),   d_bar = sum_c conj(AA x) .* y_bar
// are formed from the taped spectral input x: d_bar by shared-memory partial sums (+ one atomicAdd per CTA and
// entry, which only collide across the batch), theta_bar as one partial Gram matrix per CTA that a second tiny
// kernel reduces in a fixed order (deterministic, and no 512-way same-address atomics).
#pragma once
#include "po_fused.cuh"

struct po_zmz_params {
  const float2* X;      // (nin * Mi, nz, O)
  float2* Y;            // (nout * Mi, nz, O)
  const float2* A;      // mixing matrix as stored by the caller (ParMatrix theta)
  long long b_rs, b_cs; // B(c', c) = A[c' * b_rs + c * b_cs] (conj optional): the matrix APPLIED in this pass, nout x nin
  int conj_b;
  const float2* d;      // diagonal (ParDiagonal theta), length nd
  int conj_d;
  int diag_first;       // 0: y = dd .* (B x)   (apply)      1: y = B (dd .* x)   (pullback)
  long long nd;
  int nin, nout;
  long long Mi;         // modes per (kz, o) slab = I / channels
  long long O;
  int G;                // modes per CTA (divides Mi)
  // apply only
  float2* tape_out;     // (nin, Mi, 16, O) spectral input of the mixing, or null
  // pullback only
  const float2* xhat;   // taped spectral input of the forward pass, (nfw_in, Mi, 16, O)
  float2* gram_part;    // [CTAs][nin * nout] partial Gram matrices (cr-major: entry cr * nout + c), or null
  float2* gradd;        // cotangent of the ParDiagonal parameter (nd, zeroed by the caller) or null
  int fw_adjoint;       // forward step used theta' (s.adjoint)
  int fw_adjoint2;      // forward step used conj(d) (s.adjoint2)
  long long a_rs, a_cs; // forward-pass view AA(c', c) = A[c' * a_rs + c * a_cs] (conj: fw_conj_a), nfw_out x nfw_in
  int fw_conj_a;
};

// NCH > 0: nin == nout == NCH and G == GT are compile-time constants (shared-memory indices fold into immediates,
// the thread-role divisions disappear: the generic instantiation spends 2/3 of its instructions on them)
template <int R, bool PULLBACK, int NCH, int GT>
__global__ void __launch_bounds__(NCH ? 320 : 512, (NCH && !PULLBACK) ? 4 : 1)   // c = 20: 320 threads, 4 CTAs / SM = all 512 CTAs of C3 in one wave
po_zmz_kernel(const po_zmz_params P, const po_c16_tw tw_in, const po_c16_tw tw_out) {
  constexpr int NZ = 16 * R, NB = 16 / R;  // bins per thread in the per-bin phases
  PO_DYN_SMEM(smem_raw);
  const int nin = NCH ? NCH : P.nin, nout = NCH ? NCH : P.nout, G = NCH ? GT : P.G;
  const int NT = blockDim.x;
  float2* Pp = (float2*)smem_raw;                       // [R][16][G][nin]  partial spectra; re-used as Z2 [16][G][nout]
  float2* S1 = Pp + (size_t)R * 16 * G * (nin > nout ? nin : nout);  // [16][G][nin]
  float2* Bs = S1 + 16 * (size_t)G * nin;               // [nin][nout]      B(c', c) at Bs[c * nout + c']
  float2* twi = Bs + (size_t)nin * nout;                // [R][16] forward-type twiddles
  float2* two = twi + R * 16;                           // [R][16] inverse-type twiddles
  float2* S2 = two + R * 16;                            // pullback: [16][G][nout] taped x-hat
  float2* As2 = S2 + (PULLBACK ? 16 * (size_t)G * nout : 0);            // pullback: forward matrix AA(c'', c) at As2[c * nin + c'']
  float2* tbuf = As2 + (PULLBACK ? (size_t)nin * nout : 0);             // pullback: [16][G][nin] per-channel d-bar terms
  float2* Z2 = Pp;
  const int tid = threadIdx.x;
  const long long cta = blockIdx.x;
  const long long per_o = P.Mi / G;
  const long long o = cta / per_o;
  const long long mo0 = (cta - o * per_o) * G;
  po_pdl_trigger();
  po_pdl_wait();  // the parameters themselves may have been written by the previous kernel on the stream

  for (int e = tid; e < nin * nout; e += NT) {
    const int cp = e % nout, c = e / nout;
    float2 v = P.A[(long long)cp * P.b_rs + (long long)c * P.b_cs];
    Bs[e] = P.conj_b ? po_conj(v) : v;
  }
  for (int e = tid; e < R * 16; e += NT) { twi[e] = tw_in.tw[e]; two[e] = tw_out.tw[e]; }
  if (PULLBACK) {
    for (int e = tid; e < nin * nout; e += NT) {   // AA is nin(fw-out) x nout(fw-in)
      const int cpp = e % nin, c = e / nin;
      float2 v = P.A[(long long)cpp * P.a_rs + (long long)c * P.a_cs];
      As2[e] = P.fw_conj_a ? po_conj(v) : v;
    }
    for (int e = tid; e < 16 * G * nout; e += NT) {  // taped x-hat of the forward pass: channels = nout here
      const int c = e % nout, g = (e / nout) % G, r = e / (nout * G);
      S2[e] = P.xhat[c + (long long)nout * (mo0 + g + P.Mi * (r + 16 * o))];
    }
  }

  // thread roles: input side (c1, g1, s1), output side (c2, g2, s2); s = residue j2 or bin group q
  const int c1 = tid % nin, g1 = (tid / nin) % G, s1 = tid / (nin * G);
  const bool act1 = tid < R * G * nin;
  const int c2 = tid % nout, g2 = (tid / nout) % G, s2 = tid / (nout * G);
  const bool act2 = tid < R * G * nout;
  __syncthreads();

  // ---- phase 1: residue-s1 sub-FFT of the input column
  if (act1) {
    const long long Iin = (long long)nin * P.Mi;
    const float2* xp = P.X + (c1 + (long long)nin * (mo0 + g1)) + Iin * ((long long)NZ * o + s1);
    float2 v[16];
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) v[j1] = __ldg(xp + Iin * (long long)(R * j1));
    po_fft16<-1>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) Pp[(((size_t)s1 * 16 + r) * G + g1) * nin + c1] = po_mul(v[po_brev4(r)], twi[s1 * 16 + r]);
  }
  __syncthreads();
  // ---- phase 1b: bins r = s1*NB .. +NB of column (c1, g1)
  float2 vin[NB];  // pullback: the cotangent BEFORE the diagonal (needed again for d-bar)
#pragma unroll
  for (int i = 0; i < NB; ++i) vin[i] = make_float2(0.f, 0.f);
  if (act1) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int r = s1 * NB + i;
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int j2 = 0; j2 < R; ++j2) acc = po_add(acc, Pp[(((size_t)j2 * 16 + r) * G + g1) * nin + c1]);
      vin[i] = acc;
      const long long col = mo0 + g1 + P.Mi * (r + 16 * o);
      if (!PULLBACK && P.tape_out) P.tape_out[c1 + (long long)nin * col] = acc;
      if (P.diag_first) {
        float2 dd = P.d[col % P.nd];
        if (P.conj_d) dd = po_conj(dd);
        acc = po_mul(dd, acc);
      }
      S1[((size_t)r * G + g1) * nin + c1] = acc;
    }
  }
  __syncthreads();  // S1 complete, P free (re-used as Z2)

  // ---- phase 2: channel mixing for bins r = s2*NB .. of output column (c2, g2)
  if (act2) {
    float2 z[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) z[i] = make_float2(0.f, 0.f);
    for (int c = 0; c < nin; ++c) {
      const float2 b = Bs[c * nout + c2];
#pragma unroll
      for (int i = 0; i < NB; ++i) po_mac(z[i], b, S1[((size_t)(s2 * NB + i) * G + g2) * nin + c]);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int r = s2 * NB + i;
      if (!P.diag_first) {
        float2 dd = P.d[(mo0 + g2 + P.Mi * (r + 16 * o)) % P.nd];
        if (P.conj_d) dd = po_conj(dd);
        z[i] = po_mul(dd, z[i]);
      }
      Z2[((size_t)r * G + g2) * nout + c2] = z[i];
    }
  }
  if (PULLBACK) {
    // ---- parameter cotangents.  Forward pass: y = dd .* (AA x), x = taped (nout channels here), y has nin channels.
    //      S1 = u_bar = conj(dd) .* y_bar [nin channels], vin = y_bar for bins s1*NB.. of (c1, g1), S2 = x.
    //      No atomics in shared memory (float atomics there are CAS loops): per-channel terms go through tbuf.
    if (act1 && P.gradd) {
      float2 u[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) u[i] = make_float2(0.f, 0.f);
      for (int c = 0; c < nout; ++c) {
        const float2 a = As2[c * nin + c1];
#pragma unroll
        for (int i = 0; i < NB; ++i) po_mac(u[i], a, S2[((size_t)(s1 * NB + i) * G + g1) * nout + c]);
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        // d .* x: d_bar = sum conj(u) y_bar;  conj(d) .* x: d_bar = sum conj(y_bar) u
        tbuf[((size_t)(s1 * NB + i) * G + g1) * nin + c1] = !P.fw_adjoint2 ? po_mul(po_conj(u[i]), vin[i]) : po_mul(po_conj(vin[i]), u[i]);
      }
    }
    if (P.gram_part) {
      // partial Gram matrix of this CTA: entry (cr, c) = sum over its 16*G (bin, mode) pairs -- a 64-term dot product per thread
      float2* gp = P.gram_part + cta * (long long)(nin * nout);
      for (int e = tid; e < nin * nout; e += NT) {
        const int cr = e / nout, c = e % nout;
        float2 acc = make_float2(0.f, 0.f);
        for (int rg = 0; rg < 16 * G; ++rg) {
          const float2 ub = S1[(size_t)rg * nin + cr], x = S2[(size_t)rg * nout + c];
          // theta * x: theta_bar(cr, c) += u_bar conj(x);   theta' * x: theta_bar(c, cr) += x conj(u_bar)
          if (!P.fw_adjoint) po_mac(acc, ub, po_conj(x)); else po_mac(acc, x, po_conj(ub));
        }
        gp[e] = acc;
      }
    }
  }
  __syncthreads();

  // ---- phase 3: residue-s2 part of the zero-padded inverse of output column (c2, g2)
  if (act2) {
    const long long Iout = (long long)nout * P.Mi;
    float2* yp = P.Y + (c2 + (long long)nout * (mo0 + g2)) + Iout * ((long long)NZ * o + s2);
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = po_mul(Z2[((size_t)r * G + g2) * nout + c2], two[s2 * 16 + r]);
    po_fft16<1>(v);
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) yp[Iout * (long long)(R * j1)] = v[po_brev4(j1)];
  }
  if (PULLBACK && P.gradd) {
    for (int e = tid; e < 16 * G; e += NT) {  // e = r * G + g
      const int r = e / G, g = e % G;
      float2 acc = make_float2(0.f, 0.f);
      for (int c = 0; c < nin; ++c) acc = po_add(acc, tbuf[(size_t)e * nin + c]);
      const long long j = (mo0 + g + P.Mi * (r + 16 * o)) % P.nd;
      atomicAdd(&P.gradd[j].x, acc.x);  // collides only across the batch
      atomicAdd(&P.gradd[j].y, acc.y);
    }
  }
}

// theta_bar = sum over CTAs of the partial Gram matrices, in CTA order (deterministic); entry cr * nout + c of a
// partial is (forward output channel cr, forward input channel c); theta is prm_rows x prm_cols column-major.
__global__ void __launch_bounds__(256)
po_zmz_gram_reduce_kernel(const float2* __restrict__ part, long long nparts, int nfo, int nfi, int fw_adjoint, int prm_rows, float2* __restrict__ gradA) {
  // one warp per entry: lane l sums partials l, l+32, ... (all loads independent), then a fixed-order butterfly
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool ok = e < nfo * nfi;  // warp-uniform
  const long long ld = (long long)nfo * nfi;
  po_pdl_trigger();
  po_pdl_wait();
  float2 acc = make_float2(0.f, 0.f);
  if (ok)
    for (long long p = lane; p < nparts; p += 32) acc = po_add(acc, part[p * ld + e]);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    acc.x += PO_SHFL_XOR(acc.x, m);
    acc.y += PO_SHFL_XOR(acc.y, m);
  }
  if (ok && lane == 0) {
    const int cr = e / nfi, c = e % nfi;
    const long long idx = !fw_adjoint ? (long long)cr + (long long)prm_rows * c : (long long)c + (long long)prm_rows * cr;
    gradA[idx] = acc;
  }
}

static inline int po_zmz_group(int nz, int nin, int nout, long long Mi) {
  const int cm = nin > nout ? nin : nout, R = nz / 16;
  int G = 512 / (R * cm);
  if (G > 8) G = 8;
  while (G > 1 && (Mi % G)) --G;
  return G < 1 ? 1 : G;
}
static inline bool po_zmz_ok(long long nz, long long nin, long long nout, long long Mi) {
  if (!(nz == 32 || nz == 64 || nz == 128) || nin < 1 || nout < 1 || Mi < 1) return false;
  const long long cm = nin > nout ? nin : nout;
  return (nz / 16) * cm <= 512;
}
static inline long long po_zmz_ctas(int nz, int nin, int nout, long long Mi, long long O) { return (Mi / po_zmz_group(nz, nin, nout, Mi)) * O; }
static inline size_t po_zmz_smem(bool pullback, int R, int nin, int nout, int G) {
  const int cm = nin > nout ? nin : nout;
  size_t s = ((size_t)R * 16 * G * cm + (size_t)16 * G * nin + (size_t)nin * nout + 2 * (size_t)R * 16) * sizeof(float2);
  if (pullback) s += ((size_t)16 * G * nout + (size_t)nin * nout + (size_t)16 * G * nin) * sizeof(float2);
  return s;
}

static cudaError_t po_launch_zmz(bool pullback, int nz, po_zmz_params P, const po_c16_tw& tw_in, const po_c16_tw& tw_out, cudaStream_t st) {
  if (P.Mi * P.O == 0) return cudaSuccess;
  const int R = nz / 16;
  P.G = po_zmz_group(nz, P.nin, P.nout, P.Mi);
  const size_t smem = po_zmz_smem(pullback, R, P.nin, P.nout, P.G);
  const long long ctas = (P.Mi / P.G) * P.O;
  if (ctas > 0x7fffffffLL || smem > 200 * 1024) return cudaErrorInvalidValue;
  const int cm = P.nin > P.nout ? P.nin : P.nout;
  const int nthreads = ((R * P.G * cm + 31) / 32) * 32;
#ifndef PO_EMUL
#define PO_ZMZ_ATTR(K) { cudaError_t e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
#else
#define PO_ZMZ_ATTR(K)
#endif
#define PO_ZMZ_GO(K) { PO_ZMZ_ATTR(K) PO_LAUNCH_PDL(K, dim3((unsigned)ctas), dim3(nthreads), smem, st, P, tw_in, tw_out); return cudaGetLastError(); }
#define PO_ZMZ_CASE(r, g20)                                                                                         \
  if (R == r) {                                                                                                     \
    if (P.nin == 20 && P.nout == 20 && P.G == g20) {                                                                \
      if (!pullback) PO_ZMZ_GO((po_zmz_kernel<r, false, 20, g20>)) else PO_ZMZ_GO((po_zmz_kernel<r, true, 20, g20>)) \
    }                                                                                                               \
    if (!pullback) PO_ZMZ_GO((po_zmz_kernel<r, false, 0, 0>)) else PO_ZMZ_GO((po_zmz_kernel<r, true, 0, 0>))        \
  }
  PO_ZMZ_CASE(2, 8) PO_ZMZ_CASE(4, 4) PO_ZMZ_CASE(8, 2)
#undef PO_ZMZ_CASE
#undef PO_ZMZ_GO
#undef PO_ZMZ_ATTR
  return cudaErrorInvalidValue;
}
