// Config C2's factor pair in ONE kernel per direction (K = (R[1:16] o rDFT128) (x) DFT128 (x) DFT128 on a
// (128^3, batch) Float32 tensor; src/ParKron.jl:190-220 applies the restricted real DFT on the SLOWEST dim first, then the two
// complex DFTs; the adjoint runs the mirror image).
//
//   forward : pruned real DFT along the slow dim (n = 16 R rows -> 16 kept bins, po_r16_mode_kernel's arithmetic)
//             FUSED WITH the full complex FFT along the FASTEST dim (length 128 = one 128-thread CTA's worth of columns):
//             a CTA owns one contiguous row of 128 columns, so after the column-wise reduction it holds the whole fast
//             extent of 16 modes (16 KB) and runs their FFT128s in shared memory before anything is written;
//   inverse : FFT128^-1 along the fastest dim of a (128 x 16 modes) block in shared memory, then the zero-padded c2r along the
//             slow dim straight out of shared memory.
// Each direction saves one full read + write of the 33.5 MB complex intermediate (the separate fft16x8 pass: 20.7 us of the
// 82 us apply in profiles/r2g_secondary_configs.json) and the reduction's loads are software-pipelined: the next residue pair's
// 32 rows are in flight while the current FFT16 runs (the one-shot version alternated "32 loads outstanding" with "none":
// 3.47 TB/s).
//
// Shared-memory layouts (complex words): S[r][136] -- row stride 136 = 8 * 128 B + 64 B, so the two rows a half-warp touches in
// the stride-8 accesses of stage 1 sit on disjoint bank halves; T[r][j2][17] -- the 17-padding makes the transposing write
// (lanes along j2) and the read (lanes along k1) both conflict-free (same trick as po_cfull_mode_kernel's S[c][N + 1]).
#pragma once
#include "po_fused.cuh"

#define PO_C2_LD 136
#define PO_C2_DEFAULT_VARIANT 3  // measured (profiles/r2m_c2_fused_pair_variants.json): 0: 50.0 / 42.6 us, 1: 48.0 / 39.7, 3: 44.2 / 39.8 (forward / adjoint apply of C2)

// 16 FFT128s (one per row r of S) by 128 threads; result in natural order back in S.  twc[j2 * 16 + k1] = cis(-+ 2 pi j2 k1 / 128).
// Entered and left with the CTA synchronised by the caller (S complete on entry; S complete after the trailing barrier).
template <int INVERSE>
PO_D void po_c2_fft128_rows(float2* __restrict__ S, float2* __restrict__ T, const po_c16_tw& twc, int tid, float scale) {
  {  // stage 1: (r, j2): FFT16 over j1 of S[r][8 j1 + j2], twiddle, transposed into T[r][j2][k1]
    const int r = tid >> 3, j2 = tid & 7;
    float2 v[16];
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) v[j1] = S[r * PO_C2_LD + 8 * j1 + j2];
    if (INVERSE) po_fft16<1>(v); else po_fft16<-1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) T[r * PO_C2_LD + j2 * 17 + k1] = po_mul(v[po_brev4(k1)], twc.tw[j2 * 16 + k1]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // stage 2: (r, k1): FFT8 over j2 -> X[k1 + 16 k2]
    const int item = tid + 128 * q, r = item >> 4, k1 = item & 15;
    float2 u[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) u[t] = T[r * PO_C2_LD + t * 17 + k1];
    if (INVERSE) po_fft8<1>(u); else po_fft8<-1>(u);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) S[r * PO_C2_LD + k1 + 16 * k2] = po_scale(u[po_brev3(k2)], scale);
  }
  __syncthreads();
}

// forward: X real (I, 16 R, O) -> Y complex (I, 16, O) with the FFT128 along the fastest 128 entries of I applied; one CTA per
// run of 128 consecutive columns (I % 128 == 0)
// IC: compile-time row pitch (0 = run time), MINB: CTAs per SM the register allocation is capped for (4: 128 registers = the prefetched
// rows + 16 accumulators + one FFT16 in flight; 5: 102 registers)
template <int R, int IC, int MINB>
__global__ void __launch_bounds__(128, MINB)
po_c2_fwd_kernel(const float* __restrict__ X, float2* __restrict__ Y, const po_c16_tw twr, const po_c16_tw twc, i64 I_rt, int keep) {
  const i64 I = IC ? (i64)IC : I_rt;
  __shared__ float2 S[16 * PO_C2_LD];
  __shared__ float2 T[16 * PO_C2_LD];
  const int tid = threadIdx.x;
  const i64 col0 = (i64)blockIdx.x * 128;
  const i64 o = col0 / I, i0 = col0 - o * I;
  constexpr int N = 16 * R;
  {
    float2 acc[16];
    po_r16_fwd_column<R, IC>(X + i0 + tid + I * (i64)N * o, I, twr, acc);
#pragma unroll
    for (int r = 0; r < 16; ++r) S[r * PO_C2_LD + tid] = acc[r];
  }
  __syncthreads();
  po_c2_fft128_rows<0>(S, T, twc, tid, 1.f);
  const unsigned long long pol = keep ? po_policy_evict_last() : 0ull;  // the next (strided) FFT pass reads this out of L2
  float2* yp = Y + i0 + tid + I * 16 * o;
#pragma unroll
  for (int r = 0; r < 16; ++r) { if (keep) po_st_hint(yp + I * r, S[r * PO_C2_LD + tid], pol); else yp[I * r] = S[r * PO_C2_LD + tid]; }
}

// inverse: X complex (I, 16, O) -> FFT128^-1 along the fastest 128 entries of I (scale 1/128) -> zero-padded c2r along the mode
// dim -> Y real (I, 16 R, O).  twr = the c2r table of po_r16_mode_kernel (weights c_b / n folded in), twc = cis(+2 pi j2 k1 / 128).
template <int R, int IC>
__global__ void __launch_bounds__(128)
po_c2_inv_kernel(const float2* __restrict__ X, float* __restrict__ Y, const po_c16_tw twr, const po_c16_tw twc, i64 I_rt) {
  const i64 I = IC ? (i64)IC : I_rt;
  __shared__ float2 S[16 * PO_C2_LD];
  __shared__ float2 T[16 * PO_C2_LD];
  const int tid = threadIdx.x;
  const i64 col0 = (i64)blockIdx.x * 128;
  const i64 o = col0 / I, i0 = col0 - o * I;
  constexpr int N = 16 * R;
  const float2* xp = X + i0 + tid + I * 16 * o;
#pragma unroll
  for (int r = 0; r < 16; ++r) S[r * PO_C2_LD + tid] = __ldg(xp + I * r);
  __syncthreads();
  po_c2_fft128_rows<1>(S, T, twc, tid, 1.f / 128.f);
  float2 z[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) z[r] = S[r * PO_C2_LD + tid];
  float* yp = Y + i0 + tid + I * (i64)N * o;
  auto residue = [&](int j2) {
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = po_mul(z[r], twr.tw[j2 * 16 + r]);
    po_fft16<1>(v);
#pragma unroll
    for (int j1 = 0; j1 < 16; ++j1) yp[I * (i64)(R * j1 + j2)] = v[po_brev4(j1)].x;
  };
  if constexpr (IC != 0) {  // compile-time pitch: unrolled over the residues, every store address an immediate, twiddles constant-bank operands
#pragma unroll
    for (int j2 = 0; j2 < R; ++j2) residue(j2);
  } else {
#pragma unroll 1
    for (int j2 = 0; j2 < R; ++j2) residue(j2);
  }
}

// n = length of the real transform (32 / 64 / 128); the fused FFT is always 128 long.  cudaErrorNotSupported: shape not handled.
static cudaError_t po_launch_c2(bool inverse, int n, const void* src, void* dst, const po_c16_tw& twr, const po_c16_tw& twc, i64 I, i64 O, cudaStream_t st) {
  if (I % 128 || (n != 32 && n != 64 && n != 128)) return cudaErrorNotSupported;
  const i64 nblk = (I / 128) * O;
  if (nblk == 0) return cudaSuccess;
  if (nblk > 0x7fffffffLL) return cudaErrorNotSupported;
  static const int hints = getenv("PO_L2_HINTS") ? atoi(getenv("PO_L2_HINTS")) : 1;
  const bool trace = getenv("PO_TRACE") != nullptr;  // per call: the tests that assert on this line may not be the first users of the kernel in their process
  if (trace) fprintf(stderr, "[parops] fused C2 pair %s n=%d I=%lld O=%lld blocks=%lld\n", inverse ? "fft128^-1 -> c2r" : "r2c -> fft128", n, (long long)I, (long long)O, (long long)nblk);
  dim3 grid((unsigned)nblk), block(128);
  const int R = n / 16;
  // PO_C2_VARIANT (A/B, read per call): 0 = run-time row pitch; 1 = compile-time pitch where one is instantiated; 3 = 1 + the forward kernel
  // capped at 96 registers (5 CTAs per SM).  Measured and dropped: 6 CTAs per SM at 80 registers (54.4 / 45.0 us at C2: six 36 KB CTAs leave
  // the SM no L1) and the register cap with the run-time pitch (spills, 63.3 us).
  const char* venv = getenv("PO_C2_VARIANT");
  const int variant = venv ? atoi(venv) : PO_C2_DEFAULT_VARIANT;
  if (R == 8) {
    const bool five = (variant & 2) != 0;
    const int ic = !(variant & 1) ? 0 : (I == 16384 || I == 8192 || I == 4096) ? (int)I : 0;  // 128 x {128, 64, 32}: every row offset fits the 24-bit immediate
#define PO_C2_FWD(IC, MINB) PO_LAUNCH((po_c2_fwd_kernel<8, IC, MINB>), grid, block, 0, st, (const float*)src, (float2*)dst, twr, twc, I, hints ? 1 : 0)
#define PO_C2_INV(IC) PO_LAUNCH((po_c2_inv_kernel<8, IC>), grid, block, 0, st, (const float2*)src, (float*)dst, twr, twc, I)
#define PO_C2_IC(IC)                                                     \
  if (ic == IC) {                                                        \
    if (inverse) { PO_C2_INV(IC); }                                      \
    else if (five && IC) { PO_C2_FWD(IC, (IC ? 5 : 4)); }                \
    else { PO_C2_FWD(IC, 4); }                                           \
    return cudaGetLastError();                                           \
  }
    PO_C2_IC(16384) PO_C2_IC(8192) PO_C2_IC(4096) PO_C2_IC(0)
#undef PO_C2_IC
#undef PO_C2_INV
#undef PO_C2_FWD
  }
#define PO_C2_CASE(r)                                                                                                              \
  if (R == r) {                                                                                                                    \
    if (!inverse) { PO_LAUNCH((po_c2_fwd_kernel<r, 0, 4>), grid, block, 0, st, (const float*)src, (float2*)dst, twr, twc, I, hints ? 1 : 0); } \
    else { PO_LAUNCH((po_c2_inv_kernel<r, 0>), grid, block, 0, st, (const float2*)src, (float*)dst, twr, twc, I); }                \
    return cudaGetLastError();                                                                                                     \
  }
  PO_C2_CASE(2) PO_C2_CASE(4)
#undef PO_C2_CASE
  return cudaErrorNotSupported;
}
