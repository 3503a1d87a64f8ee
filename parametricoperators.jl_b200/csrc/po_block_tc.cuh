// Block epilogue on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a:
//     y[:, p] = gelu( s[:, p] + W x[:, p] + b )            (examples/fno.jl:34-44, see po_block.cuh)
// The pointwise branch is a real GEMM  D[p, c'] = sum_c X[p, c] W[c', c]  with M = grid points, N = c_out, K = c_in.
// On CUDA cores it costs 20 FMAs per output and makes the kernel FMA/issue-bound at 2.2 TB/s (po_block.cuh);
// here the contraction runs as 3xTF32 (hi*hi + lo*hi + hi*lo: FP32-level accuracy, rel. error ~1e-6) on
// tcgen05.mma with the accumulator in tensor memory, so the SM only stages operands and runs the epilogue:
//   * a 128-point tile of x is split into its TF32-exact high part and the remainder while it is written to
//     shared memory in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B blocks);
//   * one thread issues 3 x (K/8) tcgen05.mma (M = 128, N = 32, K = 8, kind::tf32) into 32 TMEM columns and
//     commits them to an mbarrier;
//   * each of the 4 warps reads its 32 accumulator lanes with tcgen05.ld (32x32b.x32), adds s and the bias,
//     applies gelu and stores its rows.
// No CUTLASS: descriptors are built by hand (bit layout: cute/arch/mma_sm100_desc.hpp of the vendored headers).
#pragma once
#include "po_block.cuh"
#include "po_fused.cuh"  // mbarrier helpers (po_smem_u32, po_mbar_init, po_mbar_try_wait)

#ifndef PO_EMUL
PO_D void po_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PO_D void po_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B (128 B contiguous);
// LBO = byte distance between the two core matrices of one K = 32 B step, SBO = byte distance between 8-row groups
PO_D unsigned long long po_tc_smem_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr >> 4) & 0x3FFFu);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version for sm_100
  return d;          // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
#define PO_TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24))

PO_D void po_tc_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(PO_TC_IDESC), "r"(accumulate)
      : "memory");
}

// one CTA = 128 threads = one 128-point tile at a time; several CTAs per SM (~40 KB of shared memory, 32 TMEM columns each).
// Memory pipeline: the x tile of the NEXT iteration and the s tile of the current one are loaded into registers
// (fully coalesced float4 streams) while the tensor core works; s / y go through a shared-memory tile so that the
// row-per-thread epilogue never touches global memory with an 80-byte lane stride.
template <int KQ>  // float4 per point = c_in / 4 (compile-time: register-resident prefetch)
__global__ void __launch_bounds__(128)
po_block_epilogue_tc_kernel(const float* __restrict__ W, const float* __restrict__ X, const float* __restrict__ S, const float* __restrict__ bias,
                            float* __restrict__ Y, int c_out, i64 npts, int act, int* __restrict__ err) {
  constexpr int c_in = 4 * KQ;
  constexpr int kcols = 2 * ((c_in + 7) / 8);           // 16-byte K columns (4 tf32 each), even: K = 4 * kcols
  constexpr unsigned a_bytes = 16u * kcols * 128u;      // A tile: 16 row groups x kcols core matrices
  constexpr unsigned b_bytes = 4u * kcols * 128u;       // B tile: N = 32 rows
  PO_DYN_SMEM(smem_raw);
  unsigned char* a_hi = smem_raw;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  float* tile = (float*)(b_lo + b_bytes);               // [128][c_out] s in / y out
  unsigned long long* mbar = (unsigned long long*)(tile + 128 * ((c_out + 3) & ~3));
  unsigned* tmem_slot = (unsigned*)(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (unsigned e = tid; e < (2 * a_bytes + 2 * b_bytes) / 16; e += 128) ((float4*)smem_raw)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int e = tid; e < c_in * c_out; e += 128) {       // W is c_out x c_in column-major: B[n = c'][k = c]
    const int cp = e % c_out, c = e / c_out;
    const float w = W[e];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const unsigned off = ((unsigned)(cp >> 3) * kcols + (unsigned)(c >> 2)) * 128u + (unsigned)(cp & 7) * 16u + (unsigned)(c & 3) * 4u;
    *(float*)(b_hi + off) = hi;
    *(float*)(b_lo + off) = w - hi;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(po_smem_u32(tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    po_mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  po_tc_fence_before();
  __syncthreads();
  po_tc_fence_after();
  const unsigned tmem_base = *tmem_slot;
  constexpr unsigned sbo = (unsigned)kcols * 128u;
  const int oq = c_out >> 2;                             // float4 per output row (fast path: c_out % 4 == 0)
  const bool vec_out = (c_out & 3) == 0;

  const i64 ntiles = (npts + 127) / 128;
  unsigned parity = 0;
  float4 xr[KQ];
  auto load_x = [&](i64 tile_idx) {
    const i64 p0 = tile_idx * 128;
    const i64 nv = ((npts - p0 < 128) ? (npts - p0) : 128) * KQ;
    const float4* xg = (const float4*)(X + p0 * c_in);
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
      const int e = tid + 128 * j;
      xr[j] = (e < nv) ? __ldg(xg + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if ((i64)blockIdx.x < ntiles) load_x(blockIdx.x);
  for (i64 tile_i = blockIdx.x; tile_i < ntiles; tile_i += gridDim.x) {
    const i64 p0 = tile_i * 128;
    const int np = (int)((npts - p0 < 128) ? (npts - p0) : 128);
    // ---- stage the x tile: split into TF32-exact high part and remainder, canonical core-matrix layout
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
      const int e = tid + 128 * j;
      const int p = e / KQ, kc = e - p * KQ;
      const float4 v = xr[j];
      float4 h;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
      const unsigned off = ((unsigned)(p >> 3) * kcols + (unsigned)kc) * 128u + (unsigned)(p & 7) * 16u;
      *(float4*)(a_hi + off) = h;
      *(float4*)(a_lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core (async proxy)
    po_tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      po_tc_fence_after();
      unsigned acc = 0;
      const unsigned ah = po_smem_u32(a_hi), al = po_smem_u32(a_lo), bh = po_smem_u32(b_hi), bl = po_smem_u32(b_lo);
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const unsigned aa = (term == 1) ? al : ah, bb = (term == 2) ? bl : bh;   // hi*hi, lo*hi, hi*lo
#pragma unroll
        for (int j = 0; j < kcols / 2; ++j) {
          po_tc_mma_tf32(tmem_base, po_tc_smem_desc(aa + (unsigned)j * 256u, 128u, sbo), po_tc_smem_desc(bb + (unsigned)j * 256u, 128u, sbo), acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(po_smem_u32(mbar)) : "memory");
    }
    // ---- while the tensor core works: s tile of this iteration -> shared tile (coalesced), x tile of the next -> registers
    if (S && vec_out) {
      const float4* sg = (const float4*)(S + p0 * c_out);
      const int nv = np * oq;
      for (int e = tid; e < nv; e += 128) ((float4*)tile)[e] = __ldg(sg + e);
    }
    if (tile_i + gridDim.x < ntiles) load_x(tile_i + gridDim.x);
    __syncthreads();   // s tile visible to the row owners
    // ---- wait for the accumulator (bounded), read this warp's 32 lanes
    {
      int spins = 0;
      while (!po_mbar_try_wait(mbar, parity)) {
        if (++spins > (1 << 22)) { if (tid == 0) *err = 4; break; }
      }
    }
    parity ^= 1u;
    po_tc_fence_after();
    unsigned r[32];
    const unsigned taddr = tmem_base + ((unsigned)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // ---- epilogue: row p = 32 * warp + lane
    const int p = warp * 32 + lane;
    if (vec_out) {
      float4* trow = (float4*)(tile + p * c_out);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < oq) {
          float4 sv = S ? trow[q] : make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) { const float4 bv = __ldg((const float4*)bias + q); sv.x += bv.x; sv.y += bv.y; sv.z += bv.z; sv.w += bv.w; }
          float4 ov = make_float4(__uint_as_float(r[4 * q]) + sv.x, __uint_as_float(r[4 * q + 1]) + sv.y, __uint_as_float(r[4 * q + 2]) + sv.z,
                                  __uint_as_float(r[4 * q + 3]) + sv.w);
          if (act) { ov.x = po_gelu_fast(ov.x); ov.y = po_gelu_fast(ov.y); ov.z = po_gelu_fast(ov.z); ov.w = po_gelu_fast(ov.w); }
          trow[q] = ov;
        }
      }
      po_tc_fence_before();
      __syncthreads();   // y tile complete; accumulator and operand tiles free again
      float4* yg = (float4*)(Y + p0 * c_out);
      const int nv = np * oq;
      for (int e = tid; e < nv; e += 128) yg[e] = ((const float4*)tile)[e];
      __syncthreads();   // tile free for the next iteration's s
    } else {
      if (p < np) {
        const i64 o = (i64)c_out * (p0 + p);
#pragma unroll
        for (int n = 0; n < 32; ++n) {
          if (n < c_out) {
            float v = __uint_as_float(r[n]) + (bias ? bias[n] : 0.f) + (S ? S[o + n] : 0.f);
            Y[o + n] = act ? po_gelu_fast(v) : v;
          }
        }
      }
      po_tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
}
#endif  // !PO_EMUL

// cudaErrorNotSupported: shape / build not handled here (caller uses the CUDA-core kernel of po_block.cuh)
static cudaError_t po_launch_block_epilogue_tc(const void* W, const void* X, const void* S, const void* bias, void* Y, i64 c_in, i64 c_out, i64 npts,
                                               int act, int num_sms, int* err, cudaStream_t st) {
#ifdef PO_EMUL
  (void)W; (void)X; (void)S; (void)bias; (void)Y; (void)c_in; (void)c_out; (void)npts; (void)act; (void)num_sms; (void)err; (void)st;
  return cudaErrorNotSupported;
#else
  static const int env = getenv("PO_BLOCK_TC") ? atoi(getenv("PO_BLOCK_TC")) : 1;
  if (!env || !err || num_sms <= 0 || c_in % 4 || c_in < 4 || c_in > 32 || c_out < 1 || c_out > 32 || ((size_t)X % 16) || npts < 128) return cudaErrorNotSupported;
  if ((c_out % 4 == 0) && ((((size_t)Y) % 16) || (S && ((size_t)S % 16)))) return cudaErrorNotSupported;
  const int kcols = 2 * (((int)c_in + 7) / 8);
  const size_t smem = (size_t)2 * 16 * kcols * 128 + (size_t)2 * 4 * kcols * 128 + (size_t)128 * ((c_out + 3) & ~3) * sizeof(float) + 64;
  const i64 ntiles = (npts + 127) / 128;
  int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));   // CTAs that fit one SM (shared memory; 32 TMEM columns each: <= 16)
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  i64 grid = (i64)num_sms * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (bias && ((size_t)bias % 16) && (c_out % 4 == 0)) return cudaErrorNotSupported;
#define PO_TC_GO(KQ)                                                                                                             \
  case KQ: {                                                                                                                     \
    cudaError_t e = cudaFuncSetAttribute(po_block_epilogue_tc_kernel<KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                              \
    po_block_epilogue_tc_kernel<KQ><<<dim3((unsigned)grid), dim3(128), smem, st>>>((const float*)W, (const float*)X, (const float*)S, (const float*)bias, \
                                                                                  (float*)Y, (int)c_out, npts, act, err);        \
    break;                                                                                                                       \
  }
  switch ((int)(c_in / 4)) {
    PO_TC_GO(1) PO_TC_GO(2) PO_TC_GO(3) PO_TC_GO(4) PO_TC_GO(5) PO_TC_GO(6) PO_TC_GO(7) PO_TC_GO(8)
    default: return cudaErrorNotSupported;
  }
#undef PO_TC_GO
  return cudaGetLastError();
#endif
}
