// Store-path micro-benchmark for the inverse fused kernel's pattern (32 rows of 5120 B per unit, 21 MB apart):
// which store flavour / width / burst shape reaches the HBM write ceiling on B200?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_store mb_store.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef long long i64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ void st64(float2* p, float2 v) {
  if (MODE == 0) *p = v;
  else if (MODE == 1) __stcs(p, v);
  else if (MODE == 2) __stwt(p, v);
  else if (MODE == 3) __stcg(p, v);
  else asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
// MODE: store flavour; ORDER 0: rows 0..31 in order, 1: j2-major (0,4,..,28,1,5,..) like the kernel; WORK: fake FMA work per 8-store burst
template <int MODE, int ORDER, int WORK>
__global__ void __launch_bounds__(512, 1) rows_stg(float2* __restrict__ Y, int L, i64 Mid, int nthr, float seed) {
  const unsigned rs = (unsigned)((i64)L * Mid / 2);
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    if ((int)threadIdx.x < nthr) {
      for (int v = threadIdx.x; v < L / 2; v += nthr) {
        const unsigned off0 = (unsigned)((i64)L * u / 2) + v;
        float a = seed + v, b = seed - v;
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
#pragma unroll 1
          for (int w = 0; w < WORK; ++w) { a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a); }
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) {
            const int row = ORDER ? (4 * j1 + j2) : (8 * j2 + j1);
            st64<MODE>(Y + off0 + rs * row, make_float2(a + j1, b));
          }
        }
      }
    }
    __syncthreads();
  }
}
// closer mimic of po_inv_tphase2: 8 LDS.128 per column, then per j2-step NW independent packed FMAs (ILP 8) and 8 stores
template <int NW>
__global__ void __launch_bounds__(512, 1) rows_mimic(float2* __restrict__ Y, int L, i64 Mid, int nthr, float seed) {
  extern __shared__ __align__(16) float T[];
  const unsigned rs = (unsigned)((i64)L * Mid / 2);
  for (int e = threadIdx.x; e < 16 * L; e += blockDim.x) T[e] = seed + e;
  __syncthreads();
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    if ((int)threadIdx.x < nthr) {
      for (int v = threadIdx.x; v < L / 2; v += nthr) {
        const unsigned off0 = (unsigned)((i64)L * u / 2) + v;
        float2 r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { float4 t4 = *(const float4*)(T + k * 2 * L + 4 * v); r[k] = make_float2(t4.x + t4.z, t4.y + t4.w); }
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          float2 o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = r[k];
#pragma unroll
          for (int w = 0; w < NW / 8; ++w) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = __ffma2_rn(o[k], make_float2(1.0001f, 1.0001f), r[(k + w + j2) & 7]);
          }
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) Y[off0 + rs * (4 * j1 + j2)] = o[j1];
        }
      }
    }
    __syncthreads();
  }
}
// 128-bit stores: thread owns two adjacent pair-columns
template <int MODE, int WORK>
__global__ void __launch_bounds__(512, 1) rows_stg128(float4* __restrict__ Y, int L, i64 Mid, int nthr, float seed) {
  const unsigned rs = (unsigned)((i64)L * Mid / 4);
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    if ((int)threadIdx.x < nthr) {
      for (int v = threadIdx.x; v < L / 4; v += nthr) {
        const unsigned off0 = (unsigned)((i64)L * u / 4) + v;
        float a = seed + v, b = seed - v;
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
#pragma unroll 1
          for (int w = 0; w < WORK; ++w) { a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, a); }
#pragma unroll
          for (int j1 = 0; j1 < 8; ++j1) {
            float4 val = make_float4(a + j1, b, a, b + j1);
            float4* p = Y + off0 + rs * (4 * j1 + j2);
            if (MODE == 0) *p = val; else __stcs(p, val);
          }
        }
      }
    }
    __syncthreads();
  }
}

template <class F>
static float timeit(F f, int reps = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}
int main() {
  const int L = 1280; const i64 Mid = 4096; const int nrow = 32;
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const size_t bytes = (size_t)L * Mid * nrow * 4;
  float* Y; CK(cudaMalloc(&Y, bytes));
  const double gb = bytes / 1e9;
  float ms;
#define RUN(MODE, ORDER, WORK, NTHR) ms = timeit([&] { rows_stg<MODE, ORDER, WORK><<<nsm, 512>>>((float2*)Y, L, Mid, NTHR, 1.f); }); \
  printf("stg64 mode=%d order=%d work=%d nthr=%d  %.4f ms %.0f GB/s\n", MODE, ORDER, WORK, NTHR, ms, gb / ms * 1e3);
  RUN(0, 0, 0, 320) RUN(0, 1, 0, 320) RUN(1, 1, 0, 320) RUN(2, 1, 0, 320) RUN(3, 1, 0, 320) RUN(4, 1, 0, 320)
  RUN(0, 1, 20, 320) RUN(0, 1, 60, 320) RUN(0, 1, 120, 320) RUN(1, 1, 60, 320) RUN(4, 1, 60, 320)
  RUN(0, 1, 60, 512) RUN(0, 1, 60, 160)
#define RUN128(MODE, WORK, NTHR) ms = timeit([&] { rows_stg128<MODE, WORK><<<nsm, 512>>>((float4*)Y, L, Mid, NTHR, 1.f); }); \
  printf("stg128 mode=%d work=%d nthr=%d  %.4f ms %.0f GB/s\n", MODE, WORK, NTHR, ms, gb / ms * 1e3);
  RUN128(0, 0, 320) RUN128(1, 0, 320) RUN128(0, 60, 320) RUN128(1, 60, 320) RUN128(0, 120, 320) RUN128(0, 60, 160)
#define RUNM(NW, NTHR) { size_t sm = (size_t)16 * L * 4; CK(cudaFuncSetAttribute(rows_mimic<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
  ms = timeit([&] { rows_mimic<NW><<<nsm, 512, sm>>>((float2*)Y, L, Mid, NTHR, 1.f); }); printf("mimic packed-fma=%d nthr=%d  %.4f ms %.0f GB/s\n", NW, NTHR, ms, gb / ms * 1e3); }
  RUNM(8, 320) RUNM(40, 320) RUNM(80, 320) RUNM(120, 320) RUNM(80, 480) RUNM(80, 512)
  ms = timeit([&] { cudaMemsetAsync(Y, 0, bytes); }); printf("memset %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3);
  CK(cudaGetLastError());
  return 0;
}
