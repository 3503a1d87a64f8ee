// Micro-benchmarks that bound the fused (t,x) kernels' memory patterns on the B200 (no transform maths):
//   read_lin / write_lin : plain grid-stride 128-bit streams over the whole tensor (pure read / pure write ceilings)
//   rows_ldg             : the forward kernel's pattern -- persistent CTAs, unit = NROW rows of ROWB bytes at stride
//                          `rstride`, register-staged LDG.64 bursts (32 per thread), like po_fwd_tphase
//   rows_tma<D>          : same pattern through a D-stage shared-memory ring of 8-row groups filled by cp.async.bulk
//   rows_stg             : the inverse kernel's store pattern (STG.64 per thread, 8 rows per burst)
//   rows_tma_store       : same rows written from shared memory with cp.async.bulk stores
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_stream mb_stream.cu ; run: ./mb_stream
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef long long i64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void read_lin(const float4* __restrict__ x, i64 n, float* out) {
  float acc = 0.f;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x * 4) {
    float4 a = __ldg(x + i), b = make_float4(0, 0, 0, 0), c = b, d = b;
    const i64 s = (i64)gridDim.x * blockDim.x;
    if (i + s < n) b = __ldg(x + i + s);
    if (i + 2 * s < n) c = __ldg(x + i + 2 * s);
    if (i + 3 * s < n) d = __ldg(x + i + 3 * s);
    acc += a.x + b.y + c.z + d.w;
  }
  if (acc == 123.456f) *out = acc;
}
__global__ void write_lin(float4* __restrict__ y, i64 n) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) y[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

// rows pattern: tensor = [rowlen = L floats][Mid][NROW], unit m: rows t at X + L*m + L*Mid*t
__global__ void __launch_bounds__(512, 1) rows_ldg(const float2* __restrict__ X, int L, i64 Mid, int nrow, int nthr, float* out) {
  const unsigned rs = (unsigned)((i64)L * Mid / 2);
  float acc = 0.f;
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    if ((int)threadIdx.x < nthr) {
      for (int v = threadIdx.x; v < L / 2; v += nthr) {
        unsigned off = (unsigned)((i64)L * u / 2) + v;
        for (int g = 0; g < nrow; g += 32) {
          float2 x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { x[j] = __ldg(X + off); off += rs; }
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += x[j].x * x[j].y;
        }
      }
    }
    __syncthreads();
  }
  if (acc == 123.456f) *out = acc;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(unsigned long long* b, unsigned ph) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned ph) { while (!mbar_try(b, ph)) {} }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// D-stage ring of 8-row groups; consumers (all threads but the producer warp) read every float2 once from smem
template <int D>
__global__ void __launch_bounds__(512, 1) rows_tma(const float* __restrict__ X, int L, i64 Mid, int nrow, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* ring = (float*)smem;  // [D][8][L]
  unsigned long long* full = (unsigned long long*)(ring + (size_t)D * 8 * L);
  unsigned long long* empty = full + D;
  const int tid = threadIdx.x;
  const int ncons = blockDim.x - 32;
  if (tid == 0) {
    for (int s = 0; s < D; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncons / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const i64 nunits = (Mid - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int gpu_ = nrow / 8;  // groups per unit
  const i64 ngroups = nunits * gpu_;
  const i64 rstride = (i64)L * Mid;
  float acc = 0.f;
  if (tid < 32) {
    if (tid == 0) {
      for (i64 g = 0; g < ngroups; ++g) {
        const int s = (int)(g % D);
        if (g >= D) mbar_wait(&empty[s], (unsigned)(((g / D) - 1) & 1));
        const i64 u = blockIdx.x + (g / gpu_) * gridDim.x;
        const int j2 = (int)(g % gpu_);
        mbar_expect(&full[s], 8u * L * 4u);
        for (int r = 0; r < 8; ++r) bulk_load(ring + ((size_t)s * 8 + r) * L, X + (i64)L * u + rstride * (gpu_ * r + j2), (unsigned)L * 4u, &full[s]);
      }
    }
  } else {
    const int c = tid - 32;
    for (i64 g = 0; g < ngroups; ++g) {
      const int s = (int)(g % D);
      mbar_wait(&full[s], (unsigned)((g / D) & 1));
      const float2* st = (const float2*)(ring + (size_t)s * 8 * L);
      for (int v = c; v < L / 2; v += ncons) {
#pragma unroll
        for (int r = 0; r < 8; ++r) { float2 x = st[(size_t)r * (L / 2) + v]; acc += x.x * x.y; }
      }
      __syncwarp();
      if ((c & 31) == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 123.456f) *out = acc;
}

__global__ void __launch_bounds__(512, 1) rows_stg(float2* __restrict__ Y, int L, i64 Mid, int nrow, int nthr) {
  const unsigned rs = (unsigned)((i64)L * Mid / 2);
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    if ((int)threadIdx.x < nthr) {
      for (int v = threadIdx.x; v < L / 2; v += nthr) {
        unsigned off = (unsigned)((i64)L * u / 2) + v;
        for (int j = 0; j < nrow; ++j) { Y[off] = make_float2((float)j, (float)v); off += rs; }
      }
    }
    __syncthreads();
  }
}

// rows written from smem with bulk stores: 2 staging buffers of 8 rows
__global__ void __launch_bounds__(512, 1) rows_tma_store(float* __restrict__ Y, int L, i64 Mid, int nrow) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* stage = (float*)smem;  // [2][8][L]
  const int tid = threadIdx.x;
  const i64 rstride = (i64)L * Mid;
  const int gpu_ = nrow / 8;
  i64 g = 0;
  for (i64 u = blockIdx.x; u < Mid; u += gridDim.x) {
    for (int j2 = 0; j2 < gpu_; ++j2, ++g) {
      float* st = stage + (size_t)(g & 1) * 8 * L;
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the buffer written 2 groups ago has been read
      __syncthreads();
      for (int e = tid; e < 8 * L / 4; e += blockDim.x) ((float4*)st)[e] = make_float4(1.f, 2.f, 3.f, (float)e);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        for (int r = 0; r < 8; ++r) bulk_store(Y + (i64)L * u + rstride * (gpu_ * r + j2), st + (size_t)r * L, (unsigned)L * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
static float timeit(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
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
  const int L = 1280; const i64 Mid = 4096;
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  float *X, *Y, *out;
  for (int nrow = 32; nrow <= 64; nrow += 32) {
    const size_t bytes = (size_t)L * Mid * nrow * 4;
    CK(cudaMalloc(&X, bytes)); CK(cudaMalloc(&Y, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(X, 0, bytes));
    const double gb = bytes / 1e9;
    printf("== rows %d x %d B, %lld units, %.1f MB, %d SMs\n", nrow, L * 4, Mid, bytes / 1e6, nsm);
    float ms;
    ms = timeit([&] { read_lin<<<nsm * 8, 256>>>((const float4*)X, (i64)(bytes / 16), out); }); printf("read_lin          %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { write_lin<<<nsm * 8, 256>>>((float4*)Y, (i64)(bytes / 16)); }); printf("write_lin         %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { cudaMemcpyAsync(Y, X, bytes, cudaMemcpyDeviceToDevice); }); printf("memcpy d2d        %.4f ms %.0f GB/s (r+w)\n", ms, 2 * gb / ms * 1e3);
    ms = timeit([&] { cudaMemsetAsync(Y, 0, bytes); }); printf("memset            %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3);
    for (int nthr = 320; nthr <= 512; nthr += 192) {
      ms = timeit([&] { rows_ldg<<<nsm, 512>>>((const float2*)X, L, Mid, nrow, nthr, out); }); printf("rows_ldg nthr=%d  %.4f ms %.0f GB/s\n", nthr, ms, gb / ms * 1e3);
    }
#define TMA_CASE(D) { size_t sm = (size_t)D * 8 * L * 4 + 2 * D * 8 + 64; CK(cudaFuncSetAttribute(rows_tma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    ms = timeit([&] { rows_tma<D><<<nsm, 512, sm>>>(X, L, Mid, nrow, out); }); printf("rows_tma D=%d (%zu KB ring) %.4f ms %.0f GB/s\n", D, sm / 1024, ms, gb / ms * 1e3); }
    TMA_CASE(2) TMA_CASE(3) TMA_CASE(4) TMA_CASE(5)
    for (int nthr = 320; nthr <= 512; nthr += 192) {
      ms = timeit([&] { rows_stg<<<nsm, 512>>>((float2*)Y, L, Mid, nrow, nthr); }); printf("rows_stg nthr=%d  %.4f ms %.0f GB/s\n", nthr, ms, gb / ms * 1e3);
    }
    { size_t sm = (size_t)2 * 8 * L * 4 + 64; CK(cudaFuncSetAttribute(rows_tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      ms = timeit([&] { rows_tma_store<<<nsm, 512, sm>>>(Y, L, Mid, nrow); }); printf("rows_tma_store    %.4f ms %.0f GB/s\n", ms, gb / ms * 1e3); }
    CK(cudaGetLastError());
    cudaFree(X); cudaFree(Y); cudaFree(out);
  }
  return 0;
}
