// FP64 pipe peaks on the B200 for config C5's roofline (MEASURED_PEAKS.json has no FP64 entry):
//   dfma : SIMT FP64 FMA throughput (8 independent chains per thread)
//   dmma : mma.sync.aligned.m8n8k4.row.col.f64 (the only FP64 tensor-core path; tcgen05 has no f64 kind)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_fp64 mb_fp64.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

int main() {
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; CK(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = nsm * 8;
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma_kernel<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("dfma  %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 8 * iters * 256.0 * blocks / ms / 1e9);
    cudaEventRecord(e0); dmma_kernel<<<blocks, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("dmma  %.3f ms  %.2f TFLOP/s  (m8n8k4: 512 flop per warp instruction)\n", ms, 512.0 * 4 * iters * 8.0 * blocks / ms / 1e9);
  }
  return 0;
}
