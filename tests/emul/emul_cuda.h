// Minimal CUDA-on-CPU shim used ONLY by tests/emul (never by the product library).
//
// There is no GPU in the build container and GPU time is rationed, so the kernel *logic*
// (index maths, tiling, barriers) of the plain-CUDA-C++ kernels in csrc/ is first exercised
// here: the same sources are compiled with g++ -DPO_EMUL, every CUDA block runs as real
// std::threads with std::barrier standing in for __syncthreads(), and "device memory" is
// host memory.  Kernels that use PTX (TMA, tcgen05, mbarrier, cp.async) are not emulated.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) alignas(x)

struct float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace po_emul {
struct warp_t {
  std::barrier<> bar{32};
  unsigned long long buf[32];
};
struct tls_t {
  dim3 threadIdx, blockIdx, blockDim, gridDim;
  std::barrier<>* bar = nullptr;
  warp_t* warp = nullptr;
  int lane = 0;
};
inline thread_local tls_t tls;
inline unsigned char* g_dyn_smem = nullptr;
inline unsigned char* dyn_smem() { return g_dyn_smem; }

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
  const unsigned nt = block.x * block.y * block.z;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  std::vector<unsigned char> smem_buf(smem + 64);
  g_dyn_smem = (unsigned char*)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
  std::barrier<> block_end((std::ptrdiff_t)nt);
  std::vector<std::unique_ptr<std::barrier<>>> bars(nblocks);
  for (auto& b : bars) b.reset(new std::barrier<>((std::ptrdiff_t)nt));
  std::vector<std::unique_ptr<warp_t>> warps((nt + 31) / 32);
  for (auto& w : warps) w.reset(new warp_t());
  std::vector<std::thread> threads;
  threads.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    threads.emplace_back([&, t]() {
      tls.blockDim = block;
      tls.gridDim = grid;
      tls.threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      tls.warp = warps[t / 32].get();
      tls.lane = (int)(t % 32);
      for (size_t b = 0; b < nblocks; ++b) {
        tls.blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y)));
        tls.bar = bars[b].get();
        fn();
        tls.bar->arrive_and_drop();   // exited threads no longer take part in __syncthreads
        block_end.arrive_and_wait();  // static __shared__ storage is reused by the next block
      }
    });
  }
  for (auto& th : threads) th.join();
  g_dyn_smem = nullptr;
}
// full-warp __shfl_xor_sync (all 32 lanes of the warp must call it; block size % 32 == 0)
template <class T>
inline T shfl_xor(T v, int lane_mask) {
  warp_t& w = *tls.warp;
  std::memcpy(&w.buf[tls.lane], &v, sizeof(T));
  w.bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &w.buf[tls.lane ^ lane_mask], sizeof(T));
  w.bar.arrive_and_wait();
  return r;
}
// __syncwarp(): every lane of the warp must call it
inline void syncwarp() { tls.warp->bar.arrive_and_wait(); }
}  // namespace po_emul

#define threadIdx (po_emul::tls.threadIdx)
#define blockIdx (po_emul::tls.blockIdx)
#define blockDim (po_emul::tls.blockDim)
#define gridDim (po_emul::tls.gridDim)
static inline void __syncthreads() { po_emul::tls.bar->arrive_and_wait(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
// emulated blocks run one at a time but threads of a block run concurrently: keep atomics atomic
#include <atomic>
#include <chrono>
#include <thread>
#include <mutex>
namespace po_emul { inline std::mutex atomic_mu; }
template <class T> static inline T atomicAdd(T* p, T v) { std::lock_guard<std::mutex> lk(po_emul::atomic_mu); T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { std::lock_guard<std::mutex> lk(po_emul::atomic_mu); T o = *p; *p = v; return o; }

// ---- runtime subset -------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
