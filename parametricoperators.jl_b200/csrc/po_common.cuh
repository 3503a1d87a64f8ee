// Common device/host helpers for libparops (sm_100a).
//
// Everything in the kernels is plain CUDA C++ unless a file says otherwise, so that
// tests/emul can compile the very same sources with g++ (PO_EMUL) and run the kernel
// *logic* on CPU threads before a GPU box is spent on it.  The emulator is test tooling
// only: libparops.so never contains or loads it.
#pragma once

#include <stddef.h>
#include <stdint.h>

#ifdef PO_EMUL
#include "emul_cuda.h"
#else
#include <cuda_runtime.h>
#endif

#include "../../include/parops.h"

// Translation units: libparops.so is built from several TUs (po_api.cu + tu_*.cu, one per heavy kernel family) so that
// nvcc runs in parallel; -DPO_MULTI_TU makes the family launchers external and each tu_*.cu defines the PO_TU_* macro
// of the family it owns.  Without PO_MULTI_TU (the CPU emulator build, or a plain `nvcc po_api.cu`) everything is one TU.
#ifdef PO_MULTI_TU
#define PO_XTU
#else
#define PO_XTU static
#endif

#define PO_HD __host__ __device__ __forceinline__
#define PO_D __device__ __forceinline__

#ifdef PO_EMUL
#define PO_LAUNCH(kernel, grid, block, smem, stream, ...) \
  po_emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define PO_DYN_SMEM(name) unsigned char* name = po_emul::dyn_smem()
#else
#define PO_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PO_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// The kernels of one apply form a strict chain on one stream; each is 10-140 us long, so the ~1.3 us launch gap and
// the ~1-2 us prologue (twiddle tables into shared memory, mbarrier init) of every kernel are a visible share of the
// step.  A kernel launched with PO_LAUNCH_PDL may become resident while its predecessor's last CTAs are still running
// (the predecessor calls po_pdl_trigger() first thing) and runs its prologue there; it MUST call po_pdl_wait() before
// its first access to anything a previous kernel wrote or still reads (= before touching any tensor / workspace
// memory; kernel parameters and plan-lifetime tables are fine).  The wait returns once the predecessor grid has
// completed and its writes are visible, so the chain stays transitively ordered as long as every PDL-launched kernel
// waits.  PO_PDL=0 falls back to plain launches; both instructions are no-ops in a kernel launched without the attribute.
#ifdef PO_EMUL
#define PO_LAUNCH_PDL PO_LAUNCH
PO_D void po_pdl_wait() {}
PO_D void po_pdl_trigger() {}
#else
#include <stdlib.h>
PO_D void po_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PO_D void po_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static inline bool po_pdl_enabled() {
  static const bool on = getenv("PO_PDL") ? atoi(getenv("PO_PDL")) != 0 : true;
  return on;
}
template <class... KArgs, class... Args>
static inline void po_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = po_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through cudaGetLastError() like <<<>>>
}
#define PO_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) po_launch_pdl(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)
#endif

typedef long long i64;

// ---- element-type helpers -------------------------------------------------------------
template <class T> struct po_traits;
template <> struct po_traits<float>   { typedef float real;  typedef float2 cplx;  static const bool is_cplx = false; };
template <> struct po_traits<double>  { typedef double real; typedef double2 cplx; static const bool is_cplx = false; };
template <> struct po_traits<float2>  { typedef float real;  typedef float2 cplx;  static const bool is_cplx = true; };
template <> struct po_traits<double2> { typedef double real; typedef double2 cplx; static const bool is_cplx = true; };

template <class T> PO_HD T po_zero();
template <> PO_HD float po_zero<float>() { return 0.f; }
template <> PO_HD double po_zero<double>() { return 0.0; }
template <> PO_HD float2 po_zero<float2>() { return make_float2(0.f, 0.f); }
template <> PO_HD double2 po_zero<double2>() { return make_double2(0.0, 0.0); }

PO_HD float po_conj(float a) { return a; }
PO_HD double po_conj(double a) { return a; }
PO_HD float2 po_conj(float2 a) { return make_float2(a.x, -a.y); }
PO_HD double2 po_conj(double2 a) { return make_double2(a.x, -a.y); }

PO_HD float po_fma(float a, float b, float c) { return fmaf(a, b, c); }
PO_HD double po_fma(double a, double b, double c) { return fma(a, b, c); }

// c += a*b for the type combinations the operators need
PO_HD void po_mac(float& c, float a, float b) { c = fmaf(a, b, c); }
PO_HD void po_mac(double& c, double a, double b) { c = fma(a, b, c); }
PO_HD void po_mac(float2& c, float2 a, float b) { c.x = fmaf(a.x, b, c.x); c.y = fmaf(a.y, b, c.y); }
PO_HD void po_mac(double2& c, double2 a, double b) { c.x = fma(a.x, b, c.x); c.y = fma(a.y, b, c.y); }
PO_HD void po_mac(float2& c, float2 a, float2 b) {
  c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
  c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
}
PO_HD void po_mac(double2& c, double2 a, double2 b) {
  c.x = fma(a.x, b.x, c.x); c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y); c.y = fma(a.y, b.x, c.y);
}
// real part of a complex product (truncated c2r inverse DFT)
PO_HD void po_mac(float& c, float2 a, float2 b) { c = fmaf(a.x, b.x, c); c = fmaf(-a.y, b.y, c); }
PO_HD void po_mac(double& c, double2 a, double2 b) { c = fma(a.x, b.x, c); c = fma(-a.y, b.y, c); }

PO_HD float po_mul(float a, float b) { return a * b; }
PO_HD double po_mul(double a, double b) { return a * b; }
PO_HD float2 po_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PO_HD double2 po_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PO_HD float2 po_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PO_HD double2 po_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
PO_HD float2 po_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PO_HD double2 po_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
PO_HD float po_add(float a, float b) { return a + b; }
PO_HD double po_add(double a, double b) { return a + b; }
PO_HD float2 po_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
PO_HD double2 po_scale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
PO_HD float po_scale(float a, float s) { return a * s; }
PO_HD double po_scale(double a, double s) { return a * s; }

PO_HD float2 po_make_cplx(float x, float y) { return make_float2(x, y); }
PO_HD double2 po_make_cplx(double x, double y) { return make_double2(x, y); }

static inline size_t po_dtype_size(int dt) {
  switch (dt) {
    case PO_F32: return 4;
    case PO_F64: return 8;
    case PO_C64: return 8;
    case PO_C128: return 16;
  }
  return 0;
}
static inline bool po_dtype_is_complex(int dt) { return dt == PO_C64 || dt == PO_C128; }
static inline bool po_dtype_is_double(int dt) { return dt == PO_F64 || dt == PO_C128; }
static inline int po_dtype_complex_of(int dt) { return po_dtype_is_double(dt) ? PO_C128 : PO_C64; }
static inline int po_dtype_real_of(int dt) { return po_dtype_is_double(dt) ? PO_F64 : PO_F32; }
static inline const char* po_dtype_name(int dt) {
  switch (dt) {
    case PO_F32: return "Float32";
    case PO_F64: return "Float64";
    case PO_C64: return "ComplexF32";
    case PO_C128: return "ComplexF64";
  }
  return "?";
}
