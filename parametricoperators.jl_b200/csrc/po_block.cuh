// Block epilogue of the FNO layer (examples/fno.jl:34-44), one pass:
//     y = gelu.( s .+ (I_grid (x) W) x .+ b )
// with s = S x the spectral-convolution output, W the c_out x c_in pointwise channel-mixing ParMatrix
// (src/ParMatrix.jl:42-46 inside ParKron with ParIdentity, src/ParKron.jl:190-220), b an optional bias
// (src/ParMatrix.jl:53-58).  The reference runs this as: 2 permutedims + GEMM + 2 permutedims for W x, a
// broadcast add and a broadcast gelu (>= 7 passes over the 671 MB tensor at C3); here x and s are read once and
// y is written once (3 x 671 MB, the compulsory traffic).
// Layout: channels are the fastest dim, so a grid point's channels are contiguous: x[c + c_in * p].
#pragma once
#include "po_kernels.cuh"

// gelu (tanh form, examples/fno.jl:43 / NNlib.gelu) as x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3): one MUFU.EX2 and
// one fast division instead of the ~30-instruction tanhf; |error| < 1e-6 relative (checked against the oracle in the tests).
PO_D float po_gelu_fast(float x) {
#ifdef PO_EMUL
  return po_gelu(x);
#else
  const float u2 = 1.5957691216057308f * (x + 0.044715f * x * x * x);  // 2u
  return __fdividef(x, 1.f + __expf(-u2));
#endif
}

// fast path: float, c_in % 4 == 0, c_in <= 32.  Thread (c', slot) keeps row c' of W in registers and produces
// output channel c' of every nslots-th point of the tile; the x tile sits in shared memory and is read with
// 128-bit broadcast loads (one LDS.128 per 4 FMAs).
template <int CMAX>
__global__ void __launch_bounds__(256)
po_block_epilogue_f32_kernel(const float* __restrict__ W, const float* __restrict__ X, const float* __restrict__ S, const float* __restrict__ bias,
                             float* __restrict__ Y, int c_in, int c_out, i64 npts, int tile_pts, int act) {
  PO_DYN_SMEM(smem_raw);
  float* xs = (float*)smem_raw;  // [tile_pts][c_in]
  const int tid = threadIdx.x;
  const int nslots = 256 / c_out;
  const int cp = tid % c_out, slot = tid / c_out;
  const bool worker = slot < nslots;
  float w[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) w[c] = (worker && c < c_in) ? W[cp + (i64)c_out * c] : 0.f;
  const float b = (bias && worker) ? bias[cp] : 0.f;
  const i64 ntiles = (npts + tile_pts - 1) / tile_pts;
  for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const i64 p0 = tile * tile_pts;
    const int np = (int)((npts - p0 < tile_pts) ? (npts - p0) : tile_pts);
    const float4* xg = (const float4*)(X + p0 * c_in);
    const int nv = np * c_in / 4;
    for (int e = tid; e < nv; e += 256) ((float4*)xs)[e] = __ldg(xg + e);
    __syncthreads();
    if (worker) {
      for (int p = slot; p < np; p += nslots) {
        const float4* xp = (const float4*)(xs + p * c_in);
        float acc = b;
#pragma unroll
        for (int c4 = 0; c4 < CMAX / 4; ++c4) {
          if (c4 * 4 < c_in) {
            const float4 v = xp[c4];
            acc = fmaf(w[4 * c4], v.x, acc); acc = fmaf(w[4 * c4 + 1], v.y, acc);
            acc = fmaf(w[4 * c4 + 2], v.z, acc); acc = fmaf(w[4 * c4 + 3], v.w, acc);
          }
        }
        const i64 o = cp + (i64)c_out * (p0 + p);
        if (S) acc += S[o];
        Y[o] = act ? po_gelu(acc) : acc;  // (po_gelu_fast measured slower in THIS kernel: 1.13 vs 0.83 ms)
      }
    }
    __syncthreads();
  }
}

// generic path (any real dtype / channel counts): one thread per output element
template <class T>
__global__ void __launch_bounds__(256)
po_block_epilogue_generic_kernel(const T* __restrict__ W, const T* __restrict__ X, const T* __restrict__ S, const T* __restrict__ bias,
                                 T* __restrict__ Y, int c_in, int c_out, i64 total, int act) {
  for (i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (i64)gridDim.x * blockDim.x) {
    const int cp = (int)(o % c_out);
    const i64 p = o / c_out;
    T acc = bias ? bias[cp] : (T)0;
    for (int c = 0; c < c_in; ++c) acc = po_fma(W[cp + (i64)c_out * c], X[c + (i64)c_in * p], acc);
    if (S) acc += S[o];
    Y[o] = act ? po_gelu(acc) : acc;
  }
}

static cudaError_t po_launch_block_epilogue(int dt, const void* W, const void* X, const void* S, const void* bias, void* Y, i64 c_in, i64 c_out,
                                            i64 npts, int act, int num_sms, cudaStream_t st) {
  if (npts * c_out == 0) return cudaSuccess;
  if (dt == PO_F32 && c_in % 4 == 0 && c_in <= 32 && c_out <= 256 && c_in >= 4 && ((size_t)X % 16) == 0) {
    const int nslots = 256 / (int)c_out;
    const int tile_pts = nslots * 24;
    const size_t smem = (size_t)tile_pts * c_in * sizeof(float);
    const i64 ntiles = (npts + tile_pts - 1) / tile_pts;
    i64 grid = (i64)(num_sms > 0 ? num_sms : 4) * 8;
    if (grid > ntiles) grid = ntiles;
    PO_LAUNCH((po_block_epilogue_f32_kernel<32>), dim3((unsigned)grid), dim3(256), smem, st, (const float*)W, (const float*)X, (const float*)S,
              (const float*)bias, (float*)Y, (int)c_in, (int)c_out, npts, tile_pts, act);
    return cudaGetLastError();
  }
  const i64 total = npts * c_out;
  const unsigned grid = po_grid_1d(total, 256, 148 * 32);
  if (dt == PO_F32) { PO_LAUNCH((po_block_epilogue_generic_kernel<float>), dim3(grid), dim3(256), 0, st, (const float*)W, (const float*)X, (const float*)S, (const float*)bias, (float*)Y, (int)c_in, (int)c_out, total, act); }
  else if (dt == PO_F64) { PO_LAUNCH((po_block_epilogue_generic_kernel<double>), dim3(grid), dim3(256), 0, st, (const double*)W, (const double*)X, (const double*)S, (const double*)bias, (double*)Y, (int)c_in, (int)c_out, total, act); }
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
