// Reverse mode for a compiled plan: the Zygote/ChainRules surface examples/fno.jl relies on
// (SURVEY.md 3.5).  A `ccall` is opaque to Zygote, so every plan carries its own pullback:
//   xbar = (dy/dx)^H ybar      and      thetabar for every ParMatrix / ParDiagonal / ParTensor / bias.
// Rules (matching what the reference gets from ChainRules / AbstractFFTs / NNlib / its own rrules):
//   fft      -> bfft (unnormalised backward)            ifft  -> fft / n
//   rfft     -> Re(W^H ybar) (no Hermitian doubling)    irfft -> (c_k / n) rfft(ybar)
//   ParRestriction fwd/adj -> the other direction (src/ParRestriction.jl:57-71, overwrite semantics)
//   theta*x  -> xbar = theta' ybar, thetabar = ybar x'   (conj-transposed variants for theta')
//   d .* x   -> xbar = conj(d) .* ybar, dbar = sum_cols conj(x) .* ybar
//   ParTensor (batched_mul) -> per-mode xbar / Wbar
//   ParRepartition -> its adjoint (src/ParRepartition.jl:199-205)
// The forward pass run with po_plan_apply_taped leaves the INPUT of every parametric step in
// the tape, so nothing is recomputed except the tiny A*x inside the fused mix+diag step.
#pragma once
#include "po_comm.cuh"

// ---- reduction kernels ---------------------------------------------------------------------
PO_D void po_atomic_add(float* p, float v) { atomicAdd(p, v); }
PO_D void po_atomic_add(double* p, double v) { atomicAdd(p, v); }
PO_D void po_atomic_add(float2* p, float2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
PO_D void po_atomic_add(double2* p, double2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }

// G[a + ra*b] += sum_col U(a, col) * conj(V(b, col)), U: (I, ra, O), V: (I, rb, O); col = (i, o).
// Each CTA owns a chunk of columns and a 64x64 tile of G entries (<= 16 accumulators / thread).
template <class T>
__global__ void __launch_bounds__(256)
po_gram_kernel(const T* __restrict__ U, const T* __restrict__ V, T* __restrict__ G, i64 I, int ra, int rb, i64 ncols, i64 cols_per_cta) {
  constexpr int TC = 16;
  __shared__ T Us[64][TC + 1];
  __shared__ T Vs[64][TC + 1];
  const int tid = threadIdx.x;
  const int a0 = (blockIdx.y % ((ra + 63) / 64)) * 64, b0 = (blockIdx.y / ((ra + 63) / 64)) * 64;
  const int na = (ra - a0 < 64) ? ra - a0 : 64, nb = (rb - b0 < 64) ? rb - b0 : 64;
  const i64 c_begin = (i64)blockIdx.x * cols_per_cta;
  i64 c_end = c_begin + cols_per_cta;
  if (c_end > ncols) c_end = ncols;
  T acc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = po_zero<T>();
  for (i64 c0 = c_begin; c0 < c_end; c0 += TC) {
    for (int e = tid; e < 64 * TC; e += 256) {
      int r, c;
      if (I > 1) { c = e % TC; r = e / TC; } else { r = e % 64; c = e / 64; }
      const i64 col = c0 + c;
      T u = po_zero<T>(), v = po_zero<T>();
      if (col < c_end) {
        const i64 o = (I == 1) ? col : col / I;
        if (r < na) u = U[col + I * ((i64)(a0 + r) + o * (i64)(ra - 1))];
        if (r < nb) v = V[col + I * ((i64)(b0 + r) + o * (i64)(rb - 1))];
      }
      Us[r][c] = u;
      Vs[r][c] = po_conj(v);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + 256 * q;  // entry (a, b) of the 64x64 tile
      const int a = e % 64, b = e / 64;
      if (a < na && b < nb) {
#pragma unroll
        for (int c = 0; c < TC; ++c) po_mac(acc[q], Us[a][c], Vs[b][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int e = tid + 256 * q;
    const int a = e % 64, b = e / 64;
    if (a < na && b < nb) po_atomic_add(&G[(a0 + a) + (i64)ra * (b0 + b)], acc[q]);
  }
}

static cudaError_t po_launch_gram(int dt, const void* U, const void* V, void* G, i64 I, i64 ra, i64 rb, i64 O, int num_sms, cudaStream_t st) {
  const i64 ncols = I * O;
  cudaError_t e = cudaMemsetAsync(G, 0, (size_t)(ra * rb) * po_dtype_size(dt), st);
  if (e != cudaSuccess || ncols == 0) return e;
  const int tiles = (int)(((ra + 63) / 64) * ((rb + 63) / 64));
  i64 ctas = (i64)(num_sms > 0 ? num_sms : 64) * 4 / tiles;
  if (ctas < 1) ctas = 1;
  i64 cpc = (ncols + ctas - 1) / ctas;
  cpc = ((cpc + 15) / 16) * 16;
  ctas = (ncols + cpc - 1) / cpc;
  dim3 grid((unsigned)ctas, (unsigned)tiles);
  switch (dt) {
    case PO_F32: PO_LAUNCH((po_gram_kernel<float>), grid, dim3(256), 0, st, (const float*)U, (const float*)V, (float*)G, I, (int)ra, (int)rb, ncols, cpc); break;
    case PO_F64: PO_LAUNCH((po_gram_kernel<double>), grid, dim3(256), 0, st, (const double*)U, (const double*)V, (double*)G, I, (int)ra, (int)rb, ncols, cpc); break;
    case PO_C64: PO_LAUNCH((po_gram_kernel<float2>), grid, dim3(256), 0, st, (const float2*)U, (const float2*)V, (float2*)G, I, (int)ra, (int)rb, ncols, cpc); break;
    case PO_C128: PO_LAUNCH((po_gram_kernel<double2>), grid, dim3(256), 0, st, (const double2*)U, (const double2*)V, (double2*)G, I, (int)ra, (int)rb, ncols, cpc); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// g[j] += sum_{i, o} conj(A[i, j, o]) * B[i, j, o]   (A == nullptr: sum of B -- the bias cotangent)
template <class T>
__global__ void __launch_bounds__(256)
po_modesum_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ g, i64 I, i64 n, i64 O) {
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < n * O; t += (i64)gridDim.x * blockDim.x) {
    const i64 j = t % n, o = t / n;
    const i64 base = I * (j + n * o);
    T acc = po_zero<T>();
    for (i64 i = 0; i < I; ++i) {
      if (A) po_mac(acc, po_conj(A[base + i]), B[base + i]);
      else acc = po_add(acc, B[base + i]);
    }
    po_atomic_add(&g[j], acc);
  }
}

static cudaError_t po_launch_modesum(int dt, const void* A, const void* B, void* g, i64 I, i64 n, i64 O, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(g, 0, (size_t)n * po_dtype_size(dt), st);
  if (e != cudaSuccess || I * n * O == 0) return e;
  const unsigned grid = po_grid_1d(n * O, 256);
  switch (dt) {
    case PO_F32: PO_LAUNCH((po_modesum_kernel<float>), dim3(grid), dim3(256), 0, st, (const float*)A, (const float*)B, (float*)g, I, n, O); break;
    case PO_F64: PO_LAUNCH((po_modesum_kernel<double>), dim3(grid), dim3(256), 0, st, (const double*)A, (const double*)B, (double*)g, I, n, O); break;
    case PO_C64: PO_LAUNCH((po_modesum_kernel<float2>), dim3(grid), dim3(256), 0, st, (const float2*)A, (const float2*)B, (float2*)g, I, n, O); break;
    case PO_C128: PO_LAUNCH((po_modesum_kernel<double2>), dim3(grid), dim3(256), 0, st, (const double2*)A, (const double2*)B, (double2*)g, I, n, O); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ParTensor cotangents: xbar[i,m,b] = sum_o conj(W(i,o,m)) ybar[o,m,b];  Wbar(i,o,m) = sum_b conj(x[i,m,b]) ybar[o,m,b]
template <class T>
__global__ void __launch_bounds__(256)
po_tensor_xbar_kernel(const T* __restrict__ W, const T* __restrict__ Yb, T* __restrict__ Xb, int ic, int oc, i64 M, i64 total, int w_oi) {
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const int i = (int)(idx % ic);
    const i64 t = idx / ic;
    const i64 mm = t % M, b = t / M;
    const T* yp = Yb + (i64)oc * (mm + M * b);
    const T* wp = W + (i64)ic * oc * mm;
    T acc = po_zero<T>();
    for (int o = 0; o < oc; ++o) {
      const T w = w_oi ? wp[o + (i64)oc * i] : wp[i + (i64)ic * o];
      po_mac(acc, po_conj(w), yp[o]);
    }
    Xb[idx] = acc;
  }
}
template <class T>
__global__ void __launch_bounds__(256)
po_tensor_wbar_kernel(const T* __restrict__ X, const T* __restrict__ Yb, T* __restrict__ Wb, int ic, int oc, i64 M, i64 B, i64 total, int w_oi) {
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    int i, o;
    const i64 mm = idx / ((i64)ic * oc);
    const int r = (int)(idx % ((i64)ic * oc));
    if (w_oi) { o = r % oc; i = r / oc; } else { i = r % ic; o = r / ic; }
    T acc = po_zero<T>();
    for (i64 b = 0; b < B; ++b) po_mac(acc, po_conj(X[i + (i64)ic * (mm + M * b)]), Yb[o + (i64)oc * (mm + M * b)]);
    Wb[idx] = acc;
  }
}

PO_HD float po_gelu_grad(float x) {
  const float c = 0.7978845608028654f, a = 0.044715f;
  const float u = c * (x + a * x * x * x), t = tanhf(u);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * c * (1.f + 3.f * a * x * x);
}
PO_HD double po_gelu_grad(double x) {
  const double c = 0.7978845608028654, a = 0.044715;
  const double u = c * (x + a * x * x * x), t = tanh(u);
  return 0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * c * (1.0 + 3.0 * a * x * x);
}
template <class T>
__global__ void __launch_bounds__(256)
po_gelu_grad_kernel(const T* __restrict__ X, const T* __restrict__ Yb, T* __restrict__ Xb, i64 total) {
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x)
    Xb[idx] = po_gelu_grad(X[idx]) * Yb[idx];
}

template <class T>
static cudaError_t po_tensor_pullback_t(const po_step& s, const void* W, const void* yb, void* xb, const void* xin, void* wgrad, cudaStream_t st) {
  const i64 tot_x = (i64)s.ic * s.M * s.O, tot_w = (i64)s.ic * s.oc * s.M;
  if (tot_x) PO_LAUNCH((po_tensor_xbar_kernel<T>), dim3(po_grid_1d(tot_x, 256, 148 * 32)), dim3(256), 0, st, (const T*)W, (const T*)yb, (T*)xb, s.ic, s.oc, s.M, tot_x, s.w_oi);
  if (wgrad && tot_w) {
    if (!xin) return cudaErrorInvalidValue;
    PO_LAUNCH((po_tensor_wbar_kernel<T>), dim3(po_grid_1d(tot_w, 256, 148 * 32)), dim3(256), 0, st, (const T*)xin, (const T*)yb, (T*)wgrad, s.ic, s.oc, s.M, s.O, tot_w, s.w_oi);
  }
  return cudaGetLastError();
}
static cudaError_t po_tensor_pullback(const po_step& s, const void* W, const void* yb, void* xb, const void* xin, void* wgrad, cudaStream_t st, int num_sms = 0,
                                      int* err = nullptr) {
  // weight-streaming kernels first (po_tensor.cuh): x-bar reads W once through the bulk-copy ring, W-bar is one coalesced write stream
  if (num_sms > 0 && err && (!wgrad || xin)) {
    cudaError_t e = po_launch_tensor_stream(s.in_dt, true, W, yb, xb, s.ic, s.oc, s.M, s.O, s.w_oi, num_sms, err, st);
    if (e == cudaSuccess && wgrad) {
      e = po_launch_tensor_wbar(s.in_dt, xin, yb, wgrad, s.ic, s.oc, s.M, s.O, s.w_oi, num_sms, st);
      if (e == cudaErrorNotSupported) {
        po_step s2 = s;
        (void)s2;
        e = cudaSuccess;
        const i64 tot_w = (i64)s.ic * s.oc * s.M;
        switch (s.in_dt) {
          case PO_F32: PO_LAUNCH((po_tensor_wbar_kernel<float>), dim3(po_grid_1d(tot_w, 256, 148 * 32)), dim3(256), 0, st, (const float*)xin, (const float*)yb, (float*)wgrad, s.ic, s.oc, s.M, s.O, tot_w, s.w_oi); break;
          case PO_F64: PO_LAUNCH((po_tensor_wbar_kernel<double>), dim3(po_grid_1d(tot_w, 256, 148 * 32)), dim3(256), 0, st, (const double*)xin, (const double*)yb, (double*)wgrad, s.ic, s.oc, s.M, s.O, tot_w, s.w_oi); break;
          case PO_C64: PO_LAUNCH((po_tensor_wbar_kernel<float2>), dim3(po_grid_1d(tot_w, 256, 148 * 32)), dim3(256), 0, st, (const float2*)xin, (const float2*)yb, (float2*)wgrad, s.ic, s.oc, s.M, s.O, tot_w, s.w_oi); break;
          case PO_C128: PO_LAUNCH((po_tensor_wbar_kernel<double2>), dim3(po_grid_1d(tot_w, 256, 148 * 32)), dim3(256), 0, st, (const double2*)xin, (const double2*)yb, (double2*)wgrad, s.ic, s.oc, s.M, s.O, tot_w, s.w_oi); break;
        }
        e = cudaGetLastError();
      }
    }
    if (e != cudaErrorNotSupported) return e;
  }
  switch (s.in_dt) {
    case PO_F32: return po_tensor_pullback_t<float>(s, W, yb, xb, xin, wgrad, st);
    case PO_F64: return po_tensor_pullback_t<double>(s, W, yb, xb, xin, wgrad, st);
    case PO_C64: return po_tensor_pullback_t<float2>(s, W, yb, xb, xin, wgrad, st);
    case PO_C128: return po_tensor_pullback_t<double2>(s, W, yb, xb, xin, wgrad, st);
  }
  return cudaErrorInvalidValue;
}

#define PO_DISPATCH4(dt, KERNEL, grid, st, ...)                                                       \
  switch (dt) {                                                                                       \
    case PO_F32: PO_LAUNCH((KERNEL<float>), dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;        \
    case PO_F64: PO_LAUNCH((KERNEL<double>), dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;       \
    case PO_C64: PO_LAUNCH((KERNEL<float2>), dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;       \
    case PO_C128: PO_LAUNCH((KERNEL<double2>), dim3(grid), dim3(256), 0, st, __VA_ARGS__); break;     \
  }

// ---- transposed step data (built on the first pullback) --------------------------------------
struct po_pb_step {
  int* d_in_map = nullptr;   // FFT: bin -> ybar row
  int* d_out_map = nullptr;  // FFT: xbar row -> bin (-1: zero)
  int* d_map = nullptr;      // gather: transposed map
  po_fused_tw* ftw = nullptr;
  po_fused2_tw* f2tw = nullptr;
  po_c16_tw* ctw = nullptr;
  po_c16_tw* ctw2 = nullptr;
  float2* gram_part = nullptr;  // ST_ZMZ: per-CTA partial Gram matrices
  long long gram_nparts = 0;
  po_repart_info* repart = nullptr;
};

struct po_pb_state {
  std::vector<po_pb_step> steps;
  void* tmp = nullptr;  // 2 x max(m * cols) elements for the fused mix+diag step
  size_t tmp_half = 0;
  ~po_pb_state() {
    for (auto& s : steps) { delete s.ftw; delete s.f2tw; delete s.ctw; delete s.ctw2; delete s.repart; }
  }
};

static std::map<po_plan*, po_pb_state*> g_po_pb;
static std::mutex g_po_pb_mu;

static void po_pullback_release(po_plan* pl) {
  std::lock_guard<std::mutex> lk(g_po_pb_mu);
  auto it = g_po_pb.find(pl);
  if (it != g_po_pb.end()) { delete it->second; g_po_pb.erase(it); }
}

static int po_pullback_prepare(po_plan* pl, po_pb_state** out) {
  std::lock_guard<std::mutex> lk(g_po_pb_mu);
  auto it = g_po_pb.find(pl);
  if (it != g_po_pb.end()) { *out = it->second; return PO_OK; }
  po_pb_state* pb = new po_pb_state();
  pb->steps.resize(pl->steps.size());
  int rc;
  for (size_t i = 0; i < pl->steps.size(); ++i) {
    po_step& s = pl->steps[i];
    po_pb_step& t = pb->steps[i];
    if (s.impl == IM_FFT) {
      const i64 nspec = s.dft_real ? s.dft_n / 2 + 1 : s.dft_n;
      if (!s.dft_inverse) {  // forward with out_map (row k -> bin): transposed load map bin -> k
        if (!s.out_map.empty()) {
          std::vector<int> m((size_t)nspec, -1);
          for (size_t k = 0; k < s.out_map.size(); ++k) m[(size_t)s.out_map[k]] = (int)k;
          if ((rc = po_upload(pl, m, (void**)&t.d_in_map))) return rc;
        }
      } else if (!s.in_map.empty()) {  // inverse with in_map (bin -> src row): transposed store map src -> bin
        std::vector<int> m((size_t)s.n, -1);
        for (size_t b = 0; b < s.in_map.size(); ++b) if (s.in_map[b] >= 0) m[(size_t)s.in_map[b]] = (int)b;
        if ((rc = po_upload(pl, m, (void**)&t.d_out_map))) return rc;
      }
    } else if (s.impl == IM_GATHER) {
      if ((rc = po_upload(pl, s.map_T, (void**)&t.d_map))) return rc;
    } else if (s.impl == IM_C16) {
      // forward (scale 1) <-> padded backward without 1/n ; inverse (1/n) <-> forward scaled by 1/n
      t.ctw = new po_c16_tw();
      const int R = (int)s.dft_n / 16;
      const bool fwd = !s.dft_inverse;
      for (int j2 = 0; j2 < 8; ++j2)
        for (int r = 0; r < 16; ++r) {
          double c = 0, sn = 0;
          const i64 bin = r < 8 ? r : s.dft_n - 16 + r;
          if (j2 < R) po_cis(bin * j2, s.dft_n, fwd ? +1.0 : -1.0, c, sn);
          const double w = fwd ? 1.0 : 1.0 / (double)s.dft_n;
          t.ctw->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(w * sn));
        }
    } else if (s.impl == IM_CFULL) {
      // fft -> bfft (unnormalised backward);  ifft -> fft / n
      t.ctw = new po_c16_tw();
      const int R2 = (int)s.dft_n / 16;
      for (int j2 = 0; j2 < 8; ++j2)
        for (int k1 = 0; k1 < 16; ++k1) {
          double c = 0, sn = 0;
          if (j2 < R2) po_cis((i64)j2 * k1, s.dft_n, s.dft_inverse ? -1.0 : +1.0, c, sn);
          t.ctw->tw[j2 * 16 + k1] = make_float2((float)c, (float)sn);
        }
    } else if (s.impl == IM_R16) {
      // rfft -> Re(W^H ybar) (unit weights, no Hermitian doubling);  irfft -> (c_k / n) rfft(ybar)
      t.ctw = new po_c16_tw();
      const int R = (int)s.dft_n / 16;
      const bool fwd = !s.dft_inverse;
      for (int j2 = 0; j2 < 8; ++j2)
        for (int r = 0; r < 16; ++r) {
          double c = 0, sn = 0;
          if (j2 < R) po_cis((i64)r * j2, s.dft_n, fwd ? +1.0 : -1.0, c, sn);
          const double w = fwd ? 1.0 : ((r == 0) ? 1.0 : 2.0) / (double)s.dft_n;
          t.ctw->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(w * sn));
        }
    } else if (s.impl == IM_ZMZ) {
      // reverse order: adjoint of the zero-padded inverse (forward-type, 1/n) first, adjoint of the pruned forward (inverse-type, unit weight) last
      t.ctw = new po_c16_tw(); t.ctw2 = new po_c16_tw();
      const int R = (int)s.dft_n / 16;
      for (int j2 = 0; j2 < 8; ++j2)
        for (int r = 0; r < 16; ++r) {
          double c = 0, sn = 0;
          const i64 bin = r < 8 ? r : s.dft_n - 16 + r;
          if (j2 < R) po_cis(bin * j2, s.dft_n, -1.0, c, sn);
          const double w = 1.0 / (double)s.dft_n;
          t.ctw->tw[j2 * 16 + r] = make_float2((float)(w * c), (float)(w * sn));
          t.ctw2->tw[j2 * 16 + r] = make_float2((float)c, (float)(-sn));
        }
      t.gram_nparts = po_zmz_ctas((int)s.dft_n, s.oc, s.ic, s.z_Mi, s.O);
      if (t.gram_nparts > 0) {
        PO_CUDA_CHECK(cudaMalloc((void**)&t.gram_part, (size_t)t.gram_nparts * (size_t)(s.ic * s.oc) * sizeof(float2)));
        pl->owned.push_back(t.gram_part);
      }
    } else if (s.impl == IM_FUSED_TX) {
      {  // generation 2: adjoint of the forward pair = backward pair with unit weights; adjoint of the
         // backward pair = forward pair scaled by c_k / (nt nx) per t-mode
        double ks[8];
        po_herm_kscale(s.f_nt, s.f_nx, ks);
        t.f2tw = new po_fused2_tw();
        po_fill_fused2_tw(*t.f2tw, s.f_nx, !s.dft_inverse, s.dft_inverse ? ks : nullptr);
      }
      t.ftw = new po_fused_tw();
      if (!s.dft_inverse) {
        // adjoint of the forward pair = backward pair with unit weights (no 1/n, no Hermitian doubling)
        po_fill_fused_tw(*t.ftw, s.f_nt, s.f_nx, true);
        const double undo = (double)s.f_nt * (double)s.f_nx;
        for (int j2 = 0; j2 < 8; ++j2)
          for (int k = 0; k < 8; ++k) {
            const double f = undo / ((k == 0) ? 1.0 : 2.0);
            for (int q = 0; q < 4; ++q) {
              float2& e = t.ftw->tw_t[(j2 * 8 + k) * 4 + q];
              e = make_float2((float)(e.x * f), (float)(e.y * f));
            }
          }
      } else {
        // adjoint of the backward pair = forward pair scaled by c_k / (nt * nx) per t-mode
        po_fill_fused_tw(*t.ftw, s.f_nt, s.f_nx, false);
        for (int k = 0; k < 8; ++k) t.ftw->kscale[k] = (float)(((k == 0) ? 1.0 : 2.0) / ((double)s.f_nt * (double)s.f_nx));
      }
    } else if (s.impl == IM_REPART) {
      po_repart_info* info = new po_repart_info();
      std::vector<int32_t> din(s.repart->dims_out.begin(), s.repart->dims_out.end()), dout(s.repart->dims_in.begin(), s.repart->dims_in.end());
      std::vector<int64_t> gsz(s.repart->global.begin(), s.repart->global.end());
      rc = po_repart_build(s.repart->N, gsz.data(), din.data(), dout.data(), s.repart->rank, *info);
      if (rc) { delete info; return rc; }
      const size_t esz = po_dtype_size(s.in_dt);
      i64 so = 0, ro = 0;
      for (auto& b : info->send) { info->send_off.push_back(so); if (b.peer != info->rank) so += b.count() * s.O; }
      for (auto& b : info->recv) { info->recv_off.push_back(ro); if (b.peer != info->rank) ro += b.count() * s.O; }
      if (so) { PO_CUDA_CHECK(cudaMalloc(&info->d_send, (size_t)so * esz)); pl->owned.push_back(info->d_send); }
      if (ro) { PO_CUDA_CHECK(cudaMalloc(&info->d_recv, (size_t)ro * esz)); pl->owned.push_back(info->d_recv); }
      po_repart_try_peer(pl, *info, esz, s.O);
      t.repart = info;
    } else if (s.impl == IM_MIXDIAG) {
      const size_t need = (size_t)((s.m > s.n ? s.m : s.n) * s.O) * po_dtype_size(s.in_dt);
      if (need > pb->tmp_half) pb->tmp_half = (need + 255) & ~(size_t)255;
    }
  }
  if (pb->tmp_half) {
    PO_CUDA_CHECK(cudaMalloc(&pb->tmp, 2 * pb->tmp_half));
    pl->owned.push_back(pb->tmp);
  }
  g_po_pb[pl] = pb;
  *out = pb;
  return PO_OK;
}

// cotangent of one step: src = ybar_i (output-sized), dst = xbar_i (input-sized), xin = taped input or null
static int po_pullback_step(po_plan* pl, po_pb_state* pb, size_t i, const void* src, void* dst, const void* xin,
                            const void* const* params, void* const* grads, int n_params, cudaStream_t st) {
  po_step& s = pl->steps[i];
  po_pb_step& t = pb->steps[i];
  if (s.kind != ST_REPART && (s.I * s.n * s.O == 0 || po_step_out_elems(s) == 0)) return PO_OK;  // empty shard: nothing flows back either (parameter cotangents keep their zeros)
  const void* prm = nullptr;
  void* grad = nullptr;
  if (s.param_slot >= 0) {
    if (s.param_slot >= n_params || !params || !params[s.param_slot]) return po_fail(PO_ERR_INVALID, "parameter slot %d missing", s.param_slot);
    prm = params[s.param_slot];
    grad = grads ? grads[s.param_slot] : nullptr;
  }
  const int nsm = pl->ctx->num_sms;
  cudaError_t e = cudaSuccess;
  switch (s.impl) {
    case IM_FFT: {
      const bool dbl = po_dtype_is_double(s.in_dt);
      if (!s.dft_real) {
        if (!s.dft_inverse) e = po_launch_fft(0, dbl, src, dst, s.d_table, s.I, s.dft_n, s.O, 1, 1.0, 1.0, t.d_in_map, (int)s.m, nullptr, 0, st);
        else e = po_launch_fft(0, dbl, src, dst, s.d_table, s.I, s.dft_n, s.O, 0, 1.0 / (double)s.dft_n, 1.0, nullptr, 0, t.d_out_map, (int)s.n, st);
      } else {
        if (!s.dft_inverse) e = po_launch_fft(2, dbl, src, dst, s.d_table, s.I, s.dft_n, s.O, 1, 1.0, 0.5, t.d_in_map, (int)s.m, nullptr, 0, st);
        else e = po_launch_fft(1, dbl, src, dst, s.d_table, s.I, s.dft_n, s.O, 0, 1.0 / (double)s.dft_n, 2.0, nullptr, 0, t.d_out_map, (int)s.n, st);
      }
      pl->launches += 1;
      break;
    }
    case IM_DENSE_TABLE:  // A^H: swap strides, conjugate; (out_dt, in_dt) swap picks Re(A^H y) / A^H y
      e = po_launch_dense(s.table_dt, s.out_dt, s.in_dt, s.d_table, s.a_cs, s.a_rs, !s.conj_a, src, dst, s.I, s.m, s.n, s.O, st);
      pl->launches += 1;
      break;
    case IM_C16: e = po_launch_c16(!s.dft_inverse, (int)s.dft_n, src, dst, *t.ctw, s.I, s.O, st); pl->launches += 1; break;
    case IM_CFULL: e = po_launch_cfull(!s.dft_inverse, (int)s.dft_n, src, dst, *t.ctw, s.I, s.O, s.dft_inverse ? 1.f / (float)s.dft_n : 1.f, st); pl->launches += 1; break;
    case IM_R16: e = po_launch_r16(!s.dft_inverse, (int)s.dft_n, src, dst, *t.ctw, s.I, s.O, false, st); pl->launches += 1; break;
    case IM_FUSED_TX:
      e = cudaErrorNotSupported;
      if (t.f2tw && s.dft_inverse && po_use_gen3(pl))
        e = po_launch_fwd_ring(s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && t.f2tw && s.dft_inverse && po_use_gen3(pl) && po_use_pair())
        e = po_launch_fwd_pair(s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, pl->ctx->d_err, st);
      if (t.f2tw && !s.dft_inverse && po_use_gen3(pl) && po_use_inv_stage())
        e = po_launch_inv_stage(s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && t.f2tw && !s.dft_inverse && po_use_gen3(pl) && po_use_inv_stage() && po_use_pair())
        e = po_launch_inv_pair(s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, pl->ctx->d_err, st);
      if (e == cudaErrorNotSupported && t.f2tw && po_use_gen2(pl) && po_use_tile2(!s.dft_inverse, s.f_I0, s.f_nx))
        e = po_launch_tile2(!s.dft_inverse, s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, st);
      if (e == cudaErrorNotSupported && t.f2tw && po_use_gen2(pl))
        e = po_launch_fused2(!s.dft_inverse, s.f_nt, s.f_nx, src, dst, *t.f2tw, s.f_I0, s.f_mid, s.f_B, nsm, st);
      if (e != cudaErrorNotSupported) { pl->launches += 1; break; }
      e = po_launch_fused_tx(!s.dft_inverse, s.f_nt, s.f_nx, src, dst, *t.ftw, s.f_I0, s.f_mid, s.f_B,
                             (pl->flags & PO_PLAN_NO_PIPELINE) ? 0 : nsm, st);
      pl->launches += 1;
      break;
    case IM_GATHER: e = po_launch_gather(s.in_dt, src, dst, t.d_map, s.I, s.m, s.n, s.O, st); pl->launches += 1; break;
    case IM_DENSE_PARAM: {
      e = po_launch_dense(s.table_dt, s.out_dt, s.in_dt, prm, s.a_cs, s.a_rs, !s.conj_a, src, dst, s.I, s.m, s.n, s.O, st);
      pl->launches += 1;
      if (e == cudaSuccess && grad) {
        if (!xin) return po_fail(PO_ERR_INVALID, "pullback of a ParMatrix step needs the tape");
        // theta*x: thetabar = ybar x'   (prm_rows = m);   theta'*x: thetabar = x ybar'   (prm_rows = n)
        if (!s.adjoint) e = po_launch_gram(s.in_dt, src, xin, grad, s.I, s.m, s.n, s.O, nsm, st);
        else e = po_launch_gram(s.in_dt, xin, src, grad, s.I, s.n, s.m, s.O, nsm, st);
        pl->launches += 1;
      }
      break;
    }
    case IM_DIAG: {
      e = po_launch_diag(s.in_dt, prm, !s.adjoint, src, dst, s.I, s.n, s.O, st);
      pl->launches += 1;
      if (e == cudaSuccess && grad) {
        if (!xin) return po_fail(PO_ERR_INVALID, "pullback of a ParDiagonal step needs the tape");
        // d.*x: dbar = sum conj(x) ybar ;  conj(d).*x: dbar = sum x conj(ybar) = conj(sum conj(x) ybar)
        e = !s.adjoint ? po_launch_modesum(s.in_dt, xin, src, grad, s.I, s.n, s.O, st) : po_launch_modesum(s.in_dt, src, xin, grad, s.I, s.n, s.O, st);
        pl->launches += 1;
      }
      break;
    }
    case IM_MIXDIAG: {
      if (s.param_slot2 < 0 || s.param_slot2 >= n_params || !params[s.param_slot2]) return po_fail(PO_ERR_INVALID, "parameter slot %d missing", s.param_slot2);
      const void* dprm = params[s.param_slot2];
      void* dgrad = grads ? grads[s.param_slot2] : nullptr;
      // y = dd .* (AA x), AA = theta or theta', dd = d or conj(d).   ubar = conj(dd) .* ybar ; xbar = AA' ubar
      e = po_launch_mix_diag(s.in_dt, prm, s.a_cs, s.a_rs, !s.conj_a, dprm, !s.adjoint2, s.nd, src, dst, s.m, s.n, s.O, st);
      pl->launches += 1;
      if (e == cudaSuccess && (grad || dgrad)) {
        if (!xin) return po_fail(PO_ERR_INVALID, "pullback of the mix+diag step needs the tape");
        void* u = pb->tmp;                                  // u = AA x          (m x cols)
        void* ub = (char*)pb->tmp + pb->tmp_half;           // ubar = conj(dd) .* ybar
        e = po_launch_mix_diag(s.in_dt, prm, s.a_rs, s.a_cs, s.conj_a, nullptr, 0, 1, xin, u, s.n, s.m, s.O, st);
        if (e == cudaSuccess) e = po_launch_diag(s.in_dt, dprm, !s.adjoint2, src, ub, s.m, s.nd, s.O / s.nd, st);
        pl->launches += 2;
        if (e == cudaSuccess && dgrad) {
          e = !s.adjoint2 ? po_launch_modesum(s.in_dt, u, src, dgrad, s.m, s.nd, s.O / s.nd, st) : po_launch_modesum(s.in_dt, src, u, dgrad, s.m, s.nd, s.O / s.nd, st);
          pl->launches += 1;
        }
        if (e == cudaSuccess && grad) {
          e = !s.adjoint ? po_launch_gram(s.in_dt, ub, xin, grad, 1, s.m, s.n, s.O, nsm, st) : po_launch_gram(s.in_dt, xin, ub, grad, 1, s.n, s.m, s.O, nsm, st);
          pl->launches += 1;
        }
      }
      break;
    }
    case IM_ZMZ: {
      if (s.param_slot2 < 0 || s.param_slot2 >= n_params || !params[s.param_slot2]) return po_fail(PO_ERR_INVALID, "parameter slot %d missing", s.param_slot2);
      void* dgrad = grads ? grads[s.param_slot2] : nullptr;
      if ((grad || dgrad) && !xin) return po_fail(PO_ERR_INVALID, "pullback of the z-mix-z step needs the tape");
      if (dgrad) e = cudaMemsetAsync(dgrad, 0, (size_t)s.nd * sizeof(float2), st);
      if (e != cudaSuccess) break;
      po_zmz_params P;
      memset(&P, 0, sizeof(P));
      P.X = (const float2*)src; P.Y = (float2*)dst; P.A = (const float2*)prm;
      P.b_rs = s.a_cs; P.b_cs = s.a_rs; P.conj_b = !s.conj_a;   // B = AA^H
      P.d = (const float2*)params[s.param_slot2]; P.conj_d = !s.adjoint2; P.diag_first = 1; P.nd = s.nd;
      P.nin = s.oc; P.nout = s.ic; P.Mi = s.z_Mi; P.O = s.O;
      const bool want = xin && (grad || dgrad);
      if (want) {
        P.xhat = (const float2*)xin; P.gram_part = grad ? t.gram_part : nullptr; P.gradd = (float2*)dgrad;
        P.fw_adjoint = s.adjoint; P.fw_adjoint2 = s.adjoint2; P.a_rs = s.a_rs; P.a_cs = s.a_cs; P.fw_conj_a = s.conj_a;
      }
      e = po_launch_zmz(want, (int)s.dft_n, P, *t.ctw, *t.ctw2, st);
      pl->launches += 1;
      if (e == cudaSuccess && want && grad) {
        const int ne = s.ic * s.oc;
        PO_LAUNCH_PDL(po_zmz_gram_reduce_kernel, dim3((unsigned)((ne + 7) / 8)), dim3(256), 0, st, (const float2*)t.gram_part, t.gram_nparts, s.oc, s.ic,
                  s.adjoint, (int)s.prm_rows, (float2*)grad);
        e = cudaGetLastError();
        pl->launches += 1;
      }
      break;
    }
    case IM_TENSOR: {
      e = po_tensor_pullback(s, prm, src, dst, xin, grad, st, pl->ctx->num_sms, pl->ctx->d_err);
      pl->launches += grad ? 2 : 1;
      break;
    }
    case IM_BIAS: {
      const size_t bytes = (size_t)(s.I * s.n * s.O) * po_dtype_size(s.in_dt);
      if (src != dst && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess && grad) { e = po_launch_modesum(s.in_dt, nullptr, src, grad, s.I, s.n, s.O, st); pl->launches += 1; }
      break;
    }
    case IM_GELU: {
      if (!xin) return po_fail(PO_ERR_INVALID, "pullback of gelu needs the tape");
      const i64 tot = s.I * s.n * s.O;
      if (tot) {
        if (s.in_dt == PO_F32) { PO_LAUNCH((po_gelu_grad_kernel<float>), dim3(po_grid_1d(tot, 256)), dim3(256), 0, st, (const float*)xin, (const float*)src, (float*)dst, tot); }
        else { PO_LAUNCH((po_gelu_grad_kernel<double>), dim3(po_grid_1d(tot, 256)), dim3(256), 0, st, (const double*)xin, (const double*)src, (double*)dst, tot); }
        e = cudaGetLastError();
        pl->launches += 1;
      }
      break;
    }
    case IM_REPART: {
      po_repart_info* keep = s.repart;
      std::swap(s.n, s.m);
      s.repart = t.repart;
      int rc = po_repart_run(pl, s, src, dst, st);
      s.repart = keep;
      std::swap(s.n, s.m);
      return rc;
    }
    default: return po_fail(PO_ERR_UNSUPPORTED, "no pullback for step '%s'", s.note.c_str());
  }
  if (e != cudaSuccess) return po_fail(PO_ERR_CUDA, "pullback kernel failed (%s): %s", s.note.c_str(), cudaGetErrorString(e));
  return PO_OK;
}

static int po_plan_run_pullback(po_plan* pl, const void* tape, const void* ybar, void* xbar, const void* const* params,
                                void* const* param_grads, int32_t n_params, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!pl) return po_fail(PO_ERR_INVALID, "null plan");
  if (int rc0 = po_ctx_check(pl->ctx)) return rc0;
  const size_t need = po_plan_ws_bytes(pl);
  if (need > ws_bytes || (need && !ws)) return po_fail(PO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  PO_CUDA_CHECK(cudaSetDevice(pl->ctx->device));
  pl->launches = 0;
  if (pl->batch == 0) return PO_OK;  // no columns: xbar is empty and the parameter cotangents (sums over columns) keep the caller's zeros
  const size_t nsteps = pl->steps.size();
  if (nsteps == 0) {
    const size_t bytes = (size_t)(pl->domain * pl->batch) * po_dtype_size(pl->ddt);
    if (xbar != ybar && bytes) PO_CUDA_CHECK(cudaMemcpyAsync(xbar, ybar, bytes, cudaMemcpyDeviceToDevice, st));
    return PO_OK;
  }
  if ((!ybar && pl->range * pl->batch > 0) || (!xbar && pl->domain * pl->batch > 0)) return po_fail(PO_ERR_INVALID, "null ybar or xbar");
  po_pb_state* pb = nullptr;
  int rc = po_pullback_prepare(pl, &pb);
  if (rc) return rc;
  std::vector<long long> toff;
  po_tape_layout(pl, &toff);
  const void* src = ybar;
#ifndef PO_EMUL
  std::vector<cudaEvent_t> evs;
  if (pl->profiling) {
    evs.resize(nsteps + 1);
    for (auto& e : evs) PO_CUDA_CHECK(cudaEventCreate(&e));
    PO_CUDA_CHECK(cudaEventRecord(evs[0], st));
  }
#endif
  for (size_t i = nsteps; i-- > 0;) {
    void* dst = (i == 0) ? xbar : (void*)((char*)ws + (i % 2) * pl->ws_half);
    if (i >= 1 && pl->steps[i].impl == IM_C16 && !pl->steps[i].dft_inverse && pl->steps[i - 1].impl == IM_FUSED_TX && !pl->steps[i - 1].dft_inverse) {
      // cotangents of F_y and of the fused forward (t, x): the same inverse-type pair as in po_plan_run, in z-chunks
      void* out = (i - 1 == 0) ? xbar : (void*)((char*)ws + ((i - 1) % 2) * pl->ws_half);
      if (po_inv_pair_ok(pl, pl->steps[i], pl->steps[i - 1], pb->steps[i].ctw, pb->steps[i - 1].f2tw, dst, out)) {
#ifndef PO_EMUL
        if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[nsteps - i], st));  // the pair's time is booked on the fused step
#endif
        if ((rc = po_run_inv_pair(pl, pl->steps[i], pl->steps[i - 1], *pb->steps[i].ctw, *pb->steps[i - 1].f2tw, src, dst, out, st))) return rc;
#ifndef PO_EMUL
        if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[nsteps - i + 1], st));
#endif
        src = out;
        --i;
        continue;
      }
    }
    const void* xin = (toff[i] >= 0 && tape) ? (const void*)((const char*)tape + toff[i]) : nullptr;
    if ((rc = po_pullback_step(pl, pb, i, src, dst, xin, params, param_grads, n_params, st))) return rc;
#ifndef PO_EMUL
    if (pl->profiling) PO_CUDA_CHECK(cudaEventRecord(evs[nsteps - i], st));
#endif
    src = dst;
  }
#ifndef PO_EMUL
  if (pl->profiling) pl->pending_pb_events.push_back(evs);
#endif
  return PO_OK;
}
