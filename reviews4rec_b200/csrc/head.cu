// head.cu -- K3: the small dense heads behind the TextCNN features.
//   r4r_linear_fwd / r4r_linear_bwd : nn.Linear of the heads (100->L, 2L->L, L->1, ...)
//   r4r_fm_fwd / r4r_fm_bwd         : TorchFM second-order interaction (common_pytorch_models.py:49-57)
//   r4r_mse_fwd / r4r_mse_bwd       : loss.py:7-11
// These are a few hundred FLOPs per rating; the kernels keep the weights in shared memory, give
// each rating to one thread (fm, mse) or one warp (linear), and reduce parameter gradients with
// warp shuffles before one atomicAdd per warp (Guideline 12).
#include "common.cuh"

namespace {
constexpr int THREADS = 256;

// ---------------------------------------------------------------------------------- linear
// y[n,o] = b[o] + sum_i x[n,i] W[o,i].  One warp per row n: lanes stride the in_f columns, the
// out_f partial sums are reduced with shuffles.  in_f <= 1024, out_f <= 64.
__global__ void __launch_bounds__(THREADS) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                             const float* __restrict__ b, int64_t n, int in_f, int out_f,
                                                             float* __restrict__ y) {
  extern __shared__ float sW[];                       // [out_f][in_f]
  for (int i = threadIdx.x; i < in_f * out_f; i += THREADS) sW[i] = __ldg(W + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  for (int64_t r = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); r < n; r += warps) {
    const float* xr = x + r * (int64_t)in_f;
    for (int o = 0; o < out_f; ++o) {
      float s = 0.0f;
      for (int i = lane; i < in_f; i += 32) s = fmaf(__ldg(xr + i), sW[o * in_f + i], s);
      s = warp_sum(s);
      if (lane == 0) y[r * (int64_t)out_f + o] = s + (b ? __ldg(b + o) : 0.0f);
    }
  }
}

// dx[n,i] = sum_o gy[n,o] W[o,i]
__global__ void __launch_bounds__(THREADS) linear_bwd_input_kernel(const float* __restrict__ gy, const float* __restrict__ W,
                                                                   int64_t n, int in_f, int out_f, float* __restrict__ dx) {
  extern __shared__ float sW[];
  for (int i = threadIdx.x; i < in_f * out_f; i += THREADS) sW[i] = __ldg(W + i);
  __syncthreads();
  const int64_t total = n * (int64_t)in_f;
  for (int64_t q = (int64_t)blockIdx.x * THREADS + threadIdx.x; q < total; q += (int64_t)gridDim.x * THREADS) {
    int64_t r = q / in_f;
    int i = (int)(q - r * in_f);
    float s = 0.0f;
    for (int o = 0; o < out_f; ++o) s = fmaf(__ldg(gy + r * (int64_t)out_f + o), sW[o * in_f + i], s);
    dx[q] = s;
  }
}

// dW[o,i] += sum_n gy[n,o] x[n,i];  db[o] += sum_n gy[n,o].
// CTA = (row slice); thread (o, i-chunk) accumulates over the slice, then one atomic per element.
__global__ void __launch_bounds__(THREADS) linear_bwd_params_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                    int64_t n, int in_f, int out_f,
                                                                    float* __restrict__ dW, float* __restrict__ db) {
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t n0 = (int64_t)blockIdx.x * per;
  const int64_t n1 = (n0 + per < n) ? n0 + per : n;
  const int total = in_f * out_f;
  for (int q = threadIdx.x; q < total + out_f; q += THREADS) {
    float s = 0.0f;
    if (q < total) {
      int o = q / in_f, i = q - o * in_f;
      for (int64_t r = n0; r < n1; ++r) s = fmaf(__ldg(gy + r * (int64_t)out_f + o), __ldg(x + r * (int64_t)in_f + i), s);
      if (dW && s != 0.0f) atomicAdd(dW + q, s);
    } else {
      int o = q - total;
      for (int64_t r = n0; r < n1; ++r) s += __ldg(gy + r * (int64_t)out_f + o);
      if (db && s != 0.0f) atomicAdd(db + o, s);
    }
  }
}

// ---------------------------------------------------------------------------------- FM
constexpr int FM_MAX_NF = 64, FM_MAX_K = 32;

__global__ void __launch_bounds__(THREADS) fm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ V,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         int64_t n, int nf, int k, float* __restrict__ out) {
  __shared__ float sV[FM_MAX_NF * FM_MAX_K];
  __shared__ float sw[FM_MAX_NF];
  for (int i = threadIdx.x; i < nf * k; i += THREADS) sV[i] = __ldg(V + i);
  for (int i = threadIdx.x; i < nf; i += THREADS) sw[i] = __ldg(w + i);
  __syncthreads();
  const float bias = __ldg(b);
  for (int64_t r = (int64_t)blockIdx.x * THREADS + threadIdx.x; r < n; r += (int64_t)gridDim.x * THREADS) {
    const float* xr = x + r * (int64_t)nf;
    float s1 = 0.0f, s2 = 0.0f, lin = 0.0f;
    for (int c = 0; c < k; ++c) {
      float s = 0.0f, q = 0.0f;
      for (int i = 0; i < nf; ++i) {
        float xi = __ldg(xr + i), v = sV[i * k + c];
        s = fmaf(xi, v, s);
        q = fmaf(xi * xi, v * v, q);
      }
      s1 = fmaf(s, s, s1);
      s2 += q;
    }
    for (int i = 0; i < nf; ++i) lin = fmaf(__ldg(xr + i), sw[i], lin);
    out[r] = 0.5f * (s1 - s2) + (lin + bias);
  }
}

// dx_i = g*(sum_c (s_c V_ic - x_i V_ic^2) + w_i);  dV_ic += g*(s_c x_i - x_i^2 V_ic);  dw_i += g x_i; db += g
__global__ void __launch_bounds__(THREADS) fm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ V,
                                                         const float* __restrict__ w, const float* __restrict__ gout,
                                                         int64_t n, int nf, int k, float* __restrict__ dx,
                                                         float* __restrict__ dV, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sV[FM_MAX_NF * FM_MAX_K];
  __shared__ float sw[FM_MAX_NF];
  __shared__ float sdV[FM_MAX_NF * FM_MAX_K];
  __shared__ float sdw[FM_MAX_NF + 1];
  for (int i = threadIdx.x; i < nf * k; i += THREADS) { sV[i] = __ldg(V + i); sdV[i] = 0.0f; }
  for (int i = threadIdx.x; i < nf; i += THREADS) { sw[i] = __ldg(w + i); sdw[i] = 0.0f; }
  if (threadIdx.x == 0) sdw[nf] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // all lanes of a warp iterate together (shuffle reductions inside), so bound the loop per warp
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t r0 = (int64_t)blockIdx.x * THREADS + (threadIdx.x & ~31); r0 < n; r0 += stride) {
    const int64_t r = r0 + lane;
    const bool valid = r < n;
    const float g = valid ? __ldg(gout + r) : 0.0f;
    const float* xr = x + (valid ? r : 0) * (int64_t)nf;
    float s[FM_MAX_K];
    for (int c = 0; c < k; ++c) {
      float a = 0.0f;
      for (int i = 0; i < nf; ++i) a = fmaf(__ldg(xr + i), sV[i * k + c], a);
      s[c] = a;
    }
    for (int i = 0; i < nf; ++i) {
      const float xi = valid ? __ldg(xr + i) : 0.0f;
      float d = sw[i];
      for (int c = 0; c < k; ++c) {
        float v = sV[i * k + c];
        d += s[c] * v - xi * v * v;
        float gv = warp_sum(g * (s[c] * xi - xi * xi * v));
        if (lane == 0 && dV) atomicAdd(&sdV[i * k + c], gv);
      }
      if (valid && dx) dx[r * (int64_t)nf + i] = g * d;
      float gw = warp_sum(g * xi);
      if (lane == 0) atomicAdd(&sdw[i], gw);
    }
    float gb = warp_sum(g);
    if (lane == 0) atomicAdd(&sdw[nf], gb);
  }
  __syncthreads();
  if (dV) for (int i = threadIdx.x; i < nf * k; i += THREADS) if (sdV[i] != 0.0f) atomicAdd(dV + i, sdV[i]);
  if (dw) for (int i = threadIdx.x; i < nf; i += THREADS) if (sdw[i] != 0.0f) atomicAdd(dw + i, sdw[i]);
  if (db && threadIdx.x == 0 && sdw[nf] != 0.0f) atomicAdd(db, sdw[nf]);
}

// ---------------------------------------------------------------------------------- MSE
__global__ void __launch_bounds__(THREADS) mse_fwd_kernel(const float* __restrict__ out, const float* __restrict__ y, int64_t n,
                                                          float* __restrict__ se, float* __restrict__ sum_se) {
  float local = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    float d = out[i] - y[i];
    float v = d * d;
    if (se) se[i] = v;
    local += v;
  }
  if (sum_se) {
    local = warp_sum(local);
    __shared__ float part[THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int i = 0; i < THREADS / 32; ++i) t += part[i];
      atomicAdd(sum_se, t);
    }
  }
}

__global__ void __launch_bounds__(THREADS) mse_bwd_kernel(const float* __restrict__ out, const float* __restrict__ y,
                                                          const float* __restrict__ gse, int64_t n, float* __restrict__ gout) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS)
    gout[i] = gse[i] * 2.0f * (out[i] - y[i]);
}

inline unsigned grid_for(int64_t work_items, int per_block, int cap = 148 * 8) {
  int64_t b = cdiv64(work_items, per_block);
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}
}  // namespace

extern "C" int r4r_linear_fwd(const float* x, const float* W, const float* b, int64_t n, int in_f, int out_f, float* y, void* stream) {
  R4R_REQUIRE(x && W && y, R4R_EINVAL, "linear_fwd: null pointer");
  R4R_REQUIRE(n >= 0 && in_f > 0 && out_f > 0 && (int64_t)in_f * out_f * 4 <= 48 * 1024, R4R_EUNSUP,
              "linear_fwd: in_f*out_f=%d*%d exceeds the 48 KB shared-memory weight tile", in_f, out_f);
  if (n == 0) return 0;
  linear_fwd_kernel<<<grid_for(n, THREADS / 32), THREADS, (size_t)in_f * out_f * 4, as_stream(stream)>>>(x, W, b, n, in_f, out_f, y);
  R4R_CHECK_LAUNCH("linear_fwd");
  return 0;
}

extern "C" int r4r_linear_bwd(const float* x, const float* W, const float* gy, int64_t n, int in_f, int out_f,
                              float* dx, float* dW, float* db, void* stream) {
  R4R_REQUIRE(gy && (dx == nullptr || W) && ((dW == nullptr && db == nullptr) || x), R4R_EINVAL, "linear_bwd: null pointer");
  R4R_REQUIRE(n >= 0 && in_f > 0 && out_f > 0 && (int64_t)in_f * out_f * 4 <= 48 * 1024, R4R_EUNSUP,
              "linear_bwd: in_f*out_f=%d*%d exceeds the 48 KB shared-memory weight tile", in_f, out_f);
  if (n == 0) return 0;
  cudaStream_t s = as_stream(stream);
  if (dx) {
    linear_bwd_input_kernel<<<grid_for(n * (int64_t)in_f, THREADS), THREADS, (size_t)in_f * out_f * 4, s>>>(gy, W, n, in_f, out_f, dx);
    R4R_CHECK_LAUNCH("linear_bwd_input");
  }
  if (dW || db) {
    unsigned g = grid_for(n, 64, 148 * 2);          // >= 64 rows per CTA slice
    linear_bwd_params_kernel<<<g, THREADS, 0, s>>>(x, gy, n, in_f, out_f, dW, db);
    R4R_CHECK_LAUNCH("linear_bwd_params");
  }
  return 0;
}

extern "C" int r4r_fm_fwd(const float* x, const float* V, const float* w, const float* b, int64_t n, int nf, int k, float* out, void* stream) {
  R4R_REQUIRE(x && V && w && b && out, R4R_EINVAL, "fm_fwd: null pointer");
  R4R_REQUIRE(n >= 0 && nf > 0 && nf <= FM_MAX_NF && k > 0 && k <= FM_MAX_K, R4R_EUNSUP, "fm_fwd: nf=%d (<=%d) k=%d (<=%d)", nf, FM_MAX_NF, k, FM_MAX_K);
  if (n == 0) return 0;
  fm_fwd_kernel<<<grid_for(n, THREADS), THREADS, 0, as_stream(stream)>>>(x, V, w, b, n, nf, k, out);
  R4R_CHECK_LAUNCH("fm_fwd");
  return 0;
}

extern "C" int r4r_fm_bwd(const float* x, const float* V, const float* w, const float* gout, int64_t n, int nf, int k,
                          float* dx, float* dV, float* dw, float* db, void* stream) {
  R4R_REQUIRE(x && V && w && gout, R4R_EINVAL, "fm_bwd: null pointer");
  R4R_REQUIRE(n >= 0 && nf > 0 && nf <= FM_MAX_NF && k > 0 && k <= FM_MAX_K, R4R_EUNSUP, "fm_bwd: nf=%d (<=%d) k=%d (<=%d)", nf, FM_MAX_NF, k, FM_MAX_K);
  if (n == 0) return 0;
  fm_bwd_kernel<<<grid_for(n, THREADS), THREADS, 0, as_stream(stream)>>>(x, V, w, gout, n, nf, k, dx, dV, dw, db);
  R4R_CHECK_LAUNCH("fm_bwd");
  return 0;
}

extern "C" int r4r_mse_fwd(const float* out, const float* y, int64_t n, float* se, float* sum_se, void* stream) {
  R4R_REQUIRE(out && y && (se || sum_se), R4R_EINVAL, "mse_fwd: null pointer");
  if (n <= 0) return 0;
  mse_fwd_kernel<<<grid_for(n, THREADS, 148), THREADS, 0, as_stream(stream)>>>(out, y, n, se, sum_se);
  R4R_CHECK_LAUNCH("mse_fwd");
  return 0;
}

extern "C" int r4r_mse_bwd(const float* out, const float* y, const float* gse, int64_t n, float* gout, void* stream) {
  R4R_REQUIRE(out && y && gse && gout, R4R_EINVAL, "mse_bwd: null pointer");
  if (n <= 0) return 0;
  mse_bwd_kernel<<<grid_for(n, THREADS, 148), THREADS, 0, as_stream(stream)>>>(out, y, gse, n, gout);
  R4R_CHECK_LAUNCH("mse_bwd");
  return 0;
}
