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
// CTA = a slice of LBP_ROWS rows staged in shared memory (coalesced loads); thread (o, i) then sums its
// product over the slice from shared memory and commits one atomic per element.
constexpr int LBP_ROWS = 32;
__global__ void __launch_bounds__(THREADS) linear_bwd_params_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                    int64_t n, int in_f, int out_f,
                                                                    float* __restrict__ dW, float* __restrict__ db) {
  extern __shared__ float sm[];                       // [LBP_ROWS][in_f] x, then [LBP_ROWS][out_f] gy
  float* sx = sm;
  float* sg = sm + LBP_ROWS * in_f;
  const int total = in_f * out_f;
  for (int64_t n0 = (int64_t)blockIdx.x * LBP_ROWS; n0 < n; n0 += (int64_t)gridDim.x * LBP_ROWS) {
    const int rows = (int)((n - n0 < LBP_ROWS) ? n - n0 : LBP_ROWS);
    for (int q = threadIdx.x; q < rows * in_f; q += THREADS) sx[q] = __ldg(x + n0 * in_f + q);
    for (int q = threadIdx.x; q < rows * out_f; q += THREADS) sg[q] = __ldg(gy + n0 * out_f + q);
    __syncthreads();
    for (int q = threadIdx.x; q < total + out_f; q += THREADS) {
      float s = 0.0f;
      if (q < total) {
        const int o = q / in_f, i = q - o * in_f;
        for (int r = 0; r < rows; ++r) s = fmaf(sg[r * out_f + o], sx[r * in_f + i], s);
        if (dW && s != 0.0f) atomicAdd(dW + q, s);
      } else {
        const int o = q - total;
        for (int r = 0; r < rows; ++r) s += sg[r * out_f + o];
        if (db && s != 0.0f) atomicAdd(db + o, s);
      }
    }
    __syncthreads();
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
// CTA = a slice of FMB_ROWS rows.  Phase 1: one thread per row computes s = xV, writes dx and stages
// x, g*s, g in shared memory.  Phase 2: thread (i, c) sums its parameter-gradient term over the slice
// from shared memory and commits ONE atomic per element (was: a warp-shuffle reduction per element per
// 32 rows -- 44 us for 4096 rows).
constexpr int FMB_ROWS = 64;
__global__ void __launch_bounds__(THREADS) fm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ V,
                                                         const float* __restrict__ w, const float* __restrict__ gout,
                                                         int64_t n, int nf, int k, float* __restrict__ dx,
                                                         float* __restrict__ dV, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float fsm[];
  float* sV = fsm;                                    // [nf][k]
  float* sw = sV + nf * k;                            // [nf]
  float* sx = sw + nf;                                // [FMB_ROWS][nf]
  float* sa = sx + FMB_ROWS * nf;                     // [FMB_ROWS][k]   g * s
  float* sgr = sa + FMB_ROWS * k;                     // [FMB_ROWS]      g
  for (int i = threadIdx.x; i < nf * k; i += THREADS) sV[i] = __ldg(V + i);
  for (int i = threadIdx.x; i < nf; i += THREADS) sw[i] = __ldg(w + i);
  __syncthreads();
  for (int64_t n0 = (int64_t)blockIdx.x * FMB_ROWS; n0 < n; n0 += (int64_t)gridDim.x * FMB_ROWS) {
    const int rows = (int)((n - n0 < FMB_ROWS) ? n - n0 : FMB_ROWS);
    for (int q = threadIdx.x; q < rows * nf; q += THREADS) sx[q] = __ldg(x + n0 * nf + q);
    for (int q = threadIdx.x; q < rows; q += THREADS) sgr[q] = __ldg(gout + n0 + q);
    __syncthreads();
    if ((int)threadIdx.x < rows) {
      const int r = threadIdx.x;
      const float g = sgr[r];
      const float* xr = sx + r * nf;
      float sc[FM_MAX_K];
      for (int c = 0; c < k; ++c) {
        float a = 0.0f;
        for (int i = 0; i < nf; ++i) a = fmaf(xr[i], sV[i * k + c], a);
        sc[c] = a;
        sa[r * k + c] = g * a;
      }
      if (dx) {
        for (int i = 0; i < nf; ++i) {
          const float xi = xr[i];
          float d = sw[i];
          for (int c = 0; c < k; ++c) {
            const float v = sV[i * k + c];
            d += sc[c] * v - xi * v * v;
          }
          dx[(n0 + r) * nf + i] = g * d;
        }
      }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nf * k + nf + 1; q += THREADS) {
      float s = 0.0f;
      if (q < nf * k) {                               // dV[i,c] = sum_r x_ri (g s_rc) - V_ic sum_r g x_ri^2
        const int i = q / k, c = q - i * k;
        float s2 = 0.0f;
        for (int r = 0; r < rows; ++r) {
          const float xi = sx[r * nf + i];
          s = fmaf(xi, sa[r * k + c], s);
          s2 = fmaf(sgr[r] * xi, xi, s2);
        }
        s -= sV[q] * s2;
        if (dV && s != 0.0f) atomicAdd(dV + q, s);
      } else if (q < nf * k + nf) {                   // dw[i] = sum_r g x_ri
        const int i = q - nf * k;
        for (int r = 0; r < rows; ++r) s = fmaf(sgr[r], sx[r * nf + i], s);
        if (dw && s != 0.0f) atomicAdd(dw + i, s);
      } else {                                        // db = sum_r g
        for (int r = 0; r < rows; ++r) s += sgr[r];
        if (db && s != 0.0f) atomicAdd(db, s);
      }
    }
    __syncthreads();
  }
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

// index of the (first) largest score per row: the top-1 pick of eval.py:74-78 (torch.topk(output[b], k=1))
__global__ void __launch_bounds__(THREADS) rows_argmax_kernel(const float* __restrict__ x, int64_t n, int c, int64_t* __restrict__ idx) {
  for (int64_t r = (int64_t)blockIdx.x * THREADS + threadIdx.x; r < n; r += (int64_t)gridDim.x * THREADS) {
    const float* row = x + r * (int64_t)c;
    float best = row[0];
    int at = 0;
    for (int j = 1; j < c; ++j) {
      const float v = row[j];
      if (v > best) { best = v; at = j; }
    }
    idx[r] = at;
  }
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
    unsigned g = grid_for(n, LBP_ROWS, 148 * 4);
    const size_t smem = (size_t)LBP_ROWS * (in_f + out_f) * sizeof(float);
    R4R_REQUIRE(smem <= 48 * 1024, R4R_EUNSUP, "linear_bwd: in_f + out_f = %d exceeds the shared-memory slice", in_f + out_f);
    linear_bwd_params_kernel<<<g, THREADS, smem, s>>>(x, gy, n, in_f, out_f, dW, db);
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
  const size_t smem = (size_t)(nf * k + nf + FMB_ROWS * (nf + k + 1)) * sizeof(float);     // <= 33.3 KB at nf=64, k=32
  fm_bwd_kernel<<<grid_for(n, FMB_ROWS, 148 * 2), THREADS, smem, as_stream(stream)>>>(x, V, w, gout, n, nf, k, dx, dV, dw, db);
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

extern "C" int r4r_rows_argmax(const float* x, int64_t n, int c, int64_t* idx, void* stream) {
  R4R_REQUIRE(x && idx, R4R_EINVAL, "rows_argmax: null pointer");
  R4R_REQUIRE(n >= 0 && c > 0, R4R_EINVAL, "rows_argmax: bad sizes");
  if (n == 0) return 0;
  rows_argmax_kernel<<<grid_for(n, THREADS, 148 * 4), THREADS, 0, as_stream(stream)>>>(x, n, c, idx);
  R4R_CHECK_LAUNCH("rows_argmax");
  return 0;
}
