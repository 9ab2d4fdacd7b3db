// adam.cu -- K7: fused multi-tensor dense Adam, restating torch.optim.Adam as the reference
// configures it (main.py:94-96, utils.py:70-92): betas (0.9, 0.999), eps 1e-8, L2 weight decay
// folded into the gradient (not decoupled), bias correction, and -- because the reference's
// nn.Embedding tables are sparse=False -- a DENSE update of every row of every id table each
// step (SURVEY.md finding 5).  One pass over p, g, m, v: 4 reads + 3 writes of 4 B per element,
// i.e. 28 B/element of pure HBM streaming; the kernel is a vectorised grid-stride sweep.
//
// Up to ADAM_MAX_T tensors are described BY VALUE in the kernel parameter block, so no device
// allocation or H2D copy is needed and the launch is CUDA-graph capturable.
#include "common.cuh"
#include <math.h>

namespace {
constexpr int ADAM_MAX_T = 48;
constexpr int THREADS = 256;

struct AdamPack {
  float* p[ADAM_MAX_T];
  const float* g[ADAM_MAX_T];
  float* m[ADAM_MAX_T];
  float* v[ADAM_MAX_T];
  long long start[ADAM_MAX_T + 1];   // prefix sum of float4-chunk counts (each tensor padded up to 4)
  long long numel[ADAM_MAX_T];
  int nt;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float wd, float one_m_b1, float b2,
                                          float one_m_b2, float step_size, float bc2_sqrt, float eps) {
  g = fmaf(wd, p, g);                              // grad.add(param, alpha=weight_decay)
  m = fmaf(g - m, one_m_b1, m);                    // exp_avg.lerp_(grad, 1-beta1)
  v = fmaf(g * one_m_b2, g, v * b2);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  float denom = sqrtf(v) / bc2_sqrt + eps;         // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - step_size * (m / denom);                 // param.addcdiv_(exp_avg, denom, value=-lr/bc1)
}

__global__ void __launch_bounds__(THREADS) adam_multi_kernel(const __grid_constant__ AdamPack pk, int step,
                                                             const int32_t* __restrict__ step_dev, float lr, float b1, float b2,
                                                             float eps, float wd) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    int t = step_dev ? *step_dev : step;
    double bc1 = 1.0 - pow((double)b1, (double)t);
    double bc2 = 1.0 - pow((double)b2, (double)t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float one_m_b1 = 1.0f - b1, one_m_b2 = 1.0f - b2;
  const long long total = pk.start[pk.nt];
  int t = 0;
  for (long long c = (long long)blockIdx.x * THREADS + threadIdx.x; c < total; c += (long long)gridDim.x * THREADS) {
    while (c >= pk.start[t + 1]) ++t;              // chunks are visited in increasing order per thread
    const long long off = (c - pk.start[t]) * 4;
    const long long rem = pk.numel[t] - off;
    float* p = pk.p[t] + off;
    const float* g = pk.g[t] + off;
    float* m = pk.m[t] + off;
    float* v = pk.v[t] + off;
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                      reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (rem >= 4 && al) {
      float4 P = *reinterpret_cast<float4*>(p), G = *reinterpret_cast<const float4*>(g);
      float4 M = *reinterpret_cast<float4*>(m), Vv = *reinterpret_cast<float4*>(v);
      adam_elem(P.x, G.x, M.x, Vv.x, wd, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
      adam_elem(P.y, G.y, M.y, Vv.y, wd, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
      adam_elem(P.z, G.z, M.z, Vv.z, wd, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
      adam_elem(P.w, G.w, M.w, Vv.w, wd, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
      *reinterpret_cast<float4*>(p) = P;
      *reinterpret_cast<float4*>(m) = M;
      *reinterpret_cast<float4*>(v) = Vv;
    } else {
      const int cnt = rem < 4 ? (int)rem : 4;
      for (int i = 0; i < cnt; ++i) {
        float P = p[i], M = m[i], Vv = v[i];
        adam_elem(P, g[i], M, Vv, wd, one_m_b1, b2, one_m_b2, step_size, bc2_sqrt, eps);
        p[i] = P; m[i] = M; v[i] = Vv;
      }
    }
  }
}
}  // namespace

extern "C" int r4r_adam_step(int nt, float* const* p_host, const float* const* g_host, float* const* m_host,
                             float* const* v_host, const int64_t* numel_host, int step, const int32_t* step_dev,
                             float lr, float beta1, float beta2, float eps, float weight_decay, void* stream) {
  R4R_REQUIRE(nt >= 0 && (nt == 0 || (p_host && g_host && m_host && v_host && numel_host)), R4R_EINVAL, "adam_step: null pointer");
  R4R_REQUIRE(step_dev != nullptr || step >= 1, R4R_EINVAL, "adam_step: step must be >= 1");
  cudaStream_t s = as_stream(stream);
  for (int base = 0; base < nt; base += ADAM_MAX_T) {
    AdamPack pk;
    int cnt = nt - base < ADAM_MAX_T ? nt - base : ADAM_MAX_T;
    pk.nt = cnt;
    long long acc = 0;
    for (int i = 0; i < cnt; ++i) {
      R4R_REQUIRE(p_host[base + i] && g_host[base + i] && m_host[base + i] && v_host[base + i] && numel_host[base + i] >= 0,
                  R4R_EINVAL, "adam_step: tensor %d has a null pointer", base + i);
      pk.p[i] = p_host[base + i]; pk.g[i] = g_host[base + i]; pk.m[i] = m_host[base + i]; pk.v[i] = v_host[base + i];
      pk.numel[i] = numel_host[base + i];
      pk.start[i] = acc;
      acc += (numel_host[base + i] + 3) / 4;
    }
    pk.start[cnt] = acc;
    for (int i = cnt + 1; i <= ADAM_MAX_T; ++i) pk.start[i] = acc;
    if (acc == 0) continue;
    long long blocks = (acc + THREADS - 1) / THREADS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_multi_kernel<<<(unsigned)blocks, THREADS, 0, s>>>(pk, step, step_dev, lr, beta1, beta2, eps, weight_decay);
    R4R_CHECK_LAUNCH("adam_multi");
  }
  return 0;
}

// Device-side step counter for CUDA-graph replays of the optimizer: *counter += 1.
namespace {
__global__ void counter_inc_kernel(int32_t* c) { *c += 1; }
}  // namespace

extern "C" int r4r_counter_inc(int32_t* counter, void* stream) {
  R4R_REQUIRE(counter, R4R_EINVAL, "counter_inc: null pointer");
  counter_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(counter);
  R4R_CHECK_LAUNCH("counter_inc");
  return 0;
}
