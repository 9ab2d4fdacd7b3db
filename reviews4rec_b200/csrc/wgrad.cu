// wgrad.cu -- K4: conv weight/bias gradient through ReLU + global max-pool.
//
// Only the arg-max window of each (doc, filter) receives gradient (SURVEY.md finding 4), so the
// reference's dense convolution_backward GEMM (36% of its CPU step) collapses to a re-gather of
// <= 3 rows per (doc, filter):
//     dW[f,0,j,:] += sum_n gy[n,f] * Xpad[n, argmax[n,f] + j, :]      gy = gpooled * [pooled > 0]
//     db[f]       += sum_n gy[n,f]
// Grid = F x S: CTA (f, s) owns filter f and the s-th slice of documents; each thread keeps its
// float4 slices of the 3xE window in registers across the slice, then commits them with one
// atomicAdd per element.  The gathered rows come from the fp32 word table, which is L2-resident at
// the reference's vocabulary (50,001 x 300 x 4 B = 60 MB < 126 MB L2), so this kernel is bound by
// L2 gather bandwidth, not HBM: N*F*3 row reads of 4E bytes.
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {
constexpr int THREADS = 256;
constexpr int MAXV = 3;       // float4 accumulators per thread: 3*E/4 <= MAXV*THREADS  -> E <= 1024

template <int NV>
__global__ void __launch_bounds__(THREADS) conv_wgrad_kernel(
    const float* __restrict__ table, int64_t V, int E, const int64_t* __restrict__ idx, int64_t N, int T,
    const int32_t* __restrict__ argmax, const float* __restrict__ pooled, const float* __restrict__ gpooled,
    int F, float* __restrict__ dW, float* __restrict__ db) {
  const int f = blockIdx.x;
  const int64_t per = (N + gridDim.y - 1) / gridDim.y;
  const int64_t n0 = (int64_t)blockIdx.y * per;
  const int64_t n1 = (n0 + per < N) ? n0 + per : N;
  const int e4 = E >> 2;
  const int nvec = 3 * e4;
  const int tid = threadIdx.x;

  float4 acc[NV];
  int vj[NV], vc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int q = tid + v * THREADS;
    vj[v] = q < nvec ? q / e4 : -1;
    vc[v] = q < nvec ? q % e4 : 0;
  }
  float bsum = 0.0f;

  for (int64_t n = n0; n < n1; ++n) {
    const float p = __ldg(pooled + n * F + f);
    const float g = __ldg(gpooled + n * F + f);
    if (!(p > 0.0f) || g == 0.0f) continue;            // CTA-uniform: dead ReLU or zero upstream grad
    const int a = __ldg(argmax + n * F + f);
    bsum += g;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (vj[v] < 0) continue;
      int pos = a + vj[v] - 2;                          // document row feeding window row j
      if (pos < 0 || pos >= T) continue;                // zero padding row
      int64_t tok = __ldg(idx + n * (int64_t)T + pos);
      float4 x = __ldg(reinterpret_cast<const float4*>(table + tok * (int64_t)E) + vc[v]);
      acc[v].x = fmaf(g, x.x, acc[v].x);
      acc[v].y = fmaf(g, x.y, acc[v].y);
      acc[v].z = fmaf(g, x.z, acc[v].z);
      acc[v].w = fmaf(g, x.w, acc[v].w);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (vj[v] < 0) continue;
    float* dst = dW + ((int64_t)f * 3 + vj[v]) * E + vc[v] * 4;
    atomicAdd(dst + 0, acc[v].x);
    atomicAdd(dst + 1, acc[v].y);
    atomicAdd(dst + 2, acc[v].z);
    atomicAdd(dst + 3, acc[v].w);
  }
  if (tid == 0 && bsum != 0.0f) atomicAdd(db + f, bsum);
}

// scalar fallback for E % 4 != 0 (tiny test shapes): one thread per (j, e) element, strided
__global__ void __launch_bounds__(THREADS) conv_wgrad_scalar_kernel(
    const float* __restrict__ table, int64_t V, int E, const int64_t* __restrict__ idx, int64_t N, int T,
    const int32_t* __restrict__ argmax, const float* __restrict__ pooled, const float* __restrict__ gpooled,
    int F, float* __restrict__ dW, float* __restrict__ db) {
  const int f = blockIdx.x;
  const int64_t per = (N + gridDim.y - 1) / gridDim.y;
  const int64_t n0 = (int64_t)blockIdx.y * per;
  const int64_t n1 = (n0 + per < N) ? n0 + per : N;
  for (int q = threadIdx.x; q < 3 * E; q += THREADS) {
    int j = q / E, e = q % E;
    float acc = 0.0f;
    for (int64_t n = n0; n < n1; ++n) {
      const float p = __ldg(pooled + n * F + f);
      const float g = __ldg(gpooled + n * F + f);
      if (!(p > 0.0f) || g == 0.0f) continue;
      int pos = __ldg(argmax + n * F + f) + j - 2;
      if (pos < 0 || pos >= T) continue;
      int64_t tok = __ldg(idx + n * (int64_t)T + pos);
      acc = fmaf(g, __ldg(table + tok * (int64_t)E + e), acc);
    }
    atomicAdd(dW + ((int64_t)f * 3 + j) * E + e, acc);
  }
  if (threadIdx.x == 0) {
    float bsum = 0.0f;
    for (int64_t n = n0; n < n1; ++n) {
      const float p = __ldg(pooled + n * F + f);
      if (p > 0.0f) bsum += __ldg(gpooled + n * F + f);
    }
    if (bsum != 0.0f) atomicAdd(db + f, bsum);
  }
}

// ---- half-precision rows (f16 / bf16 conv modes): the forward read the shadow table, so the exact
// gradient of what it computed is sum_n gy * shadow_row; rows are 2*Epad bytes instead of 4*E, which
// halves the L2 gather traffic that bounds this kernel.  Thread = (window row j, 16-byte chunk of
// 8 columns); documents are processed four at a time so the dependent loads (argmax -> token id ->
// row) of different documents overlap.
constexpr int HTHREADS = 128;
constexpr int HBATCH = 4;

template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&x)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& u, float (&x)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& u, float (&x)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
}

// Per batch of HBATCH documents the loads form a dependent chain  (pooled, grad, argmax) -> token id -> row.
// The loop is software-pipelined three deep: while the rows of batch k are fetched and accumulated, the token
// ids of batch k+1 and the (pooled, grad, argmax) triples of batch k+2 are already in flight, so a thread
// exposes one memory latency per batch instead of three.
template <bool RAGGED>
struct WgMeta {
  float g[HBATCH];
  int a[HBATCH];
  int64_t base[RAGGED ? HBATCH : 1];
  int len[RAGGED ? HBATCH : 1];
};

template <typename T, int NV, bool RAGGED>
__global__ void __launch_bounds__(HTHREADS) conv_wgrad_half_kernel(
    const uint8_t* __restrict__ shadow, int64_t V, int row_bytes, int E, const int64_t* __restrict__ idx,
    const int32_t* __restrict__ tok32, const int64_t* __restrict__ off, int64_t pad_id, int64_t N, int Tn,
    const int32_t* __restrict__ argmax, const float* __restrict__ pooled, const float* __restrict__ gpooled,
    int F, float* __restrict__ dW, float* __restrict__ db) {
  const int f = blockIdx.x;
  const int64_t per = (N + gridDim.y - 1) / gridDim.y;
  const int64_t n0 = (int64_t)blockIdx.y * per;
  const int64_t n1 = (n0 + per < N) ? n0 + per : N;
  const int nch = (E + 7) >> 3;                  // 16-byte chunks per row that carry data
  const int nvec = 3 * nch;
  const int tid = threadIdx.x;

  float acc[NV][8];
  int vj[NV], vc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.0f;
    const int q = tid + v * HTHREADS;
    vj[v] = q < nvec ? q / nch : -1;
    vc[v] = q < nvec ? q % nch : 0;
  }
  float bsum = 0.0f;

  // stage A: (pooled, grad, argmax) [+ document extent] of the documents nb .. nb+HBATCH-1
  auto stage_a = [&](int64_t nb, WgMeta<RAGGED>& m) {
#pragma unroll
    for (int b = 0; b < HBATCH; ++b) {
      const int64_t n = nb + b;
      const bool in = n < n1;
      if (RAGGED) {
        m.base[b] = in ? __ldg(off + n) : 0;
        m.len[b] = in ? (int)(__ldg(off + n + 1) - m.base[b]) : 0;
      }
      const float p = in ? __ldg(pooled + n * F + f) : 0.0f;
      const float gg = in ? __ldg(gpooled + n * F + f) : 0.0f;
      m.a[b] = in ? __ldg(argmax + n * F + f) : 0;
      m.g[b] = p > 0.0f ? gg : 0.0f;               // dead ReLU -> no gradient (CTA-uniform)
    }
  };
  // stage B: token id of the document row feeding this thread's window row j (-1 = nothing to add)
  auto stage_b = [&](int64_t nb, const WgMeta<RAGGED>& m, int64_t (&tok)[NV][HBATCH]) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int b = 0; b < HBATCH; ++b) {
        const int pos = m.a[b] + vj[v] - 2;
        const bool live = vj[v] >= 0 && m.g[b] != 0.0f && pos >= 0 && pos < Tn;
        int64_t t = -1;
        if (live) {
          if (RAGGED) t = pos < m.len[b] ? (int64_t)__ldg(tok32 + m.base[b] + pos) : pad_id;   // rows past the stored tokens are padding
          else t = __ldg(idx + (nb + b) * (int64_t)Tn + pos);
        }
        tok[v][b] = t;
      }
    }
  };
  // stage C: gather the rows and accumulate
  auto stage_c = [&](const WgMeta<RAGGED>& m, const int64_t (&tok)[NV][HBATCH]) {
#pragma unroll
    for (int b = 0; b < HBATCH; ++b) bsum += m.g[b];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      uint4 row[HBATCH];
#pragma unroll
      for (int b = 0; b < HBATCH; ++b) {
        row[b] = make_uint4(0u, 0u, 0u, 0u);
        if (tok[v][b] >= 0) row[b] = __ldg(reinterpret_cast<const uint4*>(shadow + tok[v][b] * (int64_t)row_bytes) + vc[v]);
      }
#pragma unroll
      for (int b = 0; b < HBATCH; ++b) {
        float x[8];
        unpack8<T>(row[b], x);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v][i] = fmaf(m.g[b], x[i], acc[v][i]);
      }
    }
  };

  WgMeta<RAGGED> m_c, m_b, m_a;                   // metadata of the batches in stages C, B, A
  int64_t t_c[NV][HBATCH], t_b[NV][HBATCH];
  stage_a(n0, m_c);
  stage_a(n0 + HBATCH, m_b);
  stage_b(n0, m_c, t_c);
  for (int64_t nb = n0; nb < n1; nb += HBATCH) {
    stage_a(nb + 2 * HBATCH, m_a);                // k+2
    stage_b(nb + HBATCH, m_b, t_b);               // k+1
    stage_c(m_c, t_c);                            // k
    m_c = m_b;
    m_b = m_a;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int b = 0; b < HBATCH; ++b) t_c[v][b] = t_b[v][b];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if (vj[v] < 0) continue;
    float* dst = dW + ((int64_t)f * 3 + vj[v]) * E + vc[v] * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (vc[v] * 8 + i < E && acc[v][i] != 0.0f) atomicAdd(dst + i, acc[v][i]);
  }
  if (tid == 0 && bsum != 0.0f) atomicAdd(db + f, bsum);
}
}  // namespace

extern "C" int r4r_conv_wgrad_argmax(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T,
                                     const int32_t* argmax, const float* pooled, const float* gpooled, int F,
                                     float* dW, float* db, void* stream) {
  R4R_REQUIRE(table && idx && argmax && pooled && gpooled && dW && db, R4R_EINVAL, "conv_wgrad: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0 && F > 0, R4R_EINVAL, "conv_wgrad: bad sizes");
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  // document slices: enough CTAs for >= 4 waves of 148 SMs x 8 resident CTAs, >= 16 docs each
  int64_t want = (148 * 8 * 4 + F - 1) / F;
  int64_t S = N / 16;
  if (S > want) S = want;
  if (S < 1) S = 1;
  if (S > 65535) S = 65535;
  dim3 grid((unsigned)F, (unsigned)S);
  const bool vec = (E % 4 == 0) && (reinterpret_cast<uintptr_t>(table) % 16 == 0);
  if (!vec) {
    conv_wgrad_scalar_kernel<<<grid, THREADS, 0, s>>>(table, V, E, idx, N, T, argmax, pooled, gpooled, F, dW, db);
  } else {
    int nvec = 3 * (E / 4);
    R4R_REQUIRE(nvec <= MAXV * THREADS, R4R_EUNSUP, "conv_wgrad: E=%d too wide (max %d)", E, MAXV * THREADS * 4 / 3);
    if (nvec <= THREADS)          conv_wgrad_kernel<1><<<grid, THREADS, 0, s>>>(table, V, E, idx, N, T, argmax, pooled, gpooled, F, dW, db);
    else if (nvec <= 2 * THREADS) conv_wgrad_kernel<2><<<grid, THREADS, 0, s>>>(table, V, E, idx, N, T, argmax, pooled, gpooled, F, dW, db);
    else                          conv_wgrad_kernel<3><<<grid, THREADS, 0, s>>>(table, V, E, idx, N, T, argmax, pooled, gpooled, F, dW, db);
  }
  R4R_CHECK_LAUNCH("conv_wgrad");
  return 0;
}

// ---- fp32 refinement of the tensor-core forward ("f16r" / "bf16r" conv modes).  The tcgen05 kernel has found, for every
// (document, filter), the window with the largest conv value -- computed from half-precision operands.  This kernel
// re-evaluates exactly that window in fp32 from the fp32 word table and the fp32 filters:
//     pooled[n,f] = relu(b[f] + sum_j sum_e table[idx[n, a+j-2], e] * W[f,0,j,e])        a = argmax[n,f]
// i.e. the reference's fp32 value of the selected window (common_pytorch_models.py:29-31).  It differs from the fp32
// max-pool only where a second window lies within the half-precision rounding noise of the winner (then by less than
// that noise).  Same gather pattern as the weight gradient: one warp per (document, filter), the filter in shared memory.
constexpr int RTHREADS = 256;
__global__ void __launch_bounds__(RTHREADS) conv_refine_kernel(const float* __restrict__ table, int64_t V, int E,
                                                               const int64_t* __restrict__ idx, int64_t N, int T,
                                                               const int32_t* __restrict__ argmax, const float* __restrict__ conv_w,
                                                               const float* __restrict__ conv_b, int F, float* __restrict__ pooled) {
  extern __shared__ __align__(16) float sw[];         // W[f, 0, :, :]: 3 x E floats
  const int f = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * E; i += RTHREADS) sw[i] = __ldg(conv_w + (int64_t)f * 3 * E + i);
  __syncthreads();
  const float bias = __ldg(conv_b + f);
  const int lane = threadIdx.x & 31;
  const int e4 = E >> 2;
  const int64_t warps = (int64_t)gridDim.y * (RTHREADS / 32);
  for (int64_t n = (int64_t)blockIdx.y * (RTHREADS / 32) + (threadIdx.x >> 5); n < N; n += warps) {
    const int a = __ldg(argmax + n * F + f);
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int pos = a + j - 2;
      if (pos < 0 || pos >= T) continue;               // conv padding row (warp-uniform)
      const int64_t tok = __ldg(idx + n * (int64_t)T + pos);
      if (tok < 0 || tok >= V) __trap();
      const float4* x = reinterpret_cast<const float4*>(table + tok * (int64_t)E);
      const float4* w = reinterpret_cast<const float4*>(sw + j * E);
      for (int c = lane; c < e4; c += 32) {
        const float4 xv = __ldg(x + c), wv = w[c];
        s = fmaf(xv.x, wv.x, s); s = fmaf(xv.y, wv.y, s); s = fmaf(xv.z, wv.z, s); s = fmaf(xv.w, wv.w, s);
      }
      for (int e = 4 * e4 + lane; e < E; e += 32) s = fmaf(__ldg(table + tok * (int64_t)E + e), sw[j * E + e], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      const float o = s + bias;
      pooled[n * F + f] = o > 0.0f ? o : 0.0f;
    }
  }
}

extern "C" int r4r_conv_refine(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T, const int32_t* argmax,
                               const float* conv_w, const float* conv_b, int F, float* pooled, void* stream) {
  R4R_REQUIRE(table && idx && argmax && conv_w && conv_b && pooled, R4R_EINVAL, "conv_refine: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && E <= 4096 && T > 0 && N >= 0 && F > 0, R4R_EINVAL, "conv_refine: bad sizes");
  R4R_REQUIRE(E % 4 != 0 || reinterpret_cast<uintptr_t>(table) % 16 == 0, R4R_EINVAL, "conv_refine: table must be 16-byte aligned");
  if (N == 0) return 0;
  int64_t S = cdiv64(N, RTHREADS / 32);
  const int64_t want = (148 * 8 * 4 + F - 1) / F;       // ~4 waves of 8 resident CTAs per SM over the F filters
  if (S > want) S = want;
  if (S < 1) S = 1;
  dim3 grid((unsigned)F, (unsigned)S);
  conv_refine_kernel<<<grid, RTHREADS, (size_t)3 * E * sizeof(float), as_stream(stream)>>>(table, V, (E % 4 == 0) ? E : E, idx, N, T, argmax, conv_w,
                                                                                           conv_b, F, pooled);
  R4R_CHECK_LAUNCH("conv_refine");
  return 0;
}

static int conv_wgrad_h_launch(const void* shadow, int64_t V, int Epad, int E, int dtype, const int64_t* idx,
                               const int32_t* tok32, const int64_t* off, int64_t pad_id, int64_t N,
                               int T, const int32_t* argmax, const float* pooled, const float* gpooled, int F,
                               float* dW, float* db, void* stream) {
  R4R_REQUIRE(shadow && (idx || (tok32 && off)) && argmax && pooled && gpooled && dW && db, R4R_EINVAL, "conv_wgrad_h: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0 && F > 0, R4R_EINVAL, "conv_wgrad_h: bad sizes");
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_wgrad_h: dtype %d", dtype);
  R4R_REQUIRE(Epad % 8 == 0 && Epad >= ((E + 7) / 8) * 8 && reinterpret_cast<uintptr_t>(shadow) % 16 == 0, R4R_EINVAL,
              "conv_wgrad_h: shadow rows must be 16-byte aligned with Epad %% 8 == 0 and Epad >= E (Epad=%d)", Epad);
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  int64_t want = (148 * 16 * 4 + F - 1) / F;     // >= 4 waves of 148 SMs x 16 resident CTAs
  int64_t S = N / 16;
  if (S > want) S = want;
  if (S < 1) S = 1;
  if (S > 65535) S = 65535;
  dim3 grid((unsigned)F, (unsigned)S);
  const int nvec = 3 * ((E + 7) / 8);
  R4R_REQUIRE(nvec <= 3 * HTHREADS, R4R_EUNSUP, "conv_wgrad_h: E=%d too wide (max %d)", E, HTHREADS * 8);
  const uint8_t* sh = static_cast<const uint8_t*>(shadow);
  const int rb = Epad * 2;
#define R4R_WG_LAUNCH(TYPE, NV) \
  do {                                                                                                                         \
    if (tok32) conv_wgrad_half_kernel<TYPE, NV, true><<<grid, HTHREADS, 0, s>>>(sh, V, rb, E, idx, tok32, off, pad_id, N, T, argmax, pooled, gpooled, F, dW, db); \
    else conv_wgrad_half_kernel<TYPE, NV, false><<<grid, HTHREADS, 0, s>>>(sh, V, rb, E, idx, tok32, off, pad_id, N, T, argmax, pooled, gpooled, F, dW, db);      \
  } while (0)
  if (dtype == R4R_DT_F16) {
    if (nvec <= HTHREADS) R4R_WG_LAUNCH(__half, 1); else if (nvec <= 2 * HTHREADS) R4R_WG_LAUNCH(__half, 2); else R4R_WG_LAUNCH(__half, 3);
  } else {
    if (nvec <= HTHREADS) R4R_WG_LAUNCH(__nv_bfloat16, 1); else if (nvec <= 2 * HTHREADS) R4R_WG_LAUNCH(__nv_bfloat16, 2); else R4R_WG_LAUNCH(__nv_bfloat16, 3);
  }
#undef R4R_WG_LAUNCH
  R4R_CHECK_LAUNCH("conv_wgrad_h");
  return 0;
}

extern "C" int r4r_conv_wgrad_argmax_h(const void* shadow, int64_t V, int Epad, int E, int dtype, const int64_t* idx, int64_t N,
                                       int T, const int32_t* argmax, const float* pooled, const float* gpooled, int F,
                                       float* dW, float* db, void* stream) {
  R4R_REQUIRE(idx, R4R_EINVAL, "conv_wgrad_h: null pointer");
  return conv_wgrad_h_launch(shadow, V, Epad, E, dtype, idx, nullptr, nullptr, 0, N, T, argmax, pooled, gpooled, F, dW, db, stream);
}

extern "C" int r4r_conv_wgrad_argmax_h_ragged(const void* shadow, int64_t V, int Epad, int E, int dtype, const int32_t* tokens,
                                              const int64_t* offsets, int64_t pad_id, int64_t N, int T, const int32_t* argmax,
                                              const float* pooled, const float* gpooled, int F, float* dW, float* db,
                                              void* stream) {
  R4R_REQUIRE(tokens && offsets && pad_id >= 0 && pad_id < V, R4R_EINVAL, "conv_wgrad_h_ragged: null pointer or pad id outside the table");
  return conv_wgrad_h_launch(shadow, V, Epad, E, dtype, nullptr, tokens, offsets, pad_id, N, T, argmax, pooled, gpooled, F, dW, db,
                             stream);
}
