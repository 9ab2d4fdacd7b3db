// conv_simt.cu -- K2 (parity mode): fused word gather -> 3xE conv -> max/argmax over positions,
// exact fp32 FMA arithmetic on CUDA cores.  This is the strict-parity implementation of
// common_pytorch_models.py:26-31 fused with the nn.Embedding gather (DeepCoNN.py:53-54); the
// tensor-core implementation lives in conv_tc.cu and is validated against this one.
//
// Work decomposition: one CTA = one document x TM consecutive output positions x all F filters.
// The CTA stages the TM+2 gathered rows it needs (zero rows outside the document, which is the
// reference's padding=(2,0)) and the filter bank in shared memory in E-slabs of EK columns, and
// every thread owns a 4-position x KF-filter register tile.  Partial maxima of the position
// tiles of one document meet in a packed 64-bit atomicMax key:
//     key = ordered(fp32 value) << 32 | (0xffffffff - position)
// so the maximum value wins and, on exact ties, the smallest position -- the "first max" rule of
// F.max_pool1d (ties are common: padded tails repeat the same window).
#include "common.cuh"

namespace {
constexpr int TM = 64;        // output positions per CTA
constexpr int EK = 16;        // embedding columns per shared-memory slab
constexpr int KF_MAX = 8;     // filters per thread (F <= 128)
constexpr int THREADS = 256;

template <int KF>
__global__ void __launch_bounds__(THREADS) conv_pool_simt_kernel(
    const float* __restrict__ table, int64_t V, int E, const int64_t* __restrict__ idx, int T,
    const float* __restrict__ conv_w, int F, unsigned long long* __restrict__ keys) {
  constexpr int FP = KF * 16;
  __shared__ float Xs[TM + 2][EK + 1];
  __shared__ float Ws[3][EK][FP];
  __shared__ int64_t tok[TM + 2];

  const int tiles = (T + 2 + TM - 1) / TM;
  const int64_t doc = blockIdx.x / tiles;
  const int p0 = (int)(blockIdx.x % tiles) * TM;  // first output position of this tile
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // tokens of the TM+2 input rows: output p reads doc rows p-2, p-1, p
  for (int r = tid; r < TM + 2; r += THREADS) {
    int pos = p0 - 2 + r;
    int64_t t = -1;
    if (pos >= 0 && pos < T) {
      t = __ldg(idx + doc * (int64_t)T + pos);
      if (t < 0 || t >= V) __trap();
    }
    tok[r] = t;
  }

  float acc[4][KF];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < KF; ++k) acc[i][k] = 0.0f;

  for (int e0 = 0; e0 < E; e0 += EK) {
    __syncthreads();
    for (int i = tid; i < (TM + 2) * EK; i += THREADS) {
      int r = i / EK, e = i - r * EK;
      int64_t t = tok[r];
      Xs[r][e] = (t >= 0 && e0 + e < E) ? __ldg(table + t * (int64_t)E + e0 + e) : 0.0f;
    }
    for (int i = tid; i < 3 * EK * FP; i += THREADS) {
      int f = i % FP;
      int je = i / FP;
      int e = je % EK, j = je / EK;
      Ws[j][e][f] = (f < F && e0 + e < E) ? __ldg(conv_w + ((int64_t)f * 3 + j) * E + e0 + e) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int e = 0; e < EK; ++e) {
      float xv[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) xv[i] = Xs[ty * 4 + i][e];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int k = 0; k < KF; ++k) {
          float w = Ws[j][e][tx + 16 * k];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][k] = fmaf(xv[i + j], w, acc[i][k]);
        }
      }
    }
  }

  const int P = T + 2;
#pragma unroll
  for (int k = 0; k < KF; ++k) {
    int f = tx + 16 * k;
    if (f >= F) continue;
    float best = 0.0f;
    int bpos = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int p = p0 + ty * 4 + i;
      if (p < P && (bpos < 0 || acc[i][k] > best)) { best = acc[i][k]; bpos = p; }
    }
    if (bpos >= 0) {
      unsigned long long key = ((unsigned long long)f32_to_ordered(best) << 32) | (0xffffffffu - (unsigned)bpos);
      atomicMax(keys + doc * (int64_t)F + f, key);
    }
  }
}

__global__ void __launch_bounds__(256) conv_pool_finalize_kernel(const unsigned long long* __restrict__ keys,
                                                                 const float* __restrict__ conv_b, int64_t total, int F,
                                                                 float* __restrict__ pooled, int32_t* __restrict__ argmax) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long key = keys[i];
    float v = ordered_to_f32((uint32_t)(key >> 32)) + __ldg(conv_b + (i % F));
    pooled[i] = v > 0.0f ? v : 0.0f;
    argmax[i] = (int32_t)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
  }
}
}  // namespace

extern "C" int r4r_conv_pool_simt(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T,
                                  const float* conv_w, const float* conv_b, int F,
                                  float* pooled, int32_t* argmax, uint64_t* keys_ws, void* stream) {
  R4R_REQUIRE(table && idx && conv_w && conv_b && pooled && argmax && keys_ws, R4R_EINVAL, "conv_pool_simt: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0, R4R_EINVAL, "conv_pool_simt: bad sizes");
  R4R_REQUIRE(F > 0 && F <= 16 * KF_MAX, R4R_EUNSUP, "conv_pool_simt: F=%d not in 1..%d", F, 16 * KF_MAX);
  const int64_t tiles = (T + 2 + TM - 1) / TM;
  R4R_REQUIRE(N * tiles <= 0x7fffffffLL, R4R_EUNSUP, "conv_pool_simt: N=%lld docs too many for one launch", (long long)N);
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  R4R_CUDA(cudaMemsetAsync(keys_ws, 0, (size_t)N * F * sizeof(uint64_t), s));
  dim3 grid((unsigned)(N * tiles));
  auto* keys = reinterpret_cast<unsigned long long*>(keys_ws);
  int kf = (F + 15) / 16;
  switch (kf) {
    case 1: conv_pool_simt_kernel<1><<<grid, THREADS, 0, s>>>(table, V, E, idx, T, conv_w, F, keys); break;
    case 2: conv_pool_simt_kernel<2><<<grid, THREADS, 0, s>>>(table, V, E, idx, T, conv_w, F, keys); break;
    case 3: case 4: conv_pool_simt_kernel<4><<<grid, THREADS, 0, s>>>(table, V, E, idx, T, conv_w, F, keys); break;
    case 5: case 6: case 7: conv_pool_simt_kernel<7><<<grid, THREADS, 0, s>>>(table, V, E, idx, T, conv_w, F, keys); break;
    default: conv_pool_simt_kernel<8><<<grid, THREADS, 0, s>>>(table, V, E, idx, T, conv_w, F, keys); break;
  }
  R4R_CHECK_LAUNCH("conv_pool_simt");
  int64_t total = N * (int64_t)F;
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv_pool_finalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(keys, conv_b, total, F, pooled, argmax);
  R4R_CHECK_LAUNCH("conv_pool_finalize");
  return 0;
}
