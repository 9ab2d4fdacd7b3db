// gather.cu -- row gathers / scatters (HBM-bound integer+copy work; no tensor cores).
//   K1 r4r_word_gather_f32   : materialising word-embedding gather (bit-exact copy)
//      r4r_shadow_build      : fp32 table -> fp16/bf16 padded shadow table
//   K5 r4r_rows_gather       : id-embedding / bias gathers
//   K6 r4r_rows_scatter_add  : their gradient scatter (warp-segmented atomics)
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// K1: one warp per token; lanes sweep the row with 128-bit loads/stores when E % 4 == 0.
// Algorithmic bytes per token: 8 (id) + 4E (row read) + 4E (row write).
template <bool VEC4>
__global__ void __launch_bounds__(256) word_gather_kernel(const float* __restrict__ table, int64_t V, int E,
                                                          const int64_t* __restrict__ idx, int64_t n,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < n; tok += warps) {
    int64_t id = __ldg(idx + tok);
    if (id < 0 || id >= V) __trap();          // the reference device-asserts on OOB ids
    if (VEC4) {
      const float4* src = reinterpret_cast<const float4*>(table + id * (int64_t)E);
      float4* dst = reinterpret_cast<float4*>(out + tok * (int64_t)E);
      const int e4 = E >> 2;
      for (int c = lane; c < e4; c += 32) stg_stream_f4(dst + c, __ldg(src + c));
    } else {
      const float* src = table + id * (int64_t)E;
      float* dst = out + tok * (int64_t)E;
      for (int c = lane; c < E; c += 32) dst[c] = __ldg(src + c);
    }
  }
}

extern "C" int r4r_word_gather_f32(const float* table, int64_t V, int E, const int64_t* idx, int64_t n,
                                   float* out, void* stream) {
  R4R_REQUIRE(table && idx && out, R4R_EINVAL, "word_gather: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && n >= 0, R4R_EINVAL, "word_gather: bad sizes V=%lld E=%d n=%lld", (long long)V, E, (long long)n);
  if (n == 0) return 0;
  const int wpb = 8;
  int64_t blocks = cdiv64(n, wpb);
  const int64_t cap = 148 * 16;               // grid-stride: multiple of the SM count
  if (blocks > cap) blocks = cap;
  bool vec = (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  if (vec) word_gather_kernel<true><<<(unsigned)blocks, wpb * 32, 0, as_stream(stream)>>>(table, V, E, idx, n, out);
  else     word_gather_kernel<false><<<(unsigned)blocks, wpb * 32, 0, as_stream(stream)>>>(table, V, E, idx, n, out);
  R4R_CHECK_LAUNCH("word_gather");
  return 0;
}

// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T cvt_from_f32(float f);
template <> __device__ __forceinline__ __half cvt_from_f32<__half>(float f) { return __float2half_rn(f); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_from_f32<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

template <typename T>
__global__ void __launch_bounds__(256) shadow_build_kernel(const float* __restrict__ table, int64_t V, int E,
                                                           T* __restrict__ shadow, int Epad) {
  const int64_t total = (V + 1) * (int64_t)Epad;        // row V = zeros (the conv's padding rows read it)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i / Epad;
    int e = (int)(i - v * Epad);
    shadow[i] = cvt_from_f32<T>((e < E && v < V) ? table[v * (int64_t)E + e] : 0.0f);
  }
}

extern "C" int r4r_shadow_build(const float* table, int64_t V, int E, void* shadow, int Epad, int dtype, void* stream) {
  R4R_REQUIRE(table && shadow, R4R_EINVAL, "shadow_build: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && Epad >= E && Epad % 8 == 0, R4R_EINVAL, "shadow_build: need Epad>=E, Epad%%8==0 (E=%d Epad=%d)", E, Epad);
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "shadow_build: dtype %d", dtype);
  int64_t total = (V + 1) * (int64_t)Epad;
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == R4R_DT_F16) shadow_build_kernel<__half><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, V, E, (__half*)shadow, Epad);
  else shadow_build_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, V, E, (__nv_bfloat16*)shadow, Epad);
  R4R_CHECK_LAUNCH("shadow_build");
  return 0;
}

// ------------------------------------------------------------------------------------------
// K5: one thread per (row, 16-byte chunk) when rows are whole float4s, else per (row, column) element;
// consecutive threads -> consecutive chunks of one gathered row.
template <int W>
__global__ void __launch_bounds__(256) rows_gather_kernel(const float* __restrict__ table, int64_t R, int L,
                                                          const int64_t* __restrict__ ids, int64_t n,
                                                          float* __restrict__ out) {
  const int lw = L / W;
  const int64_t total = n * (int64_t)lw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / lw;
    int c = (int)(i - r * lw);
    int64_t id = __ldg(ids + r);
    if (id < 0 || id >= R) __trap();
    if constexpr (W == 4) reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(table + id * (int64_t)L) + c);
    else out[i] = __ldg(table + id * (int64_t)L + c);
  }
}

extern "C" int r4r_rows_gather(const float* table, int64_t R, int L, const int64_t* ids, int64_t n, float* out, void* stream) {
  R4R_REQUIRE(table && ids && out, R4R_EINVAL, "rows_gather: null pointer");
  R4R_REQUIRE(R > 0 && L > 0 && n >= 0, R4R_EINVAL, "rows_gather: bad sizes");
  if (n == 0) return 0;
  const bool v4 = L % 4 == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  int64_t blocks = cdiv64(n * (int64_t)(v4 ? L / 4 : L), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (v4) rows_gather_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, R, L, ids, n, out);
  else rows_gather_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, R, L, ids, n, out);
  R4R_CHECK_LAUNCH("rows_gather");
  return 0;
}

// K6: one lane per batch row.  Lanes of a warp that target the same table row are found with
// match.any; the lowest such lane sums the group's rows and issues ONE reduction per W consecutive
// columns (W = 4 / 2 / 1: red.global.add.v4 / .v2 / scalar -- one L2 atomic transaction each), so a hot
// row (NARRE's pad id) costs one atomic per warp instead of 32 and a 32-column row costs
// 8 transactions instead of 32.
template <int W>
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(const float* __restrict__ gout,
                                                               const int64_t* __restrict__ ids, int64_t n, int L,
                                                               float* __restrict__ gtable, int64_t R) {
  const int lane = threadIdx.x & 31;
  const int64_t base_stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += base_stride) {
    int64_t i = i0 + lane;
    bool valid = i < n;
    int64_t id = valid ? __ldg(ids + i) : -1 - lane;     // distinct negatives: never match
    if (valid && (id < 0 || id >= R)) __trap();
    unsigned grp = __match_any_sync(0xffffffffu, id);
    int leader = __ffs(grp) - 1;
    const bool single = grp == (1u << lane);             // the common case: nobody else in the warp hits this row
    for (int c = 0; c < L; c += W) {
      float g[W];
      if constexpr (W == 4) {
        float4 t = valid ? __ldg(reinterpret_cast<const float4*>(gout + i * (int64_t)L + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
      } else if constexpr (W == 2) {
        float2 t = valid ? __ldg(reinterpret_cast<const float2*>(gout + i * (int64_t)L + c)) : make_float2(0.f, 0.f);
        g[0] = t.x; g[1] = t.y;
      } else {
        g[0] = valid ? __ldg(gout + i * (int64_t)L + c) : 0.0f;
      }
      float s[W];
      if (__all_sync(0xffffffffu, single)) {             // warp-uniform fast path: no duplicates in this warp
#pragma unroll
        for (int k = 0; k < W; ++k) s[k] = g[k];
      } else {
#pragma unroll
        for (int k = 0; k < W; ++k) s[k] = 0.0f;
        // segmented sum over the lanes in `grp`, accumulated at the leader
        unsigned rem = grp;
        while (rem) {                                      // uniform within the group
          int src = __ffs(rem) - 1;
#pragma unroll
          for (int k = 0; k < W; ++k) s[k] += __shfl_sync(grp, g[k], src);
          rem &= rem - 1;
        }
      }
      if (valid && lane == leader) {
        float* dst = gtable + id * (int64_t)L + c;
        if constexpr (W == 4) red_add_v4(dst, s[0], s[1], s[2], s[3]);
        else if constexpr (W == 2) red_add_v2(dst, s[0], s[1]);
        else atomicAdd(dst, s[0]);
      }
    }
  }
}

extern "C" int r4r_rows_scatter_add(const float* gout, const int64_t* ids, int64_t n, int L, float* gtable, int64_t R, void* stream) {
  R4R_REQUIRE(gout && ids && gtable, R4R_EINVAL, "rows_scatter_add: null pointer");
  R4R_REQUIRE(R > 0 && L > 0 && n >= 0, R4R_EINVAL, "rows_scatter_add: bad sizes");
  if (n == 0) return 0;
  int64_t blocks = cdiv64(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const uintptr_t al = reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gtable);
  if (L % 4 == 0 && al % 16 == 0) rows_scatter_add_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gout, ids, n, L, gtable, R);
  else if (L % 2 == 0 && al % 8 == 0) rows_scatter_add_kernel<2><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gout, ids, n, L, gtable, R);
  else rows_scatter_add_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gout, ids, n, L, gtable, R);
  R4R_CHECK_LAUNCH("rows_scatter_add");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Ragged -> padded documents.  The reference's fast reader keeps every rating's documents padded to
// input_length as int64 in host RAM and ships 24 KB per rating to the device each batch
// (data_fast.py:32-44,99-109; layout written by make_quick_data.py:21-44).  Our reader keeps only the
// tokens before the trailing padding run as int32 (reviews4rec_b200/readers.py), copies those, and this
// kernel rebuilds the exact [N, T] int64 tensor the models consume: out[n, t] = tokens[off[n] + t] for
// t < off[n+1] - off[n], else pad_id.
__global__ void __launch_bounds__(256) docs_expand_kernel(const int32_t* __restrict__ tokens, const int64_t* __restrict__ offsets,
                                                          int64_t N, int T, int64_t pad_id, int64_t* __restrict__ out) {
  // one warp per document: 256-byte coalesced stores, no per-element index arithmetic
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N; n += warps) {
    const int64_t lo = __ldg(offsets + n), hi = __ldg(offsets + n + 1);
    if (hi < lo || hi - lo > T) __trap();
    const int len = (int)(hi - lo);
    const int32_t* src = tokens + lo;
    int64_t* dst = out + n * (int64_t)T;
    for (int t = lane; t < T; t += 32) dst[t] = t < len ? (int64_t)__ldg(src + t) : pad_id;
  }
}

extern "C" int r4r_docs_expand(const int32_t* tokens, const int64_t* offsets, int64_t N, int T, int64_t pad_id, int64_t* out,
                               void* stream) {
  R4R_REQUIRE(offsets && out, R4R_EINVAL, "docs_expand: null pointer");
  R4R_REQUIRE(N >= 0 && T > 0, R4R_EINVAL, "docs_expand: bad sizes");
  if (N == 0) return 0;
  R4R_REQUIRE(tokens, R4R_EINVAL, "docs_expand: null token pointer");
  int64_t blocks = cdiv64(N, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  docs_expand_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(tokens, offsets, N, T, pad_id, out);
  R4R_CHECK_LAUNCH("docs_expand");
  return 0;
}
