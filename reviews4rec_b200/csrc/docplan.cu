// docplan.cu -- per-launch work plan of the fused gather+conv+pool kernel.
//
// The readers pad every document to input_length with one repeated token (id 0: data.py:198-199,
// make_quick_data.py:21-44) and that token is embedded like any other (DeepCoNN.py:53-54), so in an
// Amazon-shaped batch ~60 % of all conv windows are the SAME three rows.  Max-pooling cannot tell
// repeated values apart (F.max_pool1d returns the first maximum), which gives an exact shortcut:
//
//   rows s..T-1 of a document all equal  (trailing run, any token id)   and   T' = min(T, s+3)
//   => conv positions 0..s+1 are unchanged, position s+2 is the first all-run window, and positions
//      T'  (run,run,0) and T'+1 (run,0,0) of the document cut to T' rows reproduce positions T and T+1
//      of the full one.  Every skipped position s+3..T-1 repeats the value at s+2 and, coming later,
//      can never be the first maximum.
//
// The conv kernel therefore processes document n as if it had doc_len[n] = T' rows and maps arg-max
// positions >= T' back by + (T - T').  Values are bit-identical to processing all T rows because each
// position is an independent dot product over the same operand rows.
//
// doc_order lists the documents by decreasing tile count so that the persistent CTA pairs, which
// take work items k = cluster, cluster + nclusters, ... stay in step (longest-processing-time first).
#include "common.cuh"

namespace {
constexpr int THREADS = 256;
constexpr int NCLASS = 256;              // position tiles per document (conv_tc.cu: <= 256)

// one warp per document: scan backwards for the start of the trailing run of idx[T-1]
__global__ void __launch_bounds__(THREADS) doc_extent_kernel(const int64_t* __restrict__ idx, int64_t N, int T, int tile,
                                                             int32_t* __restrict__ doc_len, int32_t* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  for (int64_t n = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); n < N; n += warps) {
    const int64_t* row = idx + n * (int64_t)T;
    const int64_t last = __ldg(row + T - 1);
    int s = 0;                                        // start of the trailing run
    for (int hi = T; hi > 0 && s == 0; hi -= 128) {   // four 32-token windows [hi-32(w+1), hi-32w) per trip, loads issued together
      int64_t v[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int t = hi - 32 * (w + 1) + lane;
        v[w] = t >= 0 ? __ldg(row + t) : last;
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const unsigned diff = ~__ballot_sync(0xffffffffu, v[w] == last);
        if (diff && s == 0) s = hi - 32 * (w + 1) + (32 - __clz(diff));   // one past the highest differing position
      }
    }
    if (lane == 0) {
      int len = s + 3 < T ? s + 3 : T;
      doc_len[n] = len;
      int c = (len + 2 + tile - 1) / tile;
      atomicAdd(hist + (c < NCLASS ? c : NCLASS - 1), 1);
    }
  }
}

// ragged documents: the trailing padding run starts where the stored tokens end
__global__ void __launch_bounds__(THREADS) doc_extent_ragged_kernel(const int64_t* __restrict__ offsets, int64_t N, int T, int tile,
                                                                    int32_t* __restrict__ doc_len, int32_t* __restrict__ hist) {
  for (int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x; n < N; n += (int64_t)gridDim.x * THREADS) {
    const int64_t s = __ldg(offsets + n + 1) - __ldg(offsets + n);
    if (s < 0 || s > T) __trap();
    const int len = s + 3 < T ? (int)s + 3 : T;
    doc_len[n] = len;
    const int c = (len + 2 + tile - 1) / tile;
    atomicAdd(hist + (c < NCLASS ? c : NCLASS - 1), 1);
  }
}

// counting sort by tile count, descending; order within a class is arbitrary
__global__ void __launch_bounds__(THREADS) doc_order_kernel(const int32_t* __restrict__ doc_len, int64_t N, int tile,
                                                            const int32_t* __restrict__ hist, int32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ order) {
  __shared__ int base[NCLASS];
  for (int c = threadIdx.x; c < NCLASS; c += THREADS) {
    int b = 0;
    for (int c2 = c + 1; c2 < NCLASS; ++c2) b += hist[c2];
    base[c] = b;
  }
  __syncthreads();
  for (int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x; n < N; n += (int64_t)gridDim.x * THREADS) {
    int c = (doc_len[n] + 2 + tile - 1) / tile;
    if (c >= NCLASS) c = NCLASS - 1;
    order[base[c] + atomicAdd(cursor + c, 1)] = (int32_t)n;
  }
}
}  // namespace

extern "C" int64_t r4r_doc_plan_ws_bytes(void) { return 2 * NCLASS * (int64_t)sizeof(int32_t); }

extern "C" int r4r_doc_plan(const int64_t* idx, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                            void* stream) {
  R4R_REQUIRE(idx && doc_len && doc_order && ws, R4R_EINVAL, "doc_plan: null pointer");
  R4R_REQUIRE(N >= 0 && N < (1LL << 31) && T > 0, R4R_EINVAL, "doc_plan: bad sizes");
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const int tile = 256;                               // positions per CTA-pair tile of conv_pool_tc (2 * TILE_M)
  int32_t* hist = static_cast<int32_t*>(ws);
  int32_t* cursor = hist + NCLASS;
  R4R_CUDA(cudaMemsetAsync(ws, 0, (size_t)r4r_doc_plan_ws_bytes(), s));
  int64_t b = cdiv64(N, THREADS / 32);
  if (b > 148 * 8) b = 148 * 8;
  doc_extent_kernel<<<(unsigned)b, THREADS, 0, s>>>(idx, N, T, tile, doc_len, hist);
  R4R_CHECK_LAUNCH("doc_extent");
  b = cdiv64(N, THREADS);
  if (b > 148) b = 148;
  doc_order_kernel<<<(unsigned)b, THREADS, 0, s>>>(doc_len, N, tile, hist, cursor, doc_order);
  R4R_CHECK_LAUNCH("doc_order");
  return 0;
}

extern "C" int r4r_doc_plan_ragged(const int64_t* offsets, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                                   void* stream) {
  R4R_REQUIRE(offsets && doc_len && doc_order && ws, R4R_EINVAL, "doc_plan_ragged: null pointer");
  R4R_REQUIRE(N >= 0 && N < (1LL << 31) && T > 0, R4R_EINVAL, "doc_plan_ragged: bad sizes");
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const int tile = 256;
  int32_t* hist = static_cast<int32_t*>(ws);
  int32_t* cursor = hist + NCLASS;
  R4R_CUDA(cudaMemsetAsync(ws, 0, (size_t)r4r_doc_plan_ws_bytes(), s));
  int64_t b = cdiv64(N, THREADS);
  if (b > 148) b = 148;
  doc_extent_ragged_kernel<<<(unsigned)b, THREADS, 0, s>>>(offsets, N, T, tile, doc_len, hist);
  R4R_CHECK_LAUNCH("doc_extent_ragged");
  doc_order_kernel<<<(unsigned)b, THREADS, 0, s>>>(doc_len, N, tile, hist, cursor, doc_order);
  R4R_CHECK_LAUNCH("doc_order");
  return 0;
}
