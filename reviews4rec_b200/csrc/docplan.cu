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
// doc_order lists the documents by decreasing length class (256 classes of (T+2)/255 windows): the conv launch deals
// work items to its persistent CTA pairs in rounds of alternating direction, which with a sorted list gives every pair
// nearly the same number of conv windows.
// The sort is STABLE and free of global atomics: the same batch always yields the same work order.
#include "common.cuh"

namespace {
constexpr int THREADS = 256;
constexpr int NCLASS = 256;              // length classes of the sort

// one warp per document: scan backwards for the start of the trailing run of idx[T-1]
__global__ void __launch_bounds__(THREADS) doc_extent_kernel(const int64_t* __restrict__ idx, int64_t N, int T,
                                                             int32_t* __restrict__ doc_len) {
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
    if (lane == 0) doc_len[n] = s + 3 < T ? s + 3 : T;
  }
}

// ragged documents: the trailing padding run starts where the stored tokens end
__global__ void __launch_bounds__(THREADS) doc_extent_ragged_kernel(const int64_t* __restrict__ offsets, int64_t N, int T,
                                                                    int32_t* __restrict__ doc_len) {
  for (int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x; n < N; n += (int64_t)gridDim.x * THREADS) {
    const int64_t s = __ldg(offsets + n + 1) - __ldg(offsets + n);
    if (s < 0 || s > T) __trap();
    doc_len[n] = s + 3 < T ? (int)s + 3 : T;
  }
}

// class of a document of `len` rows: its len + 2 conv windows on a scale of 0 .. 255 (len <= T)
__device__ __forceinline__ int len_class(int len, int T) {
  const int c = (int)(((long long)(len + 2) * (NCLASS - 1)) / (T + 2));
  return c < NCLASS ? c : NCLASS - 1;
}

// Stable counting sort by length class, descending (documents of one class keep their batch order, so the
// work order -- and with it every timing-dependent interleaving of the conv kernel -- replays exactly).
// Pass 1: class histogram of each group of THREADS consecutive documents.
__global__ void __launch_bounds__(THREADS) doc_group_hist_kernel(const int32_t* __restrict__ doc_len, int64_t N, int T, int ncls,
                                                                 int32_t* __restrict__ ghist) {
  __shared__ int h[NCLASS];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if (n < N) atomicAdd(&h[len_class(doc_len[n], T)], 1);      // counts only: order-independent
  __syncthreads();
  if ((int)threadIdx.x < ncls) ghist[(int64_t)blockIdx.x * ncls + threadIdx.x] = h[threadIdx.x];
}

// Pass 2: slot of document n = #documents of longer classes + #documents of its class in earlier groups
//         + #documents of its class earlier in its own group.
__global__ void __launch_bounds__(THREADS) doc_order_kernel(const int32_t* __restrict__ doc_len, int64_t N, int T, int ncls,
                                                            const int32_t* __restrict__ ghist, int32_t* __restrict__ order) {
  __shared__ int total[NCLASS], before[NCLASS], base[NCLASS];
  __shared__ int wcnt[THREADS / 32][NCLASS];
  const int c0 = threadIdx.x;                         // THREADS == NCLASS: one thread per class
  if (c0 < ncls) {
    int t = 0, b = 0;
    for (int64_t g = 0; g < (int64_t)gridDim.x; ++g) {
      const int v = ghist[g * ncls + c0];
      t += v;
      if (g < (int64_t)blockIdx.x) b += v;
    }
    total[c0] = t;
    before[c0] = b;
  }
  for (int w = 0; w < THREADS / 32; ++w) wcnt[w][c0] = 0;
  __syncthreads();
  if (c0 < ncls) {
    int b = before[c0];
    for (int c2 = c0 + 1; c2 < ncls; ++c2) b += total[c2];
    base[c0] = b;
  }
  const int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = n < N ? len_class(doc_len[n], T) : -1;
  const unsigned same = __match_any_sync(0xffffffffu, c);
  const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
  if (c >= 0 && rank_in_warp == 0) wcnt[warp][c] = __popc(same);
  __syncthreads();
  if (c >= 0) {
    int slot = base[c] + rank_in_warp;
    for (int w = 0; w < warp; ++w) slot += wcnt[w][c];
    order[slot] = (int32_t)n;
  }
}
}  // namespace


extern "C" int64_t r4r_doc_plan_ws_bytes(int64_t N, int T) {
  if (N < 0 || T <= 0) return -1;
  return (cdiv64(N, THREADS) * NCLASS + 1) * (int64_t)sizeof(int32_t);
}

static int doc_order_launch(int64_t N, int T, const int32_t* doc_len, int32_t* doc_order, void* ws, cudaStream_t s) {
  const int ncls = NCLASS;
  const unsigned groups = (unsigned)cdiv64(N, THREADS);
  // (a one-launch variant in which every block recounts all documents measured 17 us against 4 + 6 us for the two passes)
  int32_t* ghist = static_cast<int32_t*>(ws);
  doc_group_hist_kernel<<<groups, THREADS, 0, s>>>(doc_len, N, T, ncls, ghist);
  R4R_CHECK_LAUNCH("doc_group_hist");
  doc_order_kernel<<<groups, THREADS, 0, s>>>(doc_len, N, T, ncls, ghist, doc_order);
  R4R_CHECK_LAUNCH("doc_order");
  return 0;
}

extern "C" int r4r_doc_plan(const int64_t* idx, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                            void* stream) {
  R4R_REQUIRE(idx && doc_len && doc_order && ws, R4R_EINVAL, "doc_plan: null pointer");
  R4R_REQUIRE(N >= 0 && N < (1LL << 31) && T > 0, R4R_EINVAL, "doc_plan: bad sizes");
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  int64_t b = cdiv64(N, THREADS / 32);
  if (b > 148 * 8) b = 148 * 8;
  doc_extent_kernel<<<(unsigned)b, THREADS, 0, s>>>(idx, N, T, doc_len);
  R4R_CHECK_LAUNCH("doc_extent");
  return doc_order_launch(N, T, doc_len, doc_order, ws, s);
}

extern "C" int r4r_doc_plan_ragged(const int64_t* offsets, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                                   void* stream) {
  R4R_REQUIRE(offsets && doc_len && doc_order && ws, R4R_EINVAL, "doc_plan_ragged: null pointer");
  R4R_REQUIRE(N >= 0 && N < (1LL << 31) && T > 0, R4R_EINVAL, "doc_plan_ragged: bad sizes");
  if (N == 0) return 0;
  cudaStream_t s = as_stream(stream);
  int64_t b = cdiv64(N, THREADS);
  if (b > 148) b = 148;
  doc_extent_ragged_kernel<<<(unsigned)b, THREADS, 0, s>>>(offsets, N, T, doc_len);
  R4R_CHECK_LAUNCH("doc_extent_ragged");
  return doc_order_launch(N, T, doc_len, doc_order, ws, s);
}
