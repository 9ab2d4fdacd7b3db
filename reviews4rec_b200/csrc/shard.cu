// shard.cu -- K8: device side of the row-sharded table lookups.
//
// The reference is single-process (no NCCL / torch.distributed call anywhere), so the
// semantics these kernels keep are simply those of nn.Embedding / Tensor.gather over the FULL table
// (DeepCoNN.py:53-54,70-71, NARRE.py:87-88,110-116, TransNet.py:108-109, MF.py:45-46,52-53) when row
// r of the table lives on rank r % P at local row r / P.
//
// A word lookup is   mark -> plan -> [ids out] -> serve -> [rows back] -> place   (the rows land in a per-step
// cache indexed by the ORIGINAL token id, so the conv / wgrad kernels read their usual ids: no remapped id
// tensors); an id-table lookup is   bucket -> [ids out] -> serve -> [rows back] -> gather.   The two bracketed
// steps are either NCCL all-to-alls (equal splits, CUDA-graph capturable) or -- fused variant --
// r4r_shard_serve_p2p, which gathers the requested rows and stores them straight into the
// requesters' receive buffers over NVLink peer mappings.
//
// Message layout (int64 words): for every destination / source rank q a block of 1+cap words,
//   msg[q*(1+cap)]       = number of valid rows n_q
//   msg[q*(1+cap)+1+j]   = local row index at the owner, j < n_q
// Row payloads are laid out [q][cap][row_bytes]; the row requested as (q, j) comes back at slot
// q*cap + j, which is what `slot[]` / `pos[]` hold.
#include "common.cuh"

namespace {
constexpr int THREADS = 256;

// ---- word-table plan, step 1: presence BITMAP (bit id of flags[], 32 ids per word) of the token ids of this
// rank's documents.  Padding makes ~60 % of all tokens the SAME id and the rest is Zipfian, so setting global
// bits directly would queue millions of accesses on a handful of L2 sectors.  Every CTA therefore collects the
// ids it sees in a shared-memory bitmap first (ids below MARK_BITS; a lane also skips an id equal to its left
// neighbour's or to the one it handled last) and ORs its non-zero words into the global bitmap once at the end:
// at most one fire-and-forget reduction per (CTA, 32-id word).
constexpr int MARK_THREADS = 512;
constexpr int64_t MARK_BITS = 1 << 20;               // 128 KB of shared memory covers ids < 1,048,576

__global__ void __launch_bounds__(MARK_THREADS) shard_mark_kernel(const int64_t* __restrict__ idx, int64_t n, int64_t V,
                                                                  uint32_t* __restrict__ flags, int64_t bits) {
  extern __shared__ uint32_t bitmap[];               // bits / 32 words
  const int words = (int)((bits + 31) >> 5);
  for (int w = threadIdx.x; w < words; w += MARK_THREADS) bitmap[w] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int U = 4;                                 // ids per lane per trip: four independent loads in flight
  const int64_t stride = (int64_t)gridDim.x * MARK_THREADS * U;
  int64_t mine = -1;                                   // the id this lane handled last
  for (int64_t i0 = ((int64_t)blockIdx.x * MARK_THREADS + (threadIdx.x & ~31)) * U; i0 < n; i0 += stride) {   // warp-uniform trip count
    int64_t id[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      id[u] = i < n ? __ldg(idx + i) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (id[u] >= V) __trap();                        // the reference device-asserts on OOB ids
      if (id[u] < -1) __trap();
      const int64_t left = __shfl_up_sync(0xffffffffu, id[u], 1);
      if (id[u] >= 0 && id[u] != mine && (lane == 0 || id[u] != left)) {
        mine = id[u];
        const uint32_t m = 1u << (id[u] & 31);
        if (id[u] < bits) {
          if ((bitmap[id[u] >> 5] & m) == 0u) atomicOr(&bitmap[id[u] >> 5], m);
        } else {
          atomicOr(flags + (id[u] >> 5), m);           // beyond the shared bitmap: rare ids of a very large vocabulary
        }
      }
    }
  }
  __syncthreads();
  for (int w = threadIdx.x; w < words; w += MARK_THREADS) {
    const uint32_t b = bitmap[w];
    if (b) atomicOr(flags + w, b);
  }
}

// ---- word-table plan, step 2: every flagged id goes into the request block of its owner (id % P) as the
// owner-local row id / P.  The order inside a block is irrelevant -- the rows are placed back by id
// (shard_place_kernel) -- so the compaction is a warp-aggregated atomic append, spread over the whole grid.
__global__ void __launch_bounds__(32) shard_plan_zero_kernel(int64_t* __restrict__ req, int P, int64_t cap) {
  if ((int)threadIdx.x < P) req[(int64_t)threadIdx.x * (1 + cap)] = 0;
}

__global__ void __launch_bounds__(THREADS) shard_plan_kernel(uint32_t* __restrict__ flags, int64_t V, int P, int64_t cap,
                                                             int64_t* __restrict__ req) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i0 = (int64_t)blockIdx.x * THREADS + (threadIdx.x & ~31); i0 < V; i0 += stride) {   // one bitmap word per warp trip
    const int64_t id = i0 + lane;
    const uint32_t word = flags[i0 >> 5];
    const bool on = id < V && ((word >> lane) & 1u);
    __syncwarp();
    if (lane == 0 && word) flags[i0 >> 5] = 0u;         // leave the bitmap clean for the next step
    const int o = on ? (int)(id % P) : -1 - lane;       // distinct negatives never match
    const unsigned grp = __match_any_sync(0xffffffffu, o);
    if (on) {
      const int leader = __ffs(grp) - 1;
      unsigned long long base = 0;
      int64_t* block = req + (int64_t)o * (1 + cap);
      if (lane == leader) base = atomicAdd(reinterpret_cast<unsigned long long*>(block), (unsigned long long)__popc(grp));
      base = __shfl_sync(grp, base, leader);
      const int64_t p = (int64_t)base + __popc(grp & ((1u << lane) - 1u));
      if (p >= cap) __trap();
      block[1 + p] = id / P;
    }
  }
}

// ---- id-table plan: no de-duplication (n is a few rows per rating); pos[i] = slot of ids[i]
__global__ void __launch_bounds__(THREADS) shard_bucket_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t R, int P,
                                                               int64_t cap, int64_t* __restrict__ req, int64_t* __restrict__ pos) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    const int64_t id = __ldg(ids + i);
    if (id < 0 || id >= R) __trap();
    const int o = (int)(id % P);
    int64_t* block = req + (int64_t)o * (1 + cap);
    const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long*>(block), 1ULL);
    block[1 + k] = id / P;
    pos[i] = (int64_t)o * cap + (int64_t)k;
  }
}

// ---- owner side: copy the requested rows of the local shard into per-requester payload blocks.
// One warp per (requester, j) row, 16-byte lanes when the row allows it.  `out_ptrs[q]` is where
// requester q's block [cap][row_bytes] starts: a local staging buffer (NCCL path) or requester q's
// receive buffer mapped over NVLink (fused path: the gather IS the all-to-all).
struct ServeArgs {
  uint8_t* out[16];
};

template <bool VEC16>
__global__ void __launch_bounds__(THREADS) shard_serve_kernel(const uint8_t* __restrict__ shard, int64_t rows_local, int row_bytes,
                                                              const int64_t* __restrict__ rreq, int P, int64_t cap,
                                                              const __grid_constant__ ServeArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  const int64_t total = (int64_t)P * cap;
  for (int64_t w = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); w < total; w += warps) {
    const int q = (int)(w / cap);
    const int64_t j = w - (int64_t)q * cap;
    const int64_t* block = rreq + (int64_t)q * (1 + cap);
    const int64_t nq = __ldg(block);
    if (nq < 0 || nq > cap) __trap();
    if (j >= nq) continue;
    const int64_t row = __ldg(block + 1 + j);
    if (row < 0 || row >= rows_local) __trap();
    const uint8_t* src = shard + row * (int64_t)row_bytes;
    uint8_t* dst = A.out[q] + j * (int64_t)row_bytes;
    if (VEC16) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (int c = lane; c < row_bytes / 16; c += 32) d4[c] = __ldg(s4 + c);
    } else {
      const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
      uint32_t* d1 = reinterpret_cast<uint32_t*>(dst);
      for (int c = lane; c < row_bytes / 4; c += 32) d1[c] = __ldg(s1 + c);
    }
  }
}

// ---- requester side: the rows came back compact ([q][j] = the j-th row asked of owner q); copy each to
// cache[id] with id = req[q][1+j] * P + q, so the cache is indexed by the original token id.
template <bool VEC16>
__global__ void __launch_bounds__(THREADS) shard_place_kernel(const uint8_t* __restrict__ rows, const int64_t* __restrict__ req, int P,
                                                              int64_t cap, int row_bytes, uint8_t* __restrict__ cache, int64_t V) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  const int64_t total = (int64_t)P * cap;
  for (int64_t w = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); w < total; w += warps) {
    const int q = (int)(w / cap);
    const int64_t j = w - (int64_t)q * cap;
    const int64_t* block = req + (int64_t)q * (1 + cap);
    if (j >= __ldg(block)) continue;
    const int64_t id = __ldg(block + 1 + j) * P + q;
    if (id < 0 || id >= V) __trap();
    const uint8_t* src = rows + w * (int64_t)row_bytes;
    uint8_t* dst = cache + id * (int64_t)row_bytes;
    if (VEC16) {
      for (int c = lane; c < row_bytes / 16; c += 32) reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
    } else {
      for (int c = lane; c < row_bytes / 4; c += 32) reinterpret_cast<uint32_t*>(dst)[c] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
    }
  }
}

// ---- owner side of the backward: gtable[row(q,j), :] += grads[q*cap + j, :] for the valid (q, j).
// Same warp-segmented combining as r4r_rows_scatter_add (one atomic per distinct row per warp).
__global__ void __launch_bounds__(THREADS) shard_scatter_add_kernel(const float* __restrict__ grads, const int64_t* __restrict__ rreq,
                                                                    int P, int64_t cap, int L, float* __restrict__ gtable,
                                                                    int64_t rows_local, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)P * cap;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i0 = (int64_t)blockIdx.x * THREADS + (threadIdx.x & ~31); i0 < total; i0 += stride) {
    const int64_t i = i0 + lane;
    bool valid = i < total;
    int64_t row = -1 - lane;                           // distinct negatives never match
    if (valid) {
      const int q = (int)(i / cap);
      const int64_t j = i - (int64_t)q * cap;
      const int64_t* block = rreq + (int64_t)q * (1 + cap);
      valid = j < __ldg(block);
      if (valid) {
        row = __ldg(block + 1 + j);
        if (row < 0 || row >= rows_local) __trap();
      }
    }
    if (__ballot_sync(0xffffffffu, valid) == 0u) continue;
    const unsigned grp = __match_any_sync(0xffffffffu, row);
    const int leader = __ffs(grp) - 1;
    for (int c = 0; c < L; ++c) {
      const float g = valid ? __ldg(grads + i * (int64_t)L + c) : 0.0f;
      float s = 0.0f;
      unsigned rem = grp;
      while (rem) {
        const int src = __ffs(rem) - 1;
        s += __shfl_sync(grp, g, src);
        rem &= rem - 1;
      }
      if (valid && lane == leader) atomicAdd(gtable + row * (int64_t)L + c, s * scale);
    }
  }
}

inline unsigned grid_of(int64_t items, int per_block, int cap_blocks = 148 * 8) {
  int64_t b = cdiv64(items, per_block);
  if (b < 1) b = 1;
  if (b > cap_blocks) b = cap_blocks;
  return (unsigned)b;
}
}  // namespace

extern "C" int r4r_shard_mark(const int64_t* idx, int64_t n, int64_t V, int32_t* flags, void* stream) {
  // flags: >= ceil(V/32) zeroed 32-bit words used as a bitmap (bit id % 32 of word id / 32)
  R4R_REQUIRE(flags && (n == 0 || idx), R4R_EINVAL, "shard_mark: null pointer");
  R4R_REQUIRE(n >= 0 && V > 0, R4R_EINVAL, "shard_mark: bad sizes");
  if (n == 0) return 0;
  const int64_t bits = V < MARK_BITS ? V : MARK_BITS;
  const size_t smem = (size_t)((bits + 31) / 32) * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set) {
    R4R_CUDA(cudaFuncSetAttribute(shard_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MARK_BITS / 8)));
    attr_set = true;
  }
  // few, fat CTAs: every CTA publishes its own bitmap, so their number bounds the global flag traffic
  int64_t blocks = cdiv64(n, MARK_THREADS * 8);
  const int64_t cap_blocks = smem > 48 * 1024 ? 148 : 148 * 4;      // small bitmaps: four CTAs per SM keep more loads in flight
  if (blocks > cap_blocks) blocks = cap_blocks;
  if (blocks < 1) blocks = 1;
  shard_mark_kernel<<<(unsigned)blocks, MARK_THREADS, smem, as_stream(stream)>>>(idx, n, V, reinterpret_cast<uint32_t*>(flags), bits);
  R4R_CHECK_LAUNCH("shard_mark");
  return 0;
}

extern "C" int r4r_shard_plan(int32_t* flags, int64_t V, int P, int64_t cap, int64_t* req, void* stream) {
  R4R_REQUIRE(flags && req, R4R_EINVAL, "shard_plan: null pointer");
  R4R_REQUIRE(V > 0 && P >= 1 && P <= 16 && cap >= (V + P - 1) / P, R4R_EINVAL,
              "shard_plan: need 1 <= P <= 16 and cap >= ceil(V/P) (V=%lld P=%d cap=%lld)", (long long)V, P, (long long)cap);
  shard_plan_zero_kernel<<<1, 32, 0, as_stream(stream)>>>(req, P, cap);
  shard_plan_kernel<<<grid_of(V, THREADS, 148 * 2), THREADS, 0, as_stream(stream)>>>(reinterpret_cast<uint32_t*>(flags), V, P, cap, req);
  R4R_CHECK_LAUNCH("shard_plan");
  return 0;
}

extern "C" int r4r_shard_bucket(const int64_t* ids, int64_t n, int64_t R, int P, int64_t cap, int64_t* req, int64_t* pos,
                                void* stream) {
  R4R_REQUIRE(req && (n == 0 || (ids && pos)), R4R_EINVAL, "shard_bucket: null pointer");
  R4R_REQUIRE(n >= 0 && R > 0 && P >= 1 && P <= 16 && cap >= n, R4R_EINVAL,
              "shard_bucket: need 1 <= P <= 16 and cap >= n (n=%lld cap=%lld)", (long long)n, (long long)cap);
  R4R_CUDA(cudaMemsetAsync(req, 0, (size_t)P * (size_t)(1 + cap) * sizeof(int64_t), as_stream(stream)));
  if (n == 0) return 0;
  shard_bucket_kernel<<<grid_of(n, THREADS), THREADS, 0, as_stream(stream)>>>(ids, n, R, P, cap, req, pos);
  R4R_CHECK_LAUNCH("shard_bucket");
  return 0;
}

static int serve_launch(const void* shard, int64_t rows_local, int row_bytes, const int64_t* rreq, int P, int64_t cap,
                        const ServeArgs& A, bool aligned, void* stream) {
  const int64_t total = (int64_t)P * cap;
  const unsigned grid = grid_of(total, THREADS / 32, 148 * 16);
  const uint8_t* s = static_cast<const uint8_t*>(shard);
  if (aligned && row_bytes % 16 == 0)
    shard_serve_kernel<true><<<grid, THREADS, 0, as_stream(stream)>>>(s, rows_local, row_bytes, rreq, P, cap, A);
  else
    shard_serve_kernel<false><<<grid, THREADS, 0, as_stream(stream)>>>(s, rows_local, row_bytes, rreq, P, cap, A);
  R4R_CHECK_LAUNCH("shard_serve");
  return 0;
}

extern "C" int r4r_shard_serve(const void* shard, int64_t rows_local, int row_bytes, const int64_t* rreq, int P, int64_t cap,
                               void* out, void* stream) {
  R4R_REQUIRE(shard && rreq && out, R4R_EINVAL, "shard_serve: null pointer");
  R4R_REQUIRE(rows_local > 0 && row_bytes > 0 && row_bytes % 4 == 0 && P >= 1 && P <= 16 && cap > 0, R4R_EINVAL,
              "shard_serve: bad sizes (row_bytes=%d must be a multiple of 4, 1 <= P <= 16)", row_bytes);
  ServeArgs A;
  for (int q = 0; q < 16; ++q) A.out[q] = q < P ? static_cast<uint8_t*>(out) + (size_t)q * (size_t)cap * (size_t)row_bytes : nullptr;
  const bool aligned = ((reinterpret_cast<uintptr_t>(shard) | reinterpret_cast<uintptr_t>(out)) % 16) == 0;
  return serve_launch(shard, rows_local, row_bytes, rreq, P, cap, A, aligned, stream);
}

extern "C" int r4r_shard_serve_p2p(const void* shard, int64_t rows_local, int row_bytes, const int64_t* rreq, int P, int64_t cap,
                                   void* const* out_ptrs_host, void* stream) {
  R4R_REQUIRE(shard && rreq && out_ptrs_host, R4R_EINVAL, "shard_serve_p2p: null pointer");
  R4R_REQUIRE(rows_local > 0 && row_bytes > 0 && row_bytes % 4 == 0 && P >= 1 && P <= 16 && cap > 0, R4R_EINVAL,
              "shard_serve_p2p: bad sizes (row_bytes=%d must be a multiple of 4, 1 <= P <= 16)", row_bytes);
  ServeArgs A;
  uintptr_t bits = reinterpret_cast<uintptr_t>(shard);
  for (int q = 0; q < 16; ++q) {
    A.out[q] = q < P ? static_cast<uint8_t*>(out_ptrs_host[q]) : nullptr;
    if (q < P) {
      R4R_REQUIRE(A.out[q], R4R_EINVAL, "shard_serve_p2p: null receive pointer for rank %d", q);
      bits |= reinterpret_cast<uintptr_t>(A.out[q]);
    }
  }
  return serve_launch(shard, rows_local, row_bytes, rreq, P, cap, A, bits % 16 == 0, stream);
}

extern "C" int r4r_shard_place(const void* rows, const int64_t* req, int P, int64_t cap, int row_bytes, void* cache, int64_t V,
                               void* stream) {
  R4R_REQUIRE(rows && req && cache, R4R_EINVAL, "shard_place: null pointer");
  R4R_REQUIRE(P >= 1 && P <= 16 && cap > 0 && row_bytes > 0 && row_bytes % 4 == 0 && V > 0, R4R_EINVAL, "shard_place: bad sizes");
  const bool v16 = row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(cache)) % 16 == 0;
  const unsigned grid = grid_of((int64_t)P * cap, THREADS / 32);
  if (v16) shard_place_kernel<true><<<grid, THREADS, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(rows), req, P, cap, row_bytes, static_cast<uint8_t*>(cache), V);
  else shard_place_kernel<false><<<grid, THREADS, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(rows), req, P, cap, row_bytes, static_cast<uint8_t*>(cache), V);
  R4R_CHECK_LAUNCH("shard_place");
  return 0;
}

extern "C" int r4r_shard_scatter_add(const float* grads, const int64_t* rreq, int P, int64_t cap, int L, float* gtable,
                                     int64_t rows_local, float scale, void* stream) {
  R4R_REQUIRE(grads && rreq && gtable, R4R_EINVAL, "shard_scatter_add: null pointer");
  R4R_REQUIRE(P >= 1 && P <= 16 && cap > 0 && L > 0 && rows_local > 0, R4R_EINVAL, "shard_scatter_add: bad sizes");
  shard_scatter_add_kernel<<<grid_of((int64_t)P * cap, THREADS), THREADS, 0, as_stream(stream)>>>(grads, rreq, P, cap, L, gtable,
                                                                                                 rows_local, scale);
  R4R_CHECK_LAUNCH("shard_scatter_add");
  return 0;
}
