// conv_tc.cu -- K2 (fast mode): fused word gather -> TextCNN conv on 5th-gen tensor cores
// (tcgen05.mma kind::f16, fp32 accumulators in TMEM) -> max/argmax over positions.
// Replaces nn.Embedding + F.conv2d(pad=(2,0)) + F.relu + F.max_pool1d
// (DeepCoNN.py:53-54, common_pytorch_models.py:26-31).
//
// GEMM view per document: Y[p, f] = sum_{j<3} X[p+j-2, :] . W[f, j, :]   (M = T+2 positions,
// N = filters, K = 3*E).  The three window rows are the SAME gathered rows shifted by one position,
// so the A operand is staged ONCE per 128-position tile and the j-th GEMM reads it through a
// shared-memory descriptor whose start address is advanced by j rows.  That only works if a row
// shift is a constant byte offset, which is why the tile uses the no-swizzle "interleaved" K-major
// UMMA layout stored chunk-major:
//
//      A slot:  [chunk c = 8 consecutive embedding columns (16 B)] [row r] [16 B]
//               address(r, c) = c * (RA*16) + r * 16
//      -> 8x16B core matrices are contiguous (SBO = 128 B between 8-row groups),
//         LBO = RA*16 B between the two K-chunks of one K=16 MMA, and row shift j = +16*j bytes.
//
// The filter bank W (B operand, N x K, K-major) uses the same layout and stays resident in shared
// memory for the whole kernel; it does not fit next to the A ring for all 100 filters at E=300, so
// the filters are split in groups of <= 64 and a CTA owns one group (documents are re-gathered
// once per group; the word table is L2-resident).
//
// Warp roles (288 threads, 1 CTA/SM, persistent over documents):
//   warps 0-3  epilogue : tcgen05.ld accumulator -> running max / argmax-tile in registers across
//                         the tiles of a document -> per-document warp-shuffle reduction
//   warps 4-7  producer : cp.async 16-byte gathers of the shadow-table rows into the A ring
//                         (zero-fill for the padding rows), mbarrier "full" per K-slab
//   warp  8    MMA      : allocates TMEM, one elected lane issues tcgen05.mma, tcgen05.commit
//                         releases ring slots ("empty") and publishes accumulators ("tmem_full")
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TILE_M = 128;            // positions per accumulator tile (UMMA M)
constexpr int RA = 131;                // rows per A slot: 130 needed (128 + 2 halo); odd => the
                                       // chunk stride RA*16 B maps 8 lanes onto 8 distinct bank groups
constexpr int CPS = 8;                 // 16-byte K-chunks per ring slot (one slab = 64 columns)
constexpr int SLOT_BYTES = CPS * RA * 16;
constexpr int NB_MAX = 64;             // filters per CTA (UMMA N), multiple of 16
constexpr int ACC_COLS = 64;           // TMEM columns per accumulator buffer
constexpr int TMEM_COLS = 128;         // two accumulator buffers
constexpr int NUM_EPI_WARPS = 4, NUM_PROD_WARPS = 4;
constexpr int NUM_THREADS = (NUM_EPI_WARPS + NUM_PROD_WARPS + 1) * 32;
constexpr int PROD_THREADS = NUM_PROD_WARPS * 32;
constexpr int MAX_SLOTS = 8;
constexpr int ROWS_PER_THREAD = 9;     // producer thread i copies rows i/8 + 16k, k < 9

struct SharedCtl {
  unsigned long long full[MAX_SLOTS];
  unsigned long long empty[MAX_SLOTS];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float red_val[NUM_EPI_WARPS][NB_MAX];
  int red_pos[NUM_EPI_WARPS][NB_MAX];
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// .ca keeps the gathered lines in L1: the pad row (token 0, ~2/3 of all positions on Amazon-shaped
// documents) and the Zipf head then hit in L1 instead of queueing on a handful of L2 lines.
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (layout_type 0), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A/B = f16 (0) or bf16 (1), both K-major, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Params {
  const uint8_t* shadow;     // [V][Epad] 2-byte elements
  long long row_bytes;       // Epad * 2
  long long V;
  const long long* idx;      // [N][T]
  long long N;
  int T;
  int Kc;                    // 16-byte chunks per window row = ceil(E/16)*2
  int F;
  const uint8_t* wpack;      // per-split operand images
  const float* bias;
  float* pooled;
  int* argmax;
  int fmt;                   // 0 f16, 1 bf16
  int nsplit;
  int nslots;
  int ld_ca;                 // 1: cp.async.ca (L1-allocating) gathers, 0: cp.async.cg
  int dbg;                   // perf experiments only (R4R_CONV_DBG): 1 no row shift, 2 no MMA, 4 no copies, 8 no epilogue compare
  int nb[2];                 // filters (padded to 16) per split
  int f0[2];                 // first filter of each split
  long long wofs[2];         // byte offset of each split image in wpack
};

// ------------------------------------------------------------------------------------------
template <int NB>
__device__ __forceinline__ void epilogue_role(const Params& P, SharedCtl* ctl, int split, int cta_in_split, int ctas_in_split,
                                              int warp, int lane) {
  const int ntiles = (P.T + 2 + TILE_M - 1) / TILE_M;
  const int npos = P.T + 2;
  const int row = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t it = 0;
  for (long long doc = cta_in_split; doc < P.N; doc += ctas_in_split) {
    // running maximum per filter column and the tile it came from (one byte per column, packed
    // four to a register so that 64 columns cost 64 + 16 registers instead of 128)
    float best[NB];
    uint32_t btile[NB / 4];
#pragma unroll
    for (int c = 0; c < NB; ++c) best[c] = -INFINITY;
#pragma unroll
    for (int c = 0; c < NB / 4; ++c) btile[c] = 0u;
    for (int t = 0; t < ntiles; ++t, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(&ctl->tmem_full[buf], ph);
      tc_fence_after();
      const bool valid = (t * TILE_M + row) < npos;
      const uint32_t taddr = ctl->tmem_base + lane_base + buf * ACC_COLS;
      uint32_t tsh[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) tsh[k] = (uint32_t)t << (8 * k);
#pragma unroll
      for (int c0 = 0; c0 < NB; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid && !(P.dbg & 8)) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float x = __uint_as_float(v[c]);
            if (x > best[c0 + c]) {
              best[c0 + c] = x;
              btile[(c0 + c) >> 2] = (btile[(c0 + c) >> 2] & ~(0xffu << (8 * (c & 3)))) | tsh[c & 3];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->tmem_empty[buf]);
    }
    // ---- per-document reduction over the 128 rows: max value, smallest position on ties
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      float v = best[c];
      int p = v == -INFINITY ? 0x7fffffff : (int)((btile[c >> 2] >> (8 * (c & 3))) & 0xffu) * TILE_M + row;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (ov > v || (ov == v && op < p)) { v = ov; p = op; }
      }
      if (lane == 0) { ctl->red_val[warp][c] = v; ctl->red_pos[warp][c] = p; }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (row < NB) {
      float v = ctl->red_val[0][row];
      int p = ctl->red_pos[0][row];
#pragma unroll
      for (int w = 1; w < NUM_EPI_WARPS; ++w) {
        float ov = ctl->red_val[w][row];
        int op = ctl->red_pos[w][row];
        if (ov > v || (ov == v && op < p)) { v = ov; p = op; }
      }
      const int f = P.f0[split] + row;
      if (f < P.F) {
        float o = v + __ldg(P.bias + f);
        P.pooled[doc * P.F + f] = o > 0.0f ? o : 0.0f;
        P.argmax[doc * P.F + f] = p;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
}

__device__ __forceinline__ void producer_role(const Params& P, SharedCtl* ctl, uint8_t* ring, int cta_in_split, int ctas_in_split, int ptid) {
  const int ntiles = (P.T + 2 + TILE_M - 1) / TILE_M;
  const int spt = (P.Kc + CPS - 1) / CPS;                // slabs per tile
  const int c8 = ptid & 7, r0 = ptid >> 3;
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t dst_thread = (uint32_t)(c8 * RA * 16 + r0 * 16);
  const int nslots = P.nslots;

  // Fully asynchronous hand-off: a thread never waits for its own copies.  After issuing the
  // 16-byte gathers of a slab it posts cp.async.mbarrier.arrive.noinc on the slab's "full"
  // barrier, which the hardware triggers once those copies have landed; the MMA warp orders the
  // (generic-proxy) writes before its tensor-core reads with fence.proxy.async after the wait.
  // Token ids of a tile's rows are fetched ONE TILE AHEAD into registers (unchecked, so the nine
  // loads are issued back to back and their HBM latency hides behind the current tile's slabs);
  // slot row r <-> document position t*128 - 2 + r, -1 marks a zero (padding) row.
  auto fetch = [&](long long doc, int t, long long (&out)[ROWS_PER_THREAD]) {
    const long long* drow = P.idx + doc * (long long)P.T;
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
      const int r = r0 + 16 * k;
      const int pos = t * TILE_M - 2 + r;
      out[k] = (r < TILE_M + 2 && pos >= 0 && pos < P.T) ? __ldg(drow + pos) : -1LL;
    }
  };
  uint32_t issued = 0;
  long long doc = cta_in_split;
  int t = 0;
  long long cur[ROWS_PER_THREAD], nxt[ROWS_PER_THREAD];
  if (doc < P.N) fetch(doc, 0, cur);
  while (doc < P.N) {
    long long ndoc = doc;
    int nt = t + 1;
    if (nt == ntiles) { nt = 0; ndoc += ctas_in_split; }
    if (ndoc < P.N) fetch(ndoc, nt, nxt);
    const uint8_t* src[ROWS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
      const long long tok = cur[k];
      if (tok < -1 || tok >= P.V) __trap();               // the reference device-asserts on OOB ids
      src[k] = tok < 0 ? nullptr : P.shadow + tok * P.row_bytes;
    }
    for (int s = 0; s < spt; ++s, ++issued) {
      const uint32_t slot = issued % nslots, round = issued / nslots;
      mbar_wait(&ctl->empty[slot], (round & 1u) ^ 1u);
      const int ch = s * CPS + c8;
      if (ch < P.Kc && !(P.dbg & 4)) {
        const uint32_t dst = ring_base + slot * SLOT_BYTES + dst_thread;
#pragma unroll
        for (int k = 0; k < ROWS_PER_THREAD; ++k) {
          if (r0 + 16 * k < TILE_M + 2) {
            const uint8_t* sp = src[k];
            if (P.ld_ca) cp_async16_ca(dst + k * 256, sp ? sp + ch * 16 : P.shadow, sp ? 16u : 0u);
            else         cp_async16(dst + k * 256, sp ? sp + ch * 16 : P.shadow, sp ? 16u : 0u);
          }
        }
      }
      cp_async_mbar_arrive_noinc(&ctl->full[slot]);
    }
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) cur[k] = nxt[k];
    doc = ndoc;
    t = nt;
  }
  cp_async_wait_all();                                    // no copy may be in flight when the CTA exits
}

__device__ __forceinline__ void mma_role(const Params& P, SharedCtl* ctl, const uint8_t* bsm, const uint8_t* ring, int split,
                                         int cta_in_split, int ctas_in_split, int lane) {
  const int ntiles = (P.T + 2 + TILE_M - 1) / TILE_M;
  const int spt = (P.Kc + CPS - 1) / CPS;
  const int nb = P.nb[split];
  const uint32_t idesc = umma_idesc(P.fmt, nb);
  const uint32_t a_base = smem_u32(ring), b_base = smem_u32(bsm);
  const uint32_t a_lbo = RA * 16, b_lbo = (uint32_t)nb * 16;
  const int nslots = P.nslots;
  uint32_t consumed = 0, it = 0;
  for (long long doc = cta_in_split; doc < P.N; doc += ctas_in_split) {
    for (int t = 0; t < ntiles; ++t, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u);       // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = ctl->tmem_base + buf * ACC_COLS;
      for (int s = 0; s < spt; ++s, ++consumed) {
        const uint32_t slot = consumed % nslots, round = consumed / nslots;
        mbar_wait(&ctl->full[slot], round & 1u);
        fence_proxy_async();                                   // producers' cp.async writes -> async proxy
        tc_fence_after();
        if (lane == 0) {
          const int nk = min(CPS, P.Kc - s * CPS) >> 1;        // K=16 steps in this slab
          for (int kk = 0; kk < nk; ++kk) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              uint64_t ad = umma_desc(a_base + slot * SLOT_BYTES + (uint32_t)(2 * kk) * a_lbo + ((P.dbg & 1) ? 0 : j * 16), a_lbo, 128);
              uint64_t bd = umma_desc(b_base + (uint32_t)(j * P.Kc + s * CPS + 2 * kk) * b_lbo, b_lbo, 128);
              if (!(P.dbg & 2)) umma_f16(d_tmem, ad, bd, idesc, (s | kk | j) ? 1u : 0u);
            }
          }
          umma_commit(&ctl->empty[slot]);                      // slot reusable once these MMAs retire
          if (s == spt - 1) umma_commit(&ctl->tmem_full[buf]); // accumulator complete
        }
        __syncwarp();
      }
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_pool_tc_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA -> (filter split, index within split)
  const int split = blockIdx.x % P.nsplit;
  const int cta_in_split = blockIdx.x / P.nsplit;
  const int ctas_in_split = (gridDim.x - split + P.nsplit - 1) / P.nsplit;
  const int nb = P.nb[split];
  const uint32_t b_bytes = (uint32_t)(3 * P.Kc * nb * 16);

  uint8_t* bsm = smem;
  uint8_t* ring = smem + ((b_bytes + 127u) & ~127u);
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(ring + (size_t)P.nslots * SLOT_BYTES);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nslots; ++i) { mbar_init(&ctl->full[i], PROD_THREADS); mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&ctl->tmem_full[i], 1); mbar_init(&ctl->tmem_empty[i], NUM_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NUM_EPI_WARPS + NUM_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&ctl->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident filter bank: plain 16-byte copies of the pre-packed operand image
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.wpack + P.wofs[split]);
    uint4* dst = reinterpret_cast<uint4*>(bsm);
    for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += NUM_THREADS) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();                      // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < NUM_EPI_WARPS) {
    if (nb > 48)      epilogue_role<64>(P, ctl, split, cta_in_split, ctas_in_split, warp, lane);
    else if (nb > 32) epilogue_role<48>(P, ctl, split, cta_in_split, ctas_in_split, warp, lane);
    else if (nb > 16) epilogue_role<32>(P, ctl, split, cta_in_split, ctas_in_split, warp, lane);
    else              epilogue_role<16>(P, ctl, split, cta_in_split, ctas_in_split, warp, lane);
  } else if (warp < NUM_EPI_WARPS + NUM_PROD_WARPS) {
    producer_role(P, ctl, ring, cta_in_split, ctas_in_split, threadIdx.x - NUM_EPI_WARPS * 32);
  } else {
    mma_role(P, ctl, bsm, ring, split, cta_in_split, ctas_in_split, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NUM_EPI_WARPS + NUM_PROD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(ctl->tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- weight packing
struct SplitPlan {
  int nsplit, Kc;
  int nb[2], f0[2];
  long long wofs[2], total;
};

inline bool make_plan(int E, int F, SplitPlan& pl) {
  if (E <= 0 || F <= 0 || F > 2 * NB_MAX) return false;
  pl.Kc = ((E + 15) / 16) * 2;
  pl.nsplit = F > NB_MAX ? 2 : 1;
  long long ofs = 0;
  for (int h = 0; h < 2; ++h) {
    pl.f0[h] = h * NB_MAX;
    int cnt = h < pl.nsplit ? ((F - pl.f0[h] < NB_MAX) ? F - pl.f0[h] : NB_MAX) : 0;
    pl.nb[h] = ((cnt + 15) / 16) * 16;
    pl.wofs[h] = ofs;
    ofs += (long long)3 * pl.Kc * pl.nb[h] * 16;
  }
  pl.total = ofs;
  return true;
}

template <typename T> __device__ __forceinline__ T cvt_w(float f);
template <> __device__ __forceinline__ __half cvt_w<__half>(float f) { return __float2half_rn(f); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_w<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

// image[h][(j*Kc + ch)][r][e8] = W[f0[h]+r][j][ch*8+e8]
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int E, int F, T* __restrict__ out, SplitPlan pl) {
  const long long total = pl.total / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int h = (pl.nsplit > 1 && i * 2 >= pl.wofs[1]) ? 1 : 0;
    long long q = i - pl.wofs[h] / 2;
    int e8 = (int)(q & 7); q >>= 3;
    int r = (int)(q % pl.nb[h]); q /= pl.nb[h];
    int ch = (int)(q % pl.Kc);
    int j = (int)(q / pl.Kc);
    int f = pl.f0[h] + r, e = ch * 8 + e8;
    float v = (f < F && e < E) ? w[((long long)f * 3 + j) * E + e] : 0.0f;
    out[i] = cvt_w<T>(v);
  }
}
}  // namespace

extern "C" int64_t r4r_conv_wpack_bytes(int E, int F) {
  SplitPlan pl;
  if (!make_plan(E, F, pl)) return -1;
  return pl.total;
}

extern "C" int r4r_conv_pack_weights(const float* conv_w, int E, int F, void* wpack, int dtype, void* stream) {
  R4R_REQUIRE(conv_w && wpack, R4R_EINVAL, "conv_pack_weights: null pointer");
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_pack_weights: dtype %d", dtype);
  SplitPlan pl;
  R4R_REQUIRE(make_plan(E, F, pl), R4R_EUNSUP, "conv_pack_weights: E=%d F=%d unsupported (F <= %d)", E, F, 2 * NB_MAX);
  long long n = pl.total / 2;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (dtype == R4R_DT_F16) pack_weights_kernel<__half><<<blocks, 256, 0, as_stream(stream)>>>(conv_w, E, F, (__half*)wpack, pl);
  else pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(conv_w, E, F, (__nv_bfloat16*)wpack, pl);
  R4R_CHECK_LAUNCH("conv_pack_weights");
  return 0;
}

extern "C" int r4r_conv_pool_tc(const void* shadow, int64_t V, int Epad, int E, int dtype,
                                const int64_t* idx, int64_t N, int T,
                                const void* wpack, const float* conv_b, int F,
                                float* pooled, int32_t* argmax, void* stream) {
  R4R_REQUIRE(shadow && idx && wpack && conv_b && pooled && argmax, R4R_EINVAL, "conv_pool_tc: null pointer");
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0, R4R_EINVAL, "conv_pool_tc: bad sizes");
  R4R_REQUIRE((T + 2 + TILE_M - 1) / TILE_M <= 256, R4R_EUNSUP, "conv_pool_tc: T=%d exceeds 256 position tiles", T);
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_pool_tc: dtype %d", dtype);
  SplitPlan pl;
  R4R_REQUIRE(make_plan(E, F, pl), R4R_EUNSUP, "conv_pool_tc: E=%d F=%d unsupported (F <= %d)", E, F, 2 * NB_MAX);
  R4R_REQUIRE(Epad % 8 == 0 && Epad >= pl.Kc * 8, R4R_EINVAL, "conv_pool_tc: shadow row width Epad=%d must be a multiple of 8 and >= %d", Epad, pl.Kc * 8);
  R4R_REQUIRE(reinterpret_cast<uintptr_t>(shadow) % 16 == 0 && reinterpret_cast<uintptr_t>(wpack) % 16 == 0, R4R_EINVAL, "conv_pool_tc: shadow/wpack must be 16-byte aligned");
  if (N == 0) return 0;

  static int sm_count = 0;
  static int smem_optin = 0;
  if (sm_count == 0) {
    int dev = 0;
    R4R_CUDA(cudaGetDevice(&dev));
    R4R_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int cc = 0;
    R4R_CUDA(cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev));
    R4R_REQUIRE(cc == 10, R4R_ENODEV, "conv_pool_tc: needs an sm_100 device (found cc %d.x)", cc);
    R4R_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long b_bytes = ((long long)3 * pl.Kc * pl.nb[0] * 16 + 127) & ~127LL;   // split 0 is the largest
  long long avail = (long long)smem_optin - b_bytes - (long long)sizeof(SharedCtl) - 1024;
  int nslots = (int)(avail / SLOT_BYTES);
  if (nslots > MAX_SLOTS) nslots = MAX_SLOTS;
  R4R_REQUIRE(nslots >= 2, R4R_EUNSUP, "conv_pool_tc: E=%d leaves no room for the A ring next to the filter bank", E);
  const size_t smem_bytes = (size_t)(b_bytes + (long long)nslots * SLOT_BYTES + sizeof(SharedCtl));

  Params P;
  P.shadow = static_cast<const uint8_t*>(shadow);
  P.row_bytes = (long long)Epad * 2;
  P.V = V;
  P.idx = reinterpret_cast<const long long*>(idx);
  P.N = N; P.T = T; P.Kc = pl.Kc; P.F = F;
  P.wpack = static_cast<const uint8_t*>(wpack);
  P.bias = conv_b; P.pooled = pooled; P.argmax = argmax;
  P.fmt = dtype; P.nsplit = pl.nsplit; P.nslots = nslots;
  {
    const char* e = getenv("R4R_CONV_LD");
    P.ld_ca = !(e && e[0] == 'c' && e[1] == 'g');
    const char* d = getenv("R4R_CONV_DBG");
    P.dbg = d ? atoi(d) : 0;
  }
  for (int h = 0; h < 2; ++h) { P.nb[h] = pl.nb[h]; P.f0[h] = pl.f0[h]; P.wofs[h] = pl.wofs[h]; }

  R4R_CUDA(cudaFuncSetAttribute(conv_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  long long grid = (long long)sm_count;
  long long work = N * pl.nsplit;
  if (grid > work) grid = work;
  conv_pool_tc_kernel<<<(unsigned)grid, NUM_THREADS, smem_bytes, as_stream(stream)>>>(P);
  R4R_CHECK_LAUNCH("conv_pool_tc");
  return 0;
}
