// conv_tc.cu -- K2 (fast mode): fused word gather -> TextCNN conv on 5th-gen tensor cores
// (tcgen05.mma cta_group::2 kind::f16, fp32 accumulators in TMEM) -> ReLU -> max/argmax over
// positions.  Replaces nn.Embedding + F.conv2d(pad=(2,0)) + F.relu + F.max_pool1d
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
// CTA pair.  The filter bank W (B operand, 3*E x 100 fp16 = 180 KB) must stay resident in shared
// memory next to the A ring, which one SM cannot hold, and a single-SM MMA of N <= 64 filters is
// bound by shared-memory operand bandwidth (measured 57 instead of 32 cycles per MMA).  The kernel
// therefore runs as clusters of two CTAs on one TPC issuing tcgen05.mma.cta_group::2 with
// M = 256 positions x N = all filters: each CTA stages its own 128 positions of A and keeps HALF of
// the filter bank; the pair's tensor cores read both halves.  Every document row is gathered
// exactly once.
//
// Warp roles per CTA (416 threads, 1 CTA/SM, persistent over documents):
//   warps 0-7   epilogue : tcgen05.ld of this CTA's 128 x N accumulator (warp w: TMEM lane quarter
//                          w&3, column half w>>2) -> running max / tile-of-max in registers across
//                          the tiles of a document -> per document two redux.sync per column (max,
//                          then the smallest tile<<5|lane key among its holders = FIRST maximum) +
//                          a four-warp shared-memory merge -> the two CTAs' partial (max, argmax)
//                          are merged through distributed shared memory (rank 1 stores into rank 0
//                          with st.async) -> bias, ReLU, store
//   warps 8-11  producer : cp.async.ca 16-byte gathers of the shadow-table rows into a ring of
//                          K=64 slabs (8 chunk columns: each lane owns one column and nine rows, so a
//                          copy's address is a per-row base + a compile-time offset); conv padding
//                          rows read the all-zero row V of the shadow table; token ids are fetched
//                          one tile ahead and document descriptors one document ahead; one mbarrier
//                          arrival per warp per slab on the LEADER's barrier, `lag` slabs after issue
//   warp  12    MMA      : allocates all 512 TMEM columns (both CTAs) = four accumulator buffers; in
//                          the leader CTA one elected lane issues tcgen05.mma, multicast
//                          tcgen05.commit releases ring slots ("empty") and publishes accumulators
//                          ("tmem_full") in both CTAs
//
// Work plan (docplan.cu): documents end in a run of one repeated padding token; every conv window
// inside the run repeats a value max-pooling has already seen, so document n is processed as if it had
// doc_len[n] = min(T, run start + 3) rows and arg-max positions >= doc_len are mapped back by
// + (T - doc_len) -- bit-identical results, ~2.5x less work on Amazon-shaped batches -- and documents
// are issued longest first.  Ragged input (tokens + offsets instead of padded ids) takes the same path.
//
// L1-allocating gathers matter: ~2/3 of all positions of Amazon-shaped documents are the pad token
// and the rest is Zipfian, so with L2-only (cp.async.cg) loads all SMs queue on a handful of L2
// lines (measured 33.7 ms vs 3.7 ms per 4096 documents).
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int TILE_M = 128;            // positions per CTA per accumulator tile (UMMA M = 256 per pair)
constexpr int RA = 131;                // rows per A slot: 130 needed (128 + 2 halo); odd => the
                                       // chunk stride RA*16 B maps 8 lanes onto 8 distinct bank groups
constexpr int N_MAX = 128;             // filters (UMMA N), multiple of 16
constexpr int ACC_STRIDE = 128;        // TMEM columns between consecutive accumulator buffers
#ifndef R4R_NACC
#define R4R_NACC 4
#endif
constexpr int NACC = R4R_NACC;         // accumulator buffers: the MMA warp may run three tiles ahead of the epilogue,
                                       // which hides the per-document reduction (the epilogue drains nothing meanwhile)
constexpr int TMEM_COLS = NACC * ACC_STRIDE;   // 512 = all of TMEM (1 CTA per SM)
constexpr int NUM_EPI_WARPS = 8, NUM_PROD_WARPS = 4;
constexpr int MMA_WARP = NUM_EPI_WARPS + NUM_PROD_WARPS;
constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
constexpr int MAX_SLOTS = 8;
constexpr int ROWS_PER_THREAD = 9;     // producer thread i copies rows i/8 + 16k, k < 9
constexpr int CPS = 8;                 // 16-byte K-chunks per ring slab: one chunk column per producer lane (K = 64 per slab)
constexpr int MAX_SPT = 16;            // slabs per position tile -> Kc <= 128 chunks (E <= 1024)
constexpr int LAG = 2;                 // default: a slab is published after the next LAG ones have been issued
constexpr int MAX_LAG = 4;             // tuning range (R4R_CONV_LAG); the ring needs lag + 2 slots

struct SharedCtl {
  unsigned long long full[MAX_SLOTS];      // leader's copy is used: 2 CTAs x NUM_PROD_WARPS arrivals
  unsigned long long empty[MAX_SLOTS];     // 1 arrival (multicast tcgen05.commit)
  unsigned long long tmem_full[NACC];      // 1 arrival (multicast tcgen05.commit)
  unsigned long long tmem_empty[NACC];     // leader's copy: 2 CTAs x NUM_EPI_WARPS arrivals
  unsigned long long xchg_full[2];         // rank 0: 1 arrival + Npad*8 transaction bytes stored by rank 1 (st.async)
  unsigned long long xchg_empty[2];        // rank 1: one arrival per column thread of rank 0
  uint32_t tmem_base;
  uint32_t pad;
  float red_val[NUM_EPI_WARPS][N_MAX / 2];
  int red_pos[NUM_EPI_WARPS][N_MAX / 2];
  float xchg_val[2][N_MAX];
  int xchg_pos[2][N_MAX];
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
// Plain (CTA-scope release) arrive on a barrier of any CTA of the cluster.  Cluster-scope release /
// acquire would be compiled to MEMBAR.GPU on the arrive and an L1 invalidation (CCTL.IVALL) on the
// wait -- the latter throws away the L1-resident hot rows of the gather.  None of the data these
// barriers guard is read through L1 by the waiting thread: operands go to the tensor cores through
// the async proxy (ordered by fence.proxy.async before the arrive) and the DSMEM exchange uses
// st.async + complete_tx.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// the same with an explicit arrival count (a register operand: lets the caller tie the arrival to earlier loads)
__device__ __forceinline__ void mbar_arrive_cluster_n(uint32_t cluster_addr, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0], %1;" :: "r"(cluster_addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// 4-byte store into another CTA's shared memory that signals completion (4 tx bytes) on a barrier there
__device__ __forceinline__ void st_async_u32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
               :: "r"(cluster_addr), "r"(v), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// .ca keeps the gathered lines in L1 (see the header comment)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}
// 16-byte copy from [src + OFS] (compile-time byte offset folded into the instruction)
template <int OFS>
__device__ __forceinline__ void cp_async16_ca_ofs(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1+%2], 16;" :: "r"(dst), "l"(src), "n"(OFS) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// UMMA shared-memory descriptor (K-major, SWIZZLE_NONE = layout_type 0, version 1 for sm_100):
// bits 0-13 start address >> 4, 16-29 LBO >> 4, 32-45 SBO >> 4, bit 46 version.  mma_role() keeps
// the two 32-bit halves and advances the address field by constants.
// instruction descriptor: D=f32, A/B = f16 (0) or bf16 (1), both K-major, M=256 (pair), N=n
__device__ __forceinline__ uint32_t umma_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TILE_M) >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Diagnostics: time spent inside a barrier wait is added to a counter when profiling is on.
#define TIMED_WAIT(acc, stmt)                         \
  do {                                                \
    if (prof_on) {                                    \
      long long t0__ = clock64();                     \
      stmt;                                           \
      acc += clock64() - t0__;                        \
    } else {                                          \
      stmt;                                           \
    }                                                 \
  } while (0)

struct Params {
  const uint8_t* shadow;     // [V][Epad] 2-byte elements
  long long row_bytes;       // Epad * 2
  long long V;
  const long long* idx;      // [N][T] padded token ids, or NULL when the documents come ragged:
  const int* tok32;          //   tokens of all documents back to back (those before each trailing padding run)
  const long long* off;      //   [N+1] offsets into tok32; rows off[n+1]-off[n] .. T-1 of document n are pad_id
  long long pad_id;
  long long N;
  int T;
  int Kc;                    // 16-byte chunks per window row = ceil(E/16)*2
  int F;
  int Npad;                  // filters padded to a multiple of 16 (UMMA N)
  const uint8_t* wpack;      // two operand images (rank 0: filters [0, Npad/2), rank 1: the rest)
  const float* bias;
  float* pooled;
  int* argmax;
  int fmt;                   // 0 f16, 1 bf16
  int nslots;
  int slot_bytes;
  unsigned long long* prof;  // diagnostics: per-role cycle counters of cluster 0 (r4r_conv_debug_profile), or NULL
  int lag;                   // slabs issued ahead of the one being published (1..MAX_LAG)
  const int* doc_len;        // [N] effective document lengths (r4r_doc_plan), or NULL = T for every document
  const int* doc_order;      // [N] processing order (longest first), or NULL = identity
};

// Work item k of the launch -> (document, its effective length).  A document whose last rows repeat
// one token (the reader's padding, data.py:198-199) is processed as if it ended three rows into that
// run: all later windows would reproduce values already seen (see r4r_doc_plan in the header).
__device__ __forceinline__ void work_item(const Params& P, long long k, long long& doc, int& Td) {
  doc = P.doc_order ? (long long)__ldg(P.doc_order + k) : k;
  Td = P.doc_len ? __ldg(P.doc_len + doc) : P.T;
}
// producer's view of a work item: also where the document's stored tokens are (ragged input)
struct DocRef {
  long long doc;
  int Td, npt;
  long long base;            // ragged: offset of the document's first token in tok32
  int len;                   // ragged: number of stored tokens (rows len..T-1 are pad_id)
};
__device__ __forceinline__ int tiles_of(int Td);
__device__ __forceinline__ DocRef doc_ref(const Params& P, long long k) {
  DocRef d;
  work_item(P, k, d.doc, d.Td);
  d.npt = tiles_of(d.Td);
  d.base = 0;
  d.len = 0;
  if (P.tok32) {
    d.base = __ldg(P.off + d.doc);
    d.len = (int)(__ldg(P.off + d.doc + 1) - d.base);
  }
  return d;
}
__device__ __forceinline__ int tiles_of(int Td) { return (Td + 2 + 2 * TILE_M - 1) / (2 * TILE_M); }

// ------------------------------------------------------------------------------------------
// does (ov, op) beat (v, p)?  larger value, then smaller position
__device__ __forceinline__ bool beats(float ov, int op, float v, int p) { return ov > v || (ov == v && op < p); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

__device__ __forceinline__ float redux_max_f32(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ uint32_t redux_min_u32(uint32_t v) {
  uint32_t m;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(v));
  return m;
}


template <int EC>   // accumulator columns handled by one epilogue warp = Npad / 2
__device__ __forceinline__ void epilogue_role(const Params& P, SharedCtl* ctl, uint32_t rank, int cluster_id, int nclusters,
                                              int warp, int lane) {
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t leader_tmem_empty0 = mapa(smem_u32(&ctl->tmem_empty[0]), 0);
  uint32_t it = 0, ndoc = 0;
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_full = 0, w_bar = 0, w_xchg = 0, t_begin = clock64();
  for (long long k = cluster_id; k < P.N; k += nclusters, ++ndoc) {
    long long doc;
    int Td;
    work_item(P, k, doc, Td);
    const int npos = Td + 2;
    const int npt = tiles_of(Td);
    // running maximum per filter column and the tile it came from (one byte per column, packed
    // four to a register)
    float best[EC];
    uint32_t btile[EC / 4];
#pragma unroll
    for (int c = 0; c < EC; ++c) best[c] = -INFINITY;
#pragma unroll
    for (int c = 0; c < EC / 4; ++c) btile[c] = 0u;
    for (int pt = 0; pt < npt; ++pt, ++it) {
      const uint32_t buf = it % NACC, ph = (it / NACC) & 1u;
      TIMED_WAIT(w_full, mbar_wait(&ctl->tmem_full[buf], ph));
      tc_fence_after();
      const bool valid = (pt * 2 * TILE_M + (int)rank * TILE_M + row) < npos;
      const uint32_t taddr = ctl->tmem_base + lane_base + buf * ACC_STRIDE + h * EC;
      uint32_t tsh[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) tsh[k] = (uint32_t)pt << (8 * k);
#pragma unroll
      for (int c0 = 0; c0 < EC; c0 += 16) {
        uint32_t v[16];
        if (c0 + 16 <= EC) {
          tmem_ld16(taddr + c0, v);
        } else {
          uint32_t v8[8];
          tmem_ld8(taddr + c0, v8);
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = v8[c];
        }
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c0 + c < EC) {
              float x = __uint_as_float(v[c]);
              if (x > best[c0 + c]) {
                best[c0 + c] = x;
                btile[(c0 + c) >> 2] = (btile[(c0 + c) >> 2] & ~(0xffu << (8 * (c & 3)))) | tsh[c & 3];
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + buf * 8u);
    }
    // ---- per-document reduction over the warp's 32 rows: max value, smallest position on ties.
    // Two warp-wide reductions per column (redux.sync -> CREDUX): the maximum, then the smallest
    // key = tile << 5 | lane among the lanes that hold it (positions within one warp are ordered by
    // tile, then lane: same CTA rank and lane quarter).  Lane c % 32 keeps column c.
    {
      float keep_v[(EC + 31) / 32];
      uint32_t keep_k[(EC + 31) / 32];
#pragma unroll
      for (int c = 0; c < EC; ++c) {
        const float m = redux_max_f32(best[c]);
        const uint32_t tile = (btile[c >> 2] >> (8 * (c & 3))) & 0xffu;
        const uint32_t kmin = redux_min_u32(best[c] == m ? ((tile << 5) | (uint32_t)lane) : 0xffffffffu);
        if (lane == (c & 31)) { keep_v[c >> 5] = m; keep_k[c >> 5] = kmin; }
      }
#pragma unroll
      for (int i = 0; i < (EC + 31) / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < EC) {
          const float v = keep_v[i];
          ctl->red_val[warp][c] = v;
          ctl->red_pos[warp][c] = v == -INFINITY ? 0x7fffffff
                                                 : (int)(keep_k[i] >> 5) * 2 * TILE_M + (int)rank * TILE_M + q * 32 + (int)(keep_k[i] & 31u);
        }
      }
    }
    TIMED_WAIT(w_bar, asm volatile("bar.sync %0, 128;" :: "r"(1 + h) : "memory"));
    const bool col_thread = row < EC;
    float v = -INFINITY;
    int p = 0x7fffffff;
    if (col_thread) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        float ov = ctl->red_val[h * 4 + w][row];
        int op = ctl->red_pos[h * 4 + w][row];
        if (beats(ov, op, v, p)) { v = ov; p = op; }
      }
    }
    asm volatile("bar.sync %0, 128;" :: "r"(1 + h) : "memory");       // red_* may be overwritten by the next document
    if (col_thread) {
      // ---- merge the two CTAs' halves of the document through distributed shared memory
      const int f = h * EC + row;
      const uint32_t b = ndoc & 1u, use = ndoc >> 1;
      if (rank == 1) {
        TIMED_WAIT(w_xchg, mbar_wait(&ctl->xchg_empty[b], (use & 1u) ^ 1u));
        const uint32_t rbar = mapa(smem_u32(&ctl->xchg_full[b]), 0);
        st_async_u32(mapa(smem_u32(&ctl->xchg_val[b][f]), 0), __float_as_uint(v), rbar);
        st_async_u32(mapa(smem_u32(&ctl->xchg_pos[b][f]), 0), (uint32_t)p, rbar);
      } else {
        if (f == 0) mbar_arrive_expect_tx(&ctl->xchg_full[b], (uint32_t)P.Npad * 8u);
        TIMED_WAIT(w_xchg, mbar_wait(&ctl->xchg_full[b], use & 1u));
        const float ov = ctl->xchg_val[b][f];
        const int op = ctl->xchg_pos[b][f];
        if (beats(ov, op, v, p)) { v = ov; p = op; }
        // Releasing the buffer lets rank 1 overwrite xchg_*[b] (it may be two documents ahead), so the arrival
        // must not leave this SM before the two loads above have RETURNED.  A CTA-scope release does not order
        // this thread's shared-memory loads against another CTA's st.async (the compiler issued the remote
        // arrive while the loads were still in flight and, with the shared-memory pipe busy feeding the tensor
        // cores, rank 1's next stores occasionally won the race: round-1 nondeterminism).  A cluster-scope
        // release would cost MEMBAR.GPU per document; instead the arrival count is made data-dependent on the
        // merged position, which is computed from both loaded values (always 1: positions are non-negative).
        mbar_arrive_cluster_n(mapa(smem_u32(&ctl->xchg_empty[b]), 1), 1u + ((uint32_t)p >> 31));
        if (f < P.F) {
          const float o = v + __ldg(P.bias + f);
          P.pooled[doc * P.F + f] = o > 0.0f ? o : 0.0f;
          // positions Td, Td+1 of the shortened document are positions T, T+1 of the full one
          P.argmax[doc * P.F + f] = (p >= Td && p < npos) ? p + (P.T - Td) : p;
        }
      }
    }
  }
  if (prof_on && warp == 0 && lane == 0) {
    unsigned long long* o = P.prof + rank * 16;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_full; o[2] = w_bar; o[3] = w_xchg;
  }
}

template <int LAGT>
__device__ __forceinline__ void producer_role(const Params& P, SharedCtl* ctl, uint8_t* ring, uint32_t rank, int cluster_id,
                                              int nclusters, int ptid) {
  const int spt = (P.Kc + CPS - 1) / CPS;                // slabs per tile
  const int c8 = ptid & 7, r0 = ptid >> 3;
  const int lane = ptid & 31;
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t dst_thread = (uint32_t)(c8 * RA * 16 + r0 * 16);
  const int nslots = P.nslots;
  const uint32_t leader_full0 = mapa(smem_u32(&ctl->full[0]), 0);
  const bool last_row = r0 + 16 * (ROWS_PER_THREAD - 1) < TILE_M + 2;   // rows 128, 129 exist for r0 < 2 only
  // every producer lane copies ONE 16-byte chunk column (c8 + 8s in slab s) of its nine rows, so the
  // global address of a copy is a per-row base + a compile-time offset (no address arithmetic per copy);
  // conv padding rows read the all-zero row V of the shadow table (no src-size operand either)
  const uint8_t* const thread_base = P.shadow + c8 * 16;
  const uint8_t* const zero_row = thread_base + P.V * P.row_bytes;
  const int last_slab_chunks = P.Kc - (spt - 1) * CPS;                  // chunk columns in the last slab
  const bool in_last = c8 < last_slab_chunks;

  // Token ids of a tile's rows are fetched ONE TILE AHEAD into registers (unchecked, so the nine
  // loads are issued back to back and their HBM latency hides behind the current tile's slabs);
  // slot row r <-> document position pt*256 + rank*128 - 2 + r, -1 marks a zero (padding) row.
  auto fetch = [&](const DocRef& d, int pt, long long (&out)[ROWS_PER_THREAD]) {
    if (P.tok32 == nullptr) {
      const long long* drow = P.idx + d.doc * (long long)P.T;
#pragma unroll
      for (int k = 0; k < ROWS_PER_THREAD; ++k) {
        const int r = r0 + 16 * k;
        const int pos = pt * 2 * TILE_M + (int)rank * TILE_M - 2 + r;
        out[k] = (r < TILE_M + 2 && pos >= 0 && pos < d.Td) ? __ldg(drow + pos) : -1LL;
      }
    } else {
      const int* drow = P.tok32 + d.base;
#pragma unroll
      for (int k = 0; k < ROWS_PER_THREAD; ++k) {
        const int r = r0 + 16 * k;
        const int pos = pt * 2 * TILE_M + (int)rank * TILE_M - 2 + r;
        const bool in_doc = r < TILE_M + 2 && pos >= 0 && pos < d.Td;
        const int t = (in_doc && pos < d.len) ? __ldg(drow + pos) : (int)P.pad_id;
        out[k] = in_doc ? (long long)t : -1LL;
      }
    }
  };
  // publish a slab: its copies have landed (wait_group), make them visible to the tensor cores
  // (async proxy), then ONE arrival per warp on the leader CTA's barrier
  uint32_t sig_slot = 0;
  auto publish = [&]() {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(leader_full0 + sig_slot * 8u);
    sig_slot = sig_slot + 1 == (uint32_t)nslots ? 0u : sig_slot + 1;
  };

  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_empty = 0, w_group = 0, t_begin = clock64();
  uint32_t slot = 0, empty_parity = 1u;                  // parity the empty barrier shows once the slot is free
  uint32_t pending = 0;                                  // slabs issued but not yet published
  // `cur` = the work item being copied, `nxt` = the one after it: its order / length / offsets are loaded
  // a whole document ahead so that the token-id prefetch of its first tile never waits on them
  long long wk = cluster_id;
  int pt = 0;
  DocRef cur_d, nxt_d;
  cur_d.doc = nxt_d.doc = 0; cur_d.Td = nxt_d.Td = 0; cur_d.npt = nxt_d.npt = 1; cur_d.base = nxt_d.base = 0; cur_d.len = nxt_d.len = 0;
  long long cur[ROWS_PER_THREAD], nxt[ROWS_PER_THREAD];
  if (wk < P.N) {
    cur_d = doc_ref(P, wk);
    if (wk + nclusters < P.N) nxt_d = doc_ref(P, wk + nclusters);
    fetch(cur_d, 0, cur);
  }
  while (wk < P.N) {
    const bool wrap = pt + 1 == cur_d.npt;
    const long long nk = wrap ? wk + nclusters : wk;
    const int pt_next = wrap ? 0 : pt + 1;
    if (nk < P.N) fetch(wrap ? nxt_d : cur_d, pt_next, nxt);
    const uint8_t* src[ROWS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
      const long long tok = cur[k];
      if (tok < -1 || tok >= P.V) __trap();               // the reference device-asserts on OOB ids
      src[k] = tok < 0 ? zero_row : thread_base + tok * P.row_bytes;
    }
    static_for<0, MAX_SPT>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      if (s < spt) {
        TIMED_WAIT(w_empty, mbar_wait(&ctl->empty[slot], empty_parity));
        const uint32_t dst = ring_base + slot * (uint32_t)P.slot_bytes + dst_thread;
        if (s < spt - 1 || in_last) {
#pragma unroll
          for (int k = 0; k < ROWS_PER_THREAD - 1; ++k) cp_async16_ca_ofs<s * CPS * 16>(dst + k * 256, src[k]);
          if (last_row) cp_async16_ca_ofs<s * CPS * 16>(dst + (ROWS_PER_THREAD - 1) * 256, src[ROWS_PER_THREAD - 1]);
        }
        cp_async_commit();
        ++pending;
        if (pending > (uint32_t)LAGT) {
          TIMED_WAIT(w_group, cp_async_wait<LAGT>(); publish());
          --pending;
        }
        if (++slot == (uint32_t)nslots) { slot = 0; empty_parity ^= 1u; }
      }
    });
#pragma unroll
    for (int k2 = 0; k2 < ROWS_PER_THREAD; ++k2) cur[k2] = nxt[k2];
    if (wrap) {
      cur_d = nxt_d;
      if (nk + nclusters < P.N) nxt_d = doc_ref(P, nk + nclusters);   // consumed a document later
    }
    wk = nk;
    pt = pt_next;
  }
  cp_async_wait<0>();
  for (; pending > 0; --pending) publish();
  if (prof_on && ptid == 0) {
    unsigned long long* o = P.prof + rank * 16 + 4;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_empty; o[2] = w_group;
  }
}

__device__ __forceinline__ void mma_role(const Params& P, SharedCtl* ctl, const uint8_t* bsm, const uint8_t* ring, int cluster_id,
                                         int nclusters, int lane) {
  const int spt = (P.Kc + CPS - 1) / CPS;
  const uint32_t idesc = umma_idesc(P.fmt, P.Npad);
  const uint32_t a_base = smem_u32(ring), b_base = smem_u32(bsm);
  const uint32_t a_lbo = RA * 16, b_lbo = (uint32_t)(P.Npad / 2) * 16;
  const int nslots = P.nslots;
  // low descriptor word = start address >> 4 | LBO >> 4 << 16, high word = SBO (128 B) >> 4 | version 1
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);
  const uint32_t a_step = a_lbo >> 4, b_step = b_lbo >> 4;      // one 16-byte K-chunk
  const uint32_t a_j = 1u;                   // window row j: +16 bytes in the A slot
  const uint32_t b_j = (uint32_t)P.Kc * b_step;                 //               +Kc chunks in the filter bank
  const uint32_t a_lo0 = ((a_base >> 4) & 0x3FFFu) | (a_step << 16);
  const uint32_t b_lo0 = ((b_base >> 4) & 0x3FFFu) | (b_step << 16);
  const bool leader = elect_one();
  uint32_t slot = 0, full_parity = 0u, it = 0;
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_tmem = 0, w_full = 0, t_begin = clock64();
  for (long long k = cluster_id; k < P.N; k += nclusters) {
    long long doc;
    int Td;
    work_item(P, k, doc, Td);
    const int npt = tiles_of(Td);
    for (int pt = 0; pt < npt; ++pt, ++it) {
      const uint32_t buf = it % NACC, use = it / NACC;
      TIMED_WAIT(w_tmem, mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u));       // both CTAs' epilogues drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = ctl->tmem_base + buf * ACC_STRIDE;
      for (int s = 0; s < spt; ++s) {
        TIMED_WAIT(w_full, mbar_wait(&ctl->full[slot], full_parity));              // both CTAs' producers published this slab
        tc_fence_after();
        if (leader) {
          // descriptors advance by constants: +2 K-chunks per K=16 step, +1 row (16 B) per window row j
          const int nk = min(CPS, P.Kc - s * CPS) >> 1;        // K=16 steps in this slab
          uint32_t a_lo = a_lo0 + ((slot * (uint32_t)P.slot_bytes) >> 4);
          uint32_t b_lo = b_lo0 + (uint32_t)(s * CPS) * b_step;
          uint32_t acc = s ? 1u : 0u;
          for (int kk = 0; kk < nk; ++kk) {
            umma_f16_pair(d_tmem, mk_desc(a_lo, desc_hi), mk_desc(b_lo, desc_hi), idesc, acc);
            umma_f16_pair(d_tmem, mk_desc(a_lo + a_j, desc_hi), mk_desc(b_lo + b_j, desc_hi), idesc, 1u);
            umma_f16_pair(d_tmem, mk_desc(a_lo + 2 * a_j, desc_hi), mk_desc(b_lo + 2 * b_j, desc_hi), idesc, 1u);
            acc = 1u;
            a_lo += 2 * a_step;
            b_lo += 2 * b_step;
          }
          umma_commit_pair(&ctl->empty[slot]);                      // slot reusable (both CTAs) once these MMAs retire
          if (s == spt - 1) umma_commit_pair(&ctl->tmem_full[buf]); // accumulator complete (both CTAs)
        }
        __syncwarp();
        if (++slot == (uint32_t)nslots) { slot = 0; full_parity ^= 1u; }
      }
    }
  }
  if (prof_on && leader) {
    unsigned long long* o = P.prof + 8;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_tmem; o[2] = w_full;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) conv_pool_tc_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nh = P.Npad / 2;
  const uint32_t b_bytes = (uint32_t)(3 * P.Kc * nh * 16);

  // identical carve-up in both CTAs: the pair's MMA addresses both through one descriptor
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(smem);
  uint8_t* bsm = smem + ((sizeof(SharedCtl) + 127u) & ~127u);
  uint8_t* ring = bsm + ((b_bytes + 127u) & ~127u);

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(&ctl->full[i], 2 * NUM_PROD_WARPS); mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < NACC; ++i) {
      mbar_init(&ctl->tmem_full[i], 1);
      mbar_init(&ctl->tmem_empty[i], 2 * NUM_EPI_WARPS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->xchg_full[i], 1);
      mbar_init(&ctl->xchg_empty[i], (uint32_t)P.Npad);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&ctl->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // resident half of the filter bank: plain 16-byte copies of the pre-packed operand image
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.wpack + (size_t)rank * b_bytes);
    uint4* dst = reinterpret_cast<uint4*>(bsm);
    for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += NUM_THREADS) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();                      // generic-proxy smem writes -> visible to the tensor cores (async proxy)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // barriers initialised and both filter halves resident before any remote arrive / MMA
  tc_fence_after();

  if (warp < NUM_EPI_WARPS) {
    switch (nh) {
      case 8:  epilogue_role<8>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 16: epilogue_role<16>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 24: epilogue_role<24>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 32: epilogue_role<32>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 40: epilogue_role<40>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 48: epilogue_role<48>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      case 56: epilogue_role<56>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
      default: epilogue_role<64>(P, ctl, rank, cluster_id, nclusters, warp, lane); break;
    }
  } else if (warp < MMA_WARP) {
    const int ptid = threadIdx.x - NUM_EPI_WARPS * 32;
    switch (P.lag) {
      case 1:  producer_role<1>(P, ctl, ring, rank, cluster_id, nclusters, ptid); break;
      case 3:  producer_role<3>(P, ctl, ring, rank, cluster_id, nclusters, ptid); break;
      case 4:  producer_role<4>(P, ctl, ring, rank, cluster_id, nclusters, ptid); break;
      default: producer_role<2>(P, ctl, ring, rank, cluster_id, nclusters, ptid); break;
    }
  } else if (rank == 0) {
    mma_role(P, ctl, bsm, ring, cluster_id, nclusters, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still read this CTA's smem / TMEM through the pair MMA
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(ctl->tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- weight packing
struct PackPlan {
  int Kc, Npad, nh;
  long long half_bytes, total;
};

inline bool make_plan(int E, int F, PackPlan& pl) {
  if (E <= 0 || F <= 0 || F > N_MAX) return false;
  pl.Kc = ((E + 15) / 16) * 2;
  pl.Npad = ((F + 15) / 16) * 16;
  pl.nh = pl.Npad / 2;
  pl.half_bytes = (long long)3 * pl.Kc * pl.nh * 16;
  pl.total = 2 * pl.half_bytes;
  return true;
}

template <typename T> __device__ __forceinline__ T cvt_w(float f);
template <> __device__ __forceinline__ __half cvt_w<__half>(float f) { return __float2half_rn(f); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_w<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

// image[rank][(j*Kc + ch)][r][e8] = W[rank*nh + r][j][ch*8+e8]
template <typename T>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int E, int F, T* __restrict__ out, PackPlan pl) {
  const long long total = pl.total / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int h = i * 2 >= pl.half_bytes ? 1 : 0;
    long long q = i - h * (pl.half_bytes / 2);
    int e8 = (int)(q & 7); q >>= 3;
    int r = (int)(q % pl.nh); q /= pl.nh;
    int ch = (int)(q % pl.Kc);
    int j = (int)(q / pl.Kc);
    int f = h * pl.nh + r, e = ch * 8 + e8;
    float v = (f < F && e < E) ? w[((long long)f * 3 + j) * E + e] : 0.0f;
    out[i] = cvt_w<T>(v);
  }
}
}  // namespace

static int g_clusters = 0;
// Persistent CTA pairs the next r4r_conv_pool_tc launches use (0 = one per SM pair).  The kernel owns every SM it
// runs on (one CTA with all of the shared memory), so kernels of a concurrent stream / graph branch -- the sharded
// word lookup of the next step and its NCCL all-to-alls (train.CapturedStep) -- only overlap if some pairs stay free.
extern "C" int r4r_conv_set_clusters(int n) {
  g_clusters = n > 0 ? n : 0;
  return 0;
}

static unsigned long long* g_prof = nullptr;
// Diagnostics: 32 x uint64 device buffer receiving the per-role cycle counters of cluster 0
// ([rank*16+0..3] epilogue total / wait tmem_full / bar.sync / exchange, [rank*16+4..6] producer
// total / wait empty / wait copies, [8..10] MMA total / wait tmem_empty / wait full); NULL = off.
extern "C" int r4r_conv_debug_profile(void* buf32_u64) {
  g_prof = static_cast<unsigned long long*>(buf32_u64);
  return 0;
}

extern "C" int64_t r4r_conv_wpack_bytes(int E, int F) {
  PackPlan pl;
  if (!make_plan(E, F, pl)) return -1;
  return pl.total;
}

extern "C" int r4r_conv_pack_weights(const float* conv_w, int E, int F, void* wpack, int dtype, void* stream) {
  R4R_REQUIRE(conv_w && wpack, R4R_EINVAL, "conv_pack_weights: null pointer");
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_pack_weights: dtype %d", dtype);
  PackPlan pl;
  R4R_REQUIRE(make_plan(E, F, pl), R4R_EUNSUP, "conv_pack_weights: E=%d F=%d unsupported (F <= %d)", E, F, N_MAX);
  long long n = pl.total / 2;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (dtype == R4R_DT_F16) pack_weights_kernel<__half><<<blocks, 256, 0, as_stream(stream)>>>(conv_w, E, F, (__half*)wpack, pl);
  else pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(conv_w, E, F, (__nv_bfloat16*)wpack, pl);
  R4R_CHECK_LAUNCH("conv_pack_weights");
  return 0;
}

static int conv_pool_tc_launch(const void* shadow, int64_t V, int Epad, int E, int dtype,
                               const int64_t* idx, const int32_t* tok32, const int64_t* off, int64_t pad_id, int64_t N, int T,
                               const void* wpack, const float* conv_b, int F,
                               float* pooled, int32_t* argmax,
                               const int32_t* doc_len, const int32_t* doc_order, void* stream) {
  R4R_REQUIRE(shadow && (idx || (tok32 && off)) && wpack && conv_b && pooled && argmax, R4R_EINVAL, "conv_pool_tc: null pointer");
  R4R_REQUIRE(idx || (pad_id >= 0 && pad_id < V), R4R_EINVAL, "conv_pool_tc: pad id %lld outside the table", (long long)pad_id);
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0, R4R_EINVAL, "conv_pool_tc: bad sizes");
  R4R_REQUIRE((T + 2 + 2 * TILE_M - 1) / (2 * TILE_M) <= 256, R4R_EUNSUP, "conv_pool_tc: T=%d exceeds 256 position tiles", T);
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_pool_tc: dtype %d", dtype);
  PackPlan pl;
  R4R_REQUIRE(make_plan(E, F, pl), R4R_EUNSUP, "conv_pool_tc: E=%d F=%d unsupported (F <= %d)", E, F, N_MAX);
  R4R_REQUIRE(Epad % 8 == 0 && Epad >= pl.Kc * 8, R4R_EINVAL, "conv_pool_tc: shadow row width Epad=%d must be a multiple of 8 and >= %d", Epad, pl.Kc * 8);
  // NOTE: the shadow table must carry V+1 rows, row V all zero (r4r_shadow_build writes it): conv padding rows are read from it
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
  // shared-memory plan: [ctl][resident filter half][A ring of nslots x (CPS chunks x RA rows x 16 B)]
  const long long ctl_bytes = (sizeof(SharedCtl) + 127) & ~127LL;
  const long long b_bytes = (pl.half_bytes + 127) & ~127LL;
  const long long avail = (long long)smem_optin - ctl_bytes - b_bytes - 1024;
  R4R_REQUIRE(pl.Kc <= CPS * MAX_SPT, R4R_EUNSUP, "conv_pool_tc: E=%d exceeds %d", E, CPS * MAX_SPT * 8);
  long long ns = avail / ((long long)CPS * RA * 16);
  if (ns > MAX_SLOTS) ns = MAX_SLOTS;
  const int nslots = (int)ns;
  int lag = LAG;
  {
    const char* e = getenv("R4R_CONV_LAG");                 // tuning override
    if (e && atoi(e) >= 1 && atoi(e) <= MAX_LAG) lag = atoi(e);
    if (lag > nslots - 2) lag = nslots - 2;
  }
  R4R_REQUIRE(nslots >= 3 && lag >= 1, R4R_EUNSUP, "conv_pool_tc: E=%d F=%d leaves no room for the A ring next to the filter bank", E, F);
  const int slot_bytes = CPS * RA * 16;
  const size_t smem_bytes = (size_t)(ctl_bytes + b_bytes + (long long)nslots * slot_bytes);

  Params P;
  P.shadow = static_cast<const uint8_t*>(shadow);
  P.row_bytes = (long long)Epad * 2;
  P.V = V;
  P.idx = reinterpret_cast<const long long*>(idx);
  P.tok32 = idx ? nullptr : tok32;
  P.off = reinterpret_cast<const long long*>(off);
  P.pad_id = pad_id;
  P.N = N; P.T = T; P.Kc = pl.Kc; P.F = F; P.Npad = pl.Npad;
  P.wpack = static_cast<const uint8_t*>(wpack);
  P.bias = conv_b; P.pooled = pooled; P.argmax = argmax;
  P.fmt = dtype; P.nslots = nslots; P.slot_bytes = slot_bytes;
  P.prof = g_prof;
  P.doc_len = doc_len;
  P.doc_order = doc_order;
  P.lag = lag;

  R4R_CUDA(cudaFuncSetAttribute(conv_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  long long nclusters = sm_count / 2;
  {
    int want = g_clusters;                                  // r4r_conv_set_clusters: leave SM pairs free for a concurrent graph branch
    const char* e = getenv("R4R_CONV_CLUSTERS");            // tuning override
    if (e && atoi(e) >= 1) want = atoi(e);
    if (want >= 1 && want < nclusters) nclusters = want;
  }
  if (nclusters > N) nclusters = N;
  conv_pool_tc_kernel<<<(unsigned)(2 * nclusters), NUM_THREADS, smem_bytes, as_stream(stream)>>>(P);
  R4R_CHECK_LAUNCH("conv_pool_tc");
  return 0;
}

extern "C" int r4r_conv_pool_tc(const void* shadow, int64_t V, int Epad, int E, int dtype,
                                const int64_t* idx, int64_t N, int T,
                                const void* wpack, const float* conv_b, int F,
                                float* pooled, int32_t* argmax,
                                const int32_t* doc_len, const int32_t* doc_order, void* stream) {
  R4R_REQUIRE(idx, R4R_EINVAL, "conv_pool_tc: null pointer");
  return conv_pool_tc_launch(shadow, V, Epad, E, dtype, idx, nullptr, nullptr, 0, N, T, wpack, conv_b, F, pooled, argmax,
                             doc_len, doc_order, stream);
}

extern "C" int r4r_conv_pool_tc_ragged(const void* shadow, int64_t V, int Epad, int E, int dtype,
                                       const int32_t* tokens, const int64_t* offsets, int64_t pad_id, int64_t N, int T,
                                       const void* wpack, const float* conv_b, int F,
                                       float* pooled, int32_t* argmax,
                                       const int32_t* doc_len, const int32_t* doc_order, void* stream) {
  R4R_REQUIRE(tokens && offsets, R4R_EINVAL, "conv_pool_tc_ragged: null pointer");
  return conv_pool_tc_launch(shadow, V, Epad, E, dtype, nullptr, tokens, offsets, pad_id, N, T, wpack, conv_b, F, pooled, argmax,
                             doc_len, doc_order, stream);
}
