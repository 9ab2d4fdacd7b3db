// conv_tc.cu -- K2 (fast mode): fused word gather -> TextCNN conv on 5th-gen tensor cores
// (tcgen05.mma cta_group::2 kind::f16, fp32 accumulators in TMEM) -> ReLU -> max/argmax over
// positions.  Replaces nn.Embedding + F.conv2d(pad=(2,0)) + F.relu + F.max_pool1d
// (DeepCoNN.py:53-54, common_pytorch_models.py:26-31).
//
// GEMM view per document: Y[p, f] = sum_{j<3} X[p+j-2, :] . W[f, j, :]   (M = T+2 positions = windows,
// N = filters, K = 3*E).  The three window rows are the SAME gathered rows shifted by one position,
// so the A operand is staged ONCE per 128-position tile and the j-th GEMM reads it through a
// shared-memory descriptor whose start address is advanced by j rows.
//
// TMA staging.  The rows are gathered by the tensor memory accelerator: one
// cp.async.bulk.tensor.2d ... tile::gather4 instruction fetches 64 columns (128 B) of FOUR arbitrary rows
// of the shadow table and writes them as four consecutive 128-byte rows of the canonical K-major
// SWIZZLE_128B layout.  A ring slab = one 64-column block of a tile's 132 rows (128 + 2 halo + 2 spare);
// row r lives at slab + r*128 with its 16-byte chunks XOR-ed by (r & 7).  The swizzle is a function of the
// ABSOLUTE shared-memory address (slabs are 1024-byte aligned), so window row j is simply the descriptor
// start address + j*128 with base_offset 0 (scripts/experiments/gather4_shift_test.cu: 2048/2048 outputs
// exact for j = 0, 1, 2; base_offset = j is wrong), and a K=16 step is +32 B inside the 128-byte row.
// Measured (scripts/experiments/gather4_bw.cu, profiles/r2_v3_gather4_staging_microbench.log; Zipf(1.0) tokens, 148
// CTAs staging only): one issuing thread 0.81 ms per 1.65 M rows, 4 warps x 8 issuing lanes 0.150 ms = 7.1 TB/s; a warp's
// copies issue lane after lane (~58 cycles each) and one SM's TMA unit retires a gather4 per ~20 cycles = 25 B/clk -- the
// hot head of the Zipf distribution costs nothing (uniform ids: the same 0.150 ms).  Inside this kernel that rate is
// BELOW the ~31 B/clk the tensor cores consume, so the TMA stages only the first slabs of every tile and a cp.async team
// writes the rest into the same swizzled layout (see the warp roles); TMA alone: 0.451 ms per 4096 documents with seven
// issuing warps, cp.async alone: 0.413 ms, the two together: 0.371 ms (all three before the window streams and the
// 16x256b epilogue; the shipped kernel: 0.25-0.27 ms, sweep of the TMA share in scripts/experiments/README.md).
//
// CTA pair.  The filter bank W (B operand, 3*E x 100 fp16 = 180 KB) must stay resident in shared
// memory next to the A ring, which one SM cannot hold, and a single-SM MMA of N <= 64 filters is
// bound by shared-memory operand bandwidth (measured 57 instead of 32 cycles per MMA).  The kernel
// therefore runs as clusters of two CTAs on one TPC issuing tcgen05.mma.cta_group::2 with
// M = 256 positions x N = all filters: each CTA stages its own 128 positions of A and keeps HALF of
// the filter bank (no-swizzle interleaved layout); the pair's tensor cores read both halves.  Every
// document row is gathered exactly once.
//
// Warp roles per CTA (512 threads x 128 registers, 1 CTA/SM, persistent over documents):
//   warps 0-7   epilogue : tcgen05.ld.16x256b of this CTA's 128 x N accumulator (warp w: TMEM lane quarter
//                          w&3, column half w>>2; a thread sees four of the warp's 32 rows and a column lives in
//                          8 lanes) -> running max / key (tile<<5|row) in registers over the windows a document
//                          owns -> per document three xor-shuffle steps for the max and three for the smallest
//                          key among its holders (= FIRST maximum), all values of a step issued back to back +
//                          a four-warp shared-memory merge -> the two CTAs' partial (max, argmax)
//                          are merged through distributed shared memory (rank 1 stores into rank 0
//                          with st.async) -> bias, ReLU, store
//   warps 8-11  cp.async : slabs tma_slabs.. of every tile: 16-byte gathers, each lane one chunk column of nine rows,
//                          written at the swizzled address; asynchronous hand-off (cp.async.mbarrier.arrive.noinc:
//                          no wait_group, no MEMBAR in the producer)
//   warps 12-14 TMA      : slabs 0..tma_slabs-1 of every tile (2 of 5 at E = 300): eleven lanes per warp issue the
//                          gather4 copies (a lane's four table rows stay in registers for the tile's slabs); both CTAs'
//                          copies complete on the LEADER's barrier (.cta_group::2); conv padding rows fetch the all-zero
//                          row V of the shadow table.  Both teams fetch the token ids of the next tile (contiguous
//                          int32 rows of the pair's window stream) while the current one is staged.
//   warp  15    MMA      : allocates all 512 TMEM columns (both CTAs) = four accumulator buffers; in
//                          the leader CTA one elected lane issues tcgen05.mma, multicast
//                          tcgen05.commit releases ring slots ("empty") and publishes accumulators
//                          ("tmem_full") in both CTAs; in rank 1 one lane relays "my cp.async slab has landed" to
//                          the leader, so the MMA warp waits on exactly one barrier per slab
//
// Work plan (docplan.cu): documents end in a run of one repeated padding token; every conv window
// inside the run repeats a value max-pooling has already seen, so document n is processed as if it had
// doc_len[n] = min(T, run start + 3) rows and arg-max positions >= doc_len are mapped back by
// + (T - doc_len) -- bit-identical results, ~2.5x less work on Amazon-shaped batches.  The launch then deals the
// documents (longest first, alternating direction) to its CTA pairs and lays each pair's documents end to end as one
// WINDOW STREAM (see Params): tiles are 256 consecutive windows of the stream, not whole tiles per document
// (2.08 -> 1.59 tiles per Amazon-shaped document).  Ragged input (tokens + offsets) takes the same path.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int TILE_M = 128;            // positions per CTA per accumulator tile (UMMA M = 256 per pair)
constexpr int RS = 132;                // rows per A slab: 130 needed (128 + 2 halo) = 33 gather4 groups of 4
constexpr int SLAB_BYTES = 17 * 1024;  // RS * 128 B rounded up to the 1024-byte swizzle period
constexpr int NGROUPS = RS / 4;        // gather4 copies (row groups) per slab
static_assert(RS * 128 <= SLAB_BYTES && RS % 4 == 0 && RS >= 130, "slab holds the tile's 128 rows + 2 halo rows in whole gather4 groups");
constexpr int N_MAX = 128;             // filters (UMMA N), multiple of 16
constexpr int ACC_STRIDE = 128;        // TMEM columns between consecutive accumulator buffers
#ifndef R4R_NACC
#define R4R_NACC 4
#endif
constexpr int NACC = R4R_NACC;         // accumulator buffers: the MMA warp may run three tiles ahead of the epilogue,
                                       // which hides the per-document reduction (the epilogue drains nothing meanwhile)
constexpr int TMEM_COLS = NACC * ACC_STRIDE;   // 512 = all of TMEM (1 CTA per SM)
// Two producer teams share the slabs of a tile (8 epilogue + 4 + 3 producer + 1 MMA warp = 512 threads x 128 registers):
//   TMA team   (3 warps): the first `tma_slabs` slabs of every tile (2 of 5 at E = 300).  A warp's TMA copies issue lane
//                         after lane (~58 cycles each) and one SM's TMA unit retires a gather4 per ~20 cycles = 25 B/clk,
//                         below the 31 B/clk the tensor cores consume -- TMA alone leaves the kernel copy-bound
//                         (measured: 0.451 ms per 4096 documents with seven issuing warps, 0.529 ms with four).
//   cp.async team (4 warps): the remaining slabs, written into the SAME swizzled layout (chunk c of row r at
//                         r*128 + ((c ^ (r & 7)) << 4)); alone it is LSU-issue-bound (0.343 ms).
// Together each team runs at about half of its ceiling and the kernel becomes tensor-pipe-bound.
#ifndef R4R_LDG_WARPS
#define R4R_LDG_WARPS 4
#endif
constexpr int NUM_EPI_WARPS = 8, NUM_LDG_WARPS = R4R_LDG_WARPS, NUM_TMA_WARPS = 7 - NUM_LDG_WARPS, NUM_PROD_WARPS = 7;
constexpr int LDG_ROW_STEP = NUM_LDG_WARPS * 4;                                     // rows covered by one pass of the cp.async team
constexpr int ROWS_PER_THREAD = (130 + LDG_ROW_STEP - 1) / LDG_ROW_STEP;          // cp.async thread i copies rows i/8 + LDG_ROW_STEP*k of ONE 16-byte chunk column
static_assert(LDG_ROW_STEP % 8 == 0, "a thread's rows must share r & 7 (constant swizzled chunk)");
constexpr int TMA_GROUPS_PER_LANE = (RS / 4 + NUM_TMA_WARPS * 32 - 1) / (NUM_TMA_WARPS * 32);   // 1 unless a single warp stages all 33 groups
constexpr int MMA_WARP = NUM_EPI_WARPS + NUM_PROD_WARPS;
constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
constexpr int MAX_SLOTS = 8;
constexpr int CPS = 8;                 // 16-byte K-chunks per ring slab (K = 64 = one 128-byte swizzled row per slab)
constexpr int MAX_SPT = 16;            // slabs per position tile -> Kc <= 128 chunks (E <= 1024)

struct SharedCtl {
  // "this CTA's slab has landed", one barrier per ring slot and producer team (a slot is used by either team, depending on
  // the slab).  In the LEADER the same barriers also take one arrival from rank 1 (relayed by its idle MMA warp once ITS
  // slab has landed), so the MMA warp waits on exactly one barrier per slab.
  unsigned long long landed_t[MAX_SLOTS];  // TMA teams of BOTH CTAs -> the leader's copy: 2 x NUM_TMA_WARPS arrive.expect_tx + the bytes of both halves
  unsigned long long landed_l[MAX_SLOTS];  // cp.async team: one asynchronous arrival per thread (cp.async.mbarrier.arrive.noinc) (+ 1 relay in the leader)
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
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" :: "r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Every wait is bounded: a protocol bug must surface as a trapped launch (an error at the caller's next
// synchronisation), never as a kernel that spins forever.  try_wait suspends the thread for a hardware-defined
// interval per poll, so 2^26 failed polls are many seconds -- orders of magnitude beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  uint32_t polls = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && ++polls == (1u << 26)) __trap();
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

// 16-byte asynchronous copies of the cp.async team (.ca: duplicate requests for a hot row merge in L1)
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// Arrival on `bar` that the hardware performs when all cp.async copies this thread has issued so far have landed:
// the producer never waits for its own copies (no wait_group, no MEMBAR) -- the hand-off is fully asynchronous.
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// One TMA gather4 copy: columns [col0, col0 + 64) of table rows r0..r3 -> four consecutive 128-byte rows at `dst`
// (SWIZZLE_128B, this CTA's shared memory), 512 transaction bytes on the barrier `bar` -- a shared::cluster address:
// with .cta_group::2 the completion may signal the PEER's barrier, so both CTAs' copies complete on the leader's.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int col0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(dst), "l"(tmap), "r"(bar), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// UMMA shared-memory descriptor (K-major, version 1 for sm_100): bits 0-13 start address >> 4, 16-29 LBO >> 4,
// 32-45 SBO >> 4, bit 46 version, 61-63 layout type (0 = no swizzle: the filter bank; 2 = SWIZZLE_128B: the
// TMA-written A slabs).  mma_role() keeps the two 32-bit halves and advances the address field by constants.
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
  int tma_slabs;             // slabs 0 .. tma_slabs-1 of every tile are staged by TMA, the rest by cp.async
  unsigned long long* prof;  // diagnostics: per-role cycle counters of cluster 0 (r4r_conv_debug_profile), or NULL
  // the launch's window streams (stream_plan_kernel / stream_fill_kernel below), one per CTA pair
  const int* stream;         // row u of cluster c at stream[c * stream_stride + u]: token id, or -1 = all-zero (conv padding) row
  long long stream_stride;   // ints per cluster (a multiple of 4)
  const int4* dlist;         // document i of cluster c at dlist[c * dlist_stride + i] = (document, effective length, first window, 0)
  long long dlist_stride;
  const int2* head;          // per cluster: (tiles of 256 windows, documents)
};

// The WINDOW STREAM of a CTA pair.  Its documents are laid end to end, each as two zero rows followed by its
// (effective) rows; the two zero rows that open document i+1 also close document i:
//     rows     z z a0 a1 .. aL-1 z z b0 b1 .. bM-1 z z ...
// Window w covers rows w, w+1, w+2.  Document a (first row index o) owns windows o .. o+L+1 = its L+2 conv positions
// (position p <-> window o + p), document b starts at window o + L + 2: every window belongs to exactly one document
// and none straddles two documents' rows.  The pair's tensor cores therefore run over tiles of 256 CONSECUTIVE windows
// of the stream -- a tile holds the end of one document and the start of the next -- instead of rounding every
// document up to whole tiles (Amazon-shaped batches: 2.08 -> 1.59 tiles per document).  The epilogue takes the
// max / arg-max per document over the window range it owns.
// ------------------------------------------------------------------------------------------
// does (ov, op) beat (v, p)?  larger value, then smaller position
__device__ __forceinline__ bool beats(float ov, int op, float v, int p) { return ov > v || (ov == v && op < p); }

// tcgen05.ld .16x256b: 16 TMEM lanes x 8 columns per block.  Register 4*blk + 2*j + e of lane L holds
// (TMEM lane  base + L/4 + 8*j,  column  8*blk + 2*(L%4) + e)  -- pinned by scripts/experiments/tmem_ld_layout.cu
// (profiles/r2_v6_tmem_ld_layout.log).  A thread therefore sees FOUR rows of its warp's 32 (two per half of 16 lanes)
// and a column is spread over only 8 lanes: the per-document cross-lane reduction is three shuffle steps on EC/4 values
// per thread instead of two warp-wide reductions on each of EC columns.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

template <int EC>   // accumulator columns handled by one epilogue warp = Npad / 2
__device__ __forceinline__ void epilogue_role(const Params& P, SharedCtl* ctl, uint32_t rank, int cluster_id, int warp, int lane) {
  constexpr int NB = EC / 8;                              // 8-column blocks of this warp's accumulator half
  const int q = warp & 3, h = warp >> 2;
  const int row = q * 32 + lane;
  const int rsub = lane >> 2, csub = (lane & 3) * 2;      // .16x256b: this thread's rows are rsub + 8*ri, its columns 8*blk + csub + e
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t leader_tmem_empty0 = mapa(smem_u32(&ctl->tmem_empty[0]), 0);
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_full = 0, w_bar = 0, w_xchg = 0, t_begin = clock64();
  long long c_pass = 0, c_fin = 0, c_merge = 0, c_mark = 0;   // diagnostics: cycles in the tile passes / reduction / merge + exchange
  const int nd = __ldg(P.head + cluster_id).y;
  const int4* dl = P.dlist + (long long)cluster_id * P.dlist_stride;
  const int wrow0 = (int)rank * TILE_M + q * 32;          // first window of this warp inside a tile
  int acquired = -1;                                      // last stream tile whose accumulator this warp has waited for
  int4 e_next = nd > 0 ? __ldg(dl) : make_int4(0, 0, 0, 0);
  for (int i = 0; i < nd; ++i) {
    const int4 e = e_next;                                // (document, effective length, first window)
    if (i + 1 < nd) e_next = __ldg(dl + i + 1);
    const long long doc = e.x;
    const int Td = e.y, o = e.z;
    const int npos = Td + 2;
    const int w_end = o + npos;                           // one past the document's last window
    const int t0 = o >> 8, t1 = (w_end - 1) >> 8;         // stream tiles holding its windows
    // running maximum of this thread's four rows per column it sees, and where it came from (key = tile << 5 | row)
    float best[NB][2];
    uint32_t bkey[NB][2];
#pragma unroll
    for (int b = 0; b < NB; ++b) { best[b][0] = best[b][1] = -INFINITY; bkey[b][0] = bkey[b][1] = 0u; }
    bool touched = false;                                 // warp-uniform: some row of this warp belongs to the document
    for (int t = t0; t <= t1; ++t) {
      const uint32_t buf = (uint32_t)t % NACC, ph = ((uint32_t)t / NACC) & 1u;
      if (t > acquired) {
        // every warp waits for every tile, also one it has no rows in: its release below must not run ahead of the
        // tile's MMAs (the arrival counts of tmem_empty assume one arrival per warp per use of the buffer)
        TIMED_WAIT(w_full, mbar_wait(&ctl->tmem_full[buf], ph));
        tc_fence_after();
        acquired = t;
      }
      const int wbase = t * 2 * TILE_M + wrow0;           // window of lane 0
      if (wbase < w_end && wbase + 32 > o) {              // warp-uniform: the tcgen05.ld below are warp-collective
        if (prof_on) c_mark = clock64();
        touched = true;
        const int wrel = wbase - o;                       // window of the warp's row 0, relative to the document's first
        bool valid[4];
        uint32_t kc[4];
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
          const int r = rsub + 8 * ri;
          valid[ri] = (unsigned)(wrel + r) < (unsigned)npos;
          kc[ri] = ((uint32_t)(t - t0) << 5) | (uint32_t)r;
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {            // rows in increasing order: strict > keeps the first maximum
          const uint32_t taddr = ctl->tmem_base + lane_base + ((uint32_t)(half * 16) << 16) + buf * ACC_STRIDE + h * EC;
          uint32_t v[NB * 4];                             // all loads of the half are in flight before the one wait
#pragma unroll
          for (int b0 = 0; b0 < NB; b0 += 4) {
            const int nb = NB - b0 < 4 ? NB - b0 : 4;
            if (nb == 4) {
              tmem_ld_16x256b_x4(taddr + b0 * 8, v + 4 * b0);
            } else {
              if (nb >= 2) tmem_ld_16x256b_x2(taddr + b0 * 8, v + 4 * b0);
              if (nb & 1) tmem_ld_16x256b_x1(taddr + (b0 + nb - 1) * 8, v + 4 * (b0 + nb - 1));
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                const float x = __uint_as_float(v[4 * b + 2 * j + e2]);
                if (valid[half * 2 + j] && x > best[b][e2]) { best[b][e2] = x; bkey[b][e2] = kc[half * 2 + j]; }
              }
        }
        if (prof_on) c_pass += clock64() - c_mark;
      }
      // the accumulator is released by the document that reaches the end of the tile (later documents start in later
      // tiles), or by the pair's last document
      if (w_end >= (t + 1) * 2 * TILE_M || i == nd - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + buf * 8u);
      }
    }
    // ---- per-document reduction over the warp's 32 rows: max value, smallest position on ties.  A column lives in the 8
    // lanes with the same lane % 4: three xor-shuffle steps for the maximum, three for the smallest key = tile << 5 | row
    // among its holders (positions within one warp are ordered by tile, then row) = FIRST maximum, as F.max_pool1d.
    // (Measured alternatives on the .32x32b layout: two redux.sync per column 0.334 ms per launch, one redux + a
    // shared-memory atomicMin by the holders 0.417 ms.)
    if (prof_on) c_mark = clock64();
    if (touched) {
      // The shuffles of one step are issued back to back for all of the thread's values: shuffles keep program order, so a
      // value-by-value formulation is one dependent chain of 6 * EC/4 shuffle latencies (measured: 4.1 k cycles per document).
      float m[NB][2];
      uint32_t k[NB][2];
#pragma unroll
      for (int b = 0; b < NB; ++b) { m[b][0] = best[b][0]; m[b][1] = best[b][1]; }
#pragma unroll
      for (int step = 4; step <= 16; step <<= 1) {
        float ov[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b) { ov[b][0] = __shfl_xor_sync(0xffffffffu, m[b][0], step); ov[b][1] = __shfl_xor_sync(0xffffffffu, m[b][1], step); }
#pragma unroll
        for (int b = 0; b < NB; ++b) { m[b][0] = fmaxf(m[b][0], ov[b][0]); m[b][1] = fmaxf(m[b][1], ov[b][1]); }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        k[b][0] = best[b][0] == m[b][0] ? bkey[b][0] : 0xffffffffu;
        k[b][1] = best[b][1] == m[b][1] ? bkey[b][1] : 0xffffffffu;
      }
#pragma unroll
      for (int step = 4; step <= 16; step <<= 1) {
        uint32_t ok[NB][2];
#pragma unroll
        for (int b = 0; b < NB; ++b) { ok[b][0] = __shfl_xor_sync(0xffffffffu, k[b][0], step); ok[b][1] = __shfl_xor_sync(0xffffffffu, k[b][1], step); }
#pragma unroll
        for (int b = 0; b < NB; ++b) { k[b][0] = min(k[b][0], ok[b][0]); k[b][1] = min(k[b][1], ok[b][1]); }
      }
      if (rsub == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int c = 8 * b + csub + e2;
            // window -> position of the document: (t0 + tile) * 256 + rank * 128 + q * 32 + row - o
            ctl->red_val[warp][c] = m[b][e2];
            ctl->red_pos[warp][c] = m[b][e2] == -INFINITY ? 0x7fffffff
                                                          : (t0 + (int)(k[b][e2] >> 5)) * 2 * TILE_M + wrow0 + (int)(k[b][e2] & 31u) - o;
          }
      }
    } else {
#pragma unroll
      for (int i2 = 0; i2 < (EC + 31) / 32; ++i2) {
        const int c = lane + 32 * i2;
        if (c < EC) { ctl->red_val[warp][c] = -INFINITY; ctl->red_pos[warp][c] = 0x7fffffff; }
      }
    }
    if (prof_on) { const long long now = clock64(); c_fin += now - c_mark; c_mark = now; }
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
      const uint32_t b = (uint32_t)i & 1u, use = (uint32_t)i >> 1;
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
          const float ob = v + __ldg(P.bias + f);
          P.pooled[doc * P.F + f] = ob > 0.0f ? ob : 0.0f;
          // positions Td, Td+1 of the shortened document are positions T, T+1 of the full one
          P.argmax[doc * P.F + f] = (p >= Td && p < npos) ? p + (P.T - Td) : p;
        }
      }
    }
    if (prof_on) c_merge += clock64() - c_mark;
  }
  if (prof_on && warp == 0 && lane == 0) {
    unsigned long long* o = P.prof + rank * 16;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_full; o[2] = w_bar; o[3] = w_xchg;
    if (rank == 0) { P.prof[24] = c_pass; P.prof[25] = c_fin; P.prof[26] = c_merge; P.prof[27] = (unsigned long long)nd; }   // slots unused by rank 1's teams
  }
}

// ---- TMA team: slabs 0 .. tma_slabs-1 of every tile.  Row group g (rows 4g .. 4g+3) of a slab belongs to lane
// g / NUM_TMA_WARPS of TMA warp g % NUM_TMA_WARPS; the four table rows stay in registers for the tile's slabs.
// Slab row r of tile t of this CTA is stream row t*256 + rank*128 + r: a lane's four tokens are one aligned 16-byte load.
__device__ __forceinline__ void tma_role(const Params& P, const CUtensorMap* tmap, SharedCtl* ctl, uint8_t* ring, uint32_t rank,
                                         int cluster_id, int pwarp, int lane) {
  // row group g of a slab (rows 4g .. 4g+3): warp g % NUM_TMA_WARPS, lane (g / NUM_TMA_WARPS) % 32, the lane's j-th group
  constexpr int GPL = TMA_GROUPS_PER_LANE;
  const int warp_groups = (NGROUPS - pwarp + NUM_TMA_WARPS - 1) / NUM_TMA_WARPS;     // groups of this warp
  int grp[GPL], my_groups = 0;
#pragma unroll
  for (int j = 0; j < GPL; ++j) {
    grp[j] = (lane + 32 * j) * NUM_TMA_WARPS + pwarp;
    if (lane + 32 * j < warp_groups) ++my_groups;
  }
  if (my_groups == 0) return;                            // TMA copies are per-thread instructions: the other lanes have nothing to do
  const int spt = (P.Kc + CPS - 1) / CPS;                // slabs per tile
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t nslots = (uint32_t)P.nslots;
  const uint32_t warp_bytes = (uint32_t)warp_groups * 512u;                           // this warp's share of a slab
  const uint32_t leader_t0 = mapa(smem_u32(&ctl->landed_t[0]), 0);                    // both CTAs' TMA copies complete on the leader's barriers
  const int zero_row = (int)P.V;                         // row V of the shadow table is all zero (conv padding)
  const unsigned issue_mask = __activemask();
  auto rows_of = [&](const int4& tok, int (&out)[4]) {
    const int t4[4] = {tok.x, tok.y, tok.z, tok.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t4[k] < -1 || t4[k] >= P.V) __trap();          // the reference device-asserts on OOB ids
      out[k] = t4[k] < 0 ? zero_row : t4[k];
    }
  };
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_empty = 0, w_issue = 0, t_begin = clock64();
  const int ntiles = __ldg(P.head + cluster_id).x;
  const int4* srow = reinterpret_cast<const int4*>(P.stream + (long long)cluster_id * P.stream_stride + (long long)rank * TILE_M);
  int cur[GPL][4];
  int4 nxt[GPL];
#pragma unroll
  for (int j = 0; j < GPL; ++j) {
    nxt[j] = (ntiles > 0 && j < my_groups) ? __ldg(srow + grp[j]) : make_int4(-1, -1, -1, -1);
    rows_of(nxt[j], cur[j]);
  }
  uint32_t slab = 0;                                     // running slab index of this CTA: slot = slab % nslots, use = slab / nslots
  for (int t = 0; t < ntiles; ++t) {
    const bool more = t + 1 < ntiles;
    if (more) {
#pragma unroll
      for (int j = 0; j < GPL; ++j)
        if (j < my_groups) nxt[j] = __ldg(srow + (long long)(t + 1) * (2 * TILE_M / 4) + grp[j]);
    }
    for (int s = 0; s < P.tma_slabs; ++s) {
      const uint32_t slot = (slab + s) % nslots, use = (slab + s) / nslots;
      TIMED_WAIT(w_empty, mbar_wait(&ctl->empty[slot], (use & 1u) ^ 1u));
      const uint32_t bar = leader_t0 + slot * 8u;
      if (lane == 0) mbar_arrive_expect_tx_cluster(bar, warp_bytes);
      __syncwarp(issue_mask);
      const long long t_i = prof_on ? clock64() : 0;
#pragma unroll
      for (int j = 0; j < GPL; ++j)
        if (j < my_groups)
          tma_gather4(ring_base + slot * (uint32_t)SLAB_BYTES + (uint32_t)grp[j] * 512u, tmap, bar, s * CPS * 8, cur[j][0], cur[j][1], cur[j][2], cur[j][3]);
      if (prof_on) w_issue += clock64() - t_i;
    }
    slab += (uint32_t)spt;
    if (more) {
#pragma unroll
      for (int j = 0; j < GPL; ++j) rows_of(nxt[j], cur[j]);
    }
  }
  if (prof_on && pwarp == 0 && lane == 0) {
    unsigned long long* o = P.prof + rank * 16 + 4;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_empty; o[2] = w_issue;
  }
}

// ---- cp.async team: slabs tma_slabs .. spt-1 of every tile, in the same 128-byte-swizzled layout.  Every lane copies ONE
// 16-byte chunk column c8 of its nine rows r0 + 16k (r & 7 is the same for all of them, so the swizzled chunk is a
// per-thread constant).  The hand-off is asynchronous: after issuing its copies of a slab every thread posts a
// cp.async.mbarrier.arrive.noinc, which the hardware turns into an arrival once those copies have landed -- no
// wait_group, no MEMBAR in the producer (round 1's producer stalled ~700 cycles per slab on exactly that); the
// consumer side (MMA warp / relay) executes the generic->async proxy fence after the barrier completes.
__device__ __forceinline__ void ldg_role(const Params& P, SharedCtl* ctl, uint8_t* ring, uint32_t rank, int cluster_id, int ptid) {
  const int spt = (P.Kc + CPS - 1) / CPS;
  if (P.tma_slabs >= spt) return;                        // narrow rows: the TMA team stages everything
  const int c8 = ptid & 7, r0 = ptid >> 3;
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t dst_thread = (uint32_t)(r0 * 128 + ((c8 ^ (r0 & 7)) << 4));
  const uint32_t nslots = (uint32_t)P.nslots;
  const bool last_row = r0 + LDG_ROW_STEP * (ROWS_PER_THREAD - 1) < TILE_M + 2;   // the last pass only reaches rows < 130
  const uint8_t* const thread_base = P.shadow + c8 * 16;
  const uint8_t* const zero_row = thread_base + P.V * P.row_bytes;
  const int last_slab_chunks = P.Kc - (spt - 1) * CPS;                  // chunk columns the MMA reads in the last slab
  const bool in_last = c8 < last_slab_chunks;
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_empty = 0, t_begin = clock64();
  const int ntiles = __ldg(P.head + cluster_id).x;
  const int* srow = P.stream + (long long)cluster_id * P.stream_stride + (long long)rank * TILE_M + r0;
  auto tokens_of = [&](int t, int (&out)[ROWS_PER_THREAD]) {
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k)
      out[k] = (k < ROWS_PER_THREAD - 1 || last_row) ? __ldg(srow + (long long)t * (2 * TILE_M) + LDG_ROW_STEP * k) : -1;
  };
  int cur[ROWS_PER_THREAD], nxt[ROWS_PER_THREAD];
  if (ntiles > 0) tokens_of(0, cur);
  uint32_t slab = 0;
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) tokens_of(t + 1, nxt);
    const uint8_t* src[ROWS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) {
      if (cur[k] < -1 || cur[k] >= P.V) __trap();        // the reference device-asserts on OOB ids
      src[k] = cur[k] < 0 ? zero_row : thread_base + (long long)cur[k] * P.row_bytes;
    }
    for (int s = P.tma_slabs; s < spt; ++s) {
      const uint32_t slot = (slab + s) % nslots, use = (slab + s) / nslots;
      TIMED_WAIT(w_empty, mbar_wait(&ctl->empty[slot], (use & 1u) ^ 1u));
      const uint32_t dst = ring_base + slot * (uint32_t)SLAB_BYTES + dst_thread;
      if (s < spt - 1 || in_last) {
        const int cofs = s * CPS * 16;
#pragma unroll
        for (int k = 0; k < ROWS_PER_THREAD - 1; ++k) cp_async16_ca(dst + k * (LDG_ROW_STEP * 128), src[k] + cofs);
        if (last_row) cp_async16_ca(dst + (ROWS_PER_THREAD - 1) * (LDG_ROW_STEP * 128), src[ROWS_PER_THREAD - 1] + cofs);
      }
      cp_async_arrive_noinc(&ctl->landed_l[slot]);
    }
    slab += (uint32_t)spt;
#pragma unroll
    for (int k = 0; k < ROWS_PER_THREAD; ++k) cur[k] = nxt[k];
  }
  if (prof_on && ptid == 0) {
    unsigned long long* o = P.prof + rank * 16 + 12;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_empty; o[2] = 0;
  }
}

// Rank 1's otherwise idle MMA warp: tells the leader when each cp.async slab of THIS CTA has landed (the asynchronous
// arrivals of cp.async can only target a barrier of their own CTA; the pair's MMA is issued by rank 0).  TMA slabs need
// no relay: their copies complete on the leader's barrier directly.
__device__ __forceinline__ void relay_role(const Params& P, SharedCtl* ctl, int cluster_id) {
  const int spt = (P.Kc + CPS - 1) / CPS;
  const uint32_t leader_l0 = mapa(smem_u32(&ctl->landed_l[0]), 0);
  const uint32_t nslots = (uint32_t)P.nslots;
  const int ntiles = __ldg(P.head + cluster_id).x;
  uint32_t slab = 0, par_l = 0u;                         // per-slot phase parity of landed_l (bit = slot)
  for (int t = 0; t < ntiles; ++t)
    for (int s = 0; s < spt; ++s, ++slab) {
      const uint32_t slot = slab % nslots;
      if (s >= P.tma_slabs) {
        mbar_wait(&ctl->landed_l[slot], (par_l >> slot) & 1u); par_l ^= 1u << slot;
        fence_proxy_async();                              // cp.async wrote through the generic proxy, the tensor cores read through the async proxy
        mbar_arrive_cluster(leader_l0 + slot * 8u);
      }
    }
}

__device__ __forceinline__ void mma_role(const Params& P, SharedCtl* ctl, const uint8_t* bsm, const uint8_t* ring, int cluster_id, int lane) {
  const int spt = (P.Kc + CPS - 1) / CPS;
  const uint32_t idesc = umma_idesc(P.fmt, P.Npad);
  const uint32_t a_base = smem_u32(ring), b_base = smem_u32(bsm);
  const uint32_t b_lbo = (uint32_t)(P.Npad / 2) * 16;
  const int nslots = P.nslots;
  // low descriptor word = start address >> 4 | LBO >> 4 << 16; high word = SBO >> 4 | version 1 | layout type << 29
  //   A (TMA-written, SWIZZLE_128B): SBO = 1024 B between 8-row groups, LBO unused (1), layout type 2
  //   B (filter bank, no swizzle)  : SBO = 128 B, LBO = distance between the two K-chunks of a K=16 step
  const uint32_t a_desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t b_desc_hi = (128u >> 4) | (1u << 14);
  const uint32_t b_step = b_lbo >> 4;                           // one 16-byte K-chunk of the filter bank
  const uint32_t a_j = 128u >> 4;                               // window row j: +128 bytes (one row) in the A slab
  const uint32_t b_j = (uint32_t)P.Kc * b_step;                 //               +Kc chunks in the filter bank
  const uint32_t a_lo0 = ((a_base >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t b_lo0 = ((b_base >> 4) & 0x3FFFu) | (b_step << 16);
  const bool leader = elect_one();
  uint32_t slot = 0, it = 0, par_t = 0u, par_l = 0u;     // par_*: per-slot phase parity of the landed barriers (bit = slot)
  const bool prof_on = P.prof != nullptr && cluster_id == 0;
  long long w_tmem = 0, w_full = 0, t_begin = clock64();
  const uint32_t ntiles = (uint32_t)__ldg(P.head + cluster_id).x;
  {
    for (; it < ntiles; ++it) {                          // tiles of 256 consecutive windows of the pair's stream
      const uint32_t buf = it % NACC, use = it / NACC;
      TIMED_WAIT(w_tmem, mbar_wait(&ctl->tmem_empty[buf], (use & 1u) ^ 1u));       // both CTAs' epilogues drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = ctl->tmem_base + buf * ACC_STRIDE;
      for (int s = 0; s < spt; ++s) {
        // both CTAs' halves of the slab have landed (this CTA's team + the relayed arrival of rank 1)
        if (s < P.tma_slabs) { TIMED_WAIT(w_full, mbar_wait(&ctl->landed_t[slot], (par_t >> slot) & 1u)); par_t ^= 1u << slot; }
        else                 { TIMED_WAIT(w_full, mbar_wait(&ctl->landed_l[slot], (par_l >> slot) & 1u)); par_l ^= 1u << slot; fence_proxy_async(); }
        tc_fence_after();
        if (leader) {
          // descriptors advance by constants: a K=16 step is +32 B inside the swizzled 128-byte A row and +2 K-chunks in
          // the filter bank; window row j is +128 B (one row) in the A slab
          const int nk = min(CPS, P.Kc - s * CPS) >> 1;        // K=16 steps in this slab
          uint32_t a_lo = a_lo0 + ((slot * (uint32_t)SLAB_BYTES) >> 4);
          uint32_t b_lo = b_lo0 + (uint32_t)(s * CPS) * b_step;
          uint32_t acc = s ? 1u : 0u;
          for (int kk = 0; kk < nk; ++kk) {
            umma_f16_pair(d_tmem, mk_desc(a_lo, a_desc_hi), mk_desc(b_lo, b_desc_hi), idesc, acc);
            umma_f16_pair(d_tmem, mk_desc(a_lo + a_j, a_desc_hi), mk_desc(b_lo + b_j, b_desc_hi), idesc, 1u);
            umma_f16_pair(d_tmem, mk_desc(a_lo + 2 * a_j, a_desc_hi), mk_desc(b_lo + 2 * b_j, b_desc_hi), idesc, 1u);
            acc = 1u;
            a_lo += 32u >> 4;
            b_lo += 2 * b_step;
          }
          umma_commit_pair(&ctl->empty[slot]);                      // slot reusable (both CTAs) once these MMAs retire
          if (s == spt - 1) umma_commit_pair(&ctl->tmem_full[buf]); // accumulator complete (both CTAs)
        }
        __syncwarp();
        if (++slot == (uint32_t)nslots) slot = 0;
      }
    }
  }
  if (prof_on && leader) {
    unsigned long long* o = P.prof + 8;
    o[0] = (unsigned long long)(clock64() - t_begin); o[1] = w_tmem; o[2] = w_full;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) conv_pool_tc_kernel(const __grid_constant__ Params P,
                                                                                                 const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int nh = P.Npad / 2;
  const uint32_t b_bytes = (uint32_t)(3 * P.Kc * nh * 16);

  // identical carve-up in both CTAs: the pair's MMA addresses both through one descriptor
  SharedCtl* ctl = reinterpret_cast<SharedCtl*>(smem);
  uint8_t* bsm = smem + ((sizeof(SharedCtl) + 127u) & ~127u);
  // the ring starts on a 1024-byte boundary of the shared-memory WINDOW (the 128-byte swizzle is a function of the
  // absolute address; both CTAs get the same offset because their dynamic segments start at the same window offset)
  uint8_t* ring = smem + (((smem_u32(bsm) + ((b_bytes + 127u) & ~127u) + 1023u) & ~1023u) - smem_u32(smem));

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_SLOTS; ++i) {
      const uint32_t relay = rank == 0 ? 1u : 0u;
      mbar_init(&ctl->landed_t[i], 2 * NUM_TMA_WARPS); mbar_init(&ctl->landed_l[i], NUM_LDG_WARPS * 32 + relay);
      mbar_init(&ctl->empty[i], 1);
    }
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap) : "memory");
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
      case 8:  epilogue_role<8>(P, ctl, rank, cluster_id, warp, lane); break;
      case 16: epilogue_role<16>(P, ctl, rank, cluster_id, warp, lane); break;
      case 24: epilogue_role<24>(P, ctl, rank, cluster_id, warp, lane); break;
      case 32: epilogue_role<32>(P, ctl, rank, cluster_id, warp, lane); break;
      case 40: epilogue_role<40>(P, ctl, rank, cluster_id, warp, lane); break;
      case 48: epilogue_role<48>(P, ctl, rank, cluster_id, warp, lane); break;
      case 56: epilogue_role<56>(P, ctl, rank, cluster_id, warp, lane); break;
      default: epilogue_role<64>(P, ctl, rank, cluster_id, warp, lane); break;
    }
  } else if (warp < NUM_EPI_WARPS + NUM_LDG_WARPS) {
    ldg_role(P, ctl, ring, rank, cluster_id, threadIdx.x - NUM_EPI_WARPS * 32);
  } else if (warp < MMA_WARP) {
    tma_role(P, &tmap, ctl, ring, rank, cluster_id, warp - NUM_EPI_WARPS - NUM_LDG_WARPS, lane);
  } else if (rank == 0) {
    mma_role(P, ctl, bsm, ring, cluster_id, lane);
  } else if (lane == 0) {
    relay_role(P, ctl, cluster_id);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer may still read this CTA's smem / TMEM through the pair MMA
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(ctl->tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- window streams
// Work item k (position in doc_order, or the document itself without a plan) -> (cluster, index inside the cluster).
// Rounds of nc items alternate direction (0..nc-1, nc-1..0, ...): with doc_order sorted by decreasing length every
// pair of rounds hands each cluster nearly the same number of windows.
__device__ __forceinline__ void item_home(long long k, int nc, int& c, int& i) {
  i = (int)(k / nc);
  const int pos = (int)(k % nc);
  c = (i & 1) ? nc - 1 - pos : pos;
}
__device__ __forceinline__ long long item_of(int c, int i, int nc) { return (long long)i * nc + ((i & 1) ? nc - 1 - c : c); }

struct StreamWs {
  int2* head;                // [nc] (tiles, documents)
  int4* dlist;               // [nc][dlist_stride]
  int* stream;               // [nc][stream_stride]
  long long dlist_stride, stream_stride;
};

constexpr int PLAN_THREADS = 256;

// one block per cluster: exclusive scan of (effective length + 2) over the cluster's documents -> dlist, head, and the
// -1 rows that close the stream (two zero rows after the last document, then filler up to the last tile's halo)
__global__ void __launch_bounds__(PLAN_THREADS) stream_plan_kernel(long long N, int T, int nc, const int* __restrict__ doc_len,
                                                                   const int* __restrict__ doc_order, StreamWs W) {
  __shared__ int wsum[PLAN_THREADS / 32];
  __shared__ int carry_s;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rounds = N / nc;
  const int rem = (int)(N % nc);
  const int last_pos = (rounds & 1) ? nc - 1 - c : c;
  const int nd = (int)rounds + (last_pos < rem ? 1 : 0);
  int4* dl = W.dlist + (long long)c * W.dlist_stride;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nd; i0 += PLAN_THREADS) {
    const int i = i0 + tid;
    int doc = 0, len = 0, w = 0;
    if (i < nd) {
      const long long k = item_of(c, i, nc);
      doc = doc_order ? __ldg(doc_order + k) : (int)k;
      len = doc_len ? __ldg(doc_len + doc) : T;
      w = len + 2;
    }
    int incl = w;                                       // inclusive scan inside the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = carry_s;
    for (int w2 = 0; w2 < warp; ++w2) before += wsum[w2];
    if (i < nd) dl[i] = make_int4(doc, len, before + incl - w, 0);
    __syncthreads();
    if (tid == PLAN_THREADS - 1) carry_s = before + incl;
    __syncthreads();
  }
  const int total = carry_s;                            // windows of this cluster
  const int ntiles = (total + 2 * TILE_M - 1) / (2 * TILE_M);
  if (tid == 0) W.head[c] = make_int2(ntiles, nd);
  int* st = W.stream + (long long)c * W.stream_stride;
  for (int u = total + tid; u < ntiles * 2 * TILE_M + 4; u += PLAN_THREADS) st[u] = -1;
}

// one block per work item: the document's two opening zero rows and its rows' token ids (validated: the reference
// device-asserts on out-of-range ids).  Every thread issues its loads of a pass together (the kernel is latency-bound:
// a warp per document with one load in flight took 22 us per 4096 documents).
constexpr int FILL_THREADS = 128, FILL_UNROLL = 4;
__global__ void __launch_bounds__(FILL_THREADS) stream_fill_kernel(long long N, int T, int nc, long long V, const long long* __restrict__ idx,
                                                                   const int* __restrict__ tok32, const long long* __restrict__ off, long long pad_id,
                                                                   StreamWs W) {
  for (long long k = blockIdx.x; k < N; k += gridDim.x) {
    int c, i;
    item_home(k, nc, c, i);
    const int4 e = __ldg(W.dlist + (long long)c * W.dlist_stride + i);
    int* st = W.stream + (long long)c * W.stream_stride + e.z;
    if (threadIdx.x < 2) st[threadIdx.x] = -1;
    st += 2;
    const long long* row = idx ? idx + (long long)e.x * T : nullptr;
    long long base = 0;
    int len = 0;
    if (!idx) {
      base = __ldg(off + e.x);
      len = (int)(__ldg(off + e.x + 1) - base);
    }
    for (int p0 = 0; p0 < e.y; p0 += FILL_THREADS * FILL_UNROLL) {
      long long t[FILL_UNROLL];
#pragma unroll
      for (int u = 0; u < FILL_UNROLL; ++u) {
        const int p = p0 + u * FILL_THREADS + (int)threadIdx.x;
        t[u] = 0;
        if (p < e.y) t[u] = idx ? __ldg(row + p) : (p < len ? (long long)__ldg(tok32 + base + p) : pad_id);
      }
#pragma unroll
      for (int u = 0; u < FILL_UNROLL; ++u) {
        const int p = p0 + u * FILL_THREADS + (int)threadIdx.x;
        if (p < e.y) {
          if (t[u] < 0 || t[u] >= V) __trap();
          st[p] = (int)t[u];
        }
      }
    }
  }
}

// workspace carve-up for `nc` clusters (host): [head][dlist][stream], every part 16-byte aligned
inline long long stream_ws_layout(long long N, int T, int nc, void* ws, StreamWs* out) {
  const long long maxnd = (N + nc - 1) / nc;
  const long long dstride = maxnd;
  const long long sstride = ((maxnd * (T + 2) + 2 * TILE_M - 1) / (2 * TILE_M)) * (2 * TILE_M) + 8;   // whole tiles + the last tile's halo
  const long long head_b = (((long long)nc * 8) + 15) & ~15LL;
  const long long dl_b = (long long)nc * dstride * 16;
  const long long st_b = (long long)nc * sstride * 4;
  if (out) {
    uint8_t* b = static_cast<uint8_t*>(ws);
    out->head = reinterpret_cast<int2*>(b);
    out->dlist = reinterpret_cast<int4*>(b + head_b);
    out->stream = reinterpret_cast<int*>(b + head_b + dl_b);
    out->dlist_stride = dstride;
    out->stream_stride = sstride;
  }
  return head_b + dl_b + st_b;
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

// Workspace of one r4r_conv_pool_tc launch over N documents of T rows: the work lists and window streams of its CTA pairs,
// sized for the worst case (no padding to skip) at any cluster count the launch may use.
extern "C" int64_t r4r_conv_stream_ws_bytes(int64_t N, int T) {
  if (N < 0 || T <= 0) return -1;
  if (N == 0) return 16;
  static int pairs = 0;
  if (pairs == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    pairs = sms / 2 > 0 ? sms / 2 : 1;
  }
  // the byte count is not monotonic in the cluster count (per-cluster rounding): take the maximum over all of them
  long long best = 0;
  const int ncmax = (int)(N < pairs ? N : pairs);
  for (int nc = 1; nc <= ncmax; ++nc) {
    const long long b = stream_ws_layout(N, T, nc, nullptr, nullptr);
    if (b > best) best = b;
  }
  return best;
}

static int conv_pool_tc_launch(const void* shadow, int64_t V, int Epad, int E, int dtype,
                               const int64_t* idx, const int32_t* tok32, const int64_t* off, int64_t pad_id, int64_t N, int T,
                               const void* wpack, const float* conv_b, int F,
                               float* pooled, int32_t* argmax,
                               const int32_t* doc_len, const int32_t* doc_order, void* ws, void* stream) {
  R4R_REQUIRE(shadow && (idx || (tok32 && off)) && wpack && conv_b && pooled && argmax, R4R_EINVAL, "conv_pool_tc: null pointer");
  R4R_REQUIRE(ws || N == 0, R4R_EINVAL, "conv_pool_tc: null workspace (r4r_conv_stream_ws_bytes)");
  R4R_REQUIRE(N < (1LL << 31), R4R_EUNSUP, "conv_pool_tc: N=%lld documents exceed the work-list index range", (long long)N);
  R4R_REQUIRE(idx || (pad_id >= 0 && pad_id < V), R4R_EINVAL, "conv_pool_tc: pad id %lld outside the table", (long long)pad_id);
  R4R_REQUIRE(V > 0 && E > 0 && T > 0 && N >= 0, R4R_EINVAL, "conv_pool_tc: bad sizes");
  R4R_REQUIRE((T + 2 + 2 * TILE_M - 1) / (2 * TILE_M) <= 256, R4R_EUNSUP, "conv_pool_tc: T=%d exceeds 256 position tiles", T);
  R4R_REQUIRE(dtype == R4R_DT_F16 || dtype == R4R_DT_BF16, R4R_EINVAL, "conv_pool_tc: dtype %d", dtype);
  PackPlan pl;
  R4R_REQUIRE(make_plan(E, F, pl), R4R_EUNSUP, "conv_pool_tc: E=%d F=%d unsupported (F <= %d)", E, F, N_MAX);
  R4R_REQUIRE(Epad % 8 == 0 && Epad >= pl.Kc * 8, R4R_EINVAL, "conv_pool_tc: shadow row width Epad=%d must be a multiple of 8 and >= %d", Epad, pl.Kc * 8);
  R4R_REQUIRE(V + 1 < (1LL << 31), R4R_EUNSUP, "conv_pool_tc: V=%lld exceeds the TMA row coordinate range", (long long)V);
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
  // shared-memory plan: [ctl][resident filter half][pad to 1024][A ring of nslots x SLAB_BYTES]
  const long long ctl_bytes = (sizeof(SharedCtl) + 127) & ~127LL;
  const long long b_bytes = (pl.half_bytes + 127) & ~127LL;
  const long long avail = (long long)smem_optin - ctl_bytes - b_bytes - 1024;      // 1024: worst-case alignment pad of the ring
  R4R_REQUIRE(pl.Kc <= CPS * MAX_SPT, R4R_EUNSUP, "conv_pool_tc: E=%d exceeds %d", E, CPS * MAX_SPT * 8);
  long long ns = avail / SLAB_BYTES;
  if (ns > MAX_SLOTS) ns = MAX_SLOTS;
  const int nslots = (int)ns;
  R4R_REQUIRE(nslots >= 3, R4R_EUNSUP, "conv_pool_tc: E=%d F=%d leaves no room for the A ring next to the filter bank", E, F);
  const size_t smem_bytes = (size_t)(ctl_bytes + b_bytes + 1024 + (long long)nslots * SLAB_BYTES);

  // tensor map of the shadow table for the gather4 copies: [V+1 rows][Epad columns] half-precision, box = 64 columns x 1 row
  // (four such rows per instruction), SWIZZLE_128B; columns past Epad (a last, partial block) are zero-filled by the TMA
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    R4R_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    R4R_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, R4R_ENODEV, "conv_pool_tc: cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  CUtensorMap tmap;
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)Epad, (cuuint64_t)V + 1};
    const cuuint64_t gstride[1] = {(cuuint64_t)Epad * 2};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&tmap, dtype == R4R_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                              const_cast<void*>(shadow), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R4R_REQUIRE(r == CUDA_SUCCESS, R4R_EINVAL, "conv_pool_tc: cuTensorMapEncodeTiled failed (%d) for V=%lld Epad=%d", (int)r, (long long)V, Epad);
  }

  long long nclusters = sm_count / 2;
  {
    int want = g_clusters;                                  // r4r_conv_set_clusters: leave SM pairs free for a concurrent graph branch
    const char* e = getenv("R4R_CONV_CLUSTERS");            // tuning override
    if (e && atoi(e) >= 1) want = atoi(e);
    if (want >= 1 && want < nclusters) nclusters = want;
  }
  if (nclusters > N) nclusters = N;

  // window streams of this launch's CTA pairs (see Params): work list + row tokens, two small kernels
  StreamWs W;
  stream_ws_layout(N, T, (int)nclusters, ws, &W);
  R4R_REQUIRE(W.stream_stride < (1LL << 31), R4R_EUNSUP, "conv_pool_tc: %lld rows per CTA pair exceed the window index range", W.stream_stride);
  stream_plan_kernel<<<(unsigned)nclusters, PLAN_THREADS, 0, as_stream(stream)>>>(N, T, (int)nclusters, doc_len, doc_order, W);
  R4R_CHECK_LAUNCH("conv_pool_tc (stream plan)");
  {
    long long b = N;
    if (b > sm_count * 64) b = sm_count * 64;
    stream_fill_kernel<<<(unsigned)b, FILL_THREADS, 0, as_stream(stream)>>>(N, T, (int)nclusters, V, reinterpret_cast<const long long*>(idx),
                                                                 idx ? nullptr : tok32, reinterpret_cast<const long long*>(off), pad_id, W);
    R4R_CHECK_LAUNCH("conv_pool_tc (stream fill)");
  }

  Params P;
  P.shadow = static_cast<const uint8_t*>(shadow);
  P.row_bytes = (long long)Epad * 2;
  P.V = V;
  P.N = N; P.T = T; P.Kc = pl.Kc; P.F = F; P.Npad = pl.Npad;
  P.wpack = static_cast<const uint8_t*>(wpack);
  P.bias = conv_b; P.pooled = pooled; P.argmax = argmax;
  P.fmt = dtype; P.nslots = nslots; P.slot_bytes = SLAB_BYTES;
  {
    const int spt = (pl.Kc + CPS - 1) / CPS;
    P.tma_slabs = (2 * spt + 4) / 5;                        // 2 of 5 slabs at E = 300; every slab of narrow rows (spt <= 2: 1 of 1, 1 of 2)
    const char* e = getenv("R4R_CONV_TMA_SLABS");           // tuning override
    if (e && *e && atoi(e) >= 0) P.tma_slabs = atoi(e) < spt ? atoi(e) : spt;
  }
  P.prof = g_prof;
  P.stream = W.stream; P.stream_stride = W.stream_stride;
  P.dlist = W.dlist; P.dlist_stride = W.dlist_stride;
  P.head = W.head;

  R4R_CUDA(cudaFuncSetAttribute(conv_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  conv_pool_tc_kernel<<<(unsigned)(2 * nclusters), NUM_THREADS, smem_bytes, as_stream(stream)>>>(P, tmap);
  R4R_CHECK_LAUNCH("conv_pool_tc");
  return 0;
}

extern "C" int r4r_conv_pool_tc(const void* shadow, int64_t V, int Epad, int E, int dtype,
                                const int64_t* idx, int64_t N, int T,
                                const void* wpack, const float* conv_b, int F,
                                float* pooled, int32_t* argmax,
                                const int32_t* doc_len, const int32_t* doc_order, void* ws, void* stream) {
  R4R_REQUIRE(idx, R4R_EINVAL, "conv_pool_tc: null pointer");
  return conv_pool_tc_launch(shadow, V, Epad, E, dtype, idx, nullptr, nullptr, 0, N, T, wpack, conv_b, F, pooled, argmax,
                             doc_len, doc_order, ws, stream);
}

extern "C" int r4r_conv_pool_tc_ragged(const void* shadow, int64_t V, int Epad, int E, int dtype,
                                       const int32_t* tokens, const int64_t* offsets, int64_t pad_id, int64_t N, int T,
                                       const void* wpack, const float* conv_b, int F,
                                       float* pooled, int32_t* argmax,
                                       const int32_t* doc_len, const int32_t* doc_order, void* ws, void* stream) {
  R4R_REQUIRE(tokens && offsets, R4R_EINVAL, "conv_pool_tc_ragged: null pointer");
  return conv_pool_tc_launch(shadow, V, Epad, E, dtype, nullptr, tokens, offsets, pad_id, N, T, wpack, conv_b, F, pooled, argmax,
                             doc_len, doc_order, ws, stream);
}
