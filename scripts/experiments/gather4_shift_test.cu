// gather4_shift_test.cu -- micro-experiment for the planned TMA producer of conv_pool_tc (NOT part of the product,
// written offline at the end of round 1, first run in round 2: see scripts/experiments/README.md).
//
// Question: can the conv's three window rows be read from ONE staged tile when the tile is filled by
// cp.async.bulk.tensor ... tile::gather4 into the canonical K-major SWIZZLE_128B layout, by advancing the UMMA
// shared-memory descriptor's start address by j rows (j*128 B) -- and does that need base_offset = j?
//
// Setup: table[V][64] f16 with exactly representable values; 132 gathered rows (33 gather4 instructions) land at
// smem row r = base + r*128; B[16][64] (no-swizzle interleaved layout, the one conv_tc.cu already uses) selects
// column 4n, so   D_j[m][n] = A[m + j][4n] = table[idx[m + j]][4n]   for m < 128, j = 0..2.
// The program prints, for base_offset = 0 and base_offset = j, how many of the 3*128*16 outputs match.
//
// build:  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather4_shift_test gather4_shift_test.cu -lcuda
// run  :  timeout 20 ./gather4_shift_test          (ALWAYS under a short timeout)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("driver error %s at %s:%d\n", s_, __FILE__, __LINE__); exit(1); } } while (0)

constexpr int V = 512, C = 64, ROWS = 132, M = 128, N = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) test_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                                      float* __restrict__ out /*[3][M][N]*/, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_sm = smem;                                   // ROWS x 128 B, SWIZZLE_128B (written by TMA)
  uint8_t* b_sm = smem + 17 * 1024;                       // B: 8 chunks x 16 rows x 16 B, no swizzle, chunk-major
  __shared__ unsigned long long bar_tma, bar_mma;
  __shared__ uint32_t tmem_base;
  __shared__ int timed_out;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    timed_out = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_tma)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B[n][k] = (k == 4n): chunk c = k / 8 holds columns 8c..8c+7; address(n, c) = c * (N*16) + n * 16
  for (int i = tid; i < 8 * N * 8; i += 128) {
    const int e = i & 7, n = (i >> 3) % N, c = i / (8 * N);
    reinterpret_cast<__half*>(b_sm)[i] = __float2half((c * 8 + e) == 4 * n ? 1.0f : 0.0f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar_tma)), "r"(ROWS * 128) : "memory");
    for (int q = 0; q < ROWS / 4; ++q) {
      const int r0 = idx[4 * q], r1 = idx[4 * q + 1], r2 = idx[4 * q + 2], r3 = idx[4 * q + 3];
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
          " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
          :: "r"(smem_u32(a_sm + q * 512)), "l"(&tmap), "r"(smem_u32(&bar_tma)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
          : "memory");
    }
    // wait for the rows (bounded: a TMA that never completes must not hang the GPU)
    uint32_t done = 0;
    for (long long spin = 0; !done && spin < (1LL << 22); ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar_tma)) : "memory");
    if (!done) { out[0] = -12345.0f; timed_out = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: D f32, A/B f16, K-major both, N, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t b_lbo = N * 16;
    for (int j = 0; j < 3; ++j) {
      for (int ks = 0; ks < 4; ++ks) {                    // K = 16 per step: 32 B inside the swizzled 128-byte row
        const uint32_t a_addr = smem_u32(a_sm) + j * 128 + ks * 32;
        uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) /* LBO unused */ |
                         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        if (use_base_offset) adesc |= (uint64_t)(j & 7) << 49;
        const uint32_t b_addr = smem_u32(b_sm) + ks * 2 * b_lbo;
        const uint64_t bdesc = (uint64_t)((b_addr >> 4) & 0x3FFFu) | ((uint64_t)(b_lbo >> 4) << 16) |
                               ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t d = tmem_base + j * N;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(ks ? 1u : 0u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma)) : "memory");
  }
  // everybody waits for the MMAs (bounded)
  {
    uint32_t done = 0;
    for (long long spin = 0; !done && spin < (1LL << 24); ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar_mma)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int j = 0; j < 3; ++j) {
    uint32_t v[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + j * N;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < N; ++n) out[(j * M + warp * 32 + lane) * N + n] = __uint_as_float(v[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem_base) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  CKD(cuInit(0));
  std::vector<__half> table((size_t)V * C);
  for (int v = 0; v < V; ++v)
    for (int c = 0; c < C; ++c) table[(size_t)v * C + c] = __float2half((float)(((v * 3 + c * 7) % 97) - 48));
  std::vector<int> idx(ROWS);
  srand(7);
  for (int r = 0; r < ROWS; ++r) idx[r] = rand() % V;
  __half* d_table; int* d_idx; float* d_out;
  CK(cudaMalloc(&d_table, table.size() * sizeof(__half)));
  CK(cudaMalloc(&d_idx, ROWS * sizeof(int)));
  CK(cudaMalloc(&d_out, 3 * M * N * sizeof(float)));
  CK(cudaMemcpy(d_table, table.data(), table.size() * sizeof(__half), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), ROWS * sizeof(int), cudaMemcpyHostToDevice));

  for (int box_rows = 1; box_rows <= 4; box_rows += 3) {          // which box height does tile::gather4 expect: 1 or 4?
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {C, V};
    cuuint64_t gstride[1] = {C * sizeof(__half)};
    cuuint32_t box[2] = {C, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_table, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box rows %d: cuTensorMapEncodeTiled failed (%d)\n", box_rows, (int)r); continue; }
    for (int use_bo = 0; use_bo < 2; ++use_bo) {
      CK(cudaMemset(d_out, 0xff, 3 * M * N * sizeof(float)));
      const size_t smem = 17 * 1024 + 8 * N * 16 + 1024;
      CK(cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      test_kernel<<<1, 128, smem>>>(tmap, d_idx, d_out, use_bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("box rows %d, base_offset %d: kernel failed: %s\n", box_rows, use_bo, cudaGetErrorString(e)); return 1; }
      std::vector<float> out(3 * M * N);
      CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
      for (int j = 0; j < 3; ++j) {
        int ok = 0;
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n)
            ok += out[(j * M + m) * N + n] == __half2float(table[(size_t)idx[m + j] * C + 4 * n]);
        printf("box rows %d  base_offset %s  shift j=%d : %d / %d outputs correct\n", box_rows, use_bo ? "= j" : "= 0", j, ok, M * N);
      }
    }
  }
  return 0;
}
