// Probe of the tcgen05.ld .16x256b register layout (scripts/experiments, not part of the product).
// Every warp writes its 32 TMEM lanes with tcgen05.st.32x32b (lane i = row i, register c = column c; value = row << 8 | column)
// and reads them back with tcgen05.ld.16x256b.x2 at lane offsets 0 and 16: the printed (row, column) per thread and
// register is the layout the conv epilogue relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_ld_layout tmem_ld_layout.cu && ./tmem_ld_layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)c;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + ((uint32_t)(half * 16) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 8; ++k) out[((warp * 2 + half) * 32 + lane) * 8 + k] = r[k];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_s), "r"(32) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
  cudaMemset(d, 0xff, 4 * 2 * 32 * 8 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int warp = 0; warp < 4; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int lane = 0; lane < 32; ++lane)
        for (int k = 0; k < 8; ++k) {
          const uint32_t v = h[((warp * 2 + half) * 32 + lane) * 8 + k];
          const int row = (int)(v >> 8), col = (int)(v & 0xff);
          // expected: register k = 4*blk + 2*j + e  ->  row = warp*32 + half*16 + lane/4 + 8*j, column = 8*blk + 2*(lane%4) + e
          const int blk = k >> 2, j = (k >> 1) & 1, e2 = k & 1;
          const int erow = warp * 32 + half * 16 + lane / 4 + 8 * j, ecol = 8 * blk + 2 * (lane % 4) + e2;
          if (row != erow || col != ecol) {
            if (bad < 40) printf("warp %d half %d lane %2d reg %d: (row %d, col %d), expected (row %d, col %d)\n", warp, half, lane, k, row, col, erow, ecol);
            ++bad;
          }
        }
  printf("tmem_ld_layout 16x256b.x2: %d of %d registers differ from the expected layout\n", bad, 4 * 2 * 32 * 8);
  if (bad) {
    printf("warp 0 half 0 dump (lane: reg0..7 as row.col):\n");
    for (int lane = 0; lane < 32; ++lane) {
      printf("lane %2d:", lane);
      for (int k = 0; k < 8; ++k) { const uint32_t v = h[lane * 8 + k]; printf(" %3d.%-2d", (int)(v >> 8), (int)(v & 0xff)); }
      printf("\n");
    }
  }
  return 0;
}
