// gather4_bw.cu -- micro-benchmark for the TMA producer decision of conv_pool_tc (NOT part of the product).
//
// Question: how fast can persistent CTAs stage gathered shadow-table rows into shared memory with
// cp.async.bulk.tensor ... tile::gather4 (SWIZZLE_128B, 4 rows x 128 B per instruction), on the token statistics of
// the benchmark (Zipf(1.0) over 50,000 ids: the head row takes ~9 % of all fetches and TMA reads bypass L1)?
// The kernel only stages (a consumer thread frees each slab as soon as it lands): this is the producer's ceiling.
// Needed by the conv kernel: 1.66 M rows x 640 B per launch in <= 0.22 ms (the tensor-pipe time) = 4.8 TB/s.
//
// build:  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather4_bw gather4_bw.cu -lcuda
// run  :  timeout 60 ./gather4_bw          (ALWAYS under a short timeout; all device waits are bounded)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("driver error %s at %s:%d\n", s_, __FILE__, __LINE__); exit(1); } } while (0)

constexpr int V = 50001, EPAD = 320, TILE_ROWS = 128, CB = EPAD / 64;      // 5 column blocks of 64 halves (128 B)
constexpr int SLOTS = 8, SLAB_BYTES = TILE_ROWS * 128;                      // one slab = 128 rows x 128 B = 16 KB
constexpr long long SPIN = 1LL << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool wait_bounded(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  for (long long s = 0; !done && s < SPIN; ++s)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}

// PW producer warps (lane 0 of each issues); producer w stages row groups g = w, w + PW, ... of every slab
template <int PW>
__global__ void __launch_bounds__(32 * (PW + 1), 1) stage_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ tok,
                                                                 long long ntiles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[SLOTS], empty[SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&full[i])), "r"(PW));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < PW) {
    if (lane == 0) {
      uint32_t slot = 0, par = 1;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int* rows = tok + t * TILE_ROWS;
        for (int cb = 0; cb < CB; ++cb) {
          if (!wait_bounded(&empty[slot], par)) { atomicExch(err, 1); return; }
          constexpr int G = TILE_ROWS / 4 / PW;             // row groups of this producer per slab
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&full[slot])), "r"(G * 512) : "memory");
          for (int g = warp; g < TILE_ROWS / 4; g += PW) {
            const int4 r = __ldg(reinterpret_cast<const int4*>(rows) + g);
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                :: "r"(smem_u32(smem + slot * SLAB_BYTES + g * 512)), "l"(&tmap), "r"(smem_u32(&full[slot])), "r"(cb * 64),
                   "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
          }
          if (++slot == SLOTS) { slot = 0; par ^= 1; }
        }
      }
    }
  } else if (lane == 0) {                                     // consumer: frees every slab as soon as it has landed
    uint32_t slot = 0, par = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int cb = 0; cb < CB; ++cb) {
        if (!wait_bounded(&full[slot], par)) { atomicExch(err, 2); return; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&empty[slot])) : "memory");
        if (++slot == SLOTS) { slot = 0; par ^= 1; }
      }
  }
}

// ONE producer warp, all 32 lanes issue: lane g stages row group g of the slab (TMA instructions are per thread)
__global__ void __launch_bounds__(64, 1) stage_lanes_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ tok,
                                                            long long ntiles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[SLOTS], empty[SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t slot = 0, par = 1;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int4 r = __ldg(reinterpret_cast<const int4*>(tok + t * TILE_ROWS) + lane);     // this lane's four rows of the tile
      for (int cb = 0; cb < CB; ++cb) {
        if (!wait_bounded(&empty[slot], par)) { atomicExch(err, 1); return; }
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&full[slot])), "r"(SLAB_BYTES) : "memory");
        __syncwarp();
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            :: "r"(smem_u32(smem + slot * SLAB_BYTES + lane * 512)), "l"(&tmap), "r"(smem_u32(&full[slot])), "r"(cb * 64),
               "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
        if (++slot == SLOTS) { slot = 0; par ^= 1; }
      }
    }
  } else if (lane == 0) {
    uint32_t slot = 0, par = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int cb = 0; cb < CB; ++cb) {
        if (!wait_bounded(&full[slot], par)) { atomicExch(err, 2); return; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&empty[slot])) : "memory");
        if (++slot == SLOTS) { slot = 0; par ^= 1; }
      }
  }
}

template <int PW, int LN>
__global__ void __launch_bounds__(32 * (PW + 1), 1) stage_wl_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ tok,
                                                                    long long ntiles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long full[SLOTS], empty[SLOTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&full[i])), "r"(PW));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp < PW) {
    if (lane < LN) {
      constexpr int NI = PW * LN, G = TILE_ROWS / 4 / NI;   // row groups per issuing thread per slab
      const int me = warp * LN + lane;
      uint32_t slot = 0, par = 1;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int4 r[G];
        for (int k = 0; k < G; ++k) r[k] = __ldg(reinterpret_cast<const int4*>(tok + t * TILE_ROWS) + me + k * NI);
        for (int cb = 0; cb < CB; ++cb) {
          if (!wait_bounded(&empty[slot], par)) { atomicExch(err, 1); return; }
          if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&full[slot])), "r"(LN * G * 512) : "memory");
          __syncwarp((1u << LN) - 1u);
          for (int k = 0; k < G; ++k)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                :: "r"(smem_u32(smem + slot * SLAB_BYTES + (me + k * NI) * 512)), "l"(&tmap), "r"(smem_u32(&full[slot])), "r"(cb * 64),
                   "r"(r[k].x), "r"(r[k].y), "r"(r[k].z), "r"(r[k].w) : "memory");
          if (++slot == SLOTS) { slot = 0; par ^= 1; }
        }
      }
    }
  } else if (lane == 0) {
    uint32_t slot = 0, par = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      for (int cb = 0; cb < CB; ++cb) {
        if (!wait_bounded(&full[slot], par)) { atomicExch(err, 2); return; }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&empty[slot])) : "memory");
        if (++slot == SLOTS) { slot = 0; par ^= 1; }
      }
  }
}

template <int PW, int LN>
static void run_wl(const char* name, const CUtensorMap& tmap, const int* d_tok, long long ntiles, int* d_err) {
  const size_t smem = SLOTS * SLAB_BYTES + 1024;
  CK(cudaFuncSetAttribute(stage_wl_kernel<PW, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) stage_wl_kernel<PW, LN><<<148, 32 * (PW + 1), smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e0));
  const int iters = 5;
  for (int it = 0; it < iters; ++it) stage_wl_kernel<PW, LN><<<148, 32 * (PW + 1), smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
  int err; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  const double bytes = (double)ntiles * TILE_ROWS * EPAD * 2;
  printf("%-34s %d warps x %d issuing lanes: %.3f ms  = %.2f TB/s staged  (%s)\n", name, PW, LN, ms, bytes / ms / 1e9, err ? "TIMED OUT" : "ok");
}

static void run_lanes(const char* name, const CUtensorMap& tmap, const int* d_tok, long long ntiles, int* d_err) {
  const size_t smem = SLOTS * SLAB_BYTES + 1024;
  CK(cudaFuncSetAttribute(stage_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) stage_lanes_kernel<<<148, 64, smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e0));
  const int iters = 5;
  for (int it = 0; it < iters; ++it) stage_lanes_kernel<<<148, 64, smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
  int err; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  const double bytes = (double)ntiles * TILE_ROWS * EPAD * 2;
  printf("%-34s one warp, 32 issuing lanes: %.3f ms per %lld rows  = %.2f TB/s staged  (%s)\n", name, ms, ntiles * TILE_ROWS, bytes / ms / 1e9,
         err ? "TIMED OUT" : "ok");
}

template <int PW>
static void run(const char* name, const CUtensorMap& tmap, const int* d_tok, long long ntiles, int* d_err) {
  const size_t smem = SLOTS * SLAB_BYTES + 1024;
  CK(cudaFuncSetAttribute(stage_kernel<PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int it = 0; it < 2; ++it) stage_kernel<PW><<<148, 32 * (PW + 1), smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e0));
  const int iters = 5;
  for (int it = 0; it < iters; ++it) stage_kernel<PW><<<148, 32 * (PW + 1), smem>>>(tmap, d_tok, ntiles, d_err);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
  int err; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  const double bytes = (double)ntiles * TILE_ROWS * EPAD * 2;
  printf("%-34s producers/CTA %d: %.3f ms per %lld rows  = %.2f TB/s staged  (%s)\n", name, PW, ms, ntiles * TILE_ROWS, bytes / ms / 1e9,
         err ? "TIMED OUT" : "ok");
}

int main() {
  CK(cudaSetDevice(0));
  CKD(cuInit(0));
  const long long nrows = 4096LL * 404, ntiles = nrows / TILE_ROWS;
  __half* d_table;
  CK(cudaMalloc(&d_table, (size_t)(V + 1) * EPAD * sizeof(__half)));
  CK(cudaMemset(d_table, 0, (size_t)(V + 1) * EPAD * sizeof(__half)));
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {EPAD, V + 1};
  cuuint64_t gstride[1] = {EPAD * sizeof(__half)};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CKD(cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  std::vector<double> cdf(V - 1);
  double acc = 0;
  for (int r = 1; r < V; ++r) { acc += 1.0 / r; cdf[r - 1] = acc; }
  int *d_tok, *d_err;
  CK(cudaMalloc(&d_tok, nrows * sizeof(int)));
  CK(cudaMalloc(&d_err, 4));
  CK(cudaMemset(d_err, 0, 4));
  std::vector<int> tok(nrows);
  srand(1);
  for (int dist = 0; dist < 3; ++dist) {
    for (long long i = 0; i < nrows; ++i) {
      const double u = (rand() + 0.5) / ((double)RAND_MAX + 1.0);
      if (dist == 0) tok[i] = 1 + (int)(u * (V - 1));                                                    // uniform
      else if (dist == 1) tok[i] = 1 + (int)(std::lower_bound(cdf.begin(), cdf.end(), u * acc) - cdf.begin());   // Zipf(1.0)
      else tok[i] = 7;                                                                                    // one row (worst case)
    }
    CK(cudaMemcpy(d_tok, tok.data(), nrows * sizeof(int), cudaMemcpyHostToDevice));
    const char* name = dist == 0 ? "uniform ids" : dist == 1 ? "Zipf(1.0) ids (benchmark workload)" : "a single hot row";
    run<1>(name, tmap, d_tok, ntiles, d_err);
    run<2>(name, tmap, d_tok, ntiles, d_err);
    run<4>(name, tmap, d_tok, ntiles, d_err);
    run<8>(name, tmap, d_tok, ntiles, d_err);
    run<16>(name, tmap, d_tok, ntiles, d_err);
    run_lanes(name, tmap, d_tok, ntiles, d_err);
    run_wl<4, 2>(name, tmap, d_tok, ntiles, d_err);
    run_wl<4, 4>(name, tmap, d_tok, ntiles, d_err);
    run_wl<4, 8>(name, tmap, d_tok, ntiles, d_err);
    run_wl<2, 8>(name, tmap, d_tok, ntiles, d_err);
    run_wl<2, 16>(name, tmap, d_tok, ntiles, d_err);
  }
  return 0;
}
