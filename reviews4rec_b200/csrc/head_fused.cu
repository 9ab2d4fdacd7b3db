// head_fused.cu -- K3: the whole DeepCoNN / DeepCoNN++ head behind the two TextCNN towers in ONE forward and ONE
// backward kernel (SURVEY.md section 2 kernel table):
//   forward : lat_t = fc_t(pooled_t) (common_pytorch_models.py:33-37) -> dropout (Philox4x32-10 in-kernel, or masks
//             handed in by a test) -> cat (DeepCoNN.py:61) -> deepconn:   global_bias + TorchFM(cat)   (DeepCoNN.py:64-66,
//             common_pytorch_models.py:49-57)  |  deepconn++: final MLP (2L -> L, ReLU, dropout, L -> 1) + user_bias[u] +
//             item_bias[i] + global_bias (DeepCoNN.py:69-72) -> rating [-> (rating - y)^2 and its batch sum, loss.py:7-11]
//   backward: analytic gradients of all of the above: d pooled_t (for the conv weight-gradient kernels), d fc_t, d FM /
//             d MLP parameters, d global_bias, d (gathered bias values)
// A few hundred FLOPs per rating: one warp per rating, lanes over the F pooled features, warp-shuffle reductions for the
// FC / FM sums (north_star: "warp-shuffle reductions for the FM sum-of-squares"); parameter gradients are accumulated
// per CTA in shared memory and committed with one atomic per element per CTA.
#include "common.cuh"

namespace {
constexpr int THREADS = 256, WARPS = THREADS / 32;
constexpr int LMAX = 32, FMAX = 128, KMAX = 16;

struct HeadArgs {
  const float* pooled[2];      // [N, F] user / item tower
  const float* fc_w[2];        // [L, F]
  const float* fc_b[2];        // [L]
  const float* fm_V;           // [2L, K]          (head 0)
  const float* fm_w;           // [2L]
  const float* fm_b;           // [1]
  const float* w0;             // [L, 2L]          (head 1)
  const float* b0;             // [L]
  const float* w3;             // [L]
  const float* b3;             // [1]
  const float* ub;             // [N] gathered user_bias values (head 1)
  const float* ib;             // [N]
  const float* global_bias;    // [1]
  const float* y;              // [N] or NULL
  const uint8_t* mask_ext;     // [N, 3L] keep masks handed in (tests), or NULL = Philox
  const int* step;             // device step counter (Philox offset; advances once per training step), or NULL = 0
  unsigned long long seed;
  float p;                     // dropout probability (0 in eval mode)
  int N, F, L, K, head;
  float* rating;               // [N]
  float* se;                   // [N] or NULL
  float* se_sum;               // += sum of se, or NULL
  float* cat;                  // [N, 2L] post-dropout latent (saved for backward)
  float* hid;                  // [N, L]  post-ReLU, post-dropout hidden (head 1, saved)
  uint32_t* keep;              // [N, 3]  keep bits: word 0 user latent, 1 item latent, 2 hidden
};

struct HeadGrads {
  const float* g_rating;       // [N] upstream gradient of rating, or NULL
  const float* g_se;           // [N] upstream gradient of se, or NULL
  float* dpooled[2];           // [N, F]
  float* dfc_w[2];             // [L, F]   (all parameter gradients are ACCUMULATED: zero them first)
  float* dfc_b[2];             // [L]
  float* dV; float* dfm_w; float* dfm_b;
  float* dw0; float* db0; float* dw3; float* db3;
  float* dub; float* dib;      // [N] gradient of the gathered bias values (written)
  float* dglobal;              // [1]
};

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// keep words of rating n: bit q of the 96-bit string = keep flag of element q (0..L-1 user latent, L..2L-1 item latent,
// 2L..3L-1 hidden).  Lane j draws the flags of elements j, 32 + j, 64 + j from ONE Philox call.
__device__ __forceinline__ void keep_words(const HeadArgs& A, int n, int lane, uint32_t (&w)[3]) {
  if (A.p <= 0.0f) { w[0] = w[1] = w[2] = 0xffffffffu; return; }
  bool k[3];
  if (A.mask_ext) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int q = 32 * i + lane;
      k[i] = q < 3 * A.L ? A.mask_ext[(size_t)n * 3 * A.L + q] != 0 : true;
    }
  } else {
    uint32_t c[4] = {(uint32_t)n, (uint32_t)lane, A.step ? (uint32_t)__ldg(A.step) : 0u, 0x5eedu};
    philox4x32_10(c, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
#pragma unroll
    for (int i = 0; i < 3; ++i) k[i] = (float)c[i] * 2.3283064365386963e-10f >= A.p;     // uniform [0,1) >= p  <=>  keep
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) w[i] = __ballot_sync(0xffffffffu, k[i]);
}
__device__ __forceinline__ bool keep_bit(const uint32_t (&w)[3], int q) { return (w[q >> 5] >> (q & 31)) & 1u; }

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(THREADS) deepconn_head_fwd_kernel(const __grid_constant__ HeadArgs A) {
  extern __shared__ float sm[];
  const int L = A.L, F = A.F, K = A.K, L2 = 2 * A.L;
  float* sW = sm;                              // [2][L][F]
  float* sH = sW + 2 * L * F;                  // head parameters: FM  V [2L][K], w [2L]   |   MLP  W0 [L][2L], b0 [L], w3 [L]
  float* sC = sH + (A.head == 0 ? L2 * K + L2 : L * L2 + 2 * L);      // per-warp cat scratch [WARPS][2L]
  for (int t = 0; t < 2; ++t)
    for (int i = threadIdx.x; i < L * F; i += THREADS) sW[t * L * F + i] = __ldg(A.fc_w[t] + i);
  if (A.head == 0) {
    for (int i = threadIdx.x; i < L2 * K; i += THREADS) sH[i] = __ldg(A.fm_V + i);
    for (int i = threadIdx.x; i < L2; i += THREADS) sH[L2 * K + i] = __ldg(A.fm_w + i);
  } else {
    for (int i = threadIdx.x; i < L * L2; i += THREADS) sH[i] = __ldg(A.w0 + i);
    for (int i = threadIdx.x; i < L; i += THREADS) { sH[L * L2 + i] = __ldg(A.b0 + i); sH[L * L2 + L + i] = __ldg(A.w3 + i); }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = A.p > 0.0f ? 1.0f / (1.0f - A.p) : 1.0f;
  const float gb = __ldg(A.global_bias);
  float* myc = sC + warp * L2;
  float se_acc = 0.0f;
  for (int n = blockIdx.x * WARPS + warp; n < A.N; n += gridDim.x * WARPS) {
    uint32_t kw[3];
    keep_words(A, n, lane, kw);
    // ---- two FC layers: lanes stride the F features; the L partial sums are reduced with shuffles
    for (int t = 0; t < 2; ++t) {
      const float* x = A.pooled[t] + (size_t)n * F;
      float xr[FMAX / 32];
#pragma unroll
      for (int k = 0; k < FMAX / 32; ++k) xr[k] = (lane + 32 * k) < F ? __ldg(x + lane + 32 * k) : 0.0f;
      for (int l = 0; l < L; ++l) {
        const float* w = sW + (t * L + l) * F;
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < FMAX / 32; ++k)
          if (lane + 32 * k < F) s = fmaf(xr[k], w[lane + 32 * k], s);
        s = warp_sum(s);
        if (lane == 0) {
          const float v = s + __ldg(A.fc_b[t] + l);
          myc[t * L + l] = keep_bit(kw, t * L + l) ? v * scale : 0.0f;
        }
      }
    }
    __syncwarp();
    if (lane < L2) A.cat[(size_t)n * L2 + lane] = myc[lane];
    float r = 0.0f;
    if (A.head == 0) {
      // TorchFM: 0.5 * (sum_c (cat V)_c^2 - sum_c (cat^2 V^2)_c) + cat . w + b      (lane c < K owns column c)
      float s1 = 0.0f, s2 = 0.0f, lin = 0.0f;
      if (lane < K) {
        float s = 0.0f, q = 0.0f;
        for (int i = 0; i < L2; ++i) {
          const float c = myc[i], v = sH[i * K + lane];
          s = fmaf(c, v, s);
          q = fmaf(c * c, v * v, q);
        }
        s1 = s * s;
        s2 = q;
      }
      if (lane < L2) lin = myc[lane] * sH[L2 * K + lane];
      s1 = warp_sum(s1); s2 = warp_sum(s2); lin = warp_sum(lin);
      r = gb + (0.5f * (s1 - s2) + (lin + __ldg(A.fm_b)));
    } else {
      // final MLP: h = dropout(relu(W0 cat + b0)); rating = w3 . h + b3 + user_bias[u] + item_bias[i] + global_bias
      float hv = 0.0f;
      if (lane < L) {
        float s = sH[L * L2 + lane];
        for (int i = 0; i < L2; ++i) s = fmaf(myc[i], sH[lane * L2 + i], s);
        s = s > 0.0f ? s : 0.0f;
        hv = keep_bit(kw, L2 + lane) ? s * scale : 0.0f;
        A.hid[(size_t)n * L + lane] = hv;
        hv *= sH[L * L2 + L + lane];
      }
      r = warp_sum(hv) + __ldg(A.b3);
      r = r + __ldg(A.ub + n) + __ldg(A.ib + n) + gb;
    }
    __syncwarp();
    if (lane == 0) {
      A.rating[n] = r;
      A.keep[(size_t)n * 3 + 0] = kw[0]; A.keep[(size_t)n * 3 + 1] = kw[1]; A.keep[(size_t)n * 3 + 2] = kw[2];
      if (A.se) {
        const float d = r - __ldg(A.y + n);
        A.se[n] = d * d;
        se_acc += d * d;
      }
    }
  }
  if (A.se_sum) {
    __shared__ float red[WARPS];
    if (lane == 0) red[warp] = se_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.0f;
      for (int w = 0; w < WARPS; ++w) s += red[w];
      if (s != 0.0f) atomicAdd(A.se_sum, s);
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(THREADS) deepconn_head_bwd_kernel(const __grid_constant__ HeadArgs A, const __grid_constant__ HeadGrads G) {
  extern __shared__ float sm[];
  const int L = A.L, F = A.F, K = A.K, L2 = 2 * A.L;
  const int nH = A.head == 0 ? L2 * K + L2 : L * L2 + 2 * L;
  float* sW = sm;                              // [2][L][F] fc weights
  float* sH = sW + 2 * L * F;                  // head parameters (as in the forward)
  float* aW = sH + nH;                         // accumulators: d fc_w [2][L][F]
  float* aB = aW + 2 * L * F;                  //               d fc_b [2][L]
  float* aH = aB + 2 * L;                      //               d head parameters (same layout as sH) + [nH] = d scalar bias, [nH+1] = d global
  float* sD = aH + nH + 2;                     // per-warp scratch: dlat [WARPS][2L]
  for (int t = 0; t < 2; ++t)
    for (int i = threadIdx.x; i < L * F; i += THREADS) sW[t * L * F + i] = __ldg(A.fc_w[t] + i);
  if (A.head == 0) {
    for (int i = threadIdx.x; i < L2 * K; i += THREADS) sH[i] = __ldg(A.fm_V + i);
    for (int i = threadIdx.x; i < L2; i += THREADS) sH[L2 * K + i] = __ldg(A.fm_w + i);
  } else {
    for (int i = threadIdx.x; i < L * L2; i += THREADS) sH[i] = __ldg(A.w0 + i);
    for (int i = threadIdx.x; i < L; i += THREADS) { sH[L * L2 + i] = __ldg(A.b0 + i); sH[L * L2 + L + i] = __ldg(A.w3 + i); }
  }
  for (int i = threadIdx.x; i < 2 * L * F + 2 * L + nH + 2; i += THREADS) aW[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = A.p > 0.0f ? 1.0f / (1.0f - A.p) : 1.0f;
  float* myd = sD + warp * L2;
  for (int n = blockIdx.x * WARPS + warp; n < A.N; n += gridDim.x * WARPS) {
    uint32_t kw[3] = {A.keep[(size_t)n * 3], A.keep[(size_t)n * 3 + 1], A.keep[(size_t)n * 3 + 2]};
    float dr = G.g_rating ? __ldg(G.g_rating + n) : 0.0f;
    if (G.g_se) dr = fmaf(2.0f * (A.rating[n] - __ldg(A.y + n)), __ldg(G.g_se + n), dr);
    const float c = lane < L2 ? A.cat[(size_t)n * L2 + lane] : 0.0f;          // lane i < 2L owns cat_i
    float dc = 0.0f;                                                             // d loss / d cat_i
    if (A.head == 0) {
      // s_c for every column (lanes broadcast cat through shuffles)
      float sc = 0.0f;                                                           // lane c < K: s_c = sum_i cat_i V_ic
      for (int i = 0; i < L2; ++i) {
        const float ci = __shfl_sync(0xffffffffu, c, i);
        if (lane < K) sc = fmaf(ci, sH[i * K + lane], sc);
      }
      for (int cc = 0; cc < K; ++cc) {
        const float s = __shfl_sync(0xffffffffu, sc, cc);
        if (lane < L2) {
          const float v = sH[lane * K + cc];
          dc += s * v - c * v * v;
          atomicAdd(aH + lane * K + cc, dr * (s * c - c * c * v));              // dV_ic
        }
      }
      if (lane < L2) {
        dc = dr * (dc + sH[L2 * K + lane]);
        atomicAdd(aH + L2 * K + lane, dr * c);                                   // d fm.lin.weight
      }
      if (lane == 0) { atomicAdd(aH + nH, dr); atomicAdd(aH + nH + 1, dr); }     // d fm.lin.bias, d global_bias
    } else {
      const float hv = lane < L ? A.hid[(size_t)n * L + lane] : 0.0f;            // post-ReLU, post-dropout
      float dh = 0.0f;                                                           // d loss / d (pre-activation of hidden l)
      if (lane < L) {
        atomicAdd(aH + L * L2 + L + lane, dr * hv);                              // d final.3.weight
        dh = hv > 0.0f ? dr * sH[L * L2 + L + lane] * scale : 0.0f;              // kept and positive (hv > 0 <=> both)
        atomicAdd(aH + L * L2 + lane, dh);                                       // d final.0.bias
      }
      for (int l = 0; l < L; ++l) {
        const float d = __shfl_sync(0xffffffffu, dh, l);
        if (lane < L2) {
          dc = fmaf(d, sH[l * L2 + lane], dc);
          atomicAdd(aH + l * L2 + lane, d * c);                                  // d final.0.weight
        }
      }
      if (lane == 0) {
        atomicAdd(aH + nH, dr);                                                  // d final.3.bias
        atomicAdd(aH + nH + 1, dr);                                              // d global_bias
        G.dub[n] = dr;
        G.dib[n] = dr;
      }
    }
    // through the latent dropout
    if (lane < L2) myd[lane] = keep_bit(kw, lane) ? dc * scale : 0.0f;
    __syncwarp();
    for (int t = 0; t < 2; ++t) {
      const float* x = A.pooled[t] + (size_t)n * F;
      float* dx = G.dpooled[t] + (size_t)n * F;
      if (lane < L) atomicAdd(aB + t * L + lane, myd[t * L + lane]);
#pragma unroll
      for (int k = 0; k < FMAX / 32; ++k) {
        const int f = lane + 32 * k;
        if (f < F) {
          const float xv = __ldg(x + f);
          float s = 0.0f;
          for (int l = 0; l < L; ++l) {
            const float d = myd[t * L + l];
            s = fmaf(d, sW[(t * L + l) * F + f], s);
            atomicAdd(aW + (t * L + l) * F + f, d * xv);
          }
          dx[f] = s;
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // the dropout stream advances once per training step: after the step's only backward launch (all forward CTAs are done)
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.step) *const_cast<int*>(A.step) += 1;
  // commit the CTA's sums: one atomic per element
  for (int i = threadIdx.x; i < L * F; i += THREADS) {
    if (aW[i] != 0.0f) atomicAdd(G.dfc_w[0] + i, aW[i]);
    if (aW[L * F + i] != 0.0f) atomicAdd(G.dfc_w[1] + i, aW[L * F + i]);
  }
  for (int i = threadIdx.x; i < L; i += THREADS) {
    atomicAdd(G.dfc_b[0] + i, aB[i]);
    atomicAdd(G.dfc_b[1] + i, aB[L + i]);
  }
  if (A.head == 0) {
    for (int i = threadIdx.x; i < L2 * K; i += THREADS) atomicAdd(G.dV + i, aH[i]);
    for (int i = threadIdx.x; i < L2; i += THREADS) atomicAdd(G.dfm_w + i, aH[L2 * K + i]);
    if (threadIdx.x == 0) atomicAdd(G.dfm_b, aH[nH]);
  } else {
    for (int i = threadIdx.x; i < L * L2; i += THREADS) atomicAdd(G.dw0 + i, aH[i]);
    for (int i = threadIdx.x; i < L; i += THREADS) { atomicAdd(G.db0 + i, aH[L * L2 + i]); atomicAdd(G.dw3 + i, aH[L * L2 + L + i]); }
    if (threadIdx.x == 0) atomicAdd(G.db3, aH[nH]);
  }
  if (threadIdx.x == 0) atomicAdd(G.dglobal, aH[nH + 1]);
}

inline size_t head_smem(int L, int F, int K, int head, bool bwd) {
  const size_t nH = head == 0 ? (size_t)2 * L * K + 2 * L : (size_t)L * 2 * L + 2 * L;
  size_t fl = (size_t)2 * L * F + nH + (size_t)WARPS * 2 * L;
  if (bwd) fl += (size_t)2 * L * F + 2 * L + nH + 2;
  return fl * sizeof(float);
}
inline unsigned head_grid(int N) {
  int64_t b = cdiv64(N, WARPS * 4);          // ~4 ratings per warp: the weights are re-staged per CTA
  if (b > 148 * 2) b = 148 * 2;
  return (unsigned)(b < 1 ? 1 : b);
}
}  // namespace

// `ptrs` = the device pointers of HeadArgs / HeadGrads in declaration order (NULL where a head does not use one);
// plain pointers and sizes only (C ABI).  Forward: 25 pointers; backward: the same 25 + the 18 of HeadGrads.
extern "C" int r4r_deepconn_head_fwd(const void* const* ptrs, int N, int F, int L, int K, int head, float p, uint64_t seed, void* stream) {
  R4R_REQUIRE(ptrs, R4R_EINVAL, "deepconn_head_fwd: null pointer");
  R4R_REQUIRE(N >= 0 && F > 0 && F <= FMAX && L > 0 && L <= LMAX && K >= 0 && K <= KMAX && (head == 0 || head == 1) && p >= 0.0f && p < 1.0f,
              R4R_EUNSUP, "deepconn_head_fwd: N=%d F=%d (<= %d) L=%d (<= %d) K=%d (<= %d) head=%d p=%g", N, F, FMAX, L, LMAX, K, KMAX, head, p);
  if (N == 0) return 0;
  HeadArgs A;
  int q = 0;
  auto f = [&]() { return static_cast<const float*>(ptrs[q++]); };
  A.pooled[0] = f(); A.pooled[1] = f(); A.fc_w[0] = f(); A.fc_w[1] = f(); A.fc_b[0] = f(); A.fc_b[1] = f();
  A.fm_V = f(); A.fm_w = f(); A.fm_b = f(); A.w0 = f(); A.b0 = f(); A.w3 = f(); A.b3 = f(); A.ub = f(); A.ib = f();
  A.global_bias = f(); A.y = f();
  A.mask_ext = static_cast<const uint8_t*>(ptrs[q++]);
  A.step = static_cast<const int*>(ptrs[q++]);
  A.rating = const_cast<float*>(f()); A.se = const_cast<float*>(f()); A.se_sum = const_cast<float*>(f());
  A.cat = const_cast<float*>(f()); A.hid = const_cast<float*>(f());
  A.keep = static_cast<uint32_t*>(const_cast<void*>(ptrs[q++]));
  A.seed = seed; A.p = p; A.N = N; A.F = F; A.L = L; A.K = K; A.head = head;
  R4R_REQUIRE(A.pooled[0] && A.pooled[1] && A.fc_w[0] && A.fc_w[1] && A.fc_b[0] && A.fc_b[1] && A.global_bias && A.rating && A.cat && A.keep,
              R4R_EINVAL, "deepconn_head_fwd: null pointer");
  R4R_REQUIRE(head == 0 ? (A.fm_V && A.fm_w && A.fm_b) : (A.w0 && A.b0 && A.w3 && A.b3 && A.ub && A.ib && A.hid), R4R_EINVAL,
              "deepconn_head_fwd: null head parameter");
  R4R_REQUIRE(!A.se || A.y, R4R_EINVAL, "deepconn_head_fwd: se wanted without y");
  const size_t smem = head_smem(L, F, K, head, false);
  R4R_CUDA(cudaFuncSetAttribute(deepconn_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  deepconn_head_fwd_kernel<<<head_grid(N), THREADS, smem, as_stream(stream)>>>(A);
  R4R_CHECK_LAUNCH("deepconn_head_fwd");
  return 0;
}

extern "C" int r4r_deepconn_head_bwd(const void* const* ptrs, const void* const* gptrs, int N, int F, int L, int K, int head, float p,
                                     void* stream) {
  R4R_REQUIRE(ptrs && gptrs, R4R_EINVAL, "deepconn_head_bwd: null pointer");
  R4R_REQUIRE(N >= 0 && F > 0 && F <= FMAX && L > 0 && L <= LMAX && K >= 0 && K <= KMAX && (head == 0 || head == 1), R4R_EUNSUP,
              "deepconn_head_bwd: unsupported sizes");
  if (N == 0) return 0;
  HeadArgs A;
  int q = 0;
  auto f = [&]() { return static_cast<const float*>(ptrs[q++]); };
  A.pooled[0] = f(); A.pooled[1] = f(); A.fc_w[0] = f(); A.fc_w[1] = f(); A.fc_b[0] = f(); A.fc_b[1] = f();
  A.fm_V = f(); A.fm_w = f(); A.fm_b = f(); A.w0 = f(); A.b0 = f(); A.w3 = f(); A.b3 = f(); A.ub = f(); A.ib = f();
  A.global_bias = f(); A.y = f();
  A.mask_ext = static_cast<const uint8_t*>(ptrs[q++]);
  A.step = static_cast<const int*>(ptrs[q++]);
  A.rating = const_cast<float*>(f()); A.se = const_cast<float*>(f()); A.se_sum = const_cast<float*>(f());
  A.cat = const_cast<float*>(f()); A.hid = const_cast<float*>(f());
  A.keep = static_cast<uint32_t*>(const_cast<void*>(ptrs[q++]));
  A.seed = 0; A.p = p; A.N = N; A.F = F; A.L = L; A.K = K; A.head = head;
  HeadGrads G;
  q = 0;
  auto g = [&]() { return static_cast<float*>(const_cast<void*>(gptrs[q++])); };
  G.g_rating = g(); G.g_se = g();
  G.dpooled[0] = g(); G.dpooled[1] = g(); G.dfc_w[0] = g(); G.dfc_w[1] = g(); G.dfc_b[0] = g(); G.dfc_b[1] = g();
  G.dV = g(); G.dfm_w = g(); G.dfm_b = g(); G.dw0 = g(); G.db0 = g(); G.dw3 = g(); G.db3 = g(); G.dub = g(); G.dib = g(); G.dglobal = g();
  R4R_REQUIRE((G.g_rating || G.g_se) && G.dpooled[0] && G.dpooled[1] && G.dfc_w[0] && G.dfc_w[1] && G.dfc_b[0] && G.dfc_b[1] && G.dglobal,
              R4R_EINVAL, "deepconn_head_bwd: null pointer");
  R4R_REQUIRE(!G.g_se || A.y, R4R_EINVAL, "deepconn_head_bwd: g_se without y");
  R4R_REQUIRE(head == 0 ? (G.dV && G.dfm_w && G.dfm_b) : (G.dw0 && G.db0 && G.dw3 && G.db3 && G.dub && G.dib), R4R_EINVAL,
              "deepconn_head_bwd: null head gradient");
  const size_t smem = head_smem(L, F, K, head, true);
  R4R_CUDA(cudaFuncSetAttribute(deepconn_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  deepconn_head_bwd_kernel<<<head_grid(N), THREADS, smem, as_stream(stream)>>>(A, G);
  R4R_CHECK_LAUNCH("deepconn_head_bwd");
  return 0;
}
