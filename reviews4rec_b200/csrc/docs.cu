// docs.cu -- on-device document assembly: the reference's slow reader builds, per rating,
//   user document  = all train reviews of the user EXCEPT the one written for this item      (data.py:212-248 remove_overlap)
//   item document  = all train reviews of the item EXCEPT the one written by this user
//   this review    = that left-out review (training) or the held-out review (evaluation)
//   neighbour ids  = the item / user behind every kept review, padded to 10                   (data.py:277-282)
// as python lists, concatenated and padded to input_length (pad_and_join, data.py:174-210) or -- NARRE -- kept
// per review, padded to narre_num_words x narre_num_reviews (pad_only, data.py:146-172), and make_quick_data.py
// freezes the result into 24 KB of HDF5 per rating.  Here the train reviews live ONCE in HBM as CSR
// (token array + review offsets + per-user / per-item review lists) and a batch is assembled from
// (list id, index to leave out) per rating: one warp per rating, HBM-bound copies of <= T tokens.
#include "common.cuh"

namespace {
constexpr int THREADS = 256;

struct AssembleArgs {
  const int32_t* tok;        // tokens of all train reviews, back to back
  const int64_t* rev_off;    // [n_reviews + 1]
  const int64_t* ptr;        // [n_lists + 1]  reviews of list l: rev[ptr[l] .. ptr[l+1])
  const int32_t* rev;        // review ids in list order
  const int64_t* nb;         // neighbour id of every list entry (item of a user's review / user of an item's review)
  const int64_t* ids;        // [B] list id per rating (user id or item id)
  const int32_t* skip;       // [B] entry of the list to leave out, or NULL / -1 = none
  int64_t n_lists;
  int64_t B;
  int mode;                  // 0 = pad_and_join -> [B, T];  1 = pad_only -> [B, R, W]
  int T, R, W;
  int64_t nb_pad;
  int nbw;                   // neighbour list width (10)
  int64_t* out_docs;
  int64_t* out_nb;           // [B, nbw] or NULL
  int64_t* out_this;         // same shape as one document, or NULL
  const int32_t* this_tok;   // evaluation: held-out reviews as CSR rows this_row0 + b; NULL = take the skipped review
  const int64_t* this_off;
  int64_t this_row0;
};

// copies up to `room` tokens of review [a, a+n) to dst (lane-strided), returns the number taken
__device__ __forceinline__ int copy_tokens(const int32_t* __restrict__ tok, int64_t a, int64_t n, int room, int64_t* __restrict__ dst, int lane) {
  const int take = n < room ? (int)n : room;
  for (int t = lane; t < take; t += 32) dst[t] = (int64_t)__ldg(tok + a + t);
  return take;
}

__global__ void __launch_bounds__(THREADS) docs_assemble_kernel(const __grid_constant__ AssembleArgs A) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  const int doc_elems = A.mode == 0 ? A.T : A.R * A.W;
  for (int64_t b = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); b < A.B; b += warps) {
    const int64_t l = __ldg(A.ids + b);
    if (l < 0 || l >= A.n_lists) __trap();
    const int64_t lo = __ldg(A.ptr + l), hi = __ldg(A.ptr + l + 1);
    const int sk = A.skip ? __ldg(A.skip + b) : -1;
    if (sk >= hi - lo) __trap();
    const int64_t kept = (hi - lo) - (sk >= 0 ? 1 : 0);
    int64_t* doc = A.out_docs + b * (int64_t)doc_elems;

    if (A.mode == 0) {
      // concatenate the kept reviews until T tokens are out; 32 reviews' offsets are fetched at a time
      int pos = 0;
      for (int64_t j0 = 0; j0 < kept && pos < A.T; j0 += 32) {
        const int64_t j = j0 + lane;
        int64_t a = 0, n = 0;
        if (j < kept) {
          const int64_t k = lo + j + ((sk >= 0 && j >= sk) ? 1 : 0);
          const int32_t r = __ldg(A.rev + k);
          a = __ldg(A.rev_off + r);
          n = __ldg(A.rev_off + r + 1) - a;
        }
        const int cnt = (int)((kept - j0) < 32 ? (kept - j0) : 32);
        for (int q = 0; q < cnt && pos < A.T; ++q) {
          const int64_t aq = __shfl_sync(0xffffffffu, a, q);
          const int64_t nq = __shfl_sync(0xffffffffu, n, q);
          pos += copy_tokens(A.tok, aq, nq, A.T - pos, doc + pos, lane);
        }
      }
      for (int t = pos + lane; t < A.T; t += 32) doc[t] = 0;
    } else {
      for (int row = 0; row < A.R; ++row) {
        int64_t* dst = doc + row * A.W;
        int got = 0;
        if (row < kept) {
          const int64_t k = lo + row + ((sk >= 0 && row >= sk) ? 1 : 0);
          const int32_t r = __ldg(A.rev + k);
          const int64_t a = __ldg(A.rev_off + r);
          got = copy_tokens(A.tok, a, __ldg(A.rev_off + r + 1) - a, A.W, dst, lane);
        }
        for (int t = got + lane; t < A.W; t += 32) dst[t] = 0;
      }
    }

    if (A.out_nb) {
      for (int j = lane; j < A.nbw; j += 32) {
        int64_t v = A.nb_pad;
        if (j < kept) v = __ldg(A.nb + lo + j + ((sk >= 0 && j >= sk) ? 1 : 0));
        A.out_nb[b * A.nbw + j] = v;
      }
    }

    if (A.out_this) {
      int64_t* dst = A.out_this + b * (int64_t)doc_elems;
      const int width = A.mode == 0 ? A.T : A.W;       // pad_only keeps the single review in row 0
      int64_t a = 0, n = 0;
      if (A.this_tok) {
        a = __ldg(A.this_off + A.this_row0 + b);
        n = __ldg(A.this_off + A.this_row0 + b + 1) - a;
      } else if (sk >= 0) {
        const int32_t r = __ldg(A.rev + lo + sk);
        a = __ldg(A.rev_off + r);
        n = __ldg(A.rev_off + r + 1) - a;
      }
      const int got = copy_tokens(A.this_tok ? A.this_tok : A.tok, a, n, width, dst, lane);
      for (int t = got + lane; t < doc_elems; t += 32) dst[t] = 0;
    }
  }
}
}  // namespace

extern "C" int r4r_docs_assemble(const int32_t* tok, const int64_t* rev_off, const int64_t* ptr, const int32_t* rev,
                                 const int64_t* nb, int64_t n_lists, const int64_t* ids, const int32_t* skip, int64_t B,
                                 int mode, int T, int R, int W, int64_t nb_pad, int nbw,
                                 int64_t* out_docs, int64_t* out_nb, int64_t* out_this,
                                 const int32_t* this_tok, const int64_t* this_off, int64_t this_row0, void* stream) {
  R4R_REQUIRE(rev_off && ptr && rev && ids && out_docs, R4R_EINVAL, "docs_assemble: null pointer");
  R4R_REQUIRE(B >= 0 && n_lists > 0 && (mode == 0 || mode == 1), R4R_EINVAL, "docs_assemble: bad sizes / mode");
  R4R_REQUIRE(mode == 0 ? T > 0 : (R > 0 && W > 0), R4R_EINVAL, "docs_assemble: bad document shape");
  R4R_REQUIRE(out_nb == nullptr || (nb && nbw > 0), R4R_EINVAL, "docs_assemble: neighbour output without neighbour ids");
  R4R_REQUIRE(this_tok == nullptr || this_off, R4R_EINVAL, "docs_assemble: held-out reviews need offsets");
  if (B == 0) return 0;
  AssembleArgs A;
  A.tok = tok; A.rev_off = rev_off; A.ptr = ptr; A.rev = rev; A.nb = nb; A.ids = ids; A.skip = skip;
  A.n_lists = n_lists; A.B = B; A.mode = mode; A.T = T; A.R = R; A.W = W; A.nb_pad = nb_pad; A.nbw = nbw;
  A.out_docs = out_docs; A.out_nb = out_nb; A.out_this = out_this;
  A.this_tok = this_tok; A.this_off = this_off; A.this_row0 = this_row0;
  int64_t blocks = cdiv64(B, THREADS / 32);
  if (blocks > 148 * 16) blocks = 148 * 16;
  docs_assemble_kernel<<<(unsigned)blocks, THREADS, 0, as_stream(stream)>>>(A);
  R4R_CHECK_LAUNCH("docs_assemble");
  return 0;
}
