/* r4r_b200.h -- C ABI of the B200-native rating-prediction training hot path.
 *
 * Drop-in boundary for noveens/reviews4rec's training path (main.py + loss.py driving
 * pytorch_models/).  The reference is pure Python/PyTorch and has no FFI of its own; every entry
 * point below therefore replaces a *library op dispatched by PyTorch on the reference's behalf*
 * and cites the reference call site (file:line under /root/reference) it stands in for.
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 *     the parameter name ends in _host.  `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).  Nothing allocates device memory; callers own every buffer.
 *   - every function returns 0 on success, a negative R4R_E* code on argument errors, or a positive
 *     cudaError_t.  r4r_last_error() returns a thread-local message for the last failure.  The
 *     reference validates nothing (SURVEY.md 8b "error conventions"); the Python host layer turns
 *     non-zero returns into RuntimeError.
 *   - fp32 everywhere except the private fp16/bf16 "shadow" word table read by the tensor-core
 *     conv; ids are int64 exactly as the reference's LongTensors (data_fast.py:102-108).
 *   - launches are asynchronous on `stream`; no host synchronisation happens inside the library.
 */
#ifndef R4R_B200_H
#define R4R_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R4R_ABI_VERSION 5

#define R4R_EINVAL   (-1)   /* bad argument (null pointer, size out of supported range)          */
#define R4R_EUNSUP   (-2)   /* shape outside what the sm_100a kernels were built for             */
#define R4R_ENODEV   (-3)   /* no sm_100 device                                                  */

#define R4R_DT_F16   0
#define R4R_DT_BF16  1

int         r4r_abi_version(void);
const char* r4r_last_error(void);
/* name[64], returns sm count / compute capability of the current device. */
int         r4r_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes);

/* ---- a4: word-embedding gather --------------------------------------------------------------
 * out[i,:] = table[idx[i],:]            replaces nn.Embedding.forward / aten::index_select at
 * DeepCoNN.py:53-54, NARRE.py:95-96, TransNet.py:52-53,100-102.  Bit-exact copy. */
int r4r_word_gather_f32(const float* table, int64_t V, int E, const int64_t* idx, int64_t n,
                        float* out, void* stream);

/* Ragged -> padded documents: out[n, t] = tokens[offsets[n] + t] for t < offsets[n+1] - offsets[n], else
 * pad_id.  Rebuilds on the device the [N, T] int64 tensors the reference's fast reader ships padded from
 * host RAM (data_fast.py:99-109); the host side keeps only the tokens before the trailing padding run. */
int r4r_docs_expand(const int32_t* tokens, const int64_t* offsets, int64_t N, int T, int64_t pad_id, int64_t* out,
                    void* stream);

/* On-device document assembly.  The train reviews live once in device memory as CSR:
 * tok (all reviews back to back), rev_off [n_reviews+1], and per side (user or item) the review lists
 * ptr [n_lists+1] / rev (review ids in list order) / nb (the item -- resp. user -- behind every list entry).
 * For rating b of list ids[b], leaving out list entry skip[b] (-1 / NULL: none), this builds what the
 * reference's slow reader builds in Python (data.py:212-248 remove_overlap, :174-210 pad_and_join,
 * :146-172 pad_only, :277-282 neighbour padding):
 *   mode 0: out_docs[b, 0:T]      = kept reviews concatenated, cut to T, zero padded
 *   mode 1: out_docs[b, r, 0:W]   = kept review r cut / zero padded to W, r < R, missing reviews all zero
 *   out_nb[b, 0:nbw]              = nb of the kept reviews, padded with nb_pad, cut to nbw          (optional)
 *   out_this[b, ...]              = the left-out review as a document of the same shape, or -- this_tok
 *                                   given -- row this_row0 + b of the held-out review CSR           (optional) */
int r4r_docs_assemble(const int32_t* tok, const int64_t* rev_off, const int64_t* ptr, const int32_t* rev,
                      const int64_t* nb, int64_t n_lists, const int64_t* ids, const int32_t* skip, int64_t B,
                      int mode, int T, int R, int W, int64_t nb_pad, int nbw,
                      int64_t* out_docs, int64_t* out_nb, int64_t* out_this,
                      const int32_t* this_tok, const int64_t* this_off, int64_t this_row0, void* stream);

/* Private reduced-precision copy of the frozen word table (SURVEY.md finding 2):
 * shadow[v, 0:E] = cvt(table[v,:]), shadow[v, E:Epad] = 0.  Epad % 8 == 0, row stride = Epad.
 * `shadow` holds V+1 rows: row V is all zero (r4r_conv_pool_tc reads the conv's zero padding from it). */
int r4r_shadow_build(const float* table, int64_t V, int E, void* shadow, int Epad, int dtype,
                     void* stream);

/* ---- a5: TextCNN conv + ReLU + global max-pool, fused with the gather ------------------------
 * For doc n (row of idx[N,T]) and filter f:
 *   y[p] = sum_{j<3} sum_e Xpad[p+j, e] * W[f,0,j,e],  Xpad = 2 zero rows | table[idx[n,:]] | 2 zero rows
 *   pooled[n,f] = relu(max_p y[p] + bias[f]),  argmax[n,f] = first p attaining the max, p in [0,T+2)
 * replaces F.conv2d(padding=(2,0)) + F.relu + F.max_pool1d at common_pytorch_models.py:26-31
 * (Conv2d built at :14-17).  window size is the reference default 3.
 * r4r_conv_pool_simt: exact fp32 FMA arithmetic (parity mode).  keys_ws: N*F uint64 workspace. */
int r4r_conv_pool_simt(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T,
                       const float* conv_w, const float* conv_b, int F,
                       float* pooled, int32_t* argmax, uint64_t* keys_ws, void* stream);

/* Tensor-core path (tcgen05.mma kind::f16, fp32 accumulate in TMEM).
 * r4r_conv_pack_weights: conv_w [F,1,3,E] fp32 -> operand image `wpack` in the kernel's smem layout
 * (r4r_conv_wpack_bytes(E,F) bytes).  Re-run after every optimizer step that changes conv_w. */
int64_t r4r_conv_wpack_bytes(int E, int F);
int r4r_conv_pack_weights(const float* conv_w, int E, int F, void* wpack, int dtype, void* stream);
/* `ws`: r4r_conv_stream_ws_bytes(N, T) bytes of device scratch per launch.  The launch first lays the documents of every
 * persistent CTA pair end to end as one stream of conv windows (two zero rows between documents, shared by the window
 * that closes one and the window that opens the next), so the tensor cores run over tiles of 256 CONSECUTIVE windows
 * instead of rounding each document up to whole tiles; max / arg-max are taken per document over the windows it owns.
 * Results are independent of the launch partition (bit-identical for any batch split / order). */
int64_t r4r_conv_stream_ws_bytes(int64_t N, int T);
int r4r_conv_pool_tc(const void* shadow, int64_t V, int Epad, int E, int dtype,
                     const int64_t* idx, int64_t N, int T,
                     const void* wpack, const float* conv_b, int F,
                     float* pooled, int32_t* argmax,
                     const int32_t* doc_len, const int32_t* doc_order, void* ws, void* stream);

/* Work plan for r4r_conv_pool_tc (optional: pass NULL, NULL to process every row of every document).
 * The readers pad documents to T with one repeated token that is embedded like any other
 * (data.py:198-199, DeepCoNN.py:53-54); conv windows inside such a trailing run all produce the same
 * value and max_pool1d keeps the first, so a document whose rows s..T-1 are equal gives bit-identical
 * (pooled, argmax) when cut to doc_len = min(T, s+3) rows, with arg-max positions >= doc_len mapped
 * back by + (T - doc_len).  doc_order = documents by decreasing length in 256 classes of (T+2)/255 windows (load balance:
 * the conv launch deals them to its CTA pairs in rounds of alternating direction).
 * The order is a stable sort (deterministic: no global atomics).  ws: r4r_doc_plan_ws_bytes(N, T) bytes of scratch. */
int64_t r4r_doc_plan_ws_bytes(int64_t N, int T);
int r4r_doc_plan(const int64_t* idx, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                 void* stream);

/* Ragged documents (reviews4rec_b200/readers.py): document n = tokens[offsets[n] .. offsets[n+1]) followed by
 * pad_id up to T rows -- the same padded document the readers build (data.py:198-202), without ever
 * materialising it.  Identical results to the padded entry points on the expanded ids. */
int r4r_conv_pool_tc_ragged(const void* shadow, int64_t V, int Epad, int E, int dtype,
                            const int32_t* tokens, const int64_t* offsets, int64_t pad_id, int64_t N, int T,
                            const void* wpack, const float* conv_b, int F,
                            float* pooled, int32_t* argmax,
                            const int32_t* doc_len, const int32_t* doc_order, void* ws, void* stream);
/* doc_len[n] = min(T, offsets[n+1] - offsets[n] + 3): the padding run starts where the stored tokens end */
int r4r_doc_plan_ragged(const int64_t* offsets, int64_t N, int T, int32_t* doc_len, int32_t* doc_order, void* ws,
                        void* stream);
int r4r_conv_wgrad_argmax_h_ragged(const void* shadow, int64_t V, int Epad, int E, int dtype, const int32_t* tokens,
                                   const int64_t* offsets, int64_t pad_id, int64_t N, int T, const int32_t* argmax,
                                   const float* pooled, const float* gpooled, int F, float* dW, float* db, void* stream);

/* Number of persistent CTA pairs later r4r_conv_pool_tc launches use (0 = one per SM pair = all SMs).  Fewer pairs
 * leave SMs to kernels of a concurrent stream / graph branch (the prefetched sharded word lookup). */
int r4r_conv_set_clusters(int n);

/* Diagnostics: when `buf32_u64` (device, 32 x uint64) is non-NULL every later r4r_conv_pool_tc launch
 * writes the per-role cycle counters of its first CTA pair there (see conv_tc.cu); NULL turns it off. */
int r4r_conv_debug_profile(void* buf32_u64);

/* ---- K3: the whole DeepCoNN / DeepCoNN++ head in one forward and one backward kernel ---------------------------
 * forward : fc_t(pooled_t) (common_pytorch_models.py:33-37) -> dropout (Philox in-kernel, or keep masks handed in) -> cat
 *           (DeepCoNN.py:61) -> head 0: global_bias + TorchFM(cat) (DeepCoNN.py:64-66)  |  head 1: final MLP + user_bias[u] +
 *           item_bias[i] + global_bias (DeepCoNN.py:69-72) -> rating [, (rating - y)^2 and its sum: loss.py:7-11]
 * `ptrs` (25 device pointers, NULL where unused): pooled_u, pooled_i [N,F]; fc_u.w, fc_i.w [L,F]; fc_u.b, fc_i.b [L];
 *   fm.V [2L,K], fm.lin.w [2L], fm.lin.b [1]; final.0.w [L,2L], final.0.b [L], final.3.w [L], final.3.b [1]; gathered
 *   user_bias / item_bias values [N]; global_bias [1]; y [N]; keep masks uint8 [N,3L] (tests) ; int32 step counter;
 *   OUT rating [N], se [N], se_sum [1] (+=), cat [N,2L], hid [N,L], keep uint32 [N,3].
 * backward: `gptrs` (18): g_rating [N], g_se [N] (either may be NULL); OUT dpooled_u, dpooled_i [N,F]; ACCUMULATED d fc_u.w,
 *   d fc_i.w, d fc_u.b, d fc_i.b, d fm.V, d fm.lin.w, d fm.lin.b, d final.0.w, d final.0.b, d final.3.w, d final.3.b; OUT
 *   d user_bias values, d item_bias values [N]; ACCUMULATED d global_bias.  Also advances the step counter by one. */
int r4r_deepconn_head_fwd(const void* const* ptrs, int N, int F, int L, int K, int head, float p, uint64_t seed, void* stream);
int r4r_deepconn_head_bwd(const void* const* ptrs, const void* const* gptrs, int N, int F, int L, int K, int head, float p,
                          void* stream);

/* ---- a11: conv weight gradient through relu+max-pool (SURVEY.md finding 4) --------------------
 * dW[f,0,j,:] += sum_n gy[n,f] * Xpad[n, argmax[n,f]+j, :],  db[f] += sum_n gy[n,f]
 * with gy = gpooled * (pooled > 0).  Replaces autograd's convolution_backward + relu/max-pool
 * backward for TextCNN (main.py:59).  dW/db are ACCUMULATED into (zero them first). */
int r4r_conv_wgrad_argmax(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T,
                          const int32_t* argmax, const float* pooled, const float* gpooled, int F,
                          float* dW, float* db, void* stream);

/* fp32 refinement of r4r_conv_pool_tc ("f16r" / "bf16r" modes): pooled[n,f] = relu(b[f] + the fp32 conv value of the window
 * at argmax[n,f]), from the fp32 word table and filters (common_pytorch_models.py:29-31 at the selected position). */
int r4r_conv_refine(const float* table, int64_t V, int E, const int64_t* idx, int64_t N, int T, const int32_t* argmax,
                    const float* conv_w, const float* conv_b, int F, float* pooled, void* stream);

/* Same gradient from the half-precision shadow rows the tensor-core forward read (f16 / bf16 modes):
 * it is the exact gradient of what r4r_conv_pool_tc computed and halves the gather traffic. */
int r4r_conv_wgrad_argmax_h(const void* shadow, int64_t V, int Epad, int E, int dtype, const int64_t* idx, int64_t N,
                            int T, const int32_t* argmax, const float* pooled, const float* gpooled, int F,
                            float* dW, float* db, void* stream);

/* OPT-IN (not reference behaviour: the reference freezes the word table, DeepCoNN.py:15): gradient of the
 * conv w.r.t. the word table through relu + max-pool,
 *   gtable[idx[n, argmax[n,f] + j - 2], :] += gy[n,f] * conv_w[f,0,j,:]     (positions outside [0,T) skipped)
 * accumulated into the dense [V,E] gradient; per document, updates of the same row are summed first
 * (sorted segments) so that each (row, column) receives one atomic. */
int r4r_conv_dgrad_scatter(const int64_t* idx, int64_t N, int T, const int32_t* argmax, const float* pooled,
                           const float* gpooled, const float* conv_w, int F, int E, float* gtable, int64_t V,
                           void* stream);

/* ---- small dense layers of the heads ----------------------------------------------------------
 * y[n,o] = sum_i x[n,i] W[o,i] + b[o]   (nn.Linear: TextCNN.fc common_pytorch_models.py:19,37;
 * DeepCoNN.final DeepCoNN.py:21-26; NARRE scorers NARRE.py:24-43; TransNet project TransNet.py:17-21) */
int r4r_linear_fwd(const float* x, const float* W, const float* b, int64_t n, int in_f, int out_f,
                   float* y, void* stream);
/* dx = gy W (may be NULL); dW += gy^T x; db += sum_n gy  (dW/db accumulated, may be NULL) */
int r4r_linear_bwd(const float* x, const float* W, const float* gy, int64_t n, int in_f, int out_f,
                   float* dx, float* dW, float* db, void* stream);

/* ---- a6: TorchFM second-order interaction (common_pytorch_models.py:49-57) --------------------
 * out[n] = 0.5*(sum_k (xV)_k^2 - sum_k (x^2 V^2)_k) + x.w + b        x [n,nf], V [nf,k], w [nf] */
int r4r_fm_fwd(const float* x, const float* V, const float* w, const float* b, int64_t n, int nf, int k,
               float* out, void* stream);
int r4r_fm_bwd(const float* x, const float* V, const float* w, const float* gout, int64_t n, int nf, int k,
               float* dx, float* dV, float* dw, float* db, void* stream);

/* ---- a3: MSE (loss.py:7-11) -------------------------------------------------------------------
 * se[n] = (out[n]-y[n])^2 ; *sum_se += sum_n se[n] (device scalar, accumulated; may be NULL) */
int r4r_mse_fwd(const float* out, const float* y, int64_t n, float* se, float* sum_se, void* stream);
/* gout[n] = gse[n] * 2*(out[n]-y[n])   (gse = upstream grad of se, e.g. 1/B for the mean) */
int r4r_mse_bwd(const float* out, const float* y, const float* gse, int64_t n, float* gout, void* stream);

/* idx[r] = index of the first largest x[r, 0:c]: the top-1 candidate of eval.eval_ranking (eval.py:74-78) */
int r4r_rows_argmax(const float* x, int64_t n, int c, int64_t* idx, void* stream);

/* ---- a7-a10: id-embedding / bias row gathers and their gradient scatter ----------------------
 * out[i,:] = table[ids[i],:]   replaces nn.Embedding / Tensor.gather at MF.py:45-46,52-53,
 * NARRE.py:87-88,110-116, TransNet.py:108-109, DeepCoNN.py:70-71  (L = 1 for the bias vectors) */
int r4r_rows_gather(const float* table, int64_t R, int L, const int64_t* ids, int64_t n, float* out,
                    void* stream);
/* gtable[ids[i],:] += gout[i,:]  (dense gradient as autograd's embedding_dense_backward produces;
 * duplicates are combined inside each warp before one atomic per distinct row) */
int r4r_rows_scatter_add(const float* gout, const int64_t* ids, int64_t n, int L, float* gtable, int64_t R,
                         void* stream);

/* ---- a12: fused dense Adam (torch.optim.Adam as configured at main.py:94-96) -----------------
 * For each tensor t < nt:  g = grad + wd*p; m,v moments; p -= lr/bc1 * m/(sqrt(v)/sqrt(bc2)+eps).
 * Arrays of DEVICE pointers are passed from the HOST (p_host[t] etc.).  `step` is the 1-based step
 * count; if step_dev != NULL the kernel reads the count from device memory instead (CUDA graphs). */
int r4r_adam_step(int nt, float* const* p_host, const float* const* g_host, float* const* m_host,
                  float* const* v_host, const int64_t* numel_host, int step, const int32_t* step_dev,
                  float lr, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* *counter += 1 on the device: the step count of a CUDA-graph-captured optimizer (torch keeps
 * state['step'] as a device tensor for capturable=True; torch/optim/adam.py semantics). */
int r4r_counter_inc(int32_t* counter, void* stream);

/* ---- K8: row-sharded tables (new: the reference is single-process) -----------
 * Row r of a table lives on rank r % P at local row r / P.  The lookups keep the semantics of the
 * same nn.Embedding / Tensor.gather call sites as r4r_word_gather_f32 / r4r_rows_gather.
 * Request message (int64 words): for every peer q a block of 1+cap words: [n_q, local_row_0 .. ].
 * Row payloads are [q][cap][row_bytes]; request (q, j) comes back at slot q*cap + j. */
/* sets bit idx[i] of the presence bitmap flags[] (>= ceil(V/32) 32-bit words, all zero before the first call of a step) */
int r4r_shard_mark(const int64_t* idx, int64_t n, int64_t V, int32_t* flags, void* stream);
/* appends every flagged id to the request block of its owner (req[id % P] gets the owner-local row id / P; the
 * order inside a block is unspecified -- rows are placed back by id) and clears flags.  cap >= ceil(V/P). */
int r4r_shard_plan(int32_t* flags, int64_t V, int P, int64_t cap, int64_t* req, void* stream);
/* id tables: no de-duplication; pos[i] = slot of ids[i]; req is zeroed inside.  cap >= n. */
int r4r_shard_bucket(const int64_t* ids, int64_t n, int64_t R, int P, int64_t cap, int64_t* req, int64_t* pos,
                     void* stream);
/* owner side: out[q][j][:] = shard[rreq(q, j)][:] for j < n_q.  row_bytes % 4 == 0. */
int r4r_shard_serve(const void* shard, int64_t rows_local, int row_bytes, const int64_t* rreq, int P, int64_t cap,
                    void* out, void* stream);
/* fused gather + all-to-all: as r4r_shard_serve, but requester q's block is written to
 * out_ptrs_host[q] -- rank q's receive buffer (+ this rank's block offset) mapped over NVLink. */
int r4r_shard_serve_p2p(const void* shard, int64_t rows_local, int row_bytes, const int64_t* rreq, int P, int64_t cap,
                        void* const* out_ptrs_host, void* stream);
/* requester side of a word lookup: rows [P][cap][row_bytes] came back compact; cache[req[q][1+j]*P + q] = rows[q][j],
 * i.e. the per-step row cache [V (+1 zero row), row_bytes] is indexed by the ORIGINAL token id and the conv / wgrad
 * kernels read their usual token ids (nn.Embedding semantics over the full table, DeepCoNN.py:53-54). */
int r4r_shard_place(const void* rows, const int64_t* req, int P, int64_t cap, int row_bytes, void* cache, int64_t V,
                    void* stream);
/* owner side of the backward: gtable[rreq(q, j)][:] += scale * grads[q*cap + j][:] for j < n_q */
int r4r_shard_scatter_add(const float* grads, const int64_t* rreq, int P, int64_t cap, int L, float* gtable,
                          int64_t rows_local, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R4R_B200_H */
