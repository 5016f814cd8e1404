/* ttr_b200.h — C ABI of libttr_b200.so: the sm_100a (B200) hot path of
 * jpe17/TwoTowerMLRetrieval (GRU two-tower encode, triplet step, exact cosine top-k,
 * hybrid rerank).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below
 * names the reference call site it replaces (file:line in the reference tree).  The
 * binding a maintainer adds on the reference side is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_h`;
 *    the library never allocates or frees caller-visible memory;
 *  - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and never
 *    synchronise unless documented;
 *  - return value: 0 = OK, non-zero = error; ttr_last_error() gives the thread-local text;
 *  - no CPU fallback: on a machine without an sm_100 device every compute call fails.
 */
#ifndef TTR_B200_H
#define TTR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTR_OK 0
#define TTR_ERR_INVALID 1
#define TTR_ERR_CUDA 2
#define TTR_ERR_UNSUPPORTED 3

const char* ttr_last_error(void);
int ttr_version(void);
/* Number of SMs of the current device (grid sizing on the host side). */
int ttr_sm_count(int* out_h);

/* ---- sequence plan -------------------------------------------------------------------
 * Replaces `lengths = (x != 0).sum(dim=1).cpu()` + `pack_padded_sequence(enforce_sorted=
 * False)` (backend/model.py:52-57) without the device->host sync.
 * ids      int64 [B, T] right-padded with 0 (collate layout, backend/main.py:50-56)
 * lengths  int32 [B]     count_nonzero per row (quirk #1: the first `length` positions are
 *                        the sequence)
 * order    int32 [B]     sorted position -> row, lengths non-increasing
 * offsets  int32 [B+1]   token offset of each sorted position in the packed layout;
 *                        offsets[B] = total tokens
 * status   int32 [4]     [0] number of zero-length rows (reference raises RuntimeError),
 *                        [1] total tokens, [2] max length, [3] reserved
 * Requires T <= 8192. */
int ttr_seq_plan(const int64_t* ids, int B, int T, int32_t* lengths, int32_t* order,
                 int32_t* offsets, int32_t* status, void* stream);

/* ---- K1: embedding gather --------------------------------------------------------------
 * Replaces `self.embedding(x)` (backend/model.py:49).  Writes the packed token matrix
 * X[offsets[s] + t, 0:E] = table[ids[order[s], t], 0:E] for t < length.  `round_tf32` != 0
 * rounds to TF32 (round-to-nearest) because X only feeds the tensor-core projection; the value 2
 * makes X an fp16 [Mtok, E] matrix (E % 4 == 0) for the kind::f16 inference pipeline. */
int ttr_embed_gather(const int64_t* ids, int B, int T, const float* table, int64_t V, int E,
                     const int32_t* order, const int32_t* offsets, float* X, int round_tf32,
                     void* stream);

/* ---- K3: GRU input projection ----------------------------------------------------------
 * Replaces the `X W_ih^T + b_ih` half of `self.rnn(packed)` (backend/model.py:59-62) for
 * all timesteps and both directions at once: C[M, N] = A[M, K] * W[N, K]^T + bias[N].
 * tcgen05 (kind::tf32, fp32 accumulate in TMEM) + TMA.  The number of valid rows is read
 * on the device from *m_valid (may be NULL = m_bound); m_bound sizes the grid.
 * Requirements: K % 4 == 0, N % 8 == 0, A/W/C 16-byte aligned. */
int ttr_gemm_tf32_bias(const float* A, const float* W, const float* bias, float* C,
                       int m_bound, const int32_t* m_valid, int N, int K, void* stream);
/* Same contraction with fp16 storage: A16 fp16 [m_bound, K], W16 fp16 [N, K] (ttr_f32_to_f16 of W_ih),
 * bias fp32 [N], C16 fp16 [m_bound, N]; kind::f16 MMAs, fp32 accumulation.  K % 8 == 0, N % 8 == 0. */
int ttr_gemm_f16_bias(const void* A16, const void* W16, const float* bias, void* C16, int m_bound,
                      const int32_t* m_valid, int N, int K, void* stream);
/* dst[i] = fp16(src[i]) (round to nearest even), n elements. */
int ttr_f32_to_f16(const float* src, void* dst, int64_t n, void* stream);
/* Same contract, plain fp32 CUDA-core kernel.  Test-only reference for the tcgen05 path. */
int ttr_debug_gemm_fp32_bias(const float* A, const float* W, const float* bias, float* C,
                             int m_bound, const int32_t* m_valid, int N, int K, void* stream);
/* C[N, K] (+)= A[M, N]^T * Bm[M, K] over the valid rows: the weight-gradient contraction
 * dW = dG^T X of the GRU backward (autograd of backend/main.py:254).  fp32 split-K. */
int ttr_gemm_tn_fp32(const float* A, const float* Bm, float* C, int m_bound,
                     const int32_t* m_valid, int N, int K, int accumulate, void* stream);
/* Tensor-core version of the same contraction on column blocks of wider matrices:
 * C[N1, N2] (+)= A[0:m_valid, 0:N1]^T * Bm[0:m_valid, 0:N2], row pitches lda/ldb/ldc (multiples of 4).
 * tcgen05 kind::tf32 with MN-major operands, split-K, TMA reduce-add epilogue.  The caller must zero
 * rows [m_valid, round_up(m_valid, 32)) of both operands (ttr_zero_tail_rows). */
int ttr_gemm_tn_tf32(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int m_bound,
                     const int32_t* m_valid, int N1, int N2, int accumulate, void* stream);
/* rows [m_valid, min(m_bound, round_up(m_valid, 32))) of a row-major [m_bound, ld] matrix := 0 */
int ttr_zero_tail_rows(float* A, int m_bound, const int32_t* m_valid, int ld, void* stream);
/* C[M, K] = A[M, N] * W[N, K]  (dX = dG W_ih), tcgen05 tf32 is not needed for parity of
 * gradients; fp32 CUDA cores. */
int ttr_gemm_nn_fp32(const float* A, const float* W, float* C, int m_bound,
                     const int32_t* m_valid, int N, int K, int accumulate, void* stream);

/* ---- K4: GRU recurrence ----------------------------------------------------------------
 * Replaces the sequential half of `self.rnn(packed)` (backend/model.py:59-62), one layer,
 * all directions.  gate order r,z,n; h0 = 0; per-row final state captured at each row's own
 * last step (forward) / after position 0 (reverse).
 * gi      fp32 [Mtok, dirs*3H]  input projections incl. b_ih (from ttr_gemm_*_bias)
 * w_hh    fp32 [dirs, 3H, H];  b_hh fp32 [dirs, 3H]
 * y       fp32 [Mtok, dirs*H]   per-step outputs (NULL if not needed)
 * h_last  fp32 [B, dirs*H]      final states in ORIGINAL row order
 * saved   fp32 [Mtok, dirs, 4, H] (r, z, n, W_hn h + b_hn) for the backward pass, or NULL
 * H == 256 runs the register-resident 8-CTA-cluster kernel (fp32 CUDA cores, warp-shuffle
 * reductions); other H use a generic kernel. */
int ttr_gru_recurrence_fwd(const float* gi, const float* w_hh, const float* b_hh,
                           const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                           float* y, float* h_last, float* saved, void* stream);
/* Same contract with a caller-owned scratch buffer of ttr_gru_fwd_workspace_bytes(B, H, dirs)
 * bytes (uninitialised is fine; 0 for H != 256): H == 256 then runs the tcgen05 recurrence —
 * W_hh slice resident in shared memory as fp16, h exchanged between the 8 CTAs of a cluster by
 * multicast bulk copies through the (L2-resident) scratch, fp32 state and accumulation. */
int64_t ttr_gru_fwd_workspace_bytes(int B, int H, int dirs);
int ttr_gru_recurrence_fwd_ws(const float* gi, const float* w_hh, const float* b_hh,
                              const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                              float* y, float* h_last, float* saved, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* Inference pipeline in fp16 storage (same 11-bit significand as the tf32 operands; fp32 accumulation,
 * bias and recurrent state): gi16 fp16 [Mtok, dirs*3H] from ttr_gemm_f16_bias, y16 fp16 [Mtok, dirs*H]
 * (NULL for the last layer) feeds the next layer's projection.  H == 256 only; needs the workspace. */
int ttr_gru_recurrence_fwd_f16(const void* gi16, const float* w_hh, const float* b_hh,
                               const int32_t* order, const int32_t* offsets, int B, int H, int dirs,
                               void* y16, float* h_last, void* workspace, int64_t workspace_bytes,
                               void* stream);
/* BPTT of one layer.  In: dy [Mtok, dirs*H] (grad wrt per-step outputs, NULL = zero),
 * dh_last [B, dirs*H] (grad wrt final states, original row order, NULL = zero), y (this
 * layer's forward outputs), saved.  Out: dgi [Mtok, dirs*3H] (grad wrt gi, also the b_ih
 * gradient rows) and dgh [Mtok, dirs*3H] (grad wrt W_hh h + b_hh). */
int ttr_gru_recurrence_bwd(const float* dy, const float* dh_last, const float* y,
                           const float* saved, const float* w_hh, const int32_t* order,
                           const int32_t* offsets, int B, int H, int dirs, float* dgi, float* dgh,
                           void* stream);
/* Same contract with a caller-owned scratch buffer of ttr_gru_bwd_workspace_bytes(B, H, dirs) bytes
 * (0 for H != 256; contents irrelevant): H == 256 then runs the tcgen05 BPTT — d(gh) W_hh as a split-K
 * product over the 8 CTAs of a cluster (fp16 hi/lo planes of the scaled gate gradients, fp16 W_hh, fp32
 * accumulation), partials reduce-added into the scratch by the copy engine.  Debug bit 23 keeps the fp32
 * CUDA-core kernel.  db_ih / db_hh (optional, fp32 [dirs*3H]) RECEIVE += the bias gradients (column sums
 * of dgi / dgh over the valid tokens; m_bound / m_valid as in ttr_colsum): fused into the tcgen05 kernel,
 * a ttr_colsum call after the others. */
int64_t ttr_gru_bwd_workspace_bytes(int B, int H, int dirs);
int ttr_gru_recurrence_bwd_ws(const float* dy, const float* dh_last, const float* y,
                              const float* saved, const float* w_hh, const int32_t* order,
                              const int32_t* offsets, int B, int H, int dirs, float* dgi, float* dgh,
                              void* workspace, int64_t workspace_bytes, int m_bound,
                              const int32_t* m_valid, float* db_ih, float* db_hh, void* stream);
/* dW_hh[dir] (+)= dgh[:, dir]^T h_prev[:, dir], where h_prev is y shifted by one step inside
 * each row (0 at a row's first step).  hprev_ws: scratch fp32 [m_bound, dirs*H].
 * Out: dw_hh [dirs, 3H, H].  (Bias gradients are column sums: ttr_colsum.)  Runs ttr_gemm_tn_tf32:
 * the caller zeroes the tail rows of dgh first; H and 3H must be multiples of 4. */
int ttr_gru_whh_grad(const float* dgh, const float* y, const int32_t* offsets, int B, int H,
                     int dirs, int m_bound, float* hprev_ws, float* dw_hh, int accumulate,
                     void* stream);
/* Embedding gradient for a trainable table (pretrained_embeddings=None, backend/model.py:24-27):
 * dtable[ids] += dX over the packed tokens; id 0 (padding_idx) gets none. */
int ttr_embed_scatter_grad(const int64_t* ids, int B, int T, int64_t V, int E,
                           const int32_t* order, const int32_t* offsets, const float* dX,
                           float* dtable, void* stream);
/* out[N] (+)= column sums of A[m_valid, N] (b_ih / b_hh / projection-bias gradients). */
int ttr_colsum(const float* A, int m_bound, const int32_t* m_valid, int N, float* out,
               int accumulate, void* stream);
/* Rank of document target[q] among all N documents for query q (1 = best; -1 if target is out of
 * range): 1 + #{j : s_qj > s_qt or (s_qj == s_qt and j < t)} — the position `torch.sort(descending)`
 * + search gives in BatchEvaluator.evaluate (backend/evaluators.py:49-73), from one streaming pass
 * without the [B, N] matrix.  Q fp32 [B, D], docs fp32 [N, D], D % 4 == 0; score_out (optional)
 * receives s_qt. */
int ttr_positive_rank(const float* Q, const float* docs, const int64_t* target, int B, int64_t N, int D,
                      int32_t* rank_out, float* score_out, void* stream);
/* HOST helper (no device work, no stream): ragged token rows -> right-padded [n_rows, T] int64, zero fill —
 * `pad_sequence(batch_first=True, padding_value=0)` of the reference collate (backend/main.py:50-56) — written into
 * a caller-owned (pinned) buffer; output row r = ragged row rows[r] (flat ids, start offsets, lengths). */
int ttr_pack_padded_i64(const int64_t* flat, const int64_t* starts, const int64_t* lengths,
                        const int64_t* rows, int64_t n_rows, int64_t T, int64_t* out);
/* The same, also returning the number of non-zero ids copied (the packed token count the device plan will find:
 * backend/model.py:52, `lengths = (x != 0).sum(1)`) and the number of rows without any (for which the reference's
 * pack_padded_sequence raises, model.py:57). */
int ttr_pack_padded_count_i64(const int64_t* flat, const int64_t* starts, const int64_t* lengths,
                              const int64_t* rows, int64_t n_rows, int64_t T, int64_t* out,
                              int64_t* nnz_total, int64_t* zero_rows);
/* Test/debug switches: bit0 = force the generic (any-H) GRU kernels, bit1 = encode TMA maps
 * as FLOAT32 instead of TFLOAT32 (hardware truncation instead of round-to-nearest),
 * bit2 = force the CUDA-core streaming scorer for every batch size, bit8 = no sample pass,
 * bit9 = select-merge re-reads candidates from L2 instead of staging them in shared memory,
 * bit10 = H=256 GRU forward on the fp32 CUDA-core cluster kernel instead of the tcgen05 one,
 * bit11/18/19 = projection-GEMM timing experiments (no stores / no staging / no weight-stationary variant),
 * bits12-17 = sample tiles per SM override, bit20 = per-CTA entry/exit times into the trace buffer,
 * bit21 = never fuse the sample pass into the main scorer launch, bit22 = always fuse it,
 * bit24 = 16-tile id segments in the tcgen05 scorer (tests cross many segment boundaries on small corpora),
 * bit25 = scorer epilogue without the max-tree fast reject (A/B timing),
 * bit26 = CTA-pair scorer: swap which CTA loads which half of a 64-document tile (bring-up switch; breaks results),
 * bit27 = one CTA per query tile for B > 128 (the round-1 layout) instead of CTA pairs,
 * bit28 = CTA-pair scorer: cta_group::2 TMA loads counted on the leader's barrier instead of plain loads + forwarding,
 * bit30 = fp16 projection GEMM for K > 256 on the single-CTA 128 x 128 kernel instead of CTA pairs (256 x 256 tiles),
 * bit29 = tcgen05 scorer without the screening warps (every tile goes through the list-keeping warps, as in round 1). */
int ttr_debug_set_flags(int flags);
int ttr_debug_get_flags(int* out);
/* Diagnostic: number of 8-CTA clusters of the tcgen05 recurrence the device holds at once. */
int ttr_debug_gru_tc_max_clusters(int* out);
/* Diagnostic: device buffer (8 * 256 int64) receiving the pipeline timeline (SM clock at five
 * events per document tile) of CTA (0,0) of the tcgen05 scorer; NULL switches it off. */
int ttr_debug_set_trace(long long* trace);

/* ---- K5/K6: projection + L2 normalise --------------------------------------------------
 * Replaces `cat(h_n[-2], h_n[-1])` -> `self.projection` -> `F.normalize(p=2, dim=1)`
 * (backend/model.py:65-74).  h_cat [B, IN]; w [H, IN] or NULL (unidirectional: IN == H, no
 * projection); raw_out [B, H] pre-normalisation values (NULL unless training);
 * normalize != 0 applies v / max(||v||, 1e-12). */
int ttr_proj_l2norm_fwd(const float* h_cat, const float* w, const float* bias, int B, int IN,
                        int H, int normalize, float* out, float* raw_out, void* stream);
/* Backward of the normalise: d_raw = d(out)/d(raw)^T d_out.  The projection's own backward is
 * two small GEMMs (ttr_gemm_nn_fp32 / ttr_gemm_tn_fp32) plus ttr_colsum. */
int ttr_l2norm_bwd(const float* d_out, const float* raw, int B, int H, int normalize,
                   float* d_raw, void* stream);

/* ---- K7: cosine triplet loss -----------------------------------------------------------
 * Replaces `triplet_loss_cosine` (backend/model.py:109-114): loss = mean(max(0, cos(q,n) -
 * cos(q,p) + margin)), F.cosine_similarity eps 1e-8.  loss_out fp32 [1]; the backward
 * writes dq, dp, dn = d loss / d inputs times *dloss (device scalar, NULL = 1). */
int ttr_triplet_fwd(const float* q, const float* p, const float* n, int B, int H, float margin,
                    float* loss_out, void* stream);
int ttr_triplet_bwd(const float* q, const float* p, const float* n, int B, int H, float margin,
                    const float* dloss, float* dq, float* dp, float* dn, void* stream);
/* `TwoTowerTrainer.compute_batch_metrics` (backend/trainer.py:38-55): out fp32 [5] =
 * accuracy, similarity_gap, magnitude, pos_similarity, neg_similarity. */
int ttr_batch_metrics(const float* q, const float* p, const float* n, int B, int H, float* out,
                      void* stream);

/* ---- K9/K10: clip_grad_norm_ + Adam ----------------------------------------------------
 * Replaces `clip_grad_norm_(params, max_norm)` + `optimizer.step()` (backend/main.py:257-259)
 * over one flat fp32 bucket.  grad_scale multiplies the gradients first (1/world after an
 * all-reduce sum).  max_norm <= 0 disables clipping (backend/trainer.py:101-110 variant).
 * norm_out fp32 [1] receives the pre-clip total norm.  step is 1-based. */
int ttr_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  int64_t n, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                  float eps, int step, float* norm_out, float* workspace /* >= 1024 floats */,
                  void* stream);

/* ---- K11/K12: exact cosine top-k -------------------------------------------------------
 * Replaces `torch.matmul(q, D.t())` + `torch.topk(sim, k)` (backend/evaluators.py:185-186,
 * :50/:62, :269-272; backend/trainer.py:62-65) and ChromaDB's ANN `collection.query(
 * n_results=50)` (frontend/main.py:153-156) with an exact fused scan that never
 * materialises [B, N].
 * Q fp32 [B, D]; docs fp32 [N, D] row-major (the document_embeddings.npy layout,
 * backend/main.py:138); D must be 256 for the fused kernels; k <= 64.
 * B <= 4: CUDA-core streaming kernel (fp32 exact products, HBM-bound);
 * B  > 4: tcgen05 kind::tf32 kernel, 128 queries per pass (scores within 1e-3 relative).
 * row_offset is added to the emitted indices (global ids of a row shard).
 * out_scores fp32 [B, k] descending; out_idx int64 [B, k]; ties: lower index first.
 * workspace: ttr_score_topk_workspace_bytes(B, N, k) bytes. */
int64_t ttr_score_topk_workspace_bytes(int B, int64_t N, int k);
int ttr_score_topk(const float* Q, int B, const float* docs, int64_t N, int D, int k,
                   int64_t row_offset, float* out_scores, int64_t* out_idx, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* Merge P candidate lists per query (rank shards after the all-gather, or CTA partials):
 * cand_scores fp32 [P, B, kin], cand_idx int64 [P, B, kin] -> top-k, same ordering rule. */
int ttr_topk_merge(const float* cand_scores, const int64_t* cand_idx, int P, int B, int kin,
                   int k, float* out_scores, int64_t* out_idx, void* stream);

/* Same merge, but the P candidate lists are read IN PLACE from peer GPUs (symmetric-memory
 * buffers mapped over NVLink): peer_*_h are HOST arrays of P device addresses, each pointing to
 * that rank's [B, kin] scores (fp32) / global ids (int64, -1 = empty) / optional TF-IDF payload
 * (fp64, array may be NULL).  The exchange and the merge are one kernel (no all-gather); the
 * caller provides the cross-rank barrier before the launch.  P <= 16. */
int ttr_topk_merge_peers(const uint64_t* peer_scores_h, const uint64_t* peer_idx_h,
                         const uint64_t* peer_tfidf_h, int P, int B, int kin, int k,
                         float* out_scores, int64_t* out_idx, double* out_tfidf, void* stream);

/* The same peer-memory merge with the cross-rank barrier inside the kernel (one launch per exchange):
 * peer_flags_h is a HOST array of P device addresses, entry p = rank p's flag array (>= P uint32 slots in
 * symmetric memory, zeroed once before the first step).  Rank `my_rank` release-stores `step` into slot
 * my_rank of every peer's array, waits until its own array shows `step` from every peer, then merges
 * the peers' [B, kin] lists in place.  `step` must increase by one per call on every rank; the caller
 * alternates two list buffers by step parity (a peer may still be reading the lists of step s-1 while
 * this rank writes those of step s).  Replaces the all-gather named in SURVEY.md 8(e). */
int ttr_topk_exchange_merge(const uint64_t* peer_scores_h, const uint64_t* peer_idx_h,
                            const uint64_t* peer_tfidf_h, const uint64_t* peer_flags_h, int P, int my_rank,
                            uint32_t step, int B, int kin, int k, float* out_scores, int64_t* out_idx,
                            double* out_tfidf, void* stream);

/* ---- K14: hybrid rerank ----------------------------------------------------------------
 * Replaces frontend/main.py:158-198: semantic = 2*cos-1 (space=0, Chroma default squared-L2)
 * or cos (space=1); tfidf = <doc TF-IDF row, query TF-IDF row> (L2-normalised CSR rows);
 * final = alpha*semantic + (1-alpha)*tfidf; stable descending sort; first `top_n`.
 * cand_idx int64 [B, kc] LOCAL row ids into the CSR (idx - csr_row_offset), cand_cos fp32;
 * doc CSR: indptr int64, indices int32 (sorted per row), data fp64; query CSR likewise
 * (q_indptr int64 [B+1]).  tfidf_in: optional fp64 [B, kc] precomputed TF-IDF scores (used
 * after a cross-rank gather); if NULL they are computed from the CSR.
 * q_sqnorm fp64 [B] / d_sqnorm fp64 [B, kc] (either may be NULL = 1.0): squared norms of the queries and of
 * the candidates; when one is given, space 0 evaluates the general 1 - (|q|^2 + |d|^2 - 2 q.d) (a token-less
 * query is the zero vector -> semantic = 1 - |d|^2 = 0, frontend/main.py:152-162) instead of the unit-norm
 * shortcut 2*cos - 1.
 * Outputs (fp64 [B, top_n] each): final, semantic, tfidf; out_pos int32 [B, top_n] = position
 * in the candidate list. */
int ttr_hybrid_rerank(const int64_t* cand_idx, const float* cand_cos, int B, int kc,
                      int64_t csr_row_offset, const int64_t* indptr, const int32_t* indices,
                      const double* data, const int64_t* q_indptr, const int32_t* q_indices,
                      const double* q_data, const double* tfidf_in, const double* q_sqnorm,
                      const double* d_sqnorm, double alpha, int space,
                      int top_n, double* out_final, double* out_sem, double* out_tfidf,
                      int32_t* out_pos, void* stream);
/* TF-IDF scores only, for candidates owned by this shard (others get 0):
 * out fp64 [B, kc]. */
int ttr_tfidf_candidates(const int64_t* cand_idx, int B, int kc, int64_t csr_row_offset,
                         int64_t csr_rows, const int64_t* indptr, const int32_t* indices,
                         const double* data, const int64_t* q_indptr, const int32_t* q_indices,
                         const double* q_data, double* out, void* stream);

/* Inter-layer dropout of nn.GRU in train mode (backend/model.py:35, DROPOUT in backend/config.json):
 * mask[i] = (u_i >= p) / (1 - p) with u_i from a counter-based generator of (seed, i);
 * out = y * mask.  The mask is kept for the backward pass and can be exported for oracle replay. */
int ttr_dropout(const float* y, int64_t n, float p, uint64_t seed, float* out, float* mask, void* stream);

/* ---- corpus-wide hybrid search ----------------------------------------------------------
 * Replaces `SimpleHybridRetriever.search` (backend/simple_hybrid.py:45-66) for ONE query:
 * combined[i] = float32(alpha * cos(q, docs[i])) + (1 - alpha) * <tfidf row i, query tfidf row>
 * over all N documents (sklearn cosine_similarity semantics: both sides normalised, zero rows
 * give 0), then the top k in `np.argsort(combined)[::-1]` order (ties: HIGHER index first).
 * q fp32 [D] (device), q_norm = ||q||; docs fp32 [N, D]; doc CSR as in ttr_hybrid_rerank; the
 * query's TF-IDF row as (q_idx int32 sorted, q_val fp64, q_nnz).  out_scores fp64 [k], out_idx
 * int64 [k]; combined_out optional fp64 [N]; workspace ttr_blend_topk_workspace_bytes(k) bytes. */
int64_t ttr_blend_topk_workspace_bytes(int k);
int ttr_blend_topk(const float* q, float q_norm, const float* docs, int64_t N, int D,
                   const int64_t* indptr, const int32_t* indices, const double* data,
                   const int32_t* q_idx, const double* q_val, int q_nnz, double alpha, int k,
                   double* out_scores, int64_t* out_idx, double* combined_out, void* workspace,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTR_B200_H */
