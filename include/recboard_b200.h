/* recboard_b200 -- C ABI of the B200-native full-catalog scoring path.
 *
 * The reference (MTandHJ/RecBoard) has no FFI layer: its hot path is a handful of PyTorch
 * lines repeated in every model file.  Each entry point below names the reference lines it
 * replaces (paths relative to the reference root); INTEGRATION.md shows the Python binding a
 * RecBoard maintainer adds (ctypes, recboard_b200/_lib.py).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch); kernels never allocate,
 *     free or retain them.  All work is enqueued on `stream`; no host synchronisation inside.
 *   - Return value: 0 = ok, <0 = argument error (RB_E_*), >0 = cudaError_t.  Nothing throws or
 *     exits; rb_last_error() returns a thread-local message for the last failure.
 *   - dtype: storage type of U / W / table.  mode: arithmetic of the contraction.
 *   - Matrices are row-major and contiguous: U (M,d), W (N,d) ("TN" GEMM, both K-major).
 *   - d must be a multiple of 8; bf16 mode supports d <= 256 (fused CE passes included),
 *     fp32x3 mode supports d <= 128.  Base pointers must be 16-byte aligned.
 *   - There is no CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef RECBOARD_B200_H
#define RECBOARD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* rb_stream_t; /* == cudaStream_t */

enum { RB_DTYPE_F32 = 0, RB_DTYPE_BF16 = 1 };
enum { RB_MODE_BF16 = 0,   /* bf16 operands, fp32 accumulate (tcgen05 kind::f16)               */
       RB_MODE_FP32X3 = 1  /* fp32 operands split hi/lo, 3 TF32 MMAs (error ~1e-6, "fp32 parity") */ };
enum { RB_OP_SCATTER_ADD = 0, RB_OP_SCORE_DENSE = 1, RB_OP_CE_FWD = 2, RB_OP_CE_BWD = 3,
       RB_OP_TOPK_EVAL = 4 };
enum { RB_E_ARG = -1, RB_E_ALIGN = -2, RB_E_UNSUPPORTED = -3, RB_E_WORKSPACE = -4, RB_E_NODEVICE = -5 };

#define RB_MASKED_SCORE (-1e23f) /* UniSRec/main.py:413 */

/* out[i,:] = table[idx[i],:]                      replaces `self.Item.embeddings(seqs)`
 * (SASRec/main.py:183, BERT4Rec/main.py:168, GRU4Rec/main.py:138, HSTU/main.py:171) and the
 * row gathers `itemEmbds[...]`, `userEmbds[users]` (MF-BPR/main.py:84-86,102; LightGCN/main.py:91-93,118). */
int rb_gather_rows(const void* table, const int64_t* idx, void* out, int64_t n_idx, int64_t n_rows,
                   int d, int dtype, rb_stream_t stream);

/* out[i,:] = x[i,:] / max(||x[i,:]||_2, eps); inv_norm[i] (nullable) = 1/max(norm, eps).  Replaces
 * `F.normalize(self.Item.embeddings.weight[NUM_PADS:], dim=-1)` and the user-side normalisation
 * (HSTU/main.py:180-184): one HBM read + one write, output optionally cast to bf16 (the operand copy of
 * the cosine scoring sweep).  d % 8 == 0, d <= 1024; in/out may alias when the dtypes match. */
int rb_normalize_rows(const void* x, void* out, float* inv_norm, int64_t n_rows, int d, int in_dtype,
                      int out_dtype, float eps, rb_stream_t stream);

/* grad_table[idx[i],:] += grad_out[i,:], row `padding_idx` untouched (pass -1 for none).
 * Deterministic (stable radix sort by row + in-order segment sums).  Replaces the autograd of the
 * gather above = ATen embedding_dense_backward, triggered by `loss.backward()` (SASRec/main.py:249). */
int rb_scatter_add_rows(const void* grad_out, const int64_t* idx, float* grad_table, int64_t n_idx,
                        int64_t n_rows, int d, int dtype, int64_t padding_idx, void* ws,
                        size_t ws_bytes, rb_stream_t stream);

/* Same, accumulating into a table gradient of either dtype (table_dtype = RB_DTYPE_F32 | RB_DTYPE_BF16): the
 * embedding gradient of a bf16 parameter is added in place -- read, fp32 add, one rounding per touched row -- instead
 * of through a dense fp32 (n_rows, d) matrix and a cast pass (what ATen's embedding_dense_backward + autograd's
 * accumulation do for the reference, SASRec/main.py:183,249). */
int rb_scatter_add_rows_into(const void* grad_out, const int64_t* idx, void* grad_table, int64_t n_idx,
                             int64_t n_rows, int d, int dtype, int table_dtype, int64_t padding_idx, void* ws,
                             size_t ws_bytes, rb_stream_t stream);

/* S[m,k] = scale * <U[m,:], table[idx[m,k],:]> (fp32), ids outside [0,n_rows) score 0.  The gathered
 * (M,K,d) tensor is never materialised.  Replaces the gather + row-wise dot of every sampled / pool path:
 * `itemEmbds[data[IUnseen]]` + einsum("BD,BKD->BK") (recommend_from_pool, SASRec/main.py:230-236,
 * MF-BPR/main.py:106-109); `itemEmbds[cat(pos,negs)]` + einsum("MD,MKD->MK") (sampled softmax,
 * HSTU/main.py:192-197); the BPR / BCE positive and negative logits (SASRec/main.py:203-206,
 * MF-BPR/main.py:84-91).  Rows must be a multiple of 16 bytes. */
int rb_gather_dot(const void* U, const void* table, const int64_t* idx, float scale, float* S, int64_t M,
                  int64_t K, int64_t n_rows, int d, int dtype, rb_stream_t stream);

/* Backward of rb_gather_dot for an upstream gradient G (M,K) fp32:
 *   dU (M,d)          = scale * sum_k G[m,k] * table[idx[m,k],:]      (nullable)
 *   dTable (n_rows,d) += scale * G[m,k] * U[m,:] at row idx[m,k]       (nullable; row padding_idx skipped;
 *                        sorted segment sums => deterministic; workspace RB_OP_SCATTER_ADD with nnz = M*K)
 * Replaces the autograd of the lines above (index backward + bmm backward). */
int rb_gather_dot_bwd(const void* U, const void* table, const int64_t* idx, const float* G, float scale,
                      float* dU, float* dTable, int64_t M, int64_t K, int64_t n_rows, int d, int dtype,
                      int64_t padding_idx, void* ws, size_t ws_bytes, rb_stream_t stream);

/* Y = A X for a CSR matrix A (n_rows x n_cols, int64 indices, fp32 values) and dense fp32 X (n_cols, d);
 * optionally acc += beta * Y in the same pass (Y or acc may be NULL, not both).  Rows are summed in
 * index order (deterministic).  Replaces the LightGCN propagation step
 * `allEmbds = self.Adj @ allEmbds; avgEmbds += allEmbds / (self.num_layers + 1)` (LightGCN/main.py:83-85),
 * the step in front of `reset_ranking_buffers`.  d % 4 == 0. */
int rb_spmm_csr(const int64_t* crow, const int64_t* col, const float* val, const float* X, float* Y,
                float* acc, float beta, int64_t n_rows, int64_t n_cols, int d, rb_stream_t stream);

/* ---- a2 over a ROW-SHARDED table on the GPUs of one box (the input side of `self.Item.embeddings(seqs)`,
 * SASRec/main.py:183, when the table is split by rows across ranks): every rank reads the rows where they live --
 * its own shard or a peer's, mapped through CUDA IPC, loaded over NVLink -- instead of gathering zeros for foreign ids
 * and all-reducing the replicated activations.
 *   rb_ipc_export        owner: handle (64 bytes) of the allocation holding `ptr` + the offset of `ptr` inside it
 *   rb_ipc_open          peer (its own device current): maps the allocation, enables peer access; *ptr_out = the
 *                        owner's `ptr` in this process, *base_out = what rb_ipc_close takes
 *   rb_gather_rows_peers out[i,:] = shard r's row (idx[i] - starts[r]) for starts[r] <= idx[i] < starts[r+1];
 *                        ids outside [starts[0], starts[n_shards]) give zero rows.  shard_ptrs / starts are HOST
 *                        arrays (n_shards <= 16 pointers, n_shards + 1 bounds). */
int rb_ipc_export(const void* ptr, void* handle64, int64_t* offset);
int rb_ipc_open(const void* handle64, int64_t offset, void** ptr_out, void** base_out);
int rb_ipc_close(void* base);
int rb_gather_rows_peers(const void* const* shard_ptrs, const int64_t* starts, int n_shards, const int64_t* idx, void* out,
                         int64_t n_idx, int d, int dtype, rb_stream_t stream);

/* Device-side row compaction: row_index[k] = position of the k-th non-zero byte of mask[0..n) for k < *count, -1
 * beyond; *count = number of non-zero bytes.  Replaces the boolean indexing `userEmbds[indices]` /
 * `positives[indices]` (SASRec/main.py:199-200), `fc(userEmbds)[masks]` (BERT4Rec/main.py:181), whose nonzero()
 * waits for the count on the host every step: gather the rows with rb_gather_rows(row_index) (entries -1 give zero
 * rows) and hand `count` to the rb_ce_* entries as m_dev. */
int rb_compact_index(const void* mask, int64_t n, int64_t* row_index, int32_t* count, rb_stream_t stream);

/* S = scale * U W^T + bias  (M,N) fp32.          replaces `torch.einsum("BD,ND->BN", ...)`
 * (SASRec/main.py:228; MF-BPR/main.py:104; LightGCN/main.py:120; HSTU/main.py:209) and
 * `self.fc(userEmbds)` (BERT4Rec/main.py:189).  Compatibility path: it materialises (M,N). */
int rb_score_dense(const void* U, const void* W, const float* bias, float scale, float* S,
                   int64_t M, int64_t N, int d, int dtype, int mode, void* ws, size_t ws_bytes,
                   rb_stream_t stream);

/* Fused full-catalog cross-entropy, forward pass.  Per-row softmax statistics of
 * S = scale*U W^T + bias over this shard's N items, without materialising S:
 *   row_max[i] ~ max_j S_ij (any reference within 2^16 of it), row_sumexp[i] = sum_j exp(S_ij - row_max[i]),
 *   label_logit[i] = S_i,(labels[i]-label_base) if that column lies in [0,N) else 0 (exact fp32 dot),
 *   dU_unnorm[i,:] = sum_j exp(S_ij - row_max[i]) W_j   (nullable; the same sweep feeds a second MMA,
 *                    so the gradient wrt U costs no extra pass -- bf16 mode, d <= 256, scale > 0).
 * loss = mean(row_max + log(row_sumexp) - label_logit) replaces
 * `einsum("MD,ND->MN")` + `self.criterion(logits, labels)` (SASRec/main.py:217-219,
 * GRU4Rec/main.py:175-178, BERT4Rec/main.py:181-182).  Row-sharded tables: merge
 * (max, sumexp, label_logit) across ranks.
 *
 * m_dev (nullable, every rb_ce_* entry): a DEVICE int32 holding how many of the M query rows exist.  M is then the
 * capacity the caller sized U / labels / the outputs for; rows >= *m_dev get neutral statistics (max 0, sumexp 1,
 * label_logit 0: lse = 0), zero dU, contribute nothing to dW / dbias, and whole tiles beyond the count are skipped.
 * This is what lets the caller compact the non-padding positions (`userEmbds[indices]`, SASRec/main.py:199-200;
 * `fc(userEmbds)[masks]`, BERT4Rec/main.py:181) on the device without reading the count back (no nonzero() sync). */
int rb_ce_fwd(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
              int64_t label_base, int64_t M, int64_t N, int d, int dtype, int mode, float* row_max,
              float* row_sumexp, float* label_logit, float* dU_unnorm, const int32_t* m_dev, void* ws,
              size_t ws_bytes, rb_stream_t stream);

/* dU (M,d) = g*scale*( dU_unnorm * exp(row_max - lse) - [label in shard] W_label ): this shard's
 * piece of the CE gradient wrt U from rb_ce_fwd's accumulator and the GLOBAL lse; pieces of
 * different shards add up (one all-reduce).  g = grad_scale * (grad_scale_dev ? *grad_scale_dev : 1).
 * Part of the autograd of SASRec/main.py:217-219 (`loss.backward()`, :249). */
int rb_ce_du_finish(const float* dU_unnorm, const float* row_max, const float* lse, const void* W,
                    const int64_t* labels, int64_t label_base, float scale, float grad_scale,
                    const float* grad_scale_dev, int64_t M, int64_t N, int d, int dtype, float* dU,
                    const int32_t* m_dev, rb_stream_t stream);

/* Gradients of g * sum_i (lse_i - S_i,label_i) given the GLOBAL lse (natural log), where
 * g = grad_scale * (grad_scale_dev ? *grad_scale_dev : 1)  (a device scalar lets autograd's
 * grad_output flow in without a host read):
 *   dW (N,d)  = g * scale * (softmax - onehot)^T U    (nullable)
 *   dbias (N) = g * column sums of (softmax - onehot) (nullable; needs dW)
 *   dU (M,d)  = g * scale * (softmax - onehot) W      (this shard's partial; nullable.  Pass NULL when
 *               rb_ce_fwd produced dU_unnorm -- rb_ce_du_finish is then all that is needed; otherwise
 *               the forward sweep is re-run here)
 * Recomputes S tile by tile; the softmax tile lives only in TMEM; the one-hot is applied exactly in
 * fp32 by an index-ordered (deterministic) row update.  Replaces the autograd of SASRec/main.py:217-219 run
 * by `loss.backward()` (:249): nll_loss_backward, _log_softmax_backward_data and the two cuBLAS
 * GEMMs.  bf16 mode: d <= 256, scale > 0.  fp32x3 mode (fp32 parity, d <= 128): the same gradients from
 * exact fp32 FFMA passes (64 x 64 softmax tiles in shared memory), any sign of scale. */
int rb_ce_bwd(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
              int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
              int64_t M, int64_t N, int d, int dtype, int mode, float* dU, float* dW, float* dbias,
              const int32_t* m_dev, void* ws, size_t ws_bytes, rb_stream_t stream);

/* rb_ce_bwd's dW (+ dbias) with the gradient stored in bf16 (the dtype of a bf16 parameter): the pass
 * writes bf16 rows directly instead of an fp32 (N,d) matrix that the caller would cast (1.5 N d bytes less
 * traffic, one pass less); rows that receive the exact fp32 one-hot correction go through an fp32 side
 * table first, so every element is the correctly rounded fp32 value rb_ce_bwd would have produced.
 * bf16 mode, d <= 256, scale > 0; workspace RB_OP_CE_BWD. */
int rb_ce_bwd_dw_bf16(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                      int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                      int64_t M, int64_t N, int d, void* dW_bf16, float* dbias, const int32_t* m_dev, void* ws,
                      size_t ws_bytes, rb_stream_t stream);

/* The same pass ADDING its rows to what dW_bf16 already holds: dW_bf16 is the parameter's existing gradient buffer
 * (e.g. already carrying the embedding gather's rows), so the scoring head's dW and the gather's scatter-add end up
 * in ONE (N+P,d) gradient, as in the reference's autograd (SASRec/main.py:183,217,249), without a separate
 * gradient tensor and an accumulation pass over it.  RB_E_UNSUPPORTED when the direct bf16 pass does not apply
 * (several splits of the query range, or M > 16384): accumulate on the caller's side then. */
int rb_ce_bwd_dw_bf16_acc(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                          int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                          int64_t M, int64_t N, int d, void* dW_bf16, float* dbias, const int32_t* m_dev, void* ws,
                          size_t ws_bytes, rb_stream_t stream);

/* Masked full-catalog top-K: for every query row the K best (score desc, id asc) items among this
 * shard's N items, skipping ids in the row's seen list (CSR over GLOBAL ids, sorted ascending per
 * row; pass NULL/NULL to keep seen items).  top_ids are global (id_base + local); missing entries
 * (fewer than K unmasked items) are (RB_MASKED_SCORE, -1).  Replaces, without a dense (B,N):
 * `recommend_from_full` + `scores[seen] = -1e23` + one `torch.topk` per metric@k
 * (UniSRec/main.py:408-435).  K <= 256. */
int rb_topk_eval(const void* U, const void* W, const float* bias, float scale,
                 const int64_t* seen_crow, const int64_t* seen_col, int64_t seen_nnz, int64_t id_base,
                 int64_t B, int64_t N, int d, int dtype, int mode, int K, float* top_vals,
                 int32_t* top_ids, void* ws, size_t ws_bytes, rb_stream_t stream);

/* Diagnostics (tests, profiling): where rb_topk_eval leaves its intermediate results inside the workspace it was
 * given, for the same (B, N, d, K, mode, nnz) on the same device.  out[8] = { n_sub, cand_cap, prefix tiles, byte
 * offset of the per-row threshold ladders (64 B each: float thr[8], uint32 cnt[8]), offset of cand_cnt
 * (int32 [B][n_sub]), offset of the overflow flags (int32 [B]), offset of the candidate lists
 * ({f32 score, u32 group id} [B][n_sub][cand_cap]), bytes used }. */
int rb_topk_debug_layout(int64_t B, int64_t N, int d, int K, int mode, int64_t nnz, int64_t* out);

/* Merge R sorted per-shard lists (vals,ids)[R][B][K] into the global top-K (rb_topk_eval order). */
int rb_topk_merge(const float* vals, const int32_t* ids, int R, int64_t B, int K, float* out_vals,
                  int32_t* out_ids, rb_stream_t stream);
/* the same merge over packed[R][2][B][K] (plane 0 = float32 values, plane 1 = int32 ids): the layout ONE all-gather of
 * the per-rank (values, ids) leaves behind */
int rb_topk_merge_packed(const int32_t* packed, int R, int64_t B, int K, float* out_vals, int32_t* out_ids, rb_stream_t stream);

/* hits[b,k] = 1.0 if top_ids[b,k] (global item id, -1 = missing) is one of row b's targets, else 0.0.
 * Targets: CSR (target_crow[B+1], target_col[nnz]) of `Item.to_csr(data[IUnseen])`, ids sorted per row
 * (UniSRec/main.py:414; the reference builds the dense (B,N) multi-hot and gathers it at the top-k ids once
 * per metric@k, :428-435).  HR/NDCG/RECALL/PRECISION/MRR@k are prefix reductions of this (B,K) matrix. */
int rb_topk_hits(const int32_t* top_ids, const int64_t* target_crow, const int64_t* target_col, int64_t B,
                 int K, float* hits, rb_stream_t stream);

/* ---- 8e: merge of the per-rank (row_max, row_sumexp, label_logit) of a row-sharded table -- stats[R][3][M] as one
 * all-gather delivers them -- into the global lse and label logit of every query row (the reference has no
 * counterpart: its only multi-GPU mode is DDP).  Ranks are combined in order (deterministic). */
int rb_rowstats_merge(const float* stats, int n_ranks, int64_t M, float* lse, float* label_logit, rb_stream_t stream);

/* a10 in one pass: batch means of n_metrics METRIC@k values straight from the ranked ids -- what
 * `self.monitor(scores, targets, n=bsz, reduction="mean", pool=[...])` (UniSRec/main.py:428-435) returns per batch,
 * without a hit matrix, per-metric top-k calls or per-metric scalar reads.  kinds (HOST array): 0 HITRATE, 1 RECALL,
 * 2 PRECISION, 3 NDCG, 4 MRR; ks (HOST array): the k of each.  w[K] = 1/log2(rank+2) and w_cum[K] = its running sum
 * (device, float32); partial = device scratch of rb_topk_metrics_blocks(B) * 32 doubles; out[n_metrics] float32
 * (device).  Per-row values in float32, batch sums in double in a fixed order (deterministic). */
int rb_topk_metrics_blocks(int64_t B);
int rb_topk_metrics(const int32_t* top_ids, const int64_t* target_crow, const int64_t* target_col, int64_t B, int K,
                    const float* w, const float* w_cum, const int32_t* kinds, const int32_t* ks, int n_metrics,
                    double* partial, float* out, rb_stream_t stream);

/* Upper bound of the workspace an op needs (bytes). nnz = seen-list entries for RB_OP_TOPK_EVAL,
 * number of indices for RB_OP_SCATTER_ADD, else ignored. */
size_t rb_workspace_bytes(int op, int64_t M, int64_t N, int d, int K, int mode, int64_t nnz);

/* Number of kernel launches issued by this library since load (bench.py's `gpu_launches`). */
int64_t rb_launch_count(void);

const char* rb_last_error(void);
const char* rb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RECBOARD_B200_H */
