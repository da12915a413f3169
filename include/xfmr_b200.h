/*
 * xfmr_b200 — C ABI of the B200 (sm_100a) query-by-item score path.
 *
 * This is the drop-in boundary for the one hot path of yxtay/matrix-factorization-torch (`xfmr_rec`):
 *   - the seven embedding losses of xfmr_rec/losses.py (forward + gradients),
 *   - exact top-k retrieval with the semantics of ItemProcessor.search (xfmr_rec/data/lightning.py:237-259),
 *   - the hashed-embedding gather that feeds both (README.md:32-36; no code in the reference).
 *
 * The reference has no FFI layer (it is pure Python over ATen); the Python host side in
 * `matrix-factorization-torch_b200/` binds these symbols with ctypes and mirrors the reference's call
 * signatures. INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns every buffer, including the workspace (size from the matching *_workspace_bytes);
 *   - all work is ordered on `stream` (a cudaStream_t passed as void*); no host synchronisation, no
 *     default-stream use, no memory allocation: calls are CUDA-graph capturable.  xb_loss_forward runs its
 *     mask builder on a library-owned helper stream that forks from `stream` and joins back into it with
 *     events before the call returns to it, also on every error return (XB_FORK=0 keeps everything on `stream`);
 *   - Threading: entry points are re-entrant; any number of host threads may call them concurrently on the same
 *     or different devices / streams.  The helper streams and their events are per host thread and device, so a
 *     thread that is capturing a CUDA graph pulls only its own helper streams into the capture and never sees
 *     work enqueued by another thread (two non-blocking streams and four events per host thread and device, created
 *     on first use and kept until the process exits).  A workspace belongs to one call sequence (forward, then its backward)
 *     at a time; xb_last_error_string() and the debug / timing hooks (xb_debug_*, xb_sweep_timing*) are the
 *     only per-thread / process-wide state;
 *   - return value: 0 = ok, < 0 = error code below; xb_last_error_string() describes the last failure on
 *     the calling thread.  Nothing throws or aborts across the ABI;
 *   - there is no CPU path: a build without a GPU still loads, but every compute entry point needs sm_100.
 */
#ifndef XFMR_B200_H_
#define XFMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XB_OK 0
#define XB_ERR_INVALID_ARG (-1)   /* null pointer, negative size, inconsistent descriptor */
#define XB_ERR_UNSUPPORTED (-2)   /* shape outside what the kernels implement */
#define XB_ERR_WORKSPACE (-3)     /* workspace too small */
#define XB_ERR_CUDA (-4)          /* a CUDA runtime / driver call failed */

/* element type of embedding inputs / gradient outputs */
#define XB_DTYPE_F32 0
#define XB_DTYPE_BF16 1

/* arithmetic of the score contraction */
#define XB_COMPUTE_BF16 0   /* operands rounded to bf16, fp32 accumulate (1 tcgen05 pass) */
#define XB_COMPUTE_SPLIT 1  /* operands split hi+lo bf16, 3 tcgen05 passes: ~2^-16 relative scores */

/* loss slots, in the order of the output vector `losses[7]`; bit (1u << slot) in loss_mask.
 * Class names are the reference's (xfmr_rec/losses.py:249-359). */
#define XB_LOSS_ALIGNMENT 0             /* AlignmentLoss                                  :249-259 */
#define XB_LOSS_CONTRASTIVE 1           /* ContrastiveLoss                                :262-274 */
#define XB_LOSS_ALIGNMENT_CONTRASTIVE 2 /* AlignmentContrastiveLoss                       :277-291 */
#define XB_LOSS_INFONCE 3               /* InfomationNoiseContrastiveEstimationLoss       :294-306 */
#define XB_LOSS_MINE 4                  /* MutualInformationNeuralEstimationLoss          :309-321 */
#define XB_LOSS_PAIRWISE_HINGE 5        /* PairwiseHingeLoss                              :357-359 */
#define XB_LOSS_PAIRWISE_LOGISTIC 6     /* PairwiseLogisticLoss                           :352-354 */
#define XB_NUM_LOSSES 7

/* negative mining order when 0 < num_negatives < N */
#define XB_MINING_SEMI_HARD 0  /* semi_hard_mining (losses.py:134-162) — what every reference loss calls */
#define XB_MINING_HARD 1       /* hard_mining (losses.py:112-132): the K largest logits; defined, never called there */

/* ------------------------------------------------------------------------------------------------
 * Losses.  Replaces EmbeddingLoss.forward(user_embed, item_embed, target, *, item_idx, pos_idx)
 * (xfmr_rec/losses.py:39-52) and its autograd backward for every class selected in loss_mask, with one
 * contraction instead of one per loss (caller loop: xfmr_rec/lightning.py:137-146).
 * ---------------------------------------------------------------------------------------------- */
typedef struct xb_loss_desc {
  int32_t batch;          /* B: rows of user_embed / target                                  */
  int32_t num_items;      /* N >= B: rows of item_embed; rows 0..B-1 are the in-batch positives */
  int32_t dim;            /* d: embedding dimension (<= 256 bf16, <= 128 split)               */
  int32_t num_pos;        /* P: columns of pos_idx (0 allowed)                                */
  int32_t in_dtype;       /* XB_DTYPE_* of user_embed / item_embed / d_user / d_item          */
  int32_t compute;        /* XB_COMPUTE_*                                                     */
  int32_t num_negatives;  /* K of semi_hard_mining (losses.py:134-162); <= 0 or >= N disables */
  uint32_t loss_mask;     /* which of the 7 losses to evaluate                                */
  float sigma;            /* losses.py:31                                                     */
  float margin;           /* losses.py:32                                                     */
  int32_t has_log_q;      /* 1: subtract log_q[j] from every logit (LogQ correction; extension) */
  int32_t mining;         /* XB_MINING_*: which K columns num_negatives keeps                  */
} xb_loss_desc;

size_t xb_loss_workspace_bytes(const xb_loss_desc* desc);

/* losses_out[XB_NUM_LOSSES] (fp32): unselected slots are written as 0.  The workspace keeps what the
 * backward needs (bf16 operands, masks, row statistics); pass the same workspace to xb_loss_backward. */
int xb_loss_forward(const xb_loss_desc* desc, const void* user_embed, const void* item_embed,
                    const float* target, const int64_t* item_idx, const int64_t* pos_idx,
                    const float* log_q, float* losses_out, void* workspace, size_t workspace_bytes,
                    void* stream);

/* d_losses[XB_NUM_LOSSES] (fp32): upstream gradient of each loss slot (unselected slots ignored).
 * Writes d_user [B, d] and d_item [N, d] in desc->in_dtype. */
int xb_loss_backward(const xb_loss_desc* desc, const float* d_losses, void* d_user, void* d_item,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Uniformity (extension; no code in the reference, which only cites DirectAU / MAWU, README.md:22-25):
 *   loss = log( 1/(n(n-1)) * sum_{i != j} exp(-t * |x_i - x_j|^2) )       (Wang & Isola 2020; DirectAU's
 * `uniformity`, t = 2).  It is the Gram contraction X.X^T through the same sweep kernel as the losses
 * (rows = columns = x, only the diagonal masked), reduced over all pairs instead of per row; the
 * gradient needs one sweep (the pair weights are symmetric, so dX = 2 * the row-side gradient).
 * ---------------------------------------------------------------------------------------------- */
typedef struct xb_uniformity_desc {
  int32_t n;         /* rows of x, >= 2                          */
  int32_t dim;       /* d (<= 256 bf16, <= 128 split)            */
  int32_t in_dtype;  /* XB_DTYPE_* of x and d_x                  */
  int32_t compute;   /* XB_COMPUTE_*                             */
  float t;           /* temperature, > 0 (DirectAU: 2)           */
  int32_t reserved;
} xb_uniformity_desc;

size_t xb_uniformity_workspace_bytes(const xb_uniformity_desc* desc);
/* loss_out[1] (fp32).  The workspace keeps what the backward needs. */
int xb_uniformity_forward(const xb_uniformity_desc* desc, const void* x, float* loss_out, void* workspace,
                          size_t workspace_bytes, void* stream);
/* d_loss[1] (fp32) upstream gradient; writes d_x [n, d] in desc->in_dtype. */
int xb_uniformity_backward(const xb_uniformity_desc* desc, const float* d_loss, void* d_x, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exact top-k retrieval.  Replaces the LanceDB query of ItemProcessor.search
 * (xfmr_rec/data/lightning.py:247-258): score = q . i (cosine for unit-norm embeddings), excluded
 * items removed BEFORE ranking, k best by (score desc, item id asc).
 * ---------------------------------------------------------------------------------------------- */
typedef struct xb_topk_desc {
  int32_t num_queries;    /* Q                                                                */
  int32_t num_items;      /* N: rows of this catalog shard                                    */
  int32_t dim;            /* d                                                                */
  int32_t k;              /* results per query, 1..256                                        */
  int32_t in_dtype;       /* XB_DTYPE_* of queries and items                                  */
  int32_t compute;        /* XB_COMPUTE_BF16: scores of bf16-rounded inputs;                   */
                          /* XB_COMPUTE_SPLIT: candidates from split-bf16 scores, then re-scored */
                          /* in fp64 over the original fp32 inputs -> ids match the oracle exactly */
  int32_t has_exclusions; /* 1: excl_mask is given                                            */
  int32_t reserved;
  int64_t id_base;        /* id of catalog row r is item_ids[r] if given, else id_base + r     */
} xb_topk_desc;

size_t xb_topk_workspace_bytes(const xb_topk_desc* desc);

/* words per row of an exclusion mask for a catalog of num_items rows (multiple of 4) */
int32_t xb_mask_words(int32_t num_items);

/* excl_mask: [ceil(Q/128)*128][xb_mask_words(N)] uint32, bit (q, r) set => catalog row r is excluded for
 * query q (build it with xb_build_pair_mask).  scores_out [Q, k] fp32, ids_out [Q, k] int64; unused
 * slots (fewer than k eligible items) get score -inf and id -1. */
int xb_topk_search(const xb_topk_desc* desc, const void* queries, const void* items,
                   const int64_t* item_ids, const uint32_t* excl_mask, float* scores_out,
                   int64_t* ids_out, void* workspace, size_t workspace_bytes, void* stream);

/* k-way merge of per-shard results (catalog row-sharded across GPUs):
 * in_scores / in_ids are [Q, num_lists * list_len]; output the k best by (score desc, id asc). */
int xb_topk_merge(int32_t num_queries, int32_t num_lists, int32_t list_len, int32_t k,
                  const float* in_scores, const int64_t* in_ids, float* scores_out, int64_t* ids_out,
                  void* stream);

/* Sparse exclusion lists (the `NOT IN (exclude_item_ids)` prefilter of data/lightning.py:247-252 for catalogs too
 * large for a dense exclusion mask): in_scores / in_ids [Q, list_len] is a ranked list (e.g. xb_topk_search with
 * k = list_len), excl_ids [Q, excl_len] int64 padded with INT64_MIN.  Writes the first k entries whose id is not
 * excluded (-inf / -1 when fewer remain).  Equal to filtering BEFORE ranking whenever
 * list_len >= k + (number of excluded ids of the query), which the caller guarantees with list_len = k + excl_len. */
int xb_topk_filter(int32_t num_queries, int32_t list_len, int32_t k, int32_t excl_len, const float* in_scores,
                   const int64_t* in_ids, const int64_t* excl_ids, float* scores_out, int64_t* ids_out,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched evaluation.  Replaces the per-user torchmetrics updates of update_metrics
 * (xfmr_rec/lightning.py:149-187; metric set :289-306) for a batch of ranked result lists:
 * ids [Q, k] int64 (-1 = empty slot), target_ids [Q, T] int64 (INT64_MIN = padding), target_vals [Q, T] fp32
 * (graded relevance; relevant = value > 0).  per_query_out [Q, 6] fp32 = {ndcg, recall, precision, map,
 * hit rate, mrr} at k (torchmetrics 1.8.2 definitions, a query without a relevant target scores 0);
 * mean_out[6] (or NULL) = their means over the Q queries.
 * ---------------------------------------------------------------------------------------------- */
int xb_retrieval_metrics(int32_t num_queries, int32_t k, int32_t num_targets, const int64_t* ids,
                         const int64_t* target_ids, const float* target_vals, float* per_query_out,
                         float* mean_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pair mask builder: bit (r, c) set  <=>  col_ids[c] == row_ids0[r]  or  col_ids[c] in row_id_lists[r, :].
 * This is ~negative_masks of losses.py:92-110 (row_ids0 = item_idx[:B], lists = pos_idx) and the
 * `NOT IN (exclude_item_ids)` prefilter of data/lightning.py:247-252 (row_ids0 = NULL).
 * mask   : [ceil(num_rows/128)*128][xb_mask_words(num_cols)]      (bits for c >= num_cols are set)
 * mask_t : [ceil(num_cols/128)*128][xb_mask_words(num_rows)] or NULL (transposed copy)
 * Ids are any int64 except INT64_MIN (the empty-slot key of the builder's hash table); that goes for item_idx / pos_idx of
 * xb_loss_forward too, which builds its mask with the same kernels.
 * ---------------------------------------------------------------------------------------------- */
size_t xb_pair_mask_workspace_bytes(int32_t num_cols);
int xb_build_pair_mask(int32_t num_rows, int32_t num_cols, int32_t list_len, const int64_t* col_ids,
                       const int64_t* row_ids0, const int64_t* row_id_lists, uint32_t* mask,
                       uint32_t* mask_t, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hashed ("Bloom") embedding gather: out[n] = sum_{h < num_hashes} table[XXH32(le64(ids[n]), seed0 + h)
 * mod 2^log2_rows].  Integer hashing is bit-exact XXH32 (xxhash 0.8 / python-xxhash 3.7.0).
 * table: [2^log2_rows, dim] bf16; out: [n, dim] bf16 (fp32 sum, one rounding); idx_out [n, num_hashes]
 * int32 or NULL.
 * ---------------------------------------------------------------------------------------------- */
int xb_hash_indices(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0, int32_t log2_rows,
                    int32_t* idx_out, void* stream);
int xb_hash_gather(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0, const void* table,
                   int32_t log2_rows, int32_t dim, void* out, int32_t* idx_out, void* stream);
/* gradient of the gather w.r.t. the table: d_table[idx] += d_out[n] (fp32 atomics; d_table fp32, zeroed
 * by the caller). */
int xb_hash_scatter_grad(const int64_t* ids, int64_t n, int32_t num_hashes, uint32_t seed0,
                         const void* d_out, int32_t log2_rows, int32_t dim, float* d_table, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Test hook: raw score tiles and the second-MMA path.  s_out [ceil(R/128)*128][ceil(C/128)*128] fp32 =
 * rows . cols^T; acc_out [ceil(R/128)*128][kp] fp32 = bf16(S masked to valid cols) . cols.
 * ---------------------------------------------------------------------------------------------- */
size_t xb_debug_workspace_bytes(int32_t num_rows, int32_t num_cols, int32_t dim, int32_t compute);
int xb_debug_scores(int32_t num_rows, int32_t num_cols, int32_t dim, int32_t in_dtype, int32_t compute,
                    const void* rows, const void* cols, float* s_out, float* acc_out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Debug hook: while `trace` (device, int64 [tiles][8]) is set, CTA (0,0) of every sweep writes SM clock stamps per
 * tile: [0..2] MMA issuer before / after the S-buffer wait / after the operand wait, [3..5] epilogue warp before /
 * after the score-tile wait / at the end of the tile, [6..7] TMA producer before / after the stage wait.  NULL = off. */
int xb_debug_set_trace(int64_t* trace, int32_t tiles);

/* Measurement hook (bench.py only): while enabled, every tensor-core sweep launch is
 * bracketed by CUDA events on its stream.  xb_sweep_timing(1) starts (and clears), xb_sweep_timing(0) stops;
 * xb_sweep_timing_read synchronises on the recorded events (the one entry point that blocks the host) and
 * returns the summed device time in milliseconds and the number of sweep launches. */
int xb_sweep_timing(int32_t enable);
int xb_sweep_timing_read(double* total_ms_host, int64_t* count_host);

/* Test hook: where a named region of the loss workspace lives (valid after xb_loss_forward).
 * region 0 = mined columns int32 [B][K] (-1 = none), 1 = their logits (log2 units) fp32 [B][K],
 * 2 = pair mask uint32 [B_pad][xb_mask_words(N)], 3 = row statistics float4 [B] {count, lse, lse+pos, 1/count},
 * 4 = mining candidates uint64 [B][2*(K+16)]. */
int xb_debug_loss_region(const xb_loss_desc* desc, int32_t region, size_t* offset_host, size_t* bytes_host);

const char* xb_last_error_string(void);
const char* xb_version(void);
/* number of kernel launches issued through this library (process-wide) since the last reset */
int64_t xb_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* XFMR_B200_H_ */
