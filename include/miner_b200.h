/*
 * miner_b200 -- C ABI of the B200-native MINER scoring path.
 *
 * Drop-in boundary.  The reference (MrRobot2211/miner) is pure
 * Python/PyTorch and has no FFI of its own; the interface it exposes for this path is the
 * nn.Module API of src/model/model.py plus the metric functions of src/evaluation.py.  Each
 * entry point below replaces the aten calls behind one of those functions; the Python mirror
 * of the reference API (miner_b200/model.py, evaluation.py, loss.py) binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless said otherwise;
 *   - no allocation inside: the caller passes outputs and (where a *_workspace_bytes query
 *     exists) a scratch buffer;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and re-entrant
 *     per stream;
 *   - return value 0 = ok, non-zero = error code below; the message of the last error on the
 *     calling thread is available from miner_last_error();
 *   - row-major, contiguous tensors; B impressions (rows), H history length, C candidates per
 *     row, K context codes, Dc context-code dim, D news dim, T total candidates (CSR).
 */
#ifndef MINER_B200_H_
#define MINER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINER_B200_ABI_VERSION 2

/* error codes */
#define MINER_OK               0
#define MINER_ERR_INVALID_ARG  1   /* bad size / null pointer / unsupported dtype                      */
#define MINER_ERR_SCORE_TYPE   2   /* reference: ValueError('Invalid method of aggregating matching score'), model.py:136 */
#define MINER_ERR_CUDA         3   /* a CUDA runtime call or launch failed                             */
#define MINER_ERR_WORKSPACE    4   /* workspace too small (see the *_workspace_bytes queries)          */
#define MINER_ERR_UNSUPPORTED  5   /* shape outside what the requested kernel family supports          */

/* element / index types */
#define MINER_F32   0
#define MINER_BF16  1
#define MINER_I32   0
#define MINER_I64   1

/* score aggregation, Miner.score_type (model.py:127-136) */
#define MINER_SCORE_MAX       0
#define MINER_SCORE_MEAN      1
#define MINER_SCORE_WEIGHTED  2

/* kernel family for the dense contractions */
#define MINER_MATH_FP32    0   /* fp32 CUDA-core kernels, reference operation order                     */
#define MINER_MATH_TENSOR  1   /* tcgen05 bf16 tensor cores (fp32 accumulate) for the two projection
                                  GEMMs; softmax / weighted sums / dot-scores stay fp32               */
#define MINER_MATH_TABLE   2   /* table-level mode: projections applied once per table row
                                  (miner_table_project) + one fused scoring kernel (miner_score_table_fwd);
                                  not a value of miner_score_params.math -- it has its own entry points    */

int         miner_abi_version(void);
const char* miner_last_error(void);
/* sm_count / compute capability of the current device (host pointers, any may be NULL) */
int miner_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in this process so far */
uint64_t miner_launch_count(void);

/* ---- (a1) embedding gather: replaces NewsEncoder.forward as called at model.py:96-97,109-110
 *      (RoBERTa restated as table[id], news_encoder.py:60-106).  out[i,:] = table[ids[i],:], bit-exact.
 *      oob_flag (nullable, device int32): set to 1 if any id is outside [0,n_rows); such rows are zero-filled. */
int miner_gather(const void* table, int64_t n_rows, int64_t dim, int dtype,
                 const void* ids, int64_t n_ids, int id_dtype,
                 void* out, int32_t* oob_flag, void* stream);

/* ---- (a2) category-similarity bias: replaces model.py:113-120 + utils.py:9-29 (pairwise cosine).
 *      cat_emb (NC,Ec) fp32; his_cat (B,H); cand_cat (B,C) dense.  Writes bias_full (B,H,C) if non-null and
 *      bias_mean (B,H) = bias.mean(dim=2) (model.py:176).  0/0 -> NaN for the zero padding row, as the reference. */
int miner_category_bias(const float* cat_emb, int64_t n_cat, int64_t ec,
                        const void* his_cat, const void* cand_cat, int id_dtype,
                        int64_t B, int64_t H, int64_t C,
                        float* bias_full, float* bias_mean, void* stream);

/* ---- (a3) PolyAttention.forward (model.py:159-185).
 *      emb (B,H,D) fp32; mask (B,H) uint8 (torch.bool); bias_mean (B,H) fp32 or NULL;
 *      w_proj (Dc,D), codes (K,Dc) fp32.  out_interests (B,K,D) fp32; out_weights (B,K,H) fp32 or NULL.
 *      Masked logits are filled with 1e-30 (model.py:180), not -inf. */
size_t miner_poly_attn_workspace_bytes(int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D);
int miner_poly_attn_fwd(const float* emb, const uint8_t* mask, const float* bias_mean,
                        const float* w_proj, const float* codes,
                        int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D,
                        float* out_interests, float* out_weights,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- (a4,a5) matching scores + TargetAwareAttention.forward (model.py:127-136,200-216).
 *      interests (B,K,D) fp32; cand (T,D) fp32 gathered candidate vectors; cand_offsets (B+1) int64 CSR, or NULL
 *      for the dense layout with C candidates per row (T = B*C); w_target (D,D) fp32 (ignored unless WEIGHTED).
 *      matching (T,K) fp32 or NULL: the `value` argument of TargetAwareAttention.forward (model.py:200); when NULL
 *      the kernel computes cand . interests^T itself (model.py:127).  out_scores (T) fp32. */
size_t miner_target_score_workspace_bytes(int64_t B, int64_t K, int64_t D, int score_type);
int miner_target_score_fwd(const float* interests, const float* cand, const float* matching,
                           const int64_t* cand_offsets, const float* w_target, int score_type,
                           int64_t B, int64_t C, int64_t K, int64_t D,
                           float* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- (a1..a6) fused table-based scoring: Miner.forward (model.py:61-138) for a block of impressions.
 *      table (N,D) fp32 or bf16; his_ids (B,H); his_mask (B,H) uint8; cand_ids (T) with cand_offsets (B+1) int64
 *      CSR, or NULL offsets for dense C per row; bias_mean (B,H) or NULL.  weights fp32 (the TENSOR family
 *      additionally takes bf16 copies made with miner_cast_f32_to_bf16).  out_scores (T) fp32;
 *      out_interests (B,K,D) fp32 or NULL (eval does not need them).
 *      math = MINER_MATH_FP32 (any table dtype) or MINER_MATH_TENSOR (bf16 table, sm_100a tcgen05). */
typedef struct miner_score_params {
  const void*    table;        int64_t n_rows;  int table_dtype;
  const void*    his_ids;      const uint8_t* his_mask;
  const void*    cand_ids;     const int64_t* cand_offsets;   int id_dtype;
  const float*   bias_mean;
  const float*   w_proj;       const float* codes;   const float* w_target;
  const void*    w_proj_bf16;  const void*  w_target_bf16;     /* TENSOR family only, else NULL */
  int64_t B, H, C, K, Dc, D, T;
  int     score_type;
  int     math;
  float*  out_scores;
  float*  out_interests;
  int     stage_mask;   /* 0 = whole path; else bit 0 history projection, 1 poly softmax + weighted sum, 2 target projection,
                           3 target attention + score (profiling: lets a caller time one kernel of the path alone) */
} miner_score_params;

size_t miner_score_workspace_bytes(const miner_score_params* p, int64_t chunk_impressions);
int    miner_score_fwd(const miner_score_params* p, int64_t chunk_impressions,
                       void* workspace, size_t workspace_bytes, void* stream);
/* fp32 (rows,cols) -> bf16 (round-to-nearest-even), used once per weight matrix / table */
int miner_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* The tcgen05 projection GEMM of the TENSOR family on its own (nn.Linear without bias, model.py:155,198):
 *      c[M,N] fp32 = epi(a[M,K] b[N,K]^T), bf16 operands, fp32 accumulation; a rows optionally gathered
 *      (row m = a[a_ids[m],:], ids checked against a_rows_in_table); epilogue 0 none, 1 tanh, 2 exact-erf gelu;
 *      c_bf16 (nullable) receives a bf16 copy.  Needs K % 64 == 0 and N >= 16. */
int miner_tc_gemm(const void* a_bf16, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* b_bf16,
                  float* c, void* c_bf16, int64_t M, int64_t N, int64_t K, int epilogue, void* stream);

/* The weight-gradient GEMM of the train variant on its own (trainer.py:246-261: d loss / d nn.Linear.weight = dY^T X):
 *      c[s][M,N] fp32 = sum over the rows of split s of a[r,:]^T b[r,:], a (R,M) and b (R,N) bf16 row-major, read MN-major by
 *      tcgen05 (no transposed copies); c holds k_splits partials, summed by the caller in split order.
 *      Needs M % 8 == 0, N % 8 == 0, N >= 16. */
int miner_tc_gemm_tn(const void* a_bf16, const void* b_bf16, float* c, int64_t R, int64_t M, int64_t N, int k_splits, void* stream);

/* The two kernels of the fused tensor-core path on their own (tests / profiling).
 *   miner_hist_interests_fwd: PolyAttention.forward (model.py:159-185) straight from a bf16 table: gathers table[his_ids],
 *      projects on tcgen05, softmax over the history, weighted sum on tcgen05.  Writes the interests split as two bf16 arrays
 *      i_hi + i_lo (B*K, D) and, if out_interests != NULL, as fp32 (B,K,D).  Needs H <= 128, K in {8,16,32}, Dc <= 208, D % 64 == 0.
 *   miner_cand_score_fwd: matching scores + TargetAwareAttention (model.py:127,200-216, score_type 'weighted') from
 *      i_hi / i_lo and table[cand_ids]; CSR offsets (B+1) or NULL for dense C per row.  out_scores (T) fp32. */
size_t miner_hist_interests_workspace_bytes(int64_t Dc);
int miner_hist_interests_fwd(const void* table_bf16, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                             const float* bias_mean, const void* w_proj_bf16, const float* codes,
                             int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D,
                             void* i_hi, void* i_lo, float* out_interests, void* workspace, size_t workspace_bytes, void* stream);
int miner_cand_score_fwd(const void* i_hi, const void* i_lo, const void* w_target_bf16, const void* table_bf16, int64_t n_rows,
                         const void* cand_ids, int id_dtype, const int64_t* cand_offsets,
                         int64_t B, int64_t C, int64_t K, int64_t D, float* out_scores, void* stream);

/* ---- Table-level mode of the scoring path (opt-in; same results to fp32 rounding, different operation order).
 *      The two nn.Linear layers of the path act row-wise on gathered news vectors (model.py:171 PolyAttention.linear,
 *      model.py:212 TargetAwareAttention.linear applied to sum_h w[k,h] E[h]), so they can be applied once per TABLE row
 *      instead of once per gathered row:
 *        lg (n_rows,K) fp32 = tanh(table Wp^T) codes^T      (model.py:171,174)
 *        tw (n_rows,D) bf16 = table Wt^T                     (model.py:212 before the gelu; NULL unless score_type WEIGHTED)
 *      miner_table_project computes them (once per weight version / table; bench.py does it inside every timed step);
 *      miner_score_table_fwd is Miner.forward (model.py:61-138) for B impressions from table, lg, tw in one kernel:
 *      masked softmax over the history (1e-30 fill, model.py:180), interests, gelu, matching scores, softmax over K, score.
 *      Needs a bf16 table, H <= 256, K <= 64, D % 64 == 0 (K <= 32: two impressions share a tile of 128 TMEM lanes; K <= 16 and
 *      H <= 32: four).
 *      out_interests (B,K,D) fp32 or NULL. */
size_t miner_table_project_workspace_bytes(int64_t n_rows, int64_t Dc);
int miner_table_project(const void* table_bf16, int64_t n_rows, int64_t D, const void* w_proj_bf16, const float* codes,
                        const void* w_target_bf16, int64_t K, int64_t Dc, float* out_lg, void* out_tw,
                        void* workspace, size_t workspace_bytes, void* stream);
int miner_score_table_supported(int64_t H, int64_t K, int64_t D);
/*      Workspace of miner_score_table_fwd: the call first packs the histories into tiles (one launch, one warp per tile): masked
 *      slots that point at the same news row -- the left padding, reader.py:368-369 -- all carry the logit 1e-30 (model.py:180)
 *      and are merged into one slot of that multiplicity, so a short history costs as many gathered rows as it has clicks + 1.
 *      The first 8 bytes of the workspace are two int32 counters the call ADDS to: history ids / candidate ids outside
 *      [0, n_rows) (such rows read as zero; the reference's indexing raises IndexError).  Zero them before the first call and
 *      read them back when convenient.  miner_score_table_tile_geometry reports the tiling of a shape (0 = unsupported). */
size_t miner_score_table_workspace_bytes(int64_t B, int64_t H, int64_t K);
int miner_score_table_tile_geometry(int64_t H, int64_t K, int* impressions_per_tile, int* halves);
int miner_score_table_fwd(const void* table_bf16, const void* tw_bf16, const float* lg, int64_t n_rows,
                          const void* his_ids, const uint8_t* his_mask, const void* cand_ids, const int64_t* cand_offsets,
                          int id_dtype, const float* bias_mean, int64_t B, int64_t H, int64_t C, int64_t K, int64_t D,
                          int score_type, float* out_scores, float* out_interests,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- (a7..a12) segmented per-impression ranking metrics: replaces SlowEvaluator/FastEvaluator +
 *      compute_scores (evaluation.py:36-84,87-175) and compute_mrr/dcg/ndcg_score, is_hit (:177-249).
 *      scores (T) fp32 logits; labels (T) int8; offsets (B+1) int64; transform: 0 none, 1 sigmoid (SlowEvaluator,
 *      :165), 2 softmax within the impression (FastEvaluator, :109).  ks: host array of n_k cut-offs (<= 8 of them).
 *      Metric order m: 0 group_auc, 1 mrr, 2.. ndcg@ks[j], 2+n_k.. hit@ks[j]   (M = 2 + 2*n_k).
 *      out_partials (2*M doubles): [sum over non-NaN impressions, count of non-NaN impressions] per metric,
 *      reduced deterministically.  out_per_impression (B*M doubles) or NULL. */
size_t miner_rank_metrics_workspace_bytes(int64_t B, int n_k);
int miner_rank_metrics(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B,
                       int transform, const int* ks, int n_k,
                       double* out_partials, double* out_per_impression,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- global `auc` over ALL candidates (evaluation.py:53-55: sklearn roc_auc_score on the flattened lists) = the tie-aware
 *      Mann-Whitney statistic U / (P N), in exact integers and without a comparison sort:
 *        miner_auc_split   probabilities (same transform as miner_rank_metrics) -> order-preserving uint32 keys, compacted into
 *                          pos_keys / neg_keys (capacity T each); counts (2 device uint64) = [P, N]
 *        miner_sort_u32    LSD radix sort (ascending, in place) -- of the positive keys; with several ranks, of the all-gathered
 *                          positive keys of every rank
 *        miner_auc_count   out_u2 (device uint64) = sum over the local negatives of 2 #{pos > neg} + #{pos == neg}
 *      auc = sum_ranks(out_u2) / (2 sum(P) sum(N)).  miner_b200/evaluation.py global_auc drives the three calls (and the NCCL
 *      all-gather / all-reduce between them). */
int miner_auc_split(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B, int64_t T, int transform,
                    uint32_t* pos_keys, uint32_t* neg_keys, uint64_t* counts, void* stream);
size_t miner_sort_u32_workspace_bytes(int64_t n);
int miner_sort_u32(uint32_t* keys, int64_t n, void* workspace, size_t workspace_bytes, void* stream);
int miner_auc_count(const uint32_t* pos_sorted, int64_t P, const uint32_t* neg_keys, int64_t N, uint64_t* out_u2, void* stream);

/* ---- (a13,a14) losses: Loss.compute (loss.py:27-44) and Loss.compute_eval_loss (loss.py:68-85).
 *      interests (B,K,D); logits (B,C); labels (B,C) fp32 (one-hot for compute, binary for eval).
 *      out (3 floats, device): [total, disagreement, rank_loss]. mode 0 = compute (CE mean), 1 = compute_eval_loss. */
size_t miner_loss_workspace_bytes(int64_t B, int64_t K);
int miner_loss_fwd(const float* interests, const float* logits, const float* labels,
                   int64_t B, int64_t C, int64_t K, int64_t D, int mode,
                   float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Train variant: Miner.forward with the intermediates its backward needs, the backward of
 *      Loss.compute (loss.py:27-44) and the backward of Miner.forward (model.py:61-138) down to the weight matrices, for every
 *      score_type (model.py:127-136) and with or without the category-bias scalar of the poly attention (model.py:176-177).
 *      fp32, reference operation order, deterministic reductions.  Dense layout: cand_ids (B,C).  The news table is a frozen buffer
 *      here (no gradient flows into its rows).
 *      miner_train_fwd writes interests (B,K,D), scores (B,C) and saves T = tanh(E Wp^T) (B*H,Dc), the softmax weights (B,K,H) and,
 *      for 'weighted', Z = I Wt^T (B*K,D).  bias_mean (B,H) or NULL is added to every code's logit of a history slot.
 *      miner_loss_bwd: d loss / d interests and d loss / d logits, scaled by *grad_out (device float, NULL = 1).
 *      miner_train_bwd: grad_w_proj (Dc,D), grad_codes (K,Dc), grad_w_target (D,D; 'weighted' only, else NULL) from d_scores (B,C)
 *      and d_interests (B,K,D, nullable); d_bias_mean (B,H) or NULL receives d loss / d bias_mean (the category embedding's
 *      gradient continues from there).  grad_table (n_rows, D) fp32 or NULL: the gradient of the table ROWS (history rows through
 *      I = w E and tanh(E Wp^T), candidate rows through the matching / attention dots) is ADDED into it with atomics -- zero it
 *      first -- so that the table can be a trainable parameter or the dense output of an upstream news encoder (the reference
 *      trains its encoder, trainer.py:146-169); it needs its own scratch of miner_train_table_grad_workspace_bytes.
 *      The same workspace size serves fwd and bwd.
 *      math = MINER_MATH_FP32, or MINER_MATH_TENSOR: the five projection-sized GEMMs of the step (E Wp^T, I Wt^T, dZ Wt,
 *      dZ^T I, dZ1^T E) on tcgen05 with bf16 operands and fp32 accumulation (bf16 table, D % 64 == 0, bf16 weight copies). */
size_t miner_train_workspace_bytes(int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D, int math);
int miner_train_fwd(const void* table, int64_t n_rows, int table_dtype, const void* his_ids, const uint8_t* his_mask,
                    const void* cand_ids, int id_dtype, const float* w_proj, const float* codes, const float* w_target,
                    int64_t B, int64_t H, int64_t C, int64_t K, int64_t Dc, int64_t D, float* out_interests, float* out_scores,
                    float* save_t, float* save_w, float* save_z,
                    int math, const void* w_proj_bf16, const void* w_target_bf16, int score_type, const float* bias_mean,
                    void* workspace, size_t workspace_bytes, void* stream);
int miner_loss_bwd(const float* interests, const float* logits, const float* labels, const float* grad_out,
                   int64_t B, int64_t C, int64_t K, int64_t D, float* d_interests, float* d_logits, void* stream);
int miner_train_bwd(const void* table, int64_t n_rows, int table_dtype, const void* his_ids, const uint8_t* his_mask,
                    const void* cand_ids, int id_dtype, const float* w_proj, const float* codes, const float* w_target,
                    const float* save_t, const float* save_w, const float* interests, const float* save_z,
                    const float* d_scores, const float* d_interests,
                    int64_t B, int64_t H, int64_t C, int64_t K, int64_t Dc, int64_t D,
                    float* grad_w_proj, float* grad_codes, float* grad_w_target,
                    int math, const void* w_proj_bf16, const void* w_target_bf16, int score_type, float* d_bias_mean,
                    float* grad_table, void* table_grad_workspace, size_t table_grad_workspace_bytes,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t miner_train_table_grad_workspace_bytes(int64_t B, int64_t H, int64_t Dc, int64_t D);

#ifdef __cplusplus
}
#endif
#endif  /* MINER_B200_H_ */
