// extern "C" entry points of libminer_b200.so (see include/miner_b200.h) and the chunked orchestration of the
// fused table-based scoring path (Miner.forward, reference src/model/model.py:61-138).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "tc/tc_gemm.cuh"
#include "tc/fused.cuh"

namespace miner {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

struct ScoreWs {
  size_t proj, interests, interests_bf16, target_proj, total;
  // fused tensor-core path (hist_kernel -> cand_kernel)
  bool fused;
  size_t i_hi, i_lo, codes_t;
};

// The fused tcgen05 path covers the tensor family for the shapes the two kernels support; everything else (and the
// fp32 family) runs the 4-kernel pipeline.
static bool use_fused(const miner_score_params* p) {
  return p->math == MINER_MATH_TENSOR && hist_kernel_supported(p->H, p->K, p->Dc, p->D) &&
         (p->score_type != MINER_SCORE_WEIGHTED || cand_kernel_supported(p->K, p->D));
}

static ScoreWs score_ws_layout(const miner_score_params* p, int64_t chunk) {
  ScoreWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  w.fused = use_fused(p);
  if (w.fused) {
    const size_t rows = (size_t)chunk * p->K;
    w.i_hi = take(2 * rows * p->D);
    w.i_lo = take(2 * rows * p->D);
    w.codes_t = take(hist_kernel_ws_bytes(p->Dc));
    // 'max' / 'mean' aggregate fp32 interests with the CUDA-core kernel
    w.interests = take((p->out_interests || p->score_type == MINER_SCORE_WEIGHTED) ? 0 : sizeof(float) * rows * p->D);
    w.total = off;
    return w;
  }
  w.proj = take(sizeof(float) * chunk * p->H * p->Dc);
  w.interests = take(p->out_interests ? 0 : sizeof(float) * chunk * p->K * p->D);
  const bool weighted = p->score_type == MINER_SCORE_WEIGHTED;
  w.interests_bf16 = take((p->math == MINER_MATH_TENSOR && weighted) ? 2 * (size_t)chunk * p->K * p->D : 0);
  w.target_proj = take(weighted ? sizeof(float) * chunk * p->K * p->D : 0);
  w.total = off;
  return w;
}

static int check_score_params(const miner_score_params* p, int64_t chunk) {
  MINER_CHECK_ARG(p != nullptr, "score: null params");
  MINER_CHECK_ARG(p->table && p->his_ids && p->his_mask && p->cand_ids && p->w_proj && p->codes && p->out_scores,
                  "score: null pointer");
  MINER_CHECK_ARG(p->B >= 0 && p->H > 0 && p->K > 0 && p->Dc > 0 && p->D > 0 && p->n_rows > 0, "score: bad sizes");
  MINER_CHECK_ARG(p->cand_offsets || p->C > 0, "score: dense layout needs C > 0");
  MINER_CHECK_ARG(chunk > 0, "score: chunk_impressions must be positive");
  MINER_CHECK_ARG(p->table_dtype == MINER_F32 || p->table_dtype == MINER_BF16, "score: table dtype must be fp32 or bf16");
  MINER_CHECK_ARG(p->id_dtype == MINER_I32 || p->id_dtype == MINER_I64, "score: id dtype must be int32 or int64");
  if (p->score_type != MINER_SCORE_MAX && p->score_type != MINER_SCORE_MEAN && p->score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  MINER_CHECK_ARG(p->score_type != MINER_SCORE_WEIGHTED || p->w_target, "score: 'weighted' needs w_target");
  if (p->math == MINER_MATH_TENSOR) {
    MINER_CHECK_ARG(p->table_dtype == MINER_BF16, "score: the tensor-core family needs a bf16 table");
    MINER_CHECK_ARG(p->w_proj_bf16 && (p->score_type != MINER_SCORE_WEIGHTED || p->w_target_bf16),
                    "score: the tensor-core family needs bf16 weight copies");
    if (!tc_gemm_supported(p->D, p->Dc) || !tc_gemm_supported(p->D, p->D)) {
      set_error("score: tensor-core family needs D %% 64 == 0 (D=%lld)", (long long)p->D);
      return MINER_ERR_UNSUPPORTED;
    }
  } else {
    MINER_CHECK_ARG(p->math == MINER_MATH_FP32, "score: unknown math family");
  }
  return MINER_OK;
}

}  // namespace miner

using namespace miner;

extern "C" int miner_abi_version(void) { return MINER_B200_ABI_VERSION; }
extern "C" const char* miner_last_error(void) { return g_err; }

extern "C" uint64_t miner_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int miner_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  MINER_CUDA_OK(cudaGetDevice(&dev));
  int v = 0;
  if (sm) { MINER_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm = v; }
  if (major) { MINER_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *major = v; }
  if (minor) { MINER_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *minor = v; }
  return MINER_OK;
}

extern "C" int miner_gather(const void* table, int64_t n_rows, int64_t dim, int dtype, const void* ids, int64_t n_ids,
                            int id_dtype, void* out, int32_t* oob_flag, void* stream) {
  MINER_CHECK_ARG(n_ids >= 0 && dim >= 0 && n_rows >= 0, "gather: negative size");
  MINER_CHECK_ARG(n_ids == 0 || dim == 0 || (table && ids && out), "gather: null pointer");
  MINER_CHECK_ARG(dtype == MINER_F32 || dtype == MINER_BF16, "gather: dtype must be fp32 or bf16");
  MINER_CHECK_ARG(id_dtype == MINER_I32 || id_dtype == MINER_I64, "gather: id dtype must be int32 or int64");
  return launch_gather(table, n_rows, dim, dtype, ids, n_ids, id_dtype, out, oob_flag, static_cast<cudaStream_t>(stream));
}

extern "C" int miner_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  MINER_CHECK_ARG(n >= 0 && (n == 0 || (src && dst)), "cast: bad arguments");
  return launch_cast_f32_to_bf16(src, dst, n, static_cast<cudaStream_t>(stream));
}

extern "C" size_t miner_poly_attn_workspace_bytes(int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D) {
  (void)K; (void)D;
  return align_up(sizeof(float) * (size_t)B * H * Dc, 256);
}

extern "C" int miner_poly_attn_fwd(const float* emb, const uint8_t* mask, const float* bias_mean, const float* w_proj,
                                   const float* codes, int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D,
                                   float* out_interests, float* out_weights, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  MINER_CHECK_ARG(B >= 0 && H > 0 && K > 0 && Dc > 0 && D > 0, "poly_attn: bad sizes");
  if (B == 0) return MINER_OK;
  MINER_CHECK_ARG(emb && mask && w_proj && codes && out_interests, "poly_attn: null pointer");
  if (!workspace || workspace_bytes < miner_poly_attn_workspace_bytes(B, H, K, Dc, D)) {
    set_error("poly_attn: workspace too small (%zu bytes needed)", miner_poly_attn_workspace_bytes(B, H, K, Dc, D));
    return MINER_ERR_WORKSPACE;
  }
  auto st = static_cast<cudaStream_t>(stream);
  float* proj = static_cast<float*>(workspace);
  int rc = launch_sgemm_nt(emb, MINER_F32, nullptr, MINER_I64, 0, w_proj, proj, B * H, Dc, D, EPI_TANH, st);   // model.py:171
  if (rc) return rc;
  return launch_poly_softmax_wsum(proj, codes, mask, bias_mean, emb, nullptr, MINER_F32, nullptr, MINER_I64, 0, B, H, K, Dc, D,
                                  out_interests, out_weights, nullptr, st);
}

extern "C" size_t miner_target_score_workspace_bytes(int64_t B, int64_t K, int64_t D, int score_type) {
  return score_type == MINER_SCORE_WEIGHTED ? align_up(sizeof(float) * (size_t)B * K * D, 256) : 0;
}

extern "C" int miner_target_score_fwd(const float* interests, const float* cand, const float* matching,
                                      const int64_t* cand_offsets, const float* w_target, int score_type, int64_t B, int64_t C, int64_t K, int64_t D,
                                      float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(B >= 0 && K > 0 && D > 0 && (cand_offsets || C > 0), "target_score: bad sizes");
  if (score_type != MINER_SCORE_MAX && score_type != MINER_SCORE_MEAN && score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  if (B == 0) return MINER_OK;
  MINER_CHECK_ARG(interests && cand && out_scores, "target_score: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  float* proj = nullptr;
  if (score_type == MINER_SCORE_WEIGHTED) {
    MINER_CHECK_ARG(w_target, "target_score: 'weighted' needs w_target");
    if (!workspace || workspace_bytes < miner_target_score_workspace_bytes(B, K, D, score_type)) {
      set_error("target_score: workspace too small (%zu bytes needed)", miner_target_score_workspace_bytes(B, K, D, score_type));
      return MINER_ERR_WORKSPACE;
    }
    proj = static_cast<float*>(workspace);
    int rc = launch_sgemm_nt(interests, MINER_F32, nullptr, MINER_I64, 0, w_target, proj, B * K, D, D, EPI_GELU, st);  // model.py:212
    if (rc) return rc;
  }
  return launch_target_score(interests, proj, matching, cand, nullptr, MINER_F32, nullptr, MINER_I64, 0, cand_offsets, B, C, K, D,
                             score_type, out_scores, st);
}

extern "C" size_t miner_score_workspace_bytes(const miner_score_params* p, int64_t chunk) {
  if (!p || chunk <= 0) return 0;
  if (p->B > 0 && chunk > p->B) chunk = p->B;
  return score_ws_layout(p, chunk).total;
}

extern "C" int miner_score_fwd(const miner_score_params* p, int64_t chunk, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_score_params(p, chunk);
  if (rc) return rc;
  if (p->B == 0) return MINER_OK;
  if (chunk > p->B) chunk = p->B;
  const ScoreWs w = score_ws_layout(p, chunk);
  if (!workspace || workspace_bytes < w.total) {
    set_error("score: workspace too small (%zu bytes needed, %zu given)", w.total, workspace_bytes);
    return MINER_ERR_WORKSPACE;
  }
  auto st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const bool weighted = p->score_type == MINER_SCORE_WEIGHTED;
  const bool tensor = p->math == MINER_MATH_TENSOR;
  const int64_t id_sz = p->id_dtype == MINER_I64 ? 8 : 4;
  float* proj = reinterpret_cast<float*>(ws + w.proj);
  float* tproj = weighted ? reinterpret_cast<float*>(ws + w.target_proj) : nullptr;
  void* i_bf16 = (tensor && weighted) ? ws + w.interests_bf16 : nullptr;

  for (int64_t b0 = 0; b0 < p->B; b0 += chunk) {
    const int64_t nb = p->B - b0 < chunk ? p->B - b0 : chunk;
    const void* his = static_cast<const char*>(p->his_ids) + b0 * p->H * id_sz;
    const uint8_t* msk = p->his_mask + b0 * p->H;
    const float* bias = p->bias_mean ? p->bias_mean + b0 * p->H : nullptr;
    float* interests = p->out_interests ? p->out_interests + b0 * p->K * p->D : reinterpret_cast<float*>(ws + w.interests);
    const int stages = p->stage_mask ? p->stage_mask : 15;
    if (w.fused) {
      float* i_f32 = p->out_interests ? interests : (weighted ? nullptr : interests);
      // steps 1+2: interests straight from the table (model.py:104-111,159-185)
      if (stages & 1) {
        rc = launch_hist_kernel(p->table, p->n_rows, his, p->id_dtype, msk, bias, p->w_proj_bf16, p->codes, nb, p->H, p->K, p->Dc, p->D,
                                ws + w.i_hi, ws + w.i_lo, i_f32, reinterpret_cast<float*>(ws + w.codes_t), st);
        if (rc) return rc;
      }
      if (!(stages & 8)) continue;
      // steps 3+4: matching scores, target-aware attention, per-candidate score (model.py:127-136,200-216)
      const void* cids = p->cand_offsets ? p->cand_ids : static_cast<const void*>(static_cast<const char*>(p->cand_ids) + b0 * p->C * id_sz);
      const int64_t* coffs = p->cand_offsets ? p->cand_offsets + b0 : nullptr;
      float* outs = p->cand_offsets ? p->out_scores : p->out_scores + b0 * p->C;
      if (weighted)
        rc = launch_cand_kernel(ws + w.i_hi, ws + w.i_lo, p->w_target_bf16, p->table, p->n_rows, cids, p->id_dtype, coffs, nb, p->C, p->K,
                                p->D, outs, st);
      else
        rc = launch_target_score(i_f32, nullptr, nullptr, nullptr, p->table, p->table_dtype, cids, p->id_dtype, p->n_rows, coffs, nb,
                                 p->cand_offsets ? 0 : p->C, p->K, p->D, p->score_type, outs, st);
      if (rc) return rc;
      continue;
    }
    // step 1+2a: proj = tanh(table[his_ids] Wp^T)     (model.py:104-111,171)
    if (!(stages & 1)) rc = MINER_OK;
    else if (tensor)
      rc = launch_tc_gemm(p->table, his, p->id_dtype, p->n_rows, p->w_proj_bf16, proj, nullptr, nb * p->H, p->Dc, p->D, EPI_TANH, st);
    else
      rc = launch_sgemm_nt(p->table, p->table_dtype, his, p->id_dtype, p->n_rows, p->w_proj, proj, nb * p->H, p->Dc, p->D, EPI_TANH, st);
    if (rc) return rc;
    // step 2b: logits, mask fill, softmax, weighted sum (model.py:174-182)
    if (stages & 2) rc = launch_poly_softmax_wsum(proj, p->codes, msk, bias, nullptr, p->table, p->table_dtype, his, p->id_dtype, p->n_rows, nb,
                                  p->H, p->K, p->Dc, p->D, interests, nullptr, i_bf16, st);
    if (rc) return rc;
    // step 3a: P = gelu(I Wt^T)                        (model.py:212)
    if (weighted && (stages & 4)) {
      if (tensor)
        rc = launch_tc_gemm(i_bf16, nullptr, p->id_dtype, 0, p->w_target_bf16, tproj, nullptr, nb * p->K, p->D, p->D, EPI_GELU, st);
      else
        rc = launch_sgemm_nt(interests, MINER_F32, nullptr, p->id_dtype, 0, p->w_target, tproj, nb * p->K, p->D, p->D, EPI_GELU, st);
      if (rc) return rc;
    }
    // step 3b+4: matching scores, target attention, per-candidate score (model.py:127-136,213-214)
    if (!(stages & 8)) continue;
    if (p->cand_offsets)
      rc = launch_target_score(interests, tproj, nullptr, nullptr, p->table, p->table_dtype, p->cand_ids, p->id_dtype, p->n_rows,
                               p->cand_offsets + b0, nb, 0, p->K, p->D, p->score_type, p->out_scores, st);
    else
      rc = launch_target_score(interests, tproj, nullptr, nullptr, p->table, p->table_dtype,
                               static_cast<const char*>(p->cand_ids) + b0 * p->C * id_sz, p->id_dtype, p->n_rows, nullptr, nb,
                               p->C, p->K, p->D, p->score_type, p->out_scores + b0 * p->C, st);
    if (rc) return rc;
  }
  return MINER_OK;
}

extern "C" size_t miner_hist_interests_workspace_bytes(int64_t Dc) { return hist_kernel_ws_bytes(Dc); }

extern "C" int miner_hist_interests_fwd(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                                        const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H,
                                        int64_t K, int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(B >= 0 && n_rows > 0, "hist_interests: bad sizes");
  if (B == 0) return MINER_OK;
  MINER_CHECK_ARG(table && his_ids && his_mask && w_proj_bf16 && codes && i_hi && i_lo, "hist_interests: null pointer");
  MINER_CHECK_ARG(id_dtype == MINER_I32 || id_dtype == MINER_I64, "hist_interests: id dtype must be int32 or int64");
  if (!workspace || workspace_bytes < hist_kernel_ws_bytes(Dc)) {
    set_error("hist_interests: workspace too small (%zu bytes needed)", hist_kernel_ws_bytes(Dc));
    return MINER_ERR_WORKSPACE;
  }
  return launch_hist_kernel(table, n_rows, his_ids, id_dtype, his_mask, bias_mean, w_proj_bf16, codes, B, H, K, Dc, D, i_hi, i_lo,
                            out_interests, static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

extern "C" int miner_cand_score_fwd(const void* i_hi, const void* i_lo, const void* w_target_bf16, const void* table, int64_t n_rows,
                                    const void* cand_ids, int id_dtype, const int64_t* cand_offsets, int64_t B, int64_t C, int64_t K,
                                    int64_t D, float* out_scores, void* stream) {
  MINER_CHECK_ARG(B >= 0 && n_rows > 0 && (cand_offsets || C > 0), "cand_score: bad sizes");
  if (B == 0) return MINER_OK;
  MINER_CHECK_ARG(i_hi && i_lo && w_target_bf16 && table && cand_ids && out_scores, "cand_score: null pointer");
  MINER_CHECK_ARG(id_dtype == MINER_I32 || id_dtype == MINER_I64, "cand_score: id dtype must be int32 or int64");
  return launch_cand_kernel(i_hi, i_lo, w_target_bf16, table, n_rows, cand_ids, id_dtype, cand_offsets, B, C, K, D, out_scores,
                            static_cast<cudaStream_t>(stream));
}

extern "C" size_t miner_table_project_workspace_bytes(int64_t n_rows, int64_t Dc) { return table_project_ws_bytes(n_rows, Dc); }

extern "C" int miner_table_project(const void* table, int64_t n_rows, int64_t D, const void* w_proj_bf16, const float* codes,
                                   const void* w_target_bf16, int64_t K, int64_t Dc, float* out_lg, void* out_tw, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(n_rows >= 0 && D > 0 && K > 0 && Dc > 0, "table_project: bad sizes");
  if (n_rows == 0) return MINER_OK;
  MINER_CHECK_ARG(table && w_proj_bf16 && codes && out_lg && (!out_tw || w_target_bf16), "table_project: null pointer");
  if (!workspace || workspace_bytes < table_project_ws_bytes(n_rows, Dc)) {
    set_error("table_project: workspace too small (%zu bytes needed)", table_project_ws_bytes(n_rows, Dc));
    return MINER_ERR_WORKSPACE;
  }
  return launch_table_project(table, n_rows, D, w_proj_bf16, codes, w_target_bf16, K, Dc, out_lg, out_tw, static_cast<float*>(workspace),
                              static_cast<cudaStream_t>(stream));
}

extern "C" int miner_score_table_supported(int64_t H, int64_t K, int64_t D) { return tscore_kernel_supported(H, K, D) ? 1 : 0; }

extern "C" size_t miner_score_table_workspace_bytes(int64_t B, int64_t H, int64_t K) { return tscore_ws_bytes(B, H, K); }

extern "C" int miner_score_table_tile_geometry(int64_t H, int64_t K, int* impressions_per_tile, int* halves) {
  return tscore_tile_geometry(H, K, impressions_per_tile, halves);
}

extern "C" int miner_score_table_fwd(const void* table, const void* tw, const float* lg, int64_t n_rows, const void* his_ids,
                                     const uint8_t* his_mask, const void* cand_ids, const int64_t* cand_offsets, int id_dtype,
                                     const float* bias_mean, int64_t B, int64_t H, int64_t C, int64_t K, int64_t D, int score_type,
                                     float* out_scores, float* out_interests, void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(B >= 0 && n_rows > 0 && H > 0 && K > 0 && D > 0 && (cand_offsets || C > 0), "score_table: bad sizes");
  if (score_type != MINER_SCORE_MAX && score_type != MINER_SCORE_MEAN && score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  if (B == 0) return MINER_OK;
  MINER_CHECK_ARG(table && lg && his_ids && his_mask && cand_ids && out_scores, "score_table: null pointer");
  MINER_CHECK_ARG(tw || score_type != MINER_SCORE_WEIGHTED, "score_table: 'weighted' needs the tw table");
  MINER_CHECK_ARG(id_dtype == MINER_I32 || id_dtype == MINER_I64, "score_table: id dtype must be int32 or int64");
  return launch_tscore_kernel(table, tw ? tw : table, lg, n_rows, his_ids, id_dtype, his_mask, bias_mean, cand_ids, cand_offsets, B, H, C, K,
                              D, score_type, out_scores, out_interests, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int miner_auc_split(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B, int64_t T, int transform,
                               uint32_t* pos_keys, uint32_t* neg_keys, uint64_t* counts, void* stream) {
  MINER_CHECK_ARG(T >= 0 && B >= 0 && transform >= 0 && transform <= 2, "auc_split: bad arguments");
  MINER_CHECK_ARG(counts && (T == 0 || (scores && labels && pos_keys && neg_keys)), "auc_split: null pointer");
  MINER_CHECK_ARG(transform != 2 || offsets, "auc_split: the softmax transform needs the impression offsets");
  return launch_auc_split(scores, labels, offsets, B, T, transform, pos_keys, neg_keys, reinterpret_cast<unsigned long long*>(counts),
                          static_cast<cudaStream_t>(stream));
}

extern "C" size_t miner_sort_u32_workspace_bytes(int64_t n) { return sort_u32_ws_bytes(n); }

extern "C" int miner_sort_u32(uint32_t* keys, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(n >= 0 && (n == 0 || keys), "sort_u32: bad arguments");
  return launch_sort_u32(keys, n, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int miner_auc_count(const uint32_t* pos_sorted, int64_t P, const uint32_t* neg_keys, int64_t N, uint64_t* out_u2, void* stream) {
  MINER_CHECK_ARG(P >= 0 && N >= 0 && out_u2 && (P == 0 || pos_sorted) && (N == 0 || neg_keys), "auc_count: bad arguments");
  return launch_auc_count(pos_sorted, P, neg_keys, N, reinterpret_cast<unsigned long long*>(out_u2), static_cast<cudaStream_t>(stream));
}

#if defined(MINER_HIST_PROF) || defined(MINER_TS_PROF)
// instrumented builds only (scripts/prof_hist.py, scripts/prof_tscore.py): device buffer of 148*5*16 int64 cycle counters.
// The release library exports nothing outside include/miner_b200.h.
extern "C" void miner_debug_set_hist_prof(void* p) { set_hist_prof_buffer(static_cast<long long*>(p)); }
#endif
