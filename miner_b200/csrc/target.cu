// (a4,a5) click predictor: matching scores + candidate-aware (target) attention
//   matching[c,k] = cand[c,:] . I[k,:]                               reference model.py:127
//   'max' / 'mean' over k                                            model.py:128-131
//   'weighted' (TargetAwareAttention.forward, model.py:200-216):
//       P = gelu(I Wt^T)   -- produced by the projection GEMM, passed in as `proj`
//       w[c,:] = softmax_k(cand[c,:] . P[k,:]) ; score[c] = sum_k w[c,k] matching[c,k]
// One CTA per impression.  Candidate rows are staged in shared memory in chunks (dense fp32 rows or gathered from the
// embedding table through cand_ids, fusing step (a1)); each warp owns a context code k, keeps I[k,:] and P[k,:]
// in registers (D/32 values per lane) and produces both dot products per candidate with warp-shuffle reductions;
// the softmax over K and the weighted sum are one warp per candidate.  CSR offsets give variable candidate counts.
#include "common.cuh"

namespace miner {

constexpr int TT = 512;   // threads per CTA
constexpr int CC = 16;    // candidates staged per chunk

template <int DPL>
__global__ void __launch_bounds__(TT) target_score_kernel(
    const float* __restrict__ interests, const float* __restrict__ proj, const float* __restrict__ matching,
    const float* __restrict__ cand, const void* __restrict__ table, int table_dtype, const void* __restrict__ cand_ids, int id_dtype, int64_t n_rows,
    const int64_t* __restrict__ cand_offsets, int64_t C, int K, int D, int score_type, float* __restrict__ out_scores, int proj_preact) {
  extern __shared__ __align__(16) float smem[];
  float* cand_s = smem;                       // [CC][D]
  float* m_s = cand_s + CC * D;               // [CC][K+1]
  float* a_s = m_s + CC * (K + 1);            // [CC][K+1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const int64_t c_begin = cand_offsets ? cand_offsets[b] : b * C;
  const int64_t c_end = cand_offsets ? cand_offsets[b + 1] : (b + 1) * C;
  const float* Ib = interests + b * static_cast<int64_t>(K) * D;
  const float* Pb = proj ? proj + b * static_cast<int64_t>(K) * D : nullptr;
  const bool weighted = score_type == MINER_SCORE_WEIGHTED;

  for (int64_t c0 = c_begin; c0 < c_end; c0 += CC) {
    const int nc = static_cast<int>(c_end - c0 < CC ? c_end - c0 : CC);
    // stage the chunk's candidate vectors (fp32 in smem); bf16 table rows with 16-byte loads, eight features per request
    if (!cand && table_dtype == MINER_BF16 && (D & 7) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0) {
      const int nv = D >> 3;
      for (int i = tid; i < nc * nv; i += TT) {
        const int c = i / nv, v = i - c * nv;
        const int64_t id = load_id(cand_ids, c0 + c, id_dtype);
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (id >= 0 && id < n_rows) raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(table) + id * D) + v);
        float4* dst = reinterpret_cast<float4*>(cand_s + c * D + 8 * v);
        dst[0] = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                             __uint_as_float(raw.y & 0xffff0000u));
        dst[1] = make_float4(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u), __uint_as_float(raw.w << 16),
                             __uint_as_float(raw.w & 0xffff0000u));
      }
    } else
    for (int i = tid; i < nc * D; i += TT) {
      const int c = i / D, d = i - c * D;
      float v;
      if (cand) {
        v = cand[(c0 + c) * D + d];
      } else {
        const int64_t id = load_id(cand_ids, c0 + c, id_dtype);
        if (id < 0 || id >= n_rows) v = 0.f;
        else if (table_dtype == MINER_F32) v = static_cast<const float*>(table)[id * D + d];
        else v = bf16_bits_to_float(static_cast<const uint16_t*>(table)[id * D + d]);
      }
      cand_s[c * D + d] = v;
    }
    __syncthreads();
    for (int k = warp; k < K; k += TT / 32) {
      float iv[DPL], pv[DPL];
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        const int d = lane + j * 32;
        iv[j] = d < D ? Ib[static_cast<int64_t>(k) * D + d] : 0.f;
        pv[j] = (weighted && d < D) ? Pb[static_cast<int64_t>(k) * D + d] : 0.f;
        if (proj_preact) pv[j] = gelu_erf(pv[j]);         // `proj` is I Wt^T itself (train forward keeps it for the backward): gelu here, model.py:212
      }
      for (int c = 0; c < nc; ++c) {
        float m = 0.f, a = 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
          const int d = lane + j * 32;
          const float x = d < D ? cand_s[c * D + d] : 0.f;
          m = fmaf(x, iv[j], m);
          a = fmaf(x, pv[j], a);
        }
        m = matching ? matching[(c0 + c) * K + k] : warp_sum(m);   // caller-supplied `value` (model.py:200) or cand . I_k
        if (weighted) a = warp_sum(a);
        if (lane == 0) {
          m_s[c * (K + 1) + k] = m;
          a_s[c * (K + 1) + k] = a;
        }
      }
    }
    __syncthreads();
    for (int c = warp; c < nc; c += TT / 32) {
      const float* mrow = m_s + c * (K + 1);
      const float* arow = a_s + c * (K + 1);
      float score;
      if (weighted) {
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, arow[k]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int k = lane; k < K; k += 32) se += expf(arow[k] - mx);
        se = warp_sum(se);
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s = fmaf(expf(arow[k] - mx) / se, mrow[k], s);   // model.py:213-214
        score = warp_sum(s);
      } else if (score_type == MINER_SCORE_MAX) {
        float mx = -INFINITY;
        bool has_nan = false;
        for (int k = lane; k < K; k += 32) { mx = fmaxf(mx, mrow[k]); has_nan |= isnan(mrow[k]); }
        score = warp_max(mx);
        if (__any_sync(0xffffffffu, has_nan)) score = NAN;      // torch.max propagates NaN
      } else {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) s += mrow[k];
        score = warp_sum(s) / static_cast<float>(K);
      }
      if (lane == 0) out_scores[c0 + c] = score;
    }
    __syncthreads();
  }
}

int launch_target_score(const float* interests, const float* proj, const float* matching, const float* cand, const void* table,
                        int table_dtype,
                        const void* cand_ids, int id_dtype, int64_t n_rows, const int64_t* cand_offsets, int64_t B, int64_t C,
                        int64_t K, int64_t D, int score_type, float* out_scores, cudaStream_t stream, bool proj_is_preactivation) {
  if (B == 0) return MINER_OK;
  if (score_type != MINER_SCORE_MAX && score_type != MINER_SCORE_MEAN && score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  if (D < 1 || D > 1024 || K < 1 || K > 1024) {
    set_error("target score: unsupported shape K=%lld D=%lld (need D<=1024)", (long long)K, (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  MINER_CHECK_ARG(score_type != MINER_SCORE_WEIGHTED || proj != nullptr, "target score: 'weighted' needs the gelu projection");
  const size_t smem = sizeof(float) * (static_cast<size_t>(CC) * D + 2 * static_cast<size_t>(CC) * (K + 1));
#define MINER_TGT(DPL)                                                                                                   \
  do {                                                                                                                   \
    MINER_CUDA_OK(cudaFuncSetAttribute(target_score_kernel<DPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    target_score_kernel<DPL><<<static_cast<unsigned>(B), TT, smem, stream>>>(                                            \
        interests, proj, matching, cand, table, table_dtype, cand_ids, id_dtype, n_rows, cand_offsets, C, (int)K, (int)D,          \
        score_type, out_scores, proj_is_preactivation ? 1 : 0);                                                          \
  } while (0)
  if (D <= 256) MINER_TGT(8);
  else if (D <= 512) MINER_TGT(16);
  else if (D <= 768) MINER_TGT(24);
  else MINER_TGT(32);
#undef MINER_TGT
  MINER_LAUNCH_OK("target_score");
  return MINER_OK;
}

}  // namespace miner
