// (a13,a14) losses of the train / eval-loss variant.
//   Loss.compute           (reference src/loss.py:27-44):  mean_{b,k,l}( cos(I_k, I_l), diagonal zeroed ) + CE_mean(logits, argmax(labels))
//   Loss.compute_eval_loss (src/loss.py:68-85):            same disagreement + ( -sum logsigmoid(logits) * labels )
// One CTA per impression: the K interest vectors are L2-normalised into shared memory (utils.py:21-23 order: divide
// first, then dot), the off-diagonal cosines are summed as |sum_k u_k|^2 - sum_k |u_k|^2; the row's cross-entropy / logsigmoid term
// is computed by warp 0.  Per-impression partials are reduced by a single block in a fixed order (deterministic).
#include "common.cuh"

namespace miner {

constexpr int LT = 256;

__global__ void __launch_bounds__(LT) loss_rows_kernel(const float* __restrict__ interests, const float* __restrict__ logits,
                                                       const float* __restrict__ labels, int C, int K, int D, int mode,
                                                       float* __restrict__ row_disagree, float* __restrict__ row_rank) {
  extern __shared__ __align__(16) float smem[];
  const int DP = D + 1;
  float* In = smem;                    // [K][D+1] normalised interests
  __shared__ float wsum[LT / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const float* Ib = interests + b * static_cast<int64_t>(K) * D;
  for (int k = warp; k < K; k += LT / 32) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = Ib[static_cast<int64_t>(k) * D + d]; ss = fmaf(v, v, ss); }
    const float nrm = sqrtf(warp_sum(ss));
    for (int d = lane; d < D; d += 32) In[k * DP + d] = Ib[static_cast<int64_t>(k) * D + d] / nrm;
  }
  __syncthreads();
  // sum over the off-diagonal pairs (zero_diagonal=True, utils.py:24-27):  sum_{k != l} u_k . u_l = |sum_k u_k|^2 - sum_k |u_k|^2
  // -- O(K D) instead of the K^2 D of the pairwise form; the diagonal terms are the COMPUTED |u_k|^2, not 1
  float local = 0.f;
  for (int d = tid; d < D; d += LT) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < K; ++k) { const float u = In[k * DP + d]; s += u; q = fmaf(u, u, q); }
    local += s * s - q;
  }
  local = warp_sum(local);
  if (lane == 0) wsum[warp] = local;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < LT / 32; ++w) s += wsum[w];
    row_disagree[b] = s;
  }
  if (warp == 0) {
    const float* lg = logits + b * C;
    const float* lb = labels + b * C;
    if (mode == 0) {
      // targets = labels.argmax(dim=1) (first maximum), CrossEntropyLoss: logsumexp - logit[target]
      float best = -INFINITY; int arg = 0;
      for (int c = 0; c < C; ++c) { const float v = lb[c]; if (v > best) { best = v; arg = c; } }
      float mx = -INFINITY;
      for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
      mx = warp_max(mx);
      float se = 0.f;
      for (int c = lane; c < C; c += 32) se += expf(lg[c] - mx);
      se = warp_sum(se);
      if (lane == 0) row_rank[b] = (logf(se) + mx) - lg[arg];
    } else {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float x = lg[c];
        const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));     // logsigmoid
        s += ls * lb[c];
      }
      s = warp_sum(s);
      if (lane == 0) row_rank[b] = -s;
    }
  }
}

__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ row_disagree, const float* __restrict__ row_rank,
                                                             int64_t B, int K, int mode, float* __restrict__ out) {
  __shared__ double sd[1024], sr[1024];
  double d = 0.0, r = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += 1024) { d += row_disagree[i]; r += row_rank[i]; }
  sd[threadIdx.x] = d; sr[threadIdx.x] = r;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sd[threadIdx.x] += sd[threadIdx.x + s]; sr[threadIdx.x] += sr[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double dis = sd[0] / (static_cast<double>(B) * K * K);            // .mean() over (B,K,K), loss.py:39
    const double rank = mode == 0 ? sr[0] / static_cast<double>(B) : sr[0];  // CE reduction='mean' (trainer.py:303) / sum
    out[0] = static_cast<float>(dis + rank);
    out[1] = static_cast<float>(dis);
    out[2] = static_cast<float>(rank);
  }
}

}  // namespace miner

extern "C" size_t miner_loss_workspace_bytes(int64_t B, int64_t K) {
  (void)K;
  return sizeof(float) * 2 * static_cast<size_t>(B > 0 ? B : 1);
}

extern "C" int miner_loss_fwd(const float* interests, const float* logits, const float* labels, int64_t B, int64_t C, int64_t K,
                              int64_t D, int mode, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace miner;
  MINER_CHECK_ARG(interests && logits && labels && out, "loss: null pointer");
  MINER_CHECK_ARG(B > 0 && C > 0 && K > 0 && D > 0, "loss: bad sizes");
  MINER_CHECK_ARG(mode == 0 || mode == 1, "loss: mode must be 0 (compute) or 1 (compute_eval_loss)");
  if (!workspace || workspace_bytes < miner_loss_workspace_bytes(B, K)) {
    set_error("loss: workspace too small (%zu bytes needed)", miner_loss_workspace_bytes(B, K));
    return MINER_ERR_WORKSPACE;
  }
  const size_t smem = sizeof(float) * static_cast<size_t>(K) * (D + 1);
  if (smem > 220 * 1024) {
    set_error("loss: K=%lld D=%lld needs %zu bytes of shared memory", (long long)K, (long long)D, smem);
    return MINER_ERR_UNSUPPORTED;
  }
  auto st = static_cast<cudaStream_t>(stream);
  float* row_d = static_cast<float*>(workspace);
  float* row_r = row_d + B;
  MINER_CUDA_OK(cudaFuncSetAttribute(loss_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  loss_rows_kernel<<<static_cast<unsigned>(B), LT, smem, st>>>(interests, logits, labels, (int)C, (int)K, (int)D, mode, row_d, row_r);
  MINER_LAUNCH_OK("loss_rows");
  loss_finalize_kernel<<<1, 1024, 0, st>>>(row_d, row_r, B, (int)K, mode, out);
  MINER_LAUNCH_OK("loss_finalize");
  return MINER_OK;
}
