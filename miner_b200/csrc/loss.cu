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

// The same with the interest rows in registers (D a multiple of 128, 16-byte aligned rows): a warp takes rows k = warp, warp + 8, ...,
// reads each once with 16-byte loads, normalises it in registers and adds it to its own column sums  s_w[d] = sum u_k[d],
// q_w[d] = sum u_k[d]^2;  the eight warps' sums meet in shared memory in warp order.
template <int NV>
__global__ void __launch_bounds__(LT) loss_rows_reg_kernel(const float* __restrict__ interests, const float* __restrict__ logits,
                                                           const float* __restrict__ labels, int C, int K, int mode,
                                                           float* __restrict__ row_disagree, float* __restrict__ row_rank) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D = NV * 128;
  float4* S = reinterpret_cast<float4*>(smem);              // [8 warps][D / 4]
  float4* Q = S + (LT / 32) * (D / 4);
  __shared__ float wsum[LT / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const float4* Ib = reinterpret_cast<const float4*>(interests + b * static_cast<int64_t>(K) * D);
  float4 s_acc[NV], q_acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) s_acc[j] = q_acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = warp; k < K; k += LT / 32) {
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j] = Ib[k * (D / 4) + lane + 32 * j];
      ss = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, ss))));
    }
    const float nrm = sqrtf(warp_sum(ss));
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 u = make_float4(v[j].x / nrm, v[j].y / nrm, v[j].z / nrm, v[j].w / nrm);      // utils.py:21-23: divide first
      s_acc[j].x += u.x, s_acc[j].y += u.y, s_acc[j].z += u.z, s_acc[j].w += u.w;
      q_acc[j].x = fmaf(u.x, u.x, q_acc[j].x), q_acc[j].y = fmaf(u.y, u.y, q_acc[j].y);
      q_acc[j].z = fmaf(u.z, u.z, q_acc[j].z), q_acc[j].w = fmaf(u.w, u.w, q_acc[j].w);
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    S[warp * (D / 4) + lane + 32 * j] = s_acc[j];
    Q[warp * (D / 4) + lane + 32 * j] = q_acc[j];
  }
  __syncthreads();
  float local = 0.f;
  for (int d4 = tid; d4 < D / 4; d4 += LT) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int w = 0; w < LT / 32; ++w) {
      const float4 a = S[w * (D / 4) + d4], c = Q[w * (D / 4) + d4];
      s.x += a.x, s.y += a.y, s.z += a.z, s.w += a.w;
      q.x += c.x, q.y += c.y, q.z += c.z, q.w += c.w;
    }
    local += (s.x * s.x - q.x) + (s.y * s.y - q.y) + (s.z * s.z - q.z) + (s.w * s.w - q.w);
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
  if (D % 128 == 0 && D <= 1024 && reinterpret_cast<uintptr_t>(interests) % 16 == 0) {
    const size_t sm2 = sizeof(float) * 2 * (LT / 32) * static_cast<size_t>(D);
    switch (D / 128) {      // rows in registers (see loss_rows_reg_kernel)
#define MINER_LR(NVV)                                                                                                             \
  case NVV:                                                                                                                       \
    MINER_CUDA_OK(cudaFuncSetAttribute(loss_rows_reg_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));        \
    loss_rows_reg_kernel<NVV><<<static_cast<unsigned>(B), LT, sm2, st>>>(interests, logits, labels, (int)C, (int)K, mode, row_d, row_r); \
    break
      MINER_LR(1); MINER_LR(2); MINER_LR(3); MINER_LR(4); MINER_LR(5); MINER_LR(6); MINER_LR(7); MINER_LR(8);
#undef MINER_LR
    }
  } else {
    MINER_CUDA_OK(cudaFuncSetAttribute(loss_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loss_rows_kernel<<<static_cast<unsigned>(B), LT, smem, st>>>(interests, logits, labels, (int)C, (int)K, (int)D, mode, row_d, row_r);
  }
  MINER_LAUNCH_OK("loss_rows");
  loss_finalize_kernel<<<1, 1024, 0, st>>>(row_d, row_r, B, (int)K, mode, out);
  MINER_LAUNCH_OK("loss_finalize");
  return MINER_OK;
}
