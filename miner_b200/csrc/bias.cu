// (a2) category-similarity attention bias (reference src/model/model.py:113-120, src/utils.py:9-29):
//   bias[b,h,c] = (x_h / ||x_h||) . (y_c / ||y_c||),  x = cat_emb[his_cat[b,h]], y = cat_emb[cand_cat[b,c]]
//   bias_mean[b,h] = mean_c bias[b,h,c]                                  (model.py:176)
// The zero padding row of nn.Embedding gives 0/0 = NaN exactly as the reference (a candidate with the pad category
// NaNs the whole row; history pads are later overwritten by the 1e-30 mask fill).  One CTA per impression; rows are
// L2-normalised into shared memory first (same operation order as utils.py:21-23), then one thread per (h,c) pair.
#include "common.cuh"

namespace miner {

constexpr int BT = 256;

__global__ void __launch_bounds__(BT) category_bias_kernel(const float* __restrict__ cat_emb, int64_t n_cat, int Ec,
                                                           const void* __restrict__ his_cat, const void* __restrict__ cand_cat,
                                                           int id_dtype, int H, int C, float* __restrict__ bias_full,
                                                           float* __restrict__ bias_mean) {
  extern __shared__ __align__(16) float smem[];
  const int EP = Ec + 1;
  float* hs = smem;                 // [H][Ec+1] normalised history category vectors
  float* cs = hs + H * EP;          // [C][Ec+1] normalised candidate category vectors
  float* bs = cs + C * EP;          // [H][C]    cosine tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;

  for (int r = warp; r < H + C; r += BT / 32) {
    const bool is_h = r < H;
    int64_t id = is_h ? load_id(his_cat, b * H + r, id_dtype) : load_id(cand_cat, b * C + (r - H), id_dtype);
    const bool ok = id >= 0 && id < n_cat;
    const float* src = cat_emb + (ok ? id : 0) * static_cast<int64_t>(Ec);
    float ss = 0.f;
    for (int e = lane; e < Ec; e += 32) { const float v = ok ? src[e] : 0.f; ss = fmaf(v, v, ss); }
    const float nrm = sqrtf(warp_sum(ss));                       // utils.py:21-22
    float* dst = is_h ? hs + r * EP : cs + (r - H) * EP;
    for (int e = lane; e < Ec; e += 32) dst[e] = (ok ? src[e] : 0.f) / nrm;   // torch.div(x, x_norm); 0/0 -> NaN
  }
  __syncthreads();
  for (int p = tid; p < H * C; p += BT) {
    const int h = p / C, c = p - h * C;
    const float* x = hs + h * EP;
    const float* y = cs + c * EP;
    float s = 0.f;
    for (int e = 0; e < Ec; ++e) s = fmaf(x[e], y[e], s);        // utils.py:23
    bs[p] = s;
    if (bias_full) bias_full[(b * H + h) * C + c] = s;
  }
  __syncthreads();
  for (int h = tid; h < H; h += BT) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += bs[h * C + c];
    bias_mean[b * H + h] = s / static_cast<float>(C);            // model.py:176
  }
}

}  // namespace miner

extern "C" int miner_category_bias(const float* cat_emb, int64_t n_cat, int64_t ec, const void* his_cat, const void* cand_cat,
                                   int id_dtype, int64_t B, int64_t H, int64_t C, float* bias_full, float* bias_mean,
                                   void* stream) {
  using namespace miner;
  MINER_CHECK_ARG(cat_emb && his_cat && cand_cat && bias_mean, "category_bias: null pointer");
  MINER_CHECK_ARG(B >= 0 && H > 0 && C > 0 && ec > 0 && n_cat > 0, "category_bias: bad sizes");
  if (B == 0) return MINER_OK;
  const size_t smem = sizeof(float) * ((H + C) * (ec + 1) + H * C);
  if (smem > 200 * 1024) {
    set_error("category_bias: H=%lld C=%lld Ec=%lld needs %zu bytes of shared memory", (long long)H, (long long)C, (long long)ec, smem);
    return MINER_ERR_UNSUPPORTED;
  }
  MINER_CUDA_OK(cudaFuncSetAttribute(category_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  category_bias_kernel<<<static_cast<unsigned>(B), BT, smem, static_cast<cudaStream_t>(stream)>>>(
      cat_emb, n_cat, (int)ec, his_cat, cand_cat, id_dtype, (int)H, (int)C, bias_full, bias_mean);
  MINER_LAUNCH_OK("category_bias");
  return MINER_OK;
}
