// Table-level projections for the table-level scoring kernel (tscore_kernel.cu): the two nn.Linear layers of the path are
// applied once per news row instead of once per gathered row,
//     lg[n,k] = tanh(table[n] Wp^T) . codes[k]     PolyAttention: linear -> tanh -> context codes   (reference model.py:171,174)
//     tw[n,:] = table[n] Wt^T                      TargetAwareAttention.linear (no bias)             (model.py:198,212)
// (N rows instead of B*(H+K): 100 k rows against 82 M at BASELINE configs[1]).  Both GEMMs run on the tcgen05 projection GEMM
// (tc_gemm.cu, bf16 operands, fp32 accumulation); the logits against the context codes are an fp32 CUDA-core pass.
#include "fused.cuh"
#include "tc_gemm.cuh"

namespace miner {

namespace {

constexpr int LG_WARPS = 8;

// one warp per table row: lane k accumulates proj[n,:] . codes[k,:] (and code k + 32 when K > 32) in fp32; codes transposed in
// shared memory
__global__ void __launch_bounds__(LG_WARPS * 32)
table_logits_kernel(const float* __restrict__ proj, const float* __restrict__ codes, int64_t n_rows, int K, int Dc, float* __restrict__ lg) {
  extern __shared__ float codes_t[];                           // [Dc][64]
  for (int i = threadIdx.x; i < Dc * 64; i += blockDim.x) {
    const int dc = i >> 6, k = i & 63;
    codes_t[i] = k < K ? codes[k * Dc + dc] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t n = static_cast<int64_t>(blockIdx.x) * LG_WARPS + warp; n < n_rows; n += static_cast<int64_t>(gridDim.x) * LG_WARPS) {
    const float* row = proj + n * Dc;
    float acc0 = 0.f, acc1 = 0.f;
    for (int d0 = 0; d0 < Dc; d0 += 32) {
      const float mine = d0 + lane < Dc ? row[d0 + lane] : 0.f;
      const int nd = Dc - d0 < 32 ? Dc - d0 : 32;
      for (int d = 0; d < nd; ++d) {
        const float p = __shfl_sync(0xffffffffu, mine, d);
        acc0 = fmaf(p, codes_t[(d0 + d) * 64 + lane], acc0);
        acc1 = fmaf(p, codes_t[(d0 + d) * 64 + 32 + lane], acc1);
      }
    }
    if (lane < K) lg[n * K + lane] = acc0;
    if (lane + 32 < K) lg[n * K + lane + 32] = acc1;
  }
}

}  // namespace

size_t table_project_ws_bytes(int64_t n_rows, int64_t Dc) { return align_up(sizeof(float) * static_cast<size_t>(n_rows) * Dc, 256); }

int launch_table_project(const void* table, int64_t n_rows, int64_t D, const void* w_proj_bf16, const float* codes,
                         const void* w_target_bf16, int64_t K, int64_t Dc, float* out_lg, void* out_tw, float* proj_ws,
                         cudaStream_t stream) {
  if (n_rows == 0) return MINER_OK;
  if (K > 64 || !tc_gemm_supported(D, Dc) || (out_tw && !tc_gemm_supported(D, D))) {
    set_error("table_project: unsupported shape K=%lld Dc=%lld D=%lld (need K <= 64, Dc >= 16, D %% 64 == 0)", (long long)K, (long long)Dc,
              (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  int rc = launch_tc_gemm(table, nullptr, MINER_I64, 0, w_proj_bf16, proj_ws, nullptr, n_rows, Dc, D, EPI_TANH, stream);     // model.py:171
  if (rc) return rc;
  const int64_t blocks = (n_rows + LG_WARPS - 1) / LG_WARPS;
  const int grid = static_cast<int>(blocks < 8 * sm_count() ? blocks : 8 * sm_count());
  const size_t smem = sizeof(float) * static_cast<size_t>(Dc) * 64;
  MINER_CUDA_OK(cudaFuncSetAttribute(table_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  table_logits_kernel<<<grid, LG_WARPS * 32, smem, stream>>>(proj_ws, codes, n_rows, static_cast<int>(K), static_cast<int>(Dc), out_lg);   // model.py:174
  MINER_LAUNCH_OK("table_logits_kernel");
  if (out_tw) {
    rc = launch_tc_gemm(table, nullptr, MINER_I64, 0, w_target_bf16, nullptr, out_tw, n_rows, D, D, EPI_NONE, stream);        // model.py:212
    if (rc) return rc;
  }
  return MINER_OK;
}

}  // namespace miner
