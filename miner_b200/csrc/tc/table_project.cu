// Table-level projections for the table-level scoring kernel (tscore_kernel.cu): the two nn.Linear layers of the path are
// applied once per news row instead of once per gathered row,
//     lg[n,k] = tanh(table[n] Wp^T) . codes[k]     PolyAttention: linear -> tanh -> context codes   (reference model.py:171,174)
//     tw[n,:] = table[n] Wt^T                      TargetAwareAttention.linear (no bias)             (model.py:198,212)
// (N rows instead of B*(H+K): 100 k rows against 82 M at BASELINE configs[1]).  Both GEMMs run on the tcgen05 projection GEMM
// (tc_gemm.cu, bf16 operands, fp32 accumulation); the logits against the context codes are an fp32 CUDA-core pass (register-tiled).
#include "fused.cuh"
#include "tc_gemm.cuh"

namespace miner {

namespace {

constexpr int LG_THREADS = 256, LG_ROWS = 64;       // a block takes 64 table rows at a time: thread = (row, group of 8 codes)

// lg[n,k] = proj[n,:] . codes[k,:] in fp32, the features summed in index order.  The block stages 64 rows of proj (coalesced) and the
// transposed codes in shared memory; a thread keeps 8 (K <= 32) or 16 accumulators of one row, so a feature step is one broadcast
// load of the row value and two (four) 16-byte loads of codes for 8 (16) FMAs.
template <bool WIDE>
__global__ void __launch_bounds__(LG_THREADS)
table_logits_kernel(const float* __restrict__ proj, const float* __restrict__ codes, int64_t n_rows, int K, int Dc, float* __restrict__ lg) {
  extern __shared__ __align__(16) float lg_smem[];
  constexpr int KC = WIDE ? 64 : 32;
  float* codes_t = lg_smem;                                    // [Dc][KC]
  const int RS = Dc | 1;                                       // odd row stride: the 8 rows of a warp fall into different banks
  float* rows = lg_smem + static_cast<size_t>(Dc) * KC;        // [LG_ROWS][RS]
  for (int i = threadIdx.x; i < Dc * KC; i += LG_THREADS) {
    const int dc = i / KC, k = i % KC;
    codes_t[i] = k < K ? codes[k * Dc + dc] : 0.f;
  }
  const int r = threadIdx.x >> 2, cg = threadIdx.x & 3;
  for (int64_t n0 = static_cast<int64_t>(blockIdx.x) * LG_ROWS; n0 < n_rows; n0 += static_cast<int64_t>(gridDim.x) * LG_ROWS) {
    const int nr = static_cast<int>(n_rows - n0 < LG_ROWS ? n_rows - n0 : LG_ROWS);
    __syncthreads();                                           // the previous tile is consumed (first pass: codes_t is written)
    const float* src = proj + n0 * Dc;
    for (int i = threadIdx.x; i < nr * Dc; i += LG_THREADS) rows[(i / Dc) * RS + i % Dc] = src[i];
    __syncthreads();
    if (r < nr) {
      float acc[WIDE ? 16 : 8];
#pragma unroll
      for (int k = 0; k < (WIDE ? 16 : 8); ++k) acc[k] = 0.f;
      const float* row = rows + r * RS;
      for (int d = 0; d < Dc; ++d) {
        const float p = row[d];
        const float4* c = reinterpret_cast<const float4*>(codes_t + d * KC + cg * 8);
        const float4 c0 = c[0], c1 = c[1];
        acc[0] = fmaf(p, c0.x, acc[0]); acc[1] = fmaf(p, c0.y, acc[1]); acc[2] = fmaf(p, c0.z, acc[2]); acc[3] = fmaf(p, c0.w, acc[3]);
        acc[4] = fmaf(p, c1.x, acc[4]); acc[5] = fmaf(p, c1.y, acc[5]); acc[6] = fmaf(p, c1.z, acc[6]); acc[7] = fmaf(p, c1.w, acc[7]);
        if (WIDE) {
          const float4 c2 = c[8], c3 = c[9];                   // codes 32 + cg * 8 ...
          acc[8] = fmaf(p, c2.x, acc[8]); acc[9] = fmaf(p, c2.y, acc[9]); acc[10] = fmaf(p, c2.z, acc[10]); acc[11] = fmaf(p, c2.w, acc[11]);
          acc[12] = fmaf(p, c3.x, acc[12]); acc[13] = fmaf(p, c3.y, acc[13]); acc[14] = fmaf(p, c3.z, acc[14]); acc[15] = fmaf(p, c3.w, acc[15]);
        }
      }
      float* out = lg + (n0 + r) * K;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (cg * 8 + k < K) out[cg * 8 + k] = acc[k];
      if (WIDE) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (32 + cg * 8 + k < K) out[32 + cg * 8 + k] = acc[8 + k];
      }
    }
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
  const int64_t blocks = (n_rows + LG_ROWS - 1) / LG_ROWS;
  const int grid = static_cast<int>(blocks < 4 * sm_count() ? blocks : 4 * sm_count());
  const bool wide = K > 32;
  const size_t smem = sizeof(float) * (static_cast<size_t>(Dc) * (wide ? 64 : 32) + static_cast<size_t>(LG_ROWS) * (Dc | 1));
  if (smem > 200 * 1024) {
    set_error("table_project: Dc=%lld too large for the logits pass", (long long)Dc);
    return MINER_ERR_UNSUPPORTED;
  }
  auto kern = wide ? table_logits_kernel<true> : table_logits_kernel<false>;
  MINER_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<grid, LG_THREADS, smem, stream>>>(proj_ws, codes, n_rows, static_cast<int>(K), static_cast<int>(Dc), out_lg);   // model.py:174
  MINER_LAUNCH_OK("table_logits_kernel");
  if (out_tw) {
    rc = launch_tc_gemm(table, nullptr, MINER_I64, 0, w_target_bf16, nullptr, out_tw, n_rows, D, D, EPI_NONE, stream);        // model.py:212
    if (rc) return rc;
  }
  return MINER_OK;
}

}  // namespace miner
