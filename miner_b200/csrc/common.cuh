// Shared helpers for the miner_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/miner_b200.h"

namespace miner {

constexpr int kWarp = 32;
constexpr float kMaskFill = 1e-30f;   // reference model.py:180

void set_error(const char* fmt, ...);
int  sm_count();
void count_launch();   // every kernel launch of the library is counted (bench.py reports it as gpu_launches)

#define MINER_CHECK_ARG(cond, ...)                                   \
  do {                                                               \
    if (!(cond)) {                                                   \
      ::miner::set_error(__VA_ARGS__);                               \
      return MINER_ERR_INVALID_ARG;                                  \
    }                                                                \
  } while (0)

#define MINER_CUDA_OK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::miner::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MINER_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

#define MINER_LAUNCH_OK(name)                                                                 \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::miner::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return MINER_ERR_CUDA;                                                                  \
    }                                                                                         \
    ::miner::count_launch();                                                                  \
  } while (0)

__device__ __forceinline__ int64_t load_id(const void* ids, int64_t i, int id_dtype) {
  return id_dtype == MINER_I64 ? reinterpret_cast<const int64_t*>(ids)[i]
                               : static_cast<int64_t>(reinterpret_cast<const int32_t*>(ids)[i]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }

// exact erf GELU, the reference's torch_f.gelu default (model.py:212)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- launchers implemented across the .cu files (all asynchronous on `stream`) ----
int launch_gather(const void* table, int64_t n_rows, int64_t dim, int dtype, const void* ids, int64_t n_ids,
                  int id_dtype, void* out, int32_t* oob_flag, cudaStream_t stream);

// C[M,N] = epi(A[M,K] * B[N,K]^T); A rows optionally gathered from a table (a_ids != null) of dtype a_dtype.
enum Epilogue { EPI_NONE = 0, EPI_TANH = 1, EPI_GELU = 2 };
int launch_sgemm_nt(const void* A, int a_dtype, const void* a_ids, int id_dtype, int64_t a_rows_in_table,
                    const float* Bm, float* Cm, int64_t M, int64_t N, int64_t K, int epilogue, cudaStream_t stream);

// logits/softmax/weighted-sum part of PolyAttention for a block of impressions.
// E is either dense fp32 (B,H,D) or gathered on the fly from table[his_ids].
int launch_poly_softmax_wsum(const float* proj, const float* codes, const uint8_t* mask, const float* bias_mean,
                             const float* emb, const void* table, int table_dtype, const void* his_ids, int id_dtype,
                             int64_t n_rows, int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D,
                             float* out_interests, float* out_weights, void* out_interests_bf16, cudaStream_t stream);

// candidate-aware aggregation + dot-product score.  cand rows dense fp32 (T,D) or gathered from table[cand_ids].
int launch_target_score(const float* interests, const float* proj /*gelu(I Wt^T) or null*/,
                        const float* matching /*(T,K) caller-supplied matching scores or null*/, const float* cand,
                        const void* table, int table_dtype, const void* cand_ids, int id_dtype, int64_t n_rows,
                        const int64_t* cand_offsets, int64_t B, int64_t C, int64_t K, int64_t D, int score_type,
                        float* out_scores, cudaStream_t stream, bool proj_is_preactivation = false);

// global AUC building blocks (auc.cu)
size_t sort_u32_ws_bytes(int64_t n);
int launch_sort_u32(uint32_t* keys, int64_t n, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_auc_split(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B, int64_t T, int transform,
                     uint32_t* pos_keys, uint32_t* neg_keys, unsigned long long* counts, cudaStream_t stream);
int launch_auc_count(const uint32_t* pos_sorted, int64_t P, const uint32_t* neg_keys, int64_t N, unsigned long long* out_u2, cudaStream_t stream);

int launch_cast_f32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t stream);

}  // namespace miner
