// Fused tensor-core scoring path (sm_100a): history kernel -> candidate kernel.  See hist_kernel.cu / cand_kernel.cu.
#pragma once
#include <cuda.h>

#include "../common.cuh"

namespace miner {

// 2D row-major bf16 tensor map, SWIZZLE_128B, box = box_rows x box_cols (box_cols * 2 bytes must be 128)
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

// History side: interests I[b,k,:] = sum_h softmax_h(tanh(E Wp^T) codes^T (+bias), mask-filled)[k,h] E[h,:]  (model.py:159-185),
// E = table[his_ids] gathered on the fly.  Writes I split as bf16 hi + lo (I = hi + lo to ~2^-17) and optionally fp32.
bool hist_kernel_supported(int64_t H, int64_t K, int64_t Dc, int64_t D);
int launch_hist_kernel(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                       const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H, int64_t K,
                       int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests, float* codes_t_ws, cudaStream_t stream);
size_t hist_kernel_ws_bytes(int64_t Dc);
// the kernel itself (hist_kernel2.cu: software-pipelined, tensor-core logits), Dc <= 208; launch_hist_kernel forwards to it
bool hist_kernel2_supported(int64_t H, int64_t K, int64_t Dc, int64_t D);
int launch_hist_kernel2(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                        const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H, int64_t K,
                        int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests, cudaStream_t stream);   // transposed, zero-padded codes [DcPad][32] fp32

void set_hist_prof_buffer(long long* p);   // -DMINER_HIST_PROF builds only: where hist_kernel2 writes its cycle counters

// Candidate side: scores for score_type = 'weighted' (model.py:127,200-216) from I_hi/I_lo and table[cand_ids].
bool cand_kernel_supported(int64_t K, int64_t D);
int launch_cand_kernel(const void* i_hi, const void* i_lo, const void* wt_bf16, const void* table, int64_t n_rows,
                       const void* cand_ids, int id_dtype, const int64_t* cand_offsets, int64_t B, int64_t C, int64_t K, int64_t D,
                       float* out_scores, cudaStream_t stream);


// Table-level mode (table_project.cu, tscore_kernel.cu): projections hoisted to the news table, one scoring kernel.
size_t table_project_ws_bytes(int64_t n_rows, int64_t Dc);
int launch_table_project(const void* table, int64_t n_rows, int64_t D, const void* w_proj_bf16, const float* codes,
                         const void* w_target_bf16, int64_t K, int64_t Dc, float* out_lg, void* out_tw, float* proj_ws,
                         cudaStream_t stream);
bool tscore_kernel_supported(int64_t H, int64_t K, int64_t D);
size_t tscore_ws_bytes(int64_t B, int64_t H, int64_t K);                 // packed tiles (tpack_kernel) + the out-of-range id counters
int tscore_tile_geometry(int64_t H, int64_t K, int* ipt, int* nh);      // impressions per tile / 128-slot halves for a shape (0 = unsupported)
int launch_tscore_kernel(const void* table, const void* tw, const float* lg, int64_t n_rows, const void* his_ids, int id_dtype,
                         const uint8_t* his_mask, const float* bias_mean, const void* cand_ids, const int64_t* cand_offsets,
                         int64_t B, int64_t H, int64_t C, int64_t K, int64_t D, int score_type, float* out_scores, float* out_interests,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace miner
