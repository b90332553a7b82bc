// tcgen05 projection GEMM of the MINER_MATH_TENSOR family (see tc_gemm.cu).
#pragma once
#include "../common.cuh"

namespace miner {

// K (reduction) must be a multiple of 64 and N >= 16
bool tc_gemm_supported(int64_t K, int64_t N);

// C[M,N] (fp32) = epi( A[M,K] * B[N,K]^T ), bf16 operands, fp32 accumulation in TMEM.
//   A: bf16 row-major; rows taken from `A` directly (a_ids == nullptr) or gathered: row m = A[a_ids[m], :] with the
//      ids validated against a_rows_in_table (invalid -> zero row).
//   B: bf16 (N,K) row-major weights (nn.Linear layout).   c_bf16 (nullable): bf16 copy of C.
int launch_tc_gemm(const void* A, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* B, float* C,
                   void* c_bf16, int64_t M, int64_t N, int64_t K, int epilogue, cudaStream_t stream);

// split-K form for GEMMs with few output tiles and a very long reduction (the weight gradients of the train variant): split s writes
// its fp32 partial to C + s M N (C must hold k_splits M N floats); the caller sums them in split order.  tc_gemm_splits picks a
// split count that fills the SMs and leaves no split empty.
int tc_gemm_splits(int64_t M, int64_t N, int64_t K);
int launch_tc_gemm_splitk(const void* A, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* B, float* C,
                          void* c_bf16, int64_t M, int64_t N, int64_t K, int epilogue, int k_splits, cudaStream_t stream);

// C[M,N] = A^T B over the rows: A (R,M) and B (R,N) bf16 row-major, read MN-major by the tensor cores (tc_gemm_tn.cu; the weight-gradient
// GEMMs of the train variant, no transposed operand copies).  Split-K over the rows as above: C holds k_splits partials of M N floats
// (k_splits from tc_gemm_splits(M, N, R rounded up to 64)).  Needs M % 8 == 0, N % 8 == 0, N >= 16.
bool tc_gemm_tn_supported(int64_t R, int64_t M, int64_t N);
int launch_tc_gemm_tn_splitk(const void* A, const void* B, float* C, int64_t R, int64_t M, int64_t N, int k_splits, cudaStream_t stream);

}  // namespace miner
