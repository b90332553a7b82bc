// Weight-gradient GEMMs of the train variant on the 5th-generation tensor cores (sm_100a, tcgen05 + TMEM + TMA), with both operands
// read as they lie in memory:
//   dWt[o,i] = sum_r dZ[r,o]  I[r,i]     TargetAwareAttention.linear   (reference src/model/model.py:198,212; trainer.py:246-261)
//   dWp[c,d] = sum_r dZ1[r,c] E[r,d]     PolyAttention.linear          (model.py:155,171)
//
// C[M,N] = A^T B with A (R,M) and B (R,N) bf16 ROW-major: the reduction runs over the rows, i.e. both operands are MN-major for the
// tensor cores.  A TMA box of 64 rows x 64 columns (SWIZZLE_128B) lands in shared memory as 64 rows of 128 bytes -- exactly the
// MN-major swizzle atom tcgen05.mma reads (the 64 contiguous elements of a row run along M or N, the rows along K), so no transposed
// copy of dZ / I / dZ1 / E is ever written (round 2, first half: four transpose kernels, ~0.7 ms of an 8 ms step).
//
// Persistent, warp-specialised, one CTA per SM, 128 x 256 output tile = 2 + 4 atoms per 64-row k-block, 4-stage ring:
//   warp 0     one thread issues the six TMA loads of a stage (atoms past the edge of the matrix are skipped: nothing reads them)
//   warp 1     one thread issues 4 x tcgen05.mma (128 x N x 16, both operands MN-major) per stage into one of two TMEM accumulators
//   warps 2-5  epilogue: tcgen05.ld the finished accumulator, store the fp32 partial of this split
// Split-K over the rows (work item = tile x split) because M x N is only a handful of tiles while R is 10^5: split s writes its partial
// to C + s M N, summed by the caller in split order (bit-reproducible).
#include <cuda.h>

#include "tc_gemm.cuh"
#include "umma.cuh"

namespace miner {

namespace {

constexpr int GM = 128, GN = 256;
constexpr int GR = 64;                        // rows (reduction) per k-block
constexpr int STAGES = 4;
constexpr int ATOM_BYTES = GR * 128;          // 64 rows x 64 bf16
constexpr int A_ATOMS = GM / 64, B_ATOMS = GN / 64;
constexpr int A_BYTES = A_ATOMS * ATOM_BYTES, B_BYTES = B_ATOMS * ATOM_BYTES;
constexpr int TMEM_COLS = 512;                // two 256-column fp32 accumulators
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + 256;

struct TnBarriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* __restrict__ Cfull, int M, int N,
                  int64_t R, int m_tiles, int n_tiles, int k_splits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tiles = smem;
  uint8_t* b_tiles = smem + STAGES * A_BYTES;
  TnBarriers* bars = reinterpret_cast<TnBarriers*>(smem + STAGES * (A_BYTES + B_BYTES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_all = static_cast<int>((R + GR - 1) / GR);
  const int kb_per = (kb_all + k_splits - 1) / k_splits;
  const int total_items = m_tiles * n_tiles * k_splits;
  auto kb_range = [&](int item, int& kb0, int& nkb) {
    const int sp = item % k_splits;
    kb0 = sp * kb_per;
    nkb = kb_all - kb0 < kb_per ? kb_all - kb0 : kb_per;
    if (nkb < 0) nkb = 0;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&bars->full[s], 1);
      tc::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&bars->tmem_full[a], 1);
      tc::mbar_init(&bars->tmem_empty[a], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    tc::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tile = item / k_splits;
        int kb0, num_kb;
        kb_range(item, kb0, num_kb);
        const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
        int na = 0, nb = 0;                                        // atoms that start inside the matrix
        for (int a = 0; a < A_ATOMS; ++a) na += m0 + 64 * a < M;
        for (int b = 0; b < B_ATOMS; ++b) nb += n0 + 64 * b < N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(&bars->empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&bars->full[s], static_cast<uint32_t>(na + nb) * ATOM_BYTES);
          const int r0 = (kb0 + kb) * GR;                          // rows past R are zero-filled by the TMA unit
          for (int a = 0; a < na; ++a) tc::tma_load_2d(&tmap_a, &bars->full[s], tc::smem_u32(a_tiles + s * A_BYTES + a * ATOM_BYTES), m0 + 64 * a, r0);
          for (int b = 0; b < nb; ++b) tc::tma_load_2d(&tmap_b, &bars->full[s], tc::smem_u32(b_tiles + s * B_BYTES + b * ATOM_BYTES), n0 + 64 * b, r0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
        const int tile = item / k_splits;
        int kb0, num_kb;
        kb_range(item, kb0, num_kb);
        const int n0 = (tile % n_tiles) * GN;
        int n_eff = N - n0 < GN ? N - n0 : GN;
        n_eff = (n_eff + 15) & ~15;                                // (columns past N inside a loaded atom are TMA zero fill)
        const uint32_t idesc = tc::make_idesc_bf16_f32_major(GM, n_eff, true, true);
        const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
        tc::mbar_wait(&bars->tmem_empty[as], aph ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * GN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(&bars->full[s], ph);
          tc::tcgen05_fence_after();
          // MN-major operands: atom a of the operand starts ATOM_BYTES after atom a - 1 (LBO), 8 reduction rows are 1 KB (SBO);
          // one MMA (K = 16) covers two 8-row groups = 2 KB
          const uint64_t a_desc = tc::make_smem_desc_sw128_mn_wide(tc::smem_u32(a_tiles + s * A_BYTES), ATOM_BYTES);
          const uint64_t b_desc = tc::make_smem_desc_sw128_mn_wide(tc::smem_u32(b_tiles + s * B_BYTES), ATOM_BYTES);
#pragma unroll
          for (int k = 0; k < GR / 16; ++k)
            tc::umma_bf16(d_tmem, a_desc + k * (2048 >> 4), b_desc + k * (2048 >> 4), idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(&bars->empty[s]);
        }
        tc::umma_commit(&bars->tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5: TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    uint32_t acc_it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_it) {
      const int tile = item / k_splits;
      int kb0, num_kb;
      kb_range(item, kb0, num_kb);
      float* C = Cfull + static_cast<int64_t>(item % k_splits) * M * N;
      const int m0 = (tile / n_tiles) * GM, n0 = (tile % n_tiles) * GN;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::tcgen05_fence_after();
      const int m = m0 + q * 32 + lane;
      const int n_cols = N - n0 < GN ? N - n0 : GN;
      const bool vec_ok = (N % 4 == 0);
      for (int c = 0; c * 32 < n_cols; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GN + c * 32, r);
        tc::tmem_ld_wait();
        if (m < M) {
          float* crow = C + static_cast<int64_t>(m) * N + n0 + c * 32;
          const int valid = n_cols - c * 32 < 32 ? n_cols - c * 32 : 32;
          if (num_kb == 0) {                                       // (a split without rows: its partial is zero)
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          if (valid == 32 && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(crow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                                 __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) crow[j] = __uint_as_float(r[j]);
          }
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars->tmem_empty[as]);
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn_tn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (rows, cols) bf16 row-major matrix, box = 64 columns x 64 rows
int encode_rowmajor(EncodeTiledFn encode, CUtensorMap* tmap, const void* base, int64_t rows, int64_t cols) {
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * 2};
  const cuuint32_t box[2] = {64, GR};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("tc_gemm_tn: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(cr));
    return MINER_ERR_CUDA;
  }
  return MINER_OK;
}

}  // namespace

bool tc_gemm_tn_supported(int64_t R, int64_t M, int64_t N) {
  return R >= 1 && R < (1ll << 31) && M >= 8 && N >= 16 && M % 8 == 0 && N % 8 == 0 && M <= (1 << 20) && N <= (1 << 20);
}

int launch_tc_gemm_tn_splitk(const void* A, const void* B, float* C, int64_t R, int64_t M, int64_t N, int k_splits, cudaStream_t stream) {
  MINER_CHECK_ARG(A && B && C, "tc_gemm_tn: null pointer");
  MINER_CHECK_ARG(k_splits >= 1, "tc_gemm_tn: k_splits must be positive");
  if (!tc_gemm_tn_supported(R, M, N)) {
    set_error("tc_gemm_tn: unsupported shape R=%lld M=%lld N=%lld (need M %% 8 == 0, N %% 8 == 0, N >= 16)", (long long)R, (long long)M, (long long)N);
    return MINER_ERR_UNSUPPORTED;
  }
  MINER_CHECK_ARG(reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0 && reinterpret_cast<uintptr_t>(C) % 16 == 0,
                  "tc_gemm_tn: operands must be 16-byte aligned");
  EncodeTiledFn encode = encode_tiled_fn_tn();
  if (!encode) {
    set_error("tc_gemm_tn: cuTensorMapEncodeTiled is not available from the driver");
    return MINER_ERR_CUDA;
  }
  CUtensorMap ta, tb;
  int rc = encode_rowmajor(encode, &ta, A, R, M);
  if (rc) return rc;
  rc = encode_rowmajor(encode, &tb, B, R, N);
  if (rc) return rc;
  const int m_tiles = static_cast<int>((M + GM - 1) / GM), n_tiles = static_cast<int>((N + GN - 1) / GN);
  const int64_t total = static_cast<int64_t>(m_tiles) * n_tiles * k_splits;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  MINER_CUDA_OK(cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  tc_gemm_tn_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(ta, tb, C, static_cast<int>(M), static_cast<int>(N), R, m_tiles, n_tiles, k_splits);
  MINER_LAUNCH_OK("tc_gemm_tn");
  return MINER_OK;
}

}  // namespace miner

// generic entry used by the tests to validate the kernel on its own: c holds k_splits partials of M x N floats
extern "C" int miner_tc_gemm_tn(const void* a_bf16, const void* b_bf16, float* c, int64_t R, int64_t M, int64_t N, int k_splits, void* stream) {
  using namespace miner;
  return launch_tc_gemm_tn_splitk(a_bf16, b_bf16, c, R, M, N, k_splits, static_cast<cudaStream_t>(stream));
}
