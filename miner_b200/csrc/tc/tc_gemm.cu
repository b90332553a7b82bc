// Projection GEMMs of the scoring path on the 5th-generation tensor cores (sm_100a, tcgen05 + TMEM + TMA):
//   PolyAttention.linear        proj = tanh(E Wp^T)   (reference src/model/model.py:171)  A = table rows gathered by his_ids
//   TargetAwareAttention.linear P    = gelu(I Wt^T)   (reference src/model/model.py:212)  A = bf16 interests
//
// C[M,N] = epi(A[M,K] B[N,K]^T), bf16 operands, fp32 accumulators in tensor memory.
//
// Persistent, warp-specialised kernel, one CTA per SM, 128 x 256 output tile, K pipelined in 64-element (128-byte)
// blocks through a 4-stage shared-memory ring:
//   warps 0-3  A producers: each k-block of the tile's 128 rows is fetched with 16-byte cp.async (8 lanes cover the
//              128 contiguous bytes of a row) straight into the 128B-swizzled K-major layout tcgen05 expects -- the
//              embedding gather is fused here: a row is table[his_id], never materialised in HBM;
//   warp 4     B producer: one thread issues a TMA 2D tile load (SWIZZLE_128B) of the weight block per stage;
//   warp 5     MMA issuer: one thread issues 4 x tcgen05.mma (128 x N x 16) per stage into one of two TMEM accumulators
//              and releases the stage with tcgen05.commit;
//   warps 6-9  epilogue: tcgen05.ld the finished accumulator (32 lanes x 32 columns at a time), apply tanh / exact-erf
//              gelu and store fp32 (and optionally bf16) rows, overlapping the next tile's main loop.
#include <cuda.h>

#include "tc_gemm.cuh"
#include "umma.cuh"

namespace miner {

namespace {

constexpr int GM = 128;           // tile rows   (UMMA M)
constexpr int GN = 256;           // tile cols   (UMMA N, <= 256)
constexpr int GK = 64;            // k-block: 64 bf16 = 128 bytes = one swizzle row
constexpr int STAGES = 4;
constexpr int LAG = 2;            // cp.async groups a producer thread keeps in flight before signalling
constexpr int A_BYTES = GM * GK * 2;
constexpr int B_BYTES = GN * GK * 2;
constexpr int N_PRODUCER = 128;
constexpr int TMEM_COLS = 512;    // two 256-column fp32 accumulators
constexpr int THREADS = 320;
constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + 256;

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float apply_epilogue(float v, int epi) {
  if (epi == EPI_TANH) return tanhf(v);
  if (epi == EPI_GELU) return gelu_erf(v);
  return v;
}

__global__ void __launch_bounds__(THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_b, const uint16_t* __restrict__ A, const void* __restrict__ a_ids,
               int id_dtype, int64_t a_rows_in_table, float* __restrict__ Cfull, __nv_bfloat16* __restrict__ Cb, int64_t M, int N,
               int K, int epi, int m_tiles, int n_tiles, int k_splits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_tiles = smem;
  uint8_t* b_tiles = smem + STAGES * A_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * (A_BYTES + B_BYTES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // split-K (k_splits > 1, weight-gradient GEMMs: few output tiles, very long K): work item = (tile, split); split s accumulates
  // its k-blocks into its own fp32 partial C + s M N, summed by the caller in split order
  const int kb_all = K / GK;
  const int kb_per = (kb_all + k_splits - 1) / k_splits;
  const int total_tiles = m_tiles * n_tiles * k_splits;
  auto kb_range = [&](int item, int& kb0, int& nkb) {
    const int sp = item % k_splits;
    kb0 = sp * kb_per;
    nkb = kb_all - kb0 < kb_per ? kb_all - kb0 : kb_per;
    if (nkb < 0) nkb = 0;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&bars->full[s], N_PRODUCER + 1);
      tc::mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&bars->tmem_full[a], 1);
      tc::mbar_init(&bars->tmem_empty[a], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&tmap_b);
  if (warp == 5) {
    tc::tmem_alloc(&bars->tmem_base, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers (gather fused)
    const int chunk = lane & 7;                 // 16-byte chunk of the 128-byte k-block row
    uint32_t issued = 0, signalled = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
      const int tile = item / k_splits;
      int kb0, num_kb;
      kb_range(item, kb0, num_kb);
      const int64_t m0 = static_cast<int64_t>(tile / n_tiles) * GM;
      const uint16_t* src[8];
      uint32_t nbytes[8];
      uint32_t dst_off[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 32 + j * 4 + (lane >> 3);
        const int64_t m = m0 + r;
        bool ok = m < M;
        int64_t row = ok ? m : 0;
        if (ok && a_ids) {
          row = load_id(a_ids, m, id_dtype);
          if (row < 0 || row >= a_rows_in_table) { ok = false; row = 0; }
        }
        src[j] = A + row * K + chunk * 8 + static_cast<int64_t>(kb0) * GK;
        nbytes[j] = ok ? 16u : 0u;
        dst_off[j] = static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        const uint32_t s = issued % STAGES, ph = (issued / STAGES) & 1;
        tc::mbar_wait(&bars->empty[s], ph ^ 1);
        const uint32_t base = tc::smem_u32(a_tiles + s * A_BYTES);
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::cp_async_16(base + dst_off[j], src[j] + kb * GK, nbytes[j]);
        tc::cp_async_commit();
        ++issued;
        if (issued - signalled > LAG) {
          tc::cp_async_wait<LAG>();
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&bars->full[signalled % STAGES]);
          ++signalled;
        }
      }
    }
    tc::cp_async_wait<0>();
    tc::fence_proxy_async_smem();
    while (signalled < issued) {
      tc::mbar_arrive(&bars->full[signalled % STAGES]);
      ++signalled;
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ B producer (TMA)
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / k_splits;
        int kb0, num_kb;
        kb_range(item, kb0, num_kb);
        const int n0 = (tile % n_tiles) * GN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(&bars->empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&bars->full[s], B_BYTES);
          tc::tma_load_2d(&tmap_b, &bars->full[s], tc::smem_u32(b_tiles + s * B_BYTES), (kb0 + kb) * GK, n0);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++acc_it) {
        const int tile = item / k_splits;
        int kb0, num_kb;
        kb_range(item, kb0, num_kb);
        const int n0 = (tile % n_tiles) * GN;
        int n_eff = N - n0 < GN ? N - n0 : GN;
        n_eff = (n_eff + 15) & ~15;
        const uint32_t idesc = tc::make_idesc_bf16_f32(GM, n_eff);
        const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
        tc::mbar_wait(&bars->tmem_empty[as], aph ^ 1);
        tc::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * GN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          tc::mbar_wait(&bars->full[s], ph);
          tc::tcgen05_fence_after();
          const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(a_tiles + s * A_BYTES));
          const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(b_tiles + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < GK / 16; ++k)     // advance 16 elements = 32 bytes inside the swizzle row: +2 in 16-byte units
            tc::umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(&bars->empty[s]);
        }
        tc::umma_commit(&bars->tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 6..9)
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    uint32_t acc_it = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++acc_it) {
      const int tile = item / k_splits;
      float* C = Cfull ? Cfull + static_cast<int64_t>(item % k_splits) * M * N : nullptr;
      const int64_t m0 = static_cast<int64_t>(tile / n_tiles) * GM;
      const int n0 = (tile % n_tiles) * GN;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      tc::mbar_wait(&bars->tmem_full[as], aph);
      tc::tcgen05_fence_after();
      const int64_t m = m0 + q * 32 + lane;
      const int n_cols = N - n0 < GN ? N - n0 : GN;
      const bool vec_ok = (N % 4 == 0);
      for (int c = 0; c * 32 < n_cols; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GN + c * 32, r);
        tc::tmem_ld_wait();
        if (m < M) {
          float* crow = C ? C + m * N + n0 + c * 32 : nullptr;
          __nv_bfloat16* brow = Cb ? Cb + m * N + n0 + c * 32 : nullptr;
          const int valid = n_cols - c * 32 < 32 ? n_cols - c * 32 : 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_epilogue(__uint_as_float(r[j]), epi);
          if (!C) {
          } else if (valid == 32 && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) crow[j] = v[j];
          }
          if (brow) {
            if (valid == 32 && N % 8 == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(brow + j) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                                 *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) brow[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars->tmem_empty[as]);
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

bool tc_gemm_supported(int64_t K, int64_t N) { return K >= GK && K % GK == 0 && N >= 16 && K <= (1 << 20) && N <= (1 << 20); }

int launch_tc_gemm(const void* A, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* B, float* C, void* c_bf16,
                   int64_t M, int64_t N, int64_t K, int epilogue, cudaStream_t stream) {
  return launch_tc_gemm_splitk(A, a_ids, id_dtype, a_rows_in_table, B, C, c_bf16, M, N, K, epilogue, 1, stream);
}

int tc_gemm_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ((M + GM - 1) / GM) * ((N + GN - 1) / GN);
  int64_t s = sm_count() / (tiles > 0 ? tiles : 1);
  const int64_t kb = K / GK;
  if (s > kb / 8) s = kb / 8;                  // at least 8 k-blocks per split
  if (s < 1) s = 1;
  const int64_t per = (kb + s - 1) / s;
  s = (kb + per - 1) / per;                    // no empty split
  return static_cast<int>(s);
}

int launch_tc_gemm_splitk(const void* A, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* B, float* C, void* c_bf16,
                          int64_t M, int64_t N, int64_t K, int epilogue, int k_splits, cudaStream_t stream) {
  if (M == 0) return MINER_OK;
  MINER_CHECK_ARG(k_splits >= 1 && (k_splits == 1 || (C && !c_bf16 && epilogue == EPI_NONE)), "tc_gemm: split-K needs fp32 partials and no epilogue");
  if (k_splits > 1) {
    const int64_t kb = K / GK, per = (kb + k_splits - 1) / k_splits;
    MINER_CHECK_ARG((kb + per - 1) / per == k_splits, "tc_gemm: a split without k-blocks (use tc_gemm_splits)");
  }
  MINER_CHECK_ARG(A && B && (C || c_bf16), "tc_gemm: null pointer");
  if (!tc_gemm_supported(K, N)) {
    set_error("tc_gemm: unsupported shape N=%lld K=%lld (need K %% 64 == 0, N >= 16)", (long long)N, (long long)K);
    return MINER_ERR_UNSUPPORTED;
  }
  MINER_CHECK_ARG(reinterpret_cast<uintptr_t>(A) % 16 == 0 && reinterpret_cast<uintptr_t>(B) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(C) % 16 == 0 && reinterpret_cast<uintptr_t>(c_bf16) % 16 == 0,
                  "tc_gemm: operands must be 16-byte aligned");
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) {
    set_error("tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
    return MINER_ERR_CUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(N)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(K) * 2};
  const cuuint32_t box[2] = {GK, GN};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("tc_gemm: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(cr));
    return MINER_ERR_CUDA;
  }
  const int m_tiles = static_cast<int>((M + GM - 1) / GM);
  const int n_tiles = static_cast<int>((N + GN - 1) / GN);
  const int64_t total = static_cast<int64_t>(m_tiles) * n_tiles * k_splits;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  MINER_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  tc_gemm_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmap, static_cast<const uint16_t*>(A), a_ids, id_dtype, a_rows_in_table, C,
                                                        static_cast<__nv_bfloat16*>(c_bf16), M, static_cast<int>(N),
                                                        static_cast<int>(K), epilogue, m_tiles, n_tiles, k_splits);
  MINER_LAUNCH_OK("tc_gemm");
  return MINER_OK;
}

}  // namespace miner

// generic entry used by the tests to validate the tensor-core GEMM on its own
extern "C" int miner_tc_gemm(const void* a_bf16, const void* a_ids, int id_dtype, int64_t a_rows_in_table, const void* b_bf16,
                             float* c, void* c_bf16, int64_t M, int64_t N, int64_t K, int epilogue, void* stream) {
  using namespace miner;
  MINER_CHECK_ARG(epilogue >= EPI_NONE && epilogue <= EPI_GELU, "tc_gemm: bad epilogue");
  return launch_tc_gemm(a_bf16, a_ids, id_dtype, a_rows_in_table, b_bf16, c, c_bf16, M, N, K, epilogue, static_cast<cudaStream_t>(stream));
}
