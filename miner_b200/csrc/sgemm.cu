// fp32 CUDA-core GEMM for the exact (MINER_MATH_FP32) family:
//   C[M,N] = epi( A[M,K] * B[N,K]^T ),  epi in {none, tanh, exact-erf gelu}
// used for the two nn.Linear layers of the path:
//   PolyAttention.linear   (reference model.py:171)  proj = tanh(E Wp^T)      A rows gathered from table[his_ids]
//   TargetAwareAttention.linear (model.py:212)       P    = gelu(I Wt^T)
// A may be a dense fp32 matrix or rows gathered on the fly from an fp32/bf16 embedding table (fusing step (a1)
// into the projection, so the gathered history tile is never materialised in HBM).
//
// 128x64x16 block tile, 256 threads, 8x4 outputs per thread, register-staged double buffering.
#include "common.cuh"

namespace miner {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int PAD = 4;

__device__ __forceinline__ float4 load4_f32(const float* p, int64_t k, int64_t K, bool vec) {
  if (vec && k + 3 < K) return *reinterpret_cast<const float4*>(p + k);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) r.x = p[k];
  if (k + 1 < K) r.y = p[k + 1];
  if (k + 2 < K) r.z = p[k + 2];
  if (k + 3 < K) r.w = p[k + 3];
  return r;
}
__device__ __forceinline__ float4 load4_bf16(const uint16_t* p, int64_t k, int64_t K, bool vec) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && k + 3 < K) {
    const uint2 u = *reinterpret_cast<const uint2*>(p + k);
    r.x = __uint_as_float(u.x << 16); r.y = __uint_as_float(u.x & 0xffff0000u);
    r.z = __uint_as_float(u.y << 16); r.w = __uint_as_float(u.y & 0xffff0000u);
    return r;
  }
  if (k < K) r.x = bf16_bits_to_float(p[k]);
  if (k + 1 < K) r.y = bf16_bits_to_float(p[k + 1]);
  if (k + 2 < K) r.z = bf16_bits_to_float(p[k + 2]);
  if (k + 3 < K) r.w = bf16_bits_to_float(p[k + 3]);
  return r;
}

template <int EPI, bool A_BF16>
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const void* __restrict__ A, const void* __restrict__ a_ids, int id_dtype,
                                                       int64_t a_rows_in_table, const float* __restrict__ Bm,
                                                       float* __restrict__ Cm, int64_t M, int64_t N, int64_t K, bool a_vec, bool b_vec) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * BN;

  // global->smem staging map: 4 consecutive threads cover 16 consecutive k of one row (64 contiguous bytes)
  const int lr = tid >> 2;          // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12
  const void* a_row[2];
  bool a_ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t m = m0 + lr + j * 64;
    a_ok[j] = m < M;
    int64_t row = m;
    if (a_ok[j] && a_ids) {
      row = load_id(a_ids, m, id_dtype);
      if (row < 0 || row >= a_rows_in_table) a_ok[j] = false;   // out-of-range id contributes zeros (gather semantics)
    }
    a_row[j] = A_BF16 ? static_cast<const void*>(static_cast<const uint16_t*>(A) + row * K)
                      : static_cast<const void*>(static_cast<const float*>(A) + row * K);
  }
  const bool b_ok = (n0 + lr) < N;
  const float* b_row = Bm + (n0 + lr) * K;

  float4 ra[2], rb;
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (!a_ok[j]) ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      else if (A_BF16) ra[j] = load4_bf16(static_cast<const uint16_t*>(a_row[j]), k0 + lk, K, a_vec);
      else ra[j] = load4_f32(static_cast<const float*>(a_row[j]), k0 + lk, K, a_vec);
    }
    rb = b_ok ? load4_f32(b_row, k0 + lk, K, b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[buf][lk + 0][lr + j * 64] = ra[j].x;
      As[buf][lk + 1][lr + j * 64] = ra[j].y;
      As[buf][lk + 2][lr + j * 64] = ra[j].z;
      As[buf][lk + 3][lr + j * 64] = ra[j].w;
    }
    Bs[buf][lk + 0][lr] = rb.x;
    Bs[buf][lk + 1][lr] = rb.y;
    Bs[buf][lk + 2][lr] = rb.z;
    Bs[buf][lk + 3][lr] = rb.w;
  };

  const int ty = tid >> 4;   // 0..15 -> rows ty*8
  const int tx = tid & 15;   // 0..15 -> cols tx*4
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  stage(0);
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stage(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == EPI_TANH) v = tanhf(v);
      if (EPI == EPI_GELU) v = gelu_erf(v);
      Cm[m * N + n] = v;
    }
  }
}

int launch_sgemm_nt(const void* A, int a_dtype, const void* a_ids, int id_dtype, int64_t a_rows_in_table,
                    const float* Bm, float* Cm, int64_t M, int64_t N, int64_t K, int epilogue, cudaStream_t stream) {
  if (M == 0 || N == 0) return MINER_OK;
  MINER_CHECK_ARG(K > 0, "sgemm: K must be positive");
  const dim3 grid(static_cast<unsigned>((N + BN - 1) / BN), static_cast<unsigned>((M + BM - 1) / BM));
  const int64_t a_elt = a_dtype == MINER_BF16 ? 2 : 4;
  const bool a_vec = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(A) % (4 * a_elt) == 0);
  const bool b_vec = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(Bm) % 16 == 0);
#define MINER_SGEMM(EPI, ABF)                                                                                          \
  sgemm_nt_kernel<EPI, ABF><<<grid, 256, 0, stream>>>(A, a_ids, id_dtype, a_rows_in_table, Bm, Cm, M, N, K, a_vec, b_vec)
  const bool bf = a_dtype == MINER_BF16;
  if (epilogue == EPI_NONE) { if (bf) MINER_SGEMM(EPI_NONE, true); else MINER_SGEMM(EPI_NONE, false); }
  else if (epilogue == EPI_TANH) { if (bf) MINER_SGEMM(EPI_TANH, true); else MINER_SGEMM(EPI_TANH, false); }
  else { if (bf) MINER_SGEMM(EPI_GELU, true); else MINER_SGEMM(EPI_GELU, false); }
#undef MINER_SGEMM
  MINER_LAUNCH_OK("sgemm_nt");
  return MINER_OK;
}

}  // namespace miner
