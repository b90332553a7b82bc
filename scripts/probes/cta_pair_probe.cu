// Probe: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, B split across the two CTAs) and whether cta_group::1 MMAs
// may be mixed into the same kernel / TMEM allocation.  2 CTAs in one cluster, operands written to smem by the threads.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc/umma.cuh"
using namespace miner;

constexpr int N2 = 64;          // N of the pair MMA (each CTA holds 32 rows of B)
constexpr int KK = 64;          // K (4 MMAs of 16)

__device__ __forceinline__ void umma_bf16_2cta(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D2, float* D1, int mix) {
  // A: [256][64], B: [64][64] fp32 in global; D2: [256][64] pair result; D1: [256][32] per-CTA result of A_r (128 rows) x B_r (32 rows)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_t = smem;                 // 128 x 64 bf16 K-major SW128 (16 KB)
  uint8_t* b_t = smem + 16384;         // 32 x 64 bf16 (4 KB)
  __shared__ uint64_t bar2, bar1;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = tc::cluster_ctarank();
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(a_t + tc::sw128_offset(r, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(A[(rank * 128 + r) * 64 + k]);
  }
  for (int i = threadIdx.x; i < 32 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(b_t + tc::sw128_offset(n, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(B[(rank * 32 + n) * 64 + k]);
  }
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) { tc::mbar_init(&bar2, 1); tc::mbar_init(&bar1, 1); tc::fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tbase)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  tc::cluster_sync_all();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tbase;
  if (warp == 0) {
    if (rank == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(256, N2);
      const uint64_t ad = tc::make_smem_desc_sw128(tc::smem_u32(a_t)), bd = tc::make_smem_desc_sw128(tc::smem_u32(b_t));
      if (tc::elect_one()) {
        for (int k = 0; k < KK / 16; ++k) umma_bf16_2cta(tmem, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
        commit_2cta(&bar2, 3);
      }
      __syncwarp();
    }
  }
  tc::mbar_wait(&bar2, 0);
  tc::tcgen05_fence_after();
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  {
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      tc::tmem_ld_32x32(tmem + lane_addr + c * 32, v);
      tc::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D2[(rank * 128 + threadIdx.x) * N2 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  if (mix) {
    // per-CTA cta_group::1 MMA into columns [64,96) of the pair allocation
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    if (warp == 0) {
      const uint32_t idesc = tc::make_idesc_bf16_f32(128, 32);
      const uint64_t ad = tc::make_smem_desc_sw128(tc::smem_u32(a_t)), bd = tc::make_smem_desc_sw128(tc::smem_u32(b_t));
      if (tc::elect_one()) {
        for (int k = 0; k < KK / 16; ++k) tc::umma_bf16(tmem + 64, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
        tc::umma_commit(&bar1);
      }
      __syncwarp();
    }
    tc::mbar_wait(&bar1, 0);
    tc::tcgen05_fence_after();
    uint32_t v[32];
    tc::tmem_ld_32x32(tmem + lane_addr + 64, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D1[(rank * 128 + threadIdx.x) * 32 + j] = __uint_as_float(v[j]);
  }
  tc::tcgen05_fence_before();
  tc::cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

int main(int argc, char** argv) {
  const int mix = argc > 1 ? atoi(argv[1]) : 1;
  static float hA[256 * 64], hB[64 * 64], hD2[256 * 64], hD1[256 * 32];
  for (int i = 0; i < 256 * 64; ++i) hA[i] = float((i * 7 + (i / 64) * 3) % 13 - 6) * 0.25f;
  for (int i = 0; i < 64 * 64; ++i) hB[i] = float((i * 5 + (i / 64)) % 9 - 4) * 0.5f;
  float *dA, *dB, *dD2, *dD1;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD2, sizeof(hD2)); cudaMalloc(&dD1, sizeof(hD1));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemset(dD2, 0, sizeof(hD2)); cudaMemset(dD1, 0, sizeof(hD1));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  probe<<<2, 128, 32 * 1024>>>(dA, dB, dD2, dD1, mix);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(hD2, dD2, sizeof(hD2), cudaMemcpyDeviceToHost); cudaMemcpy(hD1, dD1, sizeof(hD1), cudaMemcpyDeviceToHost);
  int bad2 = 0, bad1 = 0;
  for (int r = 0; r < 256; ++r) for (int n = 0; n < 64; ++n) {
    double ref = 0; for (int k = 0; k < 64; ++k) ref += double(hA[r * 64 + k]) * hB[n * 64 + k];
    if (fabs(ref - hD2[r * 64 + n]) > 1e-3) ++bad2;
  }
  for (int r = 0; r < 256; ++r) for (int n = 0; n < 32; ++n) {
    const int rank = r / 128;
    double ref = 0; for (int k = 0; k < 64; ++k) ref += double(hA[r * 64 + k]) * hB[(rank * 32 + n) * 64 + k];
    if (fabs(ref - hD1[r * 32 + n]) > 1e-3) ++bad1;
  }
  printf("cta_group::2 M=256 N=64: bad %d of %d; mixed cta_group::1 MMA (mix=%d): bad %d of %d\n", bad2, 256 * 64, mix, bad1, 256 * 32);
  return (bad2 || (mix && bad1)) ? 1 : 0;
}
