// Probe: tcgen05.mma kind::f16 with MIXED operand formats: A = fp16 in tensor memory (TS form), B = bf16 in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I miner_b200/csrc -o /tmp/mixed_probe scripts/probes/mixed_mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc/umma.cuh"
using namespace miner;

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
               "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int KS = 2;      // two k-steps (K = 32) to check the column advance
constexpr int NN = 32;

__global__ void probe(const float* A, const float* B, float* D) {   // A[128][32], B[NN][32], D[128][NN]
  __shared__ __align__(1024) uint8_t btile[NN * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  for (int i = threadIdx.x; i < NN * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(btile)[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < NN * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    *reinterpret_cast<__nv_bfloat16*>(btile + tc::sw128_offset(n, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(B[n * 32 + k]);
  }
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(&tbase, 64); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tbase;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  // A row r: 32 bf16 -> 16 packed columns [0,16)
  for (int h = 0; h < 2; ++h) {
    uint32_t pk[8];
    for (int j = 0; j < 8; ++j) {
      __half2 v = __floats2half2_rn(A[r * 32 + h * 16 + 2 * j], A[r * 32 + h * 16 + 2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&v);
    }
    tmem_st_32x8(tmem + lane_addr + h * 8, pk);
  }
  tmem_st_wait();
  tc::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::tcgen05_fence_after();
    const uint32_t idesc = (1u << 4) | (0u << 7) | (1u << 10) | (static_cast<uint32_t>(NN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);   // D f32, A f16, B bf16
    const uint64_t bdesc = tc::make_smem_desc_sw128(tc::smem_u32(btile));
    for (int k = 0; k < KS; ++k) umma_bf16_ts(tmem + 32, tmem + k * 8, bdesc + 2 * k, idesc, k ? 1u : 0u);
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tcgen05_fence_after();
  uint32_t v[32];
  tc::tmem_ld_32x32(tmem + lane_addr + 32, v);
  tc::tmem_ld_wait();
  for (int j = 0; j < NN; ++j) D[r * NN + j] = __uint_as_float(v[j]);
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  float hA[128 * 32], hB[NN * 32], hD[128 * NN];
  for (int i = 0; i < 128 * 32; ++i) hA[i] = float((i * 7 + (i / 32) * 3) % 13 - 6) * 0.25f;
  for (int i = 0; i < NN * 32; ++i) hB[i] = float((i * 5 + (i / 32)) % 9 - 4) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < NN; ++n) {
    double ref = 0; for (int k = 0; k < 16 * KS; ++k) ref += double(hA[r * 32 + k]) * hB[n * 32 + k];
    double err = fabs(ref - hD[r * NN + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
  }
  printf("mixed f16 x bf16 TS-MMA probe: max err %.3e, bad %d of %d  (D[0][0..3] = %g %g %g %g)\n", maxerr, bad, 128 * NN, hD[0], hD[1], hD[2], hD[3]);
  return bad ? 1 : 0;
}
