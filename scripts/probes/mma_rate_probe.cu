// Probe: cycles per tcgen05.mma (M=128, K=16, bf16) for several N, SS vs TS A operand, K-major vs MN-major B, and
// 1 / 2 / 4 independent accumulators (dependent-accumulate latency vs throughput).  One CTA, one issuing warp.
#include <cstdio>
#include <cuda_runtime.h>
#include "tc/umma.cuh"
using namespace miner;

__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_cfg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // bf16 small values
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) { tc::tmem_alloc(&tbase, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tbase;
  if (warp == 0) {
    uint32_t phase = 0;
    const int Ns[4] = {32, 64, 128, 208};
    int cfg = 0;
    for (int ni = 0; ni < 4; ++ni) {
      const int N = Ns[ni];
      for (int mode = 0; mode < 3; ++mode) {          // 0: SS K-major B, 1: SS MN-major B (N<=64 only), 2: TS
        for (int nacc = 1; nacc <= 4; nacc *= 2) {
          if (mode == 1 && N > 64) { if (threadIdx.x == 0) out[cfg] = -1; ++cfg; continue; }
          if (N * nacc > 416) { if (threadIdx.x == 0) out[cfg] = -1; ++cfg; continue; }
          const uint32_t idesc = mode == 1 ? tc::make_idesc_bf16_f32_major(128, N, false, true) : tc::make_idesc_bf16_f32(128, N);
          const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(smem));
          const uint64_t b_desc = mode == 1 ? tc::make_smem_desc_sw128_mn(tc::smem_u32(smem + 32768)) : tc::make_smem_desc_sw128(tc::smem_u32(smem + 32768));
          const int iters = 256;
          __syncwarp();
          const long long t0 = clock64();
          if (tc::elect_one()) {
            for (int i = 0; i < iters; ++i) {
              const uint32_t d = tmem + 96 + (i % nacc) * N;
              const uint32_t k = i & 3;
              if (mode == 2) tc::umma_bf16_ts(d, tmem + 8 * k, b_desc + 2 * k, idesc, 1u);
              else tc::umma_bf16(d, a_desc + 2 * k, b_desc + (mode == 1 ? k * 128 : 2 * k), idesc, 1u);
            }
            tc::umma_commit(&bar);
          }
          __syncwarp();
          tc::mbar_wait(&bar, phase);
          phase ^= 1;
          const long long t1 = clock64();
          if (threadIdx.x == 0) out[cfg] = (t1 - t0) * 100 / iters;      // cycles x100 per MMA
          ++cfg;
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; long long h[64];
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(d, 36);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int Ns[4] = {32, 64, 128, 208};
  const char* modes[3] = {"SS K-major B ", "SS MN-major B", "TS (A in TMEM)"};
  int cfg = 0;
  printf("cycles per tcgen05.mma (M=128,K=16), accumulating into 1/2/4 independent accumulators\n");
  for (int ni = 0; ni < 4; ++ni) for (int m = 0; m < 3; ++m) {
    printf("N=%3d %s :", Ns[ni], modes[m]);
    for (int a = 0; a < 3; ++a, ++cfg) { if (h[cfg] < 0) printf("     n/a"); else printf(" %7.1f", h[cfg] / 100.0); }
    printf("\n");
  }
  return 0;
}
