// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b.x8 and tcgen05.st.16x128b.x8 (warp w of the CTA owns TMEM
// lanes 32w..32w+31; the 16-lane shapes address one half of them through the lane field of the address).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I miner_b200/csrc -o scripts/probes/tmem_shapes_probe scripts/probes/tmem_shapes_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc/umma.cuh"
using namespace miner;

__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_16x128b_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// out_ld[warp][half][thread][32]: what ld.16x256b.x8 returned (TMEM holds lane*1000 + col in columns 0..63)
// out_st[lane][32]: columns 64..95 read back with 32x32b after st.16x128b.x8 wrote code = half*100000 + thread*100 + reg
__global__ void probe(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tc::tmem_alloc(&tbase, 128); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tbase;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = (warp * 32 + lane) * 1000 + c0 + j;
    tc::tmem_st_32x16(tmem + lane_addr + c0, v);
  }
  tc::tmem_st_wait();
  __syncwarp();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    ld_16x256b_x8(tmem + lane_addr + (static_cast<uint32_t>(half * 16) << 16), r);
    tc::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out_ld[((warp * 2 + half) * 32 + lane) * 32 + j] = r[j];
    uint32_t w[16];
    for (int j = 0; j < 16; ++j) w[j] = half * 100000 + lane * 100 + j;
    st_16x128b_x8(tmem + lane_addr + (static_cast<uint32_t>(half * 16) << 16) + 64, w);
  }
  tc::tmem_st_wait();
  __syncwarp();
  uint32_t v[32];
  tc::tmem_ld_32x32(tmem + lane_addr + 64, v);
  tc::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out_st[(warp * 32 + lane) * 32 + j] = v[j];
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

int main() {
  uint32_t *d_ld, *d_st;
  static uint32_t h_ld[4 * 2 * 32 * 32], h_st[128 * 32];
  cudaMalloc(&d_ld, sizeof(h_ld)); cudaMalloc(&d_st, sizeof(h_st));
  probe<<<1, 128>>>(d_ld, d_st);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(h_ld, d_ld, sizeof(h_ld), cudaMemcpyDeviceToHost); cudaMemcpy(h_st, d_st, sizeof(h_st), cudaMemcpyDeviceToHost);
  // predicted: ld reg 4n+{0,1} = lane base + t/4, col 8n + 2(t%4) + {0,1}; reg 4n+{2,3} = lane + 8, same columns
  int bad_ld = 0, bad_st = 0;
  for (int w = 0; w < 4; ++w) for (int hf = 0; hf < 2; ++hf) for (int t = 0; t < 32; ++t) for (int j = 0; j < 32; ++j) {
    const int n = j / 4, q = j % 4;
    const int ln = w * 32 + hf * 16 + t / 4 + (q >= 2 ? 8 : 0), col = 8 * n + 2 * (t % 4) + (q & 1);
    if (h_ld[((w * 2 + hf) * 32 + t) * 32 + j] != uint32_t(ln * 1000 + col)) ++bad_ld;
  }
  // predicted: st reg 2n -> lane base + t/4, col 4n + t%4; reg 2n+1 -> lane + 8, same column
  for (int w = 0; w < 4; ++w) for (int hf = 0; hf < 2; ++hf) for (int t = 0; t < 32; ++t) for (int j = 0; j < 16; ++j) {
    const int n = j / 2;
    const int ln = w * 32 + hf * 16 + t / 4 + ((j & 1) ? 8 : 0), col = 4 * n + t % 4;
    if (h_st[ln * 32 + col] != uint32_t(hf * 100000 + t * 100 + j)) ++bad_st;
  }
  printf("ld.16x256b.x8 mapping mismatches: %d   st.16x128b.x8 mapping mismatches: %d\n", bad_ld, bad_st);
  if (bad_ld) { printf("ld thread 0 regs:"); for (int j = 0; j < 32; ++j) printf(" %u", h_ld[j]); printf("\nld thread 5 regs:"); for (int j = 0; j < 32; ++j) printf(" %u", h_ld[5 * 32 + j]); printf("\n"); }
  if (bad_st) { for (int ln = 0; ln < 17; ln += 8) { printf("st lane %d cols:", ln); for (int j = 0; j < 32; ++j) printf(" %u", h_st[ln * 32 + j]); printf("\n"); } }
  return (bad_ld || bad_st) ? 1 : 0;
}
