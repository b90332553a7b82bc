// Probe: cp.async.bulk.tensor.2d ... tile::gather4 (TMA row gather by index) on sm_100a: which tensor-map box it wants,
// where the 4 rows land in shared memory under SWIZZLE_128B, and what an out-of-range row index gives.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc/umma.cuh"
using namespace miner;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, const int* rows, int n4, int col0, uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xDEADDEADu;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar, n4 * 4 * 128);
    for (int g = 0; g < n4; ++g) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
              tc::smem_u32(smem + g * 512)),
          "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(tc::smem_u32(&bar)), "r"(col0), "r"(rows[4 * g]), "r"(rows[4 * g + 1]), "r"(rows[4 * g + 2]),
          "r"(rows[4 * g + 3])
          : "memory");
    }
  }
  tc::mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < n4 * 4 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int NR = 1000, NCOL = 256, n4 = 4;
  static uint16_t h[NR * NCOL];
  for (int r = 0; r < NR; ++r) for (int c = 0; c < NCOL; ++c) h[r * NCOL + c] = uint16_t((r * 37 + c) & 0xffff);
  uint16_t* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  int hrows[16] = {5, 900, 17, 3, 999, 0, 512, 64, 1000 /* out of range */, 7, 7, 250, 1, 2, 3, 4};
  int* drows; cudaMalloc(&drows, sizeof(hrows)); cudaMemcpy(drows, hrows, sizeof(hrows), cudaMemcpyHostToDevice);
  uint16_t* dout; cudaMalloc(&dout, 16 * 64 * 2); cudaMemset(dout, 0, 16 * 64 * 2);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {NCOL, NR}; const cuuint64_t gstr[1] = {NCOL * 2}; const cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; const cuuint32_t es[2] = {1, 1};
  CUresult cr = ((Fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d -> %d\n", box_rows, (int)cr);
  if (cr) return 3;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024);
  const int col0 = 64;
  probe<<<1, 128, 20 * 1024>>>(tm, drows, n4, col0, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  static uint16_t ho[16 * 64]; cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
  // expected: smem row i (128 B) holds table[rows[i]][col0 .. col0+63] with 16-byte chunk c stored at chunk c ^ (i % 8)
  int bad = 0;
  for (int i = 0; i < 16; ++i) for (int c = 0; c < 64; ++c) {
    const int chunk = c >> 3, pos = ((chunk ^ (i & 7)) << 3) + (c & 7);
    const uint16_t want = hrows[i] < NR ? h[hrows[i] * NCOL + col0 + c] : 0;
    if (ho[i * 64 + pos] != want) { if (bad < 6) printf("  row %d col %d: got %04x want %04x\n", i, c, ho[i * 64 + pos], want); ++bad; }
  }
  printf("gather4 probe (box_rows=%d): bad %d of %d\n", box_rows, bad, 16 * 64);
  return bad ? 1 : 0;
}
