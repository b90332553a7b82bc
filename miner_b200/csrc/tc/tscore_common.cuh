// Shared pieces of the table-level scoring kernels (tscore_kernel.cu: accumulates the interests on the tensor cores;
// tscore_x_kernel.cu: the matching scores through X = E cand^T): packed-tile records, kernel arguments, id / range helpers,
// packed fp32x2 arithmetic.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {
namespace ts {

constexpr int TM = 128;                      // TMEM lanes; history slots per 128-slot half of a tile
constexpr int FB = 64;                       // feature block (128 bytes of bf16)

// ---- packed tiles (written by tpack_kernel, read by tscore_kernel) ------------------------------------------------------------
// record of one slot:  x = news id (30 bits) | masked << 30 | id-in-range << 31
//                      y = multiplicity (16 bits; 0 = padding, not part of any history) | original slot h << 16 | impression-in-tile << 24
// header of a tile:    end slot (exclusive) of impression i as 16-bit fields, x = end0 | end1 << 16, y = end2 | end3 << 16;
//                      impression i starts at the even slot following end(i-1)
constexpr uint32_t REC_MASKED = 1u << 30, REC_VALID = 1u << 31, REC_ID = 0x3fffffffu;
__host__ __device__ __forceinline__ int hdr_end(uint2 h, int i) { return static_cast<int>(((i < 2 ? h.x : h.y) >> (16 * (i & 1))) & 0xffffu); }
__host__ __device__ __forceinline__ int hdr_start(uint2 h, int i) { return i == 0 ? 0 : (hdr_end(h, i - 1) + 1) & ~1; }

struct TScoreArgs {
  const uint16_t* table; const uint16_t* tw; const float* lg; int64_t n_rows;
  const uint2* rec; const uint2* hdr; int* oob;
  const void* cand_ids; int id_dtype;
  const float* bias_mean; const int64_t* cand_offsets;
  int64_t B;
  int H, K, D, C, score_type;
  float* out_scores; float* out_interests;
  long long* prof;
  int dbg;   // -DMINER_TS_PROF builds only (MINER_TS_DBG): bit 0 no global reads in the gathers (zero fill), bit 1 no tw reads, bit 2 no candidate reads
};

__device__ __forceinline__ int64_t cand_off(const TScoreArgs& a, int64_t i) { return a.cand_offsets ? a.cand_offsets[i] : i * a.C; }

// An id fetched ahead of time stays RAW (the loaded bits, nothing computed from them) until the tile that uses it: any
// instruction consuming the loaded register -- a range check, a sign extension -- would wait for the load where it was issued
// and put a DRAM round trip on the gather warps' path at every tile boundary.
struct RawId { uint32_t lo, hi; };
__device__ __forceinline__ RawId load_id_raw(const void* ids, int64_t i, int id_dtype) {
  RawId r;
  if (id_dtype == MINER_I64) {
    const uint2 v = reinterpret_cast<const uint2*>(ids)[i];
    r.lo = v.x; r.hi = v.y;
  } else {
    r.lo = reinterpret_cast<const uint32_t*>(ids)[i]; r.hi = 0;
  }
  return r;
}
__device__ __forceinline__ int64_t id_of(RawId r, int id_dtype) {
  return id_dtype == MINER_I64 ? static_cast<int64_t>((static_cast<uint64_t>(r.hi) << 32) | r.lo) : static_cast<int64_t>(static_cast<int32_t>(r.lo));
}

// candidate range of a tile (two loads, nothing else: callers issue them a tile ahead and only look at the values a tile later,
// so their latency never sits on a role's critical path) and its number of passes
template <int IPT>
__device__ __forceinline__ void tile_range(const TScoreArgs& a, int tile, int64_t& cs, int64_t& ce) {
  const int64_t i0 = static_cast<int64_t>(tile) * IPT;
  const int64_t i1 = i0 + IPT < a.B ? i0 + IPT : a.B;
  cs = cand_off(a, i0);
  ce = cand_off(a, i1);
}
template <int NCM>
__device__ __forceinline__ int passes_of(int64_t cs, int64_t ce) {
  const int64_t n = ce - cs;
  return n <= NCM ? 1 : static_cast<int>((n + NCM - 1) / NCM);
}

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): the epilogue is issue-bound, two lanes per instruction
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// (s0, s1) -> packed bf16 hi pair and packed bf16 lo pair with s = hi + lo to ~2^-17
__device__ __forceinline__ void split_hi_lo(unsigned long long s, uint32_t& hi, uint32_t& lo) {
  float s0, s1;
  f2_unpack(s, s0, s1);
  hi = pack2(s0, s1);
  const unsigned long long hf = f2_pack(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  float l0, l1;
  f2_unpack(f2_sub(s, hf), l0, l1);
  lo = pack2(l0, l1);
}
// gelu of two values, tanh form with the hardware tanh (model.py:212).  The erf and tanh forms differ by <= 5e-4 absolute; G only
// feeds the softmax-over-K logits, where that is far below the bf16 rounding of tw (scripts/numerics_table_mode.py).
__device__ __forceinline__ unsigned long long gelu2(unsigned long long x) {
  const unsigned long long c1 = f2_pack(0.0356774081f, 0.0356774081f), c0 = f2_pack(0.7978845608f, 0.7978845608f), half = f2_pack(0.5f, 0.5f);
  const unsigned long long u = f2_mul(x, f2_fma(f2_mul(x, x), c1, c0));
  const unsigned long long hx = f2_mul(x, half);
  float u0, u1;
  f2_unpack(u, u0, u1);
  return f2_fma(hx, f2_pack(tc::tanh_approx(u0), tc::tanh_approx(u1)), hx);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Optional cycle accounting (build with -DMINER_TS_PROF): per CTA, 16 counters for one thread of each role (0 MMA issuer,
// 1 epilogue (first 16-lane group), 2 gather, 3 softmax, 4 epilogue (second group)), written to args.prof at the end
// (scripts/prof_tscore.py prints them).  The same build reads the gather ablation bits from MINER_TS_DBG; the release library has neither.
#ifdef MINER_TS_PROF
#define PROF_DECL long long prof_c[16] = {0}; long long prof_t0 = clock64(), prof_start = prof_t0
#define PROF_ADD(i) do { const long long prof_t1 = clock64(); prof_c[i] += prof_t1 - prof_t0; prof_t0 = prof_t1; } while (0)
#define PROF_STORE(role) do { if (args.prof) { prof_c[15] = clock64() - prof_start; for (int i_ = 0; i_ < 16; ++i_) args.prof[(blockIdx.x * 5 + (role)) * 16 + i_] = prof_c[i_]; } } while (0)
#define TS_DBG(bit) (args.dbg & (bit))
#else
#define PROF_DECL
#define PROF_ADD(i)
#define PROF_STORE(role)
#define TS_DBG(bit) 0
#endif


}  // namespace ts

// tscore_x_kernel.cu: the X-formulation kernel for two impressions per tile, K <= 32, H <= 56, scores only
bool tscore_x_supported(int ipt, int km, int nh, int64_t H, bool want_interests);
int launch_tscore_x(const ts::TScoreArgs& args, int64_t n_tiles, cudaStream_t stream);

}  // namespace miner
