// History side of the fused tensor-core scoring path, software-pipelined across tiles (sm_100a: tcgen05 + TMEM + TMA).
// Same contract as hist_kernel.cu (PolyAttention.forward, reference src/model/model.py:159-185, straight from the table),
// used when Dc <= 208.  Differences:
//   * the logits against the context codes run on the tensor cores too: the epilogue warps write tanh(proj) back to
//     tensor memory as packed bf16 hi + lo (tcgen05.st) and the MMA warp multiplies it, as a TMEM-resident A operand (TS
//     form), with the context codes (bf16 hi + lo K-major tiles in shared memory) into a 128 x 32 accumulator
//     (T_hi C_hi + T_lo C_hi + T_hi C_lo: ~2^-17 relative, i.e. fp32-level logits);
//   * as soon as the epilogue warps have turned the projection of tile t into T, the projection accumulator is free: the
//     MMA warp accumulates the projection of tile t+1 while tile t goes through softmax, weighted sum and draining, so
//     the gather / TMA pipeline keeps running across the phase boundaries.  The issue order of one step is fixed and
//     shared by the three roles:
//         LG(t)  P1(t+1)[0..7]  { P2(t)[j], P1(t+1)[8 + j] }  rest of P1(t+1)
//     P1 = one projection k-block (gathered E tile + Wp tile), LG = logits MMAs, P2 = one 64-feature block of the interests
//     (A = softmax weights hi|lo from shared memory, B = the re-gathered E tile read MN-major);
//   * TMEM map (512 columns): projection [0,208) | T_hi [208,312) | T_lo [312,416) | logits [416,448); once the logits MMAs
//     are done the T region is dead and holds three 64-column interest buffers [208,400), drained round-robin.
#include <cuda.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {

long long* hist_prof_buffer();

namespace {

constexpr int HM = 128, HKB = 64;
constexpr int NA = 3;                        // ring of gathered E k-blocks of the projection pipeline
constexpr int NA2 = 2;                       // ring of gathered E k-blocks of the weighted-sum pipeline
constexpr int NB = 2;                        // ring of Wp k-blocks (projection pass only; L2-hot)
constexpr int HA_BYTES = HM * HKB * 2;       // 16 KB gathered E k-block
constexpr int KP = 32;
constexpr int LROW = 33;
constexpr int H_THREADS = 15 * 32;           // 4 gather warps, TMA warp, 2 MMA warps (5 and 14), 8 epilogue warps (6..13)
constexpr int H_EPI = 256;
constexpr int N1_MAX = 208;                  // projection accumulator columns (Dc padded to 16)
// TMEM map (512 columns): projection [0,208) | T = tanh(proj) as packed fp16 [208,312) | logits [312,344) | two 64-column
// interest buffers [344,408), [408,472)
constexpr int TH_COL = N1_MAX, LG_COL = TH_COL + N1_MAX / 2;
constexpr int IA_COL = LG_COL + KP, IA_BUFS = 2;
constexpr int WA_BYTES = 64 * 128;           // softmax-weight atom (64 interest rows x 64 history slots); the MMA reads 8 KB past it
constexpr int WT_BYTES = 4 * WA_BYTES;
constexpr int CT_ATOM = KP * 128;            // codes tile atom: 32 codes x 64 features
constexpr int CT_BYTES = 4 * CT_ATOM;        // per hi / lo tile (Dc <= 256)

enum { OP_P1 = 0, OP_LG = 1, OP_P2 = 2 };

// Optional cycle accounting (build with -DMINER_HIST_PROF): per CTA, 16 counters each for the MMA thread, one epilogue
// thread and one gather thread, written to args.prof at the end (scripts/prof_hist.py prints them).
#ifdef MINER_HIST_PROF
#define PROF_DECL long long prof_c[16] = {0}; long long prof_t0 = clock64(), prof_start = prof_t0
#define PROF_ADD(i) do { const long long prof_t1 = clock64(); prof_c[i] += prof_t1 - prof_t0; prof_t0 = prof_t1; } while (0)
#define PROF_STORE(role) do { if (args.prof) { prof_c[15] = clock64() - prof_start; for (int i_ = 0; i_ < 16; ++i_) args.prof[(blockIdx.x * 4 + (role)) * 16 + i_] = prof_c[i_]; } } while (0)
#else
#define PROF_DECL
#define PROF_ADD(i)
#define PROF_STORE(role)
#endif

struct H2Barriers {
  uint64_t full_a[NA], empty_a[NA], full_a2[NA2], empty_a2[NA2], full_b[NB], empty_b[NB];
  uint64_t p1_full, t_ready, lg_full, w_ready;
  uint64_t ia_full[IA_BUFS], ia_free[IA_BUFS];
  uint32_t tmem_base;
};

struct Hist2Args {
  const uint16_t* table; int64_t n_rows;
  const void* his_ids; int id_dtype;
  const uint8_t* mask; const float* bias_mean;
  const float* codes;                          // (K, Dc) fp32
  int64_t B;
  int H, K, Dc, D, N1, b_bytes, first;
  __nv_bfloat16* i_hi; __nv_bfloat16* i_lo; float* out_interests;
  long long* prof;
};

// issue order of one step; `has_cur` false = prologue (only the projection of the first tile).  Every projection block of
// tile t+1 is issued before interest block `first` of tile t: from there on the epilogue warps interleave the tanh
// conversion of tile t+1 with the remaining drains of tile t.
template <class F>
__device__ __forceinline__ void for_each_op(bool has_cur, bool has_next, int KB, int first, F&& f) {
  int p1 = 0;
  auto p1n = [&](int n) {
    for (int i = 0; i < n && has_next && p1 < KB; ++i) f(OP_P1, p1++);
  };
  if (!has_cur) { p1n(KB); return; }
  f(OP_LG, 0);
  // `first` interest blocks are paired 1:1 with projection blocks (continuous ingest); what does not fit is front-loaded
  // while the epilogue warps run the softmax of the current tile
  p1n(KB > first ? KB - first : 0);
  for (int j = 0; j < KB; ++j) {
    f(OP_P2, j);
    if (j < first) p1n(1);
  }
  p1n(KB);
}

__device__ __forceinline__ float tanh_acc2(float x) {          // exp form, any x
  const float ax = fabsf(x);
  const float e = __expf(-2.0f * ax);
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
// tanh to ~1.5e-7 absolute on |x| <= 1: odd minimax polynomial, FMA pipe only (the projection pre-activations are small);
// the caller falls back to tanh_acc2 for a whole unit if any |x| exceeds 1
__device__ __forceinline__ float tanh_poly(float x, float s) {
  float p = fmaf(s, 1.045553543e-03f, -6.217069879e-03f);
  p = fmaf(p, s, 2.030551945e-02f);
  p = fmaf(p, s, -5.346517736e-02f);
  p = fmaf(p, s, 1.332539784e-01f);
  p = fmaf(p, s, -3.333285609e-01f);
  p = fmaf(p, s, 9.999999527e-01f);
  return x * p;
}
__device__ __forceinline__ uint32_t pack2b(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(H_THREADS, 1)
hist_kernel2(const __grid_constant__ CUtensorMap tmap_wp, const Hist2Args args, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* st_a = smem;                                    // [NA][16 KB]       gathered E k-block
  uint8_t* st_a2 = st_a + NA * HA_BYTES;                   // [NA2][16 KB]      gathered E k-block of the weighted-sum pipeline
  uint8_t* st_b = st_a2 + NA2 * HA_BYTES;                  // [NB][b_bytes]     Wp k-block
  uint8_t* w_t = st_b + NB * args.b_bytes;                 // 4 x 8 KB          softmax weights, bf16 {hi,lo} x {rows 0-63, 64-127}
  uint8_t* c_hi = w_t + WT_BYTES;                          // 16 KB             context codes bf16 hi, K-major SW128 atoms
  uint8_t* c_lo = c_hi + CT_BYTES;                         // 16 KB             ... lo
  float* L = reinterpret_cast<float*>(c_lo + CT_BYTES);    // [128][33]         logits / softmax scratch
  H2Barriers* bars = reinterpret_cast<H2Barriers*>(reinterpret_cast<uint8_t*>(L) + HM * LROW * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, K = args.K, D = args.D, Dc = args.Dc, N1 = args.N1;
  const int KB = D / HKB;
  const int IPT = H <= 64 ? 2 : 1;
  const int HP = HM / IPT;
  const int n_local = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n_cu = N1 / 16;                                // 16-column units of the projection
  const int first = args.first < KB ? args.first : KB - 1;   // first interest block whose drain is paired with tanh units of the next tile

  for (int i = threadIdx.x; i < (WT_BYTES + 2 * CT_BYTES) / 16; i += H_THREADS) reinterpret_cast<uint4*>(w_t)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < K * Dc; i += H_THREADS) {
    const int k = i / Dc, dc = i - k * Dc;
    const float c = args.codes[i];
    const __half hi = __float2half_rn(c);
    const uint32_t off = (dc >> 6) * CT_ATOM + tc::sw128_offset(k, (dc & 63) >> 3) + (dc & 7) * 2;
    *reinterpret_cast<__half*>(c_hi + off) = hi;
    *reinterpret_cast<__half*>(c_lo + off) = __float2half_rn(c - __half2float(hi));
  }
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NA; ++s) { tc::mbar_init(&bars->full_a[s], 64); tc::mbar_init(&bars->empty_a[s], 1); }
    for (int s = 0; s < NA2; ++s) { tc::mbar_init(&bars->full_a2[s], 64); tc::mbar_init(&bars->empty_a2[s], 1); }
    for (int s = 0; s < NB; ++s) { tc::mbar_init(&bars->full_b[s], 1); tc::mbar_init(&bars->empty_b[s], 1); }
    tc::mbar_init(&bars->p1_full, 1);
    tc::mbar_init(&bars->t_ready, H_EPI);
    tc::mbar_init(&bars->lg_full, 1);
    tc::mbar_init(&bars->w_ready, H_EPI);
    for (int b = 0; b < IA_BUFS; ++b) { tc::mbar_init(&bars->ia_full[b], 1); tc::mbar_init(&bars->ia_free[b], H_EPI); }
    tc::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&tmap_wp);
  if (warp == 5) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ E gathers: warps 0-1 feed the projection ring with the
    //        rows of tile lt + 1 ... no: of every tile in turn, warps 2-3 feed the weighted-sum ring; the two pipelines only meet
    //        through the epilogue warps' barriers.  64 threads per ring: 16 rows x one 16-byte chunk each per stage.
    const int ring = warp >> 1;                          // 0: projection (P1), 1: weighted sum (P2)
    const int t64 = (warp & 1) * 32 + lane;
    const int chunk = t64 & 7;
    uint32_t dst_off[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dst_off[j] = tc::sw128_offset((t64 >> 3) + 8 * j, chunk);
    int32_t ids_pre[16];                                 // row of the table, -1 = zero row (padding slot / out-of-range id)
    auto fetch_ids = [&](int lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = (t64 >> 3) + 8 * j;
        const int64_t imp = static_cast<int64_t>(tile) * IPT + r / HP;
        const int h = r % HP;
        const bool ok = h < H && imp < args.B;
        const int64_t id = load_id(args.his_ids, ok ? imp * H + h : 0, args.id_dtype);
        ids_pre[j] = (ok && id >= 0 && id < args.n_rows) ? static_cast<int32_t>(id) : -1;
      }
    };
    uint8_t* stage0 = ring == 0 ? st_a : st_a2;
    uint64_t* fullb = ring == 0 ? bars->full_a : bars->full_a2;
    uint64_t* emptyb = ring == 0 ? bars->empty_a : bars->empty_a2;
    const int depth = ring == 0 ? NA : NA2;
    uint32_t issued = 0;
    PROF_DECL;
    if (n_local > 0) fetch_ids(0);
    for (int lt = 0; lt < n_local; ++lt) {
      const uint16_t* src[16];
      uint32_t nb = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const bool ok = ids_pre[j] >= 0;
        src[j] = args.table + static_cast<int64_t>(ok ? ids_pre[j] : 0) * D + chunk * 8;
        nb |= ok ? (1u << j) : 0u;
      }
      if (lt + 1 < n_local) fetch_ids(lt + 1);
      PROF_ADD(0);
      for (int kb = 0; kb < KB; ++kb) {
        const uint32_t s = issued % depth, ph = (issued / depth) & 1;
        tc::mbar_wait_relaxed(&emptyb[s], ph ^ 1);
        PROF_ADD(1);
        const uint32_t base = tc::smem_u32(stage0 + s * HA_BYTES);
#pragma unroll
        for (int j = 0; j < 16; ++j) tc::cp_async_16(base + dst_off[j], src[j] + kb * HKB, ((nb >> j) & 1u) ? 16u : 0u);
        tc::cp_async_mbar_arrive_noinc(&fullb[s]);
        ++issued;
        PROF_ADD(2);
      }
    }
    tc::cp_async_wait_all();
    if (threadIdx.x == 0) PROF_STORE(2);
  } else if (warp == 4) {
    // ------------------------------------------------------------------ Wp k-blocks by TMA (projection pipeline)
    uint32_t it = 0;
    for (int lt = 0; lt < n_local; ++lt) {
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const uint32_t s = it % NB, ph = (it / NB) & 1;
        tc::mbar_wait_relaxed(&bars->empty_b[s], ph ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&bars->full_b[s], static_cast<uint32_t>(N1 * HKB * 2));
          tc::tma_load_2d(&tmap_wp, &bars->full_b[s], tc::smem_u32(st_b + s * args.b_bytes), kb * HKB, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer of the projection pipeline (P1)
    const uint32_t idesc1 = tc::make_idesc_bf16_f32(HM, N1);
    uint32_t it = 0;
    PROF_DECL;
    for (int lt = 0; lt < n_local; ++lt) {
      if (lt > 0) {
        PROF_ADD(0);
        tc::mbar_wait(&bars->t_ready, (lt - 1) & 1);       // tile lt-1 has been turned into T: the projection accumulator is free
        PROF_ADD(3);
        tc::tcgen05_fence_after();
      }
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        const uint32_t sb = it % NB, phb = (it / NB) & 1;
        PROF_ADD(0);
        tc::mbar_wait(&bars->full_a[s], ph);
        PROF_ADD(1);
        tc::mbar_wait(&bars->full_b[sb], phb);
        PROF_ADD(9);
        tc::tcgen05_fence_after();
        const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_a + s * HA_BYTES));
        const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_b + sb * args.b_bytes));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < HKB / 16; ++k) tc::umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(&bars->empty_a[s]);
          tc::umma_commit(&bars->empty_b[sb]);
          if (kb == KB - 1) tc::umma_commit(&bars->p1_full);
        }
        __syncwarp();
        PROF_ADD(2);
      }
    }
    if (lane == 0) PROF_STORE(0);
  } else if (warp == 14) {
    // ------------------------------------------------------------------ MMA issuer of the logits (LG) and weighted-sum (P2) pipeline
    const uint32_t idesc_lg = tc::make_idesc_f16_f32(HM, KP);         // T and the context codes are fp16 (11-bit mantissa)
    const uint32_t idesc2 = tc::make_idesc_bf16_f32_major(HM, HKB, false, true);     // B = E k-block read MN-major
    uint32_t it = 0, gj = 0;
    PROF_DECL;
    for (int lt = 0; lt < n_local; ++lt) {
      PROF_ADD(0);
      tc::mbar_wait(&bars->t_ready, lt & 1);               // tanh(proj) of this tile sits in T as packed fp16
      PROF_ADD(3);
      tc::tcgen05_fence_after();
      {
        const uint64_t ch_desc = tc::make_smem_desc_sw128(tc::smem_u32(c_hi));
        const uint64_t cl_desc = tc::make_smem_desc_sw128(tc::smem_u32(c_lo));
        if (tc::elect_one()) {
          for (int ks = 0; ks < N1 / 16; ++ks) {
            const uint32_t adv = (ks >> 2) * (CT_ATOM >> 4) + 2 * (ks & 3);
            tc::umma_bf16_ts(tmem + LG_COL, tmem + TH_COL + 8 * ks, ch_desc + adv, idesc_lg, ks != 0 ? 1u : 0u);
            tc::umma_bf16_ts(tmem + LG_COL, tmem + TH_COL + 8 * ks, cl_desc + adv, idesc_lg, 1u);
          }
          tc::umma_commit(&bars->lg_full);
        }
        __syncwarp();
      }
      PROF_ADD(4);
      tc::mbar_wait(&bars->w_ready, lt & 1);               // softmax weights of this tile are in shared memory
      PROF_ADD(5);
      tc::tcgen05_fence_after();
      for (int j = 0; j < KB; ++j, ++it, ++gj) {
        const uint32_t slot = gj & 1;
        tc::mbar_wait(&bars->ia_free[slot], ((gj >> 1) & 1) ^ 1);       // its previous 64-feature block is out of this accumulator
        PROF_ADD(6);
        const uint32_t s = it % NA2, ph = (it / NA2) & 1;
        tc::mbar_wait(&bars->full_a2[s], ph);
        PROF_ADD(7);
        tc::tcgen05_fence_after();
        const uint64_t e_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st_a2 + s * HA_BYTES));
        const uint64_t w_desc0 = tc::make_smem_desc_sw128(tc::smem_u32(w_t));
        if (tc::elect_one()) {
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
#pragma unroll
            for (int ks = 0; ks < HM / 16; ++ks) {
              const uint64_t w_desc = w_desc0 + ((hl * 2 + (ks >> 2)) * (WA_BYTES >> 4) + 2 * (ks & 3));
              tc::umma_bf16(tmem + IA_COL + slot * HKB, w_desc, e_desc + ks * (2048 >> 4), idesc2, (hl | ks) != 0 ? 1u : 0u);
            }
          }
          tc::umma_commit(&bars->empty_a2[s]);
          tc::umma_commit(&bars->ia_full[slot]);
        }
        __syncwarp();
        PROF_ADD(8);
      }
    }
    if (lane == 0) PROF_STORE(3);
  } else {
    // ------------------------------------------------------------------ epilogue warps 6..13
    const int ew = warp - 6;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int r = q * 32 + lane;                           // tile row = TMEM lane
    const int et = ew * 32 + lane;                         // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int cu_begin = half == 0 ? 0 : (n_cu + 1) / 2, cu_end = half == 0 ? (n_cu + 1) / 2 : n_cu;
    const int n_u = cu_end - cu_begin;
    const int upd = (n_u + (KB - first) - 1) / (KB - first);      // tanh units per paired drain
    uint32_t gj = 0;
    PROF_DECL;
    // one 16-column unit of E1a: tanh(proj) -> packed bf16 hi / lo in tensor memory (model.py:171)
    auto e1a_unit = [&](int cu) {
      uint32_t v[16];
      tc::tmem_ld_32x16(tmem + lane_addr + cu * 16, v);
      tc::tmem_ld_wait();
      uint32_t pk[8];
      float t[16], smax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = __uint_as_float(v[j]), sq = x * x;
        smax = fmaxf(smax, sq);
        t[j] = tanh_poly(x, sq);
      }
      if (smax > 1.0f) {                                                       // rare: large pre-activations
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = tanh_acc2(__uint_as_float(v[j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half2 h = __floats2half2_rn(t[2 * j], t[2 * j + 1]);           // |tanh| <= 1: fp16 keeps 11 bits
        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tc::tmem_st_32x8(tmem + lane_addr + TH_COL + cu * 8, pk);
    };
    auto e1a_done = [&]() {
      tc::tmem_st_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars->t_ready);
    };
    // prologue: the first tile is converted in one go
    if (n_local > 0) {
      tc::mbar_wait(&bars->p1_full, 0);
      PROF_ADD(1);
      tc::tcgen05_fence_after();
      for (int cu = cu_begin; cu < cu_end; ++cu) e1a_unit(cu);
      e1a_done();
      PROF_ADD(2);
    }
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const bool has_next = lt + 1 < n_local;
      // ---- E1b: logits (+bias), 1e-30 mask fill, softmax over the history (model.py:174-181)
      // mask / bias of this row are fetched before the wait so that their latency hides behind the logits MMAs
      const int64_t imp_r = static_cast<int64_t>(tile) * IPT + r / HP;
      const int h_r = r % HP;
      const bool valid = h_r < H && imp_r < args.B;
      const bool keep = valid && args.mask[imp_r * H + h_r] != 0;
      const float bias = (valid && args.bias_mean) ? args.bias_mean[imp_r * H + h_r] : 0.f;
      tc::mbar_wait(&bars->lg_full, lt & 1);
      PROF_ADD(3);
      tc::tcgen05_fence_after();
      {
        uint32_t lg[16];
        tc::tmem_ld_32x16(tmem + lane_addr + LG_COL + half * 16, lg);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = __uint_as_float(lg[j]) + bias;
          if (!keep) v = kMaskFill;                                            // model.py:180 (1e-30, not -inf)
          if (!valid) v = -INFINITY;                                           // tile padding: not part of the history
          L[r * LROW + half * 16 + j] = v;
        }
      }
      tc::tcgen05_fence_before();
      tc::named_bar_sync(1, H_EPI);
      {
        const int pair = et >> 2, part = et & 3;
        const bool active = pair < IPT * K;
        const int i = active ? pair / K : 0, k = active ? pair % K : 0;
        const float* col = L + (i * HP) * LROW + k;
        float mx = -INFINITY;
        for (int h = part; h < HP; h += 4) mx = fmaxf(mx, col[h * LROW]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const bool dead = mx == -INFINITY;                                     // impression past the end of the batch
        float ev[HM / 4];                                                      // this thread's exp values (HP / 4 <= 32 slots)
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < HM / 4; ++t) {
          const int h = part + 4 * t;
          ev[t] = (h < HP && !dead) ? expf(col[h * LROW] - mx) : 0.f;
          sum += ev[t];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (active) {
          const int R = i * K + k;
#pragma unroll
          for (int t = 0; t < HM / 4; ++t) {
            const int h = part + 4 * t;
            if (h < HP) {
              const float w = dead ? 0.f : ev[t] / sum;                        // model.py:181
              const __nv_bfloat16 whi = __float2bfloat16_rn(w);
              const __nv_bfloat16 wlo = __float2bfloat16_rn(w - __bfloat162float(whi));
              const int hc = i * HP + h;
              const uint32_t off = (hc >> 6) * WA_BYTES + tc::sw128_offset(R, (hc & 63) >> 3) + (hc & 7) * 2;
              *reinterpret_cast<__nv_bfloat16*>(w_t + off) = whi;
              *reinterpret_cast<__nv_bfloat16*>(w_t + 2 * WA_BYTES + off) = wlo;
            }
          }
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bars->w_ready);
      PROF_ADD(4);
      // ---- E2: drain the interests one 64-feature block at a time (model.py:182); from block `first` on, each drain is
      //      followed by one tanh unit of the NEXT tile, whose projection is complete by then
      const bool row_ok = r < IPT * K;
      const int64_t imp2 = static_cast<int64_t>(tile) * IPT + r / K;
      const bool store_ok = row_ok && imp2 < args.B;
      const int64_t grow = static_cast<int64_t>(tile) * IPT * K + r;            // = imp * K + k
      int next_cu = cu_begin;
      for (int j = 0; j < KB; ++j, ++gj) {
        if (has_next && j == first) {
          tc::mbar_wait(&bars->p1_full, (lt + 1) & 1);
          PROF_ADD(1);
          tc::tcgen05_fence_after();
        }
        const uint32_t slot = gj & 1;
        tc::mbar_wait(&bars->ia_full[slot], (gj >> 1) & 1);
        PROF_ADD(5);
        tc::tcgen05_fence_after();
        if (q * 32 < IPT * K) {                                                // warp-uniform: this lane quarter holds interest rows
          uint32_t v[32];
          tc::tmem_ld_32x32(tmem + lane_addr + IA_COL + slot * HKB + half * 32, v);
          tc::tmem_ld_wait();
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->ia_free[slot]);                               // the accumulator is in registers: release it first
          if (store_ok) {
            const int64_t o = grow * D + j * HKB + half * 32;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float x0 = __uint_as_float(v[2 * c]), x1 = __uint_as_float(v[2 * c + 1]);
              const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
              hi[c] = pack2b(x0, x1);
              lo[c] = pack2b(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
            }
            uint4* ph = reinterpret_cast<uint4*>(args.i_hi + o);
            uint4* pl = reinterpret_cast<uint4*>(args.i_lo + o);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              ph[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
              pl[c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            if (args.out_interests) {
              float4* pf = reinterpret_cast<float4*>(args.out_interests + o);
#pragma unroll
              for (int c = 0; c < 8; ++c)
                pf[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                    __uint_as_float(v[4 * c + 3]));
            }
          }
        } else {
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->ia_free[slot]);
        }
        PROF_ADD(6);
        if (has_next && j >= first) {
          for (int u = 0; u < upd && next_cu < cu_end; ++u) e1a_unit(next_cu++);
          PROF_ADD(2);
        }
      }
      if (has_next) {
        if (first >= KB) {                                                     // KB == 0 cannot happen; defensive: projection not yet awaited
          tc::mbar_wait(&bars->p1_full, (lt + 1) & 1);
          tc::tcgen05_fence_after();
        }
        for (; next_cu < cu_end; ++next_cu) e1a_unit(next_cu);
        e1a_done();
        PROF_ADD(2);
      }
    }
    if (et == 0) PROF_STORE(1);
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// buffer for the cycle counters of -DMINER_HIST_PROF / -DMINER_TS_PROF builds (148 CTAs x 5 roles x 16 counters), set through
// miner_debug_set_hist_prof -- an entry point only those builds export
static long long* g_hist_prof = nullptr;
long long* hist_prof_buffer() { return g_hist_prof; }
void set_hist_prof_buffer(long long* p) { g_hist_prof = p; }

bool hist_kernel2_supported(int64_t H, int64_t K, int64_t Dc, int64_t D) {
  return H >= 1 && H <= 128 && (K == 8 || K == 16 || K == 32) && Dc >= 1 && Dc <= N1_MAX && D >= 64 && D % 64 == 0 && D <= 4096;
}

int launch_hist_kernel2(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                        const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H, int64_t K,
                        int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests, cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  const int N1 = static_cast<int>((Dc + 15) / 16 * 16);
  CUtensorMap m_wp;
  int rc = make_tmap_2d_bf16(&m_wp, w_proj_bf16, static_cast<uint64_t>(Dc), static_cast<uint64_t>(D), N1, HKB);
  if (rc) return rc;
  Hist2Args a;
  a.table = static_cast<const uint16_t*>(table); a.n_rows = n_rows;
  a.his_ids = his_ids; a.id_dtype = id_dtype; a.mask = his_mask; a.bias_mean = bias_mean; a.codes = codes;
  a.B = B; a.H = static_cast<int>(H); a.K = static_cast<int>(K); a.Dc = static_cast<int>(Dc); a.D = static_cast<int>(D); a.N1 = N1;
  a.b_bytes = N1 * HKB * 2;
  {
    const int kb = static_cast<int>(D / HKB);
    a.first = (2 * kb) / 3;                                             // 2/3 of the blocks paired (measured best)
#ifdef MINER_HIST_PROF
    static const char* env_first = getenv("MINER_HIST_FIRST");          // tuning knob of the instrumented build only
    if (env_first) a.first = atoi(env_first);
#endif
    if (a.first < 0) a.first = 0;
    if (a.first > kb - 1) a.first = kb - 1;
  }
  a.i_hi = static_cast<__nv_bfloat16*>(i_hi); a.i_lo = static_cast<__nv_bfloat16*>(i_lo); a.out_interests = out_interests;
  a.prof = hist_prof_buffer();
  const int ipt = H <= 64 ? 2 : 1;
  const int64_t n_tiles = (B + ipt - 1) / ipt;
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  const int smem = 1024 + (NA + NA2) * HA_BYTES + NB * a.b_bytes + WT_BYTES + 2 * CT_BYTES + HM * LROW * 4 + 512;
  MINER_CUDA_OK(cudaFuncSetAttribute(hist_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  hist_kernel2<<<grid, H_THREADS, smem, stream>>>(m_wp, a, static_cast<int>(n_tiles));
  MINER_LAUNCH_OK("hist_kernel2");
  return MINER_OK;
}

// public names of the history kernel (miner_hist_interests_fwd, miner_score_fwd's tensor family)
bool hist_kernel_supported(int64_t H, int64_t K, int64_t Dc, int64_t D) { return hist_kernel2_supported(H, K, Dc, D); }
size_t hist_kernel_ws_bytes(int64_t Dc) { (void)Dc; return 256; }      // the kernel stages the context codes itself; kept for the ABI
int launch_hist_kernel(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                       const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H, int64_t K,
                       int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests, float* codes_t_ws, cudaStream_t stream) {
  (void)codes_t_ws;
  if (B == 0) return MINER_OK;
  if (!hist_kernel2_supported(H, K, Dc, D)) {
    set_error("hist_kernel: unsupported shape H=%lld K=%lld Dc=%lld D=%lld", (long long)H, (long long)K, (long long)Dc, (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  return launch_hist_kernel2(table, n_rows, his_ids, id_dtype, his_mask, bias_mean, w_proj_bf16, codes, B, H, K, Dc, D, i_hi, i_lo,
                             out_interests, stream);
}


}  // namespace miner
