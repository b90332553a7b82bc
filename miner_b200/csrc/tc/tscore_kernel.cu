// Table-level scoring kernel (sm_100a: tcgen05 + TMEM, cp.async gathers): Miner.forward (reference src/model/model.py:61-138)
// for a block of impressions in ONE pass over the gathered rows, after the two projections of the path have been hoisted
// from the gathered rows to the news table (table_project.cu):
//     lg[n,k]  = tanh(table[n] Wp^T) . codes[k]          PolyAttention.linear + context codes    (model.py:171,174)
//     tw[n,:]  = table[n] Wt^T                           TargetAwareAttention.linear             (model.py:212, before the gelu)
// Both are linear in the gathered row (the tanh acts per history row), so for an impression
//     logits[h,k] = lg[his[h],k] (+bias[h]); masked slots := 1e-30; w = softmax_h                (model.py:174-181)
//     I[k,:] = sum_h w[k,h] table[his[h],:]                                                       (model.py:182)
//     P[k,:] = I[k,:] Wt^T = sum_h w[k,h] tw[his[h],:]   ;  G = gelu(P)                           (model.py:212)
//     m[c,k] = cand[c].I[k] ; a[c,k] = cand[c].G[k] ; score[c] = sum_k softmax_k(a[c,:]) m[c,k]   (model.py:127,213-214)
// which turns the 2HDDc + 2KD^2 FLOP of the reference order into two more gathered rows per history slot: the kernel is
// bound by the gather (HBM / L2 ingest), not by the tensor pipe.
//
// One persistent CTA per SM.  A tile is 128 history slots: IPT = 2 impressions of 64 slots (H <= 64, K <= 32), or IPT = 1
// impression (H <= 128, or K <= 64 which needs all 128 lanes), or, for 128 < H <= 256, one impression over NH = 2 halves of 128
// slots whose products accumulate into the same D_I | D_P pair -- the same code, template parameters <IPT, KM, NH>.  TMEM lanes are
// (impression i, context code k, part hl): lane 64 i + 16 (k / 8) + 8 hl + k % 8, where hl selects the bf16 hi / lo part of the
// softmax weight -- the two lanes of a pair accumulate  w_hi . E  and  w_lo . E  and are summed in the epilogue (fp32-level
// weights).  Keeping a pair 8 lanes apart lets the 16-lane tcgen05.ld / st shapes (16x256b / 16x128b) hand both rows of a pair
// to ONE thread: the pair sum, the gelu and the hi/lo split need no shuffles and no thread repeats another's work.
//   warps 0-3   gather: per 64-feature block, table[his] and tw[his] rows (128 slots x 128 B each), 16-byte cp.async straight
//               into the 128B-swizzled layout, completion through mbarriers;  warps 4-5: the tile's candidate rows, own ring
//   warp 6      issues every tcgen05.mma:
//               S1(j): D_I = A_w . E_j, D_P = A_w . TW_j   (A_w = softmax weights in tensor memory, TS form, block diagonal over
//                      the two impressions; B = the gathered 128 x 64 tile read MN-major)
//               S2(j): D_m += A_I . cand_j^T, D_a += A_G . cand_j^T  (A = bf16 hi|lo of I / gelu(P), written IN PLACE over the
//                      fp32 accumulator by the epilogue warps; B = candidate rows, K-major)
//   warps 7-14  epilogue: per block, pair sum, gelu, bf16 hi/lo split, tcgen05.st back in place (16-lane ld / st shapes)
//   warps 15-18 softmax over the history from the lg rows (L2-resident 128-byte rows), weights to tensor memory; and, per
//               finished tile, the softmax over K and the weighted sum of the matching scores through a shared-memory
//               transpose, one thread per candidate
// Tiles with more than 96 candidates (64 when NH = 2) run several passes (the history side is recomputed; rare).
#include <cuda.h>
#include <stdlib.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {

long long* hist_prof_buffer();

namespace {

constexpr int TM = 128;                      // TMEM lanes = history slots per tile
constexpr int FB = 64;                       // feature block (128 bytes of bf16)
// IPT impressions share a tile: 2 for H <= 64 (64 slots and 64 lanes each), 1 for H <= 128; template parameter of the kernel
constexpr int KMAX = 64;                     // largest number of context codes (template parameter KM = 32 or 64 picks the scratch sizes)
#ifndef MINER_TS_S1
#define MINER_TS_S1 5
#endif
#ifndef MINER_TS_S2
#define MINER_TS_S2 3
#endif
constexpr int S1_MAX = MINER_TS_S1, S2 = MINER_TS_S2;   // ring depths: (E, TW) blocks (one less when K > 32 needs the bytes) / candidate blocks
constexpr int E_BYTES = TM * FB * 2;         // 16 KB
constexpr int ST1_BYTES = 2 * E_BYTES;
constexpr int NC_MAX = 96;                   // candidate columns per pass
constexpr int C_BYTES = NC_MAX * FB * 2;     // 12 KB
template <int KM, int NH>
struct Shape {
  static constexpr int LS = KM;              // logits scratch row stride (floats)
  static constexpr int SS = KM + 1;          // score scratch row stride (floats)
  static constexpr int SCRATCH_FLOATS = (TM * NH * LS > 2 * NC_MAX * SS ? TM * NH * LS : 2 * NC_MAX * SS + 2) & ~1;
  static constexpr int S1 = S1_MAX - (KM > 32 ? 1 : 0) - (NH > 1 ? 1 : 0);      // wider scratch takes ring stages
  static constexpr int SMEM = 1024 + S1 * ST1_BYTES + S2 * C_BYTES + SCRATCH_FLOATS * 4 + 512;
};
constexpr int T_EPI = 256, T_SMX = 128;                // 8 epilogue warps, 4 softmax / score warps
constexpr int T_G1 = 128, T_G2 = 64;                   // gather threads of the (E, TW) ring / of the candidate ring
constexpr int G1_ROWS = TM * 8 / T_G1, G2_ROWS = NC_MAX * 8 / T_G2;  // rows per thread (a thread moves one 16-byte chunk per row)
constexpr int G1_STEP = T_G1 / 8, G2_STEP = T_G2 / 8;
constexpr int W_G2 = T_G1 / 32, W_MMA = W_G2 + T_G2 / 32, W_EPI0 = W_MMA + 1, W_SMX0 = W_EPI0 + T_EPI / 32;
constexpr int T_THREADS = (W_SMX0 + T_SMX / 32) * 32;
// TMEM map (512 columns)
// (NH = 128-slot halves of a tile's history: 1, or 2 for 128 < H <= 256)
constexpr int AW_COL = 0;                    // softmax weights, packed bf16: 128 NH slots -> 64 NH columns
// then IP_COL = 64 NH: 2 buffers x (I 64 | P 64) fp32, their first 32 columns become the packed A operands;
// DM_COL = IP_COL + 256: matching scores m[(i,k,hl), c];  DA_COL = DM_COL + NCM: attention logits;  NCM = 96 (NH = 1) or 64

// Optional cycle accounting (build with -DMINER_TS_PROF): per CTA, 16 counters for one thread of each role (0 MMA issuer,
// 1 epilogue (interest half), 2 gather, 3 softmax, 4 epilogue (gelu half)), written to args.prof at the end (scripts/prof_tscore.py prints them).
#ifdef MINER_TS_PROF
#define PROF_DECL long long prof_c[16] = {0}; long long prof_t0 = clock64(), prof_start = prof_t0
#define PROF_ADD(i) do { const long long prof_t1 = clock64(); prof_c[i] += prof_t1 - prof_t0; prof_t0 = prof_t1; } while (0)
#define PROF_STORE(role) do { if (args.prof) { prof_c[15] = clock64() - prof_start; for (int i_ = 0; i_ < 16; ++i_) args.prof[(blockIdx.x * 5 + (role)) * 16 + i_] = prof_c[i_]; } } while (0)
#else
#define PROF_DECL
#define PROF_ADD(i)
#define PROF_STORE(role)
#endif

struct TBarriers {
  uint64_t full1[S1_MAX], empty1[S1_MAX], full2[S2], empty2[S2];
  uint64_t w_ready, w_free, ip_full[2], a_ready[2], dma_full, dma_free;
  uint32_t tmem_base;
#ifdef MINER_TS_PROF
  long long issue_clk[S1_MAX][4];   // when lane 0 of each (E, TW) gather warp finished issuing a stage (latency accounting)
#endif
};

struct TScoreArgs {
  const uint16_t* table; const uint16_t* tw; const float* lg; int64_t n_rows;
  const void* his_ids; const void* cand_ids; int id_dtype;
  const uint8_t* mask; const float* bias_mean; const int64_t* cand_offsets;
  int64_t B;
  int H, K, D, C, score_type;
  float* out_scores; float* out_interests;
  long long* prof;
  int dbg;   // ablations (MINER_TS_DBG): bit 0 no global reads in the gathers (zero fill), bit 1 no tw reads, bit 2 no candidate reads
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

__device__ __forceinline__ float gelu_fast(float x) {               // tanh form, hardware tanh (see cand_kernel.cu)
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  const float hx = 0.5f * x;
  return fmaf(hx, tc::tanh_approx(u), hx);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int IPT, int KM, int NH>
__global__ void __launch_bounds__(T_THREADS, 1)
tscore_kernel(const TScoreArgs args, int n_tiles) {
  constexpr int HP = TM * NH / IPT, LPI = TM / IPT;     // history slots / TMEM lanes per impression
  constexpr int LS = Shape<KM, NH>::LS, SS = Shape<KM, NH>::SS, SCRATCH_FLOATS = Shape<KM, NH>::SCRATCH_FLOATS, S1 = Shape<KM, NH>::S1;
  constexpr int NCM = NH == 1 ? NC_MAX : 64;            // candidate columns per pass
  constexpr int IP_COL = 64 * NH, DM_COL = IP_COL + 256, DA_COL = DM_COL + NCM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* st1 = smem;                                         // [S1][E 16 KB | TW 16 KB]
  uint8_t* st2 = st1 + S1 * ST1_BYTES;                         // [S2][12 KB] candidate rows
  // scratch of the softmax / score warps: the logits L of the unit being prepared and the score transposes Sm / Sa of the unit
  // being finished are never live at the same time (named barriers 1 and 2 separate the phases), so they share the bytes
  float* L = reinterpret_cast<float*>(st2 + S2 * C_BYTES);     // [128 slots][LS] logits
  float* Sm = L;                                               // [NC_MAX][SS] matching scores, transposed
  float* Sa = Sm + NC_MAX * SS;                                // [NC_MAX][SS] attention logits, transposed
  TBarriers* bars = reinterpret_cast<TBarriers*>(L + SCRATCH_FLOATS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, K = args.K, D = args.D;
  const int KB = D / FB;
  const int n_local = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { tc::mbar_init(&bars->full1[s], T_G1); tc::mbar_init(&bars->empty1[s], 1); }
    for (int s = 0; s < S2; ++s) { tc::mbar_init(&bars->full2[s], T_G2); tc::mbar_init(&bars->empty2[s], 1); }
    tc::mbar_init(&bars->w_ready, T_SMX);
    tc::mbar_init(&bars->w_free, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bars->ip_full[b], 1); tc::mbar_init(&bars->a_ready[b], T_EPI); }
    tc::mbar_init(&bars->dma_full, 1);
    tc::mbar_init(&bars->dma_free, T_SMX);
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < W_G2) {
    // ------------------------------------------------------------------ gathers of the (E, TW) ring: thread = one 16-byte chunk of rows
    //        r0 + G1_STEP jj
    const int t = threadIdx.x;
    const int chunk = t & 7, r0 = t >> 3;
    const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
    const char* table_b = reinterpret_cast<const char*>(args.table);
    const char* tw_b = reinterpret_cast<const char*>(args.tw);
    const uint32_t dst0 = tc::sw128_offset(r0, chunk);       // row r0 + G1_STEP jj sits jj * G1_STEP / 8 KB further
    RawId ids_pre[G1_ROWS * NH];                             // raw ids of the next tile (see RawId)
    auto fetch_ids = [&](int lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
#pragma unroll
      for (int jj = 0; jj < G1_ROWS * NH; ++jj) {
        const int slot = (jj / G1_ROWS) * TM + r0 + G1_STEP * (jj % G1_ROWS);      // half * 128 + row of the stage
        const int64_t imp = static_cast<int64_t>(tile) * IPT + slot / HP;
        const int h = slot % HP;
        const bool ok = h < H && imp < args.B;
        ids_pre[jj] = load_id_raw(args.his_ids, ok ? imp * H + h : 0, args.id_dtype);
      }
    };
    uint32_t g = 0;
    PROF_DECL;
    int64_t cs = 0, ce = 0;
    if (n_local > 0) { fetch_ids(0); tile_range<IPT>(args, static_cast<int>(blockIdx.x), cs, ce); }
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      uint32_t eoff[G1_ROWS * NH];                           // byte offset of this thread's 16-byte chunk in its rows (table < 4 GB, checked by the launcher)
      uint32_t emask = 0;
#pragma unroll
      for (int jj = 0; jj < G1_ROWS * NH; ++jj) {
        const int slot = (jj / G1_ROWS) * TM + r0 + G1_STEP * (jj % G1_ROWS);
        const int64_t id = id_of(ids_pre[jj], args.id_dtype);
        const bool ok = slot % HP < H && static_cast<int64_t>(tile) * IPT + slot / HP < args.B && id >= 0 && id < args.n_rows;
        eoff[jj] = static_cast<uint32_t>(ok ? id : 0) * row_bytes + chunk * 16;
        emask |= ok ? (1u << jj) : 0u;
      }
      if (args.dbg & 1) emask = 0;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) {                                // the next tile's ids and candidate range are fetched a tile ahead
        fetch_ids(lt + 1);
        tile_range<IPT>(args, static_cast<int>(blockIdx.x) + (lt + 1) * static_cast<int>(gridDim.x), cs, ce);
      }
      for (int p = 0; p < npass; ++p) {
        for (int j = 0; j < KB; ++j) {
#pragma unroll
          for (int half = 0; half < NH; ++half, ++g) {             // one ring stage per 128-slot half
            const uint32_t s = g % S1, ph = (g / S1) & 1;
            PROF_ADD(0);
            tc::mbar_wait(&bars->empty1[s], ph ^ 1);
            PROF_ADD(1);
            const uint32_t base = tc::smem_u32(st1 + s * ST1_BYTES) + dst0;
            const uint32_t jb = static_cast<uint32_t>(j) * (FB * 2);
#pragma unroll
            for (int jj = 0; jj < G1_ROWS; ++jj) {
              const uint32_t o = eoff[half * G1_ROWS + jj] + jb, sz = ((emask >> (half * G1_ROWS + jj)) & 1u) ? 16u : 0u;
              tc::cp_async_16(base + jj * (G1_STEP * 128), table_b + o, sz);
              tc::cp_async_16(base + E_BYTES + jj * (G1_STEP * 128), tw_b + o, (args.dbg & 2) ? 0u : sz);
            }
#ifdef MINER_TS_PROF
            if (lane == 0) *reinterpret_cast<volatile long long*>(&bars->issue_clk[s][warp]) = clock64();
#endif
            tc::cp_async_mbar_arrive_noinc(&bars->full1[s]);
            PROF_ADD(2);
          }
        }
      }
    }
    tc::cp_async_wait_all();
    if (threadIdx.x == 0) PROF_STORE(2);
  } else if (warp < W_MMA) {
    // ------------------------------------------------------------------ gathers of the candidate ring: thread = one 16-byte chunk of
    //        rows r0 + G2_STEP jj.  Own warps: the candidate stages are released two blocks later than the (E, TW) stages and
    //        must not hold those back.  The candidate ids of the next unit are fetched one unit ahead.
    const int t = threadIdx.x - T_G1;
    const int chunk = t & 7, r0 = t >> 3;
    const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
    const char* table_b = reinterpret_cast<const char*>(args.table);
    const uint32_t dst0 = tc::sw128_offset(r0, chunk);
    RawId cid_pre[G2_ROWS];                                  // raw candidate ids of the next unit (see RawId)
    auto fetch_cands = [&](int64_t pc0, int nc) {
#pragma unroll
      for (int jj = 0; jj < G2_ROWS; ++jj) {
        const int c = r0 + G2_STEP * jj;
        cid_pre[jj] = RawId{0u, 0u};
        if (c < nc) cid_pre[jj] = load_id_raw(args.cand_ids, pc0 + c, args.id_dtype);
      }
    };
    uint32_t g = 0;
    const int tile0 = static_cast<int>(blockIdx.x), tstep = static_cast<int>(gridDim.x);
    int64_t cs = 0, ce = 0, ncs = 0, nce = 0, n2cs = 0, n2ce = 0;      // candidate ranges of this tile, the next one, the one after
    if (n_local > 0) {
      tile_range<IPT>(args, tile0, cs, ce);
      if (n_local > 1) tile_range<IPT>(args, tile0 + tstep, ncs, nce);
      fetch_cands(cs, static_cast<int>(ce - cs < NCM ? ce - cs : NCM));
    }
    for (int lt = 0; lt < n_local; ++lt) {
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 2 < n_local) tile_range<IPT>(args, tile0 + (lt + 2) * tstep, n2cs, n2ce);
      for (int p = 0; p < npass; ++p) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        uint32_t coff[G2_ROWS];
        uint32_t cmask = 0;
#pragma unroll
        for (int jj = 0; jj < G2_ROWS; ++jj) {
          const int64_t id = id_of(cid_pre[jj], args.id_dtype);
          const bool ok = r0 + G2_STEP * jj < nc && id >= 0 && id < args.n_rows;
          coff[jj] = static_cast<uint32_t>(ok ? id : 0) * row_bytes + chunk * 16;
          cmask |= ok ? (1u << jj) : 0u;
        }
        if (args.dbg & 5) cmask = 0;
        // ids of the next unit: next pass of this tile, else first pass of the next tile (its range was loaded a tile ago)
        if (p + 1 < npass) {
          const int64_t q0 = pc0 + NCM;
          fetch_cands(q0, static_cast<int>(ce - q0 < NCM ? ce - q0 : NCM));
        } else if (lt + 1 < n_local) {
          fetch_cands(ncs, static_cast<int>(nce - ncs < NCM ? nce - ncs : NCM));
        }
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t s = g % S2, ph = (g / S2) & 1;
          tc::mbar_wait(&bars->empty2[s], ph ^ 1);
          const uint32_t base = tc::smem_u32(st2 + s * C_BYTES) + dst0;
          const uint32_t jb = static_cast<uint32_t>(j) * (FB * 2);
#pragma unroll
          for (int jj = 0; jj < G2_ROWS; ++jj)
            if (G2_STEP * jj < nc16)
              tc::cp_async_16(base + jj * (G2_STEP * 128), table_b + (coff[jj] + jb), ((cmask >> jj) & 1u) ? 16u : 0u);
          tc::cp_async_mbar_arrive_noinc(&bars->full2[s]);
        }
      }
      cs = ncs; ce = nce; ncs = n2cs; nce = n2ce;
    }
    tc::cp_async_wait_all();
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tc::make_idesc_bf16_f32_major(TM, FB, false, true);        // B = gathered tile, MN-major
    uint32_t g1 = 0, g2 = 0, u = 0, sg = 0;                      // blocks issued (S1 / S2), units, ring stages consumed
    bool pending = false;
    int pend_j = 0, pend_nc16 = 16;
    uint32_t pend_u = 0;
    PROF_DECL;
    auto stage2 = [&]() {                                                                // S2 of block g2
      const uint32_t b = g2 & 1;
      PROF_ADD(0);
      tc::mbar_wait(&bars->a_ready[b], (g2 >> 1) & 1);
      PROF_ADD(4);
      const uint32_t s = g2 % S2, ph = (g2 / S2) & 1;
      tc::mbar_wait(&bars->full2[s], ph);
      PROF_ADD(5);
      if (pend_j == 0) tc::mbar_wait(&bars->dma_free, (pend_u & 1) ^ 1);                 // previous unit's scores are out of D_m / D_a
      PROF_ADD(6);
      tc::tcgen05_fence_after();
      const uint32_t idesc2 = tc::make_idesc_bf16_f32(TM, pend_nc16);
      const uint64_t c_desc = tc::make_smem_desc_sw128(tc::smem_u32(st2 + s * C_BYTES));
      const uint32_t a_i = tmem + IP_COL + b * 128, a_g = a_i + 64;
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FB / 16; ++ks) {
          const uint32_t acc = (pend_j | ks) != 0 ? 1u : 0u;
          tc::umma_bf16_ts(tmem + DM_COL, a_i + 8 * ks, c_desc + 2 * ks, idesc2, acc);
          tc::umma_bf16_ts(tmem + DA_COL, a_g + 8 * ks, c_desc + 2 * ks, idesc2, acc);
        }
        tc::umma_commit(&bars->empty2[s]);
        if (pend_j == KB - 1) tc::umma_commit(&bars->dma_full);
      }
      __syncwarp();
      ++g2;
      PROF_ADD(7);
    };
    int64_t t_cs = 0, t_ce = 0;
    if (n_local > 0) tile_range<IPT>(args, static_cast<int>(blockIdx.x), t_cs, t_ce);
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) tile_range<IPT>(args, tile + static_cast<int>(gridDim.x), t_cs, t_ce);      // a tile ahead: off the critical path
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        PROF_ADD(0);
        tc::mbar_wait(&bars->w_ready, u & 1);
        PROF_ADD(1);
        tc::tcgen05_fence_after();
        for (int j = 0; j < KB; ++j) {
          const uint32_t b = g1 & 1;
          const uint32_t d_i = tmem + IP_COL + b * 128, d_p = d_i + 64;
#pragma unroll
          for (int half = 0; half < NH; ++half, ++sg) {
            const uint32_t s = sg % S1, ph = (sg / S1) & 1;
            PROF_ADD(0);
            tc::mbar_wait(&bars->full1[s], ph);
            PROF_ADD(2);
            tc::tcgen05_fence_after();
            const uint64_t e_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st1 + s * ST1_BYTES));
            const uint64_t t_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st1 + s * ST1_BYTES + E_BYTES));
            if (tc::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < TM / 16; ++ks) {
                const uint32_t acc = (half | ks) != 0 ? 1u : 0u;
                tc::umma_bf16_ts(d_i, tmem + AW_COL + half * 64 + 8 * ks, e_desc + ks * (2048 >> 4), idesc1, acc);
                tc::umma_bf16_ts(d_p, tmem + AW_COL + half * 64 + 8 * ks, t_desc + ks * (2048 >> 4), idesc1, acc);
              }
              tc::umma_commit(&bars->empty1[s]);
              if (half == NH - 1) {
                tc::umma_commit(&bars->ip_full[b]);
                if (j == KB - 1) tc::umma_commit(&bars->w_free);
              }
            }
            __syncwarp();
          }
          ++g1;
          PROF_ADD(3);
          if (pending) stage2();
          pending = true; pend_j = j; pend_nc16 = nc16; pend_u = u;
        }
      }
    }
    if (pending) stage2();
    if (lane == 0) PROF_STORE(0);
  } else if (warp < W_SMX0) {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - W_EPI0;
    const int q = warp & 3;                                    // TMEM lane quarter
    const int half = ew >> 2;                                  // 16-lane group of the quarter
    const int et = ew * 32 + lane;
    const int li = (q * 32) / LPI;                             // impression of this quarter's lanes
    // this warp owns lanes [16 half, 16 half + 16) of its quarter; thread t meets the (hi, lo) rows of code bk
    const uint32_t grp_addr = static_cast<uint32_t>(q * 32 + half * 16) << 16;
    const int bk = (((q * 32) % LPI) / 16 + half) * 8 + (lane >> 2);   // context code of this thread's row pair
    const int bf = 2 * (lane & 3);                              // its features inside an 8-feature group
    uint32_t g = 0;
    PROF_DECL;
    int64_t t_cs = 0, t_ce = 0;
    if (n_local > 0) tile_range<IPT>(args, static_cast<int>(blockIdx.x), t_cs, t_ce);
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) tile_range<IPT>(args, tile + static_cast<int>(gridDim.x), t_cs, t_ce);      // a tile ahead: off the critical path
      const int64_t i0 = static_cast<int64_t>(tile) * IPT;
      const int64_t my_imp = i0 + li;
      for (int p = 0; p < npass; ++p) {
        const bool want_i = args.out_interests != nullptr && p == 0 && bk < K && my_imp < args.B;
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t b = g & 1;
          PROF_ADD(0);
          tc::mbar_wait(&bars->ip_full[b], (g >> 1) & 1);
          PROF_ADD(1);
          tc::tcgen05_fence_after();
          const uint32_t acc = tmem + grp_addr + IP_COL + b * 128;
          uint32_t vi[32], vp[32];
          tc::tmem_ld_16x256b_x8(acc, vi);                                     // interests block: rows (hi, lo) x 16 of its 64 features
          tc::tmem_ld_16x256b_x8(acc + 64, vp);                                // same of P = I Wt^T
          tc::tmem_ld_wait();
          PROF_ADD(5);
          uint32_t pk[16];
          if (want_i) {                                                        // model.py:138 (interests are an output)
            float* o = args.out_interests + (my_imp * K + bk) * D + j * FB + bf;
#pragma unroll
            for (int n = 0; n < 8; ++n)
              *reinterpret_cast<float2*>(o + 8 * n) = make_float2(__uint_as_float(vi[4 * n]) + __uint_as_float(vi[4 * n + 2]),
                                                                  __uint_as_float(vi[4 * n + 1]) + __uint_as_float(vi[4 * n + 3]));
          }
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float s0 = __uint_as_float(vi[4 * n]) + __uint_as_float(vi[4 * n + 2]);       // w_hi . E + w_lo . E
            const float s1 = __uint_as_float(vi[4 * n + 1]) + __uint_as_float(vi[4 * n + 3]);
            const uint32_t hi = pack2(s0, s1);
            pk[2 * n] = hi;
            pk[2 * n + 1] = pack2(s0 - __uint_as_float(hi << 16), s1 - __uint_as_float(hi & 0xffff0000u));
          }
          tc::tmem_st_16x128b_x8(acc, pk);                                     // in place: this thread group has read all 64 columns
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            // gelu(P) only feeds the softmax-over-K logits: one bf16 (as cand_kernel.cu), the lo row of the pair stays zero
            pk[2 * n] = pack2(gelu_fast(__uint_as_float(vp[4 * n]) + __uint_as_float(vp[4 * n + 2])),           // model.py:212
                              gelu_fast(__uint_as_float(vp[4 * n + 1]) + __uint_as_float(vp[4 * n + 3])));
            pk[2 * n + 1] = 0u;
          }
          PROF_ADD(6);
          tc::tmem_st_16x128b_x8(acc + 64, pk);
          PROF_ADD(7);
          tc::tmem_st_wait();
          PROF_ADD(8);
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->a_ready[b]);
          PROF_ADD(2);
        }
      }
    }
    if (et == 0) PROF_STORE(1);
    if (et == 128) PROF_STORE(4);
  } else {
    // ------------------------------------------------------------------ softmax / score warps
    const int sw = warp - W_SMX0;
    const int q = warp & 3;
    const int li = (q * 32) / LPI;                             // impression of this quarter's lanes
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int st = sw * 32 + lane;                             // 0..127
    const int tl = q * 32 + lane;                              // TMEM lane = (i, k, hl) for the 32-lane reads of the score stage
    const int lk = ((tl % LPI) >> 4) * 8 + (tl & 7);
    const bool lo_part = (tl & 8) != 0;
    uint32_t u = 0;
    PROF_DECL;
    // ---- scores of a finished unit (model.py:127-136,213-214).  These warps have the slack: while they do this the epilogue
    //      warps are already converting the blocks of the next unit.
    int64_t f_pc0 = 0, f_cs = 0, f_ce = 0;
    int f_nc = 0;
    auto score_stage = [&](uint32_t fu) {
      const int64_t pc0 = f_pc0, my_cs = f_cs, my_ce = f_ce;
      const int nc = f_nc;
      const bool row_ok = lk < K;
      tc::mbar_wait(&bars->dma_full, fu & 1);
      PROF_ADD(5);
      tc::tcgen05_fence_after();
      tc::named_bar_sync(1, T_SMX);                                            // the previous unit's score threads are done with Sm / Sa
      {
        // columns of this lane's impression inside the pass (warp-uniform: a warp's 32 lanes belong to one impression)
        const int64_t r_lo = my_cs - pc0, r_hi = my_ce - pc0;
        const int c_lo = static_cast<int>(r_lo < 0 ? 0 : (r_lo > nc ? nc : r_lo));
        int c_hi = static_cast<int>(r_hi < 0 ? 0 : (r_hi > nc ? nc : r_hi));
        if (c_hi < c_lo) c_hi = c_lo;
        for (int c0 = c_lo & ~15; c0 < c_hi; c0 += 16) {
          uint32_t vm[16], va[16];
          tc::tmem_ld_32x16(tmem + lane_addr + DM_COL + c0, vm);
          tc::tmem_ld_32x16(tmem + lane_addr + DA_COL + c0, va);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float xm = __uint_as_float(vm[c]), xa = __uint_as_float(va[c]);
            const float sm_ = xm + __shfl_xor_sync(0xffffffffu, xm, 8);       // A_hi . cand + A_lo . cand
            const float sa_ = xa + __shfl_xor_sync(0xffffffffu, xa, 8);
            const int col = c0 + c;
            if (!lo_part && row_ok && col >= c_lo && col < c_hi) {
              Sm[col * SS + lk] = sm_;
              Sa[col * SS + lk] = sa_;
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bars->dma_free);
      tc::named_bar_sync(1, T_SMX);
      if (st < nc) {
        const float* m = Sm + st * SS;
        const float* a = Sa + st * SS;
        float score;
        if (args.score_type == MINER_SCORE_WEIGHTED) {
          float mx = -INFINITY;
          for (int k = 0; k < K; ++k) mx = fmaxf(mx, a[k]);
          float den = 0.f, num = 0.f;
          for (int k = 0; k < K; ++k) {
            const float e = __expf(a[k] - mx);
            den += e;
            num = fmaf(e, m[k], num);
          }
          score = num / den;
        } else if (args.score_type == MINER_SCORE_MAX) {
          score = -INFINITY;
          for (int k = 0; k < K; ++k) score = fmaxf(score, m[k]);
        } else {
          score = 0.f;
          for (int k = 0; k < K; ++k) score += m[k];
          score /= static_cast<float>(K);
        }
        args.out_scores[pc0 + st] = score;
      }
      PROF_ADD(6);
    };
    {
      // the off-diagonal half of A_w (slots of the other impression) stays zero for the whole kernel
      uint32_t z[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) z[c] = 0u;
#pragma unroll
      for (int cc = 0; cc < 4 * NH; ++cc) tc::tmem_st_32x16(tmem + lane_addr + AW_COL + cc * 16, z);
      tc::tmem_st_wait();
    }
    int64_t t_cs = 0, t_ce = 0;
    if (n_local > 0) tile_range<IPT>(args, static_cast<int>(blockIdx.x), t_cs, t_ce);
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) tile_range<IPT>(args, tile + static_cast<int>(gridDim.x), t_cs, t_ce);      // a tile ahead: off the critical path
      for (int p = 0; p < npass; ++p, ++u) {
        PROF_ADD(0);
        tc::named_bar_sync(2, T_SMX);                                          // previous unit's reads of L are done
        {
          // logits of 32 slots per warp: lg rows are K consecutive floats (model.py:174 hoisted to the table).  Stored scaled by
          // log2(e): the softmax below runs on ex2.
          const uint32_t Ku = static_cast<uint32_t>(K);
          const bool has_bias = args.bias_mean != nullptr;
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {                                    // 128 slots per pass, 32 per warp
            const int slot = hh * TM + sw * 32 + lane;
            const int64_t imp = static_cast<int64_t>(tile) * IPT + slot / HP;
            const int h = slot % HP;
            const bool valid = h < H && imp < args.B;
            uint32_t info = 0;                                                 // id | code << 30; code 0 padding, 1 masked, 2 kept, 3 kept with a bad id
            float bias = 0.f;
            if (valid) {
              const int64_t id = load_id(args.his_ids, imp * H + h, args.id_dtype);
              const bool keep = args.mask[imp * H + h] != 0;
              if (args.bias_mean) bias = args.bias_mean[imp * H + h];
              const bool id_ok = id >= 0 && id < args.n_rows;
              info = (id_ok ? static_cast<uint32_t>(id) : 0u) | ((keep ? (id_ok ? 2u : 3u) : 1u) << 30);
            }
#pragma unroll
            for (int kk = 0; kk < KM / 32; ++kk) {                             // 32 codes per pass
              const int kcol = lane + 32 * kk;
              const float* lgp = args.lg + kcol;
              float v[32];
#pragma unroll
              for (int ss = 0; ss < 32; ++ss) {                                // 32 independent 128-byte row loads in flight
                const uint32_t info_s = __shfl_sync(0xffffffffu, info, ss);
                v[ss] = ((info_s >> 30) == 2u && kcol < K) ? lgp[(info_s & 0x3fffffffu) * Ku] : 0.f;
              }
#pragma unroll
              for (int ss = 0; ss < 32; ++ss) {
                const uint32_t code_s = __shfl_sync(0xffffffffu, info, ss) >> 30;
                float x = v[ss];
                if (has_bias) x += __shfl_sync(0xffffffffu, bias, ss);         // model.py:174-177
                if (code_s == 1u) x = kMaskFill;                               // model.py:180 (1e-30, not -inf)
                if (code_s == 0u) x = -INFINITY;                               // tile padding: not part of the history
                L[(hh * TM + sw * 32 + ss) * LS + kcol] = x * 1.4426950408889634f;
              }
            }
          }
        }
        tc::named_bar_sync(2, T_SMX);
        PROF_ADD(1);
        // softmax over the history (model.py:181): per 16-lane group, thread t owns code k = 8 group + t/4 and the 16 slots
        // {2c, 2c+1 : c = t%4 + 4n}; the (hi, lo) rows of the pair leave through one 16x128b store
        if constexpr (IPT == 1) {
          // one impression per tile (up to 128 NH slots): too many weights to hold in registers across the wait, so wait first,
          // take max and sum in two passes over the logits, then recompute, pack and store 32 packed columns at a time
          if (u > 0) tc::mbar_wait(&bars->w_free, (u - 1) & 1);                // S1 of the previous unit no longer reads A_w
          PROF_ADD(3);
          tc::tcgen05_fence_after();
          constexpr int NCT = HP / 8;                                          // packed columns per thread: c = t%4 + 4n, n < NCT
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int k = (((q * 32) % LPI) / 16 + hf) * 8 + (lane >> 2);
            const bool row_ok = k < K;
            const float* col = L + (row_ok ? k : 0);
            float mx = -INFINITY;
            for (int n = 0; n < NCT; ++n) {
              const int c = (lane & 3) + 4 * n;
              mx = fmaxf(mx, fmaxf(col[(2 * c) * LS], col[(2 * c + 1) * LS]));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const bool dead = mx == -INFINITY || !row_ok;                      // impression past the end of the batch / unused row
            float sum = 0.f;
            for (int n = 0; n < NCT; ++n) {
              const int c = (lane & 3) + 4 * n;
              sum += dead ? 0.f : ex2_approx(col[(2 * c) * LS] - mx) + ex2_approx(col[(2 * c + 1) * LS] - mx);
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = dead ? 0.f : __fdividef(1.0f, sum);
#pragma unroll
            for (int part = 0; part < NCT / 8; ++part) {
              uint32_t o[16];
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const int c = (lane & 3) + 4 * (8 * part + n);
                const float w0 = dead ? 0.f : ex2_approx(col[(2 * c) * LS] - mx) * inv;          // model.py:181
                const float w1 = dead ? 0.f : ex2_approx(col[(2 * c + 1) * LS] - mx) * inv;
                const uint32_t hi = pack2(w0, w1);
                o[2 * n] = hi;
                o[2 * n + 1] = pack2(w0 - __uint_as_float(hi << 16), w1 - __uint_as_float(hi & 0xffff0000u));
              }
              tc::tmem_st_16x128b_x8(tmem + (static_cast<uint32_t>(q * 32 + hf * 16) << 16) + AW_COL + part * 32, o);
            }
          }
          PROF_ADD(2);
        } else {
        constexpr int NC8 = HP / 8;                                            // packed columns per thread: c = t%4 + 4n, n < NC8
        uint32_t pk[2][2 * NC8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int k = (((q * 32) % LPI) / 16 + hf) * 8 + (lane >> 2);
          const bool row_ok = k < K;
          const float* col = L + (li * HP) * LS + (row_ok ? k : 0);
          float e[2 * NC8];
          float mx = -INFINITY;
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int c = (lane & 3) + 4 * n;
            e[2 * n] = col[(2 * c) * LS];
            e[2 * n + 1] = col[(2 * c + 1) * LS];
            mx = fmaxf(mx, fmaxf(e[2 * n], e[2 * n + 1]));
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const bool dead = mx == -INFINITY || !row_ok;                        // impression past the end of the batch / unused row
          float sum = 0.f;
#pragma unroll
          for (int n = 0; n < 2 * NC8; ++n) {
            e[n] = dead ? 0.f : ex2_approx(e[n] - mx);
            sum += e[n];
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          const float inv = dead ? 0.f : __fdividef(1.0f, sum);
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const float w0 = e[2 * n] * inv, w1 = e[2 * n + 1] * inv;
            const uint32_t hi = pack2(w0, w1);
            pk[hf][2 * n] = hi;
            pk[hf][2 * n + 1] = pack2(w0 - __uint_as_float(hi << 16), w1 - __uint_as_float(hi & 0xffff0000u));
          }
        }
        PROF_ADD(2);
        if (u > 0) tc::mbar_wait(&bars->w_free, (u - 1) & 1);                  // S1 of the previous unit no longer reads A_w
        PROF_ADD(3);
        tc::tcgen05_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
          tc::tmem_st_16x128b_x8(tmem + (static_cast<uint32_t>(q * 32 + hf * 16) << 16) + AW_COL + li * (HP / 2), pk[hf]);
        }
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->w_ready);
        PROF_ADD(4);
        if (u > 0) score_stage(u - 1);
        {
          const int64_t my_imp = static_cast<int64_t>(tile) * IPT + li;
          f_pc0 = cs + static_cast<int64_t>(p) * NCM;
          f_nc = static_cast<int>(ce - f_pc0 < NCM ? ce - f_pc0 : NCM);
          f_cs = my_imp < args.B ? cand_off(args, my_imp) : ce;
          f_ce = my_imp < args.B ? cand_off(args, my_imp + 1) : ce;
        }
      }
    }
    if (u > 0) score_stage(u - 1);
    if (threadIdx.x == W_SMX0 * 32) PROF_STORE(3);
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc(tmem, 512);
}


}  // namespace

bool tscore_kernel_supported(int64_t H, int64_t K, int64_t D) {
  return H >= 1 && H <= 2 * TM && K >= 1 && K <= KMAX && D >= FB && D % FB == 0 && D <= 8192;
}

int launch_tscore_kernel(const void* table, const void* tw, const float* lg, int64_t n_rows, const void* his_ids, int id_dtype,
                         const uint8_t* his_mask, const float* bias_mean, const void* cand_ids, const int64_t* cand_offsets,
                         int64_t B, int64_t H, int64_t C, int64_t K, int64_t D, int score_type, float* out_scores, float* out_interests,
                         cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  if (!tscore_kernel_supported(H, K, D)) {
    set_error("table-level scoring: unsupported shape H=%lld K=%lld D=%lld (need H <= 256, K <= 64, D %% 64 == 0)", (long long)H, (long long)K,
              (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  if (static_cast<uint64_t>(n_rows) * static_cast<uint64_t>(D) * 2 > 0xffffffffull || n_rows >= (1ll << 30) ||
      static_cast<uint64_t>(n_rows) * static_cast<uint64_t>(K) > 0xffffffffull) {
    set_error("table-level scoring: table of %lld x %lld bf16 exceeds the 4 GB the kernel addresses with 32-bit offsets", (long long)n_rows,
              (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  TScoreArgs a;
  a.table = static_cast<const uint16_t*>(table); a.tw = static_cast<const uint16_t*>(tw); a.lg = lg; a.n_rows = n_rows;
  a.his_ids = his_ids; a.cand_ids = cand_ids; a.id_dtype = id_dtype; a.mask = his_mask; a.bias_mean = bias_mean;
  a.cand_offsets = cand_offsets; a.B = B; a.H = static_cast<int>(H); a.K = static_cast<int>(K); a.D = static_cast<int>(D);
  a.C = static_cast<int>(C); a.score_type = score_type; a.out_scores = out_scores; a.out_interests = out_interests;
  a.prof = hist_prof_buffer();
  {
    static const char* env_dbg = getenv("MINER_TS_DBG");
    a.dbg = env_dbg ? atoi(env_dbg) : 0;
  }
  // lanes per impression: 2 K (hi / lo rows of every code) -> two impressions share a tile only if K <= 32 and H <= 64
  const int ipt = (H <= TM / 2 && K <= 32) ? 2 : 1;
  const int64_t n_tiles = (B + ipt - 1) / ipt;
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
#define MINER_TS_LAUNCH(I, KMV, NHV)                                                                                                    \
  do {                                                                                                                                  \
    MINER_CUDA_OK(cudaFuncSetAttribute(tscore_kernel<I, KMV, NHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Shape<KMV, NHV>::SMEM)); \
    tscore_kernel<I, KMV, NHV><<<grid, T_THREADS, Shape<KMV, NHV>::SMEM, stream>>>(a, static_cast<int>(n_tiles));                      \
  } while (0)
  if (H > TM && K > 32) MINER_TS_LAUNCH(1, 64, 2);          // 128 < H <= 256: two 128-slot halves accumulate into one D_I | D_P pair
  else if (H > TM) MINER_TS_LAUNCH(1, 32, 2);
  else if (K > 32) MINER_TS_LAUNCH(1, 64, 1);
  else if (ipt == 2) MINER_TS_LAUNCH(2, 32, 1);
  else MINER_TS_LAUNCH(1, 32, 1);
#undef MINER_TS_LAUNCH
  MINER_LAUNCH_OK("tscore_kernel");
  return MINER_OK;
}

}  // namespace miner
