// Table-level scoring kernel, X formulation (sm_100a: tcgen05 + TMEM, cp.async gathers).  Same mathematics and same packed tiles
// as tscore_kernel.cu (Miner.forward, reference src/model/model.py:61-138, with both nn.Linear layers hoisted to the news table),
// for the shapes the headline workload uses: two impressions per tile, K <= 32, H <= 56, scores only (no interests output).
//
// The matching scores never needed the interests themselves:
//     m[k,c] = I[k] . cand[c] = sum_h w[k,h] (E_h . cand_c) = sum_h w[k,h] X[h,c],      X = E cand^T                 (model.py:127,182)
// X is ONE SS-mode MMA chain over the gathered tiles (no conversion, no epilogue pass), accumulated over all 64-feature blocks
// in tensor memory; the small contraction with the softmax weights (K x slots x candidates, ~35 kFLOP per tile) runs on the CUDA
// cores of the score warps at the end of the tile, with the weights read back from tensor memory as bf16 hi + lo.  Only
//     P[k,:] = sum_h w[k,h] tw[his[h],:],  G = gelu(P),  a[c,k] = cand[c] . G[k]                                     (model.py:212-213)
// still goes through the per-block chain  S1 -> epilogue (pair sum, gelu, hi/lo split) -> S2, which is the latency chain that
// bounds tscore_kernel: the epilogue converts half as much per block, S1 / S2 issue half the MMA work, and the D_m drain is gone.
// What the freed tensor memory buys: the softmax weights A_w and the attention logits D_a are DOUBLE-buffered, so neither the
// softmax warps (a tile ahead) nor the score warps (a tile behind) are on the MMA issuer's critical path.
//
// TMEM map (496 of 512 columns): A_w 2 x 64 | D_P 2 x 64 (the first 32 columns of a buffer become packed G hi|lo) | D_X 80 | D_a 2 x 80.
// TMEM lanes of A_w / D_P / D_a are (impression i, code k, part hl) exactly as in tscore_kernel.cu; D_X has lane = history slot.
//   warps 0-3   gather table[id] (E) and tw[id] rows of the tile's slots per 64-feature block (4-stage ring of 28 KB)
//   warp 4      S1 / S2 issuer, per block j:  S1(j): D_P[b] = A_w . TW_j (TS form, B = the TW half of the stage read MN-major);
//                                             S2(j - 1): D_a[u & 1] += A_G[b'] . cand_(j-1)^T (TS form)
//   warp 5      SX issuer:                    D_X += E_j . cand_j^T (SS form, both operands K-major as gathered)
//   warps 6-7   gather the tile's candidate rows (5-stage ring of 10 KB)
//   warps 8-15  epilogue of D_P per block (pair sum, gelu, bf16 hi/lo split, in place)
//   warps 16-19 softmax over the history (model.py:174-181) from the lg rows (cp.async) -> A_w[u & 1]
//   warps 20-23 per finished unit: A_w -> W and D_X -> Xs in shared memory, m = W X on the CUDA cores (4 codes x 4 candidates per
//               thread), drain D_a[u & 1], softmax over K, scores
#include "tscore_common.cuh"

namespace miner {

namespace {

using namespace ts;

constexpr int IPT = 2, LPI = TM / IPT;       // two impressions per tile, 64 TMEM lanes each
constexpr int KM = 32;                       // context codes covered (K <= 32)
constexpr int SL = 112;                      // slot capacity of a tile: IPT * ((H + 1) & ~1) <= 112, i.e. at most 7 16-slot K-steps
constexpr int NKS_MAX = SL / 16;
constexpr int NCM = 80;                      // candidate columns per pass
#ifndef MINER_TSX_S2
#define MINER_TSX_S2 5
#endif
constexpr int S1 = 4, S2 = MINER_TSX_S2;     // ring depths: (E, TW) blocks / candidate blocks
constexpr int E_BYTES = SL * FB * 2;         // 14 KB
constexpr int ST1_BYTES = 2 * E_BYTES;       // E | TW
constexpr int C_BYTES = NCM * FB * 2;        // 10 KB
constexpr int LS = KM + 4;                   // logits scratch row stride (floats): column-wise reads cover the 32 banks once
constexpr int SS = KM + 4;                   // score scratch row stride (16-byte aligned rows: four codes per thread)
constexpr int CC = 32;                       // candidate columns of an impression per contraction round
constexpr int XS = CC + 4;                   // X scratch row stride
constexpr int L_FLOATS = SL * LS;
constexpr int SM_FLOATS = NCM * SS;
constexpr int XA_FLOATS = (SL * XS > NCM * SS) ? SL * XS : NCM * SS;   // X scratch and the attention-logit transposes share their bytes
constexpr int SMEM = 1024 + S1 * ST1_BYTES + S2 * C_BYTES + (2 * L_FLOATS + SM_FLOATS + XA_FLOATS) * 4 + 512;   // L + W, Sm, Xs | Sa
static_assert(SMEM <= 232448, "shared memory budget");
static_assert(E_BYTES % 1024 == 0 && C_BYTES % 1024 == 0, "swizzle atoms are 1 KB");

constexpr int T_EPI = 256, T_SMX = 128, T_SCR = 128;
#ifndef MINER_TSX_TG1
#define MINER_TSX_TG1 128
#endif
constexpr int T_G1 = MINER_TSX_TG1, T_G2 = 64;
constexpr int G1_STEP = T_G1 / 8, G2_STEP = T_G2 / 8;
constexpr int G1_ROWS = SL / G1_STEP, G2_ROWS = NCM * 8 / T_G2;
static_assert(G1_STEP == 16 || G1_STEP == 8, "a gather thread's row jj lies in 16-row group jj * G1_STEP / 16: one K-step of S1");
// warp roles (a warp's scheduler and its TMEM lane quarter are warp % 4: the epilogue / softmax / score groups are multiples of four
// warps, one per quarter).  Two gather warps for the (E, TW) ring instead of four were measured slower (36.1 vs 38.4 M impressions/s).
constexpr int W_G1 = 0, W_MMA = W_G1 + T_G1 / 32, W_SX = W_MMA + 1, W_G2 = W_SX + 1, W_EPI0 = W_G2 + T_G2 / 32, W_SMX0 = W_EPI0 + T_EPI / 32,
              W_SCR0 = W_SMX0 + T_SMX / 32;
constexpr int T_THREADS = (W_SCR0 + T_SCR / 32) * 32;
#ifndef MINER_TSX_NDP
#define MINER_TSX_NDP 2
#endif
constexpr int NDP = MINER_TSX_NDP;           // D_P buffers; S2 of a block is issued LAG = NDP - 1 blocks after its S1, so the epilogue warps
constexpr int LAG = NDP - 1;                 // always have a finished S1 waiting for them
constexpr int NDA = NDP == 2 ? 2 : 1;        // D_a buffers (what fits)
constexpr uint32_t AW_COL = 0, DP_COL = 128, DX_COL = DP_COL + 64 * NDP, DA_COL = DX_COL + NCM;   // A_w x 2 | D_P x NDP | D_X | D_a x NDA
static_assert(DA_COL + NDA * NCM <= 512 && DX_COL + NCM - 1 + CC <= 512, "tensor memory map");
static_assert(S2 >= LAG + 2, "a candidate stage is held from SX of its block to S2 of its block");

// ablation switches for A/B timing runs (results are wrong by construction): -DMINER_TSX_ABL=<bits>
//   1 score warps without their arithmetic   2 softmax warps store zeros   4 epilogue without the gelu   8 no SX MMAs   16 no S2 MMAs
//   32 gathers without global reads
#ifndef MINER_TSX_ABL
#define MINER_TSX_ABL 0
#endif
constexpr int ABL = MINER_TSX_ABL;
// waits of roles that run ahead of their consumer back off between probes (bits: 1 gathers, 2 softmax, 4 score warps, 8 epilogue, 16 SX issuer):
// a failed mbarrier probe loop otherwise takes issue slots from the warps doing the work (ncu: 40 % of the issued instructions)
#ifndef MINER_TSX_RELAX
#define MINER_TSX_RELAX 7
#endif
#ifndef MINER_TSX_RELAX_NS
#define MINER_TSX_RELAX_NS 128
#endif
constexpr int RELAX = MINER_TSX_RELAX;
template <int BIT>
__device__ __forceinline__ void role_wait(uint64_t* bar, uint32_t parity) {
  if (RELAX & BIT) {
    while (!tc::mbar_try_wait(bar, parity)) __nanosleep(MINER_TSX_RELAX_NS);
  } else {
    tc::mbar_wait(bar, parity);
  }
}

struct XBarriers {
  uint64_t full1[S1], empty1[S1], full2[S2], empty2[S2];
  uint64_t w_ready[2], w_free[2], ip_full[NDP], a_ready[NDP], x_full, x_free, dma_full[NDA], dma_free[NDA];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// mirror of tmem_st_16x128b_x8: reg 2n = lane t/4, column 4n + t%4; reg 2n+1 = lane t/4 + 8, same column
__device__ __forceinline__ void tmem_ld_16x128b_x8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x128b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// what a role keeps about a tile it will work on later: loaded a tile (or two) ahead, looked at only when its turn comes
struct TileInfo {
  uint2 hd;
  int64_t cs, ce;
};
__device__ __forceinline__ void fetch_info(const TScoreArgs& a, int tile, TileInfo& t) {
  t.hd = a.hdr[tile];
  tile_range<IPT>(a, tile, t.cs, t.ce);
}

__global__ void __launch_bounds__(T_THREADS, 1)
tscore_x_kernel(const TScoreArgs args, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET (not by integer arithmetic on the pointer) so that the compiler keeps the shared address space: LDS / STS, not generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* st1 = smem;                                         // [S1][E 14 KB | TW 14 KB]
  uint8_t* st2 = st1 + S1 * ST1_BYTES;                         // [S2][10 KB] candidate rows
  float* L = reinterpret_cast<float*>(st2 + S2 * C_BYTES);     // [SL slots][LS] logits of the unit the softmax warps prepare
  float* W = L + L_FLOATS;                                     // [SL slots][LS] softmax weights of the unit the score warps finish
  float* Sm = W + L_FLOATS;                               // [NCM][SS] matching scores, transposed
  float* Xs = Sm + SM_FLOATS;                                  // [SL][XS] X rows of a contraction round; later
  float* Sa = Xs;                                              // [NCM][SS] attention logits, transposed (same bytes)
  XBarriers* bars = reinterpret_cast<XBarriers*>(Xs + XA_FLOATS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, K = args.K, D = args.D;
  const int KB = D / FB;
  const int tile0 = static_cast<int>(blockIdx.x), tstep = static_cast<int>(gridDim.x);
  const int n_local = (n_tiles - tile0 + tstep - 1) / tstep;

  if (threadIdx.x == 0) {
    // a ring stage is released by two commits: the S1 / S2 issuer's and the SX issuer's
    for (int s = 0; s < S1; ++s) { tc::mbar_init(&bars->full1[s], T_G1); tc::mbar_init(&bars->empty1[s], 2); }
    for (int s = 0; s < S2; ++s) { tc::mbar_init(&bars->full2[s], T_G2); tc::mbar_init(&bars->empty2[s], 2); }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&bars->w_ready[b], T_SMX); tc::mbar_init(&bars->w_free[b], T_SCR + 1);
    }
    for (int b = 0; b < NDP; ++b) { tc::mbar_init(&bars->ip_full[b], 1); tc::mbar_init(&bars->a_ready[b], T_EPI); }
    for (int b = 0; b < NDA; ++b) { tc::mbar_init(&bars->dma_full[b], 1); tc::mbar_init(&bars->dma_free[b], T_SCR); }
    tc::mbar_init(&bars->x_full, 1);
    tc::mbar_init(&bars->x_free, T_SCR);
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  if (bars->tmem_base != 0u) __trap();                         // the CTA owns the SM's tensor memory: every address below is a constant
  constexpr uint32_t tmem = 0u;

  if (warp < W_MMA) {
    // ------------------------------------------------------------------ gathers of the (E, TW) ring: thread = one 16-byte chunk of rows
    //        r0 + 16 jj, i.e. one row of every 16-slot group (K-step) of the tile
    const int t = threadIdx.x;
    const int chunk = t & 7, r0 = t >> 3;
    const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
    const char* table_b = reinterpret_cast<const char*>(args.table);
    const char* tw_b = reinterpret_cast<const char*>(args.tw);
    const uint32_t dst0 = tc::sw128_offset(r0, chunk);         // row r0 + 16 jj sits 2 jj KB further
    uint32_t rec_pre[G1_ROWS];                                 // raw slot records (x word) of the next tile
    TileInfo nx;
    auto fetch_recs = [&](int tile) {
      const uint32_t* rt = reinterpret_cast<const uint32_t*>(args.rec + static_cast<int64_t>(tile) * TM);
#pragma unroll
      for (int jj = 0; jj < G1_ROWS; ++jj) rec_pre[jj] = rt[2 * (r0 + G1_STEP * jj)];
      fetch_info(args, tile, nx);
    };
    uint32_t g = 0;
    PROF_DECL;
    if (n_local > 0) fetch_recs(tile0);
    for (int lt = 0; lt < n_local; ++lt) {
      uint32_t eoff[G1_ROWS];
      uint32_t emask = 0;
      const int nks = (hdr_end(nx.hd, IPT - 1) + 15) >> 4;
#pragma unroll
      for (int jj = 0; jj < G1_ROWS; ++jj) {
        const uint32_t rc = rec_pre[jj];
        eoff[jj] = (rc & REC_ID) * row_bytes + chunk * 16;
        emask |= (rc & REC_VALID) ? (1u << jj) : 0u;
      }
      const int npass = passes_of<NCM>(nx.cs, nx.ce);
      if (lt + 1 < n_local) fetch_recs(tile0 + (lt + 1) * tstep);
      for (int p = 0; p < npass; ++p) {
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t s = g % S1, ph = (g / S1) & 1;
          PROF_ADD(0);
          role_wait<1>(&bars->empty1[s], ph ^ 1);
          PROF_ADD(1);
          const uint32_t base = tc::smem_u32(st1 + s * ST1_BYTES) + dst0;
          const uint32_t jb = static_cast<uint32_t>(j) * (FB * 2);
#pragma unroll
          for (int jj = 0; jj < G1_ROWS; ++jj) {
            if (jj * G1_STEP < nks * 16) {
              const uint32_t o = eoff[jj] + jb, sz = (((emask >> jj) & 1u) && !(ABL & 32)) ? 16u : 0u;
              tc::cp_async_16(base + jj * (G1_STEP * 128), table_b + o, sz);
              tc::cp_async_16(base + E_BYTES + jj * (G1_STEP * 128), tw_b + o, sz);
            }
          }
          tc::cp_async_mbar_arrive_noinc(&bars->full1[s]);
          PROF_ADD(2);
        }
      }
    }
    tc::cp_async_wait_all();
    if (threadIdx.x == 0) PROF_STORE(2);
  } else if (warp >= W_G2 && warp < W_EPI0) {
    // ------------------------------------------------------------------ gathers of the candidate ring: thread = one 16-byte chunk of
    //        rows r0 + 8 jj; the candidate ids of the next unit are fetched one unit ahead and kept raw
    const int t = threadIdx.x - W_G2 * 32;
    const int chunk = t & 7, r0 = t >> 3;
    const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
    const char* table_b = reinterpret_cast<const char*>(args.table);
    const uint32_t dst0 = tc::sw128_offset(r0, chunk);
    RawId cid_pre[G2_ROWS];
    auto fetch_cands = [&](int64_t pc0, int nc) {
#pragma unroll
      for (int jj = 0; jj < G2_ROWS; ++jj) {
        const int c = r0 + G2_STEP * jj;
        cid_pre[jj] = RawId{0u, 0u};
        if (c < nc) cid_pre[jj] = load_id_raw(args.cand_ids, pc0 + c, args.id_dtype);
      }
    };
    uint32_t g = 0;
    int bad = 0;
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
          const bool in = r0 + G2_STEP * jj < nc, ok = in && id >= 0 && id < args.n_rows;
          coff[jj] = static_cast<uint32_t>(ok ? id : 0) * row_bytes + chunk * 16;
          cmask |= ok ? (1u << jj) : 0u;
          if (in && !ok && chunk == 0) ++bad;
        }
        if (p + 1 < npass) {
          const int64_t q0 = pc0 + NCM;
          fetch_cands(q0, static_cast<int>(ce - q0 < NCM ? ce - q0 : NCM));
        } else if (lt + 1 < n_local) {
          fetch_cands(ncs, static_cast<int>(nce - ncs < NCM ? nce - ncs : NCM));
        }
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t s = g % S2, ph = (g / S2) & 1;
          role_wait<1>(&bars->empty2[s], ph ^ 1);
          const uint32_t base = tc::smem_u32(st2 + s * C_BYTES) + dst0;
          const uint32_t jb = static_cast<uint32_t>(j) * (FB * 2);
#pragma unroll
          for (int jj = 0; jj < G2_ROWS; ++jj)
            if (G2_STEP * jj < nc16)
              tc::cp_async_16(base + jj * (G2_STEP * 128), table_b + (coff[jj] + jb), (((cmask >> jj) & 1u) && !(ABL & 32)) ? 16u : 0u);
          tc::cp_async_mbar_arrive_noinc(&bars->full2[s]);
        }
      }
      cs = ncs; ce = nce; ncs = n2cs; nce = n2ce;
    }
    tc::cp_async_wait_all();
    if (bad > 0 && args.oob) atomicAdd(args.oob + 1, bad);     // candidate ids outside the table (their rows read as zero)
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tc::make_idesc_bf16_f32_major(TM, FB, false, true);   // S1: N = 64 features, B (the TW half of the stage) MN-major
    uint32_t g1 = 0, g2 = 0, u = 0;                              // blocks issued (S1 / S2), units
    // blocks whose S1 is issued and whose S2 is still to come (at most LAG): block index in its unit | nc16 << 8 | unit << 16
    uint32_t pq[LAG] = {0u};
    int npend = 0;
    PROF_DECL;
    auto stage2 = [&]() {                                        // S2 of block g2 = the oldest pending block
      const uint32_t pend = pq[0];
#pragma unroll
      for (int i = 0; i + 1 < LAG; ++i) pq[i] = pq[i + 1];
      --npend;
      const int pend_j = static_cast<int>(pend & 0xffu), pend_nc16 = static_cast<int>((pend >> 8) & 0xffu);
      const uint32_t pend_u = pend >> 16;
      const uint32_t b = g2 % NDP;
      PROF_ADD(0);
      tc::mbar_wait(&bars->a_ready[b], (g2 / NDP) & 1);
      PROF_ADD(4);
      const uint32_t db = NDA == 2 ? (pend_u & 1) : 0;           // D_a buffer of the unit
      if (pend_j == 0) tc::mbar_wait(&bars->dma_free[db], (((NDA == 2 ? pend_u >> 1 : pend_u) & 1) ^ 1));   // the previous user's attention logits are out of it
      PROF_ADD(6);
      tc::tcgen05_fence_after();
      const uint32_t s = g2 % S2;
      tc::mbar_wait(&bars->full2[s], (g2 / S2) & 1);
      PROF_ADD(5);
      const uint32_t idesc2 = tc::make_idesc_bf16_f32(TM, pend_nc16);
      const uint64_t c_desc = tc::make_smem_desc_sw128(tc::smem_u32(st2 + s * C_BYTES));
      const uint32_t a_g = tmem + DP_COL + b * 64;
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ((ABL & 16) ? 0 : FB / 16); ++ks)
          tc::umma_bf16_ts(tmem + DA_COL + NCM * db, a_g + 8 * ks, c_desc + 2 * ks, idesc2, (pend_j | ks) != 0 ? 1u : 0u);
        tc::umma_commit(&bars->empty2[s]);
        if (pend_j == KB - 1) tc::umma_commit(&bars->dma_full[db]);
      }
      __syncwarp();
      ++g2;
      PROF_ADD(7);
    };
    TileInfo nx;
    if (n_local > 0) fetch_info(args, tile0, nx);
    for (int lt = 0; lt < n_local; ++lt) {
      const int64_t cs = nx.cs, ce = nx.ce;
      const int nks = (hdr_end(nx.hd, IPT - 1) + 15) >> 4;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) fetch_info(args, tile0 + (lt + 1) * tstep, nx);      // a tile ahead: off the critical path
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        const uint32_t wb = u & 1;
        PROF_ADD(0);
        tc::mbar_wait(&bars->w_ready[wb], (u >> 1) & 1);
        PROF_ADD(1);
        tc::tcgen05_fence_after();
        const uint32_t aw = tmem + AW_COL + 64 * wb;
        for (int j = 0; j < KB; ++j) {
          const uint32_t b = g1 % NDP;
          const uint32_t d_p = tmem + DP_COL + b * 64;
          const uint32_t s = g1 % S1, ph = (g1 / S1) & 1;
          PROF_ADD(0);
          tc::mbar_wait(&bars->full1[s], ph);
          PROF_ADD(2);
          tc::tcgen05_fence_after();
          const uint32_t st_addr = tc::smem_u32(st1 + s * ST1_BYTES);
          const uint64_t tw_desc = tc::make_smem_desc_sw128_mn(st_addr + E_BYTES);
          if (tc::elect_one()) {
#ifdef MINER_TSX_S1_UNROLL
#pragma unroll
            for (int ks = 0; ks < NKS_MAX; ++ks)
              if (ks < nks) tc::umma_bf16_ts(d_p, aw + 8 * ks, tw_desc + ks * (2048 >> 4), idesc1, ks != 0 ? 1u : 0u);
#else
            for (int ks = 0; ks < nks; ++ks) tc::umma_bf16_ts(d_p, aw + 8 * ks, tw_desc + ks * (2048 >> 4), idesc1, ks != 0 ? 1u : 0u);
#endif
            tc::umma_commit(&bars->ip_full[b]);
            tc::umma_commit(&bars->empty1[s]);
            if (j == KB - 1) tc::umma_commit(&bars->w_free[wb]);   // (the score warps read A_w too: they arrive on it themselves)
          }
          __syncwarp();
          PROF_ADD(3);
          ++g1;
          PROF_ADD(9);
          if (npend == LAG) stage2();
          pq[LAG - 1] = static_cast<uint32_t>(j) | (static_cast<uint32_t>(nc16) << 8) | (u << 16);
          if (npend < LAG - 1) {                                 // (start-up only: keep the queue packed at its front)
#pragma unroll
            for (int i = 0; i + 1 < LAG; ++i) if (i == npend) pq[i] = pq[LAG - 1];
          }
          ++npend;
        }
      }
    }
    while (npend > 0) stage2();
    if (lane == 0) PROF_STORE(0);
  } else if (warp >= W_EPI0 && warp < W_SMX0) {
    // ------------------------------------------------------------------ epilogue warps: P -> G = gelu(P) as bf16 hi | lo, in place
    const int ew = warp - W_EPI0;
    const int q = warp & 3;                                    // TMEM lane quarter
    const int half = ew >> 2;                                  // 16-lane group of the quarter
    const uint32_t grp_addr = static_cast<uint32_t>(q * 32 + half * 16) << 16;
    uint32_t g = 0;
    TileInfo nx;
    PROF_DECL;
    if (n_local > 0) fetch_info(args, tile0, nx);
    for (int lt = 0; lt < n_local; ++lt) {
      const int npass = passes_of<NCM>(nx.cs, nx.ce);
      if (lt + 1 < n_local) fetch_info(args, tile0 + (lt + 1) * tstep, nx);
      const int nblk = npass * KB;
      for (int jb = 0; jb < nblk; ++jb, ++g) {
        const uint32_t b = g % NDP;
        PROF_ADD(0);
        role_wait<8>(&bars->ip_full[b], (g / NDP) & 1);
        PROF_ADD(1);
        tc::tcgen05_fence_after();
        const uint32_t acc = tmem + grp_addr + DP_COL + b * 64;
        uint32_t vp[32];
        tc::tmem_ld_16x256b_x8(acc, vp);                       // rows (hi, lo) of this thread's code x 16 of the 64 features
        tc::tmem_ld_wait();
        PROF_ADD(5);
        uint32_t pk[16];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const unsigned long long s = f2_add(f2_pack(__uint_as_float(vp[4 * n]), __uint_as_float(vp[4 * n + 1])),
                                              f2_pack(__uint_as_float(vp[4 * n + 2]), __uint_as_float(vp[4 * n + 3])));
          if (ABL & 4) { pk[2 * n] = static_cast<uint32_t>(s); pk[2 * n + 1] = static_cast<uint32_t>(s >> 32); }
          else split_hi_lo(gelu2(s), pk[2 * n], pk[2 * n + 1]);     // model.py:212
        }
        PROF_ADD(6);
        tc::tmem_st_16x128b_x8(acc, pk);
        tc::tmem_st_wait();
        PROF_ADD(8);
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->a_ready[b]);
        PROF_ADD(2);
      }
    }
    if (ew == 0 && lane == 0) PROF_STORE(1);
  } else if (warp >= W_SMX0 && warp < W_SCR0) {
    // ------------------------------------------------------------------ softmax warps: lg rows of the tile's slots (cp.async, issued
    //        as soon as the previous unit's logits have been read) -> softmax over the history -> A_w[u & 1] in tensor memory
    const int sw = warp - W_SMX0;
    const int q = warp & 3;
    const int c4 = lane & 3;
    const int my_slot = 4 * lane + sw;                         // lane l of warp sw owns slot 4 l + sw
    const bool k_vec4 = (K & 3) == 0;                          // lg rows are then 16-byte aligned
    constexpr float LOG2E = 1.4426950408889634f;
    // tiles lt (A), lt + 1 (B), lt + 2 (C): slot record of this lane + header + candidate range, loaded two tiles ahead
    uint2 recA = make_uint2(0u, 0u), recB = recA, recC = recA;
    TileInfo tA, tB, tC;
    tA.hd = tB.hd = tC.hd = make_uint2(0u, 0u);
    tA.cs = tA.ce = tB.cs = tB.ce = tC.cs = tC.ce = 0;
    auto fetch_tile = [&](int lt, uint2& rec, TileInfo& ti) {
      const int tile = tile0 + lt * tstep;
      rec = args.rec[static_cast<int64_t>(tile) * TM + my_slot];
      fetch_info(args, tile, ti);
    };
    // start the copy of a unit's logits into L: kept slots by cp.async from their lg row (model.py:174 hoisted to the table),
    // masked records filled with the constant the reference overwrites them with (model.py:180; multiplicity n adds ln n)
    auto issue_fill = [&](uint2 rec, uint2 hd) {
      const int n_tot = hdr_end(hd, IPT - 1);
      if (my_slot < n_tot) {
        const uint32_t info = rec.x, mult = rec.y & 0xffffu;
        float* dst = L + my_slot * LS;
        if (mult != 0u && !(info & REC_MASKED)) {
          const float* row = args.lg + static_cast<size_t>(info & REC_ID) * K;
          if (!(info & REC_VALID)) {
#pragma unroll 1
            for (int k = 0; k < K; ++k) dst[k] = 0.f;          // an id outside the table reads as a zero row
          } else if (k_vec4) {
            const uint32_t d = tc::smem_u32(dst);
#pragma unroll
            for (int i = 0; i < KM / 4; ++i)
              if (4 * i < K) tc::cp_async_16(d + 16 * i, row + 4 * i, 16u);
          } else {
#pragma unroll 1
            for (int k = 0; k < K; ++k) dst[k] = row[k];
          }
        } else if (mult != 0u) {
          const float x = kMaskFill + __logf(static_cast<float>(mult));
          if (k_vec4) {
#pragma unroll 1
            for (int k = 0; k < K; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(x, x, x, x);
          } else {
#pragma unroll 1
            for (int k = 0; k < K; ++k) dst[k] = x;
          }
        }
      }
      tc::cp_async_commit();
    };
    if (n_local > 0) fetch_tile(0, recA, tA);
    if (n_local > 1) fetch_tile(1, recB, tB);
    if (n_local > 0) issue_fill(recA, tA.hd);
    uint32_t u = 0;
    PROF_DECL;
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = tile0 + lt * tstep;
      const int npass = passes_of<NCM>(tA.cs, tA.ce);
      if (lt + 2 < n_local) fetch_tile(lt + 2, recC, tC);
      const uint2 hd = tA.hd;
      const int n_tot = hdr_end(hd, IPT - 1);
      const int nks = (n_tot + 15) >> 4;
      for (int p = 0; p < npass; ++p, ++u) {
        PROF_ADD(0);
        tc::cp_async_wait<0>();                                // this thread's row of the unit has landed
        if (args.bias_mean && my_slot < n_tot) {
          // category bias of a kept slot (model.py:176-177): added to this lane's own row before anybody reads it
          const uint32_t info = recA.x, meta = recA.y;
          if ((meta & 0xffffu) != 0u && !(info & REC_MASKED)) {
            const float bias = args.bias_mean[(static_cast<int64_t>(tile) * IPT + ((meta >> 24) & 0xfu)) * H + ((meta >> 16) & 0xffu)];
            float* dst = L + my_slot * LS;
#pragma unroll 1
            for (int k = 0; k < K; ++k) dst[k] += bias;
          }
        }
        tc::named_bar_sync(2, T_SMX);                          // every row of the unit is in L
        PROF_ADD(1);
        const uint32_t wb = u & 1;
        if (u >= 2) role_wait<2>(&bars->w_free[wb], ((u >> 1) - 1) & 1);        // S1 of unit u - 2 no longer reads this buffer
        PROF_ADD(3);
        tc::tcgen05_fence_after();
        // softmax over the history (model.py:181): per 16-lane group, thread t owns code k = 8 group + t/4 and the slots
        // {2c, 2c+1 : c = t%4 + 4n} of its impression's range [s0, s1); the (hi, lo) rows of the pair leave through 16x128b stores.
        // Every column the MMAs of this tile read (8 per 16-slot group) is rewritten, zeros outside the impression's own range.
        const uint32_t aw = tmem + AW_COL + 64 * wb;
        const int nparts = (nks * 8 + 31) >> 5;                // 32-column stores per lane group
        // (the kernel's code must stay inside the instruction cache: the per-tile roles are rolled loops, not unrolled bodies)
#pragma unroll 1
        for (int hf = 0; hf < ((ABL & 2) ? 0 : 2); ++hf) {
          const int g0 = q * 32 + hf * 16;
          const int li = g0 / LPI;
          const int k = ((g0 % LPI) / 16) * 8 + (lane >> 2);
          const int s0 = hdr_start(hd, li), s1 = hdr_end(hd, li);
          const bool row_ok = k < K && s1 > s0;
          float* col = L + (row_ok ? k : 0) + 2 * c4 * LS;     // this thread's slots: {8n + 2c4, 8n + 2c4 + 1}
          const int n_lo = s0 >> 3, n_hi = (s1 + 7) >> 3;      // 8-slot groups holding the impression (warp-uniform)
          float mx = -INFINITY;
#pragma unroll 1
          for (int n = n_lo; n < n_hi; ++n) {
            const int sa = 8 * n + 2 * c4;
            const float a = (sa >= s0 && sa < s1) ? col[8 * n * LS] : -INFINITY;
            const float b = (sa + 1 >= s0 && sa + 1 < s1) ? col[(8 * n + 1) * LS] : -INFINITY;
            mx = fmaxf(mx, fmaxf(a, b));
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const bool dead = mx == -INFINITY || !row_ok;        // impression past the end of the batch / unused row
          const float nmx = -mx * LOG2E;
          float sum = 0.f;
#pragma unroll 1
          for (int n = n_lo; n < n_hi; ++n) {                  // exp2((logit - max) log2 e) goes back to L in place
            const int sa = 8 * n + 2 * c4;
            if (!dead && sa >= s0 && sa < s1) {
              const float e = ex2_approx(fmaf(col[8 * n * LS], LOG2E, nmx));
              col[8 * n * LS] = e;
              sum += e;
            }
            if (!dead && sa + 1 >= s0 && sa + 1 < s1) {
              const float e = ex2_approx(fmaf(col[(8 * n + 1) * LS], LOG2E, nmx));
              col[(8 * n + 1) * LS] = e;
              sum += e;
            }
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          const float inv = dead ? 0.f : __fdividef(1.0f, sum);
#pragma unroll 1
          for (int part = 0; part < nparts; ++part) {
            uint32_t o[16];
#pragma unroll
            for (int nn = 0; nn < 8; ++nn) {
              const int n = 8 * part + nn;
              const int sa = 8 * n + 2 * c4;
              const bool in0 = !dead && sa >= s0 && sa < s1, in1 = !dead && sa + 1 >= s0 && sa + 1 < s1;
              const float w0 = in0 ? col[8 * n * LS] * inv : 0.f;                  // model.py:181
              const float w1 = in1 ? col[(8 * n + 1) * LS] * inv : 0.f;
              split_hi_lo(f2_pack(w0, w1), o[2 * nn], o[2 * nn + 1]);
            }
            tc::tmem_st_16x128b_x8(aw + (static_cast<uint32_t>(g0) << 16) + part * 32, o);
          }
        }
        PROF_ADD(2);
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->w_ready[wb]);
        tc::named_bar_sync(2, T_SMX);                          // everybody has read L: the next unit's rows may come in
        if (p + 1 < npass) issue_fill(recA, tA.hd);
        else if (lt + 1 < n_local) issue_fill(recB, tB.hd);
        PROF_ADD(4);
      }
      recA = recB; tA = tB; recB = recC; tB = tC;
    }
    tc::cp_async_wait_all();
    if (threadIdx.x == W_SMX0 * 32) PROF_STORE(3);
  } else if (warp == W_SX) {
    // ------------------------------------------------------------------ SX issuer: D_X += E_j . cand_j^T per block (SS form, both
    //        operands K-major exactly as gathered); it touches no accumulator of the S1 / S2 chain
    uint32_t g = 0, u = 0;
    TileInfo nx;
    if (n_local > 0) fetch_info(args, tile0, nx);
    for (int lt = 0; lt < n_local; ++lt) {
      const int64_t cs = nx.cs, ce = nx.ce;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) fetch_info(args, tile0 + (lt + 1) * tstep, nx);
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        const uint32_t idescx = tc::make_idesc_bf16_f32(TM, nc16);
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t s = g % S1, s2 = g % S2;
          role_wait<16>(&bars->full1[s], (g / S1) & 1);
          role_wait<16>(&bars->full2[s2], (g / S2) & 1);
          if (j == 0) tc::mbar_wait(&bars->x_free, (u & 1) ^ 1);                // the previous unit's X is out of D_X
          tc::tcgen05_fence_after();
          const uint64_t e_desc = tc::make_smem_desc_sw128(tc::smem_u32(st1 + s * ST1_BYTES));
          const uint64_t c_desc = tc::make_smem_desc_sw128(tc::smem_u32(st2 + s2 * C_BYTES));
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < ((ABL & 8) ? 0 : FB / 16); ++ks) tc::umma_bf16(tmem + DX_COL, e_desc + 2 * ks, c_desc + 2 * ks, idescx, (j | ks) != 0 ? 1u : 0u);
            tc::umma_commit(&bars->empty1[s]);
            tc::umma_commit(&bars->empty2[s2]);
            if (j == KB - 1) tc::umma_commit(&bars->x_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ score warps, per finished unit (model.py:127-136,213-214):
    //        the softmax weights back from A_w (bf16 hi + lo) and X = D_X into shared memory, m[k,c] = sum_slots w[k,s] X[s,c] on
    //        the CUDA cores (4 codes x 4 candidates per thread), then the attention logits out of D_a, the softmax over K and the
    //        weighted sum
    const int q = warp & 3;
    const int li_q = (q * 32) / LPI;                           // impression of this quarter's TMEM lanes
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int st = (warp - W_SCR0) * 32 + lane;                // 0..127
    const int tl = q * 32 + lane;                              // TMEM lane: (i, k, hl) in A_w / D_a, history slot in D_X
    const int lk = ((tl % LPI) >> 4) * 8 + (tl & 7);
    const bool lo_part = (tl & 8) != 0;
    const bool row_ok = lk < K;
    const int c4 = lane & 3;
    // contraction / final stage: thread = (impression, 4 codes kg, 4 candidates cg); a warp holds one impression
    const int my_i = st >> 6, kg = st & 7, cg = (st >> 3) & 7;
    uint32_t u = 0;
    // next tile: header, candidate range of the tile, boundary between its two impressions
    TileInfo nx;
    int64_t nx_mid = 0;
    PROF_DECL;
    auto fetch_next = [&](int tile) {
      fetch_info(args, tile, nx);
      const int64_t i1 = static_cast<int64_t>(tile) * IPT + 1;
      nx_mid = cand_off(args, i1 < args.B ? i1 : args.B);
    };
    if (n_local > 0) fetch_next(tile0);
    for (int lt = 0; lt < n_local; ++lt) {
      const int64_t cs = nx.cs, ce = nx.ce, mid = nx_mid;
      const uint2 hd = nx.hd;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) fetch_next(tile0 + (lt + 1) * tstep);              // a tile ahead: off the critical path
      const int n_tot = hdr_end(hd, IPT - 1);
      const int n16 = (n_tot + 15) & ~15;
      const int e0 = hdr_end(hd, 0), b1 = hdr_start(hd, 1), e1 = hdr_end(hd, 1);
      const int my_imp_s = tl < e0 ? 0 : ((tl >= b1 && tl < e1) ? 1 : -1);    // impression of this thread's history slot
      const int ws0 = li_q == 0 ? 0 : b1, ws1 = li_q == 0 ? e0 : e1;          // slots of this quarter's impression (weights copy)
      const int ms0 = my_i == 0 ? 0 : b1, ms1 = my_i == 0 ? e0 : e1;          // slots of this thread's impression (contraction)
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        // pass columns of the two impressions: [0, cut) and [cut, nc)
        const int64_t m0 = mid - pc0;
        const int cut = static_cast<int>(m0 < 0 ? 0 : (m0 > nc ? nc : m0));
        const int nmax = cut > nc - cut ? cut : nc - cut;
        const int rounds = (nmax + CC - 1) / CC;
        const uint32_t db = u & 1;                             // A_w buffer of the unit
        const uint32_t dab = NDA == 2 ? (u & 1) : 0;           // its D_a buffer
        PROF_ADD(0);
        role_wait<4>(&bars->x_full, u & 1);                    // every SX of the unit is complete
        PROF_ADD(1);
        tc::tcgen05_fence_after();
        tc::named_bar_sync(1, T_SCR);                          // the previous unit's score threads are done with W / Sm / Sa
        PROF_ADD(2);
        auto copy_weights = [&]() {
          tc::mbar_wait(&bars->w_ready[db], (u >> 1) & 1);      // (long complete: S1 of the unit waited for it)
          tc::tcgen05_fence_after();
          if (!(ABL & 1)) {
            // softmax weights of this quarter's impression: A_w (lanes (i, k, hl), two slots per column) -> W[slot][k] fp32
            const int n_lo = ws0 >> 3, n_hi = (ws1 + 7) >> 3;
#pragma unroll 1
            for (int hp = 0; hp < 4; ++hp) {
              const int hf = hp >> 1, part = hp & 1;
              if (8 * part < n_hi && 8 * part + 8 > n_lo) {      // warp-uniform
                const int g0 = q * 32 + hf * 16;
                const int k = ((g0 % LPI) / 16) * 8 + (lane >> 2);
                uint32_t wr[16];
                tmem_ld_16x128b_x8(tmem + (static_cast<uint32_t>(g0) << 16) + AW_COL + 64 * db + 32 * part, wr);
                tc::tmem_ld_wait();
                float* wp = W + (64 * part + 2 * c4) * LS + k;
#pragma unroll
                for (int nn = 0; nn < 8; ++nn) {
                  const int sa = 8 * (8 * part + nn) + 2 * c4;
                  const uint32_t hi = wr[2 * nn], lo = wr[2 * nn + 1];
                  if (sa >= ws0 && sa < ws1) wp[8 * nn * LS] = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
                  if (sa + 1 >= ws0 && sa + 1 < ws1) wp[(8 * nn + 1) * LS] = __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u);
                }
              }
            }
          }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->w_free[db]);                    // A_w[u & 1] may be rewritten (unit u + 2)
        PROF_ADD(7);
        };
        for (int r = 0; r < rounds; ++r) {
          if (r > 0) tc::named_bar_sync(1, T_SCR);             // the previous round's reads of Xs are done
          // X rows of this thread's slot: the CC columns of its own impression's range that belong to this round
#pragma unroll 1
          for (int i = 0; i < IPT; ++i) {
            const int sb = i == 0 ? 0 : b1, se = i == 0 ? e0 : e1;
            const int cbase = (i == 0 ? 0 : cut) + r * CC;
            const int ncols = (i == 0 ? cut : nc) - cbase;
            if (ncols > 0 && se > q * 32 && sb < q * 32 + 32) {              // warp-uniform: slots of impression i live in this quarter
              uint32_t t8[CC / 8][8];
#pragma unroll
              for (int h = 0; h < CC / 8; ++h) tmem_ld_32x8(tmem + lane_addr + DX_COL + cbase + 8 * h, t8[h]);
              tc::tmem_ld_wait();
              if (my_imp_s == i) {                             // (columns past the impression's range hold other candidates' values:
                uint4* dst = reinterpret_cast<uint4*>(Xs + tl * XS);   //  nothing reads them)
#pragma unroll
                for (int c = 0; c < CC / 4; ++c) dst[c] = make_uint4(t8[c >> 1][(4 * c) & 7], t8[c >> 1][(4 * c + 1) & 7], t8[c >> 1][(4 * c + 2) & 7], t8[c >> 1][(4 * c + 3) & 7]);
              }
            }
          }
          if (r == rounds - 1) {                               // D_X is out of tensor memory: the next unit's SX may start
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&bars->x_free);
          }
          PROF_ADD(3);
          if (r == 0) copy_weights();
          tc::named_bar_sync(1, T_SCR);
          PROF_ADD(2);
          // m[4 kg .. + 4][4 cg .. + 4] of this thread's impression over its slots
          const int lo_c = (my_i == 0 ? 0 : cut) + r * CC;
          const int ncols = (my_i == 0 ? cut : nc) - lo_c < CC ? (my_i == 0 ? cut : nc) - lo_c : CC;
          if (4 * cg < ncols && !(ABL & 1)) {
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
            const float* wp = W + 4 * kg;
            const float* xp = Xs + 4 * cg;
#pragma unroll 4
            for (int s = ms0; s < ms1; ++s) {
              const float4 w4 = *reinterpret_cast<const float4*>(wp + s * LS);
              const float4 x4 = *reinterpret_cast<const float4*>(xp + s * XS);
              acc[0][0] = fmaf(x4.x, w4.x, acc[0][0]); acc[0][1] = fmaf(x4.x, w4.y, acc[0][1]); acc[0][2] = fmaf(x4.x, w4.z, acc[0][2]); acc[0][3] = fmaf(x4.x, w4.w, acc[0][3]);
              acc[1][0] = fmaf(x4.y, w4.x, acc[1][0]); acc[1][1] = fmaf(x4.y, w4.y, acc[1][1]); acc[1][2] = fmaf(x4.y, w4.z, acc[1][2]); acc[1][3] = fmaf(x4.y, w4.w, acc[1][3]);
              acc[2][0] = fmaf(x4.z, w4.x, acc[2][0]); acc[2][1] = fmaf(x4.z, w4.y, acc[2][1]); acc[2][2] = fmaf(x4.z, w4.z, acc[2][2]); acc[2][3] = fmaf(x4.z, w4.w, acc[2][3]);
              acc[3][0] = fmaf(x4.w, w4.x, acc[3][0]); acc[3][1] = fmaf(x4.w, w4.y, acc[3][1]); acc[3][2] = fmaf(x4.w, w4.z, acc[3][2]); acc[3][3] = fmaf(x4.w, w4.w, acc[3][3]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)                        // candidate lo_c + 4 cg + a, codes 4 kg .. 4 kg + 3
              if (4 * cg + a < ncols)
                *reinterpret_cast<float4*>(Sm + (lo_c + 4 * cg + a) * SS + 4 * kg) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
          }
          PROF_ADD(8);
        }
        if (rounds == 0) {                                     // a unit without candidates
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->x_free);
          copy_weights();
        }
        // attention logits a[c, k] out of D_a (lanes (i, k, hl), columns = candidates), transposed through shared memory
        const int c_lo = li_q == 0 ? 0 : cut, c_hi = li_q == 0 ? cut : nc;
        PROF_ADD(0);
        role_wait<4>(&bars->dma_full[dab], (NDA == 2 ? u >> 1 : u) & 1);
        PROF_ADD(5);
        tc::tcgen05_fence_after();
        tc::named_bar_sync(1, T_SCR);                          // Xs is dead, Sa (same bytes) may be written; Sm is complete
        bool released = false;
        for (int c0 = c_lo & ~15; c0 < c_hi; c0 += 16) {
          uint32_t va[16];
          tc::tmem_ld_32x16(tmem + lane_addr + DA_COL + NCM * dab + c0, va);
          tc::tmem_ld_wait();
          if (c0 + 16 >= c_hi) {                               // last chunk in registers: this D_a buffer is free again
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&bars->dma_free[dab]);
            released = true;
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float xa = __uint_as_float(va[c]);
            const float sa_ = xa + __shfl_xor_sync(0xffffffffu, xa, 8);       // G_hi . cand + G_lo . cand
            const int col = c0 + c;
            if (!lo_part && row_ok && col >= c_lo && col < c_hi) Sa[col * SS + lk] = sa_;
          }
        }
        if (!released) {
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->dma_free[dab]);
        }
        PROF_ADD(10);
        tc::named_bar_sync(1, T_SCR);
        PROF_ADD(11);
        {
          // per candidate: softmax over K of the attention logits, weighted sum of the matching scores (model.py:213-214), or
          // max / mean (model.py:128-131); the eight threads kg = 0..7 of a candidate group hold four codes each
          const int lo_c = my_i == 0 ? 0 : cut, n_i = my_i == 0 ? cut : nc - cut;
          const int nrounds_i = (n_i + CC - 1) / CC;           // warp-uniform (a warp holds one impression)
          for (int r = 0; r < ((ABL & 1) ? 0 : nrounds_i); ++r) {
            const int cb = r * CC + 4 * cg;
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {
              const bool ok = cb + a < n_i;
              const int col = lo_c + (ok ? cb + a : 0);
              const float4 mv = *reinterpret_cast<const float4*>(Sm + col * SS + 4 * kg);
              const float4 av = *reinterpret_cast<const float4*>(Sa + col * SS + 4 * kg);
              const float m4[4] = {mv.x, mv.y, mv.z, mv.w}, a4[4] = {av.x, av.y, av.z, av.w};
              float score;
              if (args.score_type == MINER_SCORE_WEIGHTED) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) mx = fmaxf(mx, 4 * kg + j < K ? a4[j] : -INFINITY);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
                float den = 0.f, num = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float e = 4 * kg + j < K ? __expf(a4[j] - mx) : 0.f;
                  den += e;
                  num = fmaf(e, 4 * kg + j < K ? m4[j] : 0.f, num);
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) { den += __shfl_xor_sync(0xffffffffu, den, o); num += __shfl_xor_sync(0xffffffffu, num, o); }
                score = num / den;
              } else if (args.score_type == MINER_SCORE_MAX) {
                score = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) score = fmaxf(score, 4 * kg + j < K ? m4[j] : -INFINITY);
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) score = fmaxf(score, __shfl_xor_sync(0xffffffffu, score, o));
              } else {
                score = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) score += 4 * kg + j < K ? m4[j] : 0.f;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) score += __shfl_xor_sync(0xffffffffu, score, o);
                score /= static_cast<float>(K);
              }
              if (ok && kg == 0) args.out_scores[pc0 + col] = score;
            }
          }
        }
        PROF_ADD(6);
      }
    }
    if (threadIdx.x == W_SCR0 * 32) PROF_STORE(4);
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool tscore_x_supported(int ipt, int km, int nh, int64_t H, bool want_interests) {
  return ipt == IPT && km == KM && nh == 1 && IPT * ((H + 1) & ~1ll) <= SL && !want_interests;
}

int launch_tscore_x(const ts::TScoreArgs& args, int64_t n_tiles, cudaStream_t stream) {
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  MINER_CUDA_OK(cudaFuncSetAttribute(tscore_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  tscore_x_kernel<<<grid, T_THREADS, SMEM, stream>>>(args, static_cast<int>(n_tiles));
  MINER_LAUNCH_OK("tscore_x_kernel");
  return MINER_OK;
}

}  // namespace miner
