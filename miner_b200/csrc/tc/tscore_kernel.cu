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
// One persistent CTA per SM; a tile is IPT = 2 impressions = 128 history slots (64 per impression, H <= 64).  TMEM lanes are
// (impression i, context code k, part hl): lane 64 i + 2 k + hl, where hl selects the bf16 hi / lo part of the softmax
// weight -- the two lanes of a pair accumulate  w_hi . E  and  w_lo . E  and are summed in the epilogue (fp32-level weights).
//   warps 0-3   gather: per 64-feature block, table[his] and tw[his] rows (128 slots x 128 B each) and the tile's candidate
//               rows, 16-byte cp.async straight into the 128B-swizzled layout, completion through mbarriers
//   warp 4      issues every tcgen05.mma:
//               S1(j): D_I = A_w . E_j, D_P = A_w . TW_j   (A_w = softmax weights in tensor memory, TS form, block diagonal over
//                      the two impressions; B = the gathered 128 x 64 tile read MN-major)
//               S2(j): D_m += A_I . cand_j^T, D_a += A_G . cand_j^T  (A = bf16 hi|lo of I / gelu(P), written IN PLACE over the
//                      fp32 accumulator by the epilogue warps; B = candidate rows, K-major)
//   warps 5-12  epilogue: per block, pair-sum (shuffle), gelu, bf16 hi/lo split, tcgen05.st; per tile, the softmax over K and
//               the weighted sum of the matching scores through a shared-memory transpose, one thread per candidate
//   warps 13-16 softmax over the history from the lg rows (L2-resident 128-byte rows), weights to tensor memory
// Tiles with more than 96 candidates run several passes (the history side is recomputed; rare).
#include <cuda.h>
#include <stdlib.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {

namespace {

constexpr int TM = 128;                      // TMEM lanes = history slots per tile
constexpr int FB = 64;                       // feature block (128 bytes of bf16)
constexpr int IPT = 2, HP = TM / IPT, LPI = TM / IPT;
constexpr int KMAX = 32;
constexpr int S1 = 4, S2 = 4;                // ring depths: (E, TW) blocks / candidate blocks
constexpr int E_BYTES = TM * FB * 2;         // 16 KB
constexpr int ST1_BYTES = 2 * E_BYTES;
constexpr int NC_MAX = 96;                   // candidate columns per pass
constexpr int C_BYTES = NC_MAX * FB * 2;     // 12 KB
constexpr int LS = KMAX;                     // logits scratch row stride (floats)
constexpr int SS = KMAX + 1;                 // score scratch row stride (floats)
constexpr int T_THREADS = 17 * 32;
constexpr int T_EPI = 256, T_SMX = 128, T_GAT = 128;
// TMEM map (512 columns)
constexpr int AW_COL = 0;                    // softmax weights, packed bf16: 128 slots -> 64 columns
constexpr int IP_COL = 64;                   // 2 buffers x (I 64 | P 64) fp32; their first 32 columns become the packed A operands
constexpr int DM_COL = IP_COL + 2 * 128;     // matching scores  m[(i,k,hl), c]
constexpr int DA_COL = DM_COL + NC_MAX;      // attention logits a[(i,k,hl), c]

struct TBarriers {
  uint64_t full1[S1], empty1[S1], full2[S2], empty2[S2];
  uint64_t w_ready, w_free, ip_full[2], a_ready[2], dma_full, dma_free;
  uint32_t tmem_base;
};

struct TScoreArgs {
  const uint16_t* table; const uint16_t* tw; const float* lg; int64_t n_rows;
  const void* his_ids; const void* cand_ids; int id_dtype;
  const uint8_t* mask; const float* bias_mean; const int64_t* cand_offsets;
  int64_t B;
  int H, K, D, C, score_type;
  float* out_scores; float* out_interests;
};

__device__ __forceinline__ int64_t cand_off(const TScoreArgs& a, int64_t i) { return a.cand_offsets ? a.cand_offsets[i] : i * a.C; }

// candidate range of a tile and its number of passes
__device__ __forceinline__ void tile_cands(const TScoreArgs& a, int tile, int64_t& cs, int64_t& ce, int& npass) {
  const int64_t i0 = static_cast<int64_t>(tile) * IPT;
  const int64_t i1 = i0 + IPT < a.B ? i0 + IPT : a.B;
  cs = cand_off(a, i0);
  ce = cand_off(a, i1);
  const int64_t n = ce - cs;
  npass = n <= NC_MAX ? 1 : static_cast<int>((n + NC_MAX - 1) / NC_MAX);
}

__device__ __forceinline__ float gelu_fast(float x) {               // tanh form, hardware tanh (see cand_kernel.cu)
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  const float hx = 0.5f * x;
  return fmaf(hx, tc::tanh_approx(u), hx);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// bf16 hi (even lanes) or lo = x - hi (odd lanes) of a pair of values
__device__ __forceinline__ uint32_t pack_part(float x0, float x1, bool lo) {
  const uint32_t hi = pack2(x0, x1);
  if (!lo) return hi;
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  return pack2(x0 - h0, x1 - h1);
}

__global__ void __launch_bounds__(T_THREADS, 1)
tscore_kernel(const TScoreArgs args, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* st1 = smem;                                         // [S1][E 16 KB | TW 16 KB]
  uint8_t* st2 = st1 + S1 * ST1_BYTES;                         // [S2][12 KB] candidate rows
  float* L = reinterpret_cast<float*>(st2 + S2 * C_BYTES);     // [128 slots][LS] logits
  float* Sm = L + TM * LS;                                     // [NC_MAX][SS] matching scores, transposed
  float* Sa = Sm + NC_MAX * SS;                                // [NC_MAX][SS] attention logits, transposed
  TBarriers* bars = reinterpret_cast<TBarriers*>(Sa + NC_MAX * SS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, K = args.K, D = args.D;
  const int KB = D / FB;
  const int n_local = (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { tc::mbar_init(&bars->full1[s], T_GAT); tc::mbar_init(&bars->empty1[s], 1); }
    for (int s = 0; s < S2; ++s) { tc::mbar_init(&bars->full2[s], T_GAT); tc::mbar_init(&bars->empty2[s], 1); }
    tc::mbar_init(&bars->w_ready, T_SMX);
    tc::mbar_init(&bars->w_free, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bars->ip_full[b], 1); tc::mbar_init(&bars->a_ready[b], T_EPI); }
    tc::mbar_init(&bars->dma_full, 1);
    tc::mbar_init(&bars->dma_free, T_EPI);
    tc::fence_barrier_init();
  }
  if (warp == 4) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ gathers
    const int t = threadIdx.x;
    const int chunk = t & 7, r0 = t >> 3;                    // 16-byte chunk of the 128-byte block row; rows r0 + 16 jj
    uint32_t dst_off[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) dst_off[jj] = tc::sw128_offset(r0 + 16 * jj, chunk);
    int32_t ids_pre[8];
    auto fetch_ids = [&](int lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int r = r0 + 16 * jj;
        const int64_t imp = static_cast<int64_t>(tile) * IPT + r / HP;
        const int h = r % HP;
        const bool ok = h < H && imp < args.B;
        const int64_t id = load_id(args.his_ids, ok ? imp * H + h : 0, args.id_dtype);
        ids_pre[jj] = (ok && id >= 0 && id < args.n_rows) ? static_cast<int32_t>(id) : -1;
      }
    };
    uint32_t g = 0;
    if (n_local > 0) fetch_ids(0);
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      int64_t eoff[8];
      uint32_t emask = 0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const bool ok = ids_pre[jj] >= 0;
        eoff[jj] = static_cast<int64_t>(ok ? ids_pre[jj] : 0) * D + chunk * 8;
        emask |= ok ? (1u << jj) : 0u;
      }
      if (lt + 1 < n_local) fetch_ids(lt + 1);
      int64_t cs, ce;
      int npass;
      tile_cands(args, tile, cs, ce, npass);
      for (int p = 0; p < npass; ++p) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NC_MAX;
        const int nc = static_cast<int>(ce - pc0 < NC_MAX ? ce - pc0 : NC_MAX);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        int64_t coff[6];
        uint32_t cmask = 0;
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
          const int c = r0 + 16 * jj;
          int64_t id = -1;
          if (c < nc) id = load_id(args.cand_ids, pc0 + c, args.id_dtype);
          const bool ok = id >= 0 && id < args.n_rows;
          coff[jj] = (ok ? id : 0) * D + chunk * 8;
          cmask |= ok ? (1u << jj) : 0u;
        }
        for (int j = 0; j < KB; ++j, ++g) {
          {
            const uint32_t s = g % S1, ph = (g / S1) & 1;
            tc::mbar_wait_relaxed(&bars->empty1[s], ph ^ 1);
            const uint32_t base = tc::smem_u32(st1 + s * ST1_BYTES);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              tc::cp_async_16(base + dst_off[jj], args.table + eoff[jj] + j * FB, ((emask >> jj) & 1u) ? 16u : 0u);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              tc::cp_async_16(base + E_BYTES + dst_off[jj], args.tw + eoff[jj] + j * FB, ((emask >> jj) & 1u) ? 16u : 0u);
            tc::cp_async_mbar_arrive_noinc(&bars->full1[s]);
          }
          {
            const uint32_t s = g % S2, ph = (g / S2) & 1;
            tc::mbar_wait_relaxed(&bars->empty2[s], ph ^ 1);
            const uint32_t base = tc::smem_u32(st2 + s * C_BYTES);
#pragma unroll
            for (int jj = 0; jj < 6; ++jj)
              if (r0 + 16 * jj < nc16)
                tc::cp_async_16(base + dst_off[jj], args.table + coff[jj] + j * FB, ((cmask >> jj) & 1u) ? 16u : 0u);
            tc::cp_async_mbar_arrive_noinc(&bars->full2[s]);
          }
        }
      }
    }
    tc::cp_async_wait_all();
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc1 = tc::make_idesc_bf16_f32_major(TM, FB, false, true);        // B = gathered tile, MN-major
    uint32_t g1 = 0, g2 = 0, u = 0;
    bool pending = false;
    int pend_j = 0, pend_nc16 = 16;
    uint32_t pend_u = 0;
    auto stage2 = [&]() {                                                                // S2 of block g2
      const uint32_t b = g2 & 1;
      tc::mbar_wait(&bars->a_ready[b], (g2 >> 1) & 1);
      const uint32_t s = g2 % S2, ph = (g2 / S2) & 1;
      tc::mbar_wait(&bars->full2[s], ph);
      if (pend_j == 0) tc::mbar_wait(&bars->dma_free, (pend_u & 1) ^ 1);                 // previous unit's scores are out of D_m / D_a
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
    };
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      int64_t cs, ce;
      int npass;
      tile_cands(args, tile, cs, ce, npass);
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NC_MAX;
        const int nc = static_cast<int>(ce - pc0 < NC_MAX ? ce - pc0 : NC_MAX);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        tc::mbar_wait(&bars->w_ready, u & 1);
        tc::tcgen05_fence_after();
        for (int j = 0; j < KB; ++j) {
          const uint32_t s = g1 % S1, ph = (g1 / S1) & 1, b = g1 & 1;
          tc::mbar_wait(&bars->full1[s], ph);
          tc::tcgen05_fence_after();
          const uint64_t e_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st1 + s * ST1_BYTES));
          const uint64_t t_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st1 + s * ST1_BYTES + E_BYTES));
          const uint32_t d_i = tmem + IP_COL + b * 128, d_p = d_i + 64;
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < TM / 16; ++ks) {
              tc::umma_bf16_ts(d_i, tmem + AW_COL + 8 * ks, e_desc + ks * (2048 >> 4), idesc1, ks != 0 ? 1u : 0u);
              tc::umma_bf16_ts(d_p, tmem + AW_COL + 8 * ks, t_desc + ks * (2048 >> 4), idesc1, ks != 0 ? 1u : 0u);
            }
            tc::umma_commit(&bars->empty1[s]);
            tc::umma_commit(&bars->ip_full[b]);
            if (j == KB - 1) tc::umma_commit(&bars->w_free);
          }
          __syncwarp();
          ++g1;
          if (pending) stage2();
          pending = true; pend_j = j; pend_nc16 = nc16; pend_u = u;
        }
      }
    }
    if (pending) stage2();
  } else if (warp < 13) {
    // ------------------------------------------------------------------ epilogue warps 5..12
    const int ew = warp - 5;
    const int q = warp & 3;                                    // TMEM lane quarter
    const int half = ew >> 2;                                  // 0: interests / matching scores, 1: gelu(P) / attention logits
    const int et = ew * 32 + lane;
    const int tl = q * 32 + lane;                              // TMEM lane = (i, k, hl)
    const int li = tl / LPI, lk = (tl % LPI) >> 1;
    const bool lo_part = (tl & 1) != 0;
    const bool row_ok = lk < K;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t g = 0, u = 0;
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      int64_t cs, ce;
      int npass;
      tile_cands(args, tile, cs, ce, npass);
      const int64_t i0 = static_cast<int64_t>(tile) * IPT;
      const int64_t my_imp = i0 + li;
      // candidate range of this lane's impression
      const int64_t my_cs = my_imp < args.B ? cand_off(args, my_imp) : ce;
      const int64_t my_ce = my_imp < args.B ? cand_off(args, my_imp + 1) : ce;
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NC_MAX;
        const int nc = static_cast<int>(ce - pc0 < NC_MAX ? ce - pc0 : NC_MAX);
        for (int j = 0; j < KB; ++j, ++g) {
          const uint32_t b = g & 1;
          tc::mbar_wait(&bars->ip_full[b], (g >> 1) & 1);
          tc::tcgen05_fence_after();
          const uint32_t acc = tmem + lane_addr + IP_COL + b * 128 + half * 64;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t v[32];
            tc::tmem_ld_32x32(acc + cc * 32, v);
            tc::tmem_ld_wait();
            float s[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float x = __uint_as_float(v[c]);
              s[c] = x + __shfl_xor_sync(0xffffffffu, x, 1);                  // w_hi . E + w_lo . E
            }
            if (half == 0) {
              if (args.out_interests && p == 0 && !lo_part && row_ok && my_imp < args.B) {     // model.py:138 (interests are an output)
                float4* o = reinterpret_cast<float4*>(args.out_interests + (my_imp * K + lk) * D + j * FB + cc * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = make_float4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) s[c] = gelu_fast(s[c]);                              // model.py:212
            }
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = pack_part(s[2 * c], s[2 * c + 1], lo_part);
            tc::tmem_st_32x16(acc + cc * 16, pk);                             // in place: these columns have been read
          }
          tc::tmem_st_wait();
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->a_ready[b]);
        }
        // ---- scores of this pass (model.py:127-136,213-214)
        tc::mbar_wait(&bars->dma_full, u & 1);
        tc::tcgen05_fence_after();
        tc::named_bar_sync(1, T_EPI);                                          // the previous pass's score threads are done with Sm / Sa
        {
          // columns of this lane's impression inside the pass (warp-uniform: a warp's 32 lanes belong to one impression)
          const int64_t r_lo = my_cs - pc0, r_hi = my_ce - pc0;
          const int c_lo = static_cast<int>(r_lo < 0 ? 0 : (r_lo > nc ? nc : r_lo));
          int c_hi = static_cast<int>(r_hi < 0 ? 0 : (r_hi > nc ? nc : r_hi));
          if (c_hi < c_lo) c_hi = c_lo;
          float* S = half == 0 ? Sm : Sa;
          const uint32_t dcol = tmem + lane_addr + (half == 0 ? DM_COL : DA_COL);
          for (int c0 = c_lo & ~15; c0 < c_hi; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld_32x16(dcol + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float x = __uint_as_float(v[c]);
              const float sum = x + __shfl_xor_sync(0xffffffffu, x, 1);       // A_hi . cand + A_lo . cand
              const int col = c0 + c;
              if (!lo_part && row_ok && col >= c_lo && col < c_hi) S[col * SS + lk] = sum;
            }
          }
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->dma_free);
        tc::named_bar_sync(1, T_EPI);
        if (et < nc) {
          const float* m = Sm + et * SS;
          const float* a = Sa + et * SS;
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
          args.out_scores[pc0 + et] = score;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps 13..16
    const int sw = warp - 13;
    const int q = warp & 3;
    const int tl = q * 32 + lane;
    const int li = tl / LPI, lk = (tl % LPI) >> 1;
    const bool lo_part = (tl & 1) != 0;
    const bool row_ok = lk < K;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t u = 0;
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      int64_t cs, ce;
      int npass;
      tile_cands(args, tile, cs, ce, npass);
      for (int p = 0; p < npass; ++p, ++u) {
        tc::named_bar_sync(2, T_SMX);                                          // previous unit's reads of L are done
        {
          // logits of 32 slots per warp: lg rows are K consecutive floats (model.py:174 hoisted to the table)
          const int slot = sw * 32 + lane;
          const int64_t imp = static_cast<int64_t>(tile) * IPT + slot / HP;
          const int h = slot % HP;
          const bool valid = h < H && imp < args.B;
          int64_t id = -1;
          bool keep = false;
          float bias = 0.f;
          if (valid) {
            id = load_id(args.his_ids, imp * H + h, args.id_dtype);
            keep = args.mask[imp * H + h] != 0;
            if (args.bias_mean) bias = args.bias_mean[imp * H + h];
            if (id < 0 || id >= args.n_rows) id = -1;
          }
          int code = valid ? (keep ? 2 : 1) : 0;
#pragma unroll 4
          for (int ss = 0; ss < 32; ++ss) {
            const long long id_s = __shfl_sync(0xffffffffu, static_cast<long long>(id), ss);
            const int code_s = __shfl_sync(0xffffffffu, code, ss);
            const float bias_s = __shfl_sync(0xffffffffu, bias, ss);
            float v = -INFINITY;                                               // tile padding: not part of the history
            if (code_s == 1) v = kMaskFill;                                    // model.py:180 (1e-30, not -inf)
            if (code_s == 2) v = ((id_s >= 0 && lane < K) ? args.lg[id_s * K + lane] : 0.f) + bias_s;   // model.py:174-177
            L[(sw * 32 + ss) * LS + lane] = v;
          }
        }
        tc::named_bar_sync(2, T_SMX);
        uint32_t pk[32];                                                       // this lane's 64 slots, packed bf16 pairs (hi or lo part)
        {
          const float* col = L + (li * HP) * LS + (row_ok ? lk : 0);
          float mx = -INFINITY;
          for (int h = 0; h < HP; ++h) mx = fmaxf(mx, col[h * LS]);
          const bool dead = mx == -INFINITY || !row_ok;
          float sum = 0.f;
          for (int h = 0; h < HP; ++h) sum += dead ? 0.f : __expf(col[h * LS] - mx);
          const float inv = dead ? 0.f : 1.0f / sum;
#pragma unroll
          for (int c = 0; c < HP / 2; ++c) {
            const float w0 = dead ? 0.f : __expf(col[(2 * c) * LS] - mx) * inv;                  // model.py:181
            const float w1 = dead ? 0.f : __expf(col[(2 * c + 1) * LS] - mx) * inv;
            pk[c] = pack_part(w0, w1, lo_part);
          }
        }
        if (u > 0) tc::mbar_wait(&bars->w_free, (u - 1) & 1);                  // S1 of the previous unit no longer reads A_w
        tc::tcgen05_fence_after();
        {
          uint32_t z[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) z[c] = 0u;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {                                     // 16 columns = 32 slots per store
            const bool mine = (cc >> 1) == li;                                 // warp-uniform
            if (mine) {
              uint32_t o[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) o[c] = (cc & 1) ? pk[16 + c] : pk[c];
              tc::tmem_st_32x16(tmem + lane_addr + AW_COL + cc * 16, o);
            } else {
              tc::tmem_st_32x16(tmem + lane_addr + AW_COL + cc * 16, z);
            }
          }
        }
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->w_ready);
      }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

constexpr int T_SMEM = 1024 + S1 * ST1_BYTES + S2 * C_BYTES + TM * LS * 4 + 2 * NC_MAX * SS * 4 + 256;

}  // namespace

bool tscore_kernel_supported(int64_t H, int64_t K, int64_t D) {
  return H >= 1 && H <= HP && K >= 1 && K <= KMAX && D >= FB && D % FB == 0 && D <= 8192;
}

int launch_tscore_kernel(const void* table, const void* tw, const float* lg, int64_t n_rows, const void* his_ids, int id_dtype,
                         const uint8_t* his_mask, const float* bias_mean, const void* cand_ids, const int64_t* cand_offsets,
                         int64_t B, int64_t H, int64_t C, int64_t K, int64_t D, int score_type, float* out_scores, float* out_interests,
                         cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  if (!tscore_kernel_supported(H, K, D)) {
    set_error("table-level scoring: unsupported shape H=%lld K=%lld D=%lld (need H <= 64, K <= 32, D %% 64 == 0)", (long long)H, (long long)K,
              (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  TScoreArgs a;
  a.table = static_cast<const uint16_t*>(table); a.tw = static_cast<const uint16_t*>(tw); a.lg = lg; a.n_rows = n_rows;
  a.his_ids = his_ids; a.cand_ids = cand_ids; a.id_dtype = id_dtype; a.mask = his_mask; a.bias_mean = bias_mean;
  a.cand_offsets = cand_offsets; a.B = B; a.H = static_cast<int>(H); a.K = static_cast<int>(K); a.D = static_cast<int>(D);
  a.C = static_cast<int>(C); a.score_type = score_type; a.out_scores = out_scores; a.out_interests = out_interests;
  const int64_t n_tiles = (B + IPT - 1) / IPT;
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  MINER_CUDA_OK(cudaFuncSetAttribute(tscore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
  tscore_kernel<<<grid, T_THREADS, T_SMEM, stream>>>(a, static_cast<int>(n_tiles));
  MINER_LAUNCH_OK("tscore_kernel");
  return MINER_OK;
}

}  // namespace miner
