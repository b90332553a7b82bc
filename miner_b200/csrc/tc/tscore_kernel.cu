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
// Two launches per call:
//   tpack_kernel   (one warp per tile) turns his_ids / his_mask into PACKED TILES in the caller's workspace.  A tile holds IPT
//       impressions; each impression's history becomes a compact list of slot records.  Every masked slot carries the SAME logit
//       (the reference overwrites it with 1e-30 for every code, model.py:180), so masked slots that point at the same news row
//       -- the left padding with the pad news, reader.py:368-369 -- have identical softmax weights and identical rows:
//       sum_pads w E_pad = (n_pad w) E_pad.  They are merged into ONE record with multiplicity n_pad (its logit becomes
//       1e-30 + ln n_pad, which is the same softmax term), so a 12-click history costs 13 gathered rows, not 50.  Same
//       mathematics as the reference to fp32 rounding; nothing is dropped (masked slots with other ids stay separate records).
//       The impressions of a tile are laid side by side (even start slots), so a tile is  sum(n_i) <= 128 NH  slots.
//       Ids outside [0, n_rows) are counted into the workspace's oob counter (rows read as zero, as before).
//   tscore_kernel  one persistent CTA per SM.  TMEM lanes are (impression i, context code k, part hl): lane
//       LPI i + 16 (k / 8) + 8 hl + k % 8 with LPI = 128 / IPT, where hl selects the bf16 hi / lo part of the softmax weight --
//       the two lanes of a pair accumulate  w_hi . E  and  w_lo . E  and are summed in the epilogue (fp32-level weights).
//       Keeping a pair 8 lanes apart lets the 16-lane tcgen05.ld / st shapes (16x256b / 16x128b) hand both rows of a pair to ONE
//       thread: the pair sum, the gelu and the hi/lo splits need no shuffles and no thread repeats another's work.
//   warps 0-3   gather: per 64-feature block, table[id] and tw[id] rows of the tile's slots (only ceil(n / 16) 16-row groups),
//               16-byte cp.async straight into the 128B-swizzled layout, completion through mbarriers;  warps 4-5: the
//               tile's candidate rows, own ring
//   warp 6      issues every tcgen05.mma:
//               S1(j): D_I = A_w . E_j, D_P = A_w . TW_j   (A_w = softmax weights in tensor memory, TS form, block "staircase"
//                      over the tile's impressions; B = the gathered tile read MN-major; ceil(n / 16) K-steps)
//               S2(j): D_m += A_I . cand_j^T, D_a += A_G . cand_j^T  (A = bf16 hi|lo of I and of gelu(P), written IN PLACE over
//                      the fp32 accumulators by the epilogue warps; B = candidate rows, K-major)
//   warps 7-14  epilogue: per block, pair sum, gelu, bf16 hi/lo splits, tcgen05.st back in place (packed f32x2 arithmetic)
//   warps 15-18 softmax over the history from the lg rows (L2-resident 128-byte rows, one slot per lane), weights to tensor memory
//   warps 19-22 per finished tile, the softmax over K and the weighted sum of the matching scores through a shared-memory
//               transpose, 1 / 2 / 4 threads per candidate
// Tiles with more than 96 candidates (64 when NH = 2) run several passes (the history side is recomputed; rare).
#include "tscore_common.cuh"

namespace miner {

long long* hist_prof_buffer();

namespace {

using namespace ts;

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
  static constexpr int LS = KM + 4;          // logits scratch row stride (floats): the softmax threads read column-wise (4 slots x 8 codes per
                                             // request) and  8 c + k  then covers the 32 banks once
  static constexpr int SS = KM + 2;          // score scratch row stride (floats): 1, 2 or 4 threads per candidate read it with at most 2-way conflicts
  static constexpr int L_FLOATS = TM * NH * LS, S_FLOATS = 2 * NC_MAX * SS;
  static constexpr int SCRATCH_FLOATS = L_FLOATS + S_FLOATS;                    // the softmax warps' logits and the score warps' transposes are live together
  static constexpr int S1 = S1_MAX - 1 - (KM > 32 ? 1 : 0) - (NH > 1 ? 1 : 0);  // wider scratch takes ring stages
  static constexpr int SMEM = 1024 + S1 * ST1_BYTES + S2 * C_BYTES + SCRATCH_FLOATS * 4 + 512;
};
constexpr int T_EPI = 256, T_SMX = 128, T_SCR = 128;   // 8 epilogue warps, 4 softmax warps, 4 score warps
constexpr int T_G1 = 128, T_G2 = 64;                   // gather threads of the (E, TW) ring / of the candidate ring
constexpr int G1_ROWS = TM * 8 / T_G1, G2_ROWS = NC_MAX * 8 / T_G2;  // rows per thread (a thread moves one 16-byte chunk per row)
constexpr int G1_STEP = T_G1 / 8, G2_STEP = T_G2 / 8;
static_assert(G1_STEP == 16 && G1_ROWS == 8, "a gather thread's row jj is 16-row group jj: one K-step of S1");
constexpr int W_G2 = T_G1 / 32, W_MMA = W_G2 + T_G2 / 32, W_EPI0 = W_MMA + 1, W_SMX0 = W_EPI0 + T_EPI / 32, W_SCR0 = W_SMX0 + T_SMX / 32;
constexpr int T_THREADS = (W_SCR0 + T_SCR / 32) * 32;
// TMEM map (512 columns)
// (NH = 128-slot halves of a tile: 1, or 2 when IPT histories can exceed 128 slots)
constexpr int AW_COL = 0;                    // softmax weights, packed bf16: 128 NH slots -> 64 NH columns
// then IP_COL = 64 NH: 2 buffers x (I 64 | P 64) fp32, their first 32 columns become the packed A operands;
// DM_COL = IP_COL + 256: matching scores m[(i,k,hl), c];  DA_COL = DM_COL + NCM: attention logits;  NCM = 96 (NH = 1) or 64

struct TBarriers {
  uint64_t full1[S1_MAX], empty1[S1_MAX], full2[S2], empty2[S2];
  uint64_t w_ready, w_free, ip_full[2], a_ready[2], dma_full, dma_free;
  uint32_t tmem_base;
};


// -------------------------------------------------------------------------------------------------------------------- tpack_kernel
__device__ __forceinline__ uint2 make_rec(int64_t id, bool masked, int64_t n_rows, int mult, int h, int il) {
  const bool valid = id >= 0 && id < n_rows;
  uint2 r;
  r.x = (valid ? static_cast<uint32_t>(id) : 0u) | (masked ? REC_MASKED : 0u) | (valid ? REC_VALID : 0u);
  r.y = static_cast<uint32_t>(mult) | (static_cast<uint32_t>(h) << 16) | (static_cast<uint32_t>(il) << 24);
  return r;
}

// One warp per tile.  CH = ceil(H / 32) chunks of 32 slots per impression.
template <int CH>
__global__ void __launch_bounds__(256)
tpack_kernel(const void* __restrict__ his_ids, int id_dtype, const uint8_t* __restrict__ mask, int64_t B, int H, int64_t n_rows, int ipt, int ts,
             int64_t n_tiles, uint2* __restrict__ rec, uint2* __restrict__ hdr, int* __restrict__ oob) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  int bad = 0;
  for (int64_t tile = w0; tile < n_tiles; tile += nw) {
    uint2* out = rec + tile * ts;
    int pos = 0;
    uint32_t ends[4] = {0u, 0u, 0u, 0u};
    for (int i = 0; i < 4; ++i) {
      if (i < ipt) {
        const int64_t imp = tile * ipt + i;
        pos = (pos + 1) & ~1;
        if (imp < B) {
          int64_t id[CH];
          bool in[CH], msk[CH];
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int h = c * 32 + lane;
            in[c] = h < H;
            id[c] = in[c] ? load_id(his_ids, imp * H + h, id_dtype) : -1;
            msk[c] = in[c] && mask[imp * H + h] == 0;
          }
          // the first masked slot names the row every other masked slot with the same id is merged into
          int64_t fm = 0;
          bool have = false;
          int fh = 0;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const unsigned b = __ballot_sync(0xffffffffu, msk[c]);
            const int l = b ? __ffs(b) - 1 : 0;
            const int64_t cand = __shfl_sync(0xffffffffu, id[c], l);
            if (!have && b) { fm = cand; have = true; fh = c * 32 + l; }
          }
          bool mrg[CH];
          int nm = 0;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            mrg[c] = have && msk[c] && id[c] == fm;
            nm += __popc(__ballot_sync(0xffffffffu, mrg[c]));
          }
          if (have) {
            if (lane == 0) {
              out[pos] = make_rec(fm, true, n_rows, nm, fh, i);
              if (fm < 0 || fm >= n_rows) bad += nm;
            }
            pos += 1;
          }
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const bool keep = in[c] && !mrg[c];
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              out[pos + __popc(b & lt_mask)] = make_rec(id[c], msk[c], n_rows, 1, c * 32 + lane, i);
              if (id[c] < 0 || id[c] >= n_rows) ++bad;
            }
            pos += __popc(b);
          }
        }
      }
      ends[i] = static_cast<uint32_t>(pos);
    }
    // padding records (multiplicity 0): the odd slot between two impressions and the tail up to the next 16-slot group
    const int n_tot = pos, n16 = (n_tot + 15) & ~15;
    if (lane < 3 && lane + 1 < ipt && (ends[lane] & 1u)) out[ends[lane]] = make_uint2(0u, 0u);
    if (n_tot + lane < n16) out[n_tot + lane] = make_uint2(0u, 0u);
    if (lane == 0) hdr[tile] = make_uint2(ends[0] | (ends[1] << 16), ends[2] | (ends[3] << 16));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if (lane == 0 && bad > 0 && oob) atomicAdd(oob, bad);
}

// -------------------------------------------------------------------------------------------------------------------- tscore_kernel
template <int IPT, int KM, int NH>
__global__ void __launch_bounds__(T_THREADS, 1)
tscore_kernel(const TScoreArgs args, int n_tiles) {
  constexpr int LPI = TM / IPT;                         // TMEM lanes per impression
  constexpr int TS = TM * NH;                           // slot records per tile
  constexpr int LS = Shape<KM, NH>::LS, SS = Shape<KM, NH>::SS, SCRATCH_FLOATS = Shape<KM, NH>::SCRATCH_FLOATS, S1 = Shape<KM, NH>::S1;
  constexpr int NCM = NH == 1 ? NC_MAX : 64;            // candidate columns per pass
  constexpr int IP_COL = 64 * NH, DM_COL = IP_COL + 256, DA_COL = DM_COL + NCM;
  static_assert(LPI >= 32, "a warp's 32 TMEM lanes belong to one impression");
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET (not by integer arithmetic on the pointer) so that the compiler keeps the shared address space: LDS / STS, not generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* st1 = smem;                                         // [S1][E 16 KB | TW 16 KB]
  uint8_t* st2 = st1 + S1 * ST1_BYTES;                         // [S2][12 KB] candidate rows
  // scratch: the logits L of the unit the softmax warps prepare, and the score transposes Sm / Sa of the unit the score warps finish
  float* L = reinterpret_cast<float*>(st2 + S2 * C_BYTES);     // [128 NH slots][LS] logits, then exp2(logit - max) in place
  float* Sm = L + Shape<KM, NH>::L_FLOATS;                     // [NC_MAX][SS] matching scores, transposed
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
    tc::mbar_init(&bars->dma_free, T_SCR);
    tc::fence_barrier_init();
  }
  if (warp == W_MMA) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  // the CTA owns all 512 columns of its SM's tensor memory, so the allocation starts at lane 0 / column 0: every TMEM address below
  // is then a compile-time constant (the MMA-issuing warp is bound by its own instruction stream, not by the tensor pipe)
  if (bars->tmem_base != 0u) __trap();
  constexpr uint32_t tmem = 0u;

  if (warp < W_G2) {
    // ------------------------------------------------------------------ gathers of the (E, TW) ring: thread = one 16-byte chunk of rows
    //        r0 + 16 jj, i.e. one row of every 16-slot group (K-step) of the tile
    const int t = threadIdx.x;
    const int chunk = t & 7, r0 = t >> 3;
    const uint32_t row_bytes = static_cast<uint32_t>(D) * 2;
    const char* table_b = reinterpret_cast<const char*>(args.table);
    const char* tw_b = reinterpret_cast<const char*>(args.tw);
    const uint32_t dst0 = tc::sw128_offset(r0, chunk);       // row r0 + 16 jj sits 2 jj KB further
    uint32_t rec_pre[G1_ROWS * NH];                          // raw slot records (x word) of the next tile (see RawId)
    uint2 hdr_pre = make_uint2(0u, 0u);
    auto fetch_recs = [&](int lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const uint32_t* rt = reinterpret_cast<const uint32_t*>(args.rec + static_cast<int64_t>(tile) * TS);
#pragma unroll
      for (int jj = 0; jj < G1_ROWS * NH; ++jj) rec_pre[jj] = rt[2 * (r0 + G1_STEP * jj)];
      hdr_pre = args.hdr[tile];
    };
    uint32_t g = 0;
    PROF_DECL;
    int64_t cs = 0, ce = 0;
    if (n_local > 0) { fetch_recs(0); tile_range<IPT>(args, static_cast<int>(blockIdx.x), cs, ce); }
    for (int lt = 0; lt < n_local; ++lt) {
      uint32_t eoff[G1_ROWS * NH];                           // byte offset of this thread's 16-byte chunk in its rows (table < 4 GB, checked by the launcher)
      uint32_t emask = 0;
      const int nks = (hdr_end(hdr_pre, IPT - 1) + 15) >> 4; // 16-slot groups of this tile
#pragma unroll
      for (int jj = 0; jj < G1_ROWS * NH; ++jj) {
        const uint32_t rc = rec_pre[jj];
        eoff[jj] = (rc & REC_ID) * row_bytes + chunk * 16;
        emask |= (rc & REC_VALID) ? (1u << jj) : 0u;
      }
      if (TS_DBG(1)) emask = 0;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) {                                // the next tile's records and candidate range are fetched a tile ahead
        fetch_recs(lt + 1);
        tile_range<IPT>(args, static_cast<int>(blockIdx.x) + (lt + 1) * static_cast<int>(gridDim.x), cs, ce);
      }
      for (int p = 0; p < npass; ++p) {
        for (int j = 0; j < KB; ++j) {
#pragma unroll
          for (int half = 0; half < NH; ++half) {                  // one ring stage per 128-slot half in use
            const int ng = nks - G1_ROWS * half;                   // 16-slot groups of this half
            if (half > 0 && ng <= 0) break;
            const uint32_t s = g % S1, ph = (g / S1) & 1;
            ++g;
            PROF_ADD(0);
            tc::mbar_wait(&bars->empty1[s], ph ^ 1);
            PROF_ADD(1);
            const uint32_t base = tc::smem_u32(st1 + s * ST1_BYTES) + dst0;
            const uint32_t jb = static_cast<uint32_t>(j) * (FB * 2);
#pragma unroll
            for (int jj = 0; jj < G1_ROWS; ++jj) {
              if (jj < ng) {
                const uint32_t o = eoff[half * G1_ROWS + jj] + jb, sz = ((emask >> (half * G1_ROWS + jj)) & 1u) ? 16u : 0u;
                tc::cp_async_16(base + jj * (G1_STEP * 128), table_b + o, sz);
                tc::cp_async_16(base + E_BYTES + jj * (G1_STEP * 128), tw_b + o, TS_DBG(2) ? 0u : sz);
              }
            }
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
    int bad = 0;
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
          const bool in = r0 + G2_STEP * jj < nc, ok = in && id >= 0 && id < args.n_rows;
          coff[jj] = static_cast<uint32_t>(ok ? id : 0) * row_bytes + chunk * 16;
          cmask |= ok ? (1u << jj) : 0u;
          if (in && !ok && chunk == 0) ++bad;
        }
        if (TS_DBG(5)) cmask = 0;
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
    if (bad > 0 && args.oob) atomicAdd(args.oob + 1, bad);   // candidate ids outside the table (their rows read as zero)
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    // S1: ONE MMA per 16-slot K-step computes D_I | D_P = A_w . [E_j | TW_j]: B is two 64-feature swizzle atoms wide along N (the E
    // and TW halves of the stage, E_BYTES apart), read MN-major; the 128 accumulator columns are the I | P halves of the buffer
    const uint32_t idesc1 = tc::make_idesc_bf16_f32_major(TM, 2 * FB, false, true);
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
    uint2 hdr_pre = make_uint2(0u, 0u);
    if (n_local > 0) { tile_range<IPT>(args, static_cast<int>(blockIdx.x), t_cs, t_ce); hdr_pre = args.hdr[blockIdx.x]; }
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int nks = (hdr_end(hdr_pre, IPT - 1) + 15) >> 4;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) {                                  // a tile ahead: off the critical path
        tile_range<IPT>(args, tile + static_cast<int>(gridDim.x), t_cs, t_ce);
        hdr_pre = args.hdr[tile + static_cast<int>(gridDim.x)];
      }
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
          const uint32_t d_i = tmem + IP_COL + b * 128;
#pragma unroll
          for (int half = 0; half < NH; ++half) {
            const int ng = nks - 8 * half < 8 ? nks - 8 * half : 8;
            if (half > 0 && ng <= 0) break;
            const bool last_half = half == NH - 1 || nks <= 8 * (half + 1);
            const uint32_t s = sg % S1, ph = (sg / S1) & 1;
            ++sg;
            PROF_ADD(0);
            tc::mbar_wait(&bars->full1[s], ph);
            PROF_ADD(2);
            tc::tcgen05_fence_after();
            const uint64_t et_desc = tc::make_smem_desc_sw128_mn_wide(tc::smem_u32(st1 + s * ST1_BYTES), E_BYTES);
            if (tc::elect_one()) {
              for (int ks = 0; ks < ng; ++ks)
                tc::umma_bf16_ts(d_i, tmem + AW_COL + half * 64 + 8 * ks, et_desc + ks * (2048 >> 4), idesc1, (half | ks) != 0 ? 1u : 0u);
              tc::umma_commit(&bars->empty1[s]);
              if (last_half) {
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
    const int g0 = q * 32 + half * 16;                         // first lane of this warp's group
    const int li = g0 / LPI;                                   // its impression inside the tile
    // this warp owns lanes [g0, g0 + 16); thread t meets the (hi, lo) rows of code bk
    const uint32_t grp_addr = static_cast<uint32_t>(g0) << 16;
    const int bk = ((g0 % LPI) / 16) * 8 + (lane >> 2);        // context code of this thread's row pair
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
      const int64_t my_imp = static_cast<int64_t>(tile) * IPT + li;
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
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            // w_hi . E + w_lo . E for two adjacent features, then the bf16 hi / lo images of the sum
            const unsigned long long s = f2_add(f2_pack(__uint_as_float(vi[4 * n]), __uint_as_float(vi[4 * n + 1])),
                                                f2_pack(__uint_as_float(vi[4 * n + 2]), __uint_as_float(vi[4 * n + 3])));
            if (want_i) {                                                      // model.py:138 (interests are an output)
              float s0, s1;
              f2_unpack(s, s0, s1);
              *reinterpret_cast<float2*>(args.out_interests + (my_imp * K + bk) * D + j * FB + bf + 8 * n) = make_float2(s0, s1);
            }
            split_hi_lo(s, pk[2 * n], pk[2 * n + 1]);
          }
          tc::tmem_st_16x128b_x8(acc, pk);                                     // in place: this thread group has read all 64 columns
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            // G = gelu(P) as hi + lo as well (model.py:212): a single bf16 G is the largest error of the path once |P| ~ 1
            const unsigned long long s = f2_add(f2_pack(__uint_as_float(vp[4 * n]), __uint_as_float(vp[4 * n + 1])),
                                                f2_pack(__uint_as_float(vp[4 * n + 2]), __uint_as_float(vp[4 * n + 3])));
            split_hi_lo(gelu2(s), pk[2 * n], pk[2 * n + 1]);
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
  } else if (warp < W_SCR0) {
    // ------------------------------------------------------------------ softmax warps: logits of the tile's slots -> softmax over the
    //        history -> A_w in tensor memory
    const int sw = warp - W_SMX0;
    const int q = warp & 3;
    const int c4 = lane & 3;
    uint32_t u = 0;
    PROF_DECL;
    // raw slot records of the next tile (see RawId): lane l of warp sw owns slot 4 l + sw of every 128-slot half
    uint2 rec_pre[NH];
    uint2 hdr_pre = make_uint2(0u, 0u);
    auto fetch_recs = [&](int tile) {
      const uint2* rt = args.rec + static_cast<int64_t>(tile) * TS;
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) rec_pre[hh] = rt[hh * TM + 4 * lane + sw];
      hdr_pre = args.hdr[tile];
    };
    int64_t t_cs = 0, t_ce = 0;
    if (n_local > 0) { tile_range<IPT>(args, static_cast<int>(blockIdx.x), t_cs, t_ce); fetch_recs(static_cast<int>(blockIdx.x)); }
    const bool k_vec4 = (K & 3) == 0;                          // lg rows are then 16-byte aligned
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int npass = passes_of<NCM>(cs, ce);
      uint2 rec_cur[NH];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) rec_cur[hh] = rec_pre[hh];
      const uint2 hd = hdr_pre;
      if (lt + 1 < n_local) {                                  // a tile ahead: off the critical path
        tile_range<IPT>(args, tile + static_cast<int>(gridDim.x), t_cs, t_ce);
        fetch_recs(tile + static_cast<int>(gridDim.x));
      }
      const int n_tot = hdr_end(hd, IPT - 1);
      const int nks = (n_tot + 15) >> 4;
      for (int p = 0; p < npass; ++p, ++u) {
        PROF_ADD(0);
        tc::named_bar_sync(2, T_SMX);                                          // previous unit's reads of L are done
        {
          // logits of the tile's slots, one slot per lane (slots interleaved over the four warps): the lg row of a slot is K
          // consecutive floats (model.py:174 hoisted to the table).  Stored scaled by log2(e): the softmax below runs on ex2.
          constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
            const int slot = hh * TM + 4 * lane + sw;
            if (hh * TM >= n_tot) break;
            if (slot < n_tot) {
              const uint32_t info = rec_cur[hh].x, meta = rec_cur[hh].y;
              const uint32_t mult = meta & 0xffffu;
              float* dst = L + slot * LS;
              if (mult != 0u && !(info & REC_MASKED)) {
                // kept slot: lg row (+ the category bias of the slot, model.py:176-177); an id outside the table reads as a zero row
                const float bias = args.bias_mean ? args.bias_mean[(static_cast<int64_t>(tile) * IPT + ((meta >> 24) & 0xfu)) * H + ((meta >> 16) & 0xffu)] : 0.f;
                const float* row = args.lg + static_cast<size_t>(info & REC_ID) * K;
                if (!(info & REC_VALID)) {
#pragma unroll 1
                  for (int k = 0; k < K; ++k) dst[k] = bias * LOG2E;
                } else if (k_vec4) {
#pragma unroll
                  for (int h8 = 0; h8 < KM / 32; ++h8) {
                    float4 x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                      if (32 * h8 + 4 * i < K) x[i] = reinterpret_cast<const float4*>(row)[8 * h8 + i];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                      if (32 * h8 + 4 * i < K)
                        reinterpret_cast<float4*>(dst)[8 * h8 + i] = make_float4((x[i].x + bias) * LOG2E, (x[i].y + bias) * LOG2E,
                                                                                 (x[i].z + bias) * LOG2E, (x[i].w + bias) * LOG2E);
                  }
                } else {
#pragma unroll 1
                  for (int k = 0; k < K; ++k) dst[k] = (row[k] + bias) * LOG2E;
                }
              } else if (mult != 0u) {
                // masked record of multiplicity n: n slots filled with 1e-30 (model.py:180) add up to n exp(1e-30) = exp(1e-30 + ln n)
                const float x = kMaskFill * LOG2E + __log2f(static_cast<float>(mult));
#pragma unroll 1
                for (int k = 0; k < K; ++k) dst[k] = x;
              }
            }
          }
        }
        tc::named_bar_sync(2, T_SMX);
        PROF_ADD(1);
        // softmax over the history (model.py:181): per 16-lane group, thread t owns code k = 8 group + t/4 and the slots
        // {2c, 2c+1 : c = t%4 + 4n} of its impression's range [s0, s1); the (hi, lo) rows of the pair leave through 16x128b stores.
        // Every column the MMAs of this tile read (8 per 16-slot group) is rewritten, zeros outside the impression's own range:
        // the block "staircase" of A_w moves from tile to tile.
        const int nparts = (nks * 8 + 31) >> 5;                                // 32-column stores per lane group
        if (nparts == 1) {
          // the whole tile is 64 slots or fewer: weights stay in registers across the wait for A_w
          uint32_t o[2][16];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int g0 = q * 32 + hf * 16;
            const int li = g0 / LPI;
            const int k = ((g0 % LPI) / 16) * 8 + (lane >> 2);
            const int s0 = hdr_start(hd, li), s1 = hdr_end(hd, li);
            const bool row_ok = k < K && s1 > s0;
            const float* col = L + (row_ok ? k : 0);
            float e[16];
            float mx = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const int sa = 2 * (c4 + 4 * n);
              e[2 * n] = (row_ok && sa >= s0 && sa < s1) ? col[sa * LS] : -INFINITY;
              e[2 * n + 1] = (row_ok && sa >= s0 && sa + 1 < s1) ? col[(sa + 1) * LS] : -INFINITY;
              mx = fmaxf(mx, fmaxf(e[2 * n], e[2 * n + 1]));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const bool dead = mx == -INFINITY;                                 // impression past the end of the batch / unused row
            float sum = 0.f;
#pragma unroll
            for (int n = 0; n < 16; ++n) {
              e[n] = dead ? 0.f : ex2_approx(e[n] - mx);
              sum += e[n];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = dead ? 0.f : __fdividef(1.0f, sum);
#pragma unroll
            for (int n = 0; n < 8; ++n) split_hi_lo(f2_pack(e[2 * n] * inv, e[2 * n + 1] * inv), o[hf][2 * n], o[hf][2 * n + 1]);   // model.py:181
          }
          PROF_ADD(2);
          if (u > 0) tc::mbar_wait(&bars->w_free, (u - 1) & 1);                // S1 of the previous unit no longer reads A_w
          PROF_ADD(3);
          tc::tcgen05_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) tc::tmem_st_16x128b_x8(tmem + (static_cast<uint32_t>(q * 32 + hf * 16) << 16) + AW_COL, o[hf]);
        } else {
          // more than 64 slots: exp2(logit - max) goes back to L in place, the weights are formed 64 slots at a time after the wait
          float inv[2];
          int s0v[2], s1v[2];
          bool deadv[2];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int g0 = q * 32 + hf * 16;
            const int li = g0 / LPI;
            const int k = ((g0 % LPI) / 16) * 8 + (lane >> 2);
            const int s0 = hdr_start(hd, li), s1 = hdr_end(hd, li);
            const bool row_ok = k < K && s1 > s0;
            float* col = L + (row_ok ? k : 0);
            const int n_lo = s0 >> 3, n_hi = (s1 + 7) >> 3;
            float mx = -INFINITY;
            for (int n = n_lo; n < n_hi; ++n) {
              const int sa = 2 * (c4 + 4 * n);
              const float a = (sa >= s0 && sa < s1) ? col[sa * LS] : -INFINITY;
              const float b = (sa + 1 < s1 && sa >= s0) ? col[(sa + 1) * LS] : -INFINITY;
              mx = fmaxf(mx, fmaxf(a, b));
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const bool dead = mx == -INFINITY || !row_ok;
            float sum = 0.f;
            for (int n = n_lo; n < n_hi; ++n) {
              const int sa = 2 * (c4 + 4 * n);
              if (!dead && sa >= s0 && sa < s1) {
                const float e = ex2_approx(col[sa * LS] - mx);
                col[sa * LS] = e;
                sum += e;
              }
              if (!dead && sa >= s0 && sa + 1 < s1) {
                const float e = ex2_approx(col[(sa + 1) * LS] - mx);
                col[(sa + 1) * LS] = e;
                sum += e;
              }
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            inv[hf] = dead ? 0.f : __fdividef(1.0f, sum);
            s0v[hf] = s0; s1v[hf] = s1; deadv[hf] = dead;
          }
          PROF_ADD(2);
          if (u > 0) tc::mbar_wait(&bars->w_free, (u - 1) & 1);
          PROF_ADD(3);
          tc::tcgen05_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int g0 = q * 32 + hf * 16;
            const int k = ((g0 % LPI) / 16) * 8 + (lane >> 2);
            const float* col = L + (deadv[hf] ? 0 : k);
            const int s0 = s0v[hf], s1 = s1v[hf];
            for (int part = 0; part < nparts; ++part) {
              uint32_t o[16];
#pragma unroll
              for (int n = 0; n < 8; ++n) {
                const int sa = 2 * (c4 + 4 * (8 * part + n));
                const bool in0 = !deadv[hf] && sa >= s0 && sa < s1, in1 = !deadv[hf] && sa >= s0 && sa + 1 < s1;
                const float w0 = in0 ? col[sa * LS] * inv[hf] : 0.f;               // model.py:181
                const float w1 = in1 ? col[(sa + 1) * LS] * inv[hf] : 0.f;
                split_hi_lo(f2_pack(w0, w1), o[2 * n], o[2 * n + 1]);
              }
              tc::tmem_st_16x128b_x8(tmem + (static_cast<uint32_t>(g0) << 16) + AW_COL + part * 32, o);
            }
          }
        }
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->w_ready);
        PROF_ADD(4);
      }
    }
    if (threadIdx.x == W_SMX0 * 32) PROF_STORE(3);
  } else {
    // ------------------------------------------------------------------ score warps: scores of a finished unit (model.py:127-136,213-214).
    //        Own warps: the MMA issuer waits for D_m / D_a to be drained before the next unit's first S2, so the drain starts the
    //        moment the last S2 of the unit completes.
    const int q = warp & 3;
    const int li_q = (q * 32) / LPI;                           // impression of this quarter's lanes (LPI >= 32)
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int st = (warp - W_SCR0) * 32 + lane;                // 0..127
    const int tl = q * 32 + lane;                              // TMEM lane = (i, k, hl) for the 32-lane reads
    const int lk = ((tl % LPI) >> 4) * 8 + (tl & 7);
    const bool lo_part = (tl & 8) != 0;
    const bool row_ok = lk < K;
    uint32_t u = 0;
    int64_t t_cs = 0, t_ce = 0, t_mcs = 0, t_mce = 0;          // next tile: candidate range of the tile and of this quarter's impression
    auto fetch_range = [&](int tile) {
      tile_range<IPT>(args, tile, t_cs, t_ce);
      const int64_t my_imp = static_cast<int64_t>(tile) * IPT + li_q;
      t_mcs = my_imp < args.B ? cand_off(args, my_imp) : -1;
      t_mce = my_imp < args.B ? cand_off(args, my_imp + 1) : -1;
    };
    if (n_local > 0) fetch_range(static_cast<int>(blockIdx.x));
    for (int lt = 0; lt < n_local; ++lt) {
      const int tile = static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x);
      const int64_t cs = t_cs, ce = t_ce;
      const int64_t my_cs = t_mcs < 0 ? ce : t_mcs, my_ce = t_mce < 0 ? ce : t_mce;
      const int npass = passes_of<NCM>(cs, ce);
      if (lt + 1 < n_local) fetch_range(tile + static_cast<int>(gridDim.x));   // a tile ahead: off the critical path
      for (int p = 0; p < npass; ++p, ++u) {
        const int64_t pc0 = cs + static_cast<int64_t>(p) * NCM;
        const int nc = static_cast<int>(ce - pc0 < NCM ? ce - pc0 : NCM);
        // columns of this lane's impression inside the pass (warp-uniform: a warp's 32 lanes belong to one impression)
        const int64_t r_lo = my_cs - pc0, r_hi = my_ce - pc0;
        const int c_lo = static_cast<int>(r_lo < 0 ? 0 : (r_lo > nc ? nc : r_lo));
        int c_hi = static_cast<int>(r_hi < 0 ? 0 : (r_hi > nc ? nc : r_hi));
        if (c_hi < c_lo) c_hi = c_lo;
        tc::mbar_wait(&bars->dma_full, u & 1);
        tc::tcgen05_fence_after();
        tc::named_bar_sync(1, T_SCR);                                          // the previous unit's score threads are done with Sm / Sa
        bool released = false;
        for (int c0 = c_lo & ~15; c0 < c_hi; c0 += 16) {
          uint32_t vm[16], va[16];
          tc::tmem_ld_32x16(tmem + lane_addr + DM_COL + c0, vm);
          tc::tmem_ld_32x16(tmem + lane_addr + DA_COL + c0, va);
          tc::tmem_ld_wait();
          if (c0 + 16 >= c_hi) {                                               // last chunk in registers: D_m / D_a are free again
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&bars->dma_free);
            released = true;
          }
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
        if (!released) {
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->dma_free);
        }
        tc::named_bar_sync(1, T_SCR);
        {
          // one candidate per group of P threads (P = 4, 2 or 1 by the number of candidates): softmax over K of the attention
          // logits, weighted sum of the matching scores (model.py:213-214), or max / mean (model.py:128-131)
          const int P = nc <= 32 ? 4 : (nc <= 64 ? 2 : 1);
          const int c = st / P, part = st % P;
          const int ce_ = c < nc ? c : nc - 1;
          const float* m = Sm + ce_ * SS;
          const float* a = Sa + ce_ * SS;
          float score;
          if (args.score_type == MINER_SCORE_WEIGHTED) {
            float mx = -INFINITY;
            for (int k = part; k < K; k += P) mx = fmaxf(mx, a[k]);
            if (P > 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            if (P > 2) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float den = 0.f, num = 0.f;
            for (int k = part; k < K; k += P) {
              const float e = __expf(a[k] - mx);
              den += e;
              num = fmaf(e, m[k], num);
            }
            if (P > 1) { den += __shfl_xor_sync(0xffffffffu, den, 1); num += __shfl_xor_sync(0xffffffffu, num, 1); }
            if (P > 2) { den += __shfl_xor_sync(0xffffffffu, den, 2); num += __shfl_xor_sync(0xffffffffu, num, 2); }
            score = num / den;
          } else if (args.score_type == MINER_SCORE_MAX) {
            score = -INFINITY;
            for (int k = part; k < K; k += P) score = fmaxf(score, m[k]);
            if (P > 1) score = fmaxf(score, __shfl_xor_sync(0xffffffffu, score, 1));
            if (P > 2) score = fmaxf(score, __shfl_xor_sync(0xffffffffu, score, 2));
          } else {
            score = 0.f;
            for (int k = part; k < K; k += P) score += m[k];
            if (P > 1) score += __shfl_xor_sync(0xffffffffu, score, 1);
            if (P > 2) score += __shfl_xor_sync(0xffffffffu, score, 2);
            score /= static_cast<float>(K);
          }
          if (c < nc && part == 0) args.out_scores[pc0 + c] = score;
        }
      }
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc(tmem, 512);
}

// tile geometry for a shape: impressions per tile and 128-slot halves.  An impression takes 2 K TMEM lanes (32 at least) and
// up to (H + 1) & ~1 slots; four impressions share a tile only when their slots fit ONE half (two halves leave 64 candidate
// columns per pass, too few for four impressions' candidates), two when they fit two halves.
struct TileGeom { int ipt, nh, km; };
bool tile_geom(int64_t H, int64_t K, TileGeom* g) {
  if (H < 1 || H > 2 * TM || K < 1 || K > KMAX) return false;
  const int ha = static_cast<int>((H + 1) & ~1ll);
  int ipt = 1;
  if (K <= 16 && 4 * ha <= TM) ipt = 4;
  else if (K <= 32 && 2 * ha <= 2 * TM) ipt = 2;
  g->ipt = ipt;
  g->nh = ipt * ha > TM ? 2 : 1;
  g->km = K > 32 ? 64 : 32;
  return true;
}

}  // namespace

bool tscore_kernel_supported(int64_t H, int64_t K, int64_t D) {
  TileGeom g;
  return tile_geom(H, K, &g) && D >= FB && D % FB == 0 && D <= 8192;
}

// workspace: [oob counters: 2 x int32, padded to 256 B][tile headers: n_tiles x uint2][slot records: n_tiles x 128 NH x uint2]
size_t tscore_ws_bytes(int64_t B, int64_t H, int64_t K) {
  TileGeom g;
  if (B <= 0 || !tile_geom(H, K, &g)) return 256;
  const size_t n_tiles = static_cast<size_t>((B + g.ipt - 1) / g.ipt);
  return 256 + align_up(n_tiles * sizeof(uint2), 256) + n_tiles * static_cast<size_t>(TM * g.nh) * sizeof(uint2);
}

int launch_tscore_kernel(const void* table, const void* tw, const float* lg, int64_t n_rows, const void* his_ids, int id_dtype,
                         const uint8_t* his_mask, const float* bias_mean, const void* cand_ids, const int64_t* cand_offsets,
                         int64_t B, int64_t H, int64_t C, int64_t K, int64_t D, int score_type, float* out_scores, float* out_interests,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  TileGeom geo;
  if (!tscore_kernel_supported(H, K, D) || !tile_geom(H, K, &geo)) {
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
  if (!workspace || workspace_bytes < tscore_ws_bytes(B, H, K)) {
    set_error("table-level scoring: workspace too small (%zu bytes needed, %zu given)", tscore_ws_bytes(B, H, K), workspace_bytes);
    return MINER_ERR_WORKSPACE;
  }
  const int64_t n_tiles = (B + geo.ipt - 1) / geo.ipt;
  if (n_tiles > 0x7fffffffll) {
    set_error("table-level scoring: %lld tiles in one call (split the batch)", (long long)n_tiles);
    return MINER_ERR_UNSUPPORTED;
  }
  char* ws = static_cast<char*>(workspace);
  int* oob = reinterpret_cast<int*>(ws);
  uint2* hdr = reinterpret_cast<uint2*>(ws + 256);
  uint2* rec = reinterpret_cast<uint2*>(ws + 256 + align_up(static_cast<size_t>(n_tiles) * sizeof(uint2), 256));
  const int ts = TM * geo.nh;
  {
    // step 0: packed tiles (one warp per tile)
    const int64_t blocks = (n_tiles + 7) / 8;
    const int grid = static_cast<int>(blocks < 16ll * sm_count() ? blocks : 16ll * sm_count());
    const int ch = static_cast<int>((H + 31) / 32);
#define MINER_TP_LAUNCH(CHV)                                                                                                                \
  tpack_kernel<CHV><<<grid, 256, 0, stream>>>(his_ids, id_dtype, his_mask, B, static_cast<int>(H), n_rows, geo.ipt, ts, n_tiles, rec, hdr, oob)
    if (ch <= 1) MINER_TP_LAUNCH(1);
    else if (ch <= 2) MINER_TP_LAUNCH(2);
    else if (ch <= 4) MINER_TP_LAUNCH(4);
    else MINER_TP_LAUNCH(8);
#undef MINER_TP_LAUNCH
    MINER_LAUNCH_OK("tpack_kernel");
  }
  TScoreArgs a;
  a.table = static_cast<const uint16_t*>(table); a.tw = static_cast<const uint16_t*>(tw); a.lg = lg; a.n_rows = n_rows;
  a.rec = rec; a.hdr = hdr; a.oob = oob;
  a.cand_ids = cand_ids; a.id_dtype = id_dtype; a.bias_mean = bias_mean;
  a.cand_offsets = cand_offsets; a.B = B; a.H = static_cast<int>(H); a.K = static_cast<int>(K); a.D = static_cast<int>(D);
  a.C = static_cast<int>(C); a.score_type = score_type; a.out_scores = out_scores; a.out_interests = out_interests;
  a.prof = nullptr;
  a.dbg = 0;
#ifdef MINER_TS_PROF
  a.prof = hist_prof_buffer();
  {
    static const char* env_dbg = getenv("MINER_TS_DBG");
    a.dbg = env_dbg ? atoi(env_dbg) : 0;
  }
#endif
#ifndef MINER_TS_NO_X
  // headline shapes (two impressions per tile, K <= 32, H <= 56, scores only): the X-formulation kernel
  if (tscore_x_supported(geo.ipt, geo.km, geo.nh, H, out_interests != nullptr)) return launch_tscore_x(a, n_tiles, stream);
#endif
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
#define MINER_TS_LAUNCH(I, KMV, NHV)                                                                                                    \
  do {                                                                                                                                  \
    MINER_CUDA_OK(cudaFuncSetAttribute(tscore_kernel<I, KMV, NHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Shape<KMV, NHV>::SMEM)); \
    tscore_kernel<I, KMV, NHV><<<grid, T_THREADS, Shape<KMV, NHV>::SMEM, stream>>>(a, static_cast<int>(n_tiles));                      \
  } while (0)
  const int key = geo.ipt * 100 + geo.km + geo.nh;
  switch (key) {
    case 4 * 100 + 32 + 1: MINER_TS_LAUNCH(4, 32, 1); break;
    case 2 * 100 + 32 + 1: MINER_TS_LAUNCH(2, 32, 1); break;
    case 2 * 100 + 32 + 2: MINER_TS_LAUNCH(2, 32, 2); break;
    case 1 * 100 + 32 + 1: MINER_TS_LAUNCH(1, 32, 1); break;
    case 1 * 100 + 32 + 2: MINER_TS_LAUNCH(1, 32, 2); break;
    case 1 * 100 + 64 + 1: MINER_TS_LAUNCH(1, 64, 1); break;
    case 1 * 100 + 64 + 2: MINER_TS_LAUNCH(1, 64, 2); break;
    default:
      set_error("table-level scoring: no kernel variant for ipt=%d km=%d nh=%d", geo.ipt, geo.km, geo.nh);
      return MINER_ERR_UNSUPPORTED;
  }
#undef MINER_TS_LAUNCH
  MINER_LAUNCH_OK("tscore_kernel");
  return MINER_OK;
}

// what a caller needs to know about the tiling of a shape (tests, bench): impressions per tile, 128-slot halves
int tscore_tile_geometry(int64_t H, int64_t K, int* ipt, int* nh) {
  TileGeom g;
  if (!tile_geom(H, K, &g)) return 0;
  if (ipt) *ipt = g.ipt;
  if (nh) *nh = g.nh;
  return 1;
}

}  // namespace miner
