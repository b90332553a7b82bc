// Candidate side of the fused tensor-core scoring path on CTA PAIRS (sm_100a: tcgen05 cta_group::2 + TMEM + TMA).
// Same contract as cand_kernel.cu (matching scores + TargetAwareAttention, reference src/model/model.py:127,200-216).
//
// Both kernels of the fused path are bound by what one SM can ingest from L2 (about 28 B/cycle); here the weight matrix
// Wt is the largest stream.  Two CTAs of a cluster (two SMs of a TPC) therefore work as a pair: each owns one group of 128
// interest rows, and the projection P = I_hi Wt^T of BOTH groups is one tcgen05.mma.cta_group::2 (M = 256) issued by the
// leader CTA, whose B operand is split across the pair: each SM loads and holds only HALF of every Wt k-block (96 of 192
// rows).  Per k-block an SM ingests 16 KB (its interest tile) + 12 KB instead of 16 + 24 KB.
//   - TMA loads of both CTAs signal the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier
//     address mapped into CTA 0); the leader's tcgen05.commit.cta_group::2 multicasts "stage free" and "chunk done" to both
//     CTAs; the peer forwards "my epilogue has drained the accumulator" with one remote mbarrier arrive per chunk.
//   - everything that depends on a CTA's own candidates stays CTA-local and cta_group::1, exactly as in cand_kernel.cu:
//     the matching-score MMAs m = (I_hi + I_lo) cand^T (their I k-blocks come through a local ring, so they never touch the
//     shared pipeline), the gelu(P) tiles, the attention MMAs a = gelu(P) cand^T, the softmax over K and the scores.
//     (Mixing cta_group::1 MMAs into a cta_group::2 TMEM allocation is verified by scripts/probes/cta_pair_probe.cu.)
//   - the pair runs in lock-step on the shared pipeline: pair iteration gp handles groups 2 gp and 2 gp + 1 with the same
//     number of 128-candidate passes (the larger of the two; the other CTA runs empty passes).
#include <cuda.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {

namespace {

constexpr int CM = 128, CKB = 64, NCH = 192, CST = 3;
constexpr int CA_BYTES = CM * CKB * 2;         // 16 KB  I k-block of this CTA's group
constexpr int BH_ROWS = NCH / 2;               // Wt rows this CTA holds of a k-block
constexpr int BH_BYTES = BH_ROWS * CKB * 2;    // 12 KB
constexpr int CMAXC = 128;
constexpr int CC_BYTES = CMAXC * CKB * 2;
constexpr int RING2 = 2;
constexpr int ML_BYTES = 2 * CA_BYTES;         // matching ring slot: I_hi k-block + I_lo k-block
constexpr int C_THREADS = 14 * 32;
constexpr int C_EPI_THREADS = 256;
constexpr int COL_P = 0, COL_A = NCH, COL_M = NCH + CMAXC;
constexpr int C2_SMEM = 1024 + CST * (CA_BYTES + BH_BYTES) + RING2 * (ML_BYTES + CA_BYTES + CC_BYTES) + 512;

struct C2Barriers {
  uint64_t full[CST], empty[CST];              // full: used in the leader only (both CTAs' TMA loads land there)
  uint64_t lo_full[RING2], lo_empty[RING2];
  uint64_t c_full[RING2], c_empty[RING2];
  uint64_t pb_full[RING2], pb_empty[RING2];
  uint64_t p_full, p_empty, p_empty_peer, fin_full, fin_empty;
  uint32_t tmem_base;
};

struct Cand2Args {
  const uint16_t* table; int64_t n_rows;
  const void* cand_ids; int id_dtype; const int64_t* cand_offsets;
  int64_t B, C;
  int K, D;
  float* out;
};

__device__ __forceinline__ int64_t cand_off2(const Cand2Args& a, int64_t i) {
  if (i > a.B) i = a.B;
  return a.cand_offsets ? a.cand_offsets[i] : i * a.C;
}
__device__ __forceinline__ int pass_count2(int64_t left) { return left <= 0 ? 0 : (left < CMAXC ? static_cast<int>(left) : CMAXC); }

struct PairWork2 { int g; int64_t c0, c1; int passes; };
template <int IPG>
__device__ __forceinline__ PairWork2 pair_work2(const Cand2Args& a, int gp, int rank) {
  PairWork2 w;
  const int64_t b0 = cand_off2(a, static_cast<int64_t>(2 * gp) * IPG), b1 = cand_off2(a, static_cast<int64_t>(2 * gp + 1) * IPG),
                b2 = cand_off2(a, static_cast<int64_t>(2 * gp + 2) * IPG);
  const int64_t n0 = b1 - b0, n1 = b2 - b1;
  const int64_t nmax = n0 > n1 ? n0 : n1;
  w.passes = static_cast<int>((nmax + CMAXC - 1) / CMAXC);
  w.g = 2 * gp + rank;
  w.c0 = rank == 0 ? b0 : b1;
  w.c1 = rank == 0 ? b1 : b2;
  return w;
}

__device__ __forceinline__ float gelu_fast2(float x) {       // tanh-form gelu, see cand_kernel.cu
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  const float hx = 0.5f * x;
  return fmaf(hx, tc::tanh_approx(u), hx);
}
__device__ __forceinline__ uint32_t pack_bf16x2b(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int W>
__device__ __forceinline__ float gmax2(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int W>
__device__ __forceinline__ float gsum2(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- pair (cta_group::2) flavours of the primitives in umma.cuh
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion is counted on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_pair(const void* tmap, uint32_t leader_bar, uint32_t dst_smem, int32_t x, int32_t y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {     // arrives on this barrier in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

template <int K>
__global__ void __launch_bounds__(C_THREADS, 1)
cand_kernel2(const __grid_constant__ CUtensorMap tmap_ihi, const __grid_constant__ CUtensorMap tmap_ilo,
             const __grid_constant__ CUtensorMap tmap_wt, const Cand2Args args, int n_groups) {
  constexpr int IPG = CM / K;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* st_a = smem;                                   // [CST][16 KB]  I_hi k-block (this CTA's 128 rows of the M = 256 MMA)
  uint8_t* st_b = st_a + CST * CA_BYTES;                  // [CST][12 KB]  this CTA's half of the Wt k-block
  uint8_t* ml_t = st_b + CST * BH_BYTES;                  // [2][32 KB]    I_hi | I_lo k-block for the matching MMAs
  uint8_t* pb_t = ml_t + RING2 * ML_BYTES;                // [2][16 KB]    gelu(P) tile
  uint8_t* cd_t = pb_t + RING2 * CA_BYTES;                // [2][16 KB]    candidate feature block
  C2Barriers* bars = reinterpret_cast<C2Barriers*>(cd_t + RING2 * CC_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = args.D;
  const int KB = D / CKB;
  const int n_chunks = (D + NCH - 1) / NCH;
  const int rank = static_cast<int>(tc::cluster_ctarank());
  const bool leader = rank == 0;
  const int cluster_id = static_cast<int>(blockIdx.x) / 2, n_clusters = static_cast<int>(gridDim.x) / 2;
  const int n_pairs = (n_groups + 1) / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CST; ++s) { tc::mbar_init(&bars->full[s], 1); tc::mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < RING2; ++s) {
      tc::mbar_init(&bars->lo_full[s], 1);  tc::mbar_init(&bars->lo_empty[s], 1);
      tc::mbar_init(&bars->c_full[s], 128); tc::mbar_init(&bars->c_empty[s], 1);
      tc::mbar_init(&bars->pb_full[s], C_EPI_THREADS); tc::mbar_init(&bars->pb_empty[s], 1);
    }
    tc::mbar_init(&bars->p_full, 1);   tc::mbar_init(&bars->p_empty, C_EPI_THREADS);
    tc::mbar_init(&bars->p_empty_peer, 1);
    tc::mbar_init(&bars->fin_full, 1); tc::mbar_init(&bars->fin_empty, C_EPI_THREADS);
    tc::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tc::tma_prefetch_desc(&tmap_ihi); tc::tma_prefetch_desc(&tmap_ilo); tc::tma_prefetch_desc(&tmap_wt); }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  tc::cluster_sync_all();                 // barriers of both CTAs are initialised before any remote signal
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ candidate gather (CTA-local, cp.async)
    const int chunk = lane & 7;
    uint32_t issued = 0;
    for (int gp = cluster_id; gp < n_pairs; gp += n_clusters) {
      const PairWork2 pw = pair_work2<IPG>(args, gp, rank);
      for (int ps = 0; ps < pw.passes; ++ps) {
        const int64_t pc0 = pw.c0 + static_cast<int64_t>(ps) * CMAXC;
        const int nc = pass_count2(pw.c1 - pc0);
        const uint16_t* src[8];
        uint32_t nbytes[8], dst_off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = warp * 32 + j * 4 + (lane >> 3);
          bool ok = r < nc;
          int64_t row = 0;
          if (ok) {
            row = load_id(args.cand_ids, pc0 + r, args.id_dtype);
            if (row < 0 || row >= args.n_rows) { ok = false; row = 0; }     // out-of-range id: zero row (gather semantics)
          }
          src[j] = args.table + row * D + chunk * 8;
          nbytes[j] = ok ? 16u : 0u;
          dst_off[j] = tc::sw128_offset(r, chunk);
        }
        for (int n = 0; n < n_chunks; ++n) {
          const int first = n == 0 ? 0 : n * (NCH / CKB);
          const int nn = (D - n * NCH < NCH ? D - n * NCH : NCH) / CKB;
          const int count = n == 0 ? KB + nn : nn;
          for (int t = 0; t < count; ++t) {
            const int fb = (n == 0) ? (t < KB ? t : t - KB) : first + t;
            const uint32_t s = issued % RING2, ph = (issued / RING2) & 1;
            tc::mbar_wait(&bars->c_empty[s], ph ^ 1);
            const uint32_t base = tc::smem_u32(cd_t + s * CC_BYTES);
#pragma unroll
            for (int j = 0; j < 8; ++j) tc::cp_async_16(base + dst_off[j], src[j] + fb * CKB, nbytes[j]);
            tc::cp_async_mbar_arrive_noinc(&bars->c_full[s]);
            ++issued;
          }
        }
      }
    }
    tc::cp_async_wait_all();
  } else if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it = 0, lo_it = 0;
    for (int gp = cluster_id; gp < n_pairs; gp += n_clusters) {
      const PairWork2 pw = pair_work2<IPG>(args, gp, rank);
      const int g = pw.g;
      for (int ps = 0; ps < pw.passes; ++ps) {
        for (int n = 0; n < n_chunks; ++n) {
          const int nn_cols = D - n * NCH < NCH ? D - n * NCH : NCH;
          for (int kb = 0; kb < KB; ++kb, ++it) {
            const uint32_t s = it % CST, ph = (it / CST) & 1;
            tc::mbar_wait(&bars->empty[s], ph ^ 1);              // the leader has released this stage in both CTAs
            if (tc::elect_one()) {
              const uint32_t lbar = map_to_cta(tc::smem_u32(&bars->full[s]), 0);
              if (leader) tc::mbar_arrive_expect_tx(&bars->full[s], 2 * (CA_BYTES + BH_BYTES));
              tma_load_2d_pair(&tmap_ihi, lbar, tc::smem_u32(st_a + s * CA_BYTES), kb * CKB, g * CM);
              tma_load_2d_pair(&tmap_wt, lbar, tc::smem_u32(st_b + s * BH_BYTES), kb * CKB, n * NCH + rank * (nn_cols / 2));
            }
            __syncwarp();
            if (n == 0) {                                        // matching ring: CTA-local
              const uint32_t ls = lo_it % RING2, lph = (lo_it / RING2) & 1;
              tc::mbar_wait(&bars->lo_empty[ls], lph ^ 1);
              if (tc::elect_one()) {
                tc::mbar_arrive_expect_tx(&bars->lo_full[ls], ML_BYTES);
                tc::tma_load_2d(&tmap_ihi, &bars->lo_full[ls], tc::smem_u32(ml_t + ls * ML_BYTES), kb * CKB, g * CM);
                tc::tma_load_2d(&tmap_ilo, &bars->lo_full[ls], tc::smem_u32(ml_t + ls * ML_BYTES + CA_BYTES), kb * CKB, g * CM);
              }
              __syncwarp();
              ++lo_it;
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform; one elected lane issues)
    uint32_t it = 0, lo_it = 0, c_it = 0, pb_it = 0, ch_it = 0, grp_it = 0;
    const uint32_t peer_bar = map_to_cta(tc::smem_u32(&bars->p_empty_peer), 0);
    for (int gp = cluster_id; gp < n_pairs; gp += n_clusters) {
      const PairWork2 pw = pair_work2<IPG>(args, gp, rank);
      for (int ps = 0; ps < pw.passes; ++ps, ++grp_it) {
        const int64_t pc0 = pw.c0 + static_cast<int64_t>(ps) * CMAXC;
        const int nc = pass_count2(pw.c1 - pc0);
        const int nc16 = nc <= 16 ? 16 : (nc + 15) & ~15;
        const uint32_t idesc_c = tc::make_idesc_bf16_f32(CM, nc16);
        tc::mbar_wait(&bars->fin_empty, (grp_it & 1) ^ 1);        // previous pass's scores have been read out of a / m
        tc::tcgen05_fence_after();
        for (int n = 0; n < n_chunks; ++n, ++ch_it) {
          const int nn_cols = D - n * NCH < NCH ? D - n * NCH : NCH;
          const uint32_t idesc_p = tc::make_idesc_bf16_f32(2 * CM, nn_cols);
          if (leader) {
            tc::mbar_wait(&bars->p_empty, (ch_it & 1) ^ 1);        // both epilogues have drained the previous chunk
            tc::mbar_wait(&bars->p_empty_peer, (ch_it & 1) ^ 1);
            tc::tcgen05_fence_after();
          }
          for (int kb = 0; kb < KB; ++kb, ++it) {
            if (n == 0) {
              // matching scores of this CTA's group: m += (I_hi + I_lo) cand^T, all operands CTA-local
              const uint32_t ls = lo_it % RING2, lph = (lo_it / RING2) & 1;
              const uint32_t cs = c_it % RING2, cph = (c_it / RING2) & 1;
              tc::mbar_wait(&bars->lo_full[ls], lph);
              tc::mbar_wait(&bars->c_full[cs], cph);
              tc::tcgen05_fence_after();
              const uint64_t h_desc = tc::make_smem_desc_sw128(tc::smem_u32(ml_t + ls * ML_BYTES));
              const uint64_t l_desc = tc::make_smem_desc_sw128(tc::smem_u32(ml_t + ls * ML_BYTES + CA_BYTES));
              const uint64_t c_desc = tc::make_smem_desc_sw128(tc::smem_u32(cd_t + cs * CC_BYTES));
              if (tc::elect_one()) {
#pragma unroll
                for (int k = 0; k < CKB / 16; ++k) {
                  tc::umma_bf16(tmem + COL_M, h_desc + 2 * k, c_desc + 2 * k, idesc_c, (kb | k) != 0 ? 1u : 0u);
                  tc::umma_bf16(tmem + COL_M, l_desc + 2 * k, c_desc + 2 * k, idesc_c, 1u);
                }
                tc::umma_commit(&bars->lo_empty[ls]);
                tc::umma_commit(&bars->c_empty[cs]);
              }
              __syncwarp();
              ++lo_it; ++c_it;
            }
            if (leader) {
              // projection of BOTH groups: P[256 x nn] += I_hi[256 x 64] Wt_block[nn x 64]^T, B split across the pair
              const uint32_t s = it % CST, ph = (it / CST) & 1;
              tc::mbar_wait(&bars->full[s], ph);
              tc::tcgen05_fence_after();
              const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_a + s * CA_BYTES));
              const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_b + s * BH_BYTES));
              if (tc::elect_one()) {
#pragma unroll
                for (int k = 0; k < CKB / 16; ++k)
                  umma_bf16_pair(tmem + COL_P, a_desc + 2 * k, b_desc + 2 * k, idesc_p, (kb | k) != 0 ? 1u : 0u);
                umma_commit_pair(&bars->empty[s]);
                if (kb == KB - 1) umma_commit_pair(&bars->p_full);
              }
              __syncwarp();
            }
          }
          // attention logits of this CTA's group: a += gelu(P)[:, 64-feature block] . cand[:, same block]^T
          for (int sb = 0; sb < nn_cols / CKB; ++sb, ++pb_it, ++c_it) {
            const uint32_t ps2 = pb_it % RING2, pph = (pb_it / RING2) & 1;
            const uint32_t cs = c_it % RING2, cph = (c_it / RING2) & 1;
            tc::mbar_wait(&bars->pb_full[ps2], pph);
            tc::mbar_wait(&bars->c_full[cs], cph);
            tc::tcgen05_fence_after();
            const uint64_t p_desc = tc::make_smem_desc_sw128(tc::smem_u32(pb_t + ps2 * CA_BYTES));
            const uint64_t c_desc = tc::make_smem_desc_sw128(tc::smem_u32(cd_t + cs * CC_BYTES));
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < CKB / 16; ++k)
                tc::umma_bf16(tmem + COL_A, p_desc + 2 * k, c_desc + 2 * k, idesc_c, (n | sb | k) != 0 ? 1u : 0u);
              tc::umma_commit(&bars->pb_empty[ps2]);
              tc::umma_commit(&bars->c_empty[cs]);
              if (n == n_chunks - 1 && sb == nn_cols / CKB - 1) tc::umma_commit(&bars->fin_full);
            }
            __syncwarp();
          }
          if (!leader) {
            // tell the leader that this CTA's epilogue is done with the projection accumulator of this chunk
            tc::mbar_wait(&bars->p_empty, ch_it & 1);
            if (tc::elect_one()) remote_arrive(peer_bar);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 6..13 (CTA-local)
    const int ew = warp - 6;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t pb_it = 0, ch_it = 0, grp_it = 0;
    for (int gp = cluster_id; gp < n_pairs; gp += n_clusters) {
      const PairWork2 pw = pair_work2<IPG>(args, gp, rank);
      const int g = pw.g;
      const int64_t c1 = pw.c1;
      for (int ps = 0; ps < pw.passes; ++ps, ++grp_it) {
        const int64_t pc0 = pw.c0 + static_cast<int64_t>(ps) * CMAXC;
        for (int n = 0; n < n_chunks; ++n, ++ch_it) {
          const int nn_cols = D - n * NCH < NCH ? D - n * NCH : NCH;
          tc::mbar_wait(&bars->p_full, ch_it & 1);
          tc::tcgen05_fence_after();
          for (int sb = 0; sb < nn_cols / CKB; ++sb, ++pb_it) {
            const uint32_t ps2 = pb_it % RING2, pph = (pb_it / RING2) & 1;
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem + lane_addr + COL_P + sb * CKB + half * 32, v);
            tc::tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = pack_bf16x2b(gelu_fast2(__uint_as_float(v[2 * j])), gelu_fast2(__uint_as_float(v[2 * j + 1])));
            tc::mbar_wait(&bars->pb_empty[ps2], pph ^ 1);
            uint8_t* tile = pb_t + ps2 * CA_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<uint4*>(tile + tc::sw128_offset(r, half * 4 + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&bars->pb_full[ps2]);
          }
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bars->p_empty);
        }
        // ---- scores of this pass: softmax over the K lanes of an impression, weighted sum of matching scores
        tc::mbar_wait(&bars->fin_full, grp_it & 1);
        tc::tcgen05_fence_after();
        if (half == 0) {
          const int64_t imp = static_cast<int64_t>(g) * IPG + r / K;
          const int k = r % K;
          const int64_t o0 = cand_off2(args, imp), o1 = cand_off2(args, imp + 1);
          const int64_t q0 = cand_off2(args, static_cast<int64_t>(g) * IPG + (q * 32) / K);
          const int64_t q1 = cand_off2(args, static_cast<int64_t>(g) * IPG + (q * 32 + 31) / K + 1);
          const int nc = pass_count2(c1 - pc0);
          int lo_col = static_cast<int>(q0 - pc0), hi_col = static_cast<int>(q1 - pc0 < nc ? q1 - pc0 : nc);
          if (q0 - pc0 < 0) lo_col = 0;
          if (q1 <= pc0) hi_col = 0;
          for (int cb = 0; cb < CMAXC / 32; ++cb) {
            if (cb * 32 >= hi_col || cb * 32 + 32 <= lo_col) continue;       // warp-uniform
            uint32_t av[32], mv[32];
            tc::tmem_ld_32x32(tmem + lane_addr + COL_A + cb * 32, av);
            tc::tmem_ld_32x32(tmem + lane_addr + COL_M + cb * 32, mv);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = cb * 32 + j;
              if (col < lo_col || col >= hi_col) continue;                     // warp-uniform
              const float a = __uint_as_float(av[j]);
              const float mx = gmax2<K>(a);
              const float e = expf(a - mx);
              const float se = gsum2<K>(e);
              const float sm = gsum2<K>(e * __uint_as_float(mv[j]));
              const int64_t f = pc0 + col;
              if (k == 0 && f >= o0 && f < o1) args.out[f] = sm / se;
            }
          }
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->fin_empty);
      }
    }
  }

  tc::tcgen05_fence_before();
  tc::cluster_sync_all();                 // the pair's shared memory / tensor memory stay alive until both CTAs are done
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

int launch_cand_kernel2(const void* i_hi, const void* i_lo, const void* wt_bf16, const void* table, int64_t n_rows,
                        const void* cand_ids, int id_dtype, const int64_t* cand_offsets, int64_t B, int64_t C, int64_t K, int64_t D,
                        float* out_scores, cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  CUtensorMap m_hi, m_lo, m_wt;
  int rc = make_tmap_2d_bf16(&m_hi, i_hi, static_cast<uint64_t>(B * K), static_cast<uint64_t>(D), CM, CKB);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&m_lo, i_lo, static_cast<uint64_t>(B * K), static_cast<uint64_t>(D), CM, CKB);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&m_wt, wt_bf16, static_cast<uint64_t>(D), static_cast<uint64_t>(D), BH_ROWS, CKB);
  if (rc) return rc;
  const int ipg = static_cast<int>(CM / K);
  const int64_t n_groups = (B + ipg - 1) / ipg;
  const int64_t n_pairs = (n_groups + 1) / 2;
  const int max_clusters = sm_count() / 2;
  const int grid = 2 * static_cast<int>(n_pairs < max_clusters ? n_pairs : max_clusters);
  Cand2Args a{static_cast<const uint16_t*>(table), n_rows, cand_ids, id_dtype, cand_offsets, B, C, static_cast<int>(K), static_cast<int>(D), out_scores};
  const int n_groups_i = static_cast<int>(n_groups);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C_THREADS);
  cfg.dynamicSmemBytes = C2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define MINER_CAND2(KK)                                                                                             \
  do {                                                                                                              \
    MINER_CUDA_OK(cudaFuncSetAttribute(cand_kernel2<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM));    \
    MINER_CUDA_OK(cudaLaunchKernelEx(&cfg, cand_kernel2<KK>, m_hi, m_lo, m_wt, a, n_groups_i));                     \
  } while (0)
  if (K == 32) MINER_CAND2(32);
  else if (K == 16) MINER_CAND2(16);
  else MINER_CAND2(8);
#undef MINER_CAND2
  MINER_LAUNCH_OK("cand_kernel2");
  return MINER_OK;
}

}  // namespace miner
