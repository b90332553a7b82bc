// History side of the fused tensor-core scoring path (sm_100a: tcgen05 + TMEM + TMA): PolyAttention.forward
// (reference src/model/model.py:159-185) straight from the embedding table, one persistent CTA per SM.
//
//   E        = table[his_ids]                       gathered k-block by k-block with cp.async (never materialised in HBM)
//   proj     = tanh(E Wp^T)                         tcgen05.mma, bf16 operands, fp32 accumulators in tensor memory     :171
//   logits   = proj codes^T (+ bias_mean), masked slots := 1e-30 (NOT -inf)                                              :174-180
//   w[k,:]   = softmax over the history                                                                                  :181
//   I[k,:]   = sum_h w[k,h] E[h,:]                  tcgen05.mma again: A = [w_hi | w_lo] (bf16 split of the fp32 softmax
//                                                   weights, K-major), B = the SAME gathered E k-block read MN-major      :182
//
// A tile is 128 history rows = 2 impressions of up to 64 clicks (or 1 impression of up to 128).  Per tile:
//   pass 1  12 k-blocks (D = 768): gather warps fill the E stage, TMA fills the Wp stage, the MMA warp accumulates the
//           128 x Dc projection in TMEM;
//   epi 1   8 epilogue warps (two per TMEM lane quarter, splitting the Dc columns): tanh, dot with the context codes on the
//           CUDA cores (fp32), bias / mask fill, then the softmax over the history through a shared-memory transpose; the
//           weights are written as bf16 hi/lo K-major tiles for the second MMA;
//   pass 2  the gather warps stream the same E k-blocks again (L2 hits) and the MMA warp produces I in rounds of three
//           64-feature blocks (2 x 192 TMEM columns, ping-pong) while the epilogue warps drain the previous round: fp32 ->
//           bf16 hi + lo (I = hi + lo to ~2^-17; hi feeds the candidate kernel's projection GEMM, hi + lo its matching
//           scores) and, on request, the fp32 interests themselves.
#include <cuda.h>
#include <stdlib.h>

#include "fused.cuh"
#include "umma.cuh"

namespace miner {

namespace {

constexpr int HM = 128;                 // history rows per tile (UMMA M of pass 1, K of pass 2)
constexpr int HKB = 64;                 // k-block
constexpr int HST = 3;                  // ring depth
constexpr int HA_BYTES = HM * HKB * 2;  // 16 KB
constexpr int KP = 32;                  // context codes are padded to 32 in the logits accumulators
constexpr int LROW = 33;                // padded row of the logits tile
constexpr int H_THREADS = 14 * 32;      // 4 gather warps, TMA warp, MMA warp, 8 epilogue warps
constexpr int H_EPI = 256;
constexpr int RCOLS = 192;              // TMEM columns of one pass-2 round (3 feature blocks)
constexpr int WA_BYTES = 64 * 128;      // one w tile atom: 64 interest rows x 64 history slots (bf16).  The MMA reads M = 128 rows, i.e.
                                        // 8 KB past each atom: those accumulator rows are garbage and never read
constexpr int WT_BYTES = 4 * WA_BYTES;  // w tiles: {hi, lo} x {history rows 0-63, 64-127}

struct HBarriers {
  uint64_t full[HST], empty[HST];
  uint64_t p_full, w_ready, tmem_free;
  uint64_t i_full[2], i_empty[2];
  uint32_t tmem_base;
};

struct HistArgs {
  const uint16_t* table; int64_t n_rows;
  const void* his_ids; int id_dtype;
  const uint8_t* mask; const float* bias_mean;
  const float* codes_t;                 // [Dc][32] fp32, zero padded
  int64_t B;
  int H, K, Dc, D, N1, b_bytes;
  __nv_bfloat16* i_hi; __nv_bfloat16* i_lo; float* out_interests;
};

__device__ __forceinline__ float tanh_acc(float x) {
  const float ax = fabsf(x);
  const float e = __expf(-2.0f * ax);
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void transpose_codes_kernel(const float* __restrict__ codes, float* __restrict__ codes_t, int K, int Dc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Dc * KP) {
    const int dc = i / KP, k = i % KP;
    codes_t[i] = k < K ? codes[static_cast<int64_t>(k) * Dc + dc] : 0.f;
  }
}

__global__ void __launch_bounds__(H_THREADS, 1)
hist_kernel(const __grid_constant__ CUtensorMap tmap_wp, const HistArgs args, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* st_a = smem;                                    // [HST][16 KB]      gathered E k-block
  uint8_t* st_b = st_a + HST * HA_BYTES;                   // [HST][b_bytes]    Wp k-block
  uint8_t* w_t = st_b + HST * args.b_bytes;                // 4 x 16 KB         softmax weights, bf16 hi/lo
  float* codes_s = reinterpret_cast<float*>(w_t + WT_BYTES);   // [Dc][32]      context codes, transposed + zero padded
  float* L = codes_s + args.Dc * KP;                       // [128][33]         logits / softmax scratch
  HBarriers* bars = reinterpret_cast<HBarriers*>(reinterpret_cast<uint8_t*>(L) + HM * LROW * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = args.H, K = args.K, D = args.D, Dc = args.Dc, N1 = args.N1;
  const int KB = D / HKB;
  const int IPT = H <= 64 ? 2 : 1;
  const int HP = HM / IPT;
  const int NR = (KB + 2) / 3;                             // pass-2 rounds per tile

  for (int i = threadIdx.x; i < WT_BYTES / 16; i += H_THREADS) reinterpret_cast<uint4*>(w_t)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < args.Dc * KP / 4; i += H_THREADS)
    reinterpret_cast<float4*>(codes_s)[i] = __ldg(reinterpret_cast<const float4*>(args.codes_t) + i);
  tc::fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < HST; ++s) { tc::mbar_init(&bars->full[s], 128 + 1); tc::mbar_init(&bars->empty[s], 1); }
    tc::mbar_init(&bars->p_full, 1);
    tc::mbar_init(&bars->w_ready, H_EPI);
    tc::mbar_init(&bars->tmem_free, H_EPI);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&bars->i_full[b], 1); tc::mbar_init(&bars->i_empty[b], H_EPI); }
    tc::fence_barrier_init();
  }
  if (warp == 4 && lane == 0) tc::tma_prefetch_desc(&tmap_wp);
  if (warp == 5) { tc::tmem_alloc(&bars->tmem_base, 512); tc::tmem_relinquish(); }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ E gather (both passes)
    const int chunk = lane & 7;
    uint32_t issued = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint16_t* src[8];
      uint32_t nbytes[8], dst_off[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = warp * 32 + j * 4 + (lane >> 3);
        const int64_t imp = static_cast<int64_t>(tile) * IPT + r / HP;
        const int h = r % HP;
        bool ok = h < H && imp < args.B;
        int64_t row = 0;
        if (ok) {
          row = load_id(args.his_ids, imp * H + h, args.id_dtype);
          if (row < 0 || row >= args.n_rows) { ok = false; row = 0; }       // out-of-range id: zero row (gather semantics)
        }
        src[j] = args.table + row * D + chunk * 8;
        nbytes[j] = ok ? 16u : 0u;
        dst_off[j] = tc::sw128_offset(r, chunk);
      }
      for (int t = 0; t < 2 * KB; ++t) {
        const int kb = t < KB ? t : t - KB;
        const uint32_t s = issued % HST, ph = (issued / HST) & 1;
        tc::mbar_wait(&bars->empty[s], ph ^ 1);
        const uint32_t base = tc::smem_u32(st_a + s * HA_BYTES);
#pragma unroll
        for (int j = 0; j < 8; ++j) tc::cp_async_16(base + dst_off[j], src[j] + kb * HKB, nbytes[j]);
        tc::cp_async_mbar_arrive_noinc(&bars->full[s]);     // arrives by itself once this thread's 8 copies have landed
        ++issued;
      }
    }
    tc::cp_async_wait_all();
  } else if (warp == 4) {
    // ------------------------------------------------------------------ Wp k-blocks by TMA (pass 1); plain arrive in pass 2
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int t = 0; t < 2 * KB; ++t, ++it) {
          const uint32_t s = it % HST, ph = (it / HST) & 1;
          tc::mbar_wait(&bars->empty[s], ph ^ 1);
          if (t < KB) {
            tc::mbar_arrive_expect_tx(&bars->full[s], static_cast<uint32_t>(N1 * HKB * 2));
            tc::tma_load_2d(&tmap_wp, &bars->full[s], tc::smem_u32(st_b + s * args.b_bytes), t * HKB, 0);
          } else {
            tc::mbar_arrive(&bars->full[s]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = tc::make_idesc_bf16_f32(HM, N1);
      const uint32_t idesc2 = tc::make_idesc_bf16_f32_major(HM, HKB, false, true);     // B = E k-block read MN-major
      uint32_t it = 0, tile_it = 0, rnd_it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
        tc::mbar_wait(&bars->tmem_free, (tile_it & 1) ^ 1);
        tc::tcgen05_fence_after();
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % HST, ph = (it / HST) & 1;
          tc::mbar_wait(&bars->full[s], ph);
          tc::tcgen05_fence_after();
          const uint64_t a_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_a + s * HA_BYTES));
          const uint64_t b_desc = tc::make_smem_desc_sw128(tc::smem_u32(st_b + s * args.b_bytes));
#pragma unroll
          for (int k = 0; k < HKB / 16; ++k) tc::umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
          tc::umma_commit(&bars->empty[s]);
        }
        tc::umma_commit(&bars->p_full);
        tc::mbar_wait(&bars->w_ready, tile_it & 1);
        tc::tcgen05_fence_after();
        for (int rd = 0; rd < NR; ++rd, ++rnd_it) {
          const uint32_t buf = rnd_it & 1;
          tc::mbar_wait(&bars->i_empty[buf], ((rnd_it >> 1) & 1) ^ 1);
          tc::tcgen05_fence_after();
          const int nd = KB - rd * 3 < 3 ? KB - rd * 3 : 3;
          for (int db = 0; db < nd; ++db, ++it) {
            const uint32_t s = it % HST, ph = (it / HST) & 1;
            tc::mbar_wait(&bars->full[s], ph);
            tc::tcgen05_fence_after();
            const uint64_t e_desc = tc::make_smem_desc_sw128_mn(tc::smem_u32(st_a + s * HA_BYTES));
            const uint32_t d_tmem = tmem + buf * RCOLS + db * HKB;
#pragma unroll
            for (int hl = 0; hl < 2; ++hl) {
#pragma unroll
              for (int ks = 0; ks < HM / 16; ++ks) {
                const uint64_t w_desc = tc::make_smem_desc_sw128(tc::smem_u32(w_t + (hl * 2 + (ks >> 2)) * WA_BYTES)) + 2 * (ks & 3);
                tc::umma_bf16(d_tmem, w_desc, e_desc + ks * (2048 >> 4), idesc2, (hl | ks) != 0 ? 1u : 0u);
              }
            }
            tc::umma_commit(&bars->empty[s]);
          }
          tc::umma_commit(&bars->i_full[buf]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 6..13
    const int ew = warp - 6;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int r = q * 32 + lane;                           // tile row = TMEM lane
    const int et = ew * 32 + lane;                         // 0..255
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int n_cc = (N1 + 31) / 32;
    const int cc_begin = half == 0 ? 0 : (n_cc + 1) / 2, cc_end = half == 0 ? (n_cc + 1) / 2 : n_cc;
    const float4* codes4 = reinterpret_cast<const float4*>(codes_s);
    uint32_t tile_it = 0, rnd_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      // ---- epilogue 1: tanh, logits against the context codes, bias / mask, softmax over the history
      tc::mbar_wait(&bars->p_full, tile_it & 1);
      tc::tcgen05_fence_after();
      float acc[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) acc[k] = 0.f;
      for (int cc = cc_begin; cc < cc_end; ++cc) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem + lane_addr + cc * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int dc = cc * 32 + j;
          if (dc < Dc) {                                                       // warp-uniform
            const float t = tanh_acc(__uint_as_float(v[j]));                   // model.py:171
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) {
              const float4 c = codes4[dc * (KP / 4) + k4];
              acc[4 * k4 + 0] = fmaf(t, c.x, acc[4 * k4 + 0]);
              acc[4 * k4 + 1] = fmaf(t, c.y, acc[4 * k4 + 1]);
              acc[4 * k4 + 2] = fmaf(t, c.z, acc[4 * k4 + 2]);
              acc[4 * k4 + 3] = fmaf(t, c.w, acc[4 * k4 + 3]);
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      if (half == 1) {
#pragma unroll
        for (int k = 0; k < KP; ++k) L[r * LROW + k] = acc[k];
      }
      tc::named_bar_sync(1, H_EPI);
      if (half == 0) {
        const int64_t imp = static_cast<int64_t>(tile) * IPT + r / HP;
        const int h = r % HP;
        const bool valid = h < H && imp < args.B;
        const bool keep = valid && args.mask[imp * H + h] != 0;
        const float bias = (valid && args.bias_mean) ? args.bias_mean[imp * H + h] : 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          float v = acc[k] + L[r * LROW + k] + bias;                           // model.py:174-177
          if (!keep) v = kMaskFill;                                            // model.py:180 (1e-30, not -inf)
          if (!valid) v = -INFINITY;                                           // tile padding: not part of the history
          L[r * LROW + k] = v;
        }
      }
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
        float sum = 0.f;
        for (int h = part; h < HP; h += 4) sum += dead ? 0.f : expf(col[h * LROW] - mx);
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (active) {
          const int R = i * K + k;
          for (int h = part; h < HP; h += 4) {
            const float w = dead ? 0.f : expf(col[h * LROW] - mx) / sum;       // model.py:181
            const __nv_bfloat16 whi = __float2bfloat16_rn(w);
            const __nv_bfloat16 wlo = __float2bfloat16_rn(w - __bfloat162float(whi));
            const int hc = i * HP + h;
            const uint32_t off = (hc >> 6) * WA_BYTES + tc::sw128_offset(R, (hc & 63) >> 3) + (hc & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(w_t + off) = whi;
            *reinterpret_cast<__nv_bfloat16*>(w_t + 2 * WA_BYTES + off) = wlo;
          }
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bars->w_ready);

      // ---- epilogue 2: drain the interests, round by round
      const bool row_ok = r < IPT * K;
      const int64_t imp = static_cast<int64_t>(tile) * IPT + r / K;
      const bool store_ok = row_ok && imp < args.B;
      const int64_t grow = static_cast<int64_t>(tile) * IPT * K + r;            // = imp * K + k
      for (int rd = 0; rd < NR; ++rd, ++rnd_it) {
        const uint32_t buf = rnd_it & 1;
        const int nd = KB - rd * 3 < 3 ? KB - rd * 3 : 3;
        const int ncc2 = nd * 2;
        const int b2 = half == 0 ? 0 : (ncc2 + 1) / 2, e2 = half == 0 ? (ncc2 + 1) / 2 : ncc2;
        tc::mbar_wait(&bars->i_full[buf], (rnd_it >> 1) & 1);
        tc::tcgen05_fence_after();
        if (q * 32 < IPT * K) {                                                // warp-uniform: this lane quarter holds interest rows
          for (int cc = b2; cc < e2; ++cc) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem + lane_addr + buf * RCOLS + cc * 32, v);
            tc::tmem_ld_wait();
            if (store_ok) {
              const int64_t o = grow * D + (rd * 3) * HKB + cc * 32;
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float x0 = __uint_as_float(v[2 * j]), x1 = __uint_as_float(v[2 * j + 1]);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                hi[j] = pack2(x0, x1);
                lo[j] = pack2(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
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
          }
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bars->i_empty[buf]);
      }
      tc::mbar_arrive(&bars->tmem_free);
    }
  }

  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool hist_kernel_supported(int64_t H, int64_t K, int64_t Dc, int64_t D) {
  return H >= 1 && H <= 128 && (K == 8 || K == 16 || K == 32) && Dc >= 1 && Dc <= 256 && D >= 64 && D % 64 == 0 && D <= 4096;
}

size_t hist_kernel_ws_bytes(int64_t Dc) { return align_up(sizeof(float) * static_cast<size_t>(Dc) * KP, 256); }

int launch_hist_kernel(const void* table, int64_t n_rows, const void* his_ids, int id_dtype, const uint8_t* his_mask,
                       const float* bias_mean, const void* w_proj_bf16, const float* codes, int64_t B, int64_t H, int64_t K,
                       int64_t Dc, int64_t D, void* i_hi, void* i_lo, float* out_interests, float* codes_t_ws, cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  if (!hist_kernel_supported(H, K, Dc, D)) {
    set_error("hist_kernel: unsupported shape H=%lld K=%lld Dc=%lld D=%lld", (long long)H, (long long)K, (long long)Dc, (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  static const bool force_v1 = getenv("MINER_HIST_V1") != nullptr;      // A/B switch for profiling
  if (!force_v1 && hist_kernel2_supported(H, K, Dc, D))
    return launch_hist_kernel2(table, n_rows, his_ids, id_dtype, his_mask, bias_mean, w_proj_bf16, codes, B, H, K, Dc, D, i_hi, i_lo,
                               out_interests, stream);
  const int N1 = static_cast<int>((Dc + 15) / 16 * 16);
  CUtensorMap m_wp;
  int rc = make_tmap_2d_bf16(&m_wp, w_proj_bf16, static_cast<uint64_t>(Dc), static_cast<uint64_t>(D), N1, HKB);
  if (rc) return rc;
  transpose_codes_kernel<<<static_cast<unsigned>((Dc * KP + 255) / 256), 256, 0, stream>>>(codes, codes_t_ws, static_cast<int>(K), static_cast<int>(Dc));
  MINER_LAUNCH_OK("transpose_codes");
  HistArgs a;
  a.table = static_cast<const uint16_t*>(table); a.n_rows = n_rows;
  a.his_ids = his_ids; a.id_dtype = id_dtype; a.mask = his_mask; a.bias_mean = bias_mean; a.codes_t = codes_t_ws;
  a.B = B; a.H = static_cast<int>(H); a.K = static_cast<int>(K); a.Dc = static_cast<int>(Dc); a.D = static_cast<int>(D); a.N1 = N1;
  a.b_bytes = N1 * HKB * 2;
  a.i_hi = static_cast<__nv_bfloat16*>(i_hi); a.i_lo = static_cast<__nv_bfloat16*>(i_lo); a.out_interests = out_interests;
  const int ipt = H <= 64 ? 2 : 1;
  const int64_t n_tiles = (B + ipt - 1) / ipt;
  const int grid = static_cast<int>(n_tiles < sm_count() ? n_tiles : sm_count());
  const int smem = 1024 + HST * (HA_BYTES + a.b_bytes) + WT_BYTES + static_cast<int>(Dc) * KP * 4 + HM * LROW * 4 + 256;
  MINER_CUDA_OK(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  hist_kernel<<<grid, H_THREADS, smem, stream>>>(m_wp, a, static_cast<int>(n_tiles));
  MINER_LAUNCH_OK("hist_kernel");
  return MINER_OK;
}

}  // namespace miner
