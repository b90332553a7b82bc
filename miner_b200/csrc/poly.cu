// (a3) PolyAttention.forward after the projection (reference src/model/model.py:172-182):
//   logits[h,k] = proj[h,:] . codes[k,:]  (+ bias_mean[h])         model.py:174-177
//   logits^T -> masked_fill(~mask, 1e-30)                            model.py:178-180   (NOT -inf: pads keep mass)
//   w[k,:]   = softmax over the history                              model.py:181
//   interests[k,:] = sum_h w[k,h] * E[h,:]                           model.py:182
// One CTA per impression.  proj/codes are staged through shared memory in Dc-chunks and the (H x K) logit tile is
// produced by a register-tiled mini-GEMM; softmax is one warp per context code with shuffle reductions; the weighted
// sum keeps K x (D/256) accumulators per thread and streams the history rows once -- from a dense (B,H,D) tensor or
// straight from the embedding table through his_ids (fp32 or bf16 rows), so the gathered history tile never goes
// to HBM.  Everything is fp32, reference operation order.
#include "common.cuh"

namespace miner {

constexpr int PT = 256;        // threads per CTA
constexpr int DCB = 32;        // Dc chunk staged per iteration
constexpr int TH = 2, TK = 4;  // logits register tile
constexpr int MAXT = 8;        // logits tiles per thread (H*K <= 256*8*8)
constexpr int KG = 32;         // context codes accumulated per pass of the weighted sum

__device__ __forceinline__ float load_emb(const float* emb_row, const void* tab_row, int table_dtype, int64_t d) {
  if (emb_row) return emb_row[d];
  if (table_dtype == MINER_F32) return static_cast<const float*>(tab_row)[d];
  return bf16_bits_to_float(static_cast<const uint16_t*>(tab_row)[d]);
}

template <int DJ>
__global__ void __launch_bounds__(PT) poly_softmax_wsum_kernel(
    const float* __restrict__ proj, const float* __restrict__ codes, const uint8_t* __restrict__ mask,
    const float* __restrict__ bias_mean, const float* __restrict__ emb, const void* __restrict__ table, int table_dtype,
    const void* __restrict__ his_ids, int id_dtype, int64_t n_rows, int H, int K, int Dc, int D,
    float* __restrict__ out_interests, float* __restrict__ out_weights, __nv_bfloat16* __restrict__ out_interests_bf16, int stage_rows) {
  extern __shared__ __align__(16) float smem[];
  const int HP = H + 1;                          // padded row of the logits / weights tile
  const int KP = (K + 3) & ~3;                   // weights^T rows padded to float4
  float* As = smem;                              // [DCB][H + 4]   proj chunk, transposed
  float* Bs = As + DCB * (H + 4);                // [DCB][KP + 4]  codes chunk, transposed
  float* L = Bs + DCB * (KP + 4);                // [K][HP]        logits, then softmax weights
  float* WT = L + ((K * HP + 3) & ~3);           // [H][KP]        weights transposed for the weighted sum
  __shared__ int64_t row_of[256];                // table row of each history slot (H <= 256)

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const float* projb = proj + b * static_cast<int64_t>(H) * Dc;
  if (his_ids && tid < H) {
    const int64_t id = load_id(his_ids, b * H + tid, id_dtype);
    row_of[tid] = (id >= 0 && id < n_rows) ? id : -1;             // out-of-range id: zero row (gather semantics)
  }
  // (staged path, see the weighted sum) the impression's H table rows start their way into shared memory now
  const bool staged = stage_rows != 0;
  uint4* Es = reinterpret_cast<uint4*>(WT + ((H * KP + 3) & ~3));   // [H][D / 8] bf16 rows
  if (staged) {
    __syncthreads();                                               // row_of
    const int NVs = D >> 3;
    for (int i = tid; i < H * NVs; i += PT) {
      const int h = i / NVs, vv = i % NVs;
      const int64_t row = row_of[h];
      const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(table) + (row >= 0 ? row : 0) * (static_cast<int64_t>(D) * 2)) + vv;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(Es + i));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(row >= 0 ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- logits tile: register-tiled (TH x TK) mini-GEMM over Dc chunks ----
  static_assert(TH == 2 && TK == 4, "the logits tile is read as one float2 and one float4");
  const bool h_even = (H & 1) == 0;
  const int tiles_h = (H + TH - 1) / TH, tiles_k = (K + TK - 1) / TK;
  const int n_tiles = tiles_h * tiles_k;
  float acc[MAXT][TH][TK];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
#pragma unroll
    for (int i = 0; i < TH; ++i)
#pragma unroll
      for (int j = 0; j < TK; ++j) acc[t][i][j] = 0.f;

  for (int c0 = 0; c0 < Dc; c0 += DCB) {
    for (int i = tid; i < DCB * H; i += PT) {            // proj chunk: coalesced along dc
      const int h = i / DCB, c = i % DCB;
      As[c * (H + 4) + h] = (c0 + c < Dc) ? projb[static_cast<int64_t>(h) * Dc + c0 + c] : 0.f;
    }
    for (int i = tid; i < DCB * K; i += PT) {
      const int k = i / DCB, c = i % DCB;
      Bs[c * (KP + 4) + k] = (c0 + c < Dc) ? codes[static_cast<int64_t>(k) * Dc + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int tile = tid + t * PT;
      if (tile < n_tiles) {
        const int h0 = (tile / tiles_k) * TH, k0 = (tile % tiles_k) * TK;
#pragma unroll 8
        for (int c = 0; c < DCB; ++c) {
          // (rows are padded: entries past H / K belong to accumulators that are never stored)
          float a[TH], bb[TK];
          if (h_even) {
            const float2 av = *reinterpret_cast<const float2*>(As + c * (H + 4) + h0);
            a[0] = av.x, a[1] = av.y;
          } else {
#pragma unroll
            for (int i = 0; i < TH; ++i) a[i] = (h0 + i < H) ? As[c * (H + 4) + h0 + i] : 0.f;
          }
          const float4 bv = *reinterpret_cast<const float4*>(Bs + c * (KP + 4) + k0);
          bb[0] = bv.x, bb[1] = bv.y, bb[2] = bv.z, bb[3] = bv.w;
#pragma unroll
          for (int i = 0; i < TH; ++i)
#pragma unroll
            for (int j = 0; j < TK; ++j) acc[t][i][j] = fmaf(a[i], bb[j], acc[t][i][j]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    const int tile = tid + t * PT;
    if (tile < n_tiles) {
      const int h0 = (tile / tiles_k) * TH, k0 = (tile % tiles_k) * TK;
#pragma unroll
      for (int i = 0; i < TH; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) {
          const int h = h0 + i, k = k0 + j;
          if (h < H && k < K) {
            float v = acc[t][i][j];
            if (bias_mean) v += bias_mean[b * H + h];                 // model.py:176-177
            if (!mask[b * H + h]) v = kMaskFill;                      // model.py:180
            L[k * HP + h] = v;
          }
        }
    }
  }
  __syncthreads();

  // ---- softmax over the history, one warp per context code (model.py:181) ----
  const int warp = tid >> 5, lane = tid & 31;
  for (int k = warp; k < K; k += PT / 32) {
    float* row = L + k * HP;
    float mx = -INFINITY;
    for (int h = lane; h < H; h += 32) mx = fmaxf(mx, row[h]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int h = lane; h < H; h += 32) {
      const float e = expf(row[h] - mx);
      row[h] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int h = lane; h < H; h += 32) {
      const float w = row[h] / sum;
      WT[h * KP + k] = w;
      if (out_weights) out_weights[(b * K + k) * H + h] = w;
    }
  }
  if (KP != K)
    for (int i = tid; i < H * (KP - K); i += PT) WT[(i / (KP - K)) * KP + K + i % (KP - K)] = 0.f;
  __syncthreads();

  // ---- interests[k,d] = sum_h w[k,h] E[h,d]   (model.py:182) ----
  const float* embb = emb ? emb + b * static_cast<int64_t>(H) * D : nullptr;
  const int64_t row_bytes = static_cast<int64_t>(D) * (table_dtype == MINER_F32 ? 4 : 2);
  // bf16 table rows staged in shared memory (cp.async issued at the top of the kernel, so the gather runs under the logits and the
  // softmax): thread = (8-feature vector v, code slice ks) keeps 8 codes x 8 features; a history row costs one LDS.128 of the row and
  // two of the weights per 64 FMAs.  The sum over the history runs in slot order, as in the generic path below.
  if (staged) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    constexpr int KC = 8;
    const int NV = D >> 3;
    const int KS = PT / NV < 8 ? PT / NV : 8;                    // code slices that work side by side
    const int v = tid % NV, ks = tid / NV;
    if (ks >= KS) return;                                          // (no block-wide barrier below)
    for (int kg = ks * KC; kg < K; kg += KS * KC) {
      float acc[KC][8];
#pragma unroll
      for (int k = 0; k < KC; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
#pragma unroll 2
      for (int h = 0; h < H; ++h) {
        const uint4 raw = Es[static_cast<size_t>(h) * NV + v];
        float e[8];
        e[0] = __uint_as_float(raw.x << 16), e[1] = __uint_as_float(raw.x & 0xffff0000u);
        e[2] = __uint_as_float(raw.y << 16), e[3] = __uint_as_float(raw.y & 0xffff0000u);
        e[4] = __uint_as_float(raw.z << 16), e[5] = __uint_as_float(raw.z & 0xffff0000u);
        e[6] = __uint_as_float(raw.w << 16), e[7] = __uint_as_float(raw.w & 0xffff0000u);
        const float4* wrow = reinterpret_cast<const float4*>(WT + h * KP + kg);
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
          if (kg + k4 * 4 < K) {
            const float4 w = wrow[k4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[k4 * 4 + 0][j] = fmaf(w.x, e[j], acc[k4 * 4 + 0][j]);
              acc[k4 * 4 + 1][j] = fmaf(w.y, e[j], acc[k4 * 4 + 1][j]);
              acc[k4 * 4 + 2][j] = fmaf(w.z, e[j], acc[k4 * 4 + 2][j]);
              acc[k4 * 4 + 3][j] = fmaf(w.w, e[j], acc[k4 * 4 + 3][j]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        if (kg + k < K) {
          const int64_t o = (b * K + kg + k) * static_cast<int64_t>(D) + 8 * v;
          if (out_interests) {
            *reinterpret_cast<float4*>(out_interests + o) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
            *reinterpret_cast<float4*>(out_interests + o + 4) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
          }
          if (out_interests_bf16) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[k][0], acc[k][1]), p1 = __floats2bfloat162_rn(acc[k][2], acc[k][3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[k][4], acc[k][5]), p3 = __floats2bfloat162_rn(acc[k][6], acc[k][7]);
            *reinterpret_cast<uint4*>(out_interests_bf16 + o) = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                                           *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
          }
        }
      }
    }
    return;
  }
  for (int kg = 0; kg < K; kg += KG) {
    float w_acc[KG][DJ];
#pragma unroll
    for (int k = 0; k < KG; ++k)
#pragma unroll
      for (int j = 0; j < DJ; ++j) w_acc[k][j] = 0.f;
    for (int h = 0; h < H; ++h) {
      const float* er = embb ? embb + static_cast<int64_t>(h) * D : nullptr;
      const bool row_ok = embb || row_of[h] >= 0;
      const void* tr = embb ? nullptr : static_cast<const char*>(table) + (row_ok ? row_of[h] : 0) * row_bytes;
      float e[DJ];
#pragma unroll
      for (int j = 0; j < DJ; ++j) {
        const int d = tid + j * PT;
        e[j] = (d < D && row_ok) ? load_emb(er, tr, table_dtype, d) : 0.f;
      }
      const float4* wrow = reinterpret_cast<const float4*>(WT + h * KP + kg);
#pragma unroll
      for (int k4 = 0; k4 < KG / 4; ++k4) {
        if (kg + k4 * 4 < K) {
          const float4 w = wrow[k4];
#pragma unroll
          for (int j = 0; j < DJ; ++j) {
            w_acc[k4 * 4 + 0][j] = fmaf(w.x, e[j], w_acc[k4 * 4 + 0][j]);
            w_acc[k4 * 4 + 1][j] = fmaf(w.y, e[j], w_acc[k4 * 4 + 1][j]);
            w_acc[k4 * 4 + 2][j] = fmaf(w.z, e[j], w_acc[k4 * 4 + 2][j]);
            w_acc[k4 * 4 + 3][j] = fmaf(w.w, e[j], w_acc[k4 * 4 + 3][j]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KG; ++k) {
      if (kg + k < K) {
#pragma unroll
        for (int j = 0; j < DJ; ++j) {
          const int d = tid + j * PT;
          if (d < D) {
            const int64_t o = (b * K + kg + k) * static_cast<int64_t>(D) + d;
            if (out_interests) out_interests[o] = w_acc[k][j];
            if (out_interests_bf16) out_interests_bf16[o] = __float2bfloat16_rn(w_acc[k][j]);
          }
        }
      }
    }
  }
}

static size_t poly_smem_bytes(int H, int K) {
  const int KP = (K + 3) & ~3;
  return sizeof(float) * (static_cast<size_t>(DCB) * (H + 4) + static_cast<size_t>(DCB) * (KP + 4) +
                          ((static_cast<size_t>(K) * (H + 1) + 3) & ~size_t(3)) + static_cast<size_t>(H) * KP);
}

int launch_poly_softmax_wsum(const float* proj, const float* codes, const uint8_t* mask, const float* bias_mean,
                             const float* emb, const void* table, int table_dtype, const void* his_ids, int id_dtype,
                             int64_t n_rows, int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D,
                             float* out_interests, float* out_weights, void* out_interests_bf16, cudaStream_t stream) {
  if (B == 0) return MINER_OK;
  if (H < 1 || H > 256 || K < 1 || K > 64 || H * K > PT * MAXT * TH * TK || D < 1 || D > 4 * PT) {
    set_error("poly attention: unsupported shape H=%lld K=%lld D=%lld (need H<=256, K<=64, D<=1024)", (long long)H, (long long)K, (long long)D);
    return MINER_ERR_UNSUPPORTED;
  }
  size_t smem = poly_smem_bytes(static_cast<int>(H), static_cast<int>(K));
  // bf16 rows of one impression staged in shared memory when they fit beside the logit tiles (two blocks per SM up to ~100 KB each)
  const size_t stage_bytes = static_cast<size_t>(H) * D * 2;
  const int stage_rows = (!emb && table && table_dtype == MINER_BF16 && D % 8 == 0 && D / 8 <= PT && smem + stage_bytes <= 110 * 1024 &&
                          reinterpret_cast<uintptr_t>(table) % 16 == 0 && (!out_interests || reinterpret_cast<uintptr_t>(out_interests) % 16 == 0) &&
                          (!out_interests_bf16 || reinterpret_cast<uintptr_t>(out_interests_bf16) % 16 == 0))
                             ? 1 : 0;
  if (stage_rows) smem += stage_bytes;
  const int dj = static_cast<int>((D + PT - 1) / PT);
  auto bf = static_cast<__nv_bfloat16*>(out_interests_bf16);
#define MINER_POLY(DJ)                                                                                                   \
  do {                                                                                                                   \
    MINER_CUDA_OK(cudaFuncSetAttribute(poly_softmax_wsum_kernel<DJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    poly_softmax_wsum_kernel<DJ><<<static_cast<unsigned>(B), PT, smem, stream>>>(                                        \
        proj, codes, mask, bias_mean, emb, table, table_dtype, his_ids, id_dtype, n_rows, (int)H, (int)K, (int)Dc, (int)D,       \
        out_interests, out_weights, bf, stage_rows);                                                                     \
  } while (0)
  switch (dj) {
    case 1: MINER_POLY(1); break;
    case 2: MINER_POLY(2); break;
    case 3: MINER_POLY(3); break;
    default: MINER_POLY(4); break;
  }
#undef MINER_POLY
  MINER_LAUNCH_OK("poly_softmax_wsum");
  return MINER_OK;
}

}  // namespace miner
