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
    float* __restrict__ out_interests, float* __restrict__ out_weights, __nv_bfloat16* __restrict__ out_interests_bf16) {
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

  // ---- logits tile: register-tiled (TH x TK) mini-GEMM over Dc chunks ----
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
          float a[TH], bb[TK];
#pragma unroll
          for (int i = 0; i < TH; ++i) a[i] = (h0 + i < H) ? As[c * (H + 4) + h0 + i] : 0.f;
#pragma unroll
          for (int j = 0; j < TK; ++j) bb[j] = (k0 + j < K) ? Bs[c * (KP + 4) + k0 + j] : 0.f;
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
  const size_t smem = poly_smem_bytes(static_cast<int>(H), static_cast<int>(K));
  const int dj = static_cast<int>((D + PT - 1) / PT);
  auto bf = static_cast<__nv_bfloat16*>(out_interests_bf16);
#define MINER_POLY(DJ)                                                                                                   \
  do {                                                                                                                   \
    MINER_CUDA_OK(cudaFuncSetAttribute(poly_softmax_wsum_kernel<DJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    poly_softmax_wsum_kernel<DJ><<<static_cast<unsigned>(B), PT, smem, stream>>>(                                        \
        proj, codes, mask, bias_mean, emb, table, table_dtype, his_ids, id_dtype, n_rows, (int)H, (int)K, (int)Dc, (int)D,       \
        out_interests, out_weights, bf);                                                                                 \
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
