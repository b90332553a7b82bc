// Train variant of the scoring path: forward with the intermediates the backward needs, the
// backward of Loss.compute (reference src/loss.py:27-44) and the backward of Miner.forward (src/model/model.py:61-138,
// score_type 'weighted', category bias off) down to the three weight matrices.  fp32 CUDA-core kernels in the reference's
// operation order; every reduction over the batch goes through per-CTA partials summed in a fixed order (deterministic).
// The news table is a frozen buffer (TableNewsEncoder): no gradient flows into its rows.
//
//   forward   E = table[his]; T = tanh(E Wp^T); logits = T codes^T, masked := 1e-30; w = softmax_H; I = w E        (model.py:159-185)
//             Z = I Wt^T; G = gelu(Z); m = Cd I^T; a = Cd G^T; ws = softmax_K(a); s = sum_k ws m                     (model.py:127,200-216)
//   loss      mean_{b,k!=l} cos(I_k, I_l) + CE_mean(s, argmax labels)                                                 (loss.py:39-42)
//   backward  ds, dI(loss) -> dm = ds ws, dws = ds m, da = ws (dws - sum ws dws) -> dI += dm^T Cd, dG = da^T Cd, dZ = dG gelu'(Z)
//             dI += dZ Wt, dWt = dZ^T I -> dw = dI E^T -> dlogits = w (dw - sum w dw), 0 on masked slots
//             dT = dlogits^T codes, dcodes = sum_b dlogits T, dZ1 = dT (1 - T^2), dWp = dZ1^T E
#include "common.cuh"
#include "tc/tc_gemm.cuh"

namespace miner {

namespace {

constexpr int TT = 256;
#ifndef MINER_POLY_BWD_DK
#define MINER_POLY_BWD_DK 128
#endif
constexpr int POLY_BWD_DK = MINER_POLY_BWD_DK;   // feature chunk of poly_bwd's dw contraction

__device__ __forceinline__ float table_elem(const void* table, int dtype, int64_t idx) {
  return dtype == MINER_F32 ? reinterpret_cast<const float*>(table)[idx]
                            : bf16_bits_to_float(reinterpret_cast<const uint16_t*>(table)[idx]);
}
// d/dz of the exact erf gelu (model.py:212)
__device__ __forceinline__ float gelu_erf_grad(float z) {
  return 0.5f * (1.0f + erff(z * 0.70710678118654752440f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    if (r0 + i < rows && c0 + threadIdx.x < cols) tile[i][threadIdx.x] = src[static_cast<int64_t>(r0 + i) * cols + c0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    if (c0 + i < cols && r0 + threadIdx.x < rows) dst[static_cast<int64_t>(c0 + i) * rows + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// dst (cols x rpad, bf16) = src (rows x cols, fp32) transposed; columns rows..rpad-1 of dst are zero (K padding of the tensor-core GEMMs)
__global__ void transpose_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int cols, int64_t rpad) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    tile[i][threadIdx.x] = (r0 + i < rows && c0 + threadIdx.x < cols) ? src[(r0 + i) * cols + c0 + threadIdx.x] : 0.f;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    if (c0 + i < cols && r0 + threadIdx.x < rpad) dst[static_cast<int64_t>(c0 + i) * rpad + r0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][i]);
}
// dst (D x rpad, bf16): dst[d][r] = table[ids[r]][d]  (bf16 table; invalid ids and r >= rows give zeros)
__global__ void gather_transpose_kernel(const uint16_t* __restrict__ table, int64_t n_rows, const void* __restrict__ ids, int id_dtype,
                                        int64_t rows, int D, int64_t rpad, uint16_t* __restrict__ dst) {
  __shared__ uint16_t tile[32][34];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    uint16_t v = 0;
    if (r0 + i < rows && c0 + threadIdx.x < D) {
      const int64_t id = load_id(ids, r0 + i, id_dtype);
      if (id >= 0 && id < n_rows) v = table[id * D + c0 + threadIdx.x];
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y)
    if (c0 + i < D && r0 + threadIdx.x < rpad) dst[static_cast<int64_t>(c0 + i) * rpad + r0 + threadIdx.x] = tile[threadIdx.x][i];
}

// ---- backward of Loss.compute: one CTA per impression
//      d_logits = g (softmax(s) - onehot(argmax labels)) / B                                   (CrossEntropyLoss, reduction 'mean')
//      d_I[k]   = g 2 / (B K K) * (S_k - (u_k . S_k) u_k) / |I_k|,  u = I / |I|, S_k = sum_{l != k} u_l   (utils.py:21-27, loss.py:39)
__global__ void __launch_bounds__(TT) loss_bwd_kernel(const float* __restrict__ interests, const float* __restrict__ logits,
                                                      const float* __restrict__ labels, const float* __restrict__ grad_out, int64_t B, int C,
                                                      int K, int D, float* __restrict__ d_interests, float* __restrict__ d_logits) {
  extern __shared__ __align__(16) float smem[];
  const int DP = D + 1;
  float* U = smem;                       // [K][D+1] normalised interests
  float* Usum = U + K * DP;              // [D]
  float* nrm = Usum + D;                 // [K]
  float* dotU = nrm + K;                 // [K]   u_k . sum_l u_l
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const float g = grad_out ? grad_out[0] : 1.0f;
  const float* Ib = interests + b * static_cast<int64_t>(K) * D;
  for (int k = warp; k < K; k += TT / 32) {
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = Ib[static_cast<int64_t>(k) * D + d]; ss = fmaf(v, v, ss); }
    const float n = sqrtf(warp_sum(ss));
    if (lane == 0) nrm[k] = n;
    for (int d = lane; d < D; d += 32) U[k * DP + d] = Ib[static_cast<int64_t>(k) * D + d] / n;
  }
  __syncthreads();
  for (int d = tid; d < D; d += TT) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += U[k * DP + d];
    Usum[d] = s;
  }
  __syncthreads();
  for (int k = warp; k < K; k += TT / 32) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(U[k * DP + d], Usum[d], s);
    s = warp_sum(s);
    if (lane == 0) dotU[k] = s;
  }
  __syncthreads();
  const float scale = g * 2.0f / (static_cast<float>(B) * K * K);
  float* dIb = d_interests + b * static_cast<int64_t>(K) * D;
  for (int i = tid; i < K * D; i += TT) {
    const int k = i / D, d = i - k * D;
    const float u = U[k * DP + d];
    // S_k = Usum - u_k ;  u_k . S_k = u_k . Usum - |u_k|^2 with |u_k|^2 = 1
    dIb[i] = scale * ((Usum[d] - u) - (dotU[k] - 1.0f) * u) / nrm[k];
  }
  if (warp == 0) {
    const float* lg = logits + b * C;
    const float* lb = labels + b * C;
    float best = -INFINITY; int arg = 0;
    for (int c = 0; c < C; ++c) { const float v = lb[c]; if (v > best) { best = v; arg = c; } }     // argmax, first maximum
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(lg[c] - mx);
    se = warp_sum(se);
    for (int c = lane; c < C; c += 32)
      d_logits[b * C + c] = g * (expf(lg[c] - mx) / se - (c == arg ? 1.0f : 0.0f)) / static_cast<float>(B);
  }
}

// The same backward with the interest rows held in REGISTERS (D = 128 NV, K <= 32): 16 warps, two rows per warp, each lane owns NV
// float4 of a row; the norms and the dots with the column sum are warp reductions, the column sum of the normalised rows goes through
// one [16][D] shared-memory stage.  Memory-bound (reads I once, writes dI once) instead of three passes over a 98 KB shared-memory copy.
constexpr int LBW = 16;                  // warps per CTA
template <int NV>
__global__ void __launch_bounds__(LBW * 32) loss_bwd_rows_kernel(const float* __restrict__ interests, const float* __restrict__ logits,
                                                                 const float* __restrict__ labels, const float* __restrict__ grad_out, int64_t B,
                                                                 int C, int K, float* __restrict__ d_interests, float* __restrict__ d_logits) {
  constexpr int D = 128 * NV;
  extern __shared__ __align__(16) float smem[];
  float (*part)[D] = reinterpret_cast<float (*)[D]>(smem);     // [LBW][D] per-warp sums of its normalised rows
  float* Usum = smem + LBW * D;                                // [D]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const float g = grad_out ? grad_out[0] : 1.0f;
  float4 u[2][NV];
  float nrm[2];
  float4 ps[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) ps[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = warp + LBW * r;
    nrm[r] = 1.f;
    if (k < K) {
      const float4* row = reinterpret_cast<const float4*>(interests + (b * K + k) * D);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = row[lane + 32 * j];
        u[r][j] = v;
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
      }
      const float n = sqrtf(warp_sum(ss));
      nrm[r] = n;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float4 v = u[r][j];
        v.x /= n; v.y /= n; v.z /= n; v.w /= n;                    // utils.py:21-23: divide first
        u[r][j] = v;
        ps[j].x += v.x; ps[j].y += v.y; ps[j].z += v.z; ps[j].w += v.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) reinterpret_cast<float4*>(part[warp])[lane + 32 * j] = ps[j];
  __syncthreads();
  for (int d = tid; d < D; d += LBW * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LBW; ++w) s += part[w][d];
    Usum[d] = s;
  }
  __syncthreads();
  const float scale = g * 2.0f / (static_cast<float>(B) * K * K);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = warp + LBW * r;
    if (k < K) {
      float dot = 0.f;
      float4 us[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        us[j] = reinterpret_cast<const float4*>(Usum)[lane + 32 * j];
        const float4 v = u[r][j];
        dot = fmaf(v.x, us[j].x, fmaf(v.y, us[j].y, fmaf(v.z, us[j].z, fmaf(v.w, us[j].w, dot))));
      }
      dot = warp_sum(dot);
      // S_k = Usum - u_k ;  u_k . S_k = u_k . Usum - |u_k|^2 with |u_k|^2 = 1
      const float c1 = dot - 1.0f, inv = scale / nrm[r];
      float4* out = reinterpret_cast<float4*>(d_interests + (b * K + k) * D);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 v = u[r][j];
        out[lane + 32 * j] = make_float4(((us[j].x - v.x) - c1 * v.x) * inv, ((us[j].y - v.y) - c1 * v.y) * inv, ((us[j].z - v.z) - c1 * v.z) * inv,
                                         ((us[j].w - v.w) - c1 * v.w) * inv);
      }
    }
  }
  if (warp == 0) {
    const float* lg = logits + b * C;
    const float* lb = labels + b * C;
    float best = -INFINITY; int arg = 0;
    for (int c = 0; c < C; ++c) { const float v = lb[c]; if (v > best) { best = v; arg = c; } }     // argmax, first maximum
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(lg[c] - mx);
    se = warp_sum(se);
    for (int c = lane; c < C; c += 32)
      d_logits[b * C + c] = g * (expf(lg[c] - mx) / se - (c == arg ? 1.0f : 0.0f)) / static_cast<float>(B);
  }
}

// ---- backward through the target-aware attention of one impression (model.py:127,200-216): one CTA per impression
__global__ void __launch_bounds__(TT) target_bwd_kernel(const void* __restrict__ table, int table_dtype, int64_t n_rows,
                                                        const void* __restrict__ cand_ids, int id_dtype, const float* __restrict__ interests,
                                                        const float* __restrict__ Z, const float* __restrict__ d_scores,
                                                        const float* __restrict__ d_interests_in, int C, int K, int D,
                                                        float* __restrict__ d_interests, float* __restrict__ dZ, float* __restrict__ grad_table) {
  extern __shared__ __align__(16) float smem[];
  float* Cd = smem;                      // [C][D] candidate vectors
  float* m = Cd + C * D;                 // [C][K] matching scores
  float* a = m + C * K;                  // [C][K] attention logits -> softmax weights
  float* dm = a + C * K;                 // [C][K]
  float* da = dm + C * K;                // [C][K]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  if (table_dtype == MINER_BF16 && (D & 7) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0) {
    const int nv = D >> 3;                                  // 16-byte loads: eight bf16 features of a candidate row per request
    for (int i = tid; i < C * nv; i += TT) {
      const int c = i / nv, v = i - c * nv;
      const int64_t id = load_id(cand_ids, b * C + c, id_dtype);
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (id >= 0 && id < n_rows) raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(table) + id * D) + v);
      float4* dst = reinterpret_cast<float4*>(Cd + c * D + 8 * v);
      dst[0] = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                           __uint_as_float(raw.y & 0xffff0000u));
      dst[1] = make_float4(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u), __uint_as_float(raw.w << 16),
                           __uint_as_float(raw.w & 0xffff0000u));
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const int64_t id = load_id(cand_ids, b * C + c, id_dtype);
      const bool ok = id >= 0 && id < n_rows;
      for (int d = tid; d < D; d += TT) Cd[c * D + d] = ok ? table_elem(table, table_dtype, id * D + d) : 0.f;
    }
  }
  __syncthreads();
  const float* Ib = interests + b * static_cast<int64_t>(K) * D;
  const float* Zb = Z + b * static_cast<int64_t>(K) * D;
  constexpr int CB = 8;                                     // candidates per sweep over an interest row: gelu(Z) is evaluated once per sweep
  const float* dIin = d_interests_in ? d_interests_in + b * static_cast<int64_t>(K) * D : nullptr;
  float* dIb = d_interests + b * static_cast<int64_t>(K) * D;
  float* dZb = dZ + b * static_cast<int64_t>(K) * D;
  // four features per request (16-byte global loads, LDS.128 of the candidate rows) when the rows allow it
  const bool vec4 = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(interests) | reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(d_interests) |
                                      reinterpret_cast<uintptr_t>(dZ) | reinterpret_cast<uintptr_t>(d_interests_in)) & 15) == 0;
  for (int k = warp; k < K; k += TT / 32) {
    for (int c0 = 0; c0 < C; c0 += CB) {
      float sm_[CB], sa_[CB];
#pragma unroll
      for (int j = 0; j < CB; ++j) sm_[j] = sa_[j] = 0.f;
      if (vec4) {
        const float4* I4 = reinterpret_cast<const float4*>(Ib + static_cast<int64_t>(k) * D);
        const float4* Z4 = reinterpret_cast<const float4*>(Zb + static_cast<int64_t>(k) * D);
        for (int d4 = lane; d4 < (D >> 2); d4 += 32) {
          const float4 iv = I4[d4], zv = Z4[d4];
          const float4 gv = make_float4(gelu_erf(zv.x), gelu_erf(zv.y), gelu_erf(zv.z), gelu_erf(zv.w));   // model.py:212
#pragma unroll
          for (int j = 0; j < CB; ++j) {
            if (c0 + j < C) {
              const float4 cd = *reinterpret_cast<const float4*>(Cd + (c0 + j) * D + 4 * d4);
              sm_[j] = fmaf(cd.x, iv.x, fmaf(cd.y, iv.y, fmaf(cd.z, iv.z, fmaf(cd.w, iv.w, sm_[j]))));     // model.py:127
              sa_[j] = fmaf(cd.x, gv.x, fmaf(cd.y, gv.y, fmaf(cd.z, gv.z, fmaf(cd.w, gv.w, sa_[j]))));     // model.py:213
            }
          }
        }
      } else {
        for (int d = lane; d < D; d += 32) {
          const float iv = Ib[static_cast<int64_t>(k) * D + d];
          const float gv = gelu_erf(Zb[static_cast<int64_t>(k) * D + d]);                // model.py:212
#pragma unroll
          for (int j = 0; j < CB; ++j) {
            const float cd = c0 + j < C ? Cd[(c0 + j) * D + d] : 0.f;
            sm_[j] = fmaf(cd, iv, sm_[j]);                                               // model.py:127
            sa_[j] = fmaf(cd, gv, sa_[j]);                                               // model.py:213
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const float vm = warp_sum(sm_[j]), va = warp_sum(sa_[j]);
        if (lane == 0 && c0 + j < C) { m[(c0 + j) * K + k] = vm; a[(c0 + j) * K + k] = va; }
      }
    }
  }
  __syncthreads();
  if (tid < C) {
    const int c = tid;
    const float ds = d_scores[b * C + c];
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, a[c * K + k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(a[c * K + k] - mx);
    float inner = 0.f;
    for (int k = 0; k < K; ++k) {
      const float w = expf(a[c * K + k] - mx) / se;                                   // model.py:213
      a[c * K + k] = w;
      inner = fmaf(w, ds * m[c * K + k], inner);
    }
    for (int k = 0; k < K; ++k) {
      const float w = a[c * K + k];
      dm[c * K + k] = ds * w;                                                         // d score / d m   (model.py:214)
      da[c * K + k] = w * (ds * m[c * K + k] - inner);                                // softmax backward
    }
  }
  __syncthreads();
  if (vec4) {
    const int nv = D >> 2;
    for (int i = tid; i < K * nv; i += TT) {
      const int k = i / nv, d = 4 * (i - k * nv);
      float4 gi = dIin ? *reinterpret_cast<const float4*>(dIin + static_cast<int64_t>(k) * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < C; ++c) {
        const float4 cd = *reinterpret_cast<const float4*>(Cd + c * D + d);
        const float wm = dm[c * K + k], wa = da[c * K + k];
        gi.x = fmaf(wm, cd.x, gi.x), gi.y = fmaf(wm, cd.y, gi.y), gi.z = fmaf(wm, cd.z, gi.z), gi.w = fmaf(wm, cd.w, gi.w);
        gg.x = fmaf(wa, cd.x, gg.x), gg.y = fmaf(wa, cd.y, gg.y), gg.z = fmaf(wa, cd.z, gg.z), gg.w = fmaf(wa, cd.w, gg.w);
      }
      const float4 zv = *reinterpret_cast<const float4*>(Zb + static_cast<int64_t>(k) * D + d);
      *reinterpret_cast<float4*>(dIb + static_cast<int64_t>(k) * D + d) = gi;
      *reinterpret_cast<float4*>(dZb + static_cast<int64_t>(k) * D + d) =
          make_float4(gg.x * gelu_erf_grad(zv.x), gg.y * gelu_erf_grad(zv.y), gg.z * gelu_erf_grad(zv.z), gg.w * gelu_erf_grad(zv.w));
    }
  } else {
    for (int i = tid; i < K * D; i += TT) {
      const int k = i / D, d = i - k * D;
      float gi = dIin ? dIin[i] : 0.f, gg = 0.f;
      for (int c = 0; c < C; ++c) {
        const float cd = Cd[c * D + d];
        gi = fmaf(dm[c * K + k], cd, gi);
        gg = fmaf(da[c * K + k], cd, gg);
      }
      dIb[i] = gi;
      dZb[i] = gg * gelu_erf_grad(Zb[i]);
    }
  }
  if (grad_table) {
    // gradient of the candidate rows: m = Cd I^T and a = Cd gelu(Z)^T are both linear in Cd (model.py:127,213)
    for (int i = tid; i < C * D; i += TT) {
      const int c = i / D, d = i - c * D;
      const int64_t id = load_id(cand_ids, b * C + c, id_dtype);
      if (id < 0 || id >= n_rows) continue;
      float gc = 0.f;
      for (int k = 0; k < K; ++k)
        gc = fmaf(dm[c * K + k], Ib[static_cast<int64_t>(k) * D + d], fmaf(da[c * K + k], gelu_erf(Zb[static_cast<int64_t>(k) * D + d]), gc));
      atomicAdd(grad_table + id * D + d, gc);
    }
  }
}

// ---- score_type 'max' / 'mean' (model.py:128-131): scores = max_k / mean_k of m[c,k] = Cd_c . I_k.  One CTA per impression:
//      dm[c,k] = ds[c] [k == argmax_k m[c,:]]  or  ds[c] / K ;  dI[k] = dI_in[k] + sum_c dm[c,k] Cd_c
__global__ void __launch_bounds__(TT) target_bwd_simple_kernel(const void* __restrict__ table, int table_dtype, int64_t n_rows,
                                                               const void* __restrict__ cand_ids, int id_dtype, const float* __restrict__ interests,
                                                               const float* __restrict__ d_scores, const float* __restrict__ d_interests_in, int C,
                                                               int K, int D, int score_type, float* __restrict__ d_interests,
                                                               float* __restrict__ grad_table) {
  extern __shared__ __align__(16) float smem[];
  float* Cd = smem;                      // [C][D] candidate vectors
  float* m = Cd + C * D;                 // [C][K] matching scores -> dm
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  for (int c = 0; c < C; ++c) {
    const int64_t id = load_id(cand_ids, b * C + c, id_dtype);
    const bool ok = id >= 0 && id < n_rows;
    for (int d = tid; d < D; d += TT) Cd[c * D + d] = ok ? table_elem(table, table_dtype, id * D + d) : 0.f;
  }
  __syncthreads();
  const float* Ib = interests + b * static_cast<int64_t>(K) * D;
  if (score_type == MINER_SCORE_MAX) {
    for (int i = warp; i < C * K; i += TT / 32) {
      const int c = i / K, k = i - c * K;
      float sm_ = 0.f;
      for (int d = lane; d < D; d += 32) sm_ = fmaf(Cd[c * D + d], Ib[static_cast<int64_t>(k) * D + d], sm_);    // model.py:127
      sm_ = warp_sum(sm_);
      if (lane == 0) m[i] = sm_;
    }
    __syncthreads();
    if (tid < C) {
      const int c = tid;
      int arg = 0;
      float best = m[c * K];
      for (int k = 1; k < K; ++k) if (m[c * K + k] > best) { best = m[c * K + k]; arg = k; }       // first maximum (model.py:129)
      const float ds = d_scores[b * C + c];
      for (int k = 0; k < K; ++k) m[c * K + k] = k == arg ? ds : 0.f;
    }
  } else {
    for (int i = tid; i < C * K; i += TT) m[i] = d_scores[b * C + i / K] / static_cast<float>(K);  // model.py:131
  }
  __syncthreads();
  const float* dIin = d_interests_in ? d_interests_in + b * static_cast<int64_t>(K) * D : nullptr;
  float* dIb = d_interests + b * static_cast<int64_t>(K) * D;
  for (int i = tid; i < K * D; i += TT) {
    const int k = i / D, d = i - k * D;
    float gi = dIin ? dIin[i] : 0.f;
    for (int c = 0; c < C; ++c) gi = fmaf(m[c * K + k], Cd[c * D + d], gi);
    dIb[i] = gi;
  }
  if (grad_table) {
    for (int i = tid; i < C * D; i += TT) {
      const int c = i / D, d = i - c * D;
      const int64_t id = load_id(cand_ids, b * C + c, id_dtype);
      if (id < 0 || id >= n_rows) continue;
      float gc = 0.f;
      for (int k = 0; k < K; ++k) gc = fmaf(m[c * K + k], Ib[static_cast<int64_t>(k) * D + d], gc);
      atomicAdd(grad_table + id * D + d, gc);
    }
  }
}

// ---- backward through the poly attention (model.py:159-185): CTAs stride over impressions and keep their share of dcodes in
//      shared memory; dZ1 = dT (1 - T^2) goes to global memory for the dWp contraction
__global__ void __launch_bounds__(TT) poly_bwd_kernel(const void* __restrict__ table, int table_dtype, int64_t n_rows,
                                                      const void* __restrict__ his_ids, int id_dtype, const uint8_t* __restrict__ mask,
                                                      const float* __restrict__ codes, const float* __restrict__ T,
                                                      const float* __restrict__ W, const float* __restrict__ dI_a,
                                                      const float* __restrict__ dI_b, int64_t B, int H, int K, int Dc, int D,
                                                      float* __restrict__ dZ1, float* __restrict__ dcodes_partial, float* __restrict__ d_bias) {
  extern __shared__ __align__(16) float smem[];
  constexpr int DK = POLY_BWD_DK;        // feature chunk (128: half as many fill / barrier rounds per impression as 64, still two blocks per SM)
  float* codes_s = smem;                 // [K][Dc]
  float* dcodes_s = codes_s + K * Dc;    // [K][Dc]
  constexpr int DS = DK + 12;            // row stride of the chunk tiles: 16-byte aligned rows; rows two apart (a warp's slot tiles) are 24 banks apart: conflict-free LDS.128
  float* dIc = smem + ((2 * K * Dc + 3) & ~3);   // [K][DS], 16-byte aligned
  float* Ec = dIc + K * DS;              // [H][DS]
  float* dw = Ec + H * DS;               // [K][H]  dw -> dlogits
  int* ids_s = reinterpret_cast<int*>(dw + K * H);   // [H] table row of each history slot, -1 = zero row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < K * Dc; i += TT) { codes_s[i] = codes[i]; dcodes_s[i] = 0.f; }
  __syncthreads();
  const int npairs = K * H;
  const bool vec_fill = table_dtype == MINER_BF16 && (D & 7) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(dI_a) & 15) == 0 && (!dI_b || (reinterpret_cast<uintptr_t>(dI_b) & 15) == 0);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int h = tid; h < H; h += TT) {
      const int64_t id = load_id(his_ids, b * H + h, id_dtype);
      ids_s[h] = (id >= 0 && id < n_rows) ? static_cast<int>(id) : -1;
    }
    for (int i = tid; i < npairs; i += TT) dw[i] = 0.f;
    __syncthreads();
    // dw[k,h] = sum_d dI[k,d] E[h,d]                                                   (model.py:182)
    for (int d0 = 0; d0 < D; d0 += DK) {
      const int dn = D - d0 < DK ? D - d0 : DK;
      if (vec_fill && dn == DK) {
        // 16-byte loads: four gradient values, eight bf16 features of a table row per request
        for (int i = tid; i < K * (DK / 4); i += TT) {
          const int k = i / (DK / 4), d4 = i - k * (DK / 4);
          const int64_t o = (b * K + k) * D + d0 + 4 * d4;
          float4 g = *reinterpret_cast<const float4*>(dI_a + o);
          if (dI_b) {
            const float4 g2 = *reinterpret_cast<const float4*>(dI_b + o);
            g.x += g2.x, g.y += g2.y, g.z += g2.z, g.w += g2.w;
          }
          *reinterpret_cast<float4*>(dIc + k * DS + 4 * d4) = g;
        }
        for (int i = tid; i < H * (DK / 8); i += TT) {
          const int h = i / (DK / 8), c = i - h * (DK / 8);
          const int id = ids_s[h];
          uint4 raw = make_uint4(0u, 0u, 0u, 0u);
          if (id >= 0) raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(table) + static_cast<int64_t>(id) * D + d0) + c);
          float4* dst = reinterpret_cast<float4*>(Ec + h * DS + 8 * c);
          dst[0] = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                               __uint_as_float(raw.y & 0xffff0000u));
          dst[1] = make_float4(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u), __uint_as_float(raw.w << 16),
                               __uint_as_float(raw.w & 0xffff0000u));
        }
      } else {
        for (int i = tid; i < K * DK; i += TT) {
          const int k = i / DK, d = i - k * DK;
          const int64_t o = (b * K + k) * D + d0 + d;
          dIc[k * DS + d] = d < dn ? dI_a[o] + (dI_b ? dI_b[o] : 0.f) : 0.f;
        }
        for (int i = tid; i < H * DK; i += TT) {
          const int h = i / DK, d = i - h * DK;
          const int id = ids_s[h];
          Ec[h * DS + d] = (d < dn && id >= 0) ? table_elem(table, table_dtype, static_cast<int64_t>(id) * D + d0 + d) : 0.f;
        }
      }
      __syncthreads();
      // 4 (codes) x 2 (slots) register tiles, four features per step: 6 LDS.128 per 32 FMAs
      const int th_n = (H + 1) / 2, tk_n = (K + 3) / 4;
      for (int t = tid; t < th_n * tk_n; t += TT) {
        const int tk = t / th_n, th = t - tk * th_n;
        const int k0 = tk * 4, h0 = th * 2;
        const float4* x[4];
        const float4* y[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = reinterpret_cast<const float4*>(dIc + (k0 + i < K ? k0 + i : K - 1) * DS);
#pragma unroll
        for (int j = 0; j < 2; ++j) y[j] = reinterpret_cast<const float4*>(Ec + (h0 + j < H ? h0 + j : H - 1) * DS);
        float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 4
        for (int d = 0; d < DK / 4; ++d) {
          const float4 y0 = y[0][d], y1 = y[1][d];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 xv = x[i][d];
            acc[i][0] = fmaf(xv.x, y0.x, fmaf(xv.y, y0.y, fmaf(xv.z, y0.z, fmaf(xv.w, y0.w, acc[i][0]))));
            acc[i][1] = fmaf(xv.x, y1.x, fmaf(xv.y, y1.y, fmaf(xv.z, y1.z, fmaf(xv.w, y1.w, acc[i][1]))));
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (k0 + i < K && h0 + j < H) dw[(k0 + i) * H + h0 + j] += acc[i][j];
      }
      __syncthreads();
    }
    // softmax backward over the history; masked slots were overwritten by a constant (model.py:180): no gradient
    for (int k = warp; k < K; k += TT / 32) {
      const float* wk = W + (b * K + k) * H;
      float inner = 0.f;
      for (int h = lane; h < H; h += 32) inner = fmaf(wk[h], dw[k * H + h], inner);
      inner = warp_sum(inner);
      for (int h = lane; h < H; h += 32) {
        const float gl = wk[h] * (dw[k * H + h] - inner);
        dw[k * H + h] = mask[b * H + h] ? gl : 0.f;
      }
    }
    __syncthreads();
    // the category bias adds one scalar per history slot to all K logits of the slot (model.py:176-177): d bias[h] = sum_k dlogits[k,h]
    if (d_bias)
      for (int h = tid; h < H; h += TT) {
        float sb = 0.f;
        for (int k = 0; k < K; ++k) sb += dw[k * H + h];
        d_bias[b * H + h] = sb;
      }
    // dT = dlogits^T codes, dZ1 = dT (1 - T^2)                                          (model.py:171,174)
    if ((Dc & 3) == 0) {
      // 2 (slots) x 4 (code dimensions) register tiles: 3 shared-memory loads per 8 FMAs
      const int tc_n = Dc / 4, th2_n = (H + 1) / 2;
      for (int t = tid; t < th2_n * tc_n; t += TT) {
        const int th = t / tc_n, c0 = (t - th * tc_n) * 4;
        const int h0 = th * 2, h1 = h0 + 1 < H ? h0 + 1 : h0;
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
          const float4 cv = *reinterpret_cast<const float4*>(codes_s + k * Dc + c0);
          const float w0 = dw[k * H + h0], w1 = dw[k * H + h1];
          a0[0] = fmaf(w0, cv.x, a0[0]); a0[1] = fmaf(w0, cv.y, a0[1]); a0[2] = fmaf(w0, cv.z, a0[2]); a0[3] = fmaf(w0, cv.w, a0[3]);
          a1[0] = fmaf(w1, cv.x, a1[0]); a1[1] = fmaf(w1, cv.y, a1[1]); a1[2] = fmaf(w1, cv.z, a1[2]); a1[3] = fmaf(w1, cv.w, a1[3]);
        }
        const int64_t o0 = (b * H + h0) * Dc + c0;
        const float4 t0 = *reinterpret_cast<const float4*>(T + o0);
        *reinterpret_cast<float4*>(dZ1 + o0) = make_float4(a0[0] * (1.0f - t0.x * t0.x), a0[1] * (1.0f - t0.y * t0.y), a0[2] * (1.0f - t0.z * t0.z),
                                                            a0[3] * (1.0f - t0.w * t0.w));
        if (h0 + 1 < H) {
          const int64_t o1 = o0 + Dc;
          const float4 t1 = *reinterpret_cast<const float4*>(T + o1);
          *reinterpret_cast<float4*>(dZ1 + o1) = make_float4(a1[0] * (1.0f - t1.x * t1.x), a1[1] * (1.0f - t1.y * t1.y), a1[2] * (1.0f - t1.z * t1.z),
                                                              a1[3] * (1.0f - t1.w * t1.w));
        }
      }
    } else {
      for (int i = tid; i < H * Dc; i += TT) {
        const int h = i / Dc, dc = i - h * Dc;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(dw[k * H + h], codes_s[k * Dc + dc], s);
        const float t = T[(b * H + h) * Dc + dc];
        dZ1[(b * H + h) * Dc + dc] = s * (1.0f - t * t);
      }
    }
    // dcodes += dlogits T                                                              (model.py:174)
    if ((Dc & 3) == 0) {
      // 4 (codes) x 4 (code dimensions) register tiles: one 16-byte T load + 4 shared-memory loads per 16 FMAs
      const int tc_n = Dc / 4;
      for (int i = tid; i < ((K + 3) / 4) * tc_n; i += TT) {
        const int k4 = i / tc_n, c0 = (i - k4 * tc_n) * 4;
        const int k0 = k4 * 4;
        const float* wr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wr[j] = dw + (k0 + j < K ? k0 + j : K - 1) * H;
        float s[4][4] = {};
        const float* tp = T + b * H * Dc + c0;
#pragma unroll 2
        for (int h = 0; h < H; ++h) {
          const float4 tv = *reinterpret_cast<const float4*>(tp + static_cast<int64_t>(h) * Dc);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float wv = wr[j][h];
            s[j][0] = fmaf(wv, tv.x, s[j][0]); s[j][1] = fmaf(wv, tv.y, s[j][1]); s[j][2] = fmaf(wv, tv.z, s[j][2]); s[j][3] = fmaf(wv, tv.w, s[j][3]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k0 + j < K) {
            float* dc4 = dcodes_s + (k0 + j) * Dc + c0;
            dc4[0] += s[j][0]; dc4[1] += s[j][1]; dc4[2] += s[j][2]; dc4[3] += s[j][3];
          }
      }
    } else {
      for (int i = tid; i < ((K + 3) / 4) * Dc; i += TT) {                             // 4 codes per thread: one T load feeds 4 FMAs
        const int k4 = i / Dc, dc = i - k4 * Dc;
        const int k0 = k4 * 4;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int h = 0; h < H; ++h) {
          const float tv = T[(b * H + h) * Dc + dc];
#pragma unroll
          for (int j = 0; j < 4; ++j) s[j] = fmaf(dw[(k0 + j < K ? k0 + j : K - 1) * H + h], tv, s[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k0 + j < K) dcodes_s[(k0 + j) * Dc + dc] += s[j];
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < K * Dc; i += TT) dcodes_partial[static_cast<int64_t>(blockIdx.x) * K * Dc + i] = dcodes_s[i];
}

// ---- gradient of the history rows of the table (the table as a trainable parameter / the dense output of an upstream encoder):
//      I = w E and T = tanh(E Wp^T) are the two places a history row enters (model.py:171,182), so
//      dE[h] = sum_k w[k,h] dI[k] + dZ1[h] Wp   (dE2 = dZ1 Wp comes from a GEMM); rows are scattered with atomics (a news id can
//      occur in many impressions).  Masked slots keep their softmax weight (the 1e-30 fill), so they get their gradient too.
__global__ void __launch_bounds__(TT) hist_table_grad_kernel(const void* __restrict__ his_ids, int id_dtype, int64_t n_rows,
                                                             const float* __restrict__ W, const float* __restrict__ dI_a,
                                                             const float* __restrict__ dI_b, const float* __restrict__ dE2, int H, int K, int D,
                                                             float* __restrict__ grad_table) {
  extern __shared__ __align__(16) float smem[];
  float* w = smem;                          // [K][H]
  int64_t* ids_s = reinterpret_cast<int64_t*>(w + ((K * H + 1) & ~1));
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  for (int i = tid; i < K * H; i += TT) w[i] = W[b * K * H + i];
  for (int h = tid; h < H; h += TT) {
    const int64_t id = load_id(his_ids, b * H + h, id_dtype);
    ids_s[h] = (id >= 0 && id < n_rows) ? id : -1;
  }
  __syncthreads();
  for (int d = tid; d < D; d += TT) {
    for (int h = 0; h < H; ++h) {
      const int64_t id = ids_s[h];
      if (id < 0) continue;
      float g = dE2[(b * H + h) * D + d];
      for (int k = 0; k < K; ++k) {
        const int64_t o = (b * K + k) * D + d;
        g = fmaf(w[k * H + h], dI_a[o] + (dI_b ? dI_b[o] : 0.f), g);
      }
      atomicAdd(grad_table + id * D + d, g);
    }
  }
}

// ---- out_partial[s][m][n] = sum over the rows r of split s of A[r][m] * Bm[r][n]   (A^T B over a row range);  Bm rows are
//      dense fp32 or gathered from the table.  64 x 64 output tile per CTA, 4 x 4 per thread, 16 rows per step.
__global__ void __launch_bounds__(TT) atb_kernel(const float* __restrict__ A, int lda, int M, const float* __restrict__ Bd,
                                                 const void* __restrict__ table, int table_dtype, int64_t n_rows, const void* __restrict__ ids,
                                                 int id_dtype, int ldb, int N, int64_t R, int64_t rows_per_split,
                                                 float* __restrict__ out_partial) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  __shared__ int64_t rid[16];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.z) * rows_per_split;
  const int64_t r_end = r_begin + rows_per_split < R ? r_begin + rows_per_split : R;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
    if (ids && tid < 16) {
      int64_t id = -1;
      if (r0 + tid < r_end) { id = load_id(ids, r0 + tid, id_dtype); if (id < 0 || id >= n_rows) id = -1; }
      rid[tid] = id;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * TT, rr = e >> 6, cc = e & 63;
      const int64_t r = r0 + rr;
      const bool rok = r < r_end;
      As[rr][cc] = (rok && m0 + cc < M) ? A[r * lda + m0 + cc] : 0.f;
      float bv = 0.f;
      if (rok && n0 + cc < N) {
        if (ids) { const int64_t id = rid[rr]; bv = id >= 0 ? table_elem(table, table_dtype, id * ldb + n0 + cc) : 0.f; }
        else bv = Bd[r * ldb + n0 + cc];
      }
      Bs[rr][cc] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* outp = out_partial + static_cast<int64_t>(blockIdx.z) * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mm = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
      if (mm < M && nn < N) outp[static_cast<int64_t>(mm) * N + nn] = acc[i][j];
    }
}

// out[i] = sum_s partial[s][i] in split order
__global__ void sum_partials_kernel(const float* __restrict__ partial, int n_splits, int64_t n, float* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < n_splits; ++p) s += partial[static_cast<int64_t>(p) * n + i];
    out[i] = s;
  }
}

int atb_splits(int64_t R, int tiles) {
  int s = (8 * sm_count() + tiles - 1) / tiles;
  if (s < 1) s = 1;
  const int64_t max_s = (R + 255) / 256;
  if (s > max_s) s = static_cast<int>(max_s);
  if (s < 1) s = 1;
  return s;
}

struct TrainWs {
  size_t g, di, di2, dz, dz1, wt_t, part_wp, part_wt, part_codes, total;
  int s_wp, s_wt, g_poly;
  // tensor-core family: bf16 operands of the five projection-sized GEMMs
  size_t i_bf16, dz_bf16, wtt_bf16, dzt, it, dz1t, et;
  int64_t bk_pad, bh_pad;
};
TrainWs train_ws(int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D, int math) {
  TrainWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t f = sizeof(float);
  w.g = take(f * B * K * D);           // dI (Wt path) in the backward (the forward applies gelu(Z) inside the scoring kernel)
  w.di = take(f * B * K * D);
  w.dz = take(f * B * K * D);
  w.dz1 = take(f * B * H * Dc);
  w.wt_t = take(f * D * D);
  w.s_wp = atb_splits(B * H, static_cast<int>(((Dc + 63) / 64) * ((D + 63) / 64)));
  w.s_wt = atb_splits(B * K, static_cast<int>(((D + 63) / 64) * ((D + 63) / 64)));
  w.g_poly = static_cast<int>(B < 4 * sm_count() ? B : 4 * sm_count());
  w.bk_pad = (B * K + 63) / 64 * 64;
  w.bh_pad = (B * H + 63) / 64 * 64;
  if (math == MINER_MATH_TENSOR) {     // split-K partials of the two weight-gradient GEMMs (tc_gemm)
    w.s_wp = tc_gemm_splits(Dc, D, w.bh_pad);
    w.s_wt = tc_gemm_splits(D, D, w.bk_pad);
  }
  w.part_wp = take(f * w.s_wp * Dc * D);
  w.part_wt = take(f * w.s_wt * D * D);
  w.part_codes = take(f * w.g_poly * K * Dc);
  w.di2 = w.g;
  if (math == MINER_MATH_TENSOR) {
    w.i_bf16 = take(2 * static_cast<size_t>(B) * K * D);
    w.dz_bf16 = w.i_bf16;              // forward / backward never need both
    w.wtt_bf16 = take(2 * static_cast<size_t>(D) * D);
    w.dzt = take(2 * static_cast<size_t>(D) * w.bk_pad);
    w.it = take(2 * static_cast<size_t>(D) * w.bk_pad);
    w.dz1t = take(2 * static_cast<size_t>(Dc) * w.bh_pad);
    w.et = take(2 * static_cast<size_t>(D) * w.bh_pad);
  }
  w.total = off;
  return w;
}

int check_train_math(int math, int table_dtype, int64_t D, int64_t Dc, const void* w_proj_bf16, const void* w_target_bf16) {
  if (math == MINER_MATH_FP32) return MINER_OK;
  MINER_CHECK_ARG(math == MINER_MATH_TENSOR, "train: math must be MINER_MATH_FP32 or MINER_MATH_TENSOR");
  MINER_CHECK_ARG(table_dtype == MINER_BF16 && w_proj_bf16 && w_target_bf16, "train: the tensor-core family needs a bf16 table and bf16 weight copies");
  if (!tc_gemm_supported(D, Dc) || !tc_gemm_supported(D, D)) {
    set_error("train: the tensor-core family needs D %% 64 == 0 and Dc >= 16 (D=%lld Dc=%lld)", (long long)D, (long long)Dc);
    return MINER_ERR_UNSUPPORTED;
  }
  return MINER_OK;
}

}  // namespace
}  // namespace miner

using namespace miner;

extern "C" size_t miner_train_workspace_bytes(int64_t B, int64_t H, int64_t K, int64_t Dc, int64_t D, int math) {
  if (B <= 0) return 256;
  return train_ws(B, H, K, Dc, D, math).total;
}

extern "C" size_t miner_train_table_grad_workspace_bytes(int64_t B, int64_t H, int64_t Dc, int64_t D) {
  if (B <= 0) return 256;
  return align_up(sizeof(float) * static_cast<size_t>(D) * Dc, 256) + align_up(sizeof(float) * static_cast<size_t>(B) * H * D, 256);
}

extern "C" int miner_train_fwd(const void* table, int64_t n_rows, int table_dtype, const void* his_ids, const uint8_t* his_mask,
                               const void* cand_ids, int id_dtype, const float* w_proj, const float* codes, const float* w_target,
                               int64_t B, int64_t H, int64_t C, int64_t K, int64_t Dc, int64_t D, float* out_interests, float* out_scores,
                               float* save_t, float* save_w, float* save_z, int math, const void* w_proj_bf16, const void* w_target_bf16,
                               int score_type, const float* bias_mean, void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(B >= 0 && H > 0 && C > 0 && K > 0 && Dc > 0 && D > 0 && n_rows > 0, "train_fwd: bad sizes");
  if (score_type != MINER_SCORE_MAX && score_type != MINER_SCORE_MEAN && score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  if (B == 0) return MINER_OK;
  const bool weighted = score_type == MINER_SCORE_WEIGHTED;
  MINER_CHECK_ARG(table && his_ids && his_mask && cand_ids && w_proj && codes && out_interests && out_scores && save_t && save_w &&
                      (!weighted || (w_target && save_z)),
                  "train_fwd: null pointer");
  MINER_CHECK_ARG(table_dtype == MINER_F32 || table_dtype == MINER_BF16, "train_fwd: table dtype must be fp32 or bf16");
  MINER_CHECK_ARG(id_dtype == MINER_I32 || id_dtype == MINER_I64, "train_fwd: id dtype must be int32 or int64");
  int rc = check_train_math(math, table_dtype, D, Dc, w_proj_bf16, weighted ? w_target_bf16 : w_proj_bf16);
  if (rc) return rc;
  const TrainWs w = train_ws(B, H, K, Dc, D, math);
  if (!workspace || workspace_bytes < w.total) {
    set_error("train_fwd: workspace too small (%zu bytes needed)", w.total);
    return MINER_ERR_WORKSPACE;
  }
  auto st = static_cast<cudaStream_t>(stream);
  char* wsb = static_cast<char*>(workspace);
  if (math == MINER_MATH_TENSOR) {
    // the two projection GEMMs on tcgen05 (bf16 operands, fp32 accumulation), everything else as in the fp32 family
    rc = launch_tc_gemm(table, his_ids, id_dtype, n_rows, w_proj_bf16, save_t, nullptr, B * H, Dc, D, EPI_TANH, st);          // model.py:171
    if (rc) return rc;
    rc = launch_poly_softmax_wsum(save_t, codes, his_mask, bias_mean, nullptr, table, table_dtype, his_ids, id_dtype, n_rows, B, H, K, Dc, D,
                                  out_interests, save_w, wsb + w.i_bf16, st);                                                  // model.py:174-182
    if (rc) return rc;
    if (weighted) rc = launch_tc_gemm(wsb + w.i_bf16, nullptr, id_dtype, 0, w_target_bf16, save_z, nullptr, B * K, D, D, EPI_NONE, st);     // model.py:212
    if (rc) return rc;
  } else {
    rc = launch_sgemm_nt(table, table_dtype, his_ids, id_dtype, n_rows, w_proj, save_t, B * H, Dc, D, EPI_TANH, st);          // model.py:171
    if (rc) return rc;
    rc = launch_poly_softmax_wsum(save_t, codes, his_mask, bias_mean, nullptr, table, table_dtype, his_ids, id_dtype, n_rows, B, H, K, Dc, D,
                                  out_interests, save_w, nullptr, st);                                                          // model.py:174-182
    if (rc) return rc;
    if (weighted) rc = launch_sgemm_nt(out_interests, MINER_F32, nullptr, id_dtype, 0, w_target, save_z, B * K, D, D, EPI_NONE, st);         // model.py:212
    if (rc) return rc;
  }
  if (!weighted)                                                                                                               // model.py:128-131
    return launch_target_score(out_interests, nullptr, nullptr, nullptr, table, table_dtype, cand_ids, id_dtype, n_rows, nullptr, B, C, K, D,
                               score_type, out_scores, st);
  // gelu(Z) is applied inside the scoring kernel (Z itself is what the backward needs): no gelu(Z) array is written
  return launch_target_score(out_interests, save_z, nullptr, nullptr, table, table_dtype, cand_ids, id_dtype, n_rows, nullptr, B, C, K, D,
                             MINER_SCORE_WEIGHTED, out_scores, st, true);                                                           // model.py:127,213-214
}

extern "C" int miner_loss_bwd(const float* interests, const float* logits, const float* labels, const float* grad_out, int64_t B,
                              int64_t C, int64_t K, int64_t D, float* d_interests, float* d_logits, void* stream) {
  MINER_CHECK_ARG(B > 0 && C > 0 && K > 0 && D > 0, "loss_bwd: bad sizes");
  MINER_CHECK_ARG(interests && logits && labels && d_interests && d_logits, "loss_bwd: null pointer");
  if (K <= 2 * LBW && D % 128 == 0 && D <= 768 && (reinterpret_cast<uintptr_t>(interests) | reinterpret_cast<uintptr_t>(d_interests)) % 16 == 0) {
    auto st = static_cast<cudaStream_t>(stream);
    switch (D / 128) {      // rows in registers (see loss_bwd_rows_kernel)
#define MINER_LB(NVV)                                                                                                                          \
  case NVV:                                                                                                                                    \
    MINER_CUDA_OK(cudaFuncSetAttribute(loss_bwd_rows_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LBW + 1) * 128 * NVV * 4));    \
    loss_bwd_rows_kernel<NVV><<<static_cast<unsigned>(B), LBW * 32, (LBW + 1) * 128 * NVV * 4, st>>>(interests, logits, labels, grad_out, B,   \
                                                                                                     (int)C, (int)K, d_interests, d_logits);     \
    break
      MINER_LB(1); MINER_LB(2); MINER_LB(3); MINER_LB(4); MINER_LB(5); MINER_LB(6);
#undef MINER_LB
    }
    MINER_LAUNCH_OK("loss_bwd_rows");
    return MINER_OK;
  }
  const size_t smem = sizeof(float) * (static_cast<size_t>(K) * (D + 1) + D + 2 * K);
  if (smem > 220 * 1024) {
    set_error("loss_bwd: K=%lld D=%lld needs %zu bytes of shared memory", (long long)K, (long long)D, smem);
    return MINER_ERR_UNSUPPORTED;
  }
  MINER_CUDA_OK(cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  loss_bwd_kernel<<<static_cast<unsigned>(B), TT, smem, static_cast<cudaStream_t>(stream)>>>(interests, logits, labels, grad_out, B, (int)C,
                                                                                             (int)K, (int)D, d_interests, d_logits);
  MINER_LAUNCH_OK("loss_bwd");
  return MINER_OK;
}

extern "C" int miner_train_bwd(const void* table, int64_t n_rows, int table_dtype, const void* his_ids, const uint8_t* his_mask,
                               const void* cand_ids, int id_dtype, const float* w_proj, const float* codes, const float* w_target,
                               const float* save_t, const float* save_w, const float* interests, const float* save_z,
                               const float* d_scores, const float* d_interests, int64_t B, int64_t H, int64_t C, int64_t K, int64_t Dc,
                               int64_t D, float* grad_w_proj, float* grad_codes, float* grad_w_target, int math, const void* w_proj_bf16,
                               const void* w_target_bf16, int score_type, float* d_bias_mean, float* grad_table, void* table_grad_ws,
                               size_t table_grad_ws_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  MINER_CHECK_ARG(B > 0 && H > 0 && C > 0 && K > 0 && Dc > 0 && D > 0 && n_rows > 0, "train_bwd: bad sizes");
  if (score_type != MINER_SCORE_MAX && score_type != MINER_SCORE_MEAN && score_type != MINER_SCORE_WEIGHTED) {
    set_error("Invalid method of aggregating matching score");
    return MINER_ERR_SCORE_TYPE;
  }
  const bool weighted = score_type == MINER_SCORE_WEIGHTED;
  MINER_CHECK_ARG(table && his_ids && his_mask && cand_ids && codes && save_t && save_w && interests && d_scores && grad_w_proj && grad_codes &&
                      (!weighted || (w_target && save_z && grad_w_target)),
                  "train_bwd: null pointer");
  {
    const int rc0 = check_train_math(math, table_dtype, D, Dc, w_proj_bf16, weighted ? w_target_bf16 : w_proj_bf16);
    if (rc0) return rc0;
  }
  const bool tensor = math == MINER_MATH_TENSOR;
  const TrainWs w = train_ws(B, H, K, Dc, D, math);
  if (!workspace || workspace_bytes < w.total) {
    set_error("train_bwd: workspace too small (%zu bytes needed)", w.total);
    return MINER_ERR_WORKSPACE;
  }
  if (grad_table && (!w_proj || !table_grad_ws || table_grad_ws_bytes < miner_train_table_grad_workspace_bytes(B, H, Dc, D))) {
    set_error("train_bwd: the table gradient needs w_proj and %zu bytes of its own workspace", miner_train_table_grad_workspace_bytes(B, H, Dc, D));
    return MINER_ERR_WORKSPACE;
  }
  auto st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* dI = reinterpret_cast<float*>(ws + w.di);
  float* dI2 = reinterpret_cast<float*>(ws + w.di2);
  float* dZ = reinterpret_cast<float*>(ws + w.dz);
  float* dZ1 = reinterpret_cast<float*>(ws + w.dz1);
  float* WtT = reinterpret_cast<float*>(ws + w.wt_t);
  float* pWp = reinterpret_cast<float*>(ws + w.part_wp);
  float* pWt = reinterpret_cast<float*>(ws + w.part_wt);
  float* pCodes = reinterpret_cast<float*>(ws + w.part_codes);
  // 1. target-aware attention: dI (direct paths), dZ   -- or, for 'max' / 'mean', dI alone
  if (!weighted) {
    const size_t smem = sizeof(float) * (static_cast<size_t>(C) * D + static_cast<size_t>(C) * K);
    if (smem > 200 * 1024 || C > TT) {
      set_error("train_bwd: C=%lld D=%lld needs %zu bytes of shared memory", (long long)C, (long long)D, smem);
      return MINER_ERR_UNSUPPORTED;
    }
    MINER_CUDA_OK(cudaFuncSetAttribute(target_bwd_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    target_bwd_simple_kernel<<<static_cast<unsigned>(B), TT, smem, st>>>(table, table_dtype, n_rows, cand_ids, id_dtype, interests, d_scores,
                                                                        d_interests, (int)C, (int)K, (int)D, score_type, dI, grad_table);
    MINER_LAUNCH_OK("target_bwd_simple");
  } else {
    const size_t smem = sizeof(float) * (static_cast<size_t>(C) * D + 4 * C * K);
    if (smem > 200 * 1024 || C > TT) {
      set_error("train_bwd: C=%lld D=%lld needs %zu bytes of shared memory", (long long)C, (long long)D, smem);
      return MINER_ERR_UNSUPPORTED;
    }
    MINER_CUDA_OK(cudaFuncSetAttribute(target_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    target_bwd_kernel<<<static_cast<unsigned>(B), TT, smem, st>>>(table, table_dtype, n_rows, cand_ids, id_dtype, interests, save_z, d_scores,
                                                                 d_interests, (int)C, (int)K, (int)D, dI, dZ, grad_table);
    MINER_LAUNCH_OK("target_bwd");
  }
  // 2. Z = I Wt^T:  dI2 = dZ Wt  (as dZ (Wt^T)^T with the NT GEMM),  dWt = dZ^T I
  if (!weighted) {
    dI2 = nullptr;
  } else if (tensor) {
    dim3 tb(32, 8);
    const int64_t R = B * K;
    __nv_bfloat16* dz16 = reinterpret_cast<__nv_bfloat16*>(ws + w.dz_bf16);
    __nv_bfloat16* wtt16 = reinterpret_cast<__nv_bfloat16*>(ws + w.wtt_bf16);
    __nv_bfloat16* dzt = reinterpret_cast<__nv_bfloat16*>(ws + w.dzt);
    __nv_bfloat16* it = reinterpret_cast<__nv_bfloat16*>(ws + w.it);
    int rc = launch_cast_f32_to_bf16(dZ, dz16, R * D, st);
    if (rc) return rc;
    transpose_cast_kernel<<<dim3(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>((D + 31) / 32)), tb, 0, st>>>(w_target, wtt16, D, (int)D, D);
    MINER_LAUNCH_OK("transpose_cast(Wt)");
    rc = launch_tc_gemm(dz16, nullptr, id_dtype, 0, wtt16, dI2, nullptr, R, D, D, EPI_NONE, st);
    if (rc) return rc;
    if (tc_gemm_tn_supported(R, D, D)) {
      // dWt[o,i] = sum_r dZ[r,o] I[r,i]: both operands row-major as they are (dz16 from above, a bf16 copy of the interests), read
      // MN-major by the tensor cores -- no transposed copies
      rc = launch_cast_f32_to_bf16(interests, it, R * D, st);
      if (rc) return rc;
      rc = launch_tc_gemm_tn_splitk(dz16, it, w.s_wt > 1 ? pWt : grad_w_target, R, D, D, w.s_wt, st);
      if (rc) return rc;
    } else {
      const dim3 tg(static_cast<unsigned>((w.bk_pad + 31) / 32), static_cast<unsigned>((D + 31) / 32));
      transpose_cast_kernel<<<tg, tb, 0, st>>>(dZ, dzt, R, (int)D, w.bk_pad);
      MINER_LAUNCH_OK("transpose_cast(dZ)");
      transpose_cast_kernel<<<tg, tb, 0, st>>>(interests, it, R, (int)D, w.bk_pad);
      MINER_LAUNCH_OK("transpose_cast(I)");
      rc = launch_tc_gemm_splitk(dzt, nullptr, id_dtype, 0, it, w.s_wt > 1 ? pWt : grad_w_target, nullptr, D, D, w.bk_pad, EPI_NONE, w.s_wt, st);   // dWt[o,i] = sum_r dZ[r,o] I[r,i]
      if (rc) return rc;
    }
    if (w.s_wt > 1) {
      sum_partials_kernel<<<static_cast<int>((D * D + 255) / 256), 256, 0, st>>>(pWt, w.s_wt, D * D, grad_w_target);
      MINER_LAUNCH_OK("sum_partials(dWt)");
    }
  } else {
    dim3 tb(32, 8), tg(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>((D + 31) / 32));
    transpose_kernel<<<tg, tb, 0, st>>>(w_target, WtT, (int)D, (int)D);
    MINER_LAUNCH_OK("transpose");
    int rc = launch_sgemm_nt(dZ, MINER_F32, nullptr, id_dtype, 0, WtT, dI2, B * K, D, D, EPI_NONE, st);
    if (rc) return rc;
    const int64_t R = B * K;
    const int64_t rps = ((R + w.s_wt - 1) / w.s_wt + 15) / 16 * 16;
    dim3 grid(static_cast<unsigned>((D + 63) / 64), static_cast<unsigned>((D + 63) / 64), static_cast<unsigned>(w.s_wt));
    atb_kernel<<<grid, TT, 0, st>>>(dZ, (int)D, (int)D, interests, nullptr, MINER_F32, 0, nullptr, id_dtype, (int)D, (int)D, R, rps, pWt);
    MINER_LAUNCH_OK("atb(dWt)");
    sum_partials_kernel<<<static_cast<int>((D * D + 255) / 256), 256, 0, st>>>(pWt, w.s_wt, D * D, grad_w_target);
    MINER_LAUNCH_OK("sum_partials(dWt)");
  }
  // 3. poly attention: dZ1, dcodes
  {
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(K) * Dc + K * (POLY_BWD_DK + 12) + H * (POLY_BWD_DK + 12) + K * H + 4) + sizeof(int) * H;
    if (smem > 220 * 1024) {
      set_error("train_bwd: H=%lld K=%lld Dc=%lld need %zu bytes of shared memory", (long long)H, (long long)K, (long long)Dc, smem);
      return MINER_ERR_UNSUPPORTED;
    }
    MINER_CUDA_OK(cudaFuncSetAttribute(poly_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    poly_bwd_kernel<<<w.g_poly, TT, smem, st>>>(table, table_dtype, n_rows, his_ids, id_dtype, his_mask, codes, save_t, save_w, dI, dI2, B,
                                                (int)H, (int)K, (int)Dc, (int)D, dZ1, pCodes, d_bias_mean);
    MINER_LAUNCH_OK("poly_bwd");
    sum_partials_kernel<<<static_cast<int>((K * Dc + 255) / 256), 256, 0, st>>>(pCodes, w.g_poly, K * Dc, grad_codes);
    MINER_LAUNCH_OK("sum_partials(dcodes)");
  }
  // 3b. gradient of the history rows of the table: dE = w^T dI + dZ1 Wp, scattered by news id
  if (grad_table) {
    char* tg = static_cast<char*>(table_grad_ws);
    float* WpT = reinterpret_cast<float*>(tg);                                                       // (D, Dc)
    float* dE2 = reinterpret_cast<float*>(tg + align_up(sizeof(float) * D * Dc, 256));               // (B H, D)
    dim3 tb(32, 8), tgd(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>((Dc + 31) / 32));
    transpose_kernel<<<tgd, tb, 0, st>>>(w_proj, WpT, (int)Dc, (int)D);
    MINER_LAUNCH_OK("transpose(Wp)");
    int rc = launch_sgemm_nt(dZ1, MINER_F32, nullptr, id_dtype, 0, WpT, dE2, B * H, D, Dc, EPI_NONE, st);
    if (rc) return rc;
    const size_t smem = sizeof(float) * ((static_cast<size_t>(K) * H + 1) & ~static_cast<size_t>(1)) + sizeof(int64_t) * H;
    MINER_CUDA_OK(cudaFuncSetAttribute(hist_table_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    hist_table_grad_kernel<<<static_cast<unsigned>(B), TT, smem, st>>>(his_ids, id_dtype, n_rows, save_w, dI, dI2, dE2, (int)H, (int)K, (int)D,
                                                                      grad_table);
    MINER_LAUNCH_OK("hist_table_grad");
  }
  // 4. Z1 = E Wp^T:  dWp = dZ1^T E  (E gathered from the table)
  if (tensor) {
    dim3 tb(32, 8);
    const int64_t R = B * H;
    __nv_bfloat16* dz1t = reinterpret_cast<__nv_bfloat16*>(ws + w.dz1t);
    uint16_t* et = reinterpret_cast<uint16_t*>(ws + w.et);
    int rc = MINER_OK;
    if (tc_gemm_tn_supported(R, Dc, D)) {
      // dWp[c,d] = sum_r dZ1[r,c] E[r,d]: dZ1 as bf16 and the gathered history rows, both row-major, read MN-major (invalid ids: zero rows)
      rc = launch_cast_f32_to_bf16(dZ1, dz1t, R * Dc, st);
      if (rc) return rc;
      rc = launch_gather(table, n_rows, D, MINER_BF16, his_ids, R, id_dtype, et, nullptr, st);
      if (rc) return rc;
      rc = launch_tc_gemm_tn_splitk(dz1t, et, w.s_wp > 1 ? pWp : grad_w_proj, R, Dc, D, w.s_wp, st);
      if (rc) return rc;
    } else {
      transpose_cast_kernel<<<dim3(static_cast<unsigned>((w.bh_pad + 31) / 32), static_cast<unsigned>((Dc + 31) / 32)), tb, 0, st>>>(dZ1, dz1t, R, (int)Dc,
                                                                                                                                w.bh_pad);
      MINER_LAUNCH_OK("transpose_cast(dZ1)");
      gather_transpose_kernel<<<dim3(static_cast<unsigned>((w.bh_pad + 31) / 32), static_cast<unsigned>((D + 31) / 32)), tb, 0, st>>>(
          static_cast<const uint16_t*>(table), n_rows, his_ids, id_dtype, R, (int)D, w.bh_pad, et);
      MINER_LAUNCH_OK("gather_transpose(E)");
      rc = launch_tc_gemm_splitk(dz1t, nullptr, id_dtype, 0, et, w.s_wp > 1 ? pWp : grad_w_proj, nullptr, Dc, D, w.bh_pad, EPI_NONE, w.s_wp, st);   // dWp[c,d] = sum_r dZ1[r,c] E[r,d]
      if (rc) return rc;
    }
    if (w.s_wp > 1) {
      sum_partials_kernel<<<static_cast<int>((Dc * D + 255) / 256), 256, 0, st>>>(pWp, w.s_wp, Dc * D, grad_w_proj);
      MINER_LAUNCH_OK("sum_partials(dWp)");
    }
  } else {
    const int64_t R = B * H;
    const int64_t rps = ((R + w.s_wp - 1) / w.s_wp + 15) / 16 * 16;
    dim3 grid(static_cast<unsigned>((D + 63) / 64), static_cast<unsigned>((Dc + 63) / 64), static_cast<unsigned>(w.s_wp));
    atb_kernel<<<grid, TT, 0, st>>>(dZ1, (int)Dc, (int)Dc, nullptr, table, table_dtype, n_rows, his_ids, id_dtype, (int)D, (int)D, R, rps, pWp);
    MINER_LAUNCH_OK("atb(dWp)");
    sum_partials_kernel<<<static_cast<int>((Dc * D + 255) / 256), 256, 0, st>>>(pWp, w.s_wp, Dc * D, grad_w_proj);
    MINER_LAUNCH_OK("sum_partials(dWp)");
  }
  return MINER_OK;
}
