// Global AUC over ALL candidates (reference src/evaluation.py:53-55: sklearn roc_auc_score on the flattened targets / predictions).
//
// roc_auc_score is the tie-aware Mann-Whitney statistic  U / (P N),  U = #{(p, n) : s_p > s_n} + 1/2 #{(p, n) : s_p == s_n}
// over positive / negative candidates.  It is a rank statistic over the whole evaluation set, so unlike the per-impression metrics
// it does not decompose into per-rank means.  Exact integer formulation, no floating-point accumulation anywhere:
//   1. auc_split:   every candidate's probability (the same transform rank_metrics applies: sigmoid of the logit for SlowEvaluator,
//                   evaluation.py:165) becomes an order-preserving 32-bit key; positives and negatives are compacted into two arrays
//   2. radix sort of the POSITIVE keys (LSD, 4 passes of 8 bits: per-warp-segment digit histograms, one exclusive scan, stable scatter
//      ranked with match.any).  With several ranks the positive keys of all ranks are all-gathered first (the minority class:
//      ~12 % of the candidates) and every rank sorts the same global array
//   3. auc_count:   every local negative binary-searches the sorted positives:  2 U += 2 #{p > n} + #{p == n}   (uint64 atomics)
// and  auc = 2U / (2 P N)  with P, N, 2U summed over ranks (exact int64 all-reduce).
#include "common.cuh"

namespace miner {

namespace {

constexpr int RS_WARPS = 8, RS_SEG = 512;      // keys per warp segment (16 per lane), warps per block

// float -> uint32 whose unsigned order is the float order; -0 and +0 compare equal as floats, so they share a key
__device__ __forceinline__ uint32_t order_key(float x) {
  uint32_t b = __float_as_uint(x);
  if (b == 0x80000000u) b = 0u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one warp per impression when the transform is the per-impression softmax (FastEvaluator, evaluation.py:109), else elementwise
__global__ void __launch_bounds__(256) auc_split_kernel(const float* __restrict__ scores, const int8_t* __restrict__ labels,
                                                      const int64_t* __restrict__ offsets, int64_t B, int64_t T, int transform,
                                                      uint32_t* __restrict__ pos_keys, uint32_t* __restrict__ neg_keys,
                                                      unsigned long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  auto emit = [&](bool valid, float p, int y) {
    // warp-aggregated append: the order inside the two arrays is irrelevant (one is sorted, the other only searched)
    const unsigned mp = __ballot_sync(0xffffffffu, valid && y > 0), mn = __ballot_sync(0xffffffffu, valid && y <= 0);
    unsigned long long bp = 0, bn = 0;
    if (lane == 0) {
      if (mp) bp = atomicAdd(&counts[0], static_cast<unsigned long long>(__popc(mp)));
      if (mn) bn = atomicAdd(&counts[1], static_cast<unsigned long long>(__popc(mn)));
    }
    bp = __shfl_sync(0xffffffffu, bp, 0);
    bn = __shfl_sync(0xffffffffu, bn, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (valid && y > 0) pos_keys[bp + __popc(mp & lt)] = order_key(p);
    if (valid && y <= 0) neg_keys[bn + __popc(mn & lt)] = order_key(p);
  };
  if (transform == 2) {
    const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t b = w0; b < B; b += nw) {
      const int64_t o0 = offsets[b];
      const int n = static_cast<int>(offsets[b + 1] - o0);
      float mx = -INFINITY;
      for (int j = lane; j < n; j += 32) mx = fmaxf(mx, scores[o0 + j]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < n; j += 32) sum += expf(scores[o0 + j] - mx);
      sum = warp_sum(sum);
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool v = j < n;
        emit(v, v ? expf(scores[o0 + j] - mx) / sum : 0.f, v ? labels[o0 + j] : 0);
      }
    }
  } else {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t_end = (T + 31) / 32 * 32;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < t_end; i += stride) {
      const bool v = i < T;
      const float s = v ? scores[i] : 0.f;
      emit(v, transform == 1 ? 1.0f / (1.0f + expf(-s)) : s, v ? labels[i] : 0);     // same arithmetic as rank_metrics' transform_score
    }
  }
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                               uint32_t* __restrict__ counts, int64_t nseg) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seg = static_cast<int64_t>(blockIdx.x) * RS_WARPS + warp;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  if (seg < nseg) {
    for (int it = 0; it < RS_SEG / 32; ++it) {
      const int64_t i = seg * RS_SEG + it * 32 + lane;
      if (i < n) atomicAdd(&cnt[warp][(keys[i] >> shift) & 255u], 1u);
    }
    __syncwarp();
    for (int d = lane; d < 256; d += 32) counts[static_cast<int64_t>(d) * nseg + seg] = cnt[warp][d];
  }
}

// exclusive scan of `total` counters in place, one block (digit-major layout: all segments of digit 0, then digit 1, ...)
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t* __restrict__ counts, int64_t total) {
  __shared__ uint32_t warp_excl[32];
  __shared__ uint32_t block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;                                       // the same running total in every thread
  for (int64_t base = 0; base < total; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < total ? counts[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_excl[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const uint32_t t = warp_excl[lane];
      uint32_t s = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_excl[lane] = s - t;
      if (lane == 31) block_total = s;
    }
    __syncthreads();
    if (i < total) counts[i] = carry + warp_excl[warp] + (x - v);
    carry += block_total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_scatter_kernel(const uint32_t* __restrict__ keys, uint32_t* __restrict__ out, int64_t n,
                                                                  int shift, const uint32_t* __restrict__ offsets, int64_t nseg) {
  __shared__ uint32_t off[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seg = static_cast<int64_t>(blockIdx.x) * RS_WARPS + warp;
  if (seg >= nseg) return;
  for (int d = lane; d < 256; d += 32) off[warp][d] = offsets[static_cast<int64_t>(d) * nseg + seg];
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  for (int it = 0; it < RS_SEG / 32; ++it) {               // in order: the sort is stable
    const int64_t i = seg * RS_SEG + it * 32 + lane;
    const bool v = i < n;
    const uint32_t key = v ? keys[i] : 0u;
    const uint32_t d = v ? ((key >> shift) & 255u) : (256u + lane);      // lanes past the end match nobody
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(m & lt);
    if (v) out[off[warp][d] + rank] = key;
    __syncwarp();
    if (v && rank == 0) off[warp][d] += __popc(m);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) auc_count_kernel(const uint32_t* __restrict__ pos_sorted, int64_t P, const uint32_t* __restrict__ neg_keys,
                                                      int64_t N, unsigned long long* __restrict__ out_u2) {
  unsigned long long acc = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride) {
    const uint32_t key = neg_keys[i];
    int64_t lo = 0, hi = P;                                 // lower bound: first positive >= key
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (pos_sorted[mid] < key) lo = mid + 1; else hi = mid;
    }
    const int64_t lb = lo;
    hi = P;                                                 // upper bound: first positive > key
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (pos_sorted[mid] <= key) lo = mid + 1; else hi = mid;
    }
    acc += 2ull * static_cast<unsigned long long>(P - lo) + static_cast<unsigned long long>(lo - lb);
  }
  // block reduction, then one atomic per block (integers: the result does not depend on the order)
  __shared__ unsigned long long red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    if (t) atomicAdd(out_u2, t);
  }
}

}  // namespace

size_t sort_u32_ws_bytes(int64_t n) {
  const int64_t nseg = (n + RS_SEG - 1) / RS_SEG;
  return align_up(sizeof(uint32_t) * static_cast<size_t>(n > 0 ? n : 1), 256) + align_up(sizeof(uint32_t) * 256 * static_cast<size_t>(nseg > 0 ? nseg : 1), 256);
}

// LSD radix sort of n uint32 keys in place (ascending); workspace = a second key buffer + the digit counters
int launch_sort_u32(uint32_t* keys, int64_t n, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 1) return MINER_OK;
  if (!workspace || workspace_bytes < sort_u32_ws_bytes(n)) {
    set_error("sort_u32: workspace too small (%zu bytes needed)", sort_u32_ws_bytes(n));
    return MINER_ERR_WORKSPACE;
  }
  if (n >= (1ll << 32)) {
    set_error("sort_u32: %lld keys (at most 2^32 - 1)", (long long)n);
    return MINER_ERR_UNSUPPORTED;
  }
  const int64_t nseg = (n + RS_SEG - 1) / RS_SEG;
  uint32_t* tmp = static_cast<uint32_t*>(workspace);
  uint32_t* counts = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + align_up(sizeof(uint32_t) * static_cast<size_t>(n), 256));
  const unsigned blocks = static_cast<unsigned>((nseg + RS_WARPS - 1) / RS_WARPS);
  uint32_t* src = keys;
  uint32_t* dst = tmp;
  for (int pass = 0; pass < 4; ++pass) {
    rs_hist_kernel<<<blocks, RS_WARPS * 32, 0, stream>>>(src, n, 8 * pass, counts, nseg);
    MINER_LAUNCH_OK("rs_hist_kernel");
    rs_scan_kernel<<<1, 1024, 0, stream>>>(counts, 256 * nseg);
    MINER_LAUNCH_OK("rs_scan_kernel");
    rs_scatter_kernel<<<blocks, RS_WARPS * 32, 0, stream>>>(src, dst, n, 8 * pass, counts, nseg);
    MINER_LAUNCH_OK("rs_scatter_kernel");
    uint32_t* t = src; src = dst; dst = t;
  }
  return MINER_OK;                                           // four passes: the result is back in `keys`
}

int launch_auc_split(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B, int64_t T, int transform,
                     uint32_t* pos_keys, uint32_t* neg_keys, unsigned long long* counts, cudaStream_t stream) {
  MINER_CUDA_OK(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), stream));
  if (T == 0) return MINER_OK;
  const int64_t work = transform == 2 ? (B + 7) / 8 : (T + 255) / 256;
  const int grid = static_cast<int>(work < 8ll * sm_count() ? (work > 0 ? work : 1) : 8ll * sm_count());
  auc_split_kernel<<<grid, 256, 0, stream>>>(scores, labels, offsets, B, T, transform, pos_keys, neg_keys, counts);
  MINER_LAUNCH_OK("auc_split_kernel");
  return MINER_OK;
}

int launch_auc_count(const uint32_t* pos_sorted, int64_t P, const uint32_t* neg_keys, int64_t N, unsigned long long* out_u2, cudaStream_t stream) {
  MINER_CUDA_OK(cudaMemsetAsync(out_u2, 0, sizeof(unsigned long long), stream));
  if (P == 0 || N == 0) return MINER_OK;
  const int64_t blocks = (N + 255) / 256;
  const int grid = static_cast<int>(blocks < 8ll * sm_count() ? blocks : 8ll * sm_count());
  auc_count_kernel<<<grid, 256, 0, stream>>>(pos_sorted, P, neg_keys, N, out_u2);
  MINER_LAUNCH_OK("auc_count_kernel");
  return MINER_OK;
}

}  // namespace miner
