// (a7..a12) segmented per-impression ranking metrics.
// Replaces SlowEvaluator / FastEvaluator + BaseEvaluator.compute_scores (reference src/evaluation.py:36-175) and
// compute_mrr_score / compute_dcg_score / compute_ndcg_score / is_hit (evaluation.py:177-249) plus sklearn's
// roc_auc_score per impression (group_auc, evaluation.py:56-59).
//
// CSR layout: impression b owns candidates offsets[b] .. offsets[b+1].  Two kernels share the arithmetic (ImpressionTally):
//   rank_metrics_lane_kernel  (transforms none / sigmoid) a warp stages the candidates of 32 consecutive impressions, one lane ranks
//                             each of them; longer impressions fall back to the whole warp
//   rank_metrics_kernel       (softmax transform) one warp per impression, grid-stride
// The transformed scores are staged in shared memory, then the rank of a candidate is derived by counting -- pure comparisons, so
// the ranking is bit-exact given the scores:
//     gt  = #{j : p_j >  p_i}
//     rank_np = gt + #{j > i : p_j == p_i}   position under np.argsort(p)[::-1] (evaluation.py:188,208) with the tie rule
//                                            "stable ascending sort, reversed" (oracle/miner_oracle.py)
//     rank_py = gt + #{j < i : p_j == p_i}   position under python's stable sorted(reverse=True) (evaluation.py:247)
// group_auc is the tie-aware Mann-Whitney statistic; mrr, ndcg@k (gains 2^y-1, discounts log2(i+2), k=min(n,k)) and
// hit@k follow.  All metric arithmetic is float64 like the reference's numpy.  Sums over impressions skip NaN
// (np.nanmean, evaluation.py:59,65,72,80) and are reduced in a fixed order (per-block partials, then one block).
#include "common.cuh"

namespace miner {

constexpr int MT = 256;            // threads per block (8 warps)
#ifndef MINER_RM_MINB
#define MINER_RM_MINB 3
#endif
constexpr int RM_MINB = MINER_RM_MINB;   // resident blocks per SM the kernel is compiled for (register cap) and launched with
#ifndef MINER_RM_LANE_MINB
#define MINER_RM_LANE_MINB 2
#endif
constexpr int RM_LANE_MINB = MINER_RM_LANE_MINB;   // the same for the lane-per-impression kernel (32 impressions in flight per warp)
constexpr int MCAP = 1024;         // candidates per impression staged in shared memory per warp
constexpr int MAXK = 8;            // cut-offs
constexpr int LOG2_TAB = 256;      // ranks below this take log2 from a shared-memory table
constexpr int MAXM = 2 + 2 * MAXK;

struct MetricKs { int n_k; int k[MAXK]; };

__device__ __forceinline__ float transform_score(float s, int transform, float mx, float sum) {
  if (transform == 1) return 1.0f / (1.0f + expf(-s));        // torch.sigmoid, evaluation.py:165
  if (transform == 2) return expf(s - mx) / sum;              // softmax(dim=1), evaluation.py:109
  return s;
}

// Per-impression tallies and the float64 arithmetic on them, shared by the warp-per-impression and the lane-per-impression kernels:
// the same operations in the same (candidate) order, so both give the same bits.
template <int NK>
struct ImpressionTally {
  long long wins2;                                                    // 2 * (#neg below + 0.5 #neg tied), summed over positives
  int n_pos, n_neg, sum_y;
  double rr;
  double dcg[NK > 0 ? NK : 1], idcg[NK > 0 ? NK : 1];
  unsigned hit;                                                       // bit q: a positive inside the top ks[q]
  __device__ __forceinline__ void init() {
    wins2 = 0, n_pos = n_neg = sum_y = 0, rr = 0.0, hit = 0;
#pragma unroll
    for (int q = 0; q < NK; ++q) dcg[q] = idcg[q] = 0.0;
  }
  // one candidate with a non-zero label (a zero label adds exactly 0 to rr, dcg, idcg, hit and the AUC wins)
  __device__ __forceinline__ void add(int yi, int n, int gt, int eq_before, int eq_after, int neg_lt, int neg_eq, int y_gt, int y_eq_before,
                                      const MetricKs& ks, const double* log2_tab) {
    const int rank_np = gt + eq_after;                               // position under np.argsort(p)[::-1] (evaluation.py:188,208)
    const int rank_py = gt + eq_before;                              // position under python's stable sorted(reverse=True) (:247)
    const int rank_ideal = y_gt + y_eq_before;
    if (yi > 0) wins2 += 2 * neg_lt + neg_eq;
    // 2 ** y_true - 1 (evaluation.py:210); small non-negative labels are exact powers of two
    const double gain = (yi > 0 && yi < 31) ? static_cast<double>((1 << yi) - 1) : exp2(static_cast<double>(yi)) - 1.0;
    rr += static_cast<double>(yi) / static_cast<double>(rank_np + 1);  // evaluation.py:190
    const double l_np = rank_np + 2 < LOG2_TAB ? log2_tab[rank_np + 2] : log2(static_cast<double>(rank_np + 2));
    const double l_id = rank_ideal + 2 < LOG2_TAB ? log2_tab[rank_ideal + 2] : log2(static_cast<double>(rank_ideal + 2));
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const int kk = ks.k[q] < n ? ks.k[q] : n;                      // k = min(len, k), evaluation.py:207
      if (rank_np < kk) dcg[q] += gain / l_np;
      if (rank_ideal < kk) idcg[q] += gain / l_id;
      if (rank_py < ks.k[q] && yi > 0) hit |= 1u << q;               // evaluation.py:247-249
    }
  }
  __device__ __forceinline__ void finish(double (&vals)[2 + 2 * NK]) const {
    vals[0] = (n_pos > 0 && n_neg > 0) ? (0.5 * static_cast<double>(wins2)) / (static_cast<double>(n_pos) * static_cast<double>(n_neg))
                                       : NAN;                             // one class: sklearn raises -> NaN here
    vals[1] = rr / static_cast<double>(sum_y);                            // 0/0 -> NaN (no positives)
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      vals[2 + q] = dcg[q] / idcg[q];                                     // evaluation.py:231
      vals[2 + NK + q] = (hit >> q) & 1u ? 1.0 : 0.0;
    }
  }
};

// One impression ranked by the whole warp: its transformed scores / labels come from shared memory (ps, ys) when they were staged,
// else from global memory (sb, yb; transformed on every read).  Only candidates with a non-zero label contribute, so the pairwise
// counting runs once per such candidate -- about 2.4 per impression -- with all 32 lanes comparing one chunk of the others against it
// (ballot + popc): every tally is warp-uniform and exact, the fp64 sums run in candidate order in every lane.
template <int NK>
__device__ __forceinline__ void warp_impression(const float* ps, const int8_t* ys, const float* __restrict__ sb, const int8_t* __restrict__ yb,
                                                int n, int transform, float mx, float sum, const MetricKs& ks, const double* log2_tab,
                                                double (&vals)[2 + 2 * NK]) {
  const int lane = threadIdx.x & 31;
  const bool staged = ps != nullptr;
  ImpressionTally<NK> t;
  t.init();
  for (int base = 0; base < n; base += 32) {
    const int i_l = base + lane;
    const bool in_l = i_l < n;
    const float p_l = in_l ? (staged ? ps[i_l] : transform_score(sb[i_l], transform, mx, sum)) : 0.f;
    const int y_l = in_l ? (staged ? ys[i_l] : yb[i_l]) : 0;
    t.n_pos += __popc(__ballot_sync(0xffffffffu, in_l && y_l > 0));
    t.n_neg += __popc(__ballot_sync(0xffffffffu, in_l && y_l <= 0));
    t.sum_y += __reduce_add_sync(0xffffffffu, y_l);
    unsigned nz = __ballot_sync(0xffffffffu, in_l && y_l != 0);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const float pi = __shfl_sync(0xffffffffu, p_l, src);
      const int yi = __shfl_sync(0xffffffffu, y_l, src);
      const int i = base + src;
      int gt = 0, eq_before = 0, eq_after = 0, neg_lt = 0, neg_eq = 0, y_gt = 0, y_eq_before = 0;
      for (int jb = 0; jb < n; jb += 32) {
        const int j = jb + lane;
        const bool vj = j < n;
        const float pj = vj ? (staged ? ps[j] : transform_score(sb[j], transform, mx, sum)) : 0.f;
        const int yj = vj ? (staged ? ys[j] : yb[j]) : 0;
        const bool eq = vj && pj == pi;
        const bool negj = vj && yj <= 0;
        gt += __popc(__ballot_sync(0xffffffffu, vj && pj > pi));
        eq_before += __popc(__ballot_sync(0xffffffffu, eq && j < i));
        eq_after += __popc(__ballot_sync(0xffffffffu, eq && j > i));
        neg_lt += __popc(__ballot_sync(0xffffffffu, negj && pj < pi));
        neg_eq += __popc(__ballot_sync(0xffffffffu, negj && eq));
        y_gt += __popc(__ballot_sync(0xffffffffu, vj && yj > yi));
        y_eq_before += __popc(__ballot_sync(0xffffffffu, vj && yj == yi && j < i));
      }
      t.add(yi, n, gt, eq_before, eq_after, neg_lt, neg_eq, y_gt, y_eq_before, ks, log2_tab);
    }
  }
  t.finish(vals);
}

// One impression of at most LANE_MAX candidates ranked by ONE lane from shared memory: the same tallies by plain loops.  A warp then
// has 32 impressions in flight instead of one, which is what hides the latencies of this otherwise serial chain.
constexpr int LANE_MAX = 64;       // the lane keeps its non-zero-label candidates as one 64-bit map
template <int NK>
__device__ __forceinline__ void lane_impression(const float* ps, const int8_t* ys, int n, const MetricKs& ks, const double* log2_tab,
                                                double (&vals)[2 + 2 * NK]) {
  ImpressionTally<NK> t;
  t.init();
  unsigned long long nz = 0;
  for (int j = 0; j < n; ++j) {
    const int y = ys[j];
    t.n_pos += y > 0;
    t.n_neg += y <= 0;
    t.sum_y += y;
    if (y != 0) nz |= 1ull << j;
  }
  while (nz) {
    const int i = __ffsll(static_cast<long long>(nz)) - 1;
    nz &= nz - 1;
    const float pi = ps[i];
    const int yi = ys[i];
    int gt = 0, eq_before = 0, eq_after = 0, neg_lt = 0, neg_eq = 0, y_gt = 0, y_eq_before = 0;
    for (int j = 0; j < n; ++j) {
      const float pj = ps[j];
      const int yj = ys[j];
      const bool eq = pj == pi;
      const bool negj = yj <= 0;
      gt += pj > pi;
      eq_before += eq && j < i;
      eq_after += eq && j > i;
      neg_lt += negj && pj < pi;
      neg_eq += negj && eq;
      y_gt += yj > yi;
      y_eq_before += yj == yi && j < i;
    }
    t.add(yi, n, gt, eq_before, eq_after, neg_lt, neg_eq, y_gt, y_eq_before, ks, log2_tab);
  }
  t.finish(vals);
}

// per-block [sum, count] partials of the warps' accumulators, in a fixed order
template <int M>
__device__ __forceinline__ void block_partials_out(double (*red)[2 * MAXM], const double (&acc_sum)[M], const double (&acc_cnt)[M], int warp, int lane,
                                                   double* __restrict__ block_partials) {
  if (lane == 0)
    for (int m = 0; m < M; ++m) { red[warp][2 * m] = acc_sum[m]; red[warp][2 * m + 1] = acc_cnt[m]; }
  __syncthreads();
  if (threadIdx.x < 2 * M) {
    double s = 0.0;
    for (int w = 0; w < MT / 32; ++w) s += red[w][threadIdx.x];
    block_partials[static_cast<int64_t>(blockIdx.x) * 2 * MAXM + threadIdx.x] = s;
  }
}

// Warp per impression: every transform, any impression length.  The product path uses it for the softmax transform (FastEvaluator), whose
// per-impression max / sum are warp reductions.
template <int NK>
__global__ void __launch_bounds__(MT, RM_MINB) rank_metrics_kernel(const float* __restrict__ scores, const int8_t* __restrict__ labels,
                                                          const int64_t* __restrict__ offsets, int64_t B, int transform,
                                                          MetricKs ks, double* __restrict__ block_partials,
                                                          double* __restrict__ per_impression) {
  __shared__ float p_s[MT / 32][MCAP];
  __shared__ int8_t y_s[MT / 32][MCAP];
  __shared__ double red[MT / 32][2 * MAXM];
  __shared__ double log2_tab[LOG2_TAB];          // log2(r) of the small ranks: an fp64 log2 costs ~100x an fp64 shared-memory load here
  for (int i = threadIdx.x; i < LOG2_TAB; i += MT) log2_tab[i] = i > 0 ? log2(static_cast<double>(i)) : 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int M = 2 + 2 * NK;
  float* ps = p_s[warp];
  int8_t* ys = y_s[warp];
  double acc_sum[M], acc_cnt[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc_sum[m] = 0.0, acc_cnt[m] = 0.0;

  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (MT / 32);
  for (int64_t b = static_cast<int64_t>(blockIdx.x) * (MT / 32) + warp; b < B; b += warps_total) {
    const int64_t o0 = offsets[b];
    const int n = static_cast<int>(offsets[b + 1] - o0);
    const float* sb = scores + o0;
    const int8_t* yb = labels + o0;
    const bool staged = n <= MCAP;
    float mx = 0.f, sum = 1.f;
    if (transform == 2) {
      mx = -INFINITY;
      for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sb[j]);
      mx = warp_max(mx);
      sum = 0.f;
      for (int j = lane; j < n; j += 32) sum += expf(sb[j] - mx);
      sum = warp_sum(sum);
    }
    if (staged) {
      for (int j = lane; j < n; j += 32) {
        ps[j] = transform_score(sb[j], transform, mx, sum);
        ys[j] = yb[j];
      }
    }
    __syncwarp();
    double vals[M];
    warp_impression<NK>(staged ? ps : nullptr, ys, sb, yb, n, transform, mx, sum, ks, log2_tab, vals);
    if (lane == 0) {
      for (int m = 0; m < M; ++m) {
        if (per_impression) per_impression[b * M + m] = vals[m];
        if (!isnan(vals[m])) { acc_sum[m] += vals[m]; acc_cnt[m] += 1.0; }
      }
    }
    __syncwarp();
  }
  block_partials_out<M>(red, acc_sum, acc_cnt, warp, lane, block_partials);
}

// Lane per impression (transforms none / sigmoid): a warp takes 32 consecutive impressions, stages their candidates -- one contiguous
// CSR range -- with coalesced loads and one transform per candidate, then every lane ranks its own impression from shared memory.
// Impressions longer than LANE_MAX, and groups whose candidates exceed the staging capacity, go through warp_impression one by one.
template <int NK>
__global__ void __launch_bounds__(MT, RM_LANE_MINB) rank_metrics_lane_kernel(const float* __restrict__ scores, const int8_t* __restrict__ labels,
                                                               const int64_t* __restrict__ offsets, int64_t B, int transform,
                                                               MetricKs ks, double* __restrict__ block_partials,
                                                               double* __restrict__ per_impression) {
  __shared__ float p_s[MT / 32][MCAP];
  __shared__ int8_t y_s[MT / 32][MCAP];
  __shared__ double red[MT / 32][2 * MAXM];
  __shared__ double log2_tab[LOG2_TAB];
  for (int i = threadIdx.x; i < LOG2_TAB; i += MT) log2_tab[i] = i > 0 ? log2(static_cast<double>(i)) : 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int M = 2 + 2 * NK;
  float* ps = p_s[warp];
  int8_t* ys = y_s[warp];
  double acc_sum[M], acc_cnt[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc_sum[m] = 0.0, acc_cnt[m] = 0.0;

  const int64_t groups = (B + 31) / 32;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (MT / 32);
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * (MT / 32) + warp; g < groups; g += warps_total) {
    const int64_t b = g * 32 + lane;
    const bool have = b < B;
    const int64_t o0 = offsets[have ? b : B];
    const int64_t o1 = offsets[have ? b + 1 : B];
    const int64_t r0 = __shfl_sync(0xffffffffu, o0, 0), r1 = __shfl_sync(0xffffffffu, o1, 31);
    const int n = static_cast<int>(o1 - o0);
    const bool staged = r1 - r0 <= MCAP;                               // warp-uniform
    if (staged) {
      const int len = static_cast<int>(r1 - r0);
      for (int j = lane; j < len; j += 32) {
        ps[j] = transform_score(scores[r0 + j], transform, 0.f, 1.f);
        ys[j] = labels[r0 + j];
      }
    }
    __syncwarp();
    double vals[M];
    const bool mine = have && staged && n <= LANE_MAX;
    if (mine) lane_impression<NK>(ps + (o0 - r0), ys + (o0 - r0), n, ks, log2_tab, vals);
    __syncwarp();
    unsigned rest = __ballot_sync(0xffffffffu, have && !mine);
    while (rest) {
      const int src = __ffs(rest) - 1;
      rest &= rest - 1;
      const int64_t so = __shfl_sync(0xffffffffu, o0, src);
      const int sn = __shfl_sync(0xffffffffu, n, src);
      double v[M];
      warp_impression<NK>(staged ? ps + (so - r0) : nullptr, ys + (staged ? so - r0 : 0), scores + so, labels + so, sn, transform, 0.f, 1.f, ks,
                          log2_tab, v);
      if (lane == src) {
#pragma unroll
        for (int m = 0; m < M; ++m) vals[m] = v[m];
      }
    }
    if (have) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (per_impression) per_impression[b * M + m] = vals[m];
        if (!isnan(vals[m])) { acc_sum[m] += vals[m]; acc_cnt[m] += 1.0; }
      }
    }
    __syncwarp();                                                     // the next group's staging overwrites ps / ys
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {                                       // lanes -> lane 0, a fixed tree
    acc_sum[m] = warp_sum(acc_sum[m]);
    acc_cnt[m] = warp_sum(acc_cnt[m]);
  }
  block_partials_out<M>(red, acc_sum, acc_cnt, warp, lane, block_partials);
}

// out[c] = sum over blocks of column c: one warp per column, lanes stride the blocks, a fixed tree at the end
__global__ void rank_metrics_finalize(const double* __restrict__ block_partials, int n_blocks, int M, double* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < 2 * M; c += blockDim.x >> 5) {
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += block_partials[static_cast<int64_t>(b) * 2 * MAXM + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
  }
}

static int metrics_lane_grid(int64_t B) {
  int64_t blocks = ((B + 31) / 32 + MT / 32 - 1) / (MT / 32);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 2 * RM_LANE_MINB;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

static int metrics_grid(int64_t B) {
  int64_t blocks = (B + MT / 32 - 1) / (MT / 32);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 2 * RM_MINB;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace miner

extern "C" size_t miner_rank_metrics_workspace_bytes(int64_t B, int n_k) {
  (void)n_k;
  return sizeof(double) * 2 * miner::MAXM * static_cast<size_t>(miner::metrics_grid(B));
}

extern "C" int miner_rank_metrics(const float* scores, const int8_t* labels, const int64_t* offsets, int64_t B, int transform,
                                  const int* ks, int n_k, double* out_partials, double* out_per_impression, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  using namespace miner;
  MINER_CHECK_ARG(offsets && out_partials, "rank_metrics: null pointer");
  MINER_CHECK_ARG(B == 0 || (scores && labels), "rank_metrics: null scores/labels");
  MINER_CHECK_ARG(n_k >= 0 && n_k <= MAXK, "rank_metrics: at most %d cut-offs", MAXK);
  MINER_CHECK_ARG(transform >= 0 && transform <= 2, "rank_metrics: transform must be 0 (none), 1 (sigmoid) or 2 (softmax)");
  MINER_CHECK_ARG(n_k == 0 || ks, "rank_metrics: null ks");
  const int grid = transform == 2 ? metrics_grid(B) : metrics_lane_grid(B);   // never more than the workspace query's metrics_grid(B)
  if (workspace_bytes < sizeof(double) * 2 * MAXM * static_cast<size_t>(grid) || !workspace) {
    set_error("rank_metrics: workspace too small (%zu bytes needed)", sizeof(double) * 2 * MAXM * static_cast<size_t>(grid));
    return MINER_ERR_WORKSPACE;
  }
  MetricKs mk;
  mk.n_k = n_k;
  for (int i = 0; i < MAXK; ++i) mk.k[i] = i < n_k ? ks[i] : 0;
  auto st = static_cast<cudaStream_t>(stream);
  double* parts = static_cast<double*>(workspace);
  // softmax needs per-impression statistics: warp per impression; the other transforms rank 32 impressions per warp at a time
#define MINER_RM_CASE(NKV)                                                                                                      \
  case NKV:                                                                                                                     \
    if (transform == 2) rank_metrics_kernel<NKV><<<grid, MT, 0, st>>>(scores, labels, offsets, B, transform, mk, parts, out_per_impression);      \
    else rank_metrics_lane_kernel<NKV><<<grid, MT, 0, st>>>(scores, labels, offsets, B, transform, mk, parts, out_per_impression);                \
    break;
  switch (n_k) {       // the cut-off count is a template parameter: fixed-size register arrays, no predicated fp64 reductions
    MINER_RM_CASE(0) MINER_RM_CASE(1) MINER_RM_CASE(2) MINER_RM_CASE(3) MINER_RM_CASE(4)
    MINER_RM_CASE(5) MINER_RM_CASE(6) MINER_RM_CASE(7) MINER_RM_CASE(8)
  }
#undef MINER_RM_CASE
  MINER_LAUNCH_OK("rank_metrics");
  rank_metrics_finalize<<<1, 512, 0, st>>>(static_cast<const double*>(workspace), grid, 2 + 2 * n_k, out_partials);
  MINER_LAUNCH_OK("rank_metrics_finalize");
  return MINER_OK;
}
