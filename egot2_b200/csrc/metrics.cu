// metrics.cu — batched, on-device versions of the metric post-processing that directly follows the translators:
//   egot2_segment_softmax_mean   TTM: one score per video segment = softmax(mean over the segment's minibatch rows of the logits)
//                                (HHI/utils/ttm/utils.py:71-80, PostProcessor._merge_output)
//   egot2_topk_correct           LTA / AR: number of rows whose label is among the k largest predictions, for several k
//                                (HOI/evaluation/lta/lta_metrics.py:39-73, topks_correct / topk_errors)
//   egot2_edit_distance_prefix   LTA: Levenshtein distance (the `editdistance` package the reference calls) between each of the K
//                                sampled action sequences and the label sequence, for EVERY prefix length z = 1..Z at once (the
//                                diagonal of one dynamic-programming table), minimised over K, summed over the clips
//                                (lta_metrics.py:87-110, edit_distance / AUED)
// Integer results (counts, distances, their sums) are exact; the reference loops clip by clip on the host with a device
// synchronisation per `.item()`.
#define EGOT2_FILE_ID 15
#include "ops.h"

namespace egot2 {
namespace {

constexpr int kMaxCls = 8;
constexpr int kMaxZ = 64;

// one warp per segment; rows [off[s], off[s+1]) of logits (rows, n_cls); lane-strided partial sums folded in a fixed order
__global__ void __launch_bounds__(256) seg_softmax_mean_kernel(int n_seg, int n_cls, const float* __restrict__ logits,
                                                               const int32_t* __restrict__ seg_off, float* __restrict__ out) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_seg) return;
  const int r0 = seg_off[s], r1 = seg_off[s + 1];
  float acc[kMaxCls];
#pragma unroll
  for (int c = 0; c < kMaxCls; ++c) acc[c] = 0.f;
  for (int r = r0 + lane; r < r1; r += 32)
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) if (c < n_cls) acc[c] += logits[(size_t)r * n_cls + c];
#pragma unroll
  for (int c = 0; c < kMaxCls; ++c) acc[c] = warp_sum(acc[c]);
  if (lane == 0) {
    const float inv = r1 > r0 ? 1.f / (float)(r1 - r0) : 0.f;
    float mx = -INFINITY;
    for (int c = 0; c < n_cls; ++c) { acc[c] *= inv; mx = fmaxf(mx, acc[c]); }
    float den = 0.f;
    for (int c = 0; c < n_cls; ++c) { acc[c] = expf(acc[c] - mx); den += acc[c]; }
    for (int c = 0; c < n_cls; ++c) out[(size_t)s * n_cls + c] = acc[c] / den;
  }
}

// one warp per row: out = softmax(in) over n columns (fp32), the eval-mode activation of the LTA MultiTaskHead
// (HOI/models/lta/head_helper.py:284-286) and lossAV's predScore (HHI/tasks/asd/loss.py:24)
__global__ void __launch_bounds__(256) row_softmax_kernel(long long rows, int n, const float* __restrict__ in, float* __restrict__ out) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = in + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float den = 0.f;
  for (int j = lane; j < n; j += 32) den += expf(x[j] - mx);
  den = warp_sum(den);
  const float inv = 1.f / den;
  for (int j = lane; j < n; j += 32) out[row * n + j] = expf(x[j] - mx) * inv;
}

struct Ks { int n; int k[8]; };

// one warp per row: rank of the label = #(larger predictions) + #(equal predictions at a smaller index)
__global__ void __launch_bounds__(256) topk_correct_kernel(int N, int C, const float* __restrict__ preds,
                                                           const int64_t* __restrict__ labels, Ks ks,
                                                           unsigned long long* __restrict__ correct) {
  EGOT2_PDL_ENTER();
  __shared__ unsigned int s_cnt[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < N) {
    const float* p = preds + (size_t)row * C;
    const long long lab = labels[row];
    int rank = C;                               // an out-of-range label is never correct
    if (lab >= 0 && lab < C) {
      const float v = p[lab];
      int cnt = 0;
      for (int j = lane; j < C; j += 32) {
        const float x = p[j];
        cnt += (x > v || (x == v && j < lab)) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      rank = cnt;
    }
    if (lane == 0)
      for (int i = 0; i < ks.n; ++i) if (rank < ks.k[i]) atomicAdd(&s_cnt[i], 1u);
  }
  __syncthreads();
  if (threadIdx.x < ks.n && s_cnt[threadIdx.x]) atomicAdd(correct + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// one thread per clip: for every sampled sequence k the Levenshtein table of preds[n,:,k] against labels[n,:], row by row;
// D[i][i] is the distance of the two length-i prefixes.  min over k -> min_dist[n, i-1]; sums over the clips -> sum_min[i-1].
__global__ void __launch_bounds__(128) edit_distance_kernel(int N, int Z, int K, const int64_t* __restrict__ preds,
                                                            const int64_t* __restrict__ labels, int32_t* __restrict__ min_dist,
                                                            unsigned long long* __restrict__ sum_min) {
  EGOT2_PDL_ENTER();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int best[kMaxZ];
  for (int z = 0; z < Z; ++z) best[z] = 0x7fffffff;
  int64_t lab[kMaxZ];
  for (int z = 0; z < Z; ++z) lab[z] = labels[(size_t)n * Z + z];
  for (int k = 0; k < K; ++k) {
    int row[kMaxZ + 1];
    for (int j = 0; j <= Z; ++j) row[j] = j;            // D[0][j]
    for (int i = 1; i <= Z; ++i) {
      const int64_t a = preds[((size_t)n * Z + (i - 1)) * K + k];
      int diag = row[0];                                  // D[i-1][0]
      row[0] = i;
      for (int j = 1; j <= Z; ++j) {
        const int up = row[j];                            // D[i-1][j]
        const int sub = diag + (a == lab[j - 1] ? 0 : 1);
        const int v = min(min(up + 1, row[j - 1] + 1), sub);
        diag = up;
        row[j] = v;
      }
      best[i - 1] = min(best[i - 1], row[i]);
    }
  }
  for (int z = 0; z < Z; ++z) {
    if (min_dist) min_dist[(size_t)n * Z + z] = best[z];
    atomicAdd(sum_min + z, (unsigned long long)best[z]);
  }
}

}  // namespace
}  // namespace egot2

using namespace egot2;

extern "C" int egot2_segment_softmax_mean(int32_t n_seg, int32_t n_cls, const float* logits, const int32_t* seg_offsets,
                                          float* out, void* stream) {
  if (n_seg == 0) return 0;
  EGOT2_CHECK(n_seg > 0 && n_cls >= 1 && n_cls <= kMaxCls && logits && seg_offsets && out, "segment_softmax_mean: bad arguments (n_cls <= %d)", kMaxCls);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, "segment_softmax_mean segs%d", n_seg);
  launch(seg_softmax_mean_kernel, dim3((n_seg + 7) / 8), dim3(256), 0, st, n_seg, n_cls, logits, seg_offsets, out);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_row_softmax(int64_t rows, int32_t n, const float* in, float* out, void* stream) {
  if (rows == 0) return 0;
  EGOT2_CHECK(rows > 0 && n >= 1 && in && out, "row_softmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, "row_softmax rows%lld n%d", (long long)rows, n);
  launch(row_softmax_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, (long long)rows, n, in, out);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_topk_correct(int32_t N, int32_t C, const float* preds, const int64_t* labels, int32_t n_k,
                                  const int32_t* ks_host, int64_t* correct, void* stream) {
  EGOT2_CHECK(N >= 0 && C >= 1 && n_k >= 1 && n_k <= 8 && ks_host && correct, "topk_correct: bad arguments (1 <= n_k <= 8)");
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_CUDA(cudaMemsetAsync(correct, 0, sizeof(int64_t) * n_k, st));
  if (N == 0) return 0;
  EGOT2_CHECK(preds && labels, "topk_correct: preds / labels required");
  Ks ks;
  ks.n = n_k;
  for (int i = 0; i < n_k; ++i) ks.k[i] = ks_host[i];
  ProfScope prof(st, "topk_correct N%d C%d", N, C);
  launch(topk_correct_kernel, dim3((N + 7) / 8), dim3(256), 0, st, N, C, preds, labels, ks, (unsigned long long*)correct);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_edit_distance_prefix(int32_t N, int32_t Z, int32_t K, const int64_t* preds, const int64_t* labels,
                                          int32_t* min_dist, int64_t* sum_min, void* stream) {
  EGOT2_CHECK(N >= 0 && Z >= 1 && Z <= kMaxZ && K >= 1 && sum_min, "edit_distance_prefix: bad arguments (1 <= Z <= %d)", kMaxZ);
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_CUDA(cudaMemsetAsync(sum_min, 0, sizeof(int64_t) * Z, st));
  if (N == 0) return 0;
  EGOT2_CHECK(preds && labels, "edit_distance_prefix: preds / labels required");
  ProfScope prof(st, "edit_distance_prefix N%d Z%d K%d", N, Z, K);
  launch(edit_distance_kernel, dim3((N + 127) / 128), dim3(128), 0, st, N, Z, K, preds, labels, min_dist, (unsigned long long*)sum_min);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
