// loss.cu — task losses on fp32 logits, forward + d(loss)/d(logits).
//   EGOT2_LOSS_CE           nn.CrossEntropyLoss(weight=w) (TTM [0.266,0.734]; ASD lossAV [1,4]; OSCC unweighted)
//   EGOT2_LOSS_BCE_SIGMOID  nn.BCELoss()(sigmoid(z), onehot) (PNR keyframe localisation)
//   EGOT2_LOSS_CE_GROUPS    sum over Z sub-rows x G class groups of mean-CE (LTA verbs/nouns x 20 steps)
// A "segment" is one softmax: (row, sub_row, group).  One warp per segment; a single-CTA tree
// reduces the per-segment (weighted nll, weight) pairs to the scalar loss, so the result is
// deterministic (no atomics).  argmax (first maximal index, as torch.argmax) is emitted per segment.
#include <math.h>

#define EGOT2_FILE_ID 2
#include "ops.h"

namespace egot2 {

namespace {

struct SegGeom {
  int n_out, sub_rows, n_groups, per_sub;
  int gsize[EGOT2_MAX_GROUPS], goff[EGOT2_MAX_GROUPS];
};

inline SegGeom geom(const egot2_head_desc& d) {
  SegGeom g{};
  g.n_out = d.n_out;
  if (d.loss == EGOT2_LOSS_CE_GROUPS) {
    g.sub_rows = d.sub_rows; g.n_groups = d.n_groups;
    int off = 0;
    for (int i = 0; i < d.n_groups; ++i) { g.gsize[i] = d.group_size[i]; g.goff[i] = off; off += d.group_size[i]; }
    g.per_sub = off;
  } else {
    g.sub_rows = 1; g.n_groups = 1; g.gsize[0] = d.n_out; g.goff[0] = 0; g.per_sub = d.n_out;
  }
  return g;
}

__device__ __forceinline__ void seg_locate(const SegGeom& g, long long seg, long long& row, int& off, int& n) {
  const int per_row = g.sub_rows * g.n_groups;
  row = seg / per_row;
  const int r = (int)(seg % per_row), z = r / g.n_groups, gi = r % g.n_groups;
  off = z * g.per_sub + g.goff[gi];
  n = g.gsize[gi];
}

// softmax-CE per segment
__global__ void ce_fwd_kernel(SegGeom g, long long segs, const float* __restrict__ logits,
                              const int64_t* __restrict__ labels, const float* __restrict__ cw,
                              float* __restrict__ seg_loss, int32_t* __restrict__ argmax) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const long long seg = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (seg >= segs) return;
  long long row; int off, n;
  seg_locate(g, seg, row, off, n);
  const float* z = logits + row * g.n_out + off;
  float mx = -INFINITY; int am = 0x7fffffff;
  for (int j = lane; j < n; j += 32) { const float v = z[j]; if (v > mx) { mx = v; am = j; } }
  // warp arg-max with first-index tie break
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(z[j] - mx);
  s = warp_sum(s);
  if (lane == 0) {
    const int y = (int)labels[seg];
    const float w = cw ? cw[y] : 1.f;
    const float nll = (mx + logf(s)) - z[y];
    seg_loss[2 * seg] = w * nll;
    seg_loss[2 * seg + 1] = w;
    if (argmax) argmax[seg] = am;
  }
}

__global__ void ce_bwd_kernel(SegGeom g, long long segs, const float* __restrict__ logits,
                              const int64_t* __restrict__ labels, const float* __restrict__ cw,
                              const float* __restrict__ loss2, float scale, float* __restrict__ dlogits) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const long long seg = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (seg >= segs) return;
  long long row; int off, n;
  seg_locate(g, seg, row, off, n);
  const float* z = logits + row * g.n_out + off;
  float* dz = dlogits + row * g.n_out + off;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, z[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += expf(z[j] - mx);
  s = warp_sum(s);
  const int y = (int)labels[seg];
  const float w = (cw ? cw[y] : 1.f) * scale / loss2[1];
  const float inv = 1.f / s;
  for (int j = lane; j < n; j += 32) dz[j] = w * (expf(z[j] - mx) * inv - (j == y ? 1.f : 0.f));
}

// sigmoid + BCE against a one-hot row; one warp per row
__global__ void bce_fwd_kernel(int rows, int n, const float* __restrict__ logits, const int64_t* __restrict__ labels,
                               float* __restrict__ seg_loss, int32_t* __restrict__ argmax) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* z = logits + row * n;
  const int y = (int)labels[row];
  float acc = 0.f, mx = -INFINITY; int am = 0x7fffffff;
  for (int j = lane; j < n; j += 32) {
    const float v = z[j];
    if (v > mx) { mx = v; am = j; }
    const float p = 1.f / (1.f + expf(-v));
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);   // BCELoss clamps at -100
    acc -= (j == y) ? lp : l1p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    seg_loss[2 * row] = acc;
    seg_loss[2 * row + 1] = (float)n;
    if (argmax) argmax[row] = am;
  }
}
__global__ void bce_bwd_kernel(long long total, int n, const float* __restrict__ logits,
                               const int64_t* __restrict__ labels, float scale, float* __restrict__ dlogits) {
  EGOT2_PDL_ENTER();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / n; const int j = (int)(i % n);
  const float p = 1.f / (1.f + expf(-logits[i]));
  dlogits[i] = scale * (p - ((int)labels[row] == j ? 1.f : 0.f)) / (float)total;
}

// loss2[0] = sum(wnll) / (sum(w) / denom_div);  loss2[1] = sum(w) / denom_div
__global__ void loss_reduce_kernel(long long segs, const float* __restrict__ seg_loss, float denom_div,
                                   float* __restrict__ loss2) {
  EGOT2_PDL_ENTER();
  __shared__ float s0[32], s1[32];
  float a = 0.f, b = 0.f;
  for (long long i = threadIdx.x; i < segs; i += blockDim.x) { a += seg_loss[2 * i]; b += seg_loss[2 * i + 1]; }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? s0[threadIdx.x] : 0.f;
    b = threadIdx.x < (blockDim.x >> 5) ? s1[threadIdx.x] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
    if (threadIdx.x == 0) { const float w = b / denom_div; loss2[0] = a / w; loss2[1] = w; }
  }
}

// Batched PNR / OSCC evaluation metrics (HOI/evaluation/pnr/metrics.py:11-80), one thread per clip + one block-wide
// reduction in a fixed order (deterministic).  out_i64 = [correct, total]; out_f64 = [sum of keyframe time errors (s)].
//   pred   = first arg-max of the clip's logits (torch.argmax)
//   label  = label_idx[b] if given, else the first arg-max of the clip's one-hot row
//   clips with sc_label != 1 are skipped (sc_label == nullptr: every clip counts - state_change_accuracy)
//   time error (only if fps != nullptr): | float32((end - start) / 16 * pred) - (pnr - start) | / fps
__global__ void __launch_bounds__(1024) pnr_metrics_kernel(int B, int n, const float* __restrict__ logits,
                                                           const int64_t* __restrict__ label_idx,
                                                           const float* __restrict__ label_onehot,
                                                           const int64_t* __restrict__ sc_label, const double* __restrict__ fps,
                                                           const int64_t* __restrict__ start, const int64_t* __restrict__ end,
                                                           const int64_t* __restrict__ pnr, double* __restrict__ err_sec,
                                                           long long* __restrict__ out_i64, double* __restrict__ out_f64) {
  EGOT2_PDL_ENTER();
  __shared__ long long s_c[32], s_t[32];
  __shared__ double s_d[32];
  long long c = 0, t = 0;
  double dsum = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {        // fixed assignment clip -> thread: deterministic sums
    const float* z = logits + (size_t)b * n;
    int am = 0; float mx = z[0];
    for (int j = 1; j < n; ++j) if (z[j] > mx) { mx = z[j]; am = j; }
    long long lab;
    if (label_idx) lab = label_idx[b];
    else {
      const float* y = label_onehot + (size_t)b * n;
      int la = 0; float lm = y[0];
      for (int j = 1; j < n; ++j) if (y[j] > lm) { lm = y[j]; la = j; }
      lab = la;
    }
    const bool counted = sc_label == nullptr || sc_label[b] == 1;
    double e = 0.0;
    if (counted) {
      ++t;
      if ((long long)am == lab) ++c;
      if (fps) {
        const float mapped = ((float)(end[b] - start[b]) / 16.0f) * (float)am;     // the reference's float32 tensor arithmetic
        e = fabs((double)mapped - (double)(pnr[b] - start[b])) / fps[b];
        dsum += e;
      }
    }
    if (err_sec) err_sec[b] = counted ? e : -1.0;
  }
  // warp + block reduction in lane / warp order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_down_sync(0xffffffffu, c, o);
    t += __shfl_down_sync(0xffffffffu, t, o);
    dsum += __shfl_down_sync(0xffffffffu, dsum, o);
  }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = c; s_t[threadIdx.x >> 5] = t; s_d[threadIdx.x >> 5] = dsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long cc = 0, tt = 0; double dd = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { cc += s_c[w]; tt += s_t[w]; dd += s_d[w]; }
    out_i64[0] = cc; out_i64[1] = tt; out_f64[0] = dd;
  }
}

}  // namespace

int pnr_metrics(int B, int n, const float* logits, const int64_t* label_idx, const float* label_onehot,
                const int64_t* sc_label, const double* fps, const int64_t* start, const int64_t* end, const int64_t* pnr,
                double* err_sec, long long* out_i64, double* out_f64, cudaStream_t st) {
  EGOT2_CHECK(B >= 0 && n >= 1 && logits && (label_idx || label_onehot) && out_i64 && out_f64, "pnr_metrics: bad arguments");
  EGOT2_CHECK(!fps || (start && end && pnr), "pnr_metrics: the time error needs start/end/pnr frames");
  ProfScope prof(st, "pnr_metrics B%d n%d", B, n);
  launch(pnr_metrics_kernel, dim3(1), dim3(1024), 0, st, B, n, logits, label_idx, label_onehot, sc_label, fps, start, end, pnr,
         err_sec, out_i64, out_f64);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int loss_fwd(const egot2_head_desc& d, int rows, const float* logits, const int64_t* labels, const float* class_weight,
             float* row_loss, float* loss, int32_t* argmax, cudaStream_t st) {
  if (d.loss == EGOT2_LOSS_NONE || rows == 0) return 0;
  EGOT2_CHECK(labels && row_loss && loss, "loss_fwd: labels/row_loss/loss buffers required");
  const SegGeom g = geom(d);
  if (d.loss == EGOT2_LOSS_CE_GROUPS)
    EGOT2_CHECK(g.per_sub * g.sub_rows == d.n_out && d.n_groups >= 1 && d.n_groups <= EGOT2_MAX_GROUPS,
                "loss: CE_GROUPS geometry %d*%d != n_out %d", g.per_sub, g.sub_rows, d.n_out);
  const long long segs = (long long)rows * g.sub_rows * g.n_groups;
  const int grid = (int)((segs * 32 + 255) / 256);
  ProfScope prof(st, "loss_fwd kind%d rows%d n%d (2 kernels)", d.loss, rows, d.n_out);
  if (d.loss == EGOT2_LOSS_BCE_SIGMOID)
    launch(bce_fwd_kernel, dim3(grid), dim3(256), 0, st, rows, d.n_out, logits, labels, row_loss, argmax);
  else
    launch(ce_fwd_kernel, dim3(grid), dim3(256), 0, st, g, segs, logits, labels, class_weight, row_loss, argmax);
  EGOT2_LAUNCH_CHECK();
  const float denom_div = d.loss == EGOT2_LOSS_CE_GROUPS ? (float)(g.sub_rows * g.n_groups) : 1.f;
  launch(loss_reduce_kernel, dim3(1), dim3(1024), 0, st, segs, row_loss, denom_div, loss);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

// only the tree reduction of per-row (weighted nll, weight) pairs that a fused head kernel already produced
int loss_reduce(const egot2_head_desc& d, int rows, const float* row_loss, float* loss, cudaStream_t st) {
  if (d.loss == EGOT2_LOSS_NONE || rows == 0) return 0;
  EGOT2_CHECK(d.loss != EGOT2_LOSS_CE_GROUPS && row_loss && loss, "loss_reduce: single-segment losses only");
  ProfScope prof(st, "loss_reduce kind%d rows%d", d.loss, rows);
  launch(loss_reduce_kernel, dim3(1), dim3(1024), 0, st, (long long)rows, row_loss, 1.f, loss);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int loss_bwd(const egot2_head_desc& d, int rows, const float* logits, const int64_t* labels, const float* class_weight,
             const float* loss2, float dloss_scale, float* dlogits, cudaStream_t st) {
  if (d.loss == EGOT2_LOSS_NONE || rows == 0) return 0;
  const SegGeom g = geom(d);
  ProfScope prof(st, "loss_bwd kind%d rows%d n%d", d.loss, rows, d.n_out);
  if (d.loss == EGOT2_LOSS_BCE_SIGMOID) {
    const long long total = (long long)rows * d.n_out;
    launch(bce_bwd_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, st, total, d.n_out, logits, labels, dloss_scale, dlogits);
  } else {
    const long long segs = (long long)rows * g.sub_rows * g.n_groups;
    launch(ce_bwd_kernel, dim3((int)((segs * 32 + 255) / 256)), dim3(256), 0, st, g, segs, logits, labels, class_weight, loss2,
                                                                   dloss_scale, dlogits);
  }
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
