// head_fused.cu — the small classification heads in one kernel per direction (one CTA per clip):
//   forward : mean over the clip's T tokens -> LayerNorm (gamma/beta, optional dropout) -> Linear(H -> n_out <= 32)
//   backward: dlogits -> d(LN output) -> LayerNorm backward -> broadcast 1/T to every token of the clip,
//             plus d(gamma, beta, W, b) accumulated with one atomic per (clip, element)
// Replaces pool + LN + cast + GEMM (+ colsum + wgrad + dgrad + LN-bwd + pool-bwd): ~15 tiny launches -> 2.
// HHI TTM head (2 logits) and HOI PNR / OSCC heads (16 / 2 logits; LayerNorm shared with the token LN).
#define EGOT2_FILE_ID 3
#include "ops.h"

namespace egot2 {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += red[i];
  return t;
}

template <typename TT>
__global__ void __launch_bounds__(NT) head_fwd_kernel(int T, int H, int n_out, const TT* __restrict__ x,
                                                      const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                      const TT* __restrict__ W, const float* __restrict__ bias, float eps,
                                                      float p_drop, uint64_t drop_key, float* __restrict__ pooled,
                                                      float* __restrict__ stat, TT* __restrict__ g_out,
                                                      float* __restrict__ logits, int loss_kind,
                                                      const int64_t* __restrict__ labels, const float* __restrict__ cw,
                                                      float* __restrict__ seg_loss, int32_t* __restrict__ argmax) {
  EGOT2_PDL_ENTER();
  extern __shared__ float sm[];
  float* ps = sm;            // pooled (H)
  float* gs = sm + H;        // LN output (H)
  __shared__ float red[NT / 32];
  __shared__ float slog[32];
  const int row = blockIdx.x;
  const TT* xb = x + (size_t)row * T * H;
  if (sizeof(TT) == 2 && H == 128) {
    // bf16, H = 128: 16 lanes x 16 B cover a token row, the CTA's 16 lane groups take tokens t, t+16, ... (all loads
    // independent and in flight together), then the 16 partial rows are summed through shared memory
    __shared__ float part[16][128 + 4];
    const int cg = threadIdx.x & 15, tg = threadIdx.x >> 4;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int t = tg; t < T; t += 16) {
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(xb) + (size_t)t * H + cg * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { acc[2 * k] += __uint_as_float(w[k] << 16); acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[tg][cg * 8 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 128) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) s += part[g][threadIdx.x];
      s /= (float)T;
      ps[threadIdx.x] = s;
      pooled[(size_t)row * H + threadIdx.x] = s;
    }
  } else {
    for (int c = threadIdx.x; c < H; c += NT) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += to_f32(xb[(size_t)t * H + c]);
      s /= (float)T;
      ps[c] = s;
      pooled[(size_t)row * H + c] = s;
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int c = threadIdx.x; c < H; c += NT) part += ps[c];
  const float mean = block_sum(part, red) / (float)H;
  part = 0.f;
  for (int c = threadIdx.x; c < H; c += NT) { const float d = ps[c] - mean; part += d * d; }
  const float rstd = rsqrtf(block_sum(part, red) / (float)H + eps);
  if (threadIdx.x == 0) { stat[2 * (size_t)row] = mean; stat[2 * (size_t)row + 1] = rstd; }
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int c = threadIdx.x; c < H; c += NT) {
    float g = (ps[c] - mean) * rstd * ln_g[c] + ln_b[c];
    if (p_drop > 0.f) g *= drop_scale(drop_key ^ egot2_ep, (uint64_t)row * H + c, p_drop, inv_keep);
    const TT gr = from_f32<TT>(g);
    g_out[(size_t)row * H + c] = gr;
    gs[c] = to_f32(gr);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < n_out; j += NT / 32) {
    float d = 0.f;
    for (int c = lane; c < H; c += 32) d += gs[c] * to_f32(W[(size_t)j * H + c]);
    d = warp_sum(d);
    if (lane == 0) { logits[(size_t)row * n_out + j] = d + bias[j]; slog[j] = d + bias[j]; }
  }
  if (loss_kind == EGOT2_LOSS_NONE) return;
  // fused per-row loss (same arithmetic and reduction order as loss.cu's ce_fwd_kernel / bce_fwd_kernel: one warp per row)
  __syncthreads();
  if (warp != 0) return;
  const int y = (int)labels[row];
  const float z = lane < n_out ? slog[lane] : -INFINITY;
  float mx = z; int am = lane < n_out ? lane : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  if (loss_kind == EGOT2_LOSS_BCE_SIGMOID) {
    float acc = 0.f;
    if (lane < n_out) {
      const float p = 1.f / (1.f + expf(-z));
      const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);   // BCELoss clamps at -100
      acc = -((lane == y) ? lp : l1p);
    }
    acc = warp_sum(acc);
    if (lane == 0) { seg_loss[2 * (size_t)row] = acc; seg_loss[2 * (size_t)row + 1] = (float)n_out; }
  } else {
    float s = lane < n_out ? expf(z - mx) : 0.f;
    s = warp_sum(s);
    if (lane == 0) {
      const float w = cw ? cw[y] : 1.f;
      seg_loss[2 * (size_t)row] = w * ((mx + logf(s)) - slog[y]);
      seg_loss[2 * (size_t)row + 1] = w;
    }
  }
  if (lane == 0 && argmax) argmax[row] = am;
}

template <typename TT>
__global__ void __launch_bounds__(NT) head_bwd_kernel(int T, int H, int n_out, const float* __restrict__ dlogits,
                                                      const float* __restrict__ pooled, const float* __restrict__ stat,
                                                      const TT* __restrict__ g_saved, const float* __restrict__ ln_g,
                                                      const TT* __restrict__ W, float p_drop, uint64_t drop_key,
                                                      TT* __restrict__ dx, float* __restrict__ d_ln_g,
                                                      float* __restrict__ d_ln_b, float* __restrict__ dW,
                                                      float* __restrict__ db, int loss_kind, int rows,
                                                      const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                      const float* __restrict__ cw, float dloss_scale,
                                                      float* __restrict__ dlogits_out) {
  EGOT2_PDL_ENTER();
  extern __shared__ float sm[];
  float* dgs = sm;           // d(LN output) (H)
  float* dps = sm + H;       // d(pooled) (H)
  __shared__ float dl[32];
  __shared__ float red[NT / 32];
  const int row = blockIdx.x;
  if (loss_kind == EGOT2_LOSS_NONE) {
    if (threadIdx.x < n_out) {
      const float v = dlogits[(size_t)row * n_out + threadIdx.x];
      dl[threadIdx.x] = v;
      if (db) atomicAdd(db + threadIdx.x, v);
    }
  } else {
    // fused d(loss)/d(logits) of this clip (loss.cu ce_bwd_kernel / bce_bwd_kernel).  The normaliser of the weighted mean
    // (sum of the class weights over ALL rows) is recomputed by every CTA - a few hundred L2-resident loads - so that the
    // backward does not wait for a separate kernel.
    float wsum = 0.f;
    if (loss_kind == EGOT2_LOSS_CE) {
      float part = 0.f;
      for (int i = threadIdx.x; i < rows; i += NT) part += cw ? cw[(int)labels[i]] : 1.f;
      wsum = block_sum(part, red);
    }
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x, y = (int)labels[row];
      const float z = lane < n_out ? logits[(size_t)row * n_out + lane] : -INFINITY;
      float v = 0.f;
      if (loss_kind == EGOT2_LOSS_BCE_SIGMOID) {
        const float p = 1.f / (1.f + expf(-z));
        v = dloss_scale * (p - (lane == y ? 1.f : 0.f)) / ((float)rows * (float)n_out);
      } else {
        const float mx = warp_max(z);
        const float e = lane < n_out ? expf(z - mx) : 0.f;
        const float s = warp_sum(e);
        const float w = (cw ? cw[y] : 1.f) * dloss_scale / wsum;
        v = w * (e * (1.f / s) - (lane == y ? 1.f : 0.f));
      }
      if (lane < n_out) {
        dl[lane] = v;
        if (dlogits_out) dlogits_out[(size_t)row * n_out + lane] = v;
        if (db) atomicAdd(db + lane, v);
      }
    }
  }
  __syncthreads();
  const float mean = stat[2 * (size_t)row], rstd = stat[2 * (size_t)row + 1];
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < H; c += NT) {
    float dg = 0.f;
    const float gsv = to_f32(g_saved[(size_t)row * H + c]);
    for (int j = 0; j < n_out; ++j) {
      dg += dl[j] * to_f32(W[(size_t)j * H + c]);
      if (dW) atomicAdd(dW + (size_t)j * H + c, dl[j] * gsv);
    }
    if (p_drop > 0.f) dg *= drop_scale(drop_key ^ egot2_ep, (uint64_t)row * H + c, p_drop, inv_keep);
    const float xh = (pooled[(size_t)row * H + c] - mean) * rstd;
    if (d_ln_g) { atomicAdd(d_ln_g + c, dg * xh); atomicAdd(d_ln_b + c, dg); }
    const float dyg = dg * ln_g[c];
    dgs[c] = dyg;
    s1 += dyg;
    s2 += dyg * xh;
  }
  s1 = block_sum(s1, red) / (float)H;
  s2 = block_sum(s2, red) / (float)H;
  const float invT = 1.f / (float)T;
  for (int c = threadIdx.x; c < H; c += NT) {
    const float xh = (pooled[(size_t)row * H + c] - mean) * rstd;
    dps[c] = rstd * (dgs[c] - s1 - xh * s2) * invT;
  }
  __syncthreads();
  TT* dxb = dx + (size_t)row * T * H;
  if (sizeof(TT) == 2 && H % 8 == 0) {          // 16 B stores: every token row of the clip receives the same H values
    const int cpr = H / 8;                       // chunks per row
    for (int e = threadIdx.x; e < T * cpr; e += NT) {
      const int c = (e % cpr) * 8;
      uint4 o;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(dps[c], dps[c + 1]), p1 = __floats2bfloat162_rn(dps[c + 2], dps[c + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(dps[c + 4], dps[c + 5]), p3 = __floats2bfloat162_rn(dps[c + 6], dps[c + 7]);
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(dxb) + (size_t)e * 8) = o;
    }
  } else {
    for (int e = threadIdx.x; e < T * H; e += NT) dxb[e] = from_f32<TT>(dps[e % H]);
  }
}

}  // namespace

bool head_fused_supported(const egot2_head_desc& d) {
  return d.pool && d.use_ln && d.n_out >= 1 && d.n_out <= 32 && d.H % 32 == 0 && d.H <= 2048 &&
         (d.loss == EGOT2_LOSS_NONE || d.loss == EGOT2_LOSS_CE || d.loss == EGOT2_LOSS_BCE_SIGMOID);
}

int head_fused_fwd(const egot2_head_desc& d, const egot2_head_in& in, const egot2_head_out& out, cudaStream_t st) {
  const float ph = d.training ? d.p_head : 0.f;
  const uint64_t key = site_key(d.seed, SITE_HEAD, 0);
  const size_t smem = 2 * (size_t)d.H * sizeof(float);
  ProfScope prof(st, "head_fwd B%d T%d H%d n%d", d.B, d.T, d.H, d.n_out);
  if (d.dtype == EGOT2_F32)
    launch(head_fwd_kernel<float>, dim3(d.B), dim3(NT), smem, st, d.T, d.H, d.n_out, (const float*)in.x, in.ln_g, in.ln_b, (const float*)in.w,
                                                  in.b, d.ln_eps, ph, key, out.pooled, out.stat, (float*)out.g, out.logits,
           d.loss, in.labels, in.class_weight, out.row_loss, out.argmax);
  else
    launch(head_fwd_kernel<bf16>, dim3(d.B), dim3(NT), smem, st, d.T, d.H, d.n_out, (const bf16*)in.x, in.ln_g, in.ln_b, (const bf16*)in.w,
                                                 in.b, d.ln_eps, ph, key, out.pooled, out.stat, (bf16*)out.g, out.logits,
           d.loss, in.labels, in.class_weight, out.row_loss, out.argmax);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int head_fused_bwd(const egot2_head_desc& d, const egot2_head_in& in, const egot2_head_out& saved, float* dlogits,
                   float dloss_scale, void* dx, const egot2_head_grads& g, cudaStream_t st) {
  const float ph = d.training ? d.p_head : 0.f;
  const uint64_t key = site_key(d.seed, SITE_HEAD, 0);
  const size_t smem = 2 * (size_t)d.H * sizeof(float);
  ProfScope prof(st, "head_bwd B%d T%d H%d n%d", d.B, d.T, d.H, d.n_out);
  if (d.dtype == EGOT2_F32)
    launch(head_bwd_kernel<float>, dim3(d.B), dim3(NT), smem, st, d.T, d.H, d.n_out, dlogits, saved.pooled, saved.stat, (const float*)saved.g,
                                                  in.ln_g, (const float*)in.w, ph, key, (float*)dx, g.ln_g, g.ln_b, g.w, g.b,
           d.loss, d.B, saved.logits, in.labels, in.class_weight, dloss_scale, dlogits);
  else
    launch(head_bwd_kernel<bf16>, dim3(d.B), dim3(NT), smem, st, d.T, d.H, d.n_out, dlogits, saved.pooled, saved.stat, (const bf16*)saved.g,
                                                 in.ln_g, (const bf16*)in.w, ph, key, (bf16*)dx, g.ln_g, g.ln_b, g.w, g.b,
           d.loss, d.B, saved.logits, in.labels, in.class_weight, dloss_scale, dlogits);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
