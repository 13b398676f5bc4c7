// attention_wide.cu — self-attention for WIDE heads over a handful of tokens: head dim > 128 (up to any width), T <= 32.
// The one shipped user is the LTA 2-task translator at TRANSLATION_INPUT_FEATURES = 2048 with 4 heads (head dim 512) over
// 4 tokens (HOI/models/lta/lta_models_lta_transfer.py:429-526, HOI/configs/lta/ts_lta_2task.yaml:79-82), which the tiled
// kernels of attention_simt.cu / attention_mma.cu (head dim <= 128 / <= 64) do not cover.  One CTA per (clip, head): the
// T x T scores live in shared memory, every dot product over the head dim is a warp reduction with coalesced loads.
// Same math, dropout indexing and saved statistics (lse) as the other attention paths; fp32 arithmetic for both dtypes.
#include <math.h>

#define EGOT2_FILE_ID 13
#include "ops.h"

namespace egot2 {
namespace {

constexpr int MAXT = 32;
constexpr int NWW = 8;         // warps per CTA

template <typename T>
__global__ void __launch_bounds__(NWW * 32) attn_wide_fwd_kernel(int Tn, int H, int heads, int dh, const T* __restrict__ qkv,
                                                                 T* __restrict__ out, float* __restrict__ lse, float p_drop,
                                                                 uint64_t drop_key) {
  EGOT2_PDL_ENTER();
  __shared__ float P[MAXT][MAXT + 1];
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = 3 * H;
  const T* qbase = qkv + (size_t)b * Tn * ld + h * dh;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const float scale = rsqrtf((float)dh);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  // scores: one warp per (query, key) pair, lanes over the head dim
  for (int pair = warp; pair < Tn * Tn; pair += NWW) {
    const int i = pair / Tn, j = pair % Tn;
    const T* q = qbase + (size_t)i * ld;
    const T* k = kbase + (size_t)j * ld;
    float d = 0.f;
    for (int c = lane; c < dh; c += 32) d = fmaf(to_f32(q[c]), to_f32(k[c]), d);
    d = warp_sum(d);
    if (lane == 0) P[i][j] = d * scale;
  }
  __syncthreads();
  // softmax over the keys: one warp per query, lane = key
  for (int i = warp; i < Tn; i += NWW) {
    const float s = lane < Tn ? P[i][lane] : -INFINITY;
    const float mx = warp_max(s);
    const float e = lane < Tn ? expf(s - mx) : 0.f;
    const float sum = warp_sum(e);
    if (lane == 0 && lse) lse[(size_t)bh * Tn + i] = mx + logf(sum);
    if (lane < Tn) {
      float p = e / sum;
      if (p_drop > 0.f) p *= attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * Tn + i, Tn, lane, p_drop, inv_keep);
      P[i][lane] = p;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Tn * dh; e += blockDim.x) {
    const int i = e / dh, c = e % dh;
    float a = 0.f;
    for (int j = 0; j < Tn; ++j) a = fmaf(P[i][j], to_f32(vbase[(size_t)j * ld + c]), a);
    out[((size_t)b * Tn + i) * H + h * dh + c] = from_f32<T>(a);
  }
}

template <typename T>
__global__ void __launch_bounds__(NWW * 32) attn_wide_bwd_kernel(int Tn, int H, int heads, int dh, const T* __restrict__ qkv,
                                                                 const T* __restrict__ out, const float* __restrict__ lse,
                                                                 const T* __restrict__ dout, T* __restrict__ dqkv,
                                                                 float p_drop, uint64_t drop_key) {
  EGOT2_PDL_ENTER();
  __shared__ float Pd[MAXT][MAXT + 1];     // probabilities with the dropout mask applied (what multiplied V)
  __shared__ float dS[MAXT][MAXT + 1];     // gradient w.r.t. the scaled scores
  __shared__ float Dv[MAXT];               // D_i = dO_i . O_i
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ld = 3 * H;
  const T* qbase = qkv + (size_t)b * Tn * ld + h * dh;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const T* obase = out + (size_t)b * Tn * H + h * dh;
  const T* dobase = dout + (size_t)b * Tn * H + h * dh;
  const float scale = rsqrtf((float)dh);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int i = warp; i < Tn; i += NWW) {
    float d = 0.f;
    for (int c = lane; c < dh; c += 32) d = fmaf(to_f32(dobase[(size_t)i * H + c]), to_f32(obase[(size_t)i * H + c]), d);
    d = warp_sum(d);
    if (lane == 0) Dv[i] = d;
  }
  __syncthreads();
  // P is recomputed from the saved log-sum-exp; dP_ij = dO_i . V_j
  for (int pair = warp; pair < Tn * Tn; pair += NWW) {
    const int i = pair / Tn, j = pair % Tn;
    const T* q = qbase + (size_t)i * ld;
    const T* k = kbase + (size_t)j * ld;
    const T* v = vbase + (size_t)j * ld;
    const T* dO = dobase + (size_t)i * H;
    float s = 0.f, dp = 0.f;
    for (int c = lane; c < dh; c += 32) {
      s = fmaf(to_f32(q[c]), to_f32(k[c]), s);
      dp = fmaf(to_f32(dO[c]), to_f32(v[c]), dp);
    }
    s = warp_sum(s);
    dp = warp_sum(dp);
    if (lane == 0) {
      const float p = expf(s * scale - lse[(size_t)bh * Tn + i]);
      float mk = 1.f;
      if (p_drop > 0.f) mk = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * Tn + i, Tn, j, p_drop, inv_keep);
      Pd[i][j] = p * mk;
      dS[i][j] = p * (dp * mk - Dv[i]);
    }
  }
  __syncthreads();
  // dQ_i = scale * sum_j dS_ij K_j ;  dK_j = scale * sum_i dS_ij Q_i ;  dV_j = sum_i Pd_ij dO_i
  for (int e = threadIdx.x; e < Tn * dh; e += blockDim.x) {
    const int r = e / dh, c = e % dh;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int t = 0; t < Tn; ++t) {
      dq = fmaf(dS[r][t], to_f32(kbase[(size_t)t * ld + c]), dq);
      dk = fmaf(dS[t][r], to_f32(qbase[(size_t)t * ld + c]), dk);
      dv = fmaf(Pd[t][r], to_f32(dobase[(size_t)t * H + c]), dv);
    }
    T* o = dqkv + ((size_t)b * Tn + r) * ld + h * dh + c;
    o[0] = from_f32<T>(dq * scale);
    o[H] = from_f32<T>(dk * scale);
    o[2 * H] = from_f32<T>(dv);
  }
}

int wide_check(int dtype, int B, int T, int H, int heads) {
  EGOT2_CHECK(dtype == EGOT2_F32 || dtype == EGOT2_BF16, "attention_wide: bad dtype %d", dtype);
  EGOT2_CHECK(heads > 0 && H % heads == 0, "attention_wide: H=%d not divisible by heads=%d", H, heads);
  EGOT2_CHECK(T >= 1 && T <= MAXT, "attention_wide: %d tokens per clip (head dim %d > 128 is only built for T <= %d)", T,
              H / heads, MAXT);
  (void)B;
  return 0;
}

}  // namespace

int attention_wide_fwd(int dtype, int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p_drop,
                       uint64_t drop_key, cudaStream_t st) {
  EGOT2_TRY(wide_check(dtype, B, T, H, heads));
  if (B == 0) return 0;
  const int dh = H / heads;
  ProfScope prof(st, "attn_wide_fwd B%d T%d H%d dh%d", B, T, H, dh);
  if (dtype == EGOT2_F32)
    launch(attn_wide_fwd_kernel<float>, dim3(B * heads), dim3(NWW * 32), 0, st, T, H, heads, dh, (const float*)qkv, (float*)out,
           lse, p_drop, drop_key);
  else
    launch(attn_wide_fwd_kernel<bf16>, dim3(B * heads), dim3(NWW * 32), 0, st, T, H, heads, dh, (const bf16*)qkv, (bf16*)out, lse,
           p_drop, drop_key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int attention_wide_bwd(int dtype, int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse,
                       const void* dout, void* dqkv, float p_drop, uint64_t drop_key, cudaStream_t st) {
  EGOT2_TRY(wide_check(dtype, B, T, H, heads));
  EGOT2_CHECK(lse != nullptr, "attention_wide_bwd: the forward's lse is required");
  if (B == 0) return 0;
  const int dh = H / heads;
  ProfScope prof(st, "attn_wide_bwd B%d T%d H%d dh%d", B, T, H, dh);
  if (dtype == EGOT2_F32)
    launch(attn_wide_bwd_kernel<float>, dim3(B * heads), dim3(NWW * 32), 0, st, T, H, heads, dh, (const float*)qkv,
           (const float*)out, lse, (const float*)dout, (float*)dqkv, p_drop, drop_key);
  else
    launch(attn_wide_bwd_kernel<bf16>, dim3(B * heads), dim3(NWW * 32), 0, st, T, H, heads, dh, (const bf16*)qkv,
           (const bf16*)out, lse, (const bf16*)dout, (bf16*)dqkv, p_drop, drop_key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
