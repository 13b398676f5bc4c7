// attention_small.cu — attention for the EgoT2-g decoder (HHI/models/multitask/task_prompt_model.py:260-269):
// a handful of query tokens (the 2-token task prompt) per row attending either to themselves under a causal mask
// (decoder self-attention) or to that row's encoder memory (cross-attention, 3 ... ~450 keys).  One warp owns one
// (row, head, query): lanes stride over the keys for the scores, then over the head columns for P.V.  fp32 arithmetic
// for both activation dtypes; the probabilities never leave shared memory; backward recomputes them.
//
// The memory of row n, key j lives at kv row   (n / inner) * outer + j * jstride + (n % inner) * istride
//   plain (B, M, .) memory:      inner = 1, outer = M,  jstride = 1, istride = 0
//   'asd' regrouping (:251-257): inner = T, outer = 3T, jstride = T, istride = 1  (row n = b*T + t sees tokens t, T+t, 2T+t)
#include <math.h>

#define EGOT2_FILE_ID 8
#include "ops.h"

namespace egot2 {

namespace {

constexpr int WARPS = 4;
constexpr int MAXK = 1024;          // keys per row (scores live in shared memory)

// dot product of one key / value row (dh contiguous elements of the activation dtype) with a per-warp fp32 vector in
// shared memory (broadcast reads).  16-byte loads when the row allows it: with one 2-byte load per element this loop was
// the whole kernel (36 / 60 us per launch for 8 clips x 2 prompt tokens x 90 memory tokens).
template <typename T>
__device__ __forceinline__ float dot_row(const T* __restrict__ r, const float* __restrict__ sv, int dh) {
  float d = 0.f;
  if ((dh & 7) == 0 && (reinterpret_cast<uintptr_t>(r) & 15) == 0) {
    if constexpr (sizeof(T) == 2) {
      for (int c = 0; c < dh; c += 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(r + c);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          d = fmaf(sv[c + 2 * k], __uint_as_float(ww[k] << 16), d);
          d = fmaf(sv[c + 2 * k + 1], __uint_as_float(ww[k] & 0xffff0000u), d);
        }
      }
    } else {
      for (int c = 0; c < dh; c += 4) {
        const float4 w = *reinterpret_cast<const float4*>(r + c);
        d = fmaf(sv[c], w.x, d); d = fmaf(sv[c + 1], w.y, d); d = fmaf(sv[c + 2], w.z, d); d = fmaf(sv[c + 3], w.w, d);
      }
    }
    return d;
  }
  for (int c = 0; c < dh; ++c) d = fmaf(sv[c], to_f32(r[c]), d);
  return d;
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32) small_attn_fwd_kernel(const SmallAttnArgs a) {
  EGOT2_PDL_ENTER();
  __shared__ float sc[WARPS][MAXK];
  __shared__ float sq[WARPS][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dh = a.H / a.heads;
  const long long total = (long long)a.rows * a.heads * a.S;
  const long long qid = (long long)blockIdx.x * WARPS + warp;
  if (qid >= total) return;
  const int s = (int)(qid % a.S), h = (int)((qid / a.S) % a.heads);
  const int n = (int)(qid / ((long long)a.S * a.heads));
  const int nk = a.causal ? s + 1 : a.M;
  const float scale = rsqrtf((float)dh);
  const T* q = (const T*)a.q + ((size_t)n * a.S + s) * a.ldq + h * dh;
  for (int c = lane; c < dh; c += 32) sq[warp][c] = to_f32(q[c]) * scale;
  __syncwarp();
  const long long kv0 = (long long)(n / a.kv_inner) * a.kv_outer + (long long)(n % a.kv_inner) * a.kv_istride;
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) {
    const T* k = (const T*)a.k + (size_t)(kv0 + (long long)j * a.kv_jstride) * a.ldkv + h * dh;
    const float d = dot_row(k, sq[warp], dh);
    sc[warp][j] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) { const float p = expf(sc[warp][j] - mx); sc[warp][j] = p; sum += p; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  for (int j = lane; j < nk; j += 32) {
    float p = sc[warp][j] * inv;
    if (a.p_drop > 0.f) p *= drop_scale(a.drop_key ^ egot2_ep, (uint64_t)qid * a.M + j, a.p_drop, inv_keep);
    sc[warp][j] = p;
  }
  if (lane == 0 && a.lse) a.lse[qid] = mx + logf(sum);
  __syncwarp();
  T* o = (T*)a.out + ((size_t)n * a.S + s) * a.ldo + h * dh;
  for (int c = lane; c < dh; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < nk; ++j) {
      const T* v = (const T*)a.v + (size_t)(kv0 + (long long)j * a.kv_jstride) * a.ldkv + h * dh;
      acc = fmaf(sc[warp][j], to_f32(v[c]), acc);
    }
    o[c] = from_f32<T>(acc);
  }
}

// dq (rows*S, ldq) / dk, dv (kv rows, ldkv): fp32, ACCUMULATED with atomics (several queries / heads / rows share keys)
template <typename T>
__global__ void __launch_bounds__(WARPS * 32) small_attn_bwd_kernel(const SmallAttnArgs a) {
  EGOT2_PDL_ENTER();
  __shared__ float sp[WARPS][MAXK];      // raw probabilities, then ds
  __shared__ float spd[WARPS][MAXK];     // dropped probabilities (what multiplied V)
  __shared__ float sq[WARPS][128];
  __shared__ float sdo[WARPS][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dh = a.H / a.heads;
  const long long total = (long long)a.rows * a.heads * a.S;
  const long long qid = (long long)blockIdx.x * WARPS + warp;
  if (qid >= total) return;
  const int s = (int)(qid % a.S), h = (int)((qid / a.S) % a.heads);
  const int n = (int)(qid / ((long long)a.S * a.heads));
  const int nk = a.causal ? s + 1 : a.M;
  const float scale = rsqrtf((float)dh);
  const T* q = (const T*)a.q + ((size_t)n * a.S + s) * a.ldq + h * dh;
  const T* dout = (const T*)a.dout + ((size_t)n * a.S + s) * a.ldo + h * dh;
  for (int c = lane; c < dh; c += 32) { sq[warp][c] = to_f32(q[c]) * scale; sdo[warp][c] = to_f32(dout[c]); }
  __syncwarp();
  const long long kv0 = (long long)(n / a.kv_inner) * a.kv_outer + (long long)(n % a.kv_inner) * a.kv_istride;
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  // recompute the probabilities (same arithmetic as the forward) and dP = dO . V^T
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) {
    const T* k = (const T*)a.k + (size_t)(kv0 + (long long)j * a.kv_jstride) * a.ldkv + h * dh;
    const float d = dot_row(k, sq[warp], dh);
    sp[warp][j] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) { const float p = expf(sp[warp][j] - mx); sp[warp][j] = p; sum += p; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float dsum = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const T* v = (const T*)a.v + (size_t)(kv0 + (long long)j * a.kv_jstride) * a.ldkv + h * dh;
    const float dp = dot_row(v, sdo[warp], dh);
    const float p = sp[warp][j] * inv;
    const float mk = a.p_drop > 0.f ? drop_scale(a.drop_key ^ egot2_ep, (uint64_t)qid * a.M + j, a.p_drop, inv_keep) : 1.f;
    spd[warp][j] = p * mk;
    dsum += p * dp * mk;                   // dp * mk = gradient w.r.t. the un-dropped probability
    sp[warp][j] = p;
  }
  dsum = warp_sum(dsum);
  // ds_j = p_j (dp_j - sum_i p_i dp_i); dp_j is recomputed (cheaper than a third key-sized array)
  for (int j = lane; j < nk; j += 32) {
    const T* v = (const T*)a.v + (size_t)(kv0 + (long long)j * a.kv_jstride) * a.ldkv + h * dh;
    const float dp = dot_row(v, sdo[warp], dh);
    const float p = sp[warp][j];
    const float mk = p > 0.f ? spd[warp][j] / p : 0.f;      // 0 or 1/(1-p): the mask that was applied
    sp[warp][j] = p * (dp * mk - dsum);
  }
  __syncwarp();
  // dq = scale * ds . K ;  dK_j += scale * ds_j * q_raw = ds_j * (q * scale) ;  dV_j += pd_j * dO
  float* dq = a.dq + ((size_t)n * a.S + s) * a.ld_dq + h * dh;
  for (int c = lane; c < dh; c += 32) {
    float acc = 0.f;
    for (int j = 0; j < nk; ++j) {
      const size_t row = (size_t)(kv0 + (long long)j * a.kv_jstride);
      const T* k = (const T*)a.k + row * a.ldkv + h * dh;
      acc = fmaf(sp[warp][j], to_f32(k[c]), acc);
      atomicAdd(a.dk + row * a.ld_dkv + h * dh + c, sp[warp][j] * sq[warp][c]);
      atomicAdd(a.dv + row * a.ld_dkv + h * dh + c, spd[warp][j] * sdo[warp][c]);
    }
    atomicAdd(dq + c, acc * scale);
  }
}

}  // namespace

static int check(const SmallAttnArgs& a) {
  EGOT2_CHECK(a.heads > 0 && a.H % a.heads == 0 && a.H / a.heads <= 128, "small_attn: head dim %d > 128", a.heads ? a.H / a.heads : 0);
  EGOT2_CHECK(a.M >= 1 && a.M <= MAXK, "small_attn: %d keys per row (max %d)", a.M, MAXK);
  EGOT2_CHECK(!a.causal || a.M == a.S, "small_attn: the causal mask needs M == S");
  EGOT2_CHECK(a.kv_inner >= 1, "small_attn: kv_inner must be >= 1");
  return 0;
}

int small_attn_fwd(const SmallAttnArgs& a, cudaStream_t st) {
  EGOT2_TRY(check(a));
  const long long total = (long long)a.rows * a.heads * a.S;
  if (total == 0) return 0;
  ProfScope prof(st, "small_attn_fwd rows%d S%d M%d H%d", a.rows, a.S, a.M, a.H);
  const int grid = (int)((total + WARPS - 1) / WARPS);
  if (a.dtype == EGOT2_F32) launch(small_attn_fwd_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, st, a);
  else launch(small_attn_fwd_kernel<bf16>, dim3(grid), dim3(WARPS * 32), 0, st, a);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int small_attn_bwd(const SmallAttnArgs& a, cudaStream_t st) {
  EGOT2_TRY(check(a));
  EGOT2_CHECK(a.dout && a.dq && a.dk && a.dv, "small_attn_bwd: gradient buffers required");
  const long long total = (long long)a.rows * a.heads * a.S;
  if (total == 0) return 0;
  ProfScope prof(st, "small_attn_bwd rows%d S%d M%d H%d", a.rows, a.S, a.M, a.H);
  const int grid = (int)((total + WARPS - 1) / WARPS);
  if (a.dtype == EGOT2_F32) launch(small_attn_bwd_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, st, a);
  else launch(small_attn_bwd_kernel<bf16>, dim3(grid), dim3(WARPS * 32), 0, st, a);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
