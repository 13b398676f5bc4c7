// rowops.cu — HBM-bound row-wise kernels: LayerNorm fwd/bwd, token pooling, column reductions,
// dropout, casts, Adam, token-table helpers, SlowFast map pooling.
// One warp owns one row of H (H = 32*NPL, NPL in {1,2,4,8,16,32}); all statistics in fp32.
// Loads/stores are coalesced (lane-strided columns); grids are sized to a multiple of the SM count.
#include <math.h>

#include "ops.h"

namespace egot2 {

namespace {

constexpr int kWarpsPerCta = 8;

inline int row_grid(int rows) {
  int ctas = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  int cap = sm_count() * 8;
  return ctas < cap ? (ctas > 0 ? ctas : 1) : cap;
}

// ------------------------------------------------------------------ LayerNorm forward
template <typename TX, typename TY, int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) ln_fwd_kernel(const LayerNormArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = NPL * 32;
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  float g[NPL], b[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) { g[i] = a.g[lane + 32 * i]; b[i] = a.b[lane + 32 * i]; }
  for (int row = blockIdx.x * kWarpsPerCta + warp; row < a.rows; row += gridDim.x * kWarpsPerCta) {
    const TX* x = (const TX*)a.x + (size_t)row * H;
    float v[NPL], s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { v[i] = to_f32(x[lane + 32 * i]); s += v[i]; }
    const float mean = warp_sum(s) * (1.f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / H) + a.eps);
    if (a.stat && lane == 0) { a.stat[2 * (size_t)row] = mean; a.stat[2 * (size_t)row + 1] = rstd; }
    const float* tab = a.table ? a.table + (size_t)(row % a.table_rows) * H : nullptr;
    TY* y = (TY*)a.y + (size_t)row * H;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      float o = (v[i] - mean) * rstd * g[i] + b[i];
      if (tab) o += tab[c];
      if (a.p_drop > 0.f) o *= drop_scale(a.drop_key, (uint64_t)row * H + c, a.p_drop, inv_keep);
      y[c] = from_f32<TY>(o);
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward
template <typename TX, typename TDY, typename TDX, typename TR, int NPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32) ln_bwd_kernel(const LayerNormBwdArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = NPL * 32;
  __shared__ float sdg[NPL * 32], sdb[NPL * 32];
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sdg[i] = 0.f; sdb[i] = 0.f; }
  __syncthreads();
  float g[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) { g[i] = a.g[lane + 32 * i]; dg[i] = 0.f; db[i] = 0.f; }
  for (int row = blockIdx.x * kWarpsPerCta + warp; row < a.rows; row += gridDim.x * kWarpsPerCta) {
    const TX* x = (const TX*)a.x + (size_t)row * H;
    const TDY* dy = (const TDY*)a.dy + (size_t)row * H;
    const float mean = a.stat[2 * (size_t)row], rstd = a.stat[2 * (size_t)row + 1];
    float xh[NPL], dyg[NPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      const float d = to_f32(dy[c]);
      xh[i] = (to_f32(x[c]) - mean) * rstd;
      dyg[i] = d * g[i];
      dg[i] += d * xh[i];
      db[i] += d;
      s1 += dyg[i];
      s2 += dyg[i] * xh[i];
    }
    s1 = warp_sum(s1) * (1.f / H);
    s2 = warp_sum(s2) * (1.f / H);
    TDX* dx = (TDX*)a.dx + (size_t)row * H;
    const TR* dres = a.dres ? (const TR*)a.dres + (size_t)row * H : nullptr;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int c = lane + 32 * i;
      float o = rstd * (dyg[i] - s1 - xh[i] * s2);
      if (dres) o += to_f32(dres[c]);
      dx[c] = from_f32<TDX>(o);
    }
  }
  if (a.dg) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) { atomicAdd(&sdg[lane + 32 * i], dg[i]); atomicAdd(&sdb[lane + 32 * i], db[i]); }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += blockDim.x) { atomicAdd(a.dg + i, sdg[i]); atomicAdd(a.db + i, sdb[i]); }
  }
}

// ------------------------------------------------------------------ misc elementwise / reductions
template <typename TT>
__global__ void table_grad_kernel(int B, int T, int H, int clips_per_cta, const TT* __restrict__ dy,
                                  float* __restrict__ dtable) {
  // CTA (t, chunk): sums its chunk of clips for token t, then one atomic per column
  const int t = blockIdx.x;
  const int b0 = blockIdx.y * clips_per_cta, b1 = min(B, b0 + clips_per_cta);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float s = 0.f;
    for (int b = b0; b < b1; ++b) s += to_f32(dy[((size_t)b * T + t) * H + c]);
    atomicAdd(dtable + (size_t)t * H + c, s);
  }
}

template <typename T> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<bf16> { typedef __nv_bfloat162 type; };
__device__ __forceinline__ float2 pair_f32(float2 v) { return v; }
__device__ __forceinline__ float2 pair_f32(__nv_bfloat162 v) { return __bfloat1622float2(v); }

template <typename T>
__global__ void colsum_kernel(int M, int N, const T* __restrict__ x, int ldx, int rpg, int gstride,
                              float* __restrict__ out, int rows_per_cta) {
  // blockDim = (32 column pairs, 8 row lanes): a warp reads 64 consecutive columns of one row per load
  typedef typename Pair<T>::type P2;
  const int n = (blockIdx.x * 32 + threadIdx.x) * 2;
  const int mbeg = blockIdx.y * rows_per_cta, mend = min(M, mbeg + rows_per_cta);
  float s0 = 0.f, s1 = 0.f;
  const bool vec = (n + 1 < N) && ((ldx & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & (2 * sizeof(T) - 1)) == 0);
  if (n < N) {
    for (int mb = mbeg + threadIdx.y; mb < mend; mb += 32) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = mb + 8 * u;
        v[u] = make_float2(0.f, 0.f);
        if (m < mend) {
          const long long r = rpg > 0 ? (long long)(m / rpg) * gstride + (m % rpg) : m;
          if (vec) v[u] = pair_f32(*reinterpret_cast<const P2*>(x + r * ldx + n));
          else { v[u].x = to_f32(x[r * ldx + n]); if (n + 1 < N) v[u].y = to_f32(x[r * ldx + n + 1]); }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s0 += v[u].x; s1 += v[u].y; }
    }
  }
  __shared__ float red[8][65];
  red[threadIdx.y][2 * threadIdx.x] = s0;
  red[threadIdx.y][2 * threadIdx.x + 1] = s1;
  __syncthreads();
  if (threadIdx.y < 2) {
    const int c = 2 * threadIdx.x + threadIdx.y, nn = blockIdx.x * 64 + c;
    if (nn < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][c];
      atomicAdd(out + nn, t);
    }
  }
}

template <typename T>
__global__ void dropout_kernel(T* x, size_t n, float p, float inv_keep, uint64_t key) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = from_f32<T>(to_f32(x[i]) * drop_scale(key, i, p, inv_keep));
}

template <typename TT>
__global__ void pool_fwd_kernel(int B, int T, int H, int pool, int row_tokens, const TT* __restrict__ x,
                                float* __restrict__ pooled) {
  if (pool) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += to_f32(x[((size_t)b * T + t) * H + c]);
      pooled[(size_t)b * H + c] = s / (float)T;
    }
  } else {
    const int r = blockIdx.x, b = r / row_tokens, t = r % row_tokens;
    for (int c = threadIdx.x; c < H; c += blockDim.x)
      pooled[(size_t)r * H + c] = to_f32(x[((size_t)b * T + t) * H + c]);
  }
}

template <typename TT>
__global__ void pool_bwd_kernel(int B, int T, int H, int pool, int row_tokens, const float* __restrict__ dpooled,
                                TT* __restrict__ dx) {
  const int r = blockIdx.x, b = r / T, t = r % T;     // one CTA per token row of dx
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v;
    if (pool) v = dpooled[(size_t)b * H + c] / (float)T;
    else v = t < row_tokens ? dpooled[((size_t)b * row_tokens + t) * H + c] : 0.f;
    dx[(size_t)r * H + c] = from_f32<TT>(v);
  }
}

__global__ void cast_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, size_t n) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(s + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(d + i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t j = n & ~(size_t)3; j < n; ++j) d[j] = __float2bfloat16_rn(s[j]);
}
__global__ void cast_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}
__global__ void copy_f32_kernel(const float* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
__global__ void zero_kernel(float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float grad = g[i] * gscale;
    const float w = p[i];
    if (wd != 0.f) grad += wd * w;
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi; v[i] = vi;
    // torch.optim.Adam: denom = sqrt(v)/sqrt(bias_correction2) + eps ; step = lr / bias_correction1
    p[i] = w - (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

struct SegList { int n; int tokens[EGOT2_MAX_SEG]; int task[EGOT2_MAX_SEG]; };

__global__ void hhi_tok_table_fwd_kernel(const float* __restrict__ task_embed, const float* __restrict__ pe,
                                         SegList s, int H, float* __restrict__ table) {
  // one CTA per token; d restarts at 0 for every task (PositionalEncoding applied per task before the concat)
  int t = blockIdx.x, seg = 0, d = t;
  while (seg < s.n - 1 && d >= s.tokens[seg]) { d -= s.tokens[seg]; ++seg; }
  for (int c = threadIdx.x; c < H; c += blockDim.x)
    table[(size_t)t * H + c] = task_embed[(size_t)s.task[seg] * H + c] + pe[(size_t)d * H + c];
}
__global__ void hhi_tok_table_bwd_kernel(const float* __restrict__ dtable, SegList s, int H,
                                         float* __restrict__ d_task_embed) {
  // one CTA per segment: d_task_embed[task] += sum over the segment's tokens
  const int seg = blockIdx.x;
  int off = 0;
  for (int i = 0; i < seg; ++i) off += s.tokens[i];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int d = 0; d < s.tokens[seg]; ++d) acc += dtable[(size_t)(off + d) * H + c];
    atomicAdd(d_task_embed + (size_t)s.task[seg] * H + c, acc);
  }
}

// AdaptiveAvgPool3d((Tout,1,1)) + permute: in (B,C,Tin,hw) -> out (B,Tout,C).  One warp per (b,c,to):
// the Tin/Tout*hw contiguous elements are read coalesced.
template <typename TI, typename TO>
__global__ void slowfast_pool_kernel(const TI* __restrict__ in, int B, int C, int Tin, int hw, int Tout,
                                     TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t total = (size_t)B * C * Tout;
  if (warp >= total) return;
  const int to = warp % Tout;
  const int c = (warp / Tout) % C;
  const int b = warp / ((size_t)Tout * C);
  const int win = Tin / Tout, n = win * hw;
  const TI* src = in + (((size_t)b * C + c) * Tin + (size_t)to * win) * hw;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += to_f32(src[i]);
  s = warp_sum(s);
  if (lane == 0) out[((size_t)b * Tout + to) * C + c] = from_f32<TO>(s / (float)n);
}

template <int NPL> int ln_fwd_dispatch(const LayerNormArgs& a, cudaStream_t st) {
  const int grid = row_grid(a.rows);
  ProfScope prof(st, "ln_fwd rows%d H%d", a.rows, a.H);
  if (a.dtype == EGOT2_F32) ln_fwd_kernel<float, float, NPL><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  else if (a.x_is_f32) ln_fwd_kernel<float, bf16, NPL><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  else ln_fwd_kernel<bf16, bf16, NPL><<<grid, kWarpsPerCta * 32, 0, st>>>(a);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

template <int NPL> int ln_bwd_dispatch(const LayerNormBwdArgs& a, cudaStream_t st) {
  const int grid = row_grid(a.rows);
  const int nt = kWarpsPerCta * 32;
  ProfScope prof(st, "ln_bwd rows%d H%d", a.rows, a.H);
  if (a.dtype == EGOT2_F32) {
    ln_bwd_kernel<float, float, float, float, NPL><<<grid, nt, 0, st>>>(a);
  } else {
    // bf16 activations; the pooled-vector LN of the head runs with fp32 x / dy / dx
    if (a.x_is_f32 && a.dy_is_f32 && a.dx_is_f32) ln_bwd_kernel<float, float, float, bf16, NPL><<<grid, nt, 0, st>>>(a);
    else if (a.x_is_f32 && !a.dy_is_f32 && a.dx_is_f32) ln_bwd_kernel<float, bf16, float, bf16, NPL><<<grid, nt, 0, st>>>(a);
    else if (!a.x_is_f32 && !a.dy_is_f32 && !a.dx_is_f32) ln_bwd_kernel<bf16, bf16, bf16, bf16, NPL><<<grid, nt, 0, st>>>(a);
    else EGOT2_CHECK(false, "layernorm_bwd: unsupported dtype mix");
  }
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int layernorm_fwd(const LayerNormArgs& a, cudaStream_t st) {
  EGOT2_CHECK(a.H % 32 == 0 && a.H <= 1024, "layernorm: H=%d must be a multiple of 32 and <= 1024", a.H);
  if (a.rows == 0) return 0;
  switch (a.H / 32) {
    case 1: return ln_fwd_dispatch<1>(a, st);
    case 2: return ln_fwd_dispatch<2>(a, st);
    case 4: return ln_fwd_dispatch<4>(a, st);
    case 8: return ln_fwd_dispatch<8>(a, st);
    case 16: return ln_fwd_dispatch<16>(a, st);
    case 32: return ln_fwd_dispatch<32>(a, st);
  }
  EGOT2_CHECK(false, "layernorm: H=%d not in {32,64,128,256,512,1024}", a.H);
}

int layernorm_bwd(const LayerNormBwdArgs& a, cudaStream_t st) {
  EGOT2_CHECK(a.H % 32 == 0 && a.H <= 1024, "layernorm_bwd: H=%d must be a multiple of 32 and <= 1024", a.H);
  if (a.rows == 0) return 0;
  switch (a.H / 32) {
    case 1: return ln_bwd_dispatch<1>(a, st);
    case 2: return ln_bwd_dispatch<2>(a, st);
    case 4: return ln_bwd_dispatch<4>(a, st);
    case 8: return ln_bwd_dispatch<8>(a, st);
    case 16: return ln_bwd_dispatch<16>(a, st);
    case 32: return ln_bwd_dispatch<32>(a, st);
  }
  EGOT2_CHECK(false, "layernorm_bwd: H=%d not in {32,64,128,256,512,1024}", a.H);
}

int table_grad(int dtype, int B, int T, int H, const void* dy, float* dtable, cudaStream_t st) {
  const int nt = H < 256 ? ((H + 31) / 32 * 32) : 256;
  int chunks = (4 * sm_count() + T - 1) / T;           // ~4 CTAs per SM
  if (chunks > B) chunks = B;
  if (chunks < 1) chunks = 1;
  const int per = (B + chunks - 1) / chunks;
  dim3 grid(T, (B + per - 1) / per);
  ProfScope prof(st, "table_grad B%d T%d H%d", B, T, H);
  if (dtype == EGOT2_F32) table_grad_kernel<float><<<grid, nt, 0, st>>>(B, T, H, per, (const float*)dy, dtable);
  else table_grad_kernel<bf16><<<grid, nt, 0, st>>>(B, T, H, per, (const bf16*)dy, dtable);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int colsum_accum(int dtype, int M, int N, const void* x, int ldx, int rpg, int gstride, float* out, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  const int col_blocks = (N + 63) / 64;
  int row_blocks = (sm_count() * 8 + col_blocks - 1) / col_blocks;
  int rows_per_cta = (M + row_blocks - 1) / row_blocks;
  if (rows_per_cta < 64) rows_per_cta = 64;
  row_blocks = (M + rows_per_cta - 1) / rows_per_cta;
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  ProfScope prof(st, "colsum M%d N%d", M, N);
  if (dtype == EGOT2_F32) colsum_kernel<float><<<grid, block, 0, st>>>(M, N, (const float*)x, ldx, rpg, gstride, out, rows_per_cta);
  else colsum_kernel<bf16><<<grid, block, 0, st>>>(M, N, (const bf16*)x, ldx, rpg, gstride, out, rows_per_cta);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

static inline int ew_grid(size_t n, int per_thread = 1) {
  size_t ctas = (n / per_thread + 255) / 256;
  size_t cap = (size_t)sm_count() * 16;
  if (ctas > cap) ctas = cap;
  return ctas ? (int)ctas : 1;
}

int dropout_inplace(int dtype, void* x, size_t n, float p, uint64_t key, cudaStream_t st) {
  if (p <= 0.f || n == 0) return 0;
  const float inv_keep = 1.f / (1.f - p);
  ProfScope prof(st, "dropout_inplace n%zu", n);
  if (dtype == EGOT2_F32) dropout_kernel<float><<<ew_grid(n), 256, 0, st>>>((float*)x, n, p, inv_keep, key);
  else dropout_kernel<bf16><<<ew_grid(n), 256, 0, st>>>((bf16*)x, n, p, inv_keep, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int pool_fwd(int dtype, int B, int T, int H, int pool, int row_tokens, const void* x, float* pooled, cudaStream_t st) {
  const int rows = pool ? B : B * row_tokens;
  if (rows == 0) return 0;
  const int nt = H < 256 ? ((H + 31) / 32 * 32) : 256;
  ProfScope prof(st, "pool_fwd B%d T%d H%d", B, T, H);
  if (dtype == EGOT2_F32) pool_fwd_kernel<float><<<rows, nt, 0, st>>>(B, T, H, pool, row_tokens, (const float*)x, pooled);
  else pool_fwd_kernel<bf16><<<rows, nt, 0, st>>>(B, T, H, pool, row_tokens, (const bf16*)x, pooled);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int pool_bwd(int dtype, int B, int T, int H, int pool, int row_tokens, const float* dpooled, void* dx, cudaStream_t st) {
  if (B * T == 0) return 0;
  const int nt = H < 256 ? ((H + 31) / 32 * 32) : 256;
  ProfScope prof(st, "pool_bwd B%d T%d H%d", B, T, H);
  if (dtype == EGOT2_F32) pool_bwd_kernel<float><<<B * T, nt, 0, st>>>(B, T, H, pool, row_tokens, dpooled, (float*)dx);
  else pool_bwd_kernel<bf16><<<B * T, nt, 0, st>>>(B, T, H, pool, row_tokens, dpooled, (bf16*)dx);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int cast_f32_to(int dtype, const float* src, void* dst, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ProfScope prof(st, "cast_f32_to n%zu", n);
  if (dtype == EGOT2_BF16) {
    EGOT2_CHECK(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), "cast: unaligned buffers");
    cast_to_bf16_kernel<<<ew_grid(n, 4), 256, 0, st>>>(src, (bf16*)dst, n);
  } else {
    copy_f32_kernel<<<ew_grid(n), 256, 0, st>>>(src, (float*)dst, n);
  }
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int cast_to_f32(int dtype, const void* src, float* dst, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ProfScope prof(st, "cast_to_f32 n%zu", n);
  if (dtype == EGOT2_BF16) cast_to_f32_kernel<<<ew_grid(n), 256, 0, st>>>((const bf16*)src, dst, n);
  else copy_f32_kernel<<<ew_grid(n), 256, 0, st>>>((const float*)src, dst, n);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int zero_f32(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  EGOT2_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), st));
  return 0;
}

}  // namespace egot2

// ---------------------------------------------------------------------------------- C ABI (utilities)
using namespace egot2;

extern "C" int egot2_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  return cast_f32_to(EGOT2_BF16, src, dst, n, (cudaStream_t)stream);
}
extern "C" int egot2_cast_bf16_to_f32(const void* src, float* dst, size_t n, void* stream) {
  if (n == 0) return 0;
  cast_to_f32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                               void* stream) {
  EGOT2_CHECK(step >= 1, "adam: step must be >= 1");
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  ProfScope prof((cudaStream_t)stream, "adam n%zu", n);
  adam_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                           weight_decay, bc1, sqrtf(bc2), grad_scale);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_hhi_tok_table_fwd(const float* task_embed, const float* pe, int32_t pe_len, int32_t n_seg,
                                       const int32_t* seg_tokens, const int32_t* seg_task_id, int32_t H,
                                       float* tok_table, void* stream) {
  EGOT2_CHECK(n_seg >= 1 && n_seg <= EGOT2_MAX_SEG, "tok_table: n_seg=%d", n_seg);
  SegList s; s.n = n_seg; int T = 0;
  for (int i = 0; i < n_seg; ++i) {
    EGOT2_CHECK(seg_tokens[i] <= pe_len, "tok_table: segment of %d tokens exceeds the pe table (%d)", seg_tokens[i], pe_len);
    s.tokens[i] = seg_tokens[i]; s.task[i] = seg_task_id[i]; T += seg_tokens[i];
  }
  if (T == 0) return 0;
  ProfScope prof((cudaStream_t)stream, "hhi_tok_table_fwd T%d H%d", T, H);
  hhi_tok_table_fwd_kernel<<<T, 128, 0, (cudaStream_t)stream>>>(task_embed, pe, s, H, tok_table);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
extern "C" int egot2_hhi_tok_table_bwd(const float* d_tok_table, int32_t n_seg, const int32_t* seg_tokens,
                                       const int32_t* seg_task_id, int32_t H, float* d_task_embed, void* stream) {
  EGOT2_CHECK(n_seg >= 1 && n_seg <= EGOT2_MAX_SEG, "tok_table: n_seg=%d", n_seg);
  SegList s; s.n = n_seg;
  for (int i = 0; i < n_seg; ++i) { s.tokens[i] = seg_tokens[i]; s.task[i] = seg_task_id[i]; }
  ProfScope prof((cudaStream_t)stream, "hhi_tok_table_bwd H%d", H);
  hhi_tok_table_bwd_kernel<<<n_seg, 128, 0, (cudaStream_t)stream>>>(d_tok_table, s, H, d_task_embed);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_slowfast_pool_fwd(const void* in, int32_t in_dtype, int32_t B, int32_t C, int32_t Tin, int32_t hw,
                                       int32_t Tout, void* out, int32_t out_dtype, void* stream) {
  EGOT2_CHECK(Tout > 0 && Tin % Tout == 0, "slowfast_pool: Tin=%d must be a multiple of Tout=%d", Tin, Tout);
  const size_t warps = (size_t)B * C * Tout;
  if (warps == 0) return 0;
  const int grid = (int)((warps * 32 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, "slowfast_pool B%d C%d Tin%d hw%d", B, C, Tin, hw);
  if (in_dtype == EGOT2_F32 && out_dtype == EGOT2_F32)
    slowfast_pool_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, B, C, Tin, hw, Tout, (float*)out);
  else if (in_dtype == EGOT2_F32 && out_dtype == EGOT2_BF16)
    slowfast_pool_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)in, B, C, Tin, hw, Tout, (bf16*)out);
  else if (in_dtype == EGOT2_BF16 && out_dtype == EGOT2_BF16)
    slowfast_pool_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)in, B, C, Tin, hw, Tout, (bf16*)out);
  else EGOT2_CHECK(false, "slowfast_pool: unsupported dtypes %d -> %d", in_dtype, out_dtype);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_layernorm_fwd(int32_t dtype, int32_t rows, int32_t H, const void* x, const float* g, const float* b,
                                   float eps, void* y, float* stat, void* stream) {
  LayerNormArgs a; a.rows = rows; a.H = H; a.dtype = dtype; a.x = x; a.g = g; a.b = b; a.eps = eps; a.y = y; a.stat = stat;
  return layernorm_fwd(a, (cudaStream_t)stream);
}
