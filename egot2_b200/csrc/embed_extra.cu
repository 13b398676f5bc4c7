// embed_extra.cu — elementwise variants of the embed stage used by sibling translators only:
//   add_table:               tokens = projected features + positional table, NO LayerNorm
//                            (2-task simple_vit translator, HOI/models/pnr/video_model_transfer.py:63)
//   dropout_prefix_inplace:  feature dropout on the leading tokens of every clip only
//                            (2-task PNR translator with FEAT_DROPOUT_MODE > 0 drops the PNR segment alone, :95-96)
#define EGOT2_FILE_ID 14
#include "ops.h"

namespace egot2 {
namespace {

template <typename T>
__global__ void add_table_kernel(const T* __restrict__ z, const float* __restrict__ table, T* __restrict__ x, size_t n,
                                 size_t table_elems) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = from_f32<T>(to_f32(z[i]) + table[i % table_elems]);
}

template <typename T>
__global__ void dropout_prefix_kernel(T* __restrict__ x, size_t n, size_t period, size_t prefix, float p, float inv_keep,
                                      uint64_t key) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (i % period < prefix) x[i] = from_f32<T>(to_f32(x[i]) * drop_scale(key ^ egot2_ep, i, p, inv_keep));
}

// ---- embed_finish: everything of the embed stage that follows the projections, in ONE pass over the tokens.
// The long-K projections (PNR / OSCC: K = 8192 onto H = 128, M = clips x 16 rows: a 32-tile GEMM that cannot fill 148 SMs)
// run a DETERMINISTIC split-K: split s of segment k writes its partial product (bias added by the first split) to slab s, an
// fp32 (B*D_k, H) matrix (the slabs of PNR b256 are 32 MB: they stay in the 126 MB L2); this kernel sums the slabs in a fixed
// order - no atomics, bit-reproducible - reads pass-through segments (already H wide) in place, and applies feature dropout -> z (bf16, the LayerNorm input saved for backward) -> LayerNorm -> + token table -> embedding
// dropout -> x.  One warp per token row, S = H / 128 segments of 4 consecutive columns per lane (16 B loads, 8 B stores).
// HOI/models/pnr/video_model_transfer_3task.py:249-253 (dp(proj), cat, ln, + pe); HHI model_taskspecific.py:217-222.
template <int S>
__global__ void __launch_bounds__(256) embed_finish_kernel(const EmbedSrc src, int B, int T, int drop_tokens,
                                                           float p_feat, uint64_t key_feat, const float* __restrict__ g,
                                                           const float* __restrict__ b, float eps,
                                                           const float* __restrict__ table, float p_embed, uint64_t key_embed,
                                                           bf16* __restrict__ z, float* __restrict__ stat, bf16* __restrict__ x) {
  EGOT2_PDL_ENTER();
  constexpr int HH = S * 128;
  const int lane = threadIdx.x & 31;
  const float ik_f = p_feat > 0.f ? 1.f / (1.f - p_feat) : 1.f, ik_e = p_embed > 0.f ? 1.f / (1.f - p_embed) : 1.f;
  float gg[S][4], bb[S][4];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const float4 g4 = *reinterpret_cast<const float4*>(g + s * 128 + lane * 4), b4 = *reinterpret_cast<const float4*>(b + s * 128 + lane * 4);
    gg[s][0] = g4.x; gg[s][1] = g4.y; gg[s][2] = g4.z; gg[s][3] = g4.w;
    bb[s][0] = b4.x; bb[s][1] = b4.y; bb[s][2] = b4.z; bb[s][3] = b4.w;
  }
  const int warps = gridDim.x * (blockDim.x >> 5), rows = B * T;
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const int t = row % T, bclip = row / T;
    int k = 0;
#pragma unroll
    for (int i = 1; i < EGOT2_MAX_SEG; ++i) if (i < src.n && t >= src.tok_begin[i] && src.tokens[i] > 0) k = i;
    const size_t srow = (size_t)bclip * src.tokens[k] + (t - src.tok_begin[k]);       // row inside the segment's (B*D_k, H) matrix
    const size_t slab_elems = (size_t)B * src.tokens[k] * HH;
    const float* slab = src.slab[k];
    const bf16* direct = (const bf16*)src.direct[k];
    const int ns = src.splits[k];
    const bool fdrop = p_feat > 0.f && (drop_tokens <= 0 || t < drop_tokens);
    float v[S][4], sum = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int c0 = s * 128 + lane * 4;
      float w[4];
      if (direct && src.direct_f32[k]) {
        const float4 a4 = *reinterpret_cast<const float4*>((const float*)src.direct[k] + srow * HH + c0);
        w[0] = a4.x; w[1] = a4.y; w[2] = a4.z; w[3] = a4.w;
      } else if (direct) {
        const uint2 dw = *reinterpret_cast<const uint2*>(direct + srow * HH + c0);
        w[0] = __uint_as_float(dw.x << 16); w[1] = __uint_as_float(dw.x & 0xffff0000u);
        w[2] = __uint_as_float(dw.y << 16); w[3] = __uint_as_float(dw.y & 0xffff0000u);
      } else {
        const float4 a4 = *reinterpret_cast<const float4*>(slab + srow * HH + c0);
        w[0] = a4.x; w[1] = a4.y; w[2] = a4.z; w[3] = a4.w;
        for (int sp = 1; sp < ns; ++sp) {
          const float4 p4 = *reinterpret_cast<const float4*>(slab + sp * slab_elems + srow * HH + c0);
          w[0] += p4.x; w[1] += p4.y; w[2] += p4.z; w[3] += p4.w;
        }
      }
      if (fdrop) {
        float dmf[4];
        drop_scale_n<4>(key_feat ^ egot2_ep, (uint64_t)row * HH + c0, p_feat, ik_f, dmf);
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] *= dmf[i];
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(w[0], w[1]), p1 = __floats2bfloat162_rn(w[2], w[3]);
      uint2 zw; zw.x = *reinterpret_cast<uint32_t*>(&p0); zw.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(z + (size_t)row * HH + c0) = zw;
      // LayerNorm sees the rounded values (what backward re-reads)
      v[s][0] = __uint_as_float(zw.x << 16); v[s][1] = __uint_as_float(zw.x & 0xffff0000u);
      v[s][2] = __uint_as_float(zw.y << 16); v[s][3] = __uint_as_float(zw.y & 0xffff0000u);
      sum += (v[s][0] + v[s][1]) + (v[s][2] + v[s][3]);
    }
    const float mean = warp_sum(sum) * (1.f / HH);
    float q = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float d = v[s][i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / HH) + eps);
    if (stat && lane == 0) { stat[2 * (size_t)row] = mean; stat[2 * (size_t)row + 1] = rstd; }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int c0 = s * 128 + lane * 4;
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = (v[s][i] - mean) * rstd * gg[s][i] + bb[s][i];
      if (table) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(table + (size_t)t * HH + c0));
        o[0] += t4.x; o[1] += t4.y; o[2] += t4.z; o[3] += t4.w;
      }
      if (p_embed > 0.f) {
        float dme[4];
        drop_scale_n<4>(key_embed ^ egot2_ep, (uint64_t)row * HH + c0, p_embed, ik_e, dme);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] *= dme[i];
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
      uint2 xw; xw.x = *reinterpret_cast<uint32_t*>(&p0); xw.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(x + (size_t)row * HH + c0) = xw;
    }
  }
}

// dst += a (+ b), then a (and b) are cleared: the per-branch gradient arenas of an EgoT2-g step (three forwards of one model
// running side by side on three streams) meet in the shared arena and are left clean for the next step, one pass
__global__ void __launch_bounds__(256) sum_into_clear_kernel(float* __restrict__ dst, float* __restrict__ a, float* __restrict__ b, size_t n4) {
  EGOT2_PDL_ENTER();
  // every thread owns 4 float4 per pass and requests all of its (up to 12) loads before the first store: written as
  // load-store-load-store per element the pass ran at 1.2 TB/s (175 us for the 8.9 M parameters of the HHI EgoT2-g model)
  for (size_t c = blockIdx.x; c * 1024 < n4; c += gridDim.x) {
    const size_t i0 = c * 1024 + threadIdx.x;
    float4 d[4], x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t i = i0 + (size_t)k * 256;
      if (i < n4) {
        d[k] = __ldcs(reinterpret_cast<const float4*>(dst) + i);
        x[k] = __ldcs(reinterpret_cast<const float4*>(a) + i);
        y[k] = b ? __ldcs(reinterpret_cast<const float4*>(b) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t i = i0 + (size_t)k * 256;
      if (i < n4) {
        d[k].x += x[k].x + y[k].x; d[k].y += x[k].y + y[k].y; d[k].z += x[k].z + y[k].z; d[k].w += x[k].w + y[k].w;
        reinterpret_cast<float4*>(dst)[i] = d[k];
        reinterpret_cast<float4*>(a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b) reinterpret_cast<float4*>(b)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

int ew_grid(size_t n) {
  size_t ctas = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (ctas > cap) ctas = cap;
  return ctas ? (int)ctas : 1;
}

}  // namespace

int add_table(int dt, size_t n, size_t table_elems, const void* z, const float* table, void* x, cudaStream_t st) {
  if (n == 0) return 0;
  EGOT2_CHECK(table_elems > 0 && table, "add_table: empty table");
  ProfScope prof(st, "add_table n%zu", n);
  if (dt == EGOT2_F32) launch(add_table_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (const float*)z, table, (float*)x, n, table_elems);
  else launch(add_table_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (const bf16*)z, table, (bf16*)x, n, table_elems);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int sum_into_clear(float* dst, float* a, float* b, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  EGOT2_CHECK(dst && a && n % 4 == 0, "sum_into_clear: buffers / n %% 4");
  ProfScope prof(st, "sum_into_clear n%zu", n);
  // a pure streaming kernel (three reads, three writes per element): enough CTAs to keep every SM's load queue full
  size_t ctas = (n / 4 + 1023) / 1024;
  const size_t cap = (size_t)sm_count() * 8;
  launch(sum_into_clear_kernel, dim3((unsigned)(ctas > cap ? cap : ctas)), dim3(256), 0, st, dst, a, b, n / 4);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

bool embed_finish_supported(int dt, int H) { return dt == EGOT2_BF16 && (H == 128 || H == 256 || H == 512 || H == 1024); }

int embed_finish(int B, int T, int H, const EmbedSrc& src, int drop_tokens, float p_feat, uint64_t key_feat, const float* g,
                 const float* b, float eps, const float* table, float p_embed, uint64_t key_embed, void* z, float* stat, void* x,
                 cudaStream_t st) {
  const int rows = B * T;
  if (rows == 0) return 0;
  EGOT2_CHECK(embed_finish_supported(EGOT2_BF16, H), "embed_finish: H=%d not supported", H);
  ProfScope prof(st, "embed_finish rows%d H%d", rows, H);
  int grid = (rows + 7) / 8;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
#define EGOT2_EF(S) launch(embed_finish_kernel<S>, dim3(grid), dim3(256), 0, st, src, B, T, drop_tokens, p_feat, key_feat, g, b, \
                           eps, table, p_embed, key_embed, (bf16*)z, stat, (bf16*)x)
  switch (H) {
    case 128: EGOT2_EF(1); break;
    case 256: EGOT2_EF(2); break;
    case 512: EGOT2_EF(4); break;
    default: EGOT2_EF(8); break;
  }
#undef EGOT2_EF
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int dropout_prefix_inplace(int dt, void* x, size_t n, size_t period, size_t prefix, float p, uint64_t key, cudaStream_t st) {
  if (p <= 0.f || n == 0 || prefix == 0) return 0;
  EGOT2_CHECK(period > 0 && prefix <= period && p < 1.f, "dropout_prefix: bad geometry (period %zu prefix %zu p %f)", period, prefix, p);
  const float inv_keep = 1.f / (1.f - p);
  ProfScope prof(st, "dropout_prefix n%zu", n);
  if (dt == EGOT2_F32) launch(dropout_prefix_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (float*)x, n, period, prefix, p, inv_keep, key);
  else launch(dropout_prefix_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (bf16*)x, n, period, prefix, p, inv_keep, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
