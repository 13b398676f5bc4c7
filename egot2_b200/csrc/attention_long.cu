// attention_long.cu — tensor-core self-attention for clips of 129 .. 512 tokens (bf16 mma.sync m16n8k16, fp32 softmax):
// the HHI translators at more than 42 frames per task (HHI/dataset/ttm/data_loader_2task.py:119,150-162 reaches 453 tokens;
// PositionalEncoding allows 1000, HHI/models/ttm/model_taskspecific.py:137).  attention_mma.cu keeps a warp's whole 16 x T
// score tile in registers, which stops at T = 128; here
//   * one CTA per (clip, head) still stages the clip's Q, K, V (and dO) ONCE in shared memory (head dim 32: T <= 512,
//     head dim 64: T <= 384 - the backward's four matrices must fit 227 KB),
//   * 8 warps walk the 16-row tiles of the clip, and every tile is streamed over the keys in blocks with a fixed register
//     footprint: forward = online softmax over 32-key blocks (running row maximum / sum, accumulator rescaled per block),
//     backward = the two passes of attention_mma.cu over 16-key / 16-query blocks (P recomputed from the saved lse, so
//     nothing needs rescaling),
//   * dropout masks are the same pure function of (key, clip, head, query, key index) as everywhere else (common.cuh
//     attn_drop_keep), generated one 32-key word at a time.
// Same lse / dropout / layout conventions as attention_mma.cu and attention_simt.cu: the three are interchangeable.
#include <math.h>

#define EGOT2_FILE_ID 17
#include "ops.h"

namespace egot2 {
namespace {

constexpr int NWARPS = 8;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH> struct Tile {
  static constexpr int LD = DH + 8;          // padded row length (elements): conflict-free ldmatrix
  static constexpr int KS = DH / 16;
  static constexpr int ND = DH / 8;
};

template <int DH>
__device__ __forceinline__ void load_a(uint32_t sbase, int r0, uint32_t (&a)[Tile<DH>::KS][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < Tile<DH>::KS; ++ks)
    ldsm_x4(sbase + (uint32_t)(((r0 + (lane & 15)) * Tile<DH>::LD + ks * 16 + 8 * (lane >> 4)) * 2), a[ks]);
}
// acc2 (16 x 16) += A(16 x DH) . Bs[kb*16 .. kb*16+15]^T   (row index of Bs = output column)
template <int DH>
__device__ __forceinline__ void gemm_rc_kb(float (&acc2)[2][4], const uint32_t (&a)[Tile<DH>::KS][4], uint32_t sB, int kb) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < Tile<DH>::KS; ++ks) {
    uint32_t b[4];
    ldsm_x4(sB + (uint32_t)(((kb * 16 + (lane & 7) + 8 * (lane >> 4)) * Tile<DH>::LD + ks * 16 + 8 * ((lane >> 3) & 1)) * 2), b);
    mma16816(acc2[0], a[ks], b[0], b[1]);
    mma16816(acc2[1], a[ks], b[2], b[3]);
  }
}
// o (16 x DH) += P_kb(16 x 16) . Bs[kb*16 .. kb*16+15]   (row index of Bs = reduction index)
template <int DH>
__device__ __forceinline__ void gemm_pv_kb(float (&o)[Tile<DH>::ND][4], const uint32_t (&p)[4], uint32_t sB, int kb) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d2 = 0; d2 < Tile<DH>::ND / 2; ++d2) {
    uint32_t b[4];
    ldsm_x4_t(sB + (uint32_t)(((kb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * Tile<DH>::LD + d2 * 16 + 8 * (lane >> 4)) * 2), b);
    mma16816(o[2 * d2], p, b[0], b[1]);
    mma16816(o[2 * d2 + 1], p, b[2], b[3]);
  }
}

// rows [0, TP) x DH of a strided global matrix -> smem [TP][LD]; rows >= T are zero (cp.async zero-fill)
template <int DH>
__device__ __forceinline__ void stage(uint32_t dst, const bf16* __restrict__ src, int ld_src, int T, int TP) {
  constexpr int CH = DH / 8;                 // 16-byte chunks per row
  for (int i = threadIdx.x; i < TP * CH; i += NWARPS * 32) {
    const int r = i / CH, c = (i % CH) * 8;
    const bool ok = r < T;
    const bf16* g = src + (size_t)(ok ? r : 0) * ld_src + c;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)((r * Tile<DH>::LD + c) * 2)), "l"(g),
                 "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ forward
// MODE 0 no dropout, 1 p == 0.5 (one random bit per pair, kept entries' factor 2 folded into the row scale), 2 general p
template <int DH, int MODE>
__global__ void __launch_bounds__(NWARPS * 32, 1) attn_long_fwd_kernel(int T, int TP, int H, int heads, const bf16* __restrict__ qkv,
                                                                       bf16* __restrict__ out, float* __restrict__ lse,
                                                                       float p_drop, uint64_t drop_key) {
  constexpr int LD = Tile<DH>::LD;
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(smem), uK = uQ + TP * LD * 2, uV = uK + TP * LD * 2;
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + (size_t)b * T * 3 * H + h * DH;
  EGOT2_PDL_ENTER();
  stage<DH>(uQ, base, 3 * H, T, TP);
  stage<DH>(uK, base + H, 3 * H, T, TP);
  stage<DH>(uV, base + 2 * H, 3 * H, T, TP);
  stage_wait();
  __syncthreads();
  const float scn = rsqrtf((float)DH);
  const float sc = scn * 1.4426950408889634f;          // scale * log2(e)
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int wpr = (T + 31) >> 5;                       // 32-key mask words per query row
  bf16* ob = out + (size_t)b * T * H + h * DH;
  for (int r0 = warp * 16; r0 < T; r0 += NWARPS * 16) {
    uint32_t aq[Tile<DH>::KS][4];
    load_a<DH>(uQ, r0, aq);
    const int q0 = r0 + (lane >> 2), q1 = q0 + 8;      // this thread's two query rows
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // running maximum (raw scores) / partial sums (this thread's columns)
    float o[Tile<DH>::ND][4];
#pragma unroll
    for (int i = 0; i < Tile<DH>::ND; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll 1
    for (int k32 = 0; k32 < wpr; ++k32) {              // 32 keys per iteration = two 16-key blocks = four 8-key tiles
      float sa[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, sb[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      gemm_rc_kb<DH>(sa, aq, uK, 2 * k32);             // TP is a multiple of 32: both 16-key blocks exist (zero rows past T)
      gemm_rc_kb<DH>(sb, aq, uK, 2 * k32 + 1);
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float (&sn)[4] = nt < 2 ? sa[nt & 1] : sb[nt & 1];
        const int c = k32 * 32 + nt * 8 + 2 * (lane & 3);
        if (c >= T) { sn[0] = -INFINITY; sn[2] = -INFINITY; }
        if (c + 1 >= T) { sn[1] = -INFINITY; sn[3] = -INFINITY; }
        bm0 = fmaxf(bm0, fmaxf(sn[0], sn[1]));
        bm1 = fmaxf(bm1, fmaxf(sn[2], sn[3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);          // finite: every block holds at least one real key
      const float a0 = ex2((m0 - n0) * sc), a1 = ex2((m1 - n1) * sc); // first block: ex2(-inf) = 0
      m0 = n0; m1 = n1;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int i = 0; i < Tile<DH>::ND; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
      const float ms0 = m0 * sc, ms1 = m1 * sc;
      uint32_t w0 = 0, w1 = 0;
      if (MODE == 1) {
        w0 = drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + q0) * (uint64_t)wpr + k32) >> (2 * (lane & 3));
        w1 = drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + q1) * (uint64_t)wpr + k32) >> (2 * (lane & 3));
      }
      uint32_t p[2][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float (&sn)[4] = nt < 2 ? sa[nt & 1] : sb[nt & 1];
        float e0 = ex2(fmaf(sn[0], sc, -ms0)), e1 = ex2(fmaf(sn[1], sc, -ms0));
        float e2 = ex2(fmaf(sn[2], sc, -ms1)), e3 = ex2(fmaf(sn[3], sc, -ms1));
        l0 += e0 + e1; l1 += e2 + e3;
        if (MODE == 1) {
          const uint32_t k0 = w0 >> (nt * 8), k1 = w1 >> (nt * 8);
          e0 = (k0 & 1u) ? e0 : 0.f; e1 = (k0 & 2u) ? e1 : 0.f;
          e2 = (k1 & 1u) ? e2 : 0.f; e3 = (k1 & 2u) ? e3 : 0.f;
        } else if (MODE == 2) {
          const int c = k32 * 32 + nt * 8 + 2 * (lane & 3);
          attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + q0, T, c, p_drop, inv_keep, e0, e1);
          attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + q1, T, c, p_drop, inv_keep, e2, e3);
        }
        p[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
        p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
      }
      gemm_pv_kb<DH>(o, p[0], uV, 2 * k32);
      gemm_pv_kb<DH>(o, p[1], uV, 2 * k32 + 1);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float fold = MODE == 1 ? 2.f : 1.f;
    const float i0 = fold / l0, i1 = fold / l1;
#pragma unroll
    for (int nd = 0; nd < Tile<DH>::ND; ++nd) {
      const int c = nd * 8 + 2 * (lane & 3);
      if (q0 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)q0 * H + c) = pack_bf16(o[nd][0] * i0, o[nd][1] * i0);
      if (q1 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)q1 * H + c) = pack_bf16(o[nd][2] * i1, o[nd][3] * i1);
    }
    if ((lane & 3) == 0) {
      if (q0 < T) lse[(size_t)bh * T + q0] = m0 * scn + logf(l0);
      if (q1 < T) lse[(size_t)bh * T + q1] = m1 * scn + logf(l1);
    }
  }
}

// ------------------------------------------------------------------ backward (the two passes of attention_mma.cu, runtime loops)
template <int DH, int MODE>
__global__ void __launch_bounds__(NWARPS * 32, 1) attn_long_bwd_kernel(int T, int TP, int H, int heads, const bf16* __restrict__ qkv,
                                                                       const bf16* __restrict__ out, const float* __restrict__ lse,
                                                                       const bf16* __restrict__ dout, bf16* __restrict__ dqkv,
                                                                       float p_drop, uint64_t drop_key) {
  constexpr int LD = Tile<DH>::LD;
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(smem), uK = uQ + TP * LD * 2, uV = uK + TP * LD * 2,
                 udO = uV + TP * LD * 2;
  float* sL = reinterpret_cast<float*>(smem + (size_t)4 * TP * LD * 2);      // lse * log2(e) per query
  float* sD = sL + TP;                                                       // D = rowsum(dO * O) per query
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* base = qkv + (size_t)b * T * 3 * H + h * DH;
  const bf16* ob = out + (size_t)b * T * H + h * DH;
  const bf16* dob = dout + (size_t)b * T * H + h * DH;
  constexpr float l2e = 1.4426950408889634f;
  EGOT2_PDL_ENTER();
  stage<DH>(uQ, base, 3 * H, T, TP);
  stage<DH>(uK, base + H, 3 * H, T, TP);
  stage<DH>(uV, base + 2 * H, 3 * H, T, TP);
  stage<DH>(udO, dob, H, T, TP);
  // D_i and lse_i while the copies are in flight: 4 lanes per query row, 16 B loads
  for (int r = threadIdx.x >> 2; r < TP; r += NWARPS * 8) {
    float d = 0.f;
    if (r < T) {
      for (int c = (threadIdx.x & 3) * 8; c < DH; c += 32) {
        const uint4 a4 = *reinterpret_cast<const uint4*>(dob + (size_t)r * H + c);
        const uint4 o4 = *reinterpret_cast<const uint4*>(ob + (size_t)r * H + c);
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, ow[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          d = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(ow[k] << 16), d);
          d = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(ow[k] & 0xffff0000u), d);
        }
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    if ((threadIdx.x & 3) == 0) { sD[r] = d; sL[r] = r < T ? lse[(size_t)bh * T + r] * l2e : 0.f; }
  }
  stage_wait();
  __syncthreads();
  const float scn = rsqrtf((float)DH);
  const float sc2 = scn * l2e;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const int wpr = (T + 31) >> 5;
  const int nkb = (T + 15) >> 4;                       // 16-wide blocks of the clip
  bf16* dq = dqkv + (size_t)b * T * 3 * H + h * DH;

  for (int r0 = warp * 16; r0 < T; r0 += NWARPS * 16) {
    const int ra = r0 + (lane >> 2), rb = ra + 8;
    // ---------------- pass 1: rows = queries -> dQ
    {
      uint32_t a1[Tile<DH>::KS][4], a2[Tile<DH>::KS][4];
      load_a<DH>(uQ, r0, a1);
      load_a<DH>(udO, r0, a2);
      const float la = sL[ra], lb = sL[rb], da = sD[ra], db = sD[rb];
      uint32_t wa = 0, wb = 0;
      float o[Tile<DH>::ND][4];
#pragma unroll
      for (int i = 0; i < Tile<DH>::ND; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        if (MODE == 1 && (kb & 1) == 0) {
          wa = drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + ra) * (uint64_t)wpr + (kb >> 1)) >> (2 * (lane & 3));
          wb = drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + rb) * (uint64_t)wpr + (kb >> 1)) >> (2 * (lane & 3));
        }
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        gemm_rc_kb<DH>(s, a1, uK, kb);
        gemm_rc_kb<DH>(dp, a2, uV, kb);
        uint32_t ds[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float p0 = ex2(fmaf(s[h2][0], sc2, -la)), p1 = ex2(fmaf(s[h2][1], sc2, -la));
          float p2 = ex2(fmaf(s[h2][2], sc2, -lb)), p3 = ex2(fmaf(s[h2][3], sc2, -lb));
          const int c = kb * 16 + h2 * 8 + 2 * (lane & 3);
          if (c >= T) { p0 = 0.f; p2 = 0.f; }
          if (c + 1 >= T) { p1 = 0.f; p3 = 0.f; }
          float g0 = dp[h2][0], g1 = dp[h2][1], g2 = dp[h2][2], g3 = dp[h2][3];
          if (MODE == 1) {
            const uint32_t k0 = wa >> ((kb & 1) * 16 + h2 * 8), k1 = wb >> ((kb & 1) * 16 + h2 * 8);
            g0 = (k0 & 1u) ? g0 + g0 : 0.f; g1 = (k0 & 2u) ? g1 + g1 : 0.f;
            g2 = (k1 & 1u) ? g2 + g2 : 0.f; g3 = (k1 & 2u) ? g3 + g3 : 0.f;
          } else if (MODE == 2) {
            attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + ra, T, c, p_drop, inv_keep, g0, g1);
            attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + rb, T, c, p_drop, inv_keep, g2, g3);
          }
          ds[h2 * 2] = pack_bf16(p0 * (g0 - da), p1 * (g1 - da));
          ds[h2 * 2 + 1] = pack_bf16(p2 * (g2 - db), p3 * (g3 - db));
        }
        gemm_pv_kb<DH>(o, ds, uK, kb);
      }
#pragma unroll
      for (int nd = 0; nd < Tile<DH>::ND; ++nd) {
        const int c = nd * 8 + 2 * (lane & 3);
        if (ra < T) *reinterpret_cast<uint32_t*>(dq + (size_t)ra * 3 * H + c) = pack_bf16(o[nd][0] * scn, o[nd][1] * scn);
        if (rb < T) *reinterpret_cast<uint32_t*>(dq + (size_t)rb * 3 * H + c) = pack_bf16(o[nd][2] * scn, o[nd][3] * scn);
      }
    }
    // ---------------- pass 2: rows = keys -> dK, dV   (tile element (r, c) = (key r, query c), streamed over query blocks)
    {
      uint32_t a1[Tile<DH>::KS][4], a2[Tile<DH>::KS][4];
      load_a<DH>(uK, r0, a1);
      load_a<DH>(uV, r0, a2);
      uint32_t wq = 0;       // p == 0.5: lane L holds the mask word of (query 32*j + L, this tile's 32-key block); fetched by shuffle
      float ov[Tile<DH>::ND][4], ok[Tile<DH>::ND][4];
#pragma unroll
      for (int i = 0; i < Tile<DH>::ND; ++i) { ov[i][0] = ov[i][1] = ov[i][2] = ov[i][3] = 0.f; ok[i][0] = ok[i][1] = ok[i][2] = ok[i][3] = 0.f; }
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        if (MODE == 1 && (kb & 1) == 0) {
          const int qy = lane + 32 * (kb >> 1);
          wq = qy < T ? drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + qy) * (uint64_t)wpr + (uint32_t)(r0 >> 5)) : 0u;
        }
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        gemm_rc_kb<DH>(s, a1, uQ, kb);
        gemm_rc_kb<DH>(dp, a2, udO, kb);
        uint32_t pf[4], ds[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int c = kb * 16 + h2 * 8 + 2 * (lane & 3);
          const float2 l01 = *reinterpret_cast<const float2*>(sL + c), d01 = *reinterpret_cast<const float2*>(sD + c);
          float p0 = ex2(fmaf(s[h2][0], sc2, -l01.x)), p1 = ex2(fmaf(s[h2][1], sc2, -l01.y));
          float p2 = ex2(fmaf(s[h2][2], sc2, -l01.x)), p3 = ex2(fmaf(s[h2][3], sc2, -l01.y));
          if (c >= T) { p0 = 0.f; p2 = 0.f; }
          if (c + 1 >= T) { p1 = 0.f; p3 = 0.f; }
          float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
          if (MODE == 1) {
            const uint32_t u0 = __shfl_sync(0xffffffffu, wq, c & 31) >> (ra & 31);      // bit 0: key ra, bit 8: key rb
            const uint32_t u1 = __shfl_sync(0xffffffffu, wq, (c + 1) & 31) >> (ra & 31);
            k0 = (u0 & 1u) ? 2.f : 0.f;
            k1 = (u1 & 1u) ? 2.f : 0.f;
            k2 = (u0 & 0x100u) ? 2.f : 0.f;
            k3 = (u1 & 0x100u) ? 2.f : 0.f;
          } else if (MODE == 2) {
            k0 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c, T, ra, p_drop, inv_keep);
            k1 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c + 1, T, ra, p_drop, inv_keep);
            k2 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c, T, rb, p_drop, inv_keep);
            k3 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c + 1, T, rb, p_drop, inv_keep);
          }
          pf[h2 * 2] = pack_bf16(p0 * k0, p1 * k1);
          pf[h2 * 2 + 1] = pack_bf16(p2 * k2, p3 * k3);
          ds[h2 * 2] = pack_bf16(p0 * fmaf(dp[h2][0], k0, -d01.x), p1 * fmaf(dp[h2][1], k1, -d01.y));
          ds[h2 * 2 + 1] = pack_bf16(p2 * fmaf(dp[h2][2], k2, -d01.x), p3 * fmaf(dp[h2][3], k3, -d01.y));
        }
        gemm_pv_kb<DH>(ov, pf, udO, kb);
        gemm_pv_kb<DH>(ok, ds, uQ, kb);
      }
#pragma unroll
      for (int nd = 0; nd < Tile<DH>::ND; ++nd) {
        const int c = nd * 8 + 2 * (lane & 3);
        if (ra < T) {
          *reinterpret_cast<uint32_t*>(dq + (size_t)ra * 3 * H + H + c) = pack_bf16(ok[nd][0] * scn, ok[nd][1] * scn);
          *reinterpret_cast<uint32_t*>(dq + (size_t)ra * 3 * H + 2 * H + c) = pack_bf16(ov[nd][0], ov[nd][1]);
        }
        if (rb < T) {
          *reinterpret_cast<uint32_t*>(dq + (size_t)rb * 3 * H + H + c) = pack_bf16(ok[nd][2] * scn, ok[nd][3] * scn);
          *reinterpret_cast<uint32_t*>(dq + (size_t)rb * 3 * H + 2 * H + c) = pack_bf16(ov[nd][2], ov[nd][3]);
        }
      }
    }
  }
}

inline int drop_mode(float p) { return p <= 0.f ? 0 : (p == 0.5f ? 1 : 2); }
inline int padded(int T) { return (T + 31) / 32 * 32; }      // whole 32-key iterations of the forward
template <int DH> size_t bwd_smem(int TP) { return (size_t)4 * TP * Tile<DH>::LD * 2 + 2 * (size_t)TP * 4; }

template <int DH>
int launch_fwd(int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p, uint64_t key, cudaStream_t st) {
  const int TP = padded(T);
  const size_t smem = (size_t)3 * TP * Tile<DH>::LD * 2;
  typedef void (*Kern)(int, int, int, int, const bf16*, bf16*, float*, float, uint64_t);
  static const Kern kerns[3] = {attn_long_fwd_kernel<DH, 0>, attn_long_fwd_kernel<DH, 1>, attn_long_fwd_kernel<DH, 2>};
  static bool set = false;
  if (!set) {
    for (int i = 0; i < 3; ++i) EGOT2_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  ProfScope prof(st, "attn_long_fwd<dh%d> B%d T%d H%d", DH, B, T, H);
  launch(kerns[drop_mode(p)], dim3(B * heads), dim3(NWARPS * 32), smem, st, T, TP, H, heads, (const bf16*)qkv, (bf16*)out, lse, p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
template <int DH>
int launch_bwd(int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
               void* dqkv, float p, uint64_t key, cudaStream_t st) {
  const int TP = padded(T);
  const size_t smem = bwd_smem<DH>(TP);
  typedef void (*Kern)(int, int, int, int, const bf16*, const bf16*, const float*, const bf16*, bf16*, float, uint64_t);
  static const Kern kerns[3] = {attn_long_bwd_kernel<DH, 0>, attn_long_bwd_kernel<DH, 1>, attn_long_bwd_kernel<DH, 2>};
  static bool set = false;
  if (!set) {
    for (int i = 0; i < 3; ++i) EGOT2_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  ProfScope prof(st, "attn_long_bwd<dh%d> B%d T%d H%d", DH, B, T, H);
  launch(kerns[drop_mode(p)], dim3(B * heads), dim3(NWARPS * 32), smem, st, T, TP, H, heads, (const bf16*)qkv, (const bf16*)out, lse,
         (const bf16*)dout, (bf16*)dqkv, p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// bf16, head dim 16 / 32 / 64, 128 < T and the backward's four staged matrices within 227 KB (dh 32: T <= 512, dh 64: T <= 384)
bool attention_long_supported(int dtype, int T, int H, int heads) {
  if (dtype != EGOT2_BF16 || heads <= 0 || H % heads || H % 8) return false;
  const int dh = H / heads;
  if (!(dh == 16 || dh == 32 || dh == 64) || T <= 128 || T > 512) return false;
  const int TP = padded(T);
  const size_t need = dh == 16 ? bwd_smem<16>(TP) : (dh == 32 ? bwd_smem<32>(TP) : bwd_smem<64>(TP));
  return need <= 227 * 1024;
}

int attention_long_fwd(int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p_drop, uint64_t drop_key,
                       cudaStream_t st) {
  if (B * T == 0) return 0;
  switch (H / heads) {
    case 16: return launch_fwd<16>(B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
    case 32: return launch_fwd<32>(B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
    case 64: return launch_fwd<64>(B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
  }
  EGOT2_CHECK(false, "attention_long_fwd: unsupported head dim %d", H / heads);
}
int attention_long_bwd(int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
                       void* dqkv, float p_drop, uint64_t drop_key, cudaStream_t st) {
  if (B * T == 0) return 0;
  switch (H / heads) {
    case 16: return launch_bwd<16>(B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
    case 32: return launch_bwd<32>(B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
    case 64: return launch_bwd<64>(B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
  }
  EGOT2_CHECK(false, "attention_long_bwd: unsupported head dim %d", H / heads);
}

}  // namespace egot2
