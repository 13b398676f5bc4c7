// attention_mma.cu — short-sequence self-attention on tensor cores (bf16 mma.sync m16n8k16, fp32 softmax).
//
// One CTA per (clip, head).  The clip's whole token set for that head (Q, K, V, and in backward dO) is staged
// ONCE into shared memory (rows padded by 8 bf16 -> conflict-free ldmatrix); each warp owns a 16-row tile and keeps
// its full 16 x T score tile in registers, so there is no online-softmax rescaling, no T x T traffic, and no
// cross-warp reduction or atomic anywhere:
//   forward            S = Q_i K^T, P = softmax(scale S) (quad shuffles), O_i = P V, lse
//   backward, pass 1   rows = queries: recompute P from lse, dP = dO_i V^T, dS = P (dP - D), dQ_i = scale dS K
//   backward, pass 2   rows = keys:    S^T = K_j Q^T, dP^T = V_j dO^T, dV_j = P^T dO, dK_j = scale dS^T Q
// T <= 128 with head dims 16 / 32 / 64 (the HOI translators: 48 / 8 tokens; HHI at D <= 42 frames x 3 tasks) and T <= 32 with
// head dim 128 (LTA at H = 1024); longer clips use the shape-general kernels of attention_simt.cu.  Dropout masks are regenerated from (key, b, h, query, key index).
#include <math.h>
#include <stdlib.h>

#define EGOT2_FILE_ID 6
#include "ops.h"

#ifndef EGOT2_ATTN_MINB
#define EGOT2_ATTN_MINB 4      // resident CTAs per SM the register allocation aims at (T <= 96 variants)
#endif

namespace egot2 {

namespace {

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

template <int DH> struct Tile {
  static constexpr int LD = DH + 8;          // padded row length (elements)
  static constexpr int KS = DH / 16;         // k16 steps over the head dim
  static constexpr int ND = DH / 8;          // n8 tiles over the head dim
};

// A fragments of the 16 x DH row tile starting at row r0 of a [rows][LD] smem matrix
template <int DH>
__device__ __forceinline__ void load_a(uint32_t sbase, int r0, uint32_t (&a)[Tile<DH>::KS][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < Tile<DH>::KS; ++ks)
    ldsm_x4(sbase + (uint32_t)(((r0 + (lane & 15)) * Tile<DH>::LD + ks * 16 + 8 * (lane >> 4)) * 2), a[ks]);
}
// acc[nt] (16 x 8 each) = A(16 x DH) . Bs^T, Bs = [NT*8 rows][LD]  (row index of Bs = output column)
template <int DH, int NT>
__device__ __forceinline__ void gemm_rc(float (&acc)[NT][4], const uint32_t (&a)[Tile<DH>::KS][4], uint32_t sB) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int n2 = 0; n2 < NT / 2; ++n2) {
#pragma unroll
    for (int ks = 0; ks < Tile<DH>::KS; ++ks) {
      uint32_t b[4];
      ldsm_x4(sB + (uint32_t)(((n2 * 16 + (lane & 7) + 8 * (lane >> 4)) * Tile<DH>::LD + ks * 16 + 8 * ((lane >> 3) & 1)) * 2), b);
      mma16816(acc[2 * n2], a[ks], b[0], b[1]);
      mma16816(acc[2 * n2 + 1], a[ks], b[2], b[3]);
    }
  }
}
// o[nd] (16 x 8 each over the head dim) += P(16 x NT*8) . Bs, Bs = [NT*8 rows][LD]  (row index of Bs = reduction index)
template <int DH, int NT>
__device__ __forceinline__ void gemm_pv(float (&o)[Tile<DH>::ND][4], const uint32_t (&p)[NT / 2][4], uint32_t sB) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kb = 0; kb < NT / 2; ++kb) {
#pragma unroll
    for (int d2 = 0; d2 < Tile<DH>::ND / 2; ++d2) {
      uint32_t b[4];
      ldsm_x4_t(sB + (uint32_t)(((kb * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * Tile<DH>::LD + d2 * 16 + 8 * (lane >> 4)) * 2), b);
      mma16816(o[2 * d2], p[kb], b[0], b[1]);
      mma16816(o[2 * d2 + 1], p[kb], b[2], b[3]);
    }
  }
}

// one 16-column block kb of gemm_rc: acc2 (16 x 16) = A(16 x DH) . Bs[kb*16 .. kb*16+15]^T
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
// one reduction block kb of gemm_pv: o += P_kb(16 x 16) . Bs[kb*16 .. kb*16+15]
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

// stage rows [0,T) x DH of a strided global matrix into smem [TK16*16][LD]; rows >= T are zero.  Asynchronous (cp.async,
// 16 B per request, zero-fill for the padding rows): a thread's requests for ALL staged matrices are in flight together,
// so the CTA pays one global round trip, not one per matrix; the caller commits and waits (stage_wait) before its
// barrier.  The CTA has exactly TK16*32 threads, so the trip count (DH/16 chunks per thread) is a compile-time constant.
template <int DH, int TK16>
__device__ __forceinline__ void stage(uint32_t dst, const bf16* __restrict__ src, int ld_src, int T) {
  constexpr int CH = DH / 8;                 // 16-byte chunks per row
  constexpr int RPI = TK16 * 32 / CH;        // rows covered by one pass of the head's TK16 warps
  const int tid = threadIdx.x % (TK16 * 32); // several heads share a CTA (HPC): each head's warps stage that head's matrices
  const int r0 = tid / CH, c = (tid % CH) * 8;
#pragma unroll
  for (int i = 0; i < DH / 16; ++i) {
    const int r = r0 + i * RPI;
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
// 2^x on the SFU, one instruction (the arguments here are <= 0 up to rounding; -inf -> 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// p == 0.5: the 32-key mask words of one query row (common.cuh attn_drop_keep), pre-shifted by `sh` so that the bit
// of a thread's element sits at a compile-time position
template <int NW>
__device__ __forceinline__ void mask_words(uint64_t key, uint64_t row, int wpr, int sh, uint32_t (&w)[NW]) {
#pragma unroll
  for (int k = 0; k < NW; ++k) w[k] = k < wpr ? drop_bits(key, row * (uint64_t)wpr + k) >> sh : 0u;
}

// Dropout handling is a template parameter (the three variants share nothing on the per-element path and the unused ones
// would only dilute the instruction cache): MODE 0 no dropout, 1 p == 0.5 (one random bit per pair), 2 general p.
// ------------------------------------------------------------------ forward
// resident CTAs per SM the register allocation aims at: 4 for the small tiles, 2 where the score tile (T = 128) or the
// accumulators (head dim 64: the backward keeps two DH-wide tiles) need more than 80 registers - the dh-64 / T-96 backward
// spilled 2.4 KB per thread at 4 - divided by the heads a CTA carries
template <int DH, int TK16, int HPC> constexpr int min_blocks() {
  constexpr int base = (TK16 == 8 || DH >= 64) ? 2 : EGOT2_ATTN_MINB;
  return base / HPC > 0 ? base / HPC : 1;
}
// HPC heads of one clip per CTA (each head on its own TK16 warps): one staging round trip and one CTA launch serve HPC heads -
// with 8 heads of 16 dims over 48 tokens (HOI PNR) a (clip, head) CTA moved 4.5 KB and was all fixed cost.
template <int DH, int TK16, int MODE, int HPC>
__global__ void __launch_bounds__(TK16 * 32 * HPC, min_blocks<DH, TK16, HPC>())
attn_mma_fwd_kernel(int T, int H, int heads, const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                    float p_drop, uint64_t drop_key) {
  constexpr int NT = TK16 * 2, LD = Tile<DH>::LD, TP = TK16 * 16;
  extern __shared__ __align__(16) uint8_t smem[];
  const int hl = (threadIdx.x >> 5) / TK16;                       // which of this CTA's heads
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(smem) + hl * (3 * TP * LD * 2), uK = uQ + TP * LD * 2, uV = uK + TP * LD * 2;
  const int hg = heads / HPC, b = blockIdx.x / hg, h = (blockIdx.x % hg) * HPC + hl, bh = b * heads + h;
  const int warp = (threadIdx.x >> 5) % TK16, lane = threadIdx.x & 31;
  const bf16* base = qkv + (size_t)b * T * 3 * H + h * DH;
  EGOT2_PDL_ENTER();
  stage<DH, TK16>(uQ, base, 3 * H, T);
  stage<DH, TK16>(uK, base + H, 3 * H, T);
  stage<DH, TK16>(uV, base + 2 * H, 3 * H, T);
  stage_wait();
  __syncthreads();
  const int r0 = warp * 16;
  if (r0 >= T) return;
  uint32_t aq[Tile<DH>::KS][4];
  load_a<DH>(uQ, r0, aq);
  float s[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
  gemm_rc<DH, NT>(s, aq, uK);
  // softmax over keys; thread holds rows (lane/4) and (lane/4 + 8), columns nt*8 + 2*(lane%4) + {0,1}
  const float sc = rsqrtf((float)DH) * 1.4426950408889634f;        // scale * log2(e)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt * 8 + 8 > T) {                      // warp-uniform: only the last tile(s) hold padded keys
      const int c = nt * 8 + 2 * (lane & 3);
      if (c >= T) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (c + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    }
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
  const int q0 = r0 + (lane >> 2), q1 = q0 + 8;
  const float ms0 = mx0 * sc, ms1 = mx1 * sc;
  constexpr int NW = (TK16 + 1) / 2;           // 32-key mask words per query row
  uint32_t w0[NW], w1[NW];
  if (MODE == 1) {
    mask_words<NW>(drop_key ^ egot2_ep, (uint64_t)bh * T + q0, (T + 31) >> 5, 2 * (lane & 3), w0);
    mask_words<NW>(drop_key ^ egot2_ep, (uint64_t)bh * T + q1, (T + 31) >> 5, 2 * (lane & 3), w1);
  }
  // MODE 1 keeps e or zeroes it; the 1/(1-p) = 2 of the kept entries is folded into the final row scale (exact: power of 2)
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  uint32_t p[NT / 2][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float e0 = ex2(fmaf(s[nt][0], sc, -ms0)), e1 = ex2(fmaf(s[nt][1], sc, -ms0));
    float e2 = ex2(fmaf(s[nt][2], sc, -ms1)), e3 = ex2(fmaf(s[nt][3], sc, -ms1));
    sum0 += e0 + e1; sum1 += e2 + e3;
    if (MODE == 1) {
      const uint32_t m0 = w0[nt >> 2] >> ((nt & 3) * 8), m1 = w1[nt >> 2] >> ((nt & 3) * 8);
      e0 = (m0 & 1u) ? e0 : 0.f; e1 = (m0 & 2u) ? e1 : 0.f;
      e2 = (m1 & 1u) ? e2 : 0.f; e3 = (m1 & 2u) ? e3 : 0.f;
    } else if (MODE == 2) {
      const int c = nt * 8 + 2 * (lane & 3);
      attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + q0, T, c, p_drop, inv_keep, e0, e1);
      attn_drop_scale2(drop_key ^ egot2_ep, (uint64_t)bh * T + q1, T, c, p_drop, inv_keep, e2, e3);
    }
    p[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
    p[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  float o[Tile<DH>::ND][4];
#pragma unroll
  for (int i = 0; i < Tile<DH>::ND; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  gemm_pv<DH, NT>(o, p, uV);
  const float fold = MODE == 1 ? 2.f : 1.f;
  const float i0 = fold / sum0, i1 = fold / sum1;
  bf16* ob = out + (size_t)b * T * H + h * DH;
#pragma unroll
  for (int nd = 0; nd < Tile<DH>::ND; ++nd) {
    const int c = nd * 8 + 2 * (lane & 3);
    if (q0 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)q0 * H + c) = pack_bf16(o[nd][0] * i0, o[nd][1] * i0);
    if (q1 < T) *reinterpret_cast<uint32_t*>(ob + (size_t)q1 * H + c) = pack_bf16(o[nd][2] * i1, o[nd][3] * i1);
  }
  if ((lane & 3) == 0) {
    const float scn = rsqrtf((float)DH);
    if (q0 < T) lse[(size_t)bh * T + q0] = mx0 * scn + logf(sum0);
    if (q1 < T) lse[(size_t)bh * T + q1] = mx1 * scn + logf(sum1);
  }
}

// ------------------------------------------------------------------ backward
template <int DH, int TK16, int MODE, int HPC>
__global__ void __launch_bounds__(TK16 * 32 * HPC, min_blocks<DH, TK16, HPC>())
attn_mma_bwd_kernel(int T, int H, int heads, const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                    const float* __restrict__ lse, const bf16* __restrict__ dout, bf16* __restrict__ dqkv, float p_drop,
                    uint64_t drop_key) {
  constexpr int LD = Tile<DH>::LD, TP = TK16 * 16;
  constexpr int HEAD_BYTES = 4 * TP * LD * 2 + 2 * TP * 4;                   // Q, K, V, dO tiles + lse / D vectors of one head
  extern __shared__ __align__(16) uint8_t smem[];
  const int hl = (threadIdx.x >> 5) / TK16;                                  // which of this CTA's heads
  const uint32_t uQ = (uint32_t)__cvta_generic_to_shared(smem) + hl * HEAD_BYTES, uK = uQ + TP * LD * 2, uV = uK + TP * LD * 2,
                 udO = uV + TP * LD * 2;
  float* sL = reinterpret_cast<float*>(smem + (size_t)hl * HEAD_BYTES + (size_t)4 * TP * LD * 2);      // lse * log2(e) per query
  float* sD = sL + TP;                                                       // D = rowsum(dO * O) per query
  const int hg = heads / HPC, b = blockIdx.x / hg, h = (blockIdx.x % hg) * HPC + hl, bh = b * heads + h;
  const int warp = (threadIdx.x >> 5) % TK16, lane = threadIdx.x & 31;
  const int tid = threadIdx.x % (TK16 * 32);                                 // thread index inside the head's warp group
  const bf16* base = qkv + (size_t)b * T * 3 * H + h * DH;
  const bf16* ob = out + (size_t)b * T * H + h * DH;
  const bf16* dob = dout + (size_t)b * T * H + h * DH;
  constexpr float l2e = 1.4426950408889634f;
  EGOT2_PDL_ENTER();
  stage<DH, TK16>(uQ, base, 3 * H, T);
  stage<DH, TK16>(uK, base + H, 3 * H, T);
  stage<DH, TK16>(uV, base + 2 * H, 3 * H, T);
  stage<DH, TK16>(udO, dob, H, T);
  // D_i and lse_i while the copies are in flight: 4 lanes per query row, 16 B loads
  {
    constexpr int RPI = TK16 * 8;              // rows per pass (4 lanes each)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = (tid >> 2) + i * RPI;
      float d = 0.f;
      if (r < T) {
#pragma unroll
        for (int c = (tid & 3) * 8; c < DH; c += 32) {
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
      if ((tid & 3) == 0) { sD[r] = d; sL[r] = r < T ? lse[(size_t)bh * T + r] * l2e : 0.f; }
    }
  }
  stage_wait();
  __syncthreads();
  const int r0 = warp * 16;
  if (r0 >= T) return;
  const float scn = rsqrtf((float)DH);
  const float sc2 = scn * l2e;
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  constexpr int NW = (TK16 + 1) / 2;                       // 32-key (or 32-query) blocks of the clip
  const int ra = r0 + (lane >> 2), rb = ra + 8;            // this thread's two tile rows
  bf16* dq = dqkv + (size_t)b * T * 3 * H + h * DH;

  // ---------------- pass 1: rows = queries -> dQ   (streamed over 16-key blocks: nothing T-wide stays in registers)
  {
    uint32_t a1[Tile<DH>::KS][4], a2[Tile<DH>::KS][4];
    load_a<DH>(uQ, r0, a1);
    load_a<DH>(udO, r0, a2);
    const float la = sL[ra], lb = sL[rb], da = sD[ra], db = sD[rb];
    uint32_t wa[NW], wb[NW];
    if (MODE == 1) {
      mask_words<NW>(drop_key ^ egot2_ep, (uint64_t)bh * T + ra, (T + 31) >> 5, 2 * (lane & 3), wa);
      mask_words<NW>(drop_key ^ egot2_ep, (uint64_t)bh * T + rb, (T + 31) >> 5, 2 * (lane & 3), wb);
    }
    float o[Tile<DH>::ND][4];
#pragma unroll
    for (int i = 0; i < Tile<DH>::ND; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < TK16; ++kb) {
      if (kb * 16 >= T) break;
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      gemm_rc_kb<DH>(s, a1, uK, kb);
      gemm_rc_kb<DH>(dp, a2, uV, kb);
      uint32_t ds[4];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        float p0 = ex2(fmaf(s[h2][0], sc2, -la)), p1 = ex2(fmaf(s[h2][1], sc2, -la));
        float p2 = ex2(fmaf(s[h2][2], sc2, -lb)), p3 = ex2(fmaf(s[h2][3], sc2, -lb));
        if (kb * 16 + 16 > T) {                // warp-uniform: padded keys only exist in the last block
          const int c = kb * 16 + h2 * 8 + 2 * (lane & 3);
          if (c >= T) { p0 = 0.f; p2 = 0.f; }
          if (c + 1 >= T) { p1 = 0.f; p3 = 0.f; }
        }
        float g0 = dp[h2][0], g1 = dp[h2][1], g2 = dp[h2][2], g3 = dp[h2][3];
        if (MODE == 1) {
          const uint32_t m0 = wa[kb >> 1] >> ((kb & 1) * 16 + h2 * 8), m1 = wb[kb >> 1] >> ((kb & 1) * 16 + h2 * 8);
          g0 = (m0 & 1u) ? g0 + g0 : 0.f; g1 = (m0 & 2u) ? g1 + g1 : 0.f;
          g2 = (m1 & 1u) ? g2 + g2 : 0.f; g3 = (m1 & 2u) ? g3 + g3 : 0.f;
        } else if (MODE == 2) {
          const int c = kb * 16 + h2 * 8 + 2 * (lane & 3);
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
    // p == 0.5: the mask word of (query c, this warp's 32-key block) serves all 16 key rows of the warp: lane L hashes
    // the words of queries L, L+32, ... once and every element fetches its word with a shuffle
    uint32_t wq[NW];
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const int qy = lane + 32 * j;
        wq[j] = qy < T ? drop_bits(drop_key ^ egot2_ep, ((uint64_t)bh * T + qy) * (uint64_t)((T + 31) >> 5) + (uint32_t)(r0 >> 5)) : 0u;
      }
    }
    float ov[Tile<DH>::ND][4], ok[Tile<DH>::ND][4];
#pragma unroll
    for (int i = 0; i < Tile<DH>::ND; ++i) { ov[i][0] = ov[i][1] = ov[i][2] = ov[i][3] = 0.f; ok[i][0] = ok[i][1] = ok[i][2] = ok[i][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < TK16; ++kb) {
      if (kb * 16 >= T) break;
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
        if (kb * 16 + 16 > T) {                // warp-uniform: padded queries only exist in the last block
          if (c >= T) { p0 = 0.f; p2 = 0.f; }
          if (c + 1 >= T) { p1 = 0.f; p3 = 0.f; }
        }
        float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
        if (MODE == 1) {
          const uint32_t u0 = __shfl_sync(0xffffffffu, wq[kb >> 1], c & 31) >> (ra & 31);   // bit 0: key ra, bit 8: key rb
          const uint32_t u1 = __shfl_sync(0xffffffffu, wq[kb >> 1], (c + 1) & 31) >> (ra & 31);
          m0 = (u0 & 1u) ? 2.f : 0.f;
          m1 = (u1 & 1u) ? 2.f : 0.f;
          m2 = (u0 & 0x100u) ? 2.f : 0.f;
          m3 = (u1 & 0x100u) ? 2.f : 0.f;
        } else if (MODE == 2) {
          m0 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c, T, ra, p_drop, inv_keep);
          m1 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c + 1, T, ra, p_drop, inv_keep);
          m2 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c, T, rb, p_drop, inv_keep);
          m3 = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * T + c + 1, T, rb, p_drop, inv_keep);
        }
        pf[h2 * 2] = pack_bf16(p0 * m0, p1 * m1);
        pf[h2 * 2 + 1] = pack_bf16(p2 * m2, p3 * m3);
        ds[h2 * 2] = pack_bf16(p0 * fmaf(dp[h2][0], m0, -d01.x), p1 * fmaf(dp[h2][1], m1, -d01.y));
        ds[h2 * 2 + 1] = pack_bf16(p2 * fmaf(dp[h2][2], m2, -d01.x), p3 * fmaf(dp[h2][3], m3, -d01.y));
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

inline int drop_mode(float p) { return p <= 0.f ? 0 : (p == 0.5f ? 1 : 2); }

// MEASURED (B200, B 256): carrying several heads per CTA is SLOWER - T 48 / 8 heads of 16: forward 20.0 vs 13.0 us, backward
// 35.5 vs 24.8 with 4 heads per CTA; T 90 / 4 heads of 32: 13.4 vs 11.9 and 32.6 vs 28.6 with 2 - the many small CTAs hide
// the staging round trip better than fewer large ones do.  So one head per CTA it stays; the HPC parameter is kept (set
// hpc_max to try again) but nothing above 1 is instantiated.
template <int DH, int TK16> constexpr int hpc_max() { return 1; }

template <int DH, int TK16, int HPC>
int launch_fwd_h(int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p, uint64_t key, cudaStream_t st) {
  constexpr size_t smem = (size_t)HPC * 3 * TK16 * 16 * Tile<DH>::LD * 2;
  typedef void (*Kern)(int, int, int, const bf16*, bf16*, float*, float, uint64_t);
  static const Kern kerns[3] = {attn_mma_fwd_kernel<DH, TK16, 0, HPC>, attn_mma_fwd_kernel<DH, TK16, 1, HPC>, attn_mma_fwd_kernel<DH, TK16, 2, HPC>};
  static bool set = false;
  if (!set) {
    for (int i = 0; i < 3; ++i) EGOT2_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  ProfScope prof(st, "attn_mma_fwd<dh%d,tk%d,hpc%d> B%d T%d H%d", DH, TK16 * 16, HPC, B, T, H);
  launch(kerns[drop_mode(p)], dim3(B * heads / HPC), dim3(TK16 * 32 * HPC), smem, st, T, H, heads, (const bf16*)qkv, (bf16*)out, lse, p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
inline bool hpc_off() { const char* e = getenv("EGOT2_ATTN_HPC"); return e && e[0] == '1' && e[1] == 0; }
template <int DH, int TK16>
int launch_fwd(int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p, uint64_t key, cudaStream_t st) {
  constexpr int M = hpc_max<DH, TK16>();
  if (!hpc_off()) {
    if constexpr (M >= 4) if (heads % 4 == 0) return launch_fwd_h<DH, TK16, 4>(B, T, H, heads, qkv, out, lse, p, key, st);
    if constexpr (M >= 2) if (heads % 2 == 0) return launch_fwd_h<DH, TK16, 2>(B, T, H, heads, qkv, out, lse, p, key, st);
  }
  return launch_fwd_h<DH, TK16, 1>(B, T, H, heads, qkv, out, lse, p, key, st);
}
template <int DH, int TK16, int HPC>
int launch_bwd_h(int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
                 void* dqkv, float p, uint64_t key, cudaStream_t st) {
  constexpr size_t smem = (size_t)HPC * ((size_t)4 * TK16 * 16 * Tile<DH>::LD * 2 + 2 * TK16 * 16 * 4);
  typedef void (*Kern)(int, int, int, const bf16*, const bf16*, const float*, const bf16*, bf16*, float, uint64_t);
  static const Kern kerns[3] = {attn_mma_bwd_kernel<DH, TK16, 0, HPC>, attn_mma_bwd_kernel<DH, TK16, 1, HPC>, attn_mma_bwd_kernel<DH, TK16, 2, HPC>};
  static bool set = false;
  if (!set) {
    for (int i = 0; i < 3; ++i) EGOT2_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  ProfScope prof(st, "attn_mma_bwd<dh%d,tk%d,hpc%d> B%d T%d H%d", DH, TK16 * 16, HPC, B, T, H);
  launch(kerns[drop_mode(p)], dim3(B * heads / HPC), dim3(TK16 * 32 * HPC), smem, st, T, H, heads, (const bf16*)qkv, (const bf16*)out, lse,
         (const bf16*)dout, (bf16*)dqkv, p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
template <int DH, int TK16>
int launch_bwd(int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
               void* dqkv, float p, uint64_t key, cudaStream_t st) {
  constexpr int M = hpc_max<DH, TK16>();
  if (!hpc_off()) {
    if constexpr (M >= 4) if (heads % 4 == 0) return launch_bwd_h<DH, TK16, 4>(B, T, H, heads, qkv, out, lse, dout, dqkv, p, key, st);
    if constexpr (M >= 2) if (heads % 2 == 0) return launch_bwd_h<DH, TK16, 2>(B, T, H, heads, qkv, out, lse, dout, dqkv, p, key, st);
  }
  return launch_bwd_h<DH, TK16, 1>(B, T, H, heads, qkv, out, lse, dout, dqkv, p, key, st);
}

inline int pick_tk16(int T) { return T <= 16 ? 1 : (T <= 32 ? 2 : (T <= 64 ? 4 : (T <= 96 ? 6 : 8))); }

}  // namespace

bool attention_mma_supported(int dtype, int T, int H, int heads) {
  if (dtype != EGOT2_BF16 || heads <= 0 || H % heads) return false;
  const int dh = H / heads;
  if (dh == 128) return T >= 1 && T <= 32 && (H % 8 == 0);      // LTA at its shipped width (H 1024 / 8 heads, 8 tokens): the
                                                                 // output tile alone is 64 registers, so short clips only
  return T >= 1 && T <= 128 && (dh == 16 || dh == 32 || dh == 64) && (H % 8 == 0);
}

#define EGOT2_ATT_SWITCH(FN, ...)                                                     \
  switch (dh * 16 + tk) {                                                             \
    case 16 * 16 + 1: return FN<16, 1>(__VA_ARGS__);  case 16 * 16 + 2: return FN<16, 2>(__VA_ARGS__); \
    case 16 * 16 + 4: return FN<16, 4>(__VA_ARGS__);  case 16 * 16 + 6: return FN<16, 6>(__VA_ARGS__); \
    case 16 * 16 + 8: return FN<16, 8>(__VA_ARGS__);                                  \
    case 32 * 16 + 1: return FN<32, 1>(__VA_ARGS__);  case 32 * 16 + 2: return FN<32, 2>(__VA_ARGS__); \
    case 32 * 16 + 4: return FN<32, 4>(__VA_ARGS__);  case 32 * 16 + 6: return FN<32, 6>(__VA_ARGS__); \
    case 32 * 16 + 8: return FN<32, 8>(__VA_ARGS__);                                  \
    case 64 * 16 + 1: return FN<64, 1>(__VA_ARGS__);  case 64 * 16 + 2: return FN<64, 2>(__VA_ARGS__); \
    case 64 * 16 + 4: return FN<64, 4>(__VA_ARGS__);  case 64 * 16 + 6: return FN<64, 6>(__VA_ARGS__); \
    case 64 * 16 + 8: return FN<64, 8>(__VA_ARGS__);                                  \
    case 128 * 16 + 1: return FN<128, 1>(__VA_ARGS__); case 128 * 16 + 2: return FN<128, 2>(__VA_ARGS__); \
  }

int attention_mma_fwd(int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p_drop,
                      uint64_t drop_key, cudaStream_t st) {
  if (B * T == 0) return 0;
  const int dh = H / heads, tk = pick_tk16(T);
  EGOT2_ATT_SWITCH(launch_fwd, B, T, H, heads, qkv, out, lse, p_drop, drop_key, st)
  EGOT2_CHECK(false, "attention_mma_fwd: unsupported dh=%d T=%d", dh, T);
}
int attention_mma_bwd(int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
                      void* dqkv, float p_drop, uint64_t drop_key, cudaStream_t st) {
  if (B * T == 0) return 0;
  const int dh = H / heads, tk = pick_tk16(T);
  EGOT2_ATT_SWITCH(launch_bwd, B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st)
  EGOT2_CHECK(false, "attention_mma_bwd: unsupported dh=%d T=%d", dh, T);
}

}  // namespace egot2
