// attention_simt.cu — shape-general self-attention over one clip's tokens (fp32 arithmetic).
//
// Scores, softmax statistics and accumulation are fp32 for both activation dtypes.  K/V (forward,
// dQ pass) or Q/dO (dK/dV pass) are streamed through shared memory in 64-row blocks with a +1 pad
// (conflict-free lane-per-row dot products); softmax is online (running max / sum), so any T
// fits; nothing of size T x T is ever stored: backward recomputes the probabilities from the
// saved log-sum-exp.  Dropout on the probabilities is regenerated from (key, element index).
// Layout: qkv (B, T, 3H) with q|k|v packed along the last dim, head h = columns h*dh..(h+1)*dh.
#include <math.h>

#define EGOT2_FILE_ID 9
#include "ops.h"

namespace egot2 {

namespace {

constexpr int NW = 8;         // warps per CTA
// per head-dim tiling, chosen so that static shared memory stays below 48 KB:
//   KB  rows of the streamed operand per shared-memory block
//   RPW rows (queries, or keys in the dK/dV pass) owned by one warp;  RB = NW*RPW rows per CTA
template <int DH> struct Cfg {
  static constexpr int KB = DH <= 32 ? 64 : 32;
  static constexpr int RPW = DH <= 64 ? 4 : 1;
  static constexpr int RB = NW * RPW;
};

template <int DH> struct Cols { static constexpr int N = (DH + 31) / 32; };

// stage `rows` rows x DH columns starting at row r0 of a (B,T,3H)-packed tensor into smem (fp32)
template <typename T, int DH, int KB>
__device__ __forceinline__ void stage(float (*dst)[DH + 1], const T* __restrict__ base, int ld, int r0, int rows,
                                      int Ttot) {
  for (int e = threadIdx.x; e < KB * DH; e += blockDim.x) {
    const int r = e / DH, c = e % DH;
    dst[r][c] = (r < rows && r0 + r < Ttot) ? to_f32(base[(size_t)(r0 + r) * ld + c]) : 0.f;
  }
}

// ------------------------------------------------------------------ forward
template <typename T, int DH>
__global__ void __launch_bounds__(NW * 32) attn_fwd_kernel(int Tn, int H, int heads, const T* __restrict__ qkv,
                                                           T* __restrict__ out, float* __restrict__ lse, float p_drop,
                                                           uint64_t drop_key) {
  EGOT2_PDL_ENTER();
  constexpr int NC = Cols<DH>::N;
  constexpr int KB = Cfg<DH>::KB, RPW = Cfg<DH>::RPW, RB = Cfg<DH>::RB;
  __shared__ float Ks[KB][DH + 1];
  __shared__ float Vs[KB][DH + 1];
  __shared__ float Qs[RB][DH + 1];
  __shared__ float Ps[NW][KB];
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.y * RB;
  const int ld = 3 * H;
  const T* qbase = qkv + (size_t)b * Tn * ld + h * DH;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const float scale = rsqrtf((float)DH);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;

  for (int e = threadIdx.x; e < RB * DH; e += blockDim.x) {
    const int r = e / DH, c = e % DH;
    Qs[r][c] = (q0 + r < Tn) ? to_f32(qbase[(size_t)(q0 + r) * ld + c]) * scale : 0.f;
  }
  float m[RPW], l[RPW], acc[RPW][NC];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[r][i] = 0.f;
  }
  for (int k0 = 0; k0 < Tn; k0 += KB) {
    __syncthreads();
    stage<T, DH, KB>(Ks, kbase, ld, k0, KB, Tn);
    stage<T, DH, KB>(Vs, vbase, ld, k0, KB, Tn);
    __syncthreads();
    const int nk = min(KB, Tn - k0);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int qr = warp * RPW + r, qi = q0 + qr;
      if (qi >= Tn) continue;                               // warp-uniform
      float s[KB / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < KB / 32; ++jj) {
        const int j = lane + 32 * jj;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) d = fmaf(Qs[qr][c], Ks[j][c], d);
        s[jj] = j < nk ? d : -INFINITY;
        mx = fmaxf(mx, s[jj]);
      }
      mx = warp_max(mx);
      const float m_new = fmaxf(m[r], mx);
      const float corr = expf(m[r] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < KB / 32; ++jj) {
        const int j = lane + 32 * jj;
        float p = j < nk ? expf(s[jj] - m_new) : 0.f;
        psum += p;
        if (p_drop > 0.f)
          p *= attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * Tn + qi, Tn, k0 + j, p_drop, inv_keep);
        Ps[warp][j] = p;
      }
      l[r] = l[r] * corr + warp_sum(psum);
      m[r] = m_new;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        float a = acc[r][i] * corr;
        if (c < DH)
          for (int j = 0; j < nk; ++j) a = fmaf(Ps[warp][j], Vs[j][c], a);
        acc[r][i] = a;
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int qi = q0 + warp * RPW + r;
    if (qi >= Tn) continue;
    const float inv_l = 1.f / l[r];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      if (c < DH) out[((size_t)b * Tn + qi) * H + h * DH + c] = from_f32<T>(acc[r][i] * inv_l);
    }
    if (lane == 0) lse[(size_t)bh * Tn + qi] = m[r] + logf(l[r]);
  }
}

// ------------------------------------------------------------------ backward pass 1: D_i and dQ_i (one warp per query row)
template <typename T, int DH>
__global__ void __launch_bounds__(NW * 32) attn_bwd_dq_kernel(int Tn, int H, int heads, const T* __restrict__ qkv,
                                                              const T* __restrict__ out, const float* __restrict__ lse,
                                                              const T* __restrict__ dout, T* __restrict__ dqkv,
                                                              float* __restrict__ Dvec, float p_drop, uint64_t drop_key) {
  EGOT2_PDL_ENTER();
  constexpr int NC = Cols<DH>::N;
  constexpr int KB = Cfg<DH>::KB, RPW = Cfg<DH>::RPW, RB = Cfg<DH>::RB;
  __shared__ float Ks[KB][DH + 1];
  __shared__ float Vs[KB][DH + 1];
  __shared__ float Qs[RB][DH + 1];
  __shared__ float dOs[RB][DH + 1];
  __shared__ float Ps[NW][KB];
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = blockIdx.y * RB;
  const int ld = 3 * H;
  const T* qbase = qkv + (size_t)b * Tn * ld + h * DH;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const T* obase = out + (size_t)b * Tn * H + h * DH;
  const T* dobase = dout + (size_t)b * Tn * H + h * DH;
  const float scale = rsqrtf((float)DH);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;

  for (int e = threadIdx.x; e < RB * DH; e += blockDim.x) {
    const int r = e / DH, c = e % DH;
    const bool ok = q0 + r < Tn;
    Qs[r][c] = ok ? to_f32(qbase[(size_t)(q0 + r) * ld + c]) * scale : 0.f;
    dOs[r][c] = ok ? to_f32(dobase[(size_t)(q0 + r) * H + c]) : 0.f;
  }
  __syncthreads();
  float Di[RPW], L[RPW], acc[RPW][NC];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int qr = warp * RPW + r, qi = q0 + qr;
    float d = 0.f;
    if (qi < Tn)
      for (int c = lane; c < DH; c += 32) d += dOs[qr][c] * to_f32(obase[(size_t)qi * H + c]);
    Di[r] = warp_sum(d);
    L[r] = qi < Tn ? lse[(size_t)bh * Tn + qi] : 0.f;
    if (qi < Tn && lane == 0) Dvec[(size_t)bh * Tn + qi] = Di[r];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[r][i] = 0.f;
  }
  for (int k0 = 0; k0 < Tn; k0 += KB) {
    __syncthreads();
    stage<T, DH, KB>(Ks, kbase, ld, k0, KB, Tn);
    stage<T, DH, KB>(Vs, vbase, ld, k0, KB, Tn);
    __syncthreads();
    const int nk = min(KB, Tn - k0);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int qr = warp * RPW + r, qi = q0 + qr;
      if (qi >= Tn) continue;
#pragma unroll
      for (int jj = 0; jj < KB / 32; ++jj) {
        const int j = lane + 32 * jj;
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          s = fmaf(Qs[qr][c], Ks[j][c], s);
          dp = fmaf(dOs[qr][c], Vs[j][c], dp);
        }
        float ds = 0.f;
        if (j < nk) {
          const float p = expf(s - L[r]);
          if (p_drop > 0.f)
            dp *= attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * Tn + qi, Tn, k0 + j, p_drop, inv_keep);
          ds = p * (dp - Di[r]);
        }
        Ps[warp][j] = ds;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        float a = acc[r][i];
        if (c < DH)
          for (int j = 0; j < nk; ++j) a = fmaf(Ps[warp][j], Ks[j][c], a);
        acc[r][i] = a;
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int qi = q0 + warp * RPW + r;
    if (qi >= Tn) continue;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      if (c < DH) dqkv[((size_t)b * Tn + qi) * ld + h * DH + c] = from_f32<T>(acc[r][i] * scale);
    }
  }
}

// ------------------------------------------------------------------ backward pass 2: dK_j, dV_j (one warp per key row)
template <typename T, int DH>
__global__ void __launch_bounds__(NW * 32) attn_bwd_dkv_kernel(int Tn, int H, int heads, const T* __restrict__ qkv,
                                                               const float* __restrict__ lse, const T* __restrict__ dout,
                                                               T* __restrict__ dqkv, const float* __restrict__ Dvec,
                                                               float p_drop, uint64_t drop_key) {
  EGOT2_PDL_ENTER();
  constexpr int NC = Cols<DH>::N;
  constexpr int KB = Cfg<DH>::KB, RPW = Cfg<DH>::RPW, RB = Cfg<DH>::RB;
  __shared__ float Qs[KB][DH + 1];     // streamed: scaled queries
  __shared__ float dOs[KB][DH + 1];    // streamed: dO rows
  __shared__ float Ks[RB][DH + 1];     // resident: this CTA's keys
  __shared__ float Vs[RB][DH + 1];
  __shared__ float Ls[KB], Ds[KB];
  __shared__ float Ps[NW][KB], dSs[NW][KB];
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.y * RB;
  const int ld = 3 * H;
  const T* qbase = qkv + (size_t)b * Tn * ld + h * DH;
  const T* kbase = qbase + H;
  const T* vbase = qbase + 2 * H;
  const T* dobase = dout + (size_t)b * Tn * H + h * DH;
  const float scale = rsqrtf((float)DH);
  const float inv_keep = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;

  for (int e = threadIdx.x; e < RB * DH; e += blockDim.x) {
    const int r = e / DH, c = e % DH;
    const bool ok = j0 + r < Tn;
    Ks[r][c] = ok ? to_f32(kbase[(size_t)(j0 + r) * ld + c]) : 0.f;
    Vs[r][c] = ok ? to_f32(vbase[(size_t)(j0 + r) * ld + c]) : 0.f;
  }
  float dk[RPW][NC], dv[RPW][NC];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int i = 0; i < NC; ++i) { dk[r][i] = 0.f; dv[r][i] = 0.f; }

  for (int i0 = 0; i0 < Tn; i0 += KB) {
    __syncthreads();
    for (int e = threadIdx.x; e < KB * DH; e += blockDim.x) {
      const int r = e / DH, c = e % DH;
      const bool ok = i0 + r < Tn;
      Qs[r][c] = ok ? to_f32(qbase[(size_t)(i0 + r) * ld + c]) * scale : 0.f;
      dOs[r][c] = ok ? to_f32(dobase[(size_t)(i0 + r) * H + c]) : 0.f;
    }
    for (int e = threadIdx.x; e < KB; e += blockDim.x) {
      const bool ok = i0 + e < Tn;
      Ls[e] = ok ? lse[(size_t)bh * Tn + i0 + e] : 0.f;
      Ds[e] = ok ? Dvec[(size_t)bh * Tn + i0 + e] : 0.f;
    }
    __syncthreads();
    const int nq = min(KB, Tn - i0);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int kr = warp * RPW + r, kj = j0 + kr;
      if (kj >= Tn) continue;
#pragma unroll
      for (int ii = 0; ii < KB / 32; ++ii) {
        const int i = lane + 32 * ii;
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) {
          s = fmaf(Qs[i][c], Ks[kr][c], s);
          dp = fmaf(dOs[i][c], Vs[kr][c], dp);
        }
        float pd = 0.f, ds = 0.f;
        if (i < nq) {
          const float p = expf(s - Ls[i]);
          float mk = 1.f;
          if (p_drop > 0.f) mk = attn_drop_scale(drop_key ^ egot2_ep, (uint64_t)bh * Tn + (i0 + i), Tn, kj, p_drop, inv_keep);
          pd = p * mk;
          ds = p * (dp * mk - Ds[i]);
        }
        Ps[warp][i] = pd;
        dSs[warp][i] = ds;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        if (c < DH) {
          float a = dv[r][i], bq = dk[r][i];
          for (int q = 0; q < nq; ++q) {
            a = fmaf(Ps[warp][q], dOs[q][c], a);
            bq = fmaf(dSs[warp][q], Qs[q][c], bq);     // Qs already carries the 1/sqrt(dh) factor
          }
          dv[r][i] = a; dk[r][i] = bq;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int kj = j0 + warp * RPW + r;
    if (kj >= Tn) continue;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      if (c < DH) {
        dqkv[((size_t)b * Tn + kj) * ld + H + h * DH + c] = from_f32<T>(dk[r][i]);
        dqkv[((size_t)b * Tn + kj) * ld + 2 * H + h * DH + c] = from_f32<T>(dv[r][i]);
      }
    }
  }
}

template <typename T, int DH>
int fwd_launch(int B, int Tn, int H, int heads, const void* qkv, void* out, float* lse, float p, uint64_t key,
               cudaStream_t st) {
  constexpr int RB = Cfg<DH>::RB;
  dim3 grid(B * heads, (Tn + RB - 1) / RB);
  ProfScope prof(st, "attn_simt_fwd<dh%d> B%d T%d H%d", DH, B, Tn, H);
  launch(attn_fwd_kernel<T, DH>, dim3(grid), dim3(NW * 32), 0, st, Tn, H, heads, (const T*)qkv, (T*)out, lse, p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
template <typename T, int DH>
int bwd_launch(int B, int Tn, int H, int heads, const void* qkv, const void* out, const float* lse, const void* dout,
               void* dqkv, float* Dvec, float p, uint64_t key, cudaStream_t st) {
  constexpr int RB = Cfg<DH>::RB;
  dim3 grid(B * heads, (Tn + RB - 1) / RB);
  ProfScope prof(st, "attn_simt_bwd<dh%d> B%d T%d H%d (2 kernels)", DH, B, Tn, H);
  launch(attn_bwd_dq_kernel<T, DH>, dim3(grid), dim3(NW * 32), 0, st, Tn, H, heads, (const T*)qkv, (const T*)out, lse, (const T*)dout,
                                                      (T*)dqkv, Dvec, p, key);
  EGOT2_LAUNCH_CHECK();
  launch(attn_bwd_dkv_kernel<T, DH>, dim3(grid), dim3(NW * 32), 0, st, Tn, H, heads, (const T*)qkv, lse, (const T*)dout, (T*)dqkv, Dvec,
                                                       p, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

#define EGOT2_DH_SWITCH(dh, T, FN, ...)                                        \
  switch (dh) {                                                                \
    case 16: return FN<T, 16>(__VA_ARGS__);                                    \
    case 32: return FN<T, 32>(__VA_ARGS__);                                    \
    case 64: return FN<T, 64>(__VA_ARGS__);                                    \
    case 128: return FN<T, 128>(__VA_ARGS__);                                  \
    default: EGOT2_CHECK(false, "attention: head dim %d not in {16,32,64,128}", dh); \
  }

int attention_simt_fwd(int dtype, int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p_drop,
                       uint64_t drop_key, cudaStream_t st) {
  EGOT2_CHECK(heads > 0 && H % heads == 0, "attention: H=%d not divisible by heads=%d", H, heads);
  if (B * T == 0) return 0;
  const int dh = H / heads;
  if (dtype == EGOT2_F32) { EGOT2_DH_SWITCH(dh, float, fwd_launch, B, T, H, heads, qkv, out, lse, p_drop, drop_key, st); }
  else { EGOT2_DH_SWITCH(dh, bf16, fwd_launch, B, T, H, heads, qkv, out, lse, p_drop, drop_key, st); }
}

size_t attention_bwd_workspace(int dtype, int B, int T, int H, int heads) {
  (void)dtype; (void)H;
  return align_up((size_t)B * heads * T * sizeof(float));
}

int attention_simt_bwd(int dtype, int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse,
                       const void* dout, void* dqkv, float p_drop, uint64_t drop_key, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  EGOT2_CHECK(heads > 0 && H % heads == 0, "attention: H=%d not divisible by heads=%d", H, heads);
  if (B * T == 0) return 0;
  EGOT2_CHECK(ws && ws_bytes >= attention_bwd_workspace(dtype, B, T, H, heads), "attention_bwd: workspace too small");
  const int dh = H / heads;
  float* Dvec = (float*)ws;
  if (dtype == EGOT2_F32) { EGOT2_DH_SWITCH(dh, float, bwd_launch, B, T, H, heads, qkv, out, lse, dout, dqkv, Dvec, p_drop, drop_key, st); }
  else { EGOT2_DH_SWITCH(dh, bf16, bwd_launch, B, T, H, heads, qkv, out, lse, dout, dqkv, Dvec, p_drop, drop_key, st); }
}

}  // namespace egot2
