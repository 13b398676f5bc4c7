// gemm_simt.cu — generic CUDA-core GEMM (fp32 accumulate).
//
// This is the arithmetic of the EGOT2_F32 parity mode (fp32 FMA, so logits stay within 1e-3 of
// the reference and argmax stays bit-exact) and the shape-general fallback *inside the CUDA
// library* for bf16 GEMMs whose shape the tcgen05 kernel (gemm_sm100.cu) does not take.
// Register-tiled: BMxBN block tile, BK=16 k-slab staged k-major in (double-buffered) shared
// memory, TMxTN outputs per thread, 256 threads.  Handles all four operand orientations, storage-row
// remapping (segment scatter inside (B,T,H) tensors), the fused epilogue of ops.h and split-K
// with fp32 atomics for the weight-gradient GEMMs (K = all tokens of the batch).
#include <stdlib.h>
#define EGOT2_FILE_ID 10
#include "ops.h"

namespace egot2 {

namespace {

template <typename T> __device__ __forceinline__ float ld(const T* p) { return to_f32(__ldg(p)); }

__device__ __forceinline__ long long remap(int r, int rpg, int gstride) {
  return rpg > 0 ? (long long)(r / rpg) * gstride + (r % rpg) : (long long)r;
}

// vector / scalar global loads of E consecutive elements (E = 4 or 8)
template <typename TI, int E>
__device__ __forceinline__ void load_run(const TI* __restrict__ src, bool vec_ok, int valid, float (&r)[E]) {
  if constexpr (sizeof(TI) == 4) {
    if (vec_ok && valid >= E) {
#pragma unroll
      for (int v = 0; v < E / 4; ++v) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src) + v);
        r[4 * v] = t.x; r[4 * v + 1] = t.y; r[4 * v + 2] = t.z; r[4 * v + 3] = t.w;
      }
      return;
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = e < valid ? ld(src + e) : 0.f;
}

// Register-tiled CUDA-core GEMM.  Per 16-deep k-slab a thread reads its A and B fragments with 128-bit shared-memory loads
// (its TM rows / TN columns are split in groups of 4 that lie BM/2 / BN/2 apart, so the 8 threads of a load phase read 128
// contiguous bytes: conflict-free), the next slab's global loads (128-bit where alignment allows) are in flight in registers
// while the current one is multiplied, and the two shared-memory buffers alternate with ONE barrier per slab.  The fp32 FMA
// order along k is ascending per output element, exactly as in the plain version this replaces (17 -> ~45 TFLOP/s on B200).
template <typename TI, typename TO, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__(256, 2) gemm_simt_kernel(const GemmArgs a) {
  EGOT2_PDL_ENTER();
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT && TM % 4 == 0 && TN % 4 == 0, "thread tiling");
  constexpr int GM = TM / 4, GN = TN / 4;            // groups of 4 rows / columns per thread
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // split-K range of this CTA
  const int kchunk = ((a.K + a.split_k - 1) / a.split_k + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(a.K, kbeg + kchunk);

  const TI* __restrict__ A = (const TI*)a.A;
  const TI* __restrict__ B = (const TI*)a.B;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ty = tid / (BN / TN), tx = tid % (BN / TN);

  constexpr int EA = BM * BK / NT;   // elements of the A tile per thread (4 or 8), consecutive along the operand's contiguous axis
  constexpr int EB = BN * BK / NT;
  const bool a_vec = sizeof(TI) == 4 && (a.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool b_vec = sizeof(TI) == 4 && (a.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  // this thread's slice of the A / B tile: (row r, first k kk) when k is contiguous, (k row kk, first column r) otherwise
  const int a_r = !a.trans_a ? tid / (BK / EA) : (tid % (BM / EA)) * EA, a_k = !a.trans_a ? (tid % (BK / EA)) * EA : tid / (BM / EA);
  const int b_r = a.trans_b ? tid / (BK / EB) : (tid % (BN / EB)) * EB, b_k = a.trans_b ? (tid % (BK / EB)) * EB : tid / (BN / EB);
  // row pointers that do not depend on k (k-contiguous operands)
  const TI* a_row = (!a.trans_a && m0 + a_r < a.M) ? A + remap(m0 + a_r, a.a_rpg, a.a_gstride) * a.lda : nullptr;
  const TI* b_row = (a.trans_b && n0 + b_r < a.N) ? B + remap(n0 + b_r, a.b_rpg, a.b_gstride) * a.ldb : nullptr;

  float ra[EA], rb[EB];
  auto fetch = [&](int k0) {
    if (!a.trans_a) {                       // stored (M,K): k contiguous
      const int k = k0 + a_k;
      load_run<TI, EA>(a_row ? a_row + k : A, a_vec && a_row && ((k & 3) == 0), a_row ? kend - k : 0, ra);
    } else {                                // stored (K,M): m contiguous
      const int k = k0 + a_k;
      const bool ok = k < kend;
      const TI* src = A + (ok ? remap(k, a.a_rpg, a.a_gstride) * a.lda : 0) + m0 + a_r;
      load_run<TI, EA>(src, a_vec && (((m0 + a_r) & 3) == 0), ok ? a.M - (m0 + a_r) : 0, ra);
    }
    if (a.trans_b) {                        // stored (N,K): k contiguous
      const int k = k0 + b_k;
      load_run<TI, EB>(b_row ? b_row + k : B, b_vec && b_row && ((k & 3) == 0), b_row ? kend - k : 0, rb);
    } else {                                // stored (K,N): n contiguous
      const int k = k0 + b_k;
      const bool ok = k < kend;
      const TI* src = B + (ok ? remap(k, a.b_rpg, a.b_gstride) * a.ldb : 0) + n0 + b_r;
      load_run<TI, EB>(src, b_vec && (((n0 + b_r) & 3) == 0), ok ? a.N - (n0 + b_r) : 0, rb);
    }
  };
  auto stash = [&](int buf) {
    if (!a.trans_a) {
#pragma unroll
      for (int e = 0; e < EA; ++e) As[buf][a_k + e][a_r] = ra[e];
    } else {
#pragma unroll
      for (int v = 0; v < EA / 4; ++v)
        *reinterpret_cast<float4*>(&As[buf][a_k][a_r + 4 * v]) = make_float4(ra[4 * v], ra[4 * v + 1], ra[4 * v + 2], ra[4 * v + 3]);
    }
    if (a.trans_b) {
#pragma unroll
      for (int e = 0; e < EB; ++e) Bs[buf][b_k + e][b_r] = rb[e];
    } else {
#pragma unroll
      for (int v = 0; v < EB / 4; ++v)
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_r + 4 * v]) = make_float4(rb[4 * v], rb[4 * v + 1], rb[4 * v + 2], rb[4 * v + 3]);
    }
  };

  int cur = 0;
  if (kbeg < kend) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) fetch(k0 + BK);               // in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int g = 0; g < GM; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&As[cur][kk][g * (BM / GM) + ty * 4]);
        av[4 * g] = t.x; av[4 * g + 1] = t.y; av[4 * g + 2] = t.z; av[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int g = 0; g < GN; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[cur][kk][g * (BN / GN) + tx * 4]);
        bv[4 * g] = t.x; bv[4 * g + 1] = t.y; bv[4 * g + 2] = t.z; bv[4 * g + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(cur ^ 1);               // the other buffer: nobody reads it during this iteration
    __syncthreads();
    cur ^= 1;
  }

  // ---- epilogue
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const bool first_split = blockIdx.z == 0;
  TO* __restrict__ C = (TO*)a.C;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i >> 2) * (BM / GM) + ty * 4 + (i & 3);       // the thread's rows / columns come in groups of 4 (see above)
    if (m >= a.M) continue;
    const long long crow = remap(m, a.c_rpg, a.c_gstride);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j >> 2) * (BN / GN) + tx * 4 + (j & 3);
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias && first_split) v += __ldg(a.bias + n);
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.mask) v = to_f32(((const TO*)a.mask)[(long long)m * a.ldm + n]) > 0.f ? v * a.mask_scale : 0.f;
      if (a.p_drop > 0.f) v *= drop_scale(a.drop_key ^ egot2_ep, (uint64_t)m * a.N + n, a.p_drop, inv_keep, a.drop_bit_mode != 0);
      if (a.residual && first_split) v += to_f32(((const TO*)a.residual)[(long long)m * a.ldr + n]);
      TO* dst = C + crow * a.ldc + n;
      if (a.accumulate) {
        if constexpr (sizeof(TO) == 4) {
          if (a.split_k > 1) atomicAdd((float*)dst, v);
          else *dst = from_f32<TO>(to_f32(*dst) + v);
        }
      } else {
        *dst = from_f32<TO>(v);
      }
    }
  }
}

// ---------------------------------------------------------------- skinny problems (the 7-word vocabulary head of HHI EgoT2-g)
// The register-tiled kernel above pays a 16-deep k-slab round trip per 16 k and a 64- or 128-wide tile whatever N is: 27 us
// for (1200 x 256) . (256 x 7), 45-56 us for (1200 x 7) . (7 x 256).  Two direct kernels instead (bf16 operands only: the
// fp32 parity mode keeps the summation order of the tiled kernel).
// N <= 8, B stored (N, K): one warp per output row, lanes stride over k, warp-shuffle reduction.
template <typename TO>
__global__ void __launch_bounds__(256) gemm_skinny_n_kernel(const GemmArgs a) {
  EGOT2_PDL_ENTER();
  const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  const bf16* __restrict__ A = (const bf16*)a.A + (size_t)m * a.lda;
  const bf16* __restrict__ B = (const bf16*)a.B;
  float acc[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = 0.f;
  for (int k = lane; k < a.K; k += 32) {
    const float x = to_f32(A[k]);
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < a.N) acc[n] = fmaf(x, to_f32(B[(size_t)n * a.ldb + k]), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n] = warp_sum(acc[n]);
  if (lane < a.N) {
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) v = lane == n ? acc[n] : v;
    if (a.bias) v += a.bias[lane];
    ((TO*)a.C)[(size_t)m * a.ldc + lane] = from_f32<TO>(v);
  }
}
// K <= 8, B stored (K, N): one thread per output element, k ascending (the tiled kernel's order).
template <typename TO>
__global__ void __launch_bounds__(256) gemm_skinny_k_kernel(const GemmArgs a) {
  EGOT2_PDL_ENTER();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)a.M * a.N) return;
  const int m = (int)(idx / a.N), n = (int)(idx % a.N);
  const bf16* __restrict__ A = (const bf16*)a.A + (size_t)m * a.lda;
  const bf16* __restrict__ B = (const bf16*)a.B + n;
  float v = 0.f;
  for (int k = 0; k < a.K; ++k) v = fmaf(to_f32(A[k]), to_f32(B[(size_t)k * a.ldb]), v);
  if (a.bias) v += a.bias[n];
  ((TO*)a.C)[(size_t)m * a.ldc + n] = from_f32<TO>(v);
}

template <typename TO>
int launch_skinny(const GemmArgs& a, cudaStream_t st) {
  if (a.N <= 8 && a.trans_b) {
    ProfScope prof(st, "gemm_skinny_n M%d N%d K%d", a.M, a.N, a.K);
    launch(gemm_skinny_n_kernel<TO>, dim3((a.M + 7) / 8), dim3(256), 0, st, a);
  } else {
    ProfScope prof(st, "gemm_skinny_k M%d N%d K%d", a.M, a.N, a.K);
    launch(gemm_skinny_k_kernel<TO>, dim3((unsigned)(((long long)a.M * a.N + 255) / 256)), dim3(256), 0, st, a);
  }
  EGOT2_LAUNCH_CHECK();
  return 0;
}

template <typename TI, typename TO>
int launch(const GemmArgs& a, cudaStream_t st) {
  const bool small = (a.M <= 64 || a.N <= 64);
  ProfScope prof(st, "gemm_simt<%s> M%d N%d K%d ta%d tb%d sk%d", sizeof(TI) == 4 ? "f32" : "bf16", a.M, a.N, a.K, a.trans_a,
                 a.trans_b, a.split_k);
  if (small) {
    dim3 grid((a.N + 63) / 64, (a.M + 63) / 64, a.split_k);
    launch(gemm_simt_kernel<TI, TO, 64, 64, 16, 4, 4>, dim3(grid), dim3(256), 0, st, a);
  } else {
    dim3 grid((a.N + 127) / 128, (a.M + 127) / 128, a.split_k);
    launch(gemm_simt_kernel<TI, TO, 128, 128, 16, 8, 8>, dim3(grid), dim3(256), 0, st, a);
  }
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gemm_simt(const GemmArgs& a, cudaStream_t st) {
  EGOT2_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem %dx%dx%d", a.M, a.N, a.K);
  EGOT2_CHECK(!(a.accumulate && a.out_dtype != EGOT2_F32), "gemm: accumulate needs fp32 C");
  EGOT2_CHECK(a.split_k == 1 || (a.accumulate && !a.relu && !a.mask && a.p_drop == 0.f),
              "gemm: split-K only for plain accumulating GEMMs");
  const bool plain = !a.trans_a && !a.relu && !a.mask && a.p_drop <= 0.f && !a.residual && !a.accumulate && a.split_k <= 1 &&
                     a.split_stride == 0 && !a.a_rpg && !a.b_rpg && !a.c_rpg && !a.trans_c;
  static const bool skinny_on = !(getenv("EGOT2_GEMM_SKINNY") && getenv("EGOT2_GEMM_SKINNY")[0] == '0');
  if (skinny_on && plain && a.in_dtype == EGOT2_BF16 && ((a.N <= 8 && a.trans_b) || (a.K <= 8 && !a.trans_b)))
    return a.out_dtype == EGOT2_F32 ? launch_skinny<float>(a, st) : launch_skinny<bf16>(a, st);
  if (a.in_dtype == EGOT2_F32 && a.out_dtype == EGOT2_F32) return launch<float, float>(a, st);
  if (a.in_dtype == EGOT2_BF16 && a.out_dtype == EGOT2_BF16) return launch<bf16, bf16>(a, st);
  if (a.in_dtype == EGOT2_BF16 && a.out_dtype == EGOT2_F32) return launch<bf16, float>(a, st);
  EGOT2_CHECK(false, "gemm: unsupported dtype combination in=%d out=%d", a.in_dtype, a.out_dtype);
}

}  // namespace egot2
