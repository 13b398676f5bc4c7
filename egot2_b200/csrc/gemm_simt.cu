// gemm_simt.cu — generic CUDA-core GEMM (fp32 accumulate).
//
// This is the arithmetic of the EGOT2_F32 parity mode (fp32 FMA, so logits stay within 1e-3 of
// the reference and argmax stays bit-exact) and the shape-general fallback *inside the CUDA
// library* for bf16 GEMMs whose shape the tcgen05 kernel (gemm_sm100.cu) does not take.
// Register-tiled: BMxBN block tile, BK=16 k-slab staged k-major in shared memory, TMxTN
// outputs per thread, 256 threads.  Handles all four operand orientations, storage-row
// remapping (segment scatter inside (B,T,H) tensors), the fused epilogue of ops.h and split-K
// with fp32 atomics for the weight-gradient GEMMs (K = all tokens of the batch).
#define EGOT2_FILE_ID 10
#include "ops.h"

namespace egot2 {

namespace {

template <typename T> __device__ __forceinline__ float ld(const T* p) { return to_f32(__ldg(p)); }

__device__ __forceinline__ long long remap(int r, int rpg, int gstride) {
  return rpg > 0 ? (long long)(r / rpg) * gstride + (r % rpg) : (long long)r;
}

template <typename TI, typename TO, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs a) {
  EGOT2_PDL_ENTER();
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "thread tiling");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

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

  constexpr int EA = BM * BK / NT;   // elements of the A tile per thread
  constexpr int EB = BN * BK / NT;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile into As[k][m]
    if (!a.trans_a) {                       // stored (M,K): k contiguous
      constexpr int TPR = BK / EA;          // threads per row
      const int r = tid / TPR, kk = (tid % TPR) * EA;
      const int m = m0 + r;
      const TI* src = A + remap(m, a.a_rpg, a.a_gstride) * a.lda + k0 + kk;
#pragma unroll
      for (int e = 0; e < EA; ++e)
        As[kk + e][r] = (m < a.M && k0 + kk + e < kend) ? ld(src + e) : 0.f;
    } else {                                // stored (K,M): m contiguous
      constexpr int TPR = BM / EA;
      const int kk = tid / TPR, r = (tid % TPR) * EA;
      const int k = k0 + kk;
      const TI* src = A + remap(k, a.a_rpg, a.a_gstride) * a.lda + m0 + r;
#pragma unroll
      for (int e = 0; e < EA; ++e)
        As[kk][r + e] = (k < kend && m0 + r + e < a.M) ? ld(src + e) : 0.f;
    }
    // ---- stage B tile into Bs[k][n]
    if (a.trans_b) {                        // stored (N,K): k contiguous
      constexpr int TPR = BK / EB;
      const int r = tid / TPR, kk = (tid % TPR) * EB;
      const int n = n0 + r;
      const TI* src = B + remap(n, a.b_rpg, a.b_gstride) * a.ldb + k0 + kk;
#pragma unroll
      for (int e = 0; e < EB; ++e)
        Bs[kk + e][r] = (n < a.N && k0 + kk + e < kend) ? ld(src + e) : 0.f;
    } else {                                // stored (K,N): n contiguous
      constexpr int TPR = BN / EB;
      const int kk = tid / TPR, r = (tid % TPR) * EB;
      const int k = k0 + kk;
      const TI* src = B + remap(k, a.b_rpg, a.b_gstride) * a.ldb + n0 + r;
#pragma unroll
      for (int e = 0; e < EB; ++e)
        Bs[kk][r + e] = (k < kend && n0 + r + e < a.N) ? ld(src + e) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
  const bool first_split = blockIdx.z == 0;
  TO* __restrict__ C = (TO*)a.C;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= a.M) continue;
    const long long crow = remap(m, a.c_rpg, a.c_gstride);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
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
  if (a.in_dtype == EGOT2_F32 && a.out_dtype == EGOT2_F32) return launch<float, float>(a, st);
  if (a.in_dtype == EGOT2_BF16 && a.out_dtype == EGOT2_BF16) return launch<bf16, bf16>(a, st);
  if (a.in_dtype == EGOT2_BF16 && a.out_dtype == EGOT2_F32) return launch<bf16, float>(a, st);
  EGOT2_CHECK(false, "gemm: unsupported dtype combination in=%d out=%d", a.in_dtype, a.out_dtype);
}

}  // namespace egot2
