// ffn_sm100.cu — fused feed-forward block of the post-norm encoder layer on tcgen05 tensor cores (H = 128):
//
//     x_out = LayerNorm2( x1 + dropout2( dropout(relu(x1 W1^T + b1)) W2^T + b2 ) )
//
// One CTA owns 128 tokens.  The (tokens x FF) hidden activation never makes a round trip through HBM between the
// two GEMMs: it is produced 128 columns at a time in TMEM, bias+ReLU(+dropout)'d in registers, written as a bf16
// K-major swizzled A operand into shared memory and immediately consumed by the second GEMM, whose (128 x 128)
// accumulator stays in TMEM for the whole FF loop.  The epilogue adds b2 and the residual (read back from the x1
// tile that is still in shared memory) and applies LayerNorm2 with one token row per thread pair.
//
//   warp 0      TMA producer: x1 tile once, then per FF chunk the W1 rows and W2 columns (2-stage ring)
//   warp 1      MMA issuer:   GEMM1(c) -> acc1[c&1]; GEMM2(c-1) -> acc2   (GEMM1 of chunk c overlaps epilogue c-1)
//   warps 2-9   epilogue:     two warps per TMEM lane quadrant, 64 hidden columns each
// TMEM: acc1 double-buffered (2 x 128 columns) + acc2 (128 columns).
#include "ops.h"
#include "sm100.cuh"

namespace egot2 {

using namespace sm100;

namespace {

constexpr int H = 128;
constexpr int BM = 128;
constexpr int FC = 128;                    // hidden columns per chunk
constexpr int NST = 2;                     // weight ring stages
constexpr int NTHREADS = 320;
constexpr uint32_t TILE = 128 * 128 * 2;   // one 128x128 bf16 operand tile = two 16 KB K-halves
constexpr uint32_t HALF = 16384;

struct FfnArgs {
  int M, FF;
  const float* b1; const float* b2; const float* ln_g; const float* ln_b; float eps;
  bf16* hid; bf16* y2; float* stat2; bf16* x_out;
  float p_drop; uint64_t key_ffn, key_drop2;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major SW128 descriptor for k16 step kk (0..7) of a 128x128 tile made of two 64-wide halves
__device__ __forceinline__ uint64_t kdesc(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + (uint32_t)(kk >> 2) * HALF + (uint32_t)(kk & 3) * 32, 16, 1024);
}

__global__ void __launch_bounds__(NTHREADS, 1)
ffn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base;                         // 32 KB
  const uint32_t sW1 = sX + TILE;                   // NST x 32 KB
  const uint32_t sW2 = sW1 + NST * TILE;            // NST x 32 KB
  const uint32_t sHid = sW2 + NST * TILE;           // 32 KB
  const uint32_t bars = sHid + TILE;
  const uint32_t x_full = bars, w_full = bars + 8, w_empty = w_full + 8 * NST, a1_full = w_empty + 8 * NST,
                 a1_empty = a1_full + 16, h_full = a1_empty + 16, h_empty = h_full + 8, a2_full = h_empty + 8,
                 tmem_slot = a2_full + 8;
  const uint32_t red_off = (tmem_slot + 8 + 15u) & ~15u;   // float red[2][128] for the LayerNorm row statistics (16 B aligned)
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to `base`
  float* red = reinterpret_cast<float*>(gen + (red_off - base));
  // bias / LayerNorm vectors live in shared memory: every lane of an epilogue warp reads the same column values, so
  // these are conflict-free broadcasts instead of a chain of dependent global loads on the per-chunk critical path
  float* sVec = red + 256;                 // b2[128], ln_g[128], ln_b[128]
  float* sB1 = sVec + 384;                 // b1[FF]
  for (int i = threadIdx.x; i < a.FF; i += NTHREADS) sB1[i] = a.b1[i];
  for (int i = threadIdx.x; i < 128; i += NTHREADS) { sVec[i] = a.b2[i]; sVec[128 + i] = a.ln_g[i]; sVec[256 + i] = a.ln_b[i]; }
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int NC = a.FF / FC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w1); tma_prefetch_desc(&tm_w2);
    mbar_init(x_full, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(a1_full + 8 * s, 1); mbar_init(a1_empty + 8 * s, 8); }
    mbar_init(h_full, 8); mbar_init(h_empty, 1); mbar_init(a2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t acc2 = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(x_full, TILE);
      tma_load_2d(sX, &tm_x, x_full, 0, m0);
      tma_load_2d(sX + HALF, &tm_x, x_full, 64, m0);
      for (int c = 0; c < NC; ++c) {
        const int s = c % NST;
        mbar_wait(w_empty + 8 * s, ((c / NST) & 1) ^ 1);
        mbar_expect_tx(w_full + 8 * s, 2 * TILE);
        tma_load_2d(sW1 + s * TILE, &tm_w1, w_full + 8 * s, 0, c * FC);            // W1 rows c*128.., k 0..63
        tma_load_2d(sW1 + s * TILE + HALF, &tm_w1, w_full + 8 * s, 64, c * FC);    //                  k 64..127
        tma_load_2d(sW2 + s * TILE, &tm_w2, w_full + 8 * s, c * FC, 0);            // W2 all 128 rows, k = ff c*128..+63
        tma_load_2d(sW2 + s * TILE + HALF, &tm_w2, w_full + 8 * s, c * FC + 64, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, 128, false, false);
      mbar_wait(x_full, 0);
      for (int c = 0; c <= NC; ++c) {
        if (c < NC) {                                   // GEMM1(c): acc1[c&1] = x1 . W1c^T
          const int s = c % NST, bsel = c & 1;
          mbar_wait(w_full + 8 * s, (c / NST) & 1);
          mbar_wait(a1_empty + 8 * bsel, ((c >> 1) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem + bsel * 128, kdesc(sX, kk), kdesc(sW1 + s * TILE, kk), idesc, kk > 0 ? 1u : 0u);
          umma_commit(a1_full + 8 * bsel);
        }
        if (c > 0) {                                    // GEMM2(c-1): acc2 += hid(c-1) . W2c^T
          const int cp = c - 1, s = cp % NST;
          mbar_wait(h_full, cp & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(acc2, kdesc(sHid, kk), kdesc(sW2 + s * TILE, kk), idesc, (cp > 0 || kk > 0) ? 1u : 0u);
          umma_commit(h_empty);                         // hid tile may be overwritten
          umma_commit(w_empty + 8 * s);                 // weight stage may be refilled
        }
      }
      umma_commit(a2_full);
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                  // TMEM lane quadrant
    const int ch = (warp - 2) >> 2;          // which 64-column half of the 128-column tile
    const int r = q * 32 + lane;             // row inside the tile == TMEM lane
    const int m = m0 + r;
    const bool row_ok = m < a.M;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint32_t hrow = sHid + ch * HALF + (uint32_t)r * 128;
    for (int c = 0; c < NC; ++c) {
      const int bsel = c & 1;
      mbar_wait(a1_full + 8 * bsel, (c >> 1) & 1);
      tc_fence_after();
      uint32_t packed[32];
      uint32_t rr0[32], rr1[32];
      tmem_ld_32x32(tmem + lane_addr + bsel * 128 + ch * 64, rr0);
      tmem_ld_32x32(tmem + lane_addr + bsel * 128 + ch * 64 + 32, rr1);
      tmem_ld_wait();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const uint32_t (&rr)[32] = h2 ? rr1 : rr0;
        const int n0 = c * FC + ch * 64 + h2 * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sB1 + n0 + j);
          float v0 = fmaxf(__uint_as_float(rr[j]) + b4.x, 0.f), v1 = fmaxf(__uint_as_float(rr[j + 1]) + b4.y, 0.f);
          float v2 = fmaxf(__uint_as_float(rr[j + 2]) + b4.z, 0.f), v3 = fmaxf(__uint_as_float(rr[j + 3]) + b4.w, 0.f);
          if (a.p_drop > 0.f) {
            const uint64_t idx = (uint64_t)m * a.FF + n0 + j;
            v0 *= drop_scale(a.key_ffn, idx, a.p_drop, inv_keep); v1 *= drop_scale(a.key_ffn, idx + 1, a.p_drop, inv_keep);
            v2 *= drop_scale(a.key_ffn, idx + 2, a.p_drop, inv_keep); v3 *= drop_scale(a.key_ffn, idx + 3, a.p_drop, inv_keep);
          }
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1), p1 = __floats2bfloat162_rn(v2, v3);
          packed[h2 * 16 + j / 2] = *reinterpret_cast<uint32_t*>(&p0);
          packed[h2 * 16 + j / 2 + 1] = *reinterpret_cast<uint32_t*>(&p1);
        }
      }
      // acc1[bsel] fully read by this warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_empty + 8 * bsel);
      // the hid tile is free once GEMM2(c-1) retired
      if (c > 0) mbar_wait(h_empty, (c - 1) & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {           // 8 x 16 B chunks = this thread's 64 columns, 128B-swizzled row
        const uint32_t dst = hrow + (uint32_t)((j ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * j]), "r"(packed[4 * j + 1]),
                     "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3]) : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_full);
      if (a.hid && row_ok) {
        uint4* gp = reinterpret_cast<uint4*>(a.hid + (size_t)m * a.FF + c * FC + ch * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) gp[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
    }
    // ------------------------------------------------------------ final: + b2, dropout2, + residual, LayerNorm2
    mbar_wait(a2_full, 0);
    tc_fence_after();
    float y[64];
    const int nb = ch * 64;
    const uint32_t xrow = sX + ch * HALF + (uint32_t)r * 128;
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint32_t rr[32];
      tmem_ld_32x32(acc2 + lane_addr + nb + h2 * 32, rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) y[h2 * 32 + j] = __uint_as_float(rr[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j8 = 0; j8 < 8; ++j8) {
      uint32_t w0, w1, w2, w3;                  // residual: 8 bf16 of x1 from the swizzled smem tile
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                   : "r"(xrow + (uint32_t)((j8 ^ (r & 7)) << 4)));
      const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j8 * 8 + 2 * k;
        float v0 = y[j] + sVec[nb + j], v1 = y[j + 1] + sVec[nb + j + 1];
        if (a.p_drop > 0.f) {
          v0 *= drop_scale(a.key_drop2, (uint64_t)m * H + nb + j, a.p_drop, inv_keep);
          v1 *= drop_scale(a.key_drop2, (uint64_t)m * H + nb + j + 1, a.p_drop, inv_keep);
        }
        v0 += __uint_as_float(w[k] << 16);
        v1 += __uint_as_float(w[k] & 0xffff0000u);
        // LayerNorm sees the bf16-rounded pre-norm activation, exactly what is saved for backward
        v0 = __bfloat162float(__float2bfloat16_rn(v0));
        v1 = __bfloat162float(__float2bfloat16_rn(v1));
        y[j] = v0; y[j + 1] = v1;
        sum += v0 + v1;
      }
    }
    red[ch * 128 + r] = sum;
    named_bar_sync(1, 256);
    const float mean = (red[r] + red[128 + r]) * (1.f / H);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { const float d = y[j] - mean; sq += d * d; }
    named_bar_sync(2, 256);
    red[ch * 128 + r] = sq;
    named_bar_sync(1, 256);
    const float rstd = rsqrtf((red[r] + red[128 + r]) * (1.f / H) + a.eps);
    if (row_ok) {
      if (ch == 0) { a.stat2[2 * (size_t)m] = mean; a.stat2[2 * (size_t)m + 1] = rstd; }
      uint4* yp = reinterpret_cast<uint4*>(a.y2 + (size_t)m * H + nb);
      uint4* op = reinterpret_cast<uint4*>(a.x_out + (size_t)m * H + nb);
#pragma unroll
      for (int j8 = 0; j8 < 8; ++j8) {
        uint32_t yy[4], oo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j8 * 8 + 2 * k;
          __nv_bfloat162 t = __floats2bfloat162_rn(y[j], y[j + 1]);
          yy[k] = *reinterpret_cast<uint32_t*>(&t);
          const float o0 = (y[j] - mean) * rstd * sVec[128 + nb + j] + sVec[256 + nb + j];
          const float o1 = (y[j + 1] - mean) * rstd * sVec[128 + nb + j + 1] + sVec[256 + nb + j + 1];
          __nv_bfloat162 u = __floats2bfloat162_rn(o0, o1);
          oo[k] = *reinterpret_cast<uint32_t*>(&u);
        }
        yp[j8] = make_uint4(yy[0], yy[1], yy[2], yy[3]);
        op[j8] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ffn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
int kmajor_map(CUtensorMap* map, const void* basep, int inner, int rows, int ld, int box_rows) {
  EncodeTiledFn fn = ffn_encode_fn();
  EGOT2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(basep), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGOT2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace

bool ffn_fused_supported(int dtype, int Hdim, int FF) {
  return dtype == EGOT2_BF16 && Hdim == H && FF >= FC && FF % FC == 0;
}

// x_out = LN2(x1 + drop2(drop(relu(x1 W1^T + b1)) W2^T + b2)); also writes hid (if non-null), y2, stat2
int ffn_fused_fwd(int M, int FF, const void* x1, const void* W1, const float* b1, const void* W2, const float* b2,
                  const float* ln_g, const float* ln_b, float eps, void* hid, void* y2, float* stat2, void* x_out,
                  float p_drop, uint64_t key_ffn, uint64_t key_drop2, cudaStream_t st) {
  if (M == 0) return 0;
  CUtensorMap tx, tw1, tw2;
  EGOT2_TRY(kmajor_map(&tx, x1, H, M, H, 128));
  EGOT2_TRY(kmajor_map(&tw1, W1, H, FF, H, 128));
  EGOT2_TRY(kmajor_map(&tw2, W2, FF, H, FF, 128));
  FfnArgs a;
  a.M = M; a.FF = FF; a.b1 = b1; a.b2 = b2; a.ln_g = ln_g; a.ln_b = ln_b; a.eps = eps;
  a.hid = (bf16*)hid; a.y2 = (bf16*)y2; a.stat2 = stat2; a.x_out = (bf16*)x_out;
  a.p_drop = p_drop; a.key_ffn = key_ffn; a.key_drop2 = key_drop2;
  const size_t smem = 1024 + (size_t)TILE * (1 + 2 * NST + 1) + 256 + (256 + 384 + (size_t)FF) * 4;
  EGOT2_CHECK(smem <= 227 * 1024, "ffn_fused_fwd: FF=%d does not fit the bias stage in shared memory", FF);
  static size_t set_for = 0;
  if (set_for < smem) {
    EGOT2_CUDA(cudaFuncSetAttribute(ffn_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set_for = smem;
  }
  ProfScope prof(st, "ffn_fwd_sm100 M%d H128 FF%d", M, FF);
  ffn_fwd_sm100_kernel<<<(M + BM - 1) / BM, NTHREADS, smem, st>>>(tx, tw1, tw2, a);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
