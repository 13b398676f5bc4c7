// ffn_sm100.cu — fused feed-forward block of the post-norm encoder layer on tcgen05 tensor cores (H = 128):
//
//     x_out = LayerNorm2( x1 + dropout2( dropout(relu(x1 W1^T + b1)) W2^T + b2 ) )
//
// A CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) owns 256 tokens, 128 per CTA.  The (tokens x FF) hidden activation
// never makes a round trip through HBM between the two GEMMs: it is produced 128 columns at a time in TMEM,
// bias+ReLU(+dropout)'d in registers, written as a bf16 K-major swizzled A operand into shared memory and immediately
// consumed by the second GEMM, whose (128 x 128 per CTA) accumulator stays in TMEM for the whole FF loop.  The same
// swizzled tile is what TMA stores to HBM as the saved activation for backward.  The final epilogue adds b2 and the
// residual (read back from the x1 tile still in shared memory), applies LayerNorm2 with one token row per thread pair,
// and leaves through shared memory + TMA as well.
//
// Why a CTA pair: every token tile re-reads all of W1 and W2 (1 MB) from L2, and with one CTA per 128 tokens that
// L2 -> SM fill (64 KB per 128-column chunk per SM, ~28 B/clk/SM measured, all 148 SMs pulling) bounded the kernel,
// not the MMAs.  With cta_group::2 the B operand (the weights) is SPLIT across the pair - each SM fetches and holds only
// 64 of the 128 rows of every weight stage - so the fill per SM halves, and the M=256 MMA runs at the 64-cycle floor
// instead of the 83 cycles a lone M=128 x N=128 SS-MMA takes (shared-memory operand bandwidth; tools/ubench/umma_rate.cu).
//
//   warp 0      TMA producer (both CTAs): x1 tile, then this CTA's half of every weight stage; two rings (W1: 4, W2: 3
//               stages of 16 KB = 64 rows x 128 k), polled so that neither ring holds the other back; ~4 chunks of loads
//               in flight (TMA latency from L2 under load was measured at 2.3-3k cycles)
//   warp 1      leader: MMA issuer A, GEMM1(c) -> acc1[c&1]      warp 11  leader: MMA issuer B, GEMM2(c) -> acc2
//               (two issuing threads: every mbarrier wait / commit costs its thread 100-150 cycles in which it queues no
//               MMAs - tools/ubench/umma_issue.cu - so one stream's synchronisation overlaps the other stream's MMAs)
//   warps 2-9   epilogue (both CTAs): two warps per TMEM lane quadrant, 64 hidden columns each
//   warp 10     TMA store of the hidden tile (both CTAs)
// TMEM (per CTA): acc1 double-buffered (AB x 128 columns) + acc2 (128 columns).
// Barriers that the leader's MMA threads wait on (operands of BOTH CTAs ready, accumulators drained by BOTH epilogues)
// live in the leader and receive remote arrivals; completion barriers are signalled in both CTAs by multicast commits.
#include <stdlib.h>
#include <string.h>

#define EGOT2_FILE_ID 5
#include "ops.h"
#include "sm100.cuh"

namespace egot2 {

using namespace sm100;

namespace {

constexpr int H = 128;
constexpr int BM = 128;                    // tokens per CTA (256 per pair)
constexpr int FC = 128;                    // hidden columns per chunk
// Pipeline depths (compile-time knobs for experiments).  Measured on B200 (HHI b256): 2/2/4/3, 3/2/4/3 and 3/3/3/2 (acc1
// buffers / hidden buffers / W1 ring / W2 ring) all run the forward in 53-55 us - the chunk period (~2.3k cycles for 16
// M=256 x N=128 x K=16 MMAs) is not set by the depth of the GEMM1 -> epilogue -> GEMM2 loop but by SHARED-MEMORY BANDWIDTH:
// per chunk and CTA the SS-mode MMAs read 96 KB of operands (x tile 32, hidden tile 32, weight halves 32), the epilogue
// writes the 32 KB hidden tile, TMA reads it back for the saved activation and writes 32 KB of weight stages: 192 KB at
// 128 B/clk = 1.5k cycles before bank conflicts.  Going further needs operands in TMEM (tcgen05 TS mode), not more stages.
#ifndef EGOT2_FFN_AB
#define EGOT2_FFN_AB 2
#define EGOT2_FFN_HB 2
#define EGOT2_FFN_R1 4
#define EGOT2_FFN_R2 3
#endif
constexpr int R1 = EGOT2_FFN_R1, R2 = EGOT2_FFN_R2;   // W1 / W2 ring stages (16 KB: this CTA's 64 rows x 128 k, two 8 KB k-halves)
constexpr int RING = R1 + R2;
constexpr int HB = EGOT2_FFN_HB;           // hidden-tile buffers in shared memory
constexpr int AB = EGOT2_FFN_AB;           // acc1 buffers in TMEM (acc2 follows them)
// EGOT2_FFN_TS=1: GEMM1's A operand (the x1 / d2 token tile, the same for every chunk) lives in TENSOR MEMORY - the epilogue
// warps copy it there once per tile (64 columns of packed bf16 pairs, tcgen05.st) and the per-chunk MMAs take A from TMEM
// (tcgen05.mma TS form) instead of re-reading 32 KB of shared memory.  Parity-green on B200 but no faster (53.4 vs 52.9 us
// forward), so the SS form stays the default; kept as the starting point for moving the hidden tile to TMEM as well.
#ifndef EGOT2_FFN_TS
#define EGOT2_FFN_TS 0
#endif
constexpr bool TS = EGOT2_FFN_TS != 0;
constexpr int XB = AB * 128 + 128;         // TMEM column of the A tile (after acc1 buffers and acc2)
static_assert(AB * 128 + 128 + (TS ? 64 : 0) <= 512, "TMEM: acc1 buffers + acc2 (+ A tile) must fit 512 columns");
// Epilogue parallelism: EG column groups x 4 TMEM lane quadrants = 4*EG epilogue warps per CTA, each thread one token row x
// CW = 128/EG hidden columns of every chunk.  EG = 2 (8 warps x 64 columns, 144 registers) left two epilogue warps per
// scheduler: each warp's own chain per chunk (accumulator wait -> tcgen05.ld -> ~330 dependent ALU instructions -> st.shared ->
// fence -> arrive) took ~1.9k cycles at 32 % issue utilisation and set the chunk period, not the MMAs (1.0k) or shared memory
// (1.5k).  EG = 4 (16 warps x 32 columns, <= 96 registers) halves that chain and doubles the warps that hide it.
#ifndef EGOT2_FFN_EG
#define EGOT2_FFN_EG 4
#endif
// EGOT2_FFN_DBG (timing experiments only - results are WRONG): bit 0 = the epilogue does not write the hidden tile to shared
// memory, bit 1 = no TMA store of the hidden tile, bit 2 = the epilogue skips its arithmetic, bit 3 = GEMM2 is not issued,
// bit 4 = GEMM1 is not issued, bit 5 = the weight stages are loaded only while the rings first fill (stale weights afterwards:
// no L2 -> SM weight traffic in the steady state)
#ifndef EGOT2_FFN_DBG
#define EGOT2_FFN_DBG 0
#endif
constexpr int DBG = EGOT2_FFN_DBG;
constexpr int EG = EGOT2_FFN_EG;
constexpr int CW = 128 / EG;               // hidden columns per epilogue thread and chunk
constexpr int EPW = 4 * EG;                // epilogue warps per CTA (warps 2 .. 2+EPW-1)
constexpr int WS = 2 + EPW;                // TMA-store warp
constexpr int WB = 3 + EPW;                // MMA issuer B
constexpr int NTHREADS = 32 * (4 + EPW);
static_assert(EG == 2 || EG == 4, "EG: 2 or 4 column groups");
static_assert(!TS || EG == 2, "the TS (A tile in tensor memory) experiment is written for EG = 2");
constexpr uint32_t TILE = 128 * 128 * 2;   // one 128x128 bf16 operand tile = two 16 KB K-halves
constexpr uint32_t HALF = 16384;
constexpr uint32_t STAGE = 16384;          // one weight stage

struct FfnArgs {
  int M, FF;
  const float* b1; const float* b2; const float* ln_g; const float* ln_b; float eps;
  float* stat2;
  int save_hid;
  int hid_tiled;         // hidden tensor layout: 0 = row-major (M, FF); 1 = tile-major [M/128][FF/64][128][64] (16 KB contiguous per TMA box)
  uint2* hmask;          // [FF/64][M] x 64 bits: hidden activation != 0 (ReLU gate x dropout keep), for ffn_bwd_dx
  const bf16* d1;        // backward: gradient arriving through the residual branch (added to d3), (M,128)
  float* db1;            // backward, optional: db1[FF] += column sums of dhid (the linear1 bias gradient), so that no
                         // separate kernel has to re-read the (M,FF) dhid tensor for it
  float p_drop; uint64_t key_ffn, key_drop2;
  // FF-split tail (see ffn_launch): CTA pairs >= split_pair0 each own 1/S of the FF range of a tail token tile and add
  // their partial GEMM2 accumulator into `partial` (fp32, rows relative to token split_pair0*256); a fix-up kernel finishes
  int split_pair0, S;
  float* partial;
  int* tile_ctr;         // one arrival counter per 128-row tail tile (zero on entry and on exit); nullptr: separate fix-up kernels
  const bf16* fix_x1; const bf16* fix_d1; bf16* fix_y2; bf16* fix_out;     // global tensors the fix-up kernel reads / writes
  long long* trace;      // EGOT2_FFN_TRACE builds only: per-chunk clock64 stamps of CTA 0
};

#ifdef EGOT2_FFN_TRACE
#define TR(c, slot) do { if (blockIdx.x == 0 && a.trace) a.trace[(c) * 16 + (slot)] = clock64(); } while (0)
// per-warp stamps of the first cluster (both CTAs): [cta][warp][chunk][k], k = 0 accumulator seen, 1 hidden tile handed over
// (globaltimer: comparable across the two SMs)
#define TRW(c, k) do { if (blockIdx.x < 2 && a.trace && lane == 0 && (c) < 16) { unsigned long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.trace[5120 + (((blockIdx.x * 24 + warp) * 16 + (c)) * 2 + (k))] = (long long)t_; } } while (0)
#else
#define TR(c, slot) do { } while (0)
#define TRW(c, k) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// K-major SW128 descriptor for k16 step kk (0..7) of a 128-row x 128-k tile made of two 64-wide k-halves
__device__ __forceinline__ uint64_t kdesc(uint32_t tile, int kk) {
  return make_smem_desc_sw128(tile + (uint32_t)(kk >> 2) * HALF + (uint32_t)(kk & 3) * 32, 16, 1024);
}
// weight stage = this CTA's half of the B operand.  K-major (forward): 64 n-rows x 128 k as two 8 KB k-halves.
// MN-major (backward): 128 k-rows x 64 n (one 128 B swizzle atom wide): k16 step kk starts 16 rows further down.
template <bool MN>
__device__ __forceinline__ uint64_t wdesc(uint32_t stage, int kk) {
  return MN ? make_smem_desc_sw128(stage + (uint32_t)kk * 2048, 8192, 1024)
            : make_smem_desc_sw128(stage + (uint32_t)(kk >> 2) * 8192 + (uint32_t)(kk & 3) * 32, 16, 1024);
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// Finishing one token row of an FF-split tail tile once all S partial sums are in `prow` (fp32, this row's 128 columns): what
// the main kernel's final epilogue does for a regular tile, by ONE WARP (4 columns per lane); the scratch row is left zeroed.
// Called by the stand-alone fix-up kernels (default) or by the last-arriving slice of the tile (EGOT2_FFN_FIXUP=inkernel).
__device__ __forceinline__ void fixup_row_fwd(int lane, int m, float* __restrict__ prow, const bf16* __restrict__ x1,
                                              const float* __restrict__ b2, const float* __restrict__ ln_g,
                                              const float* __restrict__ ln_b, float eps, float p_drop, uint64_t key_drop2,
                                              unsigned long long ep, bf16* __restrict__ y2, float* __restrict__ stat2,
                                              bf16* __restrict__ x_out) {
  const int c0 = lane * 4;
  float4* pp = reinterpret_cast<float4*>(prow + c0);
  const float4 acc = ld_cg_f4(prow + c0);                   // the partials were added in L2 (red.global): read them there
  *pp = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint2 xw = *reinterpret_cast<const uint2*>(x1 + (size_t)m * H + c0);
  const float4 bb = *reinterpret_cast<const float4*>(b2 + c0);
  float v[4] = {acc.x + bb.x, acc.y + bb.y, acc.z + bb.z, acc.w + bb.w};
  if (p_drop > 0.f) {
    float dm[4];
    drop_scale_n<4>(key_drop2 ^ ep, (uint64_t)m * H + c0, p_drop, 1.f / (1.f - p_drop), dm);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = dm[i] * v[i];
  }
  v[0] += __uint_as_float(xw.x << 16); v[1] += __uint_as_float(xw.x & 0xffff0000u);
  v[2] += __uint_as_float(xw.y << 16); v[3] += __uint_as_float(xw.y & 0xffff0000u);
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  uint2 yw; yw.x = *reinterpret_cast<uint32_t*>(&t0); yw.y = *reinterpret_cast<uint32_t*>(&t1);
  *reinterpret_cast<uint2*>(y2 + (size_t)m * H + c0) = yw;
  v[0] = __uint_as_float(yw.x << 16); v[1] = __uint_as_float(yw.x & 0xffff0000u);
  v[2] = __uint_as_float(yw.y << 16); v[3] = __uint_as_float(yw.y & 0xffff0000u);
  float sum = (v[0] + v[1]) + (v[2] + v[3]);
  float sq = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])));
  sum = warp_sum(sum); sq = warp_sum(sq);
  const float mean = sum * (1.f / H);
  const float var = fmaxf(sq * (1.f / H) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) { stat2[2 * (size_t)m] = mean; stat2[2 * (size_t)m + 1] = rstd; }
  const float4 g4 = *reinterpret_cast<const float4*>(ln_g + c0), e4 = *reinterpret_cast<const float4*>(ln_b + c0);
  __nv_bfloat162 o0 = __floats2bfloat162_rn((v[0] - mean) * rstd * g4.x + e4.x, (v[1] - mean) * rstd * g4.y + e4.y);
  __nv_bfloat162 o1 = __floats2bfloat162_rn((v[2] - mean) * rstd * g4.z + e4.z, (v[3] - mean) * rstd * g4.w + e4.w);
  uint2 ow; ow.x = *reinterpret_cast<uint32_t*>(&o0); ow.y = *reinterpret_cast<uint32_t*>(&o1);
  *reinterpret_cast<uint2*>(x_out + (size_t)m * H + c0) = ow;
}
// backward: d3 = partial + d1 (the gradient arriving through the residual branch)
__device__ __forceinline__ void fixup_row_bwd(int lane, int m, float* __restrict__ prow, const bf16* __restrict__ d1,
                                              bf16* __restrict__ d3) {
  const int c0 = lane * 4;
  const float4 acc = ld_cg_f4(prow + c0);
  *reinterpret_cast<float4*>(prow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint2 dw = *reinterpret_cast<const uint2*>(d1 + (size_t)m * H + c0);
  __nv_bfloat162 t0 = __floats2bfloat162_rn(acc.x + __uint_as_float(dw.x << 16), acc.y + __uint_as_float(dw.x & 0xffff0000u));
  __nv_bfloat162 t1 = __floats2bfloat162_rn(acc.z + __uint_as_float(dw.y << 16), acc.w + __uint_as_float(dw.y & 0xffff0000u));
  uint2 ow; ow.x = *reinterpret_cast<uint32_t*>(&t0); ow.y = *reinterpret_cast<uint32_t*>(&t1);
  *reinterpret_cast<uint2*>(d3 + (size_t)m * H + c0) = ow;
}

// BWD = false: the forward block described above.
// BWD = true : the data-gradient half of its backward, same skeleton with the roles swapped -
//     dhid_c = (d2 . W2[:, c]) * gate_c / (1-p)   (GEMM1: A = d2 tile, B = W2 columns, MN-major; gate bits from hmask)
//     d3     = sum_c dhid_c . W1[c, :] + d1        (GEMM2: A = dhid tile, B = W1 rows, MN-major)
//   dhid leaves through TMA (it feeds the two weight-gradient GEMMs), d3 replaces the forward's y2 output.
// DROP: 0 = no dropout (inference), 1 = p == 0.5 (one random bit per hidden element), 2 = general p (hash per element).
// A template parameter, not a runtime branch: the three variants of the fully unrolled per-chunk epilogue would triple a
// body that already exceeds the 32 KB L1.5 instruction cache.
template <bool BWD, int DROP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
ffn_sm100_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_hid,
                     const __grid_constant__ CUtensorMap tm_y2, const __grid_constant__ CUtensorMap tm_out,
                     const __grid_constant__ CUtensorMap tm_part, const FfnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sX = base;                         // 32 KB
  const uint32_t sW1 = sX + TILE;                   // R1 x 16 KB
  const uint32_t sW2 = sW1 + R1 * STAGE;            // R2 x 16 KB
  const uint32_t sHid = sW2 + R2 * STAGE;           // HB x 32 KB (chunk c uses buffer c % HB)
  const uint32_t bars = sHid + HB * TILE;
  // local barriers
  const uint32_t x_full = bars, r1_empty = x_full + 8, r2_empty = r1_empty + 8 * R1, a1_full = r2_empty + 8 * R2,
                 hl_full = a1_full + 8 * AB, h_empty = hl_full + 8 * HB, hs_empty = h_empty + 8 * HB, a2_full = hs_empty + 8 * HB;
  // barriers used in the leader only (arrivals from both CTAs)
  const uint32_t x_pair = a2_full + 8, r1_full = x_pair + 8, r2_full = r1_full + 8 * R1, a1_empty = r2_full + 8 * R2,
                 h_full = a1_empty + 8 * AB, xt_full = h_full + 8 * HB, tmem_slot = xt_full + 8;
  const uint32_t red_off = (tmem_slot + 8 + 15u) & ~15u;   // float red[2][EG][128] for the LayerNorm row statistics (16 B aligned)
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to `base`
  float* red = reinterpret_cast<float*>(gen + (red_off - base));
  // bias / LayerNorm vectors live in shared memory: every lane of an epilogue warp reads the same column values, so
  // these are conflict-free broadcasts instead of a chain of dependent global loads on the per-chunk critical path
  float* sVec = red + 2 * EG * 128;        // b2[128], ln_g[128], ln_b[128]   (red: sum[EG][128], sumsq[EG][128])
  float* sB1 = sVec + 384;                 // b1[FF]
  pdl_launch_dependents();       // the next kernel may become resident and run its prologue under this one
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // which token tile and which chunk range [c_begin, c_begin + NC) of the FF loop this pair owns (NC = all of it for a
  // regular pair; the loops below count LOCAL chunks, cg = c_begin + c is the chunk's position in FF)
  int tile_pair = (int)blockIdx.x >> 1, c_begin = 0, NC = a.FF / FC;
  const bool split = tile_pair >= a.split_pair0;
  if (split) {
    const int u = tile_pair - a.split_pair0;
    tile_pair = a.split_pair0 + u / a.S;
    NC /= a.S;
    c_begin = (u % a.S) * NC;
  }
  const int m0 = (tile_pair * 2 + ((int)blockIdx.x & 1)) * BM;
#ifdef EGOT2_FFN_TRACE
  if (threadIdx.x == 0 && a.trace) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    a.trace[1024 + blockIdx.x * 4] = (long long)t; a.trace[1024 + blockIdx.x * 4 + 2] = sm;
  }
#endif
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 64) TR(60, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w1); tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_hid); tma_prefetch_desc(&tm_y2); tma_prefetch_desc(&tm_out); tma_prefetch_desc(&tm_part);
    mbar_init(x_full, 1); mbar_init(x_pair, 2); mbar_init(a2_full, 1); mbar_init(xt_full, 2 * EPW);
    for (int s = 0; s < R1; ++s) { mbar_init(r1_full + 8 * s, 1); mbar_init(r1_empty + 8 * s, 1); }
    for (int s = 0; s < R2; ++s) { mbar_init(r2_full + 8 * s, 1); mbar_init(r2_empty + 8 * s, 1); }
    for (int s = 0; s < AB; ++s) { mbar_init(a1_full + 8 * s, 1); mbar_init(a1_empty + 8 * s, 2 * EPW); }
    for (int s = 0; s < HB; ++s) {
      mbar_init(h_full + 8 * s, 2 * EPW); mbar_init(hl_full + 8 * s, EPW); mbar_init(h_empty + 8 * s, 1); mbar_init(hs_empty + 8 * s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cg2<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // both CTAs' barriers are initialised before anything targets them remotely
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t acc2 = tmem + AB * 128;
  // this CTA's half of the weight stage of local chunk i (the complete_tx goes to the LEADER's stage barrier, which expects
  // the bytes of both halves)
  auto load_w1 = [&](int i) {                 // GEMM1 operand of chunk i into W1-ring stage i % R1
    const int s = i % R1;
    const uint32_t bar = mapa(r1_full + 8 * s, 0);
    if (leader) mbar_expect_tx(r1_full + 8 * s, 2 * STAGE);
    if (!BWD) {       // GEMM1 B = W1 rows (ff) c*128 + rank*64 .., K-major: two 64-k halves
      tma_load_2d_cg2(sW1 + s * STAGE, &tm_w1, bar, 0, (c_begin + i) * FC + rank * 64);
      tma_load_2d_cg2(sW1 + s * STAGE + 8192, &tm_w1, bar, 64, (c_begin + i) * FC + rank * 64);
    } else {          // GEMM1 B = W2[:, ff], MN-major: 128 k-rows (h) x this CTA's 64 ff columns, one box
      tma_load_2d_cg2(sW1 + s * STAGE, &tm_w2, bar, (c_begin + i) * FC + rank * 64, 0);
    }
  };
  auto load_w2 = [&](int i) {                 // GEMM2 operand of chunk i into W2-ring stage i % R2
    const int s = i % R2;
    const uint32_t bar = mapa(r2_full + 8 * s, 0);
    if (leader) mbar_expect_tx(r2_full + 8 * s, 2 * STAGE);
    if (!BWD) {       // GEMM2 B = W2 rows (h) rank*64 .., K-major over ff c*128 ..
      tma_load_2d_cg2(sW2 + s * STAGE, &tm_w2, bar, (c_begin + i) * FC, rank * 64);
      tma_load_2d_cg2(sW2 + s * STAGE + 8192, &tm_w2, bar, (c_begin + i) * FC + 64, rank * 64);
    } else {          // GEMM2 B = W1[ff, :], MN-major: 128 k-rows (ff) x this CTA's 64 h columns
      tma_load_2d_cg2(sW2 + s * STAGE, &tm_w1, bar, rank * 64, (c_begin + i) * FC);
    }
  };
  // The first fill of both weight rings is requested BEFORE the programmatic-dependent-launch wait: the weights (and biases)
  // are parameters - nothing the two launches in front of this one write (LayerNorm / GEMM / head kernels of the same layer;
  // the optimizer that updates them is at least a whole stage away) - so their L2 -> SM latency (2-4k cycles at kernel start,
  // when every pair asks for the same lines) overlaps the tail of the previous kernel instead of this kernel's first chunk.
  const int pre1 = NC < R1 ? NC : R1, pre2 = NC < R2 ? NC : R2;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < pre1; ++i) load_w1(i);
    for (int i = 0; i < pre2; ++i) load_w2(i);
  }
  pdl_wait();                    // first access to anything a previous kernel produced is below
  EGOT2_TL(EGOT2_FILE_ID);
  const unsigned long long egot2_ep = epoch_xor();
  // The bias / LayerNorm vectors (forward) and the db1 accumulators (backward) are touched by the epilogue warps only, and
  // those idle until the first accumulator arrives: THEY stage them (named barrier among themselves) while the producer and
  // the MMA threads start the first chunk - a CTA-wide fill + __syncthreads here sat in front of the first TMA load.
  if (warp >= 2 && warp < 2 + EPW) {
    const int et = (int)threadIdx.x - 64;
    if (BWD) {                   // sB1 doubles as this CTA's db1 accumulator (the forward's bias stage is not needed)
      for (int i = et; i < a.FF; i += EPW * 32) sB1[i] = 0.f;
    } else {
      const float s1 = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;    // dropout scale folded into the bias (see epilogue)
      for (int i = et; i < a.FF; i += EPW * 32) sB1[i] = a.b1[i] * s1;
      for (int i = et; i < 128; i += EPW * 32) { sVec[i] = a.b2[i]; sVec[128 + i] = a.ln_g[i]; sVec[256 + i] = a.ln_b[i]; }
    }
    named_bar_sync(3, EPW * 32);
  }
  if (threadIdx.x == 64) TR(60, 1);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(x_full, TILE);
      tma_load_2d(sX, &tm_x, x_full, 0, m0);
      tma_load_2d(sX + HALF, &tm_x, x_full, 64, m0);
      // the rings' first fill was requested in the prologue; from there on the two rings are polled (try_wait) so that a
      // full W2 ring never holds back W1 loads
      int i1 = pre1, i2 = pre2;
      while (i1 < NC || i2 < NC) {
        if (i1 < NC && mbar_try_wait(r1_empty + 8 * (i1 % R1), ((i1 / R1) & 1) ^ 1)) {
          TR(i1, 0);
          if ((DBG & 32)) { if (leader) mbar_arrive(r1_full + 8 * (i1 % R1)); }
          else load_w1(i1);
          ++i1;
        }
        if (i2 < NC && i2 < i1 && mbar_try_wait(r2_empty + 8 * (i2 % R2), ((i2 / R2) & 1) ^ 1)) {
          TR(i2, 1);
          if ((DBG & 32)) { if (leader) mbar_arrive(r2_full + 8 * (i2 % R2)); }
          else load_w2(i2);
          ++i2;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer A (leader): GEMM1(c): acc1[c&1] = x1 . W1c^T
    if (lane == 0) {
      mbar_wait(x_full, 0);                           // this CTA's x1 tile landed ...
      mbar_arrive_cluster(mapa(x_pair, 0));           // ... tell the leader
      if (leader) {
        constexpr uint32_t idesc = make_idesc_bf16(2 * BM, 128, false, BWD);
        mbar_wait(x_pair, 0);
        if (TS) { mbar_wait(xt_full, 0); tc_fence_after(); }      // both CTAs' A tiles are in tensor memory
        for (int c = 0; c < NC; ++c) {
          const int bsel = c % AB, s = c % R1;
          {   // both probes in flight together; the blocking waits only where a probe said "not yet"
            const bool p0 = mbar_test_wait(a1_empty + 8 * bsel, ((c / AB) & 1) ^ 1), p1 = mbar_test_wait(r1_full + 8 * s, (c / R1) & 1);
            if (!p0) mbar_wait(a1_empty + 8 * bsel, ((c / AB) & 1) ^ 1);
            TR(c, 3);
            if (!p1) mbar_wait(r1_full + 8 * s, (c / R1) & 1);
            TR(c, 2);
          }
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < ((DBG & 16) ? 0 : 8); ++kk) {
            if (TS) umma_bf16_cg2_ts(tmem + bsel * 128, tmem + XB + kk * 8, wdesc<BWD>(sW1 + s * STAGE, kk), idesc, kk > 0 ? 1u : 0u);
            else umma_bf16_cg2(tmem + bsel * 128, kdesc(sX, kk), wdesc<BWD>(sW1 + s * STAGE, kk), idesc, kk > 0 ? 1u : 0u);
          }
          umma_commit_cg2(r1_empty + 8 * s, 3);       // stage reusable (both CTAs) once these MMAs retire
          umma_commit_cg2(a1_full + 8 * bsel, 3);
        }
      }
    }
  } else if (warp == WB) {
    // ------------------------------------------------------------ MMA issuer B (leader): GEMM2(c): acc2 += hid(c) . W2c^T
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, 128, false, BWD);
      for (int cp = 0; cp < NC; ++cp) {
        const int hb = cp % HB, s = cp % R2;
        {
          const bool p0 = mbar_test_wait(h_full + 8 * hb, (cp / HB) & 1), p1 = mbar_test_wait(r2_full + 8 * s, (cp / R2) & 1);
          if (!p0) mbar_wait(h_full + 8 * hb, (cp / HB) & 1);
          TR(cp, 5);
          if (!p1) mbar_wait(r2_full + 8 * s, (cp / R2) & 1);
          TR(cp, 4);
        }
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < ((DBG & 8) ? 0 : 8); ++kk)
          umma_bf16_cg2(acc2, kdesc(sHid + hb * TILE, kk), wdesc<BWD>(sW2 + s * STAGE, kk), idesc, (cp > 0 || kk > 0) ? 1u : 0u);
        umma_commit_cg2(r2_empty + 8 * s, 3);
        umma_commit_cg2(h_empty + 8 * hb, 3);         // hid buffers may be overwritten (once the TMA stores have read them too)
      }
      umma_commit_cg2(a2_full, 3);
    }
  } else if (warp == WS) {
    // ------------------------------------------------------------ TMA store of the hidden tile (saved for backward)
    if (lane == 0) {
      for (int c = 0; c < NC; ++c) {
        const int hb = c % HB;
        mbar_wait(hl_full + 8 * hb, (c / HB) & 1);      // this CTA's epilogue wrote (and proxy-fenced) hid(c)
        TR(c, 11);
        if (a.save_hid && !(DBG & 2)) {
          if (a.hid_tiled) {      // box (tile, ff-half) is one contiguous 16 KB block
            const int row = ((m0 / BM) * (a.FF / 64) + (c_begin + c) * 2) * 128;
            tma_store_2d(&tm_hid, sHid + hb * TILE, 0, row);
            tma_store_2d(&tm_hid, sHid + hb * TILE + HALF, 0, row + 128);
          } else {
            tma_store_2d(&tm_hid, sHid + hb * TILE, (c_begin + c) * FC, m0);
            tma_store_2d(&tm_hid, sHid + hb * TILE + HALF, (c_begin + c) * FC + 64, m0);
          }
          tma_store_commit();
          tma_store_wait_read();
        }
        TR(c, 12);
        mbar_arrive(hs_empty + 8 * hb);
      }
      tma_store_wait_read();                  // shared memory may be released; the writes are complete when the grid is
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                  // TMEM lane quadrant
    const int ch = (warp - 2) >> 2;          // which CW-column group of the 128-column tile
    const int half = (ch * CW) >> 6;         // the 64-column K-half of the swizzled tiles that group lives in ...
    const int jb = ((ch * CW) & 63) >> 3;    // ... and its first 16-byte chunk inside the 128-byte swizzle row
    const int r = q * 32 + lane;             // row inside the tile == TMEM lane
    const int m = m0 + r;
    const bool row_ok = m < a.M;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float inv_keep = a.p_drop > 0.f ? 1.f / (1.f - a.p_drop) : 1.f;
    const uint32_t thr16 = drop_threshold16(a.p_drop);
    const uint32_t a1_empty_ldr = mapa(a1_empty, 0), h_full_ldr = mapa(h_full, 0);
    // backward: the ReLU/dropout gate bits of chunk c+1 are fetched while chunk c is processed (a dependent global
    // load in front of every chunk's arithmetic was ~1 us of exposed latency per chunk)
    if constexpr (TS) {
      // this thread's half row of the A tile (64 bf16 = 32 packed words, k ascending) from the swizzled smem tile into TMEM
      mbar_wait(x_full, 0);
      uint32_t xw[32];
      const uint32_t xr = sX + ch * HALF + (uint32_t)r * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(xw[4 * j]), "=r"(xw[4 * j + 1]), "=r"(xw[4 * j + 2]), "=r"(xw[4 * j + 3]) : "r"(xr + (uint32_t)((j ^ (r & 7)) << 4)));
      tmem_st_32x32(tmem + lane_addr + XB + ch * 32, xw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(xt_full, 0));
    }
    // gate words: hmask is [FF/64][M] x 64 bit; this thread owns CW/32 consecutive 32-bit words of its row's 64-bit entry
    const uint32_t* hm32 = reinterpret_cast<const uint32_t*>(a.hmask);
    auto gate_idx = [&](int cgl) { return ((size_t)(cgl * 2 + half) * a.M + m) * 2 + (((ch * CW) & 63) >> 5); };
    uint32_t gate_next[CW / 32] = {};
    if (BWD && row_ok) {
#pragma unroll
      for (int w = 0; w < CW / 32; ++w) gate_next[w] = __ldg(hm32 + gate_idx(c_begin) + w);
    }
    for (int c = 0; c < NC; ++c) {
      const int bsel = c % AB;
      const int hb = c % HB;
      uint32_t gate_cur[CW / 32];
#pragma unroll
      for (int w = 0; w < CW / 32; ++w) gate_cur[w] = gate_next[w];
      const int cg = c_begin + c;
      if (BWD && row_ok && c + 1 < NC) {
#pragma unroll
        for (int w = 0; w < CW / 32; ++w) gate_next[w] = __ldg(hm32 + gate_idx(cg + 1) + w);
      }
      const uint32_t hrow = sHid + hb * TILE + half * HALF + (uint32_t)r * 128;
      mbar_wait(a1_full + 8 * bsel, (c / AB) & 1);
      if (threadIdx.x == 64) TR(c, 6);
      TRW(c, 0);
      tc_fence_after();
      // probe the hidden-tile buffer's two release barriers now, consume the answer after the arithmetic
      bool buf_free = true;
      if (c >= HB) buf_free = mbar_test_wait(h_empty + 8 * hb, ((c - HB) / HB) & 1) & mbar_test_wait(hs_empty + 8 * hb, ((c - HB) / HB) & 1);
      uint32_t packed[CW / 2];
      uint32_t rra[CW / 32][32];
#pragma unroll
      for (int w = 0; w < CW / 32; ++w) tmem_ld_32x32(tmem + lane_addr + bsel * 128 + ch * CW + w * 32, rra[w]);
      tmem_ld_wait();
      if (threadIdx.x == 64) TR(c, 7);
      // acc1[bsel] is in registers: hand the accumulator back (to the leader's MMA thread) before the arithmetic
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(a1_empty_ldr + 8 * bsel);
      if constexpr ((DBG & 4) != 0) {
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) packed[j] = rra[j / 16][(2 * j) % 32];
      } else if (!BWD) {
        uint32_t gate[CW / 32];
#pragma unroll
        for (int h2 = 0; h2 < CW / 32; ++h2) {
          const uint32_t (&rr)[32] = rra[h2];
          const int n0 = cg * FC + ch * CW + h2 * 32;
          const uint64_t idx0 = (uint64_t)m * a.FF + n0;          // multiple of 32: this thread owns one 32-bit mask word
          // pre-activations (dropout scale folded in: relu(acc + b) / (1-p) == relu(acc/(1-p) + b/(1-p)), sB1 holds the
          // pre-scaled bias in training), and the word of their sign bits gathered with one funnel shift per element
          float v[32];
          uint32_t neg = 0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sB1 + n0 + j);
            v[j] = fmaf(__uint_as_float(rr[j]), inv_keep, b4.x);
            v[j + 1] = fmaf(__uint_as_float(rr[j + 1]), inv_keep, b4.y);
            v[j + 2] = fmaf(__uint_as_float(rr[j + 2]), inv_keep, b4.z);
            v[j + 3] = fmaf(__uint_as_float(rr[j + 3]), inv_keep, b4.w);
          }
#pragma unroll
          for (int j = 31; j >= 0; --j) neg = __funnelshift_l(__float_as_uint(v[j]), neg, 1);     // bit j = sign of v[j]
          // gate bit = positive pre-activation that survives the dropout, i.e. the saved activation is non-zero (bf16 has
          // fp32's exponent range, so a positive normal value never rounds to zero; an exactly-zero pre-activation has
          // its bit set and the value 0, which only matters for the measure-zero relu'(0) convention)
          uint32_t keep;
          if constexpr (DROP == 1) {
            keep = drop_bits(a.key_ffn ^ egot2_ep, idx0 >> 5);
          } else if constexpr (DROP == 2) {
            keep = 0;             // general p: a 16-bit field per element, two elements per hash (common.cuh drop_keep)
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t hh = drop_bits(a.key_ffn ^ egot2_ep, (idx0 + j) >> 1);
              keep |= ((hh & 0xffffu) >= thr16 ? 1u : 0u) << j;
              keep |= ((hh >> 16) >= thr16 ? 1u : 0u) << (j + 1);
            }
          } else {
            keep = 0xFFFFFFFFu;
          }
          uint32_t gw = ~neg & keep;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float v0 = (gw >> j) & 1u ? v[j] : 0.f, v1 = (gw >> (j + 1)) & 1u ? v[j + 1] : 0.f;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1);
            const uint32_t u0 = *reinterpret_cast<uint32_t*>(&p0);
            packed[h2 * 16 + j / 2] = u0;
          }
          gate[h2] = gw;
        }
        if (a.hmask && row_ok) {
          uint32_t* hw = reinterpret_cast<uint32_t*>(a.hmask) + gate_idx(cg);
          if constexpr (CW == 64) *reinterpret_cast<uint2*>(hw) = make_uint2(gate[0], gate[CW / 32 - 1]);
          else hw[0] = gate[0];
        }
      } else {
#pragma unroll
        for (int h2 = 0; h2 < CW / 32; ++h2) {
          const uint32_t (&rr)[32] = rra[h2];
          const uint32_t gw = gate_cur[h2];
          float cs[32];          // this row's 32 dhid values; reduced over the warp's 32 rows below
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float v0 = (gw >> j) & 1u ? __uint_as_float(rr[j]) * inv_keep : 0.f;
            const float v1 = (gw >> (j + 1)) & 1u ? __uint_as_float(rr[j + 1]) * inv_keep : 0.f;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1);
            packed[h2 * 16 + j / 2] = *reinterpret_cast<uint32_t*>(&p0);
            cs[j] = v0; cs[j + 1] = v1;
          }
          if (a.db1) {
            // db1 (linear1 bias gradient) = column sums of dhid, taken here so that no other kernel re-reads the (M,FF)
            // tensor: a transposing butterfly over the warp's 32 rows - 31 shuffles for 32 columns - after which lane L
            // holds the sum of column L (rows past M have zero gate bits and contribute nothing).  Costs the kernel ~9 us
            // (HHI b256) against a 17 us separate pass; summing the staged tile in the idle TMA-store warp instead was
            // slower still (it delays the hidden-buffer hand-back).
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
              const bool up = (lane & s) != 0;
#pragma unroll
              for (int i = 0; i < s; ++i) {
                const float send = up ? cs[i] : cs[i + s];
                const float recv = __shfl_xor_sync(0xffffffffu, send, s);
                cs[i] = (up ? cs[i + s] : cs[i]) + recv;
              }
            }
            atomicAdd(sB1 + cg * FC + ch * CW + h2 * 32 + lane, cs[0]);
          }
        }
      }
      // the hid buffer is free once GEMM2(c-HB) retired and the TMA store of chunk c-HB has read it
      if (threadIdx.x == 64) TR(c, 8);
      if (c >= HB && !buf_free) {
        mbar_wait(h_empty + 8 * hb, ((c - HB) / HB) & 1);
        mbar_wait(hs_empty + 8 * hb, ((c - HB) / HB) & 1);
      }
      if (threadIdx.x == 64) TR(c, 9);
#pragma unroll
      for (int j = 0; j < CW / 8; ++j)        // CW/8 x 16 B chunks = this thread's CW columns, 128B-swizzled row
        if (!(DBG & 1) || a.M < 0) sts128(hrow + (uint32_t)(((jb + j) ^ (r & 7)) << 4), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive_cluster(h_full_ldr + 8 * hb); mbar_arrive(hl_full + 8 * hb); }
      if (threadIdx.x == 64) TR(c, 10);
      TRW(c, 1);
    }
    // ------------------------------------------------------------ final: + b2, dropout2, + residual, LayerNorm2
    mbar_wait(a2_full, 0);                    // every GEMM of the pair retired: sHid / sX are free to stage the outputs
    if (threadIdx.x == 64) TR(60, 2);
    for (int c = NC - HB; c < NC; ++c)        // the last hidden tiles have left through TMA
      if (c >= 0) mbar_wait(hs_empty + 8 * (c % HB), (c / HB) & 1);
    tc_fence_after();
    if (threadIdx.x == 64) TR(60, 3);
    // The final epilogue runs ONCE per CTA: fully unrolled over 64 columns it was ~2k straight-line instructions whose
    // cold instruction-cache misses cost ~20k cycles per CTA (stall_no_inst in ncu).  It is therefore written as two
    // compact loops over 8-column groups (8-column TMEM loads, nothing indexed dynamically).
    const int nb = ch * CW;
    const uint32_t xrow = sX + half * HALF + (uint32_t)r * 128;
    const uint32_t yrow = sHid + half * HALF + (uint32_t)r * 128;
    if (split && !a.tile_ctr) {
      // FF-split tail unit: this pair's partial sum over its chunk range joins the others' in fp32; the fix-up kernel
      // applies everything that follows the second GEMM once all S partials are in.  The tile leaves through the (idle)
      // hidden-tile buffers and bulk tensor reductions - [column group of 32][128 rows][128 B], 128B-swizzled, one 32 x 32
      // box per warp - because per-thread red.v4 is LSU-bound at ~1.8 cycles per lane (4096 lane-reds per CTA, ~7k cycles).
      static_assert(HB * TILE >= 128 * 128 * 4, "the fp32 partial tile is staged in the hidden-tile buffers");
#pragma unroll 1
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        uint32_t rr[8];
        tmem_ld_32x8(acc2 + lane_addr + nb + j8 * 8, rr);
        tmem_ld_wait();
        if (!row_ok) {            // rows past M carry relu(b1) . W2, not zero: the scratch must stay clean for the next launch
#pragma unroll
          for (int k = 0; k < 8; ++k) rr[k] = 0u;
        }
        const int col = nb + j8 * 8;                                     // this thread's 8 columns = two 16 B chunks of its row
        const uint32_t srow = sHid + (uint32_t)(col >> 5) * 16384u + (uint32_t)r * 128u;
        const int jc = (col & 31) >> 2;
        sts128(srow + (uint32_t)((jc ^ (r & 7)) << 4), rr[0], rr[1], rr[2], rr[3]);
        sts128(srow + (uint32_t)(((jc + 1) ^ (r & 7)) << 4), rr[4], rr[5], rr[6], rr[7]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int row0 = m0 - a.split_pair0 * 2 * BM + q * 32;
#pragma unroll
        for (int g = 0; g < (CW + 31) / 32; ++g)
          tma_reduce_add_2d(&tm_part, sHid + (uint32_t)((nb >> 5) + g) * 16384u + (uint32_t)(q * 32) * 128u, ((nb >> 5) + g) * 32, row0);
        tma_store_commit();
        tma_store_wait_read();
      }
    } else if (split) {
      float* prow = a.partial + ((size_t)(m - a.split_pair0 * 2 * BM)) * H + nb;
#pragma unroll 1
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        uint32_t rr[8];
        tmem_ld_32x8(acc2 + lane_addr + nb + j8 * 8, rr);
        tmem_ld_wait();
        if (row_ok) {
          red_add_v4(prow + j8 * 8, __uint_as_float(rr[0]), __uint_as_float(rr[1]), __uint_as_float(rr[2]), __uint_as_float(rr[3]));
          red_add_v4(prow + j8 * 8 + 4, __uint_as_float(rr[4]), __uint_as_float(rr[5]), __uint_as_float(rr[6]), __uint_as_float(rr[7]));
        }
      }
      if (a.tile_ctr) {
        // The slice that arrives LAST at this CTA's 128-row tile finishes it right here (what used to be a separate fix-up
        // launch behind the kernel): fence the partial sums, count this slice in, and if all S are in, one warp per row.
        __threadfence();
        named_bar_sync(2, EPW * 32);
        volatile uint32_t* s_flag = reinterpret_cast<volatile uint32_t*>(red);
        int* ctr = a.tile_ctr + (m0 - a.split_pair0 * 2 * BM) / BM;
        if (warp == 2 && lane == 0) s_flag[0] = (uint32_t)atomicAdd(ctr, 1);
        named_bar_sync(1, EPW * 32);
        if ((int)s_flag[0] == a.S - 1) {
          __threadfence();
          float* pbase = a.partial + (size_t)(m0 - a.split_pair0 * 2 * BM) * H;
          for (int rloc = warp - 2; rloc < BM && m0 + rloc < a.M; rloc += EPW) {
            if constexpr (BWD) fixup_row_bwd(lane, m0 + rloc, pbase + (size_t)rloc * H, a.fix_d1, a.fix_out);
            else fixup_row_fwd(lane, m0 + rloc, pbase + (size_t)rloc * H, a.fix_x1, sVec, sVec + 128, sVec + 256, a.eps, a.p_drop,
                               a.key_drop2, egot2_ep, a.fix_y2, a.stat2, a.fix_out);
          }
          if (warp == 2 && lane == 0) *ctr = 0;
        }
      }
    } else if constexpr (BWD) {
      // d3 = acc2 + d1 (gradient through the residual branch); with no dropout2, d1 IS the d2 tile still in sX
      const bf16* d1row = (a.d1 && row_ok) ? a.d1 + (size_t)m * H + nb : nullptr;      // nb is a multiple of 32: 16 B aligned
#pragma unroll 1
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        uint32_t rr[8];
        tmem_ld_32x8(acc2 + lane_addr + nb + j8 * 8, rr);
        const uint32_t sw = (uint32_t)(((jb + j8) ^ (r & 7)) << 4);
        uint32_t w0, w1, w2, w3;
        if (a.d1) {
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (d1row) t = __ldg(reinterpret_cast<const uint4*>(d1row) + j8);
          w0 = t.x; w1 = t.y; w2 = t.z; w3 = t.w;
        } else {
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(xrow + sw));
        }
        tmem_ld_wait();
        const uint32_t w[4] = {w0, w1, w2, w3};
        uint32_t oo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 u = __floats2bfloat162_rn(__uint_as_float(rr[2 * k]) + __uint_as_float(w[k] << 16),
                                                   __uint_as_float(rr[2 * k + 1]) + __uint_as_float(w[k] & 0xffff0000u));
          oo[k] = *reinterpret_cast<uint32_t*>(&u);
        }
        sts128(yrow + sw, oo[0], oo[1], oo[2], oo[3]);
      }
      fence_proxy_async();
      named_bar_sync(2, EPW * 32);
      if (warp == 2 && lane == 0) {
        tma_store_2d(&tm_y2, sHid, 0, m0);
        tma_store_2d(&tm_y2, sHid + HALF, 64, m0);
        tma_store_commit();
        tma_store_wait_read();
      }
    } else {
      // pass 1: y = acc2 + b2 -> dropout2 -> + x1, rounded to bf16 (what LayerNorm2 sees and what is saved); staged in the
      // first hid buffer; row sum / sum of squares on the rounded values
      float sum = 0.f, sq = 0.f;
      static_assert(CW <= 32, "the drop2 keep word covers one aligned run of 32 columns per thread");
      const uint32_t keep2 = DROP == 1 ? drop_word(a.key_drop2 ^ egot2_ep, (uint64_t)m * H + nb) : 0u;
#pragma unroll 1
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        uint32_t rr[8];
        tmem_ld_32x8(acc2 + lane_addr + nb + j8 * 8, rr);
        const uint32_t sw = (uint32_t)(((jb + j8) ^ (r & 7)) << 4);
        uint32_t w0, w1, w2, w3;                  // residual: 8 bf16 of x1 from the swizzled smem tile
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(xrow + sw));
        tmem_ld_wait();
        const uint32_t w[4] = {w0, w1, w2, w3};
        uint32_t yy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = j8 * 8 + 2 * k;
          float v0 = __uint_as_float(rr[2 * k]) + sVec[nb + j], v1 = __uint_as_float(rr[2 * k + 1]) + sVec[nb + j + 1];
          if constexpr (DROP == 1) {          // p == 0.5: this thread's CW columns share one keep word (nb % 32 == 0)
            v0 = (keep2 >> j) & 1u ? v0 * inv_keep : 0.f;
            v1 = (keep2 >> (j + 1)) & 1u ? v1 * inv_keep : 0.f;
          } else if constexpr (DROP == 2) {       // columns j, j + 1 (j even) are the two halves of one hash
            const uint32_t hh = drop_bits(a.key_drop2 ^ egot2_ep, ((uint64_t)m * H + nb + j) >> 1);
            v0 = (hh & 0xffffu) >= thr16 ? v0 * inv_keep : 0.f;
            v1 = (hh >> 16) >= thr16 ? v1 * inv_keep : 0.f;
          }
          v0 += __uint_as_float(w[k] << 16);
          v1 += __uint_as_float(w[k] & 0xffff0000u);
          __nv_bfloat162 t = __floats2bfloat162_rn(v0, v1);
          yy[k] = *reinterpret_cast<uint32_t*>(&t);
          v0 = __uint_as_float(yy[k] << 16);
          v1 = __uint_as_float(yy[k] & 0xffff0000u);
          sum += v0 + v1;
          sq = fmaf(v0, v0, fmaf(v1, v1, sq));
        }
        sts128(yrow + sw, yy[0], yy[1], yy[2], yy[3]);
      }
      red[ch * 128 + r] = sum;
      red[EG * 128 + ch * 128 + r] = sq;
      named_bar_sync(1, EPW * 32);
      float tsum = 0.f, tsq = 0.f;
#pragma unroll
      for (int e = 0; e < EG; ++e) { tsum += red[e * 128 + r]; tsq += red[EG * 128 + e * 128 + r]; }
      const float mean = tsum * (1.f / H);
      const float var = fmaxf(tsq * (1.f / H) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      if (row_ok && ch == 0) { a.stat2[2 * (size_t)m] = mean; a.stat2[2 * (size_t)m + 1] = rstd; }
      // pass 2: x_out = (y - mean) * rstd * g + b, staged in the x1 tile (its residual reads are done)
      const uint32_t orow = xrow;
#pragma unroll 1
      for (int j8 = 0; j8 < CW / 8; ++j8) {
        const uint32_t sw = (uint32_t)(((jb + j8) ^ (r & 7)) << 4);
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(yrow + sw));
        const uint32_t w[4] = {w0, w1, w2, w3};
        uint32_t oo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = nb + j8 * 8 + 2 * k;
          const float o0 = (__uint_as_float(w[k] << 16) - mean) * rstd * sVec[128 + j] + sVec[256 + j];
          const float o1 = (__uint_as_float(w[k] & 0xffff0000u) - mean) * rstd * sVec[128 + j + 1] + sVec[256 + j + 1];
          __nv_bfloat162 u = __floats2bfloat162_rn(o0, o1);
          oo[k] = *reinterpret_cast<uint32_t*>(&u);
        }
        sts128(orow + sw, oo[0], oo[1], oo[2], oo[3]);
      }
      fence_proxy_async();
      named_bar_sync(2, EPW * 32);
      if (warp == 2 && lane == 0) {
        tma_store_2d(&tm_y2, sHid, 0, m0);
        tma_store_2d(&tm_y2, sHid + HALF, 64, m0);
        tma_store_2d(&tm_out, sX, 0, m0);
        tma_store_2d(&tm_out, sX + HALF, 64, m0);
        tma_store_commit();
        tma_store_wait_read();
      }
    }
  }
  if (threadIdx.x == 64) TR(60, 4);
  tc_fence_before();
  __syncthreads();
  if (BWD && a.db1) {            // this CTA's column sums of its chunk range -> global (one reduction per column)
    for (int i = threadIdx.x; i < NC * FC; i += NTHREADS) {
      const float v = sB1[c_begin * FC + i];
      if (v != 0.f) atomicAdd(a.db1 + c_begin * FC + i, v);
    }
  }
  if (threadIdx.x == 64) TR(60, 5);
  cluster_sync_all();                         // the peer may still be signalling this CTA's barriers / using the pair's TMEM
#ifdef EGOT2_FFN_TRACE
  if (threadIdx.x == 0 && a.trace) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[1024 + blockIdx.x * 4 + 1] = (long long)t;
  }
#endif
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_cg2<512>(tmem);
  }
}

// ---------------------------------------------------------------- stand-alone fix-up of the FF-split tail tiles
// (the default; EGOT2_FFN_FIXUP=inkernel lets the last-arriving slice of a tile finish it inside the main kernel instead)
__global__ void __launch_bounds__(256) ffn_fixup_fwd_kernel(int rows, int m_base, float* __restrict__ partial,
                                                            const bf16* __restrict__ x1, const float* __restrict__ b2,
                                                            const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                            float eps, float p_drop, uint64_t key_drop2,
                                                            bf16* __restrict__ y2, float* __restrict__ stat2,
                                                            bf16* __restrict__ x_out) {
  EGOT2_PDL_ENTER();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  fixup_row_fwd(threadIdx.x & 31, m_base + r, partial + (size_t)r * H, x1, b2, ln_g, ln_b, eps, p_drop, key_drop2, egot2_ep, y2,
                stat2, x_out);
}
__global__ void __launch_bounds__(256) ffn_fixup_bwd_kernel(int rows, int m_base, float* __restrict__ partial,
                                                            const bf16* __restrict__ d1, bf16* __restrict__ d3) {
  EGOT2_PDL_ENTER();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  fixup_row_bwd(threadIdx.x & 31, m_base + r, partial + (size_t)r * H, d1, d3);
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

static bool env_off(const char* name) { const char* v = getenv(name); return v && v[0] == '0'; }

size_t ffn_scratch_bytes(int M) {
  // the tail is always shorter than one wave of pairs, whatever the occupancy query says at launch time
  const int pairs = (M + 2 * BM - 1) / (2 * BM), slots = sm_count() / 2;
  const int tail = pairs < slots ? pairs : slots;
  // fp32 partial sums of the tail tiles' rows + one arrival counter per 128-row tile (both zero on entry and on exit)
  return pairs > 1 ? (size_t)tail * 2 * BM * H * sizeof(float) + (size_t)tail * 2 * sizeof(int) : 0;
}

bool ffn_fused_supported(int dtype, int Hdim, int FF) {
  return dtype == EGOT2_BF16 && Hdim == H && FF >= FC && FF % FC == 0;
}

static int ffn_launch(bool bwd, int M, int FF, const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2,
                      const CUtensorMap& thid, const CUtensorMap& ty2, const CUtensorMap& tout, FfnArgs& a, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)TILE * (1 + HB) + (size_t)RING * STAGE + 512 + (2 * EG * 128 + 384 + (size_t)FF) * 4;
  EGOT2_CHECK(smem <= 227 * 1024, "ffn_fused: FF=%d does not fit the bias stage in shared memory", FF);
  const int drop = bwd ? 0 : (a.p_drop <= 0.f ? 0 : (a.p_drop == 0.5f ? 1 : 2));
  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, FfnArgs);
  static const KernelFn kernels[4] = {ffn_sm100_kernel<false, 0>, ffn_sm100_kernel<false, 1>, ffn_sm100_kernel<false, 2>,
                                      ffn_sm100_kernel<true, 0>};
  const int ki = bwd ? 3 : drop;
  static size_t set_for[4] = {0, 0, 0, 0};
  if (set_for[ki] < smem) {
    EGOT2_CUDA(cudaFuncSetAttribute(kernels[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set_for[ki] = smem;
  }
  a.trace = nullptr;
#ifdef EGOT2_FFN_TRACE
  static long long* dtrace = nullptr;
  if (!dtrace) cudaMalloc(&dtrace, (5120 + 2 * 24 * 16 * 2) * 8);
  cudaMemsetAsync(dtrace, 0, (5120 + 2 * 24 * 16 * 2) * 8, st);
  a.trace = dtrace;
#endif
  {
    ProfScope prof(st, bwd ? "ffn_bwd_dx_sm100 M%d H128 FF%d" : "ffn_fwd_sm100 M%d H128 FF%d", M, FF);
    const int tiles = (M + BM - 1) / BM, pairs = (tiles + 1) / 2;         // whole CTA pairs
    // Wave quantisation: `slots` pairs are resident at once (one CTA per SM); the pairs of the last, partial wave would
    // each run a full-length tile on a mostly idle GPU.  Instead each tail tile is cut into S slices of the FF loop
    // (S * tail <= slots, S | FF/128) that run side by side and meet in an fp32 scratch; a small fix-up kernel applies
    // what follows the second GEMM (ffn_fixup_*).  HHI b256: 90 pairs on 74 slots -> 74 full + 16 x 4 quarter units.
    static int slots_cached[4] = {0, 0, 0, 0};
    if (slots_cached[ki] == 0) {
      int n = 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sm_count()); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kernels[ki], &cfg) != cudaSuccess || n <= 0) { n = sm_count() / 2; (void)cudaGetLastError(); }
      slots_cached[ki] = n;
    }
    int slots = slots_cached[ki];
    if (getenv("EGOT2_FFN_SLOTS") && atoi(getenv("EGOT2_FFN_SLOTS")) > 0) slots = atoi(getenv("EGOT2_FFN_SLOTS"));   // tests: force a tail at small M
    const int NCH = FF / FC;
    int S = 1, tail = pairs % slots;
    if (a.partial && pairs > slots && tail > 0 && !env_off("EGOT2_FFN_SPLIT")) {
      for (int cand = 2; cand <= NCH && cand * tail <= slots; cand *= 2)
        if (NCH % cand == 0) S = cand;
    }
    a.S = S;
    a.split_pair0 = S > 1 ? pairs - tail : pairs;
    // the counters sit behind the partial rows of the largest tail this scratch was sized for
    // EGOT2_FFN_FIXUP=inkernel: the last-arriving slice of a tail tile finishes it inside the main kernel (no fix-up launch).
    // Measured on B200 (HHI b256): 62.1 / 64.2 us (forward / backward-dx launchers) against 54.1 / 58.6 with the separate
    // fix-up kernels - the last CTA walks its 128 rows with 16 warps at the very end of the kernel, the separate launch
    // spreads them over 512 CTAs - so the separate launches stay the default.
    static const bool separate = !(getenv("EGOT2_FFN_FIXUP") && strcmp(getenv("EGOT2_FFN_FIXUP"), "inkernel") == 0);
    {
      const int slots_max = sm_count() / 2, tail_max = pairs < slots_max ? pairs : slots_max;
      a.tile_ctr = (S > 1 && !separate) ? reinterpret_cast<int*>(a.partial + (size_t)tail_max * 2 * BM * H) : nullptr;
    }
    const int grid_pairs = S > 1 ? (pairs - tail) + tail * S : pairs;
    // fp32 map over the tail's partial rows (the target of the split units' bulk reductions); unused when S == 1
    CUtensorMap tpart = tx;
    if (S > 1) {
      const int tail_rows = tail * 2 * BM;
      EncodeTiledFn fn = ffn_encode_fn();
      EGOT2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
      cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)tail_rows};
      cuuint64_t strides[1] = {(cuuint64_t)H * 4};
      cuuint32_t box[2] = {32, 32};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = fn(&tpart, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.partial, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      EGOT2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(partial) failed (%d)", (int)r);
    }
    launch(kernels[ki], dim3(2 * grid_pairs), dim3(NTHREADS), smem, st, tx, tw1, tw2, thid, ty2, tout, tpart, a);
    EGOT2_LAUNCH_CHECK();
    if (S > 1 && separate) {
      const int m_base = a.split_pair0 * 2 * BM, rows = M - m_base;
      if (bwd)
        launch(ffn_fixup_bwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, rows, m_base, a.partial, a.fix_d1, a.fix_out);
      else
        launch(ffn_fixup_fwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, rows, m_base, a.partial, a.fix_x1, a.b2, a.ln_g, a.ln_b,
               a.eps, a.p_drop, a.key_drop2, a.fix_y2, a.stat2, a.fix_out);
      EGOT2_LAUNCH_CHECK();
    }
  }
#ifdef EGOT2_FFN_TRACE
  {
    static int printed = 0;
    static const int trace_at = getenv("EGOT2_FFN_TRACE_AT") ? atoi(getenv("EGOT2_FFN_TRACE_AT")) : 3;
    if (printed++ == trace_at) {
      static long long h[5120 + 2 * 24 * 16 * 2];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, dtrace, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = h[0];
      for (int i = 0; i < 64 * 16; ++i) if (h[i] > 0 && h[i] < t0) t0 = h[i];
      printf("ffn trace (cycles since first stamp): c | prod w1e w2e | mma w1f a1e w2f hf | epi a1f ldtm comp hwait hfull | st hf rd\n");
      printf("ffn stamps: entry %lld setup %lld a2_full %lld hs_done %lld | resid+sum %lld bar1 %lld staged %lld bar2 %lld tma_issued %lld | epi_done %lld cta_sync %lld\n", h[60*16+0]-t0, h[60*16+1]-t0, h[60*16+2]-t0, h[60*16+3]-t0, h[60*16+6]-t0, h[60*16+7]-t0, h[60*16+8]-t0, h[60*16+9]-t0, h[60*16+10]-t0, h[60*16+4]-t0, h[60*16+5]-t0);
      for (int c = 0; c < FF / FC && c < 64; ++c) {
        printf("%2d |", c);
        for (int k = 0; k < 13; ++k) printf(" %7lld%s", h[c * 16 + k] ? h[c * 16 + k] - t0 : -1LL, (k == 1 || k == 5 || k == 10) ? " |" : "");
        printf("\n");
      }
      // per-warp view: ns since the earliest stamp; rows = (cta, warp), columns = chunks, "seen/handed"
      long long w0 = 0;
      for (int i = 5120; i < 5120 + 2 * 24 * 16 * 2; ++i) if (h[i] > 0 && (w0 == 0 || h[i] < w0)) w0 = h[i];
      for (int cta = 0; cta < 2; ++cta)
        for (int w = 2; w < 2 + EPW; ++w) {
          printf("cta%d w%02d q%d g%d |", cta, w, w & 3, (w - 2) >> 2);
          for (int c = 0; c < 16 && c < FF / FC; ++c) {
            const long long* e = h + 5120 + ((cta * 24 + w) * 16 + c) * 2;
            printf(" %5lld/%5lld", e[0] ? e[0] - w0 : -1LL, e[1] ? e[1] - w0 : -1LL);
          }
          printf("\n");
        }
    }
  }
#endif
  return 0;
}

// x_out = LN2(x1 + drop2(drop(relu(x1 W1^T + b1)) W2^T + b2)); also writes hid (if non-null), its non-zero gate bits
// hmask [FF/64][M] x 64 bit (if non-null), y2 and stat2
int ffn_fused_fwd(int M, int FF, const void* x1, const void* W1, const float* b1, const void* W2, const float* b2,
                  const float* ln_g, const float* ln_b, float eps, void* hid, void* hmask, void* y2, float* stat2, void* x_out,
                  float p_drop, uint64_t key_ffn, uint64_t key_drop2, float* scratch, cudaStream_t st) {
  if (M == 0) return 0;
  CUtensorMap tx, tw1, tw2, thid, ty2, tout;
  EGOT2_TRY(kmajor_map(&tx, x1, H, M, H, 128));
  EGOT2_TRY(kmajor_map(&tw1, W1, H, FF, H, 64));      // each CTA of the pair fetches 64 rows of a stage
  EGOT2_TRY(kmajor_map(&tw2, W2, FF, H, FF, 64));
  if (getenv("EGOT2_FFN_NOHID")) hid = nullptr;                 // EXPERIMENT: how much of the kernel is the hidden-tensor write?
  const bool tiled = getenv("EGOT2_HID_TILED") != nullptr;      // EXPERIMENT: tile-major hidden layout (consumers not adapted)
  const int tiles128 = (M + 127) / 128;
  if (tiled && hid) EGOT2_TRY(kmajor_map(&thid, hid, 64, tiles128 * (FF / 64) * 128, 64, 128));
  else EGOT2_TRY(kmajor_map(&thid, hid ? hid : y2, FF, M, FF, 128));    // never dereferenced when hid == nullptr
  EGOT2_TRY(kmajor_map(&ty2, y2, H, M, H, 128));
  EGOT2_TRY(kmajor_map(&tout, x_out, H, M, H, 128));
  FfnArgs a;
  a.M = M; a.FF = FF; a.b1 = b1; a.b2 = b2; a.ln_g = ln_g; a.ln_b = ln_b; a.eps = eps;
  a.stat2 = stat2; a.save_hid = hid != nullptr; a.hmask = (uint2*)hmask; a.d1 = nullptr; a.hid_tiled = tiled && hid;
  a.db1 = nullptr;
  a.p_drop = p_drop; a.key_ffn = key_ffn; a.key_drop2 = key_drop2;
  a.partial = scratch; a.fix_x1 = (const bf16*)x1; a.fix_d1 = nullptr; a.fix_y2 = (bf16*)y2; a.fix_out = (bf16*)x_out;
  return ffn_launch(false, M, FF, tx, tw1, tw2, thid, ty2, tout, a, st);
}

// Data-gradient half of the block's backward (see the kernel comment):
//   dhid = (d2 . W2) * gate / (1-p)  -> (M,FF) bf16;   d3 = dhid . W1 + d1  -> (M,128) bf16   (d1 == nullptr: d1 is d2)
int ffn_fused_bwd_dx(int M, int FF, const void* d2, const void* d1, const void* hmask, const void* W1, const void* W2,
                     float p_drop, void* dhid, void* d3, float* scratch, float* db1, cudaStream_t st) {
  if (M == 0) return 0;
  CUtensorMap tx, tw1, tw2, thid, ty2;
  EGOT2_TRY(kmajor_map(&tx, d2, H, M, H, 128));
  EGOT2_TRY(kmajor_map(&tw1, W1, H, FF, H, 128));     // MN-major stages: 128 k-rows x 64 columns per CTA
  EGOT2_TRY(kmajor_map(&tw2, W2, FF, H, FF, 128));
  EGOT2_TRY(kmajor_map(&thid, dhid, FF, M, FF, 128));
  EGOT2_TRY(kmajor_map(&ty2, d3, H, M, H, 128));
  FfnArgs a;
  a.M = M; a.FF = FF; a.b1 = nullptr; a.b2 = nullptr; a.ln_g = nullptr; a.ln_b = nullptr; a.eps = 0.f;
  a.stat2 = nullptr; a.save_hid = 1; a.hmask = (uint2*)const_cast<void*>(hmask); a.d1 = (const bf16*)d1; a.hid_tiled = 0;
  a.db1 = db1;
  a.p_drop = p_drop; a.key_ffn = 0; a.key_drop2 = 0;
  a.partial = scratch; a.fix_x1 = nullptr; a.fix_d1 = (const bf16*)(d1 ? d1 : d2); a.fix_y2 = nullptr; a.fix_out = (bf16*)d3;
  return ffn_launch(true, M, FF, tx, tw1, tw2, thid, ty2, ty2, a, st);
}

}  // namespace egot2
