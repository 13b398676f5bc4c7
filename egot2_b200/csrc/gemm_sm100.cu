// gemm_sm100.cu — bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem -> tcgen05.mma -> TMEM
// -> tcgen05.ld -> fused epilogue (bias / ReLU / ReLU-gate / dropout / residual / accumulate) -> HBM.
//
//   C[M,N] = epilogue( op(A)[M,K] . op(B)[K,N] ),  A,B bf16, accumulation fp32 in TMEM, C bf16 or fp32.
//
// One CTA owns one 128 x BN output tile (UMMA M=128, N=BN, K=16 per instruction, cta_group::1) and walks its
// K range in 64-wide blocks through a 4-stage TMA/mbarrier ring.  Warp roles (192 threads):
//   warp 0  TMA producer (one elected lane)       warp 1  TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5  epilogue: each owns the TMEM lane quadrant (warp_id % 4), one accumulator row per thread
// Operand orientation is handled by the descriptors, not by copies:
//   K-major  (row-major, K contiguous; activations (M,K) / nn.Linear weights (N,K)): one TMA box {64 k, rows}
//   MN-major (K rows, M|N contiguous; used by dX = dY.W and dW = dY^T.X):            boxes of {64 mn, 64 k}
// gridDim.z > 1 = split-K for the weight-gradient GEMMs (K = every token of the batch): fp32 atomics into C.
#include <stdio.h>
#include <stdlib.h>

#define EGOT2_FILE_ID 4
#include "ops.h"
#include "sm100.cuh"

namespace egot2 {

using namespace sm100;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements = one 128-byte swizzle row
// warp 0 TMA producer, warp 1 MMA issuer, warps 2.. epilogue: EPG column groups x 4 TMEM lane quadrants.  EPG = 2 (8 epilogue
// warps, each thread one accumulator row x half of the tile's columns) was tried because most GEMMs of the translators are one
// or two tiles per CTA, i.e. the serial chain load -> MMA -> epilogue -> store IS the kernel - and MEASURED SLOWER on every
// workload (B200, us per step, EPG 2 vs 1: HHI 372 vs 343, PNR 1071 vs 1003, LTA 1730 vs 1628): 320 threads x 2 CTAs per SM cap
// the kernel at 96 registers and the epilogue spills.  So 4 epilogue warps it stays.
#ifndef EGOT2_GEMM_EPG
#define EGOT2_GEMM_EPG 1
#endif
constexpr int EPG = EGOT2_GEMM_EPG;
constexpr int NTHREADS = 64 + 128 * EPG;

struct EpiArgs {
  int M, N;
  void* C; int ldc; int c_rpg, c_gstride;
  const float* bias; int relu;
  const void* mask; int ldm; float mask_scale;
  float p_drop; uint64_t drop_key; int drop_bit_mode;
  const void* residual; int ldr;
  int accumulate; int atomic;
  long long split_stride;        // > 0: split z writes its own slab at C + z * split_stride (no atomics)
  int tma_store;     // bf16 C, plain rows: the tile leaves through shared memory + TMA (coalesced) instead of per-row stores
  // grouped-K addressing for gathered MN-major operands (3-D tensor maps): 64-row k-block = kg groups x kdpad rows
  int k_grouped, kg, kdblocks;
  // LayerNorm fused behind the epilogue (BN = N = 128, bf16 C): ln_out[row] = LN(C[row]) * g + b (+ table[row % table_rows])
  // (* dropout); the row statistics go to ln_stat.  Row indices of stat / table / dropout are (storage row of C) + ln_row0.
  const float* ln_g; const float* ln_b; float ln_eps;
  void* ln_out; float* ln_stat; const float* ln_table; int ln_table_rows; int ln_row0;
  float ln_p_drop; uint64_t ln_drop_key;
  int trans_c;       // tma_red only: the tile is added to the transpose of C (C is (N, M) row-major, ld = ldc)
  int tma_red;       // split-K fp32 accumulation: the tile is staged in shared memory and added to C by bulk tensor reductions (tma_c = fp32 map)
  int trace_slot;    // -DEGOT2_GEMM_TRACE builds: CTA (0,0,0) stamps its phases into g_gtrace[trace_slot]
};

// -DEGOT2_GEMM_TRACE: where one CTA of a launch spends its time (clock64 stamps of CTA (0,0,0); egot2_gemm_trace_dump prints them)
#ifdef EGOT2_GEMM_TRACE
__device__ long long g_gtrace[512][16];
#define GTR(k) do { if (blockIdx.x == 0 && blockIdx.z == 0 && e.trace_slot >= 0) g_gtrace[e.trace_slot][k] = clock64(); } while (0)
#else
#define GTR(k) do { } while (0)
#endif

__device__ __forceinline__ long long remap(int r, int rpg, int gstride) {
  return rpg > 0 ? (long long)(r / rpg) * gstride + (r % rpg) : (long long)r;
}


// one 128-bit reduction instead of four scalar atomics (split-K partial sums)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 32 contiguous elements <-> 32 fp32 registers with 128-bit accesses (pointer must be 16 B aligned)
__device__ __forceinline__ void load32(const float* p, float (&o)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + j);
    o[j] = t.x; o[j + 1] = t.y; o[j + 2] = t.z; o[j + 3] = t.w;
  }
}
__device__ __forceinline__ void load32(const bf16* p, float (&o)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + j);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[j + 2 * k] = __uint_as_float(w[k] << 16);
      o[j + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ void store32(float* p, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
__device__ __forceinline__ void store32(bf16* p, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(p + j) = o;
  }
}

template <int BN, typename TO> struct TileCfg {   // stages chosen so that BN<=128 tiles fit two CTAs per SM
  // bf16 output tile staged in shared memory for the TMA-store epilogue (bf16 C, BN <= 128 only: 128 rows x BN x 2 B);
  // fp32 outputs (split-K weight gradients: long K loops) spend that shared memory on a deeper operand ring instead
  static constexpr uint32_t OUT_BYTES = (BN <= 128 && sizeof(TO) == 2) ? 128 * BN * 2 : 0;
  static constexpr int STAGES = BN == 128 ? (OUT_BYTES ? 2 : 3) : (BN == 64 ? (OUT_BYTES ? 3 : 4) : 4);
  static constexpr int MIN_CTAS = BN == 256 ? 1 : 2;
};
__device__ __forceinline__ void sts128(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Persistent over output tiles: CTA b takes tiles b, b + gridDim.x, ... (n fastest, so neighbouring CTAs share A rows in
// L2).  The smem ring keeps running across tiles and the accumulator is double-buffered in TMEM (2 x BN columns), so the
// TMA fill and the MMAs of tile i+1 overlap the epilogue of tile i - for the short-K GEMMs of this model (K = 128..384)
// the per-tile fill + epilogue latency is several times the MMA time.  Split-K launches (gridDim.z > 1) give every
// CTA exactly one tile.
template <int BN, bool A_MN, bool B_MN, typename TO>
__global__ void __launch_bounds__(NTHREADS, TileCfg<BN, TO>::MIN_CTAS)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const EpiArgs e, const int k_blocks_total, const int k_blocks_per_split, const int tiles_n, const int num_tiles) {
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t B_BYTES = BN * BK * 2;
  constexpr int STAGES = TileCfg<BN, TO>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024 B alignment
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + STAGES * A_BYTES;
  const uint32_t sOut = sB + STAGES * B_BYTES;                            // [BN/64][128 rows][128 B], 128B-swizzled (TMA-store epilogue)
  const uint32_t bars = sOut + TileCfg<BN, TO>::OUT_BYTES;                          // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tmem_full = bars + 16 * STAGES, tmem_empty = tmem_full + 16;
  const uint32_t tmem_slot = tmem_empty + 16;
  const uint32_t bias_off = (tmem_slot + 8 + 15u) & ~15u;        // float sbias[2][BN]: the tile's bias columns (broadcast reads in the epilogue)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_begin = blockIdx.z * k_blocks_per_split;
  const int kb_end = min(k_blocks_total, kb_begin + k_blocks_per_split);
  const int nkb = kb_end - kb_begin;

  pdl_launch_dependents();      // the next kernel's prologue may overlap this one's main loop
  if (threadIdx.x == 0) GTR(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (e.tma_store || e.tma_red) tma_prefetch_desc(&tma_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + 8 * s, 1); mbar_init(tmem_empty + 8 * s, 4 * EPG); }
    fence_barrier_init();
  }
  // split-K CTAs own exactly one tile: a single accumulator (BN columns) - the double buffer would only keep other
  // tensor-core kernels' CTAs (the data-gradient chain beside these weight-gradient GEMMs) from allocating TMEM on this SM
  const bool one_tile = gridDim.z > 1;
  if (warp == 1) { if (one_tile) tmem_alloc<BN>(tmem_slot); else tmem_alloc<2 * BN>(tmem_slot); }
  if (threadIdx.x == 0) GTR(1);
  pdl_wait();                   // everything above touched only kernel parameters, shared memory and TMEM
  if (threadIdx.x == 0) GTR(2);
  EGOT2_TL(EGOT2_FILE_ID);
  const unsigned long long egot2_ep = epoch_xor();
  float* sbias = reinterpret_cast<float*>(smem_raw + (bias_off - smem_u32(smem_raw)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) GTR(3);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int it = 0;                                                   // ring position, continuous across tiles
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
          const int kb = kb_begin + i;
          const int k = kb * BK;
          if (A_MN && B_MN && e.k_grouped) {
            // token rows live in (clip, frame) groups inside a (B,T,H) tensor: one k-block = kg clips x kdpad frames,
            // frames beyond the segment are zero-filled by TMA (they contribute nothing to dW)
            const int grp = (kb / e.kdblocks) * e.kg, d0 = (kb % e.kdblocks) * 64;
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_3d(sA + s * A_BYTES + j * 8192, &tma_a, full0 + 8 * s, m0 + 64 * j, d0, grp);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(sB + s * B_BYTES + j * 8192, &tma_b, full0 + 8 * s, n0 + 64 * j, d0, grp);
            continue;
          }
          if (!A_MN) {
            tma_load_2d(sA + s * A_BYTES, &tma_a, full0 + 8 * s, k, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + s * A_BYTES + j * 8192, &tma_a, full0 + 8 * s, m0 + 64 * j, k);
          }
          if (!B_MN) {
            tma_load_2d(sB + s * B_BYTES, &tma_b, full0 + 8 * s, k, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + s * B_BYTES + j * 8192, &tma_b, full0 + 8 * s, n0 + 64 * j, k);
          }
          if (tile == (int)blockIdx.x && (i == 0 || i == nkb - 1)) GTR(i == 0 ? 4 : 5);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(tmem_empty + 8 * buf, ((lt >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator (tile lt-2)
        tc_fence_after();
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          if (lt == 0 && (i == 0 || i == nkb - 1)) GTR(i == 0 ? 6 : 7);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // K-major: 16 k = 32 B inside the 128 B swizzle row; MN-major: 16 k = 16 rows of 128 B
            const uint64_t ad = A_MN ? make_smem_desc_sw128(sA + s * A_BYTES + kk * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sA + s * A_BYTES + kk * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc_sw128(sB + s * B_BYTES + kk * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sB + s * B_BYTES + kk * 32, 16, 1024);
            umma_bf16(tmem_base + buf * BN, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * s);          // smem slot reusable once these MMAs retire
        }
        umma_commit(tmem_full + 8 * buf);       // accumulator complete
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5)
    // One accumulator row per thread.  Fast path (full 32-column chunk, 16 B-aligned rows): every global access is
    // a 128-bit vector, the one large auxiliary operand (ReLU gate or residual) is fetched while the TMEM load is
    // in flight, and there is no per-element control flow.  Edge chunks take the scalar path.
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    const int c_lo = ((warp - 2) >> 2) * (BN / EPG), c_hi = c_lo + BN / EPG;      // this warp's columns of every tile
    const float inv_keep = e.p_drop > 0.f ? 1.f / (1.f - e.p_drop) : 1.f;
    const bool first_split = blockIdx.z == 0;
    const float* bias = first_split ? e.bias : nullptr;
    // warp-uniform: plain fp32 split-K accumulation (what every weight-gradient GEMM is) leaves by TMA reduction (below)
    const bool tma_red = sizeof(TO) == 4 && e.tma_red && gridDim.z > 1 && nkb > 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int m = m0 + q * 32 + lane;
      float* sb = sbias + buf * BN;
      if (bias) {
        for (int i = threadIdx.x - 64; i < BN; i += NTHREADS - 64) sb[i] = n0 + i < e.N ? __ldg(bias + n0 + i) : 0.f;
        named_bar_sync(1, NTHREADS - 64);
      }
      const bool row_ok = (m < e.M) && nkb > 0;
      const long long crow = row_ok ? remap(m, e.c_rpg, e.c_gstride) : 0;
      TO* __restrict__ crow_ptr = (TO*)e.C + (long long)blockIdx.z * e.split_stride + crow * e.ldc;
      const TO* mask_row = e.mask ? (const TO*)e.mask + (long long)(row_ok ? m : 0) * e.ldm : nullptr;
      const TO* res_row = (e.residual && first_split) ? (const TO*)e.residual + (long long)(row_ok ? m : 0) * e.ldr : nullptr;
      const bool vec_ok = aligned16(crow_ptr) && aligned16(mask_row) && aligned16(res_row);
      const TO* aux_row = mask_row ? mask_row : res_row;
      // the first chunk's auxiliary operand is requested BEFORE waiting for the accumulator, every later chunk's one
      // chunk ahead: its global-load latency hides behind the MMAs / the previous chunk instead of in front of each chunk
      float aux[32], aux_next[32];
      const bool aux_vec = row_ok && vec_ok && aux_row != nullptr;
      if (aux_vec && n0 + c_lo + 32 <= e.N) load32(aux_row + n0 + c_lo, aux_next);
      if (e.tma_store && lt > 0) {              // the previous tile's TMA store has finished reading the staged tile
        if (warp == 2 && lane == 0) tma_store_wait_read();
        named_bar_sync(3, NTHREADS - 64);
      }
      mbar_wait(tmem_full + 8 * buf, (lt >> 1) & 1);
      if (lt == 0 && threadIdx.x == 64) GTR(8);
      tc_fence_after();
      float ln_sum = 0.f, ln_sq = 0.f;
      if (tma_red) {
        // Split-K partial sums leave through shared memory + bulk tensor reductions instead of a red.v4 per thread and 4
        // columns: the LSU pays ~1.8 cycles per lane and red instruction, which made the epilogue of the weight-gradient
        // GEMMs (128 x 256 fp32 per CTA = 8192 lane-reds, ~15k cycles) as long as their K loop.  The operand ring is idle
        // by now (one tile per CTA, all of its MMAs have retired) and takes the fp32 tile, [column group of 32][128 rows]
        // [128 B], 128B-swizzled like every other TMA tile.  Every warp sends its own 32 rows x 32 columns (4 KB box) as
        // soon as they are staged - no CTA-wide barrier - and the TMEM load of the next group is in flight meanwhile.
        const int rr = q * 32 + lane;
        uint32_t ra[32], rb[32];
        auto stage = [&](const uint32_t (&r)[32], int c0) {
          const uint32_t wbox = sA + (uint32_t)(c0 >> 5) * 16384u + (uint32_t)(q * 32) * 128u;      // this warp's 4 KB box
          if (e.trans_c) {
            // C is stored transposed ((N, M) row-major): the box is [32 columns n][32 rows m], one 128-byte smem row per
            // column of the accumulator; lane = m writes one float of every row (32 lanes = one conflict-free 128 B row)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(wbox + (uint32_t)j * 128u + (uint32_t)((((lane >> 2) ^ (j & 7)) << 4) | ((lane & 3) << 2))), "r"(r[j]) : "memory");
          } else {
            const uint32_t srow = wbox + (uint32_t)lane * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128(srow + (uint32_t)((j ^ (rr & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && n0 + c0 < e.N) {
            if (e.trans_c) tma_reduce_add_2d(&tma_c, wbox, m0 + q * 32, n0 + c0);
            else tma_reduce_add_2d(&tma_c, wbox, n0 + c0, m0 + q * 32);
          }
        };
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
        tmem_ld_32x32(tcol + c_lo, ra);
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 64) {
          tmem_ld_wait();
          if (c0 + 32 < c_hi) tmem_ld_32x32(tcol + c0 + 32, rb);
          stage(ra, c0);
          if (c0 + 32 < c_hi) {
            tmem_ld_wait();
            if (c0 + 64 < c_hi) tmem_ld_32x32(tcol + c0 + 64, ra);
            stage(rb, c0 + 32);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(tmem_empty + 8 * buf); tma_store_commit(); tma_store_wait_read(); }    // shared memory may be released
        if (lt == 0 && threadIdx.x == 64) GTR(9);
        continue;
      }
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
        const int nb = n0 + c0;
        const bool active = row_ok && nb < e.N;
        const bool fast = active && vec_ok && (nb + 32 <= e.N);
#pragma unroll
        for (int j = 0; j < 32; ++j) aux[j] = aux_next[j];
        if (aux_vec && c0 + 32 < c_hi && nb + 64 <= e.N) load32(aux_row + nb + 32, aux_next);
        tmem_ld_wait();
        if (c0 + 32 >= c_hi) {                  // this warp's part of the accumulator is in registers: release the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + 8 * buf);
        }
        if (!active) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (fast) {
          if (bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + j);
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (mask_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = aux[j] > 0.f ? v[j] * e.mask_scale : 0.f;
          }
          if (e.p_drop > 0.f) {
            const uint64_t idx0 = (uint64_t)m * e.N + nb;
            if (e.p_drop == 0.5f && (idx0 & 31) == 0) {     // one hash for the chunk's 32 columns (common.cuh drop_keep)
              const uint32_t kw = drop_word(e.drop_key ^ egot2_ep, idx0);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (kw >> j) & 1u ? v[j] * inv_keep : 0.f;
            } else if (e.p_drop != 0.5f && (idx0 & 1) == 0) {      // general p: two columns per hash
              float dm[32];
              drop_scale_n<32>(e.drop_key ^ egot2_ep, idx0, e.p_drop, inv_keep, dm);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= dm[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= drop_scale(e.drop_key ^ egot2_ep, idx0 + j, e.p_drop, inv_keep);
            }
          }
          if (res_row) {
            if (mask_row) load32(res_row + nb, aux);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += aux[j];
          }
          if (e.accumulate) {
            if constexpr (sizeof(TO) == 4) {
              float* dst = (float*)crow_ptr + nb;
              if (e.atomic) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                float old[32];
                load32(dst, old);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += old[j];
                store32(dst, v);
              }
            }
          } else if (BN <= 128 && sizeof(TO) == 2 && (e.tma_store || e.ln_g)) {
            // this thread's 32 columns = 4 x 16 B chunks of its 128 B swizzle row in the staged tile (with a fused
            // LayerNorm the staged row is also where its second pass reads the rounded values back from)
            const int r = q * 32 + lane;
            const uint32_t srow = sOut + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)r * 128u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 t = __floats2bfloat162_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&t);
                if (e.ln_g) {          // row statistics on the ROUNDED values: what a separate LayerNorm kernel would read
                  const float y0 = __uint_as_float(w[k] << 16), y1 = __uint_as_float(w[k] & 0xffff0000u);
                  ln_sum += y0 + y1;
                  ln_sq = fmaf(y0, y0, fmaf(y1, y1, ln_sq));
                }
              }
              sts128(srow + (uint32_t)(((((c0 & 63) >> 3) + j) ^ (r & 7)) << 4), w[0], w[1], w[2], w[3]);
            }
            if (!e.tma_store) store32(crow_ptr + nb, v);      // scattered rows (embedding segments): C goes straight to global
          } else {
            store32(crow_ptr + nb, v);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (n >= e.N) continue;
            float x = v[j];
            if (bias) x += sb[c0 + j];
            if (e.relu) x = fmaxf(x, 0.f);
            if (mask_row) x = to_f32(mask_row[n]) > 0.f ? x * e.mask_scale : 0.f;
            if (e.p_drop > 0.f) x *= drop_scale(e.drop_key ^ egot2_ep, (uint64_t)m * e.N + n, e.p_drop, inv_keep, e.drop_bit_mode != 0);
            if (res_row) x += to_f32(res_row[n]);
            if (e.accumulate) {
              if constexpr (sizeof(TO) == 4) {
                if (e.atomic) atomicAdd((float*)crow_ptr + n, x);
                else ((float*)crow_ptr)[n] += x;
              }
            } else {
              crow_ptr[n] = from_f32<TO>(x);
            }
          }
        }
      }
      if constexpr (BN == 128 && sizeof(TO) == 2) {
        if (e.ln_g && row_ok) {
          // fused LayerNorm, second pass: this thread's row is in the staged tile (it wrote it itself: no barrier)
          const float mean = ln_sum * (1.f / 128.f);
          const float rstd = rsqrtf(fmaxf(ln_sq * (1.f / 128.f) - mean * mean, 0.f) + e.ln_eps);
          const long long lrow = crow + e.ln_row0;
          if (e.ln_stat) { e.ln_stat[2 * lrow] = mean; e.ln_stat[2 * lrow + 1] = rstd; }
          bf16* orow = (bf16*)e.ln_out + crow * e.ldc;
          const float* trow = e.ln_table ? e.ln_table + (lrow % e.ln_table_rows) * 128 : nullptr;
          const float ln_keep = e.ln_p_drop > 0.f ? 1.f / (1.f - e.ln_p_drop) : 1.f;
          const int r = q * 32 + lane;
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            const uint32_t srow = sOut + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)r * 128u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                           : "r"(srow + (uint32_t)(((((c0 & 63) >> 3) + j) ^ (r & 7)) << 4)));
              const uint32_t w[4] = {w0, w1, w2, w3};
              const int c = c0 + 8 * j;
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.ln_g + c)), g1 = __ldg(reinterpret_cast<const float4*>(e.ln_g + c + 4));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.ln_b + c)), b1 = __ldg(reinterpret_cast<const float4*>(e.ln_b + c + 4));
              const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float o[8];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                o[2 * k] = (__uint_as_float(w[k] << 16) - mean) * rstd * gg[2 * k] + bb[2 * k];
                o[2 * k + 1] = (__uint_as_float(w[k] & 0xffff0000u) - mean) * rstd * gg[2 * k + 1] + bb[2 * k + 1];
              }
              if (trow) {
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(trow + c)), t1 = __ldg(reinterpret_cast<const float4*>(trow + c + 4));
                o[0] += t0.x; o[1] += t0.y; o[2] += t0.z; o[3] += t0.w; o[4] += t1.x; o[5] += t1.y; o[6] += t1.z; o[7] += t1.w;
              }
              if (e.ln_p_drop > 0.f) {
                { float dm[8]; drop_scale_n<8>(e.ln_drop_key ^ egot2_ep, (uint64_t)lrow * 128 + c, e.ln_p_drop, ln_keep, dm);
#pragma unroll
                  for (int k = 0; k < 8; ++k) o[k] *= dm[k]; }
              }
              uint4 pk;
              __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
              pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
              pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(orow + c) = pk;
            }
          }
        }
      }
      if (lt == 0 && threadIdx.x == 64) GTR(9);
      if (BN <= 128 && e.tma_store) {
        fence_proxy_async();                      // the generic-proxy smem writes become visible to the TMA engine
        named_bar_sync(2, NTHREADS - 64);
        if (lt == 0 && threadIdx.x == 64) GTR(10);
        if (warp == 2 && lane == 0) {
#pragma unroll
          for (int h = 0; h < BN / 64; ++h) tma_store_2d(&tma_c, sOut + h * 16384, n0 + h * 64, m0);
          tma_store_commit();
        }
      }
    }
    if (e.tma_store && warp == 2 && lane == 0) tma_store_wait_all();
    if (threadIdx.x == 64) GTR(11);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) GTR(12);
  if (warp == 1) {
    __syncwarp();
    if (one_tile) tmem_dealloc<BN>(tmem_base); else tmem_dealloc<2 * BN>(tmem_base);
    if (lane == 0) GTR(13);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 2-D bf16 tensor map over a row-major (rows, inner) matrix with leading dimension ld (elements).
int make_map(CUtensorMap* map, const void* base, int inner, int rows, int ld, int box_inner, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  EGOT2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGOT2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%d rows=%d ld=%d", (int)r, inner, rows, ld);
  return 0;
}

// 2-D fp32 map over a row-major (rows, inner) matrix (the target of the split-K bulk reductions)
int make_map_f32(CUtensorMap* map, const void* base, int inner, int rows, int ld, int box_inner, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  EGOT2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGOT2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed (%d) inner=%d rows=%d ld=%d", (int)r, inner, rows, ld);
  return 0;
}

// 3-D bf16 map over rows that live in groups: element (c, d, g) at base + ((g * gstride_rows + d) * ld + c)
int make_map3(CUtensorMap* map, const void* base, int inner, int rpg, int groups, int ld, long long gstride_rows,
              int box_inner, int box_d, int box_g) {
  EncodeTiledFn fn = encode_fn();
  EGOT2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rpg, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)gstride_rows * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_d, (cuuint32_t)box_g};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGOT2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d) inner=%d rpg=%d groups=%d ld=%d", (int)r, inner,
              rpg, groups, ld);
  return 0;
}

struct KGroup { int on = 0, g = 1, dblocks = 1, kb_total = 0; };
#ifdef EGOT2_GEMM_TRACE
struct GTraceTag { char txt[96]; };
GTraceTag g_gtrace_tags[512];
int g_gtrace_n = 0;
#endif
static bool host_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN, bool A_MN, bool B_MN, typename TO>
int launch(const GemmArgs& a, const CUtensorMap& ma, const CUtensorMap& mb, const KGroup& kg, cudaStream_t st) {
  constexpr int STAGES = TileCfg<BN, TO>::STAGES;
  constexpr size_t smem = 1024 + STAGES * (BM * BK * 2 + BN * BK * 2) + TileCfg<BN, TO>::OUT_BYTES + 16 * STAGES + 96 + 2 * BN * 4;
  static bool attr_set = false;
  auto kern = gemm_sm100_kernel<BN, A_MN, B_MN, TO>;
  if (!attr_set) {
    EGOT2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  EpiArgs e;
  e.M = a.M; e.N = a.N; e.C = a.C; e.ldc = a.ldc; e.c_rpg = a.c_rpg; e.c_gstride = a.c_gstride;
  e.bias = a.bias; e.relu = a.relu; e.mask = a.mask; e.ldm = a.ldm; e.mask_scale = a.mask_scale;
  e.p_drop = a.p_drop; e.drop_key = a.drop_key; e.drop_bit_mode = a.drop_bit_mode; e.residual = a.residual; e.ldr = a.ldr;
  e.accumulate = a.accumulate; e.atomic = a.split_k > 1;
  e.ln_g = a.ln_g; e.ln_b = a.ln_b; e.ln_eps = a.ln_eps; e.ln_out = a.ln_out; e.ln_stat = a.ln_stat; e.ln_table = a.ln_table;
  e.ln_table_rows = a.ln_table_rows > 0 ? a.ln_table_rows : 1; e.ln_row0 = a.ln_row0; e.ln_p_drop = a.ln_p_drop; e.ln_drop_key = a.ln_drop_key;
  e.k_grouped = kg.on; e.kg = kg.g; e.kdblocks = kg.dblocks;
  // TMA-store epilogue: bf16 output in plain row order, whole 32-column chunks, every per-row operand 16 B aligned (the
  // vector path then never falls back to direct stores that would race with the tile store)
  CUtensorMap mc = ma;
  e.tma_store = 0;
  if (BN <= 128 && sizeof(TO) == 2 && !a.accumulate && a.c_rpg == 0 && a.N % 32 == 0 && a.ldc % 8 == 0 && host_al16(a.C) &&
      (!a.mask || (host_al16(a.mask) && a.ldm % 8 == 0)) && (!a.residual || (host_al16(a.residual) && a.ldr % 8 == 0)) &&
      !(getenv("EGOT2_GEMM_TMASTORE") && getenv("EGOT2_GEMM_TMASTORE")[0] == '0')) {
    EGOT2_TRY(make_map(&mc, a.C, a.N, a.M, a.ldc, 64, 128));
    e.tma_store = 1;
  }
  const int kb_total = kg.on ? kg.kb_total : (a.K + BK - 1) / BK;
  int splits = a.split_k < 1 ? 1 : a.split_k;
  if (splits > kb_total) splits = kb_total;
  {
    // split-K grids are ONE wave: tiles x splits <= resident CTAs (19 splits x 16 tiles = 304 CTAs on 296 slots ran the
    // weight-gradient GEMMs in two waves, the second one 3 % full)
    const int nt = ((a.N + BN - 1) / BN) * ((a.M + BM - 1) / BM), slots = sm_count() * TileCfg<BN, TO>::MIN_CTAS;
    if (splits > 1 && nt * splits > slots) splits = slots / nt > 0 ? slots / nt : 1;
  }
  const int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;
  e.atomic = splits > 1 && a.split_stride == 0;
  e.split_stride = a.split_stride;
  // split-K accumulation by bulk tensor reduction (EGOT2_GEMM_TMARED=0: per-thread red.v4 instead)
  static const bool tma_red_on = !(getenv("EGOT2_GEMM_TMARED") && getenv("EGOT2_GEMM_TMARED")[0] == '0');
  e.tma_red = 0;
  if (sizeof(TO) == 4 && tma_red_on && e.atomic && a.accumulate && !a.bias && !a.residual && !a.relu && !a.mask && a.p_drop <= 0.f &&
      a.c_rpg == 0 && a.N % 32 == 0 && a.ldc % 4 == 0 && host_al16(a.C) && (size_t)(BN / 32) * 16384 <= (size_t)STAGES * (BM * BK * 2 + BN * BK * 2)) {
    if (a.trans_c) EGOT2_TRY(make_map_f32(&mc, a.C, a.M, a.N, a.ldc, 32, 32));
    else EGOT2_TRY(make_map_f32(&mc, a.C, a.N, a.M, a.ldc, 32, 32));
    e.tma_red = 1;
  }
  e.trans_c = a.trans_c;
  if (a.trans_c && !e.tma_red) return -3;      // only the bulk-reduction epilogue can transpose: the caller uses the plain orientation
  if (a.splits_out) *a.splits_out = splits;
  const int tiles_n = (a.N + BN - 1) / BN, num_tiles = tiles_n * ((a.M + BM - 1) / BM);
  // persistent CTAs: as many as are resident at once, evened out so that every CTA walks the same number of tiles
  int ctas = num_tiles;
  if (splits == 1) {
    const int slots = sm_count() * TileCfg<BN, TO>::MIN_CTAS;
    const int rounds = (num_tiles + slots - 1) / slots;
    ctas = (num_tiles + rounds - 1) / rounds;
  }
  dim3 grid(ctas, 1, splits);
  e.trace_slot = -1;
#ifdef EGOT2_GEMM_TRACE
  if (g_gtrace_n < 512) {
    snprintf(g_gtrace_tags[g_gtrace_n].txt, 96, "bn%d %s%s %s M%d N%d K%d grid(%d,%d) tma_store%d%s%s", BN, A_MN ? "mn" : "k", B_MN ? "mn" : "k",
             sizeof(TO) == 4 ? "f32" : "bf16", a.M, a.N, a.K, ctas, splits, e.tma_store, a.residual ? " +res" : "", a.p_drop > 0.f ? " +drop" : "");
    e.trace_slot = g_gtrace_n++;
  }
#endif
  ProfScope prof(st, "gemm_sm100<bn%d,%s%s,%s> M%d N%d K%d sk%d%s", BN, A_MN ? "mn" : "k", B_MN ? "mn" : "k",
                 sizeof(TO) == 4 ? "f32" : "bf16", a.M, a.N, a.K, splits, a.mask ? " +mask" : (a.residual ? " +res" : ""));
  ::egot2::launch(kern, grid, dim3(NTHREADS), smem, st, ma, mb, mc, e, kb_total, kb_per, tiles_n, num_tiles);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

template <bool A_MN, bool B_MN, typename TO>
int pick_bn(const GemmArgs& a, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const KGroup& kg, cudaStream_t st) {
  switch (bn) {
    case 64: return launch<64, A_MN, B_MN, TO>(a, ma, mb, kg, st);
    case 128: return launch<128, A_MN, B_MN, TO>(a, ma, mb, kg, st);
    default: return launch<256, A_MN, B_MN, TO>(a, ma, mb, kg, st);
  }
}

template <typename TO>
int pick_major(const GemmArgs& a, int bn, const CUtensorMap& ma, const CUtensorMap& mb, const KGroup& kg, cudaStream_t st) {
  if (!a.trans_a && a.trans_b) return pick_bn<false, false, TO>(a, bn, ma, mb, kg, st);
  if (!a.trans_a && !a.trans_b) return pick_bn<false, true, TO>(a, bn, ma, mb, kg, st);
  if (a.trans_a && !a.trans_b) return pick_bn<true, true, TO>(a, bn, ma, mb, kg, st);
  return pick_bn<true, false, TO>(a, bn, ma, mb, kg, st);
}

bool host_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Can the LayerNorm behind this GEMM run inside its epilogue?  One 128-column tile holds whole rows (N = 128), bf16 in and out,
// plain (non-accumulating, single-split) output with vector-aligned rows, no ReLU gate operand.
bool gemm_sm100_ln_ok(const GemmArgs& a) {
  // Opt-in (EGOT2_GEMM_LN=1): measured on B200 the fused tail costs ~1 % of the step (352.0 vs 347.6 us HHI, 1033 vs 1022 us PNR,
  // profiles/r02_ln_epilogue.txt) - the 60-CTA GEMM grid normalises its rows more slowly than the 960-CTA LayerNorm kernel it replaces.
  if (!(getenv("EGOT2_GEMM_LN") && getenv("EGOT2_GEMM_LN")[0] == '1')) return false;
  return a.in_dtype == EGOT2_BF16 && a.out_dtype == EGOT2_BF16 && a.N == 128 && a.K >= 16 && !a.accumulate && a.split_k <= 1 &&
         a.split_stride == 0 && !a.mask && !a.trans_a && a.a_rpg == 0 && a.b_rpg == 0 && a.ldc == 128 && host_al16(a.C) &&
         host_al16(a.A) && host_al16(a.B) && a.lda % 8 == 0 && a.ldb % 8 == 0 && (!a.residual || (host_al16(a.residual) && a.ldr % 8 == 0)) &&
         (a.c_rpg == 0 || (a.c_rpg * 128 * 2) % 16 == 0);
}

#ifdef EGOT2_GEMM_TRACE
// prints, for every traced launch since the last dump, the cycles CTA (0,0,0) spent up to each phase (relative to its entry)
extern "C" int egot2_gemm_trace_dump() {
  static long long h[512][16];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_gtrace, sizeof(h));
  printf("# cycles since CTA entry: prologue | pdl_wait done | sync | tma first,last issued | mma saw first,last stage | epi saw acc | epi done | bar | stored | cta sync | dealloc\n");
  for (int i = 0; i < g_gtrace_n; ++i) {
    const long long t0 = h[i][0];
    printf("%-64s", g_gtrace_tags[i].txt);
    for (int k = 1; k < 14; ++k) printf(" %6lld", h[i][k] ? h[i][k] - t0 : -1LL);
    printf("\n");
  }
  g_gtrace_n = 0;
  fflush(stdout);
  return 0;
}
#endif

// returns -1 when this kernel does not take the problem (caller falls back to the CUDA-core GEMM)
int gemm_sm100(const GemmArgs& a, cudaStream_t st) {
  if (a.in_dtype != EGOT2_BF16) return -1;
  if (a.ln_g && !gemm_sm100_ln_ok(a)) return -2;            // the caller asked for a fusion this kernel cannot do: a bug upstream
  KGroup kg;
  if (a.a_rpg > 0 || a.b_rpg > 0) {
    // gathered token rows are supported for the weight-gradient orientation (both operands MN-major, K = tokens)
    if (!(a.trans_a && !a.trans_b)) return -1;
    const int rpg = a.a_rpg > 0 ? a.a_rpg : a.b_rpg;
    if ((a.a_rpg > 0 && a.b_rpg > 0 && a.a_rpg != a.b_rpg) || a.K % rpg) return -1;
    kg.on = 1;
    // rows per group padded to the next power of two (a divisor of the 64-row k-block); the padding rows are zero-filled by
    // TMA.  (A floor of 16 here made the 2-token segments of the LTA translators walk 8x the k-blocks they needed.)
    int dpad = 1;
    while (dpad < rpg && dpad < 64) dpad <<= 1;
    kg.g = 64 / dpad;
    kg.dblocks = (rpg + 63) / 64;
    const int groups = a.K / rpg;
    kg.kb_total = ((groups + kg.g - 1) / kg.g) * kg.dblocks;
  }
  if (a.M < 1 || a.N < 32 || a.K < 16) return -1;                  // degenerate tiles (tiny heads) stay on CUDA cores
  if (!host_aligned16(a.A) || !host_aligned16(a.B) || (a.lda % 8) || (a.ldb % 8)) return -1;
  if (a.accumulate && a.out_dtype != EGOT2_F32) return -1;
  if (a.trans_c && (!a.accumulate || a.out_dtype != EGOT2_F32 || a.split_k < 2 || a.split_stride > 0 || a.c_rpg > 0)) return -3;
  if (a.split_k > 1 && a.split_stride == 0 && (!a.accumulate || a.relu || a.mask || a.p_drop > 0.f)) return -1;
  if (a.split_stride > 0 && (a.accumulate || a.out_dtype != EGOT2_F32 || a.relu || a.mask || a.p_drop > 0.f || a.residual)) return -1;
  // short-K problems are epilogue/store bound: 128-wide tiles run two CTAs per SM; long-K ones amortise A over 256 columns
  int bn = a.N <= 64 ? 64 : ((a.N <= 128 || a.K <= 512) ? 128 : ((a.N % 256 == 0 || a.N > 512) ? 256 : 128));
  {
    // few output tiles (the EgoT2-g branches: 16 .. 1800 token rows; LTA: 4096 rows x 512 columns): a long K loop on a
    // handful of SMs is bound by what ONE SM can pull from L2, so narrower tiles on more SMs win although they re-read A
    static const bool small_on = !(getenv("EGOT2_GEMM_BN_SMALL") && getenv("EGOT2_GEMM_BN_SMALL")[0] == '0');
    const long long mt = (a.M + BM - 1) / BM;
    if (small_on && !(a.accumulate && a.split_k > 1) && a.K >= 1024)
      while (bn > 64 && mt * ((a.N + bn - 1) / bn) < sm_count() / 4) bn >>= 1;      // (at 64 tiles of 256 columns LTA is faster as is)
  }
  if (a.accumulate && a.split_k > 1) {
    // split-K weight gradients with a small output: narrower tiles until the grid covers the SMs
    const long long mt = (a.M + BM - 1) / BM;
    while (bn > 64 && mt * ((a.N + bn - 1) / bn) * a.split_k < sm_count()) bn >>= 1;
  }
  CUtensorMap ma, mb;
  // A: K-major (M,K) -> box {64 k, 128 m};  MN-major (K,M) -> box {64 m, 64 k}
  if (kg.on) {
    const int rpg = a.a_rpg > 0 ? a.a_rpg : a.b_rpg, groups = a.K / rpg, dpad = 64 / kg.g;
    EGOT2_TRY(make_map3(&ma, a.A, a.M, rpg, groups, a.lda, a.a_rpg > 0 ? a.a_gstride : rpg, 64, dpad, kg.g));
    EGOT2_TRY(make_map3(&mb, a.B, a.N, rpg, groups, a.ldb, a.b_rpg > 0 ? a.b_gstride : rpg, 64, dpad, kg.g));
    if (a.out_dtype == EGOT2_F32) return pick_major<float>(a, bn, ma, mb, kg, st);
    return pick_major<bf16>(a, bn, ma, mb, kg, st);
  }
  if (!a.trans_a) EGOT2_TRY(make_map(&ma, a.A, a.K, a.M, a.lda, BK, BM));
  else EGOT2_TRY(make_map(&ma, a.A, a.M, a.K, a.lda, 64, BK));
  // B: K-major (N,K) -> box {64 k, BN n};   MN-major (K,N) -> box {64 n, 64 k}
  if (a.trans_b) EGOT2_TRY(make_map(&mb, a.B, a.K, a.N, a.ldb, BK, bn));
  else EGOT2_TRY(make_map(&mb, a.B, a.N, a.K, a.ldb, 64, BK));
  if (a.out_dtype == EGOT2_F32) return pick_major<float>(a, bn, ma, mb, kg, st);
  return pick_major<bf16>(a, bn, ma, mb, kg, st);
}

}  // namespace egot2
