// Microbenchmark: tcgen05.mma issue/throughput floor on B200 for the tile shapes our kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o egot2_b200/build/umma_rate tools/ubench/umma_rate.cu -lcuda
// Each CTA (1 per SM) issues `iters` MMAs (M=128, N, K=16, bf16, SS mode, SW128 K-major or MN-major operands) back to
// back on resident smem tiles and reports clock64 cycles per MMA.  Optional: 4 extra warps doing tcgen05.ld in a loop
// (epilogue pressure) and/or 4 warps doing STS.128 (smem write pressure).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../egot2_b200/csrc/sm100.cuh"
using namespace egot2::sm100;

template <int N, bool MN>
__global__ void __launch_bounds__(320, 1) k(int iters, int ld_pressure, int sts_pressure, int nslots, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 4 * 16384;          // 4 A slots (128x64 bf16), then B slots (N x 64)
  const uint32_t bar = sB + 4 * N * 128, slot = bar + 8;
  volatile uint32_t* slot_ptr = (volatile uint32_t*)(smem_raw + (slot - smem_u32(smem_raw)));
  volatile int* stop = (volatile int*)(smem_raw + (slot + 8 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); *stop = 0; }
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MN, MN);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = (i >> 2) % nslots, kk = i & 3;
      const uint64_t ad = MN ? make_smem_desc_sw128(sA + s * 16384 + kk * 2048, 8192, 1024)
                             : make_smem_desc_sw128(sA + s * 16384 + kk * 32, 16, 1024);
      const uint64_t bd = MN ? make_smem_desc_sw128(sB + s * N * 128 + kk * 2048, 8192, 1024)
                             : make_smem_desc_sw128(sB + s * N * 128 + kk * 32, 16, 1024);
      umma_bf16(tmem, ad, bd, idesc, i > 0);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    *stop = 1;
    out[blockIdx.x] = t1 - t0;
  } else if (warp >= 2 && warp < 6 && ld_pressure) {
    const int q = warp & 3;
    uint32_t acc = 0;
    while (!*stop) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + 256, r);
      tmem_ld_wait();
      acc += r[0] + r[31];
    }
    if (acc == 0x12345) out[1000] = acc;
  } else if (warp >= 6 && sts_pressure) {
    uint32_t dst = base + 200 * 1024 + (threadIdx.x - 192) * 16;
    while (!*stop) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(lane) : "memory");
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int N, bool MN> void run(const char* name, int grid, int ldp, int stsp, int nslots) {
  long long* out; cudaMalloc(&out, 2048 * 8); cudaMemset(out, 0, 2048 * 8);
  const int iters = 2048;
  size_t smem = 1024 + 4 * 16384 + 4 * N * 128 + 64; if (stsp) smem = 220 * 1024;
  cudaFuncSetAttribute(k<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<N, MN><<<grid, 320, smem>>>(iters, ldp, stsp, nslots, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60; for (int i = 0; i < grid && i < 148; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  printf("%-28s grid=%3d ld=%d sts=%d slots=%d: %.1f..%.1f cyc/MMA (floor %d)  [%s]\n", name, grid, ldp, stsp, nslots,
         (double)mn / iters, (double)mx / iters, 128 * N / 256, cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, false>("M128 N64  K-major", grid, 0, 0, 4);
    run<128, false>("M128 N128 K-major", grid, 0, 0, 4);
    run<256, false>("M128 N256 K-major", grid, 0, 0, 4);
    run<128, false>("M128 N128 K-major 1slot", grid, 0, 0, 1);
    run<128, true>("M128 N128 MN-major", grid, 0, 0, 4);
    run<256, true>("M128 N256 MN-major", grid, 0, 0, 4);
    run<128, false>("M128 N128 K-major +ldtm", grid, 1, 0, 4);
    run<128, false>("M128 N128 K-major +sts", grid, 0, 1, 4);
    run<128, false>("M128 N128 K-major +both", grid, 1, 1, 4);
    run<256, false>("M128 N256 K-major +both", grid, 1, 1, 4);
  }
  return 0;
}
