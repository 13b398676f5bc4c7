// Microbenchmark: what does the MMA-issuing thread pay for tcgen05.commit / fences between groups of MMAs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o egot2_b200/build/umma_issue tools/ubench/umma_issue.cu -lcuda
#include <cstdio>
#include <cuda_runtime.h>
#include "../../egot2_b200/csrc/sm100.cuh"
using namespace egot2::sm100;

// mode bit0: commit after every group of G MMAs (to a private barrier nobody waits on)
// mode bit1: tcgen05.fence::after_thread_sync before every group
// mode bit2: commit is the multicast form (cluster of 1 -> mask 1)
// mode bit3: mbarrier try_wait (already-completed barrier) before every group
template <int G>
__global__ void __launch_bounds__(128, 1) k(int groups, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar = sB + 16384, dummy = bar + 8, done = bar + 16, slot = bar + 24;
  volatile uint32_t* slot_ptr = (volatile uint32_t*)(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(dummy, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && lane == 0) {
    mbar_arrive(done);                 // phase 0 of `done` completes: waits on parity 0 pass immediately
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, false, false);
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (mode & 8) mbar_wait(done, 0);
      if (mode & 2) tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < G; ++kk)
        umma_bf16(tmem, make_smem_desc_sw128(sA + (kk & 3) * 32, 16, 1024), make_smem_desc_sw128(sB + (kk & 3) * 32, 16, 1024), idesc, 1);
      if (mode & 1) { if (mode & 4) umma_commit_mc(dummy, 1); else umma_commit(dummy); }
    }
    long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  const int groups = 256;
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 16; ++mode) { if (mode & 4) continue;
    for (int G : {4, 8}) {
      if (G == 4) k<4><<<1, 128, 64 * 1024>>>(groups, mode, out); else k<8><<<1, 128, 64 * 1024>>>(groups, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("G=%d commit=%d fence=%d mc=%d wait=%d: issue %.1f cyc/MMA, total %.1f cyc/MMA [%s]\n", G, mode & 1, (mode >> 1) & 1,
             (mode >> 2) & 1, (mode >> 3) & 1, (double)h[0] / (groups * G), (double)h[1] / (groups * G), cudaGetErrorString(e));
    }
  }
  return 0;
}
