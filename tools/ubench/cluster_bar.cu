// Microbenchmark: what does one hand-off across a CTA pair cost?  (the fused FFN's GEMM1 -> epilogue -> GEMM2 chain crosses
// the cluster twice per ring traversal; profiles/r02_ffn_what_bounds_it.txt)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_bar tools/ubench/cluster_bar.cu
// mode 0: ping-pong between two warps of ONE CTA through two local mbarriers
// mode 1: ping-pong between one thread of each CTA of a pair: remote mbarrier.arrive one way, remote arrive the other way
// mode 2: leader -> peer by tcgen05.commit (cta_group::2, multicast to both CTAs), peer -> leader by remote arrive
// mode 3: fan-in: W warps of EACH CTA arrive on a leader barrier (count 2W), the leader answers with a multicast commit
//         that all of them wait on (the shape of one FFN chunk hand-off); reports the period per round
#include <cstdio>
#include <cuda_runtime.h>
#include "../../egot2_b200/csrc/sm100.cuh"
using namespace egot2::sm100;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) k(int mode, int W, int iters, long long* out) {
  __shared__ __align__(8) unsigned long long bars[4];
  __shared__ uint32_t slot;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(b0, mode == 3 ? 2 * W : 1);
    mbar_init(b1, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_cg2<32>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (mode == 0) {
    if (rank == 0 && lane == 0 && warp == 1) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) { mbar_arrive(b0); mbar_wait(b1, i & 1); }
      t1 = clock64();
      out[0] = t1 - t0;
    } else if (rank == 0 && lane == 0 && warp == 2) {
      for (int i = 0; i < iters; ++i) { mbar_wait(b0, i & 1); mbar_arrive(b1); }
    }
  } else if (mode == 1) {
    // leader arrives on the PEER's b0; the peer waits its b0 and arrives on the LEADER's b1
    if (lane == 0 && warp == 1) {
      if (rank == 0) {
        const uint32_t peer_b0 = mapa(b0, 1);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { mbar_arrive_cluster(peer_b0); mbar_wait_cluster(b1, i & 1); }
        t1 = clock64();
        out[0] = t1 - t0;
      } else {
        const uint32_t ldr_b1 = mapa(b1, 0);
        for (int i = 0; i < iters; ++i) { mbar_wait_cluster(b0, i & 1); mbar_arrive_cluster(ldr_b1); }
      }
    }
  } else if (mode == 2) {
    if (lane == 0 && warp == 1) {
      if (rank == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { umma_commit_cg2(b0, 3); mbar_wait_cluster(b1, i & 1); }
        t1 = clock64();
        out[0] = t1 - t0;
      } else {
        const uint32_t ldr_b1 = mapa(b1, 0);
        for (int i = 0; i < iters; ++i) { mbar_wait(b0, i & 1); mbar_arrive_cluster(ldr_b1); }
      }
    }
  } else {
    // b0 (leader): fan-in barrier, count 2W.  b1 (both CTAs): "go", signalled by the leader's multicast commit.
    const uint32_t ldr_b0 = mapa(b0, 0);
    if (warp == 1 && lane == 0 && rank == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) { mbar_wait_cluster(b0, i & 1); umma_commit_cg2(b1, 3); }
      t1 = clock64();
      out[0] = t1 - t0;
    } else if (warp >= 2 && warp < 2 + W) {
      for (int i = 0; i < iters; ++i) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ldr_b0);
        mbar_wait(b1, i & 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_cg2<32>(slot);
}

int main() {
  long long* out; cudaMalloc(&out, 64);
  const int iters = 2000;
  struct { int mode, W; const char* what; } runs[] = {
      {0, 0, "local ping-pong (arrive + wait, two warps of one CTA)"},
      {1, 0, "pair ping-pong (remote arrive both ways)"},
      {2, 0, "pair ping-pong (multicast tcgen05.commit one way, remote arrive back)"},
      {3, 1, "fan-in  1 warp  per CTA + multicast commit back"},
      {3, 4, "fan-in  4 warps per CTA + multicast commit back"},
      {3, 8, "fan-in  8 warps per CTA + multicast commit back"},
      {3, 16, "fan-in 16 warps per CTA + multicast commit back"}};
  for (auto& r : runs) {
    cudaMemset(out, 0, 64);
    k<<<2, 640>>>(r.mode, r.W, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-72s %8.1f cycles per round trip [%s]\n", r.what, (double)h / iters, cudaGetErrorString(e));
  }
  return 0;
}
