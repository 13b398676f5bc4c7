// common.cuh — shared device/host helpers for libegot2 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egot2.h"

namespace egot2 {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define EGOT2_CHECK(cond, ...)                   \
  do {                                           \
    if (!(cond)) {                               \
      ::egot2::set_error(__VA_ARGS__);           \
      return 1;                                  \
    }                                            \
  } while (0)
#define EGOT2_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::egot2::set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)
// every kernel launch of the library passes through here exactly once (also feeds egot2_launch_count())
extern unsigned long long g_launch_count;
#define EGOT2_LAUNCH_CHECK()              \
  do {                                    \
    ++::egot2::g_launch_count;            \
    EGOT2_CUDA(cudaGetLastError());       \
  } while (0)
#define EGOT2_TRY(expr)      \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != 0) return rc__; \
  } while (0)

int sm_count();

// ---------------------------------------------------------------- per-launcher CUDA-event timing (egot2_prof_*)
// When profiling is enabled (egot2_prof_enable(1); eager launches only, never during graph capture) every launcher
// brackets the kernels it enqueues with a pair of CUDA events recorded on the launch stream.
bool prof_enabled();
struct ProfScope {
  cudaStream_t st; int slot;
  ProfScope(cudaStream_t stream, const char* fmt, ...);
  ~ProfScope();
};

static inline size_t dtype_size(int dtype) { return dtype == EGOT2_BF16 ? 2 : 4; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- scalar conversions
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- counter-based dropout RNG
// mask(seed, site, idx): stateless, so forward and backward regenerate identical masks.
// site ids (combined with layer index by the callers)
enum DropSite : uint32_t {
  SITE_FEAT = 1, SITE_EMBED = 2, SITE_ATTN = 3, SITE_DROP1 = 4, SITE_FFN = 5, SITE_DROP2 = 6, SITE_HEAD = 7
};
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint64_t site_key(uint64_t seed, uint32_t site, uint32_t layer) {
  return mix64(seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)(site + 16u * layer + 1u)));
}
// uniform in [0,1) with 24 bits
__device__ __forceinline__ float uniform01(uint64_t key, uint64_t idx) {
  uint64_t h = mix64(key + idx * 0xD6E8FEB86659FD93ULL);
  return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}
// returns the multiplier: 0 if dropped, 1/(1-p) if kept
__device__ __forceinline__ float drop_scale(uint64_t key, uint64_t idx, float p, float inv_keep) {
  return uniform01(key, idx) >= p ? inv_keep : 0.0f;
}

}  // namespace egot2
