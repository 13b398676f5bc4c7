// common.cuh — shared device/host helpers for libegot2 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <limits.h>
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

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the library starts with pdl_launch_dependents() - the NEXT kernel on the stream may become resident
// and run its prologue (barrier init, TMEM alloc, tensormap prefetch) while this one is still running - and executes
// pdl_wait() before its first global-memory access: griddepcontrol.wait returns only when every earlier grid has
// completed and flushed, so data hazards are exactly those of plain stream order.  Because the chain is only
// transitive when EVERY kernel waits, launch() (which sets the stream-serialization attribute) must only be used with
// kernels that call pdl_wait(); nothing before the wait may read memory another kernel of the step writes.
// EGOT2_PDL=0 launches without the attribute (griddepcontrol.* are then no-ops).
// ---- in-graph timeline (EGOT2_TIMELINE builds only; tools/timeline.py).  Block 0 / thread 0 of every kernel stamps
// %globaltimer right after its griddepcontrol.wait, i.e. when everything before it on the stream has finished; consecutive
// stamps of the chain therefore give each kernel's in-graph duration including launch gaps.  Each translation unit owns a
// copy of the buffer pointer (no relocatable device code); tl_register collects their setters.
#ifdef EGOT2_TIMELINE
void tl_register(void (*setter)(unsigned long long*));
static __device__ unsigned long long* tl_buf_dev = nullptr;
namespace {
struct TlReg {
  TlReg() { tl_register([](unsigned long long* p) { cudaMemcpyToSymbol(tl_buf_dev, &p, sizeof(p)); }); }
};
static TlReg tl_reg_instance;
}  // namespace
__device__ __forceinline__ void tl_stamp(unsigned loc) {
  if (tl_buf_dev && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned i = atomicAdd(reinterpret_cast<unsigned*>(tl_buf_dev), 1u);
    if (i < 4000) { tl_buf_dev[1 + 2 * i] = t; tl_buf_dev[2 + 2 * i] = loc; }
  }
}
#define EGOT2_TL(id) ::egot2::tl_stamp((unsigned)(id) * 100000u + (unsigned)__LINE__)
#else
#define EGOT2_TL(id) do { } while (0)
#endif
#ifndef EGOT2_FILE_ID
#define EGOT2_FILE_ID 0
#endif
// ---- dropout epoch: a device-resident counter folded into every dropout key at execution time, so that a CUDA graph that
// is replayed step after step draws a fresh mask each time although its kernels' key arguments are frozen at capture.
// effective key = key ^ (epoch * golden ratio); the epoch lives in a library-owned 8-byte slot (egot2_dropout_epoch_*):
// advanced by a one-thread kernel (graph-capturable), or, for tests, emulated on the host side (the launchers then fold
// the same term into the keys they pass).  Off (pointer null) unless enabled: keys are then exactly the host-computed ones.
// Each translation unit holds its own copy of the slot pointer (no relocatable device code), set through epoch_register.
void epoch_register(void (*setter)(const unsigned long long*));
static __device__ const unsigned long long* epoch_slot_dev = nullptr;
namespace {
struct EpochReg {
  EpochReg() { epoch_register([](const unsigned long long* p) { cudaMemcpyToSymbol(epoch_slot_dev, &p, sizeof(p)); }); }
};
static EpochReg epoch_reg_instance;
}  // namespace
constexpr unsigned long long kEpochMul = 0x9E3779B97F4A7C15ULL;
__device__ __forceinline__ unsigned long long epoch_xor() {
  const unsigned long long* p = epoch_slot_dev;
  return p ? *p * kEpochMul : 0ull;
}
extern unsigned long long g_host_epoch;      // host-side emulation (tests): folded into site_key() on the host
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// first statement(s) of every kernel; also defines egot2_ep, the dropout-epoch term every key use is XORed with
#define EGOT2_PDL_ENTER()                                                                    \
  ::egot2::pdl_launch_dependents(); ::egot2::pdl_wait(); EGOT2_TL(EGOT2_FILE_ID);            \
  const unsigned long long egot2_ep = ::egot2::epoch_xor(); (void)egot2_ep
bool pdl_enabled();
// Launch priority: kernels on the caller's stream (the data-gradient / forward chain, i.e. the critical path) outrank the
// library's side-stream kernels (parameter gradients), so when both have CTAs waiting for an SM the chain goes first and
// the side work fills what is left.  Returns INT_MIN when priorities are disabled (EGOT2_PRIO=0) or unavailable.
int launch_priority(cudaStream_t st);
template <typename... Params, typename... Args>
inline void launch(void (*kern)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  const int prio = launch_priority(st);
  if (prio != INT_MIN) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = prio;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);      // errors surface through EGOT2_LAUNCH_CHECK()
}

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
  SITE_FEAT = 1, SITE_EMBED = 2, SITE_ATTN = 3, SITE_DROP1 = 4, SITE_FFN = 5, SITE_DROP2 = 6, SITE_HEAD = 7,
  SITE_DEC_SELF = 8, SITE_DEC_CROSS = 9, SITE_DEC_DROP1 = 10, SITE_DEC_DROP2 = 11, SITE_DEC_DROP3 = 12, SITE_DEC_FFN = 13,
  SITE_PROMPT = 14
};
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint64_t site_key(uint64_t seed, uint32_t site, uint32_t layer) {
  const uint64_t k = mix64(seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)(site + 16u * layer + 1u)));
#ifndef __CUDA_ARCH__
  return k ^ (g_host_epoch * kEpochMul);     // host-side epoch emulation (0 unless a test sets it)
#else
  return k;
#endif
}
// 32 random bits for element `idx` of the dropout site `key`: a two-multiply xorshift hash (lowbias32-style) of the
// 32-bit element index, keyed by both halves of the 64-bit site key.  ~7 integer instructions per element.
__host__ __device__ __forceinline__ uint32_t drop_bits(uint64_t key, uint64_t idx) {
  uint32_t x = (uint32_t)idx ^ (uint32_t)key ^ ((uint32_t)(idx >> 32) * 0x9E3779B1u);
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= (x >> 15) ^ (uint32_t)(key >> 32);
  x *= 0x846ca68bu;
  return x;
}
// keep threshold: an element is DROPPED iff drop_bits < p * 2^32
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const float t = p * 4294967296.0f;
  return t >= 4294967040.0f ? 0xFFFFFF00u : (uint32_t)(t + 0.5f);
}
// Bernoulli(1-p) keep decision for element idx.
//   p == 0.5 (the HHI / HOI encoder default): ONE random bit per element - bit idx%32 of the hash of idx/32 - at EVERY
//   elementwise dropout site, so a thread that owns an aligned run of elements pays one hash per run (drop_scale_n below)
//   instead of one per element: the per-element hash was the larger half of the out-projection epilogue and of the
//   LayerNorm-backward / FFN final epilogues in training.
//   other p: a 16-bit field per element - half idx%2 of the hash of idx/2 - compared with round(p * 2^16): one hash per TWO
//   elements (p is honoured to 2^-16: 0.1 becomes 0.100006).
// (thr and bit_mode are kept for source compatibility; neither rule depends on them any more.)
__host__ __device__ __forceinline__ uint32_t drop_threshold16(float p) {
  const float t = p * 65536.0f;
  return t >= 65535.0f ? 65535u : (uint32_t)(t + 0.5f);
}
__host__ __device__ __forceinline__ uint32_t drop_field16(uint32_t h, uint32_t odd) { return odd ? (h >> 16) : (h & 0xffffu); }
__host__ __device__ __forceinline__ bool drop_keep(uint64_t key, uint64_t idx, float p, uint32_t thr, bool bit_mode = false) {
  (void)bit_mode; (void)thr;
  if (p == 0.5f) return (drop_bits(key, idx >> 5) >> (uint32_t)(idx & 31)) & 1u;
  return drop_field16(drop_bits(key, idx >> 1), (uint32_t)(idx & 1)) >= drop_threshold16(p);
}
// keep bits of the aligned word that holds element idx (p == 0.5 rule), shifted so that bit i belongs to element idx + i;
// valid for the elements up to the next multiple of 32
__host__ __device__ __forceinline__ uint32_t drop_word(uint64_t key, uint64_t idx) {
  return drop_bits(key, idx >> 5) >> (uint32_t)(idx & 31);
}
// Attention-probability dropout for one (query row, key) pair; row = (b*heads + h)*T + query.
// p == 0.5 (the HHI default) takes ONE random bit per pair: bit key%32 of the hash of (row * ceil(T/32) + key/32), so a
// thread that owns several keys of one query row pays one hash per 32 keys (attention_mma.cu); other p: one hash per pair.
// other p: the 16-bit field key%2 of the hash of (row * ceil(T/2) + key/2): the two adjacent keys of an mma accumulator pair
// share one hash.
__host__ __device__ __forceinline__ bool attn_drop_keep(uint64_t key, uint64_t row, int T, int c, float p, uint32_t thr) {
  (void)thr;
  if (p == 0.5f) return (drop_bits(key, row * (uint64_t)((T + 31) >> 5) + (uint32_t)(c >> 5)) >> (c & 31)) & 1u;
  return drop_field16(drop_bits(key, row * (uint64_t)((T + 1) >> 1) + (uint32_t)(c >> 1)), (uint32_t)(c & 1)) >= drop_threshold16(p);
}
__device__ __forceinline__ float attn_drop_scale(uint64_t key, uint64_t row, int T, int c, float p, float inv_keep) {
  return attn_drop_keep(key, row, T, c, p, drop_threshold(p)) ? inv_keep : 0.0f;
}
// the two adjacent keys c (EVEN) and c + 1 of one query row - an mma accumulator pair - with one hash (general p; the p == 0.5
// kernels read whole keep words instead): a *= multiplier(c), b *= multiplier(c + 1), the same decisions as attn_drop_scale
__device__ __forceinline__ void attn_drop_scale2(uint64_t key, uint64_t row, int T, int c, float p, float inv_keep, float& a, float& b) {
  if (p == 0.5f) {
    a *= attn_drop_scale(key, row, T, c, p, inv_keep);
    b *= attn_drop_scale(key, row, T, c + 1, p, inv_keep);
    return;
  }
  const uint32_t h = drop_bits(key, row * (uint64_t)((T + 1) >> 1) + (uint32_t)(c >> 1)), thr = drop_threshold16(p);
  a = (h & 0xffffu) >= thr ? a * inv_keep : 0.0f;
  b = (h >> 16) >= thr ? b * inv_keep : 0.0f;
}
// returns the multiplier: 0 if dropped, 1/(1-p) if kept
__device__ __forceinline__ float drop_scale(uint64_t key, uint64_t idx, float p, float inv_keep, bool bit_mode = false) {
  return drop_keep(key, idx, p, drop_threshold(p), bit_mode) ? inv_keep : 0.0f;
}
// multipliers of N consecutive elements idx0 .. idx0+N-1 that do not straddle a multiple of 32 (N | 32, idx0 % N == 0):
// one hash for all of them at p == 0.5, one per element otherwise - the same decisions as drop_scale element by element
template <int N>
__device__ __forceinline__ void drop_scale_n(uint64_t key, uint64_t idx0, float p, float inv_keep, float (&m)[N]) {
  static_assert(N >= 1 && N <= 32 && (32 % N) == 0, "drop_scale_n: N must divide 32");
  if (p == 0.5f) {
    const uint32_t w = drop_word(key, idx0);
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = (w >> i) & 1u ? inv_keep : 0.0f;
  } else if constexpr ((N & 1) == 0) {       // idx0 is even: elements 2i, 2i+1 are the two halves of one hash
    const uint32_t thr = drop_threshold16(p);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const uint32_t h = drop_bits(key, (idx0 + i) >> 1);
      m[i] = (h & 0xffffu) >= thr ? inv_keep : 0.0f;
      m[i + 1] = (h >> 16) >= thr ? inv_keep : 0.0f;
    }
  } else {
    m[0] = drop_keep(key, idx0, p, 0u) ? inv_keep : 0.0f;
  }
}

}  // namespace egot2
