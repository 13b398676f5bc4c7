// api.cu — the stage-level C ABI of include/egot2.h: each entry point enqueues the kernels of one
// stage of the translator (embed, encoder layer, head+loss) on the caller's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#define EGOT2_FILE_ID 11
#include "ops.h"

namespace egot2 {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#ifdef EGOT2_TIMELINE
static void (*g_tl_setters[32])(unsigned long long*);
static int g_tl_n = 0;
void tl_register(void (*setter)(unsigned long long*)) { if (g_tl_n < 32) g_tl_setters[g_tl_n++] = setter; }
#endif

// ---- dropout epoch (common.cuh): registry of the per-translation-unit slot pointers + the library-owned slot
unsigned long long g_host_epoch = 0;
static void (*g_epoch_setters[32])(const unsigned long long*);
static int g_epoch_n = 0;
void epoch_register(void (*setter)(const unsigned long long*)) { if (g_epoch_n < 32) g_epoch_setters[g_epoch_n++] = setter; }
namespace {
__global__ void epoch_advance_kernel(unsigned long long* slot, unsigned long long add) {
  EGOT2_PDL_ENTER();
  if (threadIdx.x == 0 && blockIdx.x == 0) *slot += add;
}
constexpr int kMaxDev = 64;
unsigned long long* g_epoch_slots[kMaxDev] = {};      // one slot per device (allocated on first enable on that device)
int cur_dev() { int d = 0; return (cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < kMaxDev) ? d : 0; }
#define g_epoch_slot (g_epoch_slots[cur_dev()])
}  // namespace

bool pdl_enabled() {
  static const bool on = !(getenv("EGOT2_PDL") && atoi(getenv("EGOT2_PDL")) == 0);
  return on;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// split-K factor for a weight-gradient GEMM (small M x N output, K = all tokens): fill ~2 waves.
int suggest_split_k(int M, int N, int K) {
  const int tile = (M <= 64 || N <= 64) ? 64 : 128;
  const long long tiles = (long long)((M + tile - 1) / tile) * ((N + tile - 1) / tile);
  long long want = (2LL * sm_count() + tiles - 1) / tiles;
  long long max_by_k = K / 128;                 // at least two 64-row k-blocks per split
  if (want > max_by_k) want = max_by_k;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return (int)want;
}

// ---------------------------------------------------------------- launcher timing (CUDA events on the launch stream)
namespace {
struct ProfRec { char tag[96]; cudaEvent_t e0, e1; };
constexpr int kProfMax = 16384;
ProfRec* g_prof = nullptr;
int g_prof_n = 0;
bool g_prof_on = false;
}  // namespace
bool prof_enabled() { return g_prof_on; }
ProfScope::ProfScope(cudaStream_t stream, const char* fmt, ...) : st(stream), slot(-1) {
  static const bool trace = getenv("EGOT2_TRACE_LAUNCH") != nullptr;    // debugging: print every launcher tag to stderr
  if (trace) {
    char tag[96];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tag, sizeof(tag), fmt, ap);
    va_end(ap);
    fprintf(stderr, "[egot2] %s\n", tag);
    fflush(stderr);
  }
  if (!g_prof_on || g_prof_n >= kProfMax) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  ProfRec& r = g_prof[g_prof_n];
  if (!r.e0) { if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return; }
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(r.tag, sizeof(r.tag), fmt, ap);
  va_end(ap);
  slot = g_prof_n++;
  cudaEventRecord(r.e0, st);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, st);
}

static thread_local const char* g_gemm_impl = "none";
const char* gemm_last_impl() { return g_gemm_impl; }

static int gemm_impl_mode() {   // 0 auto, 1 force CUDA-core (diagnostics: EGOT2_GEMM=simt)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("EGOT2_GEMM");
    mode = (e && strcmp(e, "simt") == 0) ? 1 : 0;
  }
  return mode;
}

int gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.in_dtype == EGOT2_BF16 && gemm_impl_mode() == 0) {
    const int rc = gemm_sm100(a, st);
    if (rc >= 0) { if (rc == 0) g_gemm_impl = "tcgen05"; return rc; }
  }
  g_gemm_impl = "simt";
  EGOT2_CHECK(!a.trans_c, "gemm: transposed accumulation exists in the tcgen05 kernel only");
  EGOT2_CHECK(a.ln_g == nullptr, "gemm: a fused LayerNorm was requested but the tcgen05 kernel did not take the problem");
  if (a.split_stride > 0) {      // the CUDA-core GEMM has no split-K: one split, written to slab 0
    GemmArgs b = a;
    b.split_k = 1; b.split_stride = 0;
    if (a.splits_out) *a.splits_out = 1;
    return gemm_simt(b, st);
  }
  return gemm_simt(a, st);
}

static bool env_is(const char* name, const char* value) {
  const char* e = getenv(name);
  return e && strcmp(e, value) == 0;
}

static bool force_simt_attention() {   // diagnostics: EGOT2_ATTN=simt
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("EGOT2_ATTN"); mode = (e && strcmp(e, "simt") == 0) ? 1 : 0; }
  return mode == 1;
}

int attention_fwd(int dtype, int B, int T, int H, int heads, const void* qkv, void* out, float* lse, float p_drop,
                  uint64_t drop_key, cudaStream_t st) {
  if (heads > 0 && H % heads == 0 && H / heads > 128)
    return attention_wide_fwd(dtype, B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
  if (attention_mma_supported(dtype, T, H, heads) && !force_simt_attention())
    return attention_mma_fwd(B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
  if (attention_long_supported(dtype, T, H, heads) && !force_simt_attention())
    return attention_long_fwd(B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
  return attention_simt_fwd(dtype, B, T, H, heads, qkv, out, lse, p_drop, drop_key, st);
}
int attention_bwd(int dtype, int B, int T, int H, int heads, const void* qkv, const void* out, const float* lse,
                  const void* dout, void* dqkv, float p_drop, uint64_t drop_key, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  if (heads > 0 && H % heads == 0 && H / heads > 128)
    return attention_wide_bwd(dtype, B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
  if (attention_mma_supported(dtype, T, H, heads) && !force_simt_attention())
    return attention_mma_bwd(B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
  if (attention_long_supported(dtype, T, H, heads) && !force_simt_attention())
    return attention_long_bwd(B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, st);
  return attention_simt_bwd(dtype, B, T, H, heads, qkv, out, lse, dout, dqkv, p_drop, drop_key, ws, ws_bytes, st);
}

namespace {

struct Carver {   // bump allocator over the caller's workspace
  char* base; size_t size, off = 0;
  Carver(void* p, size_t n) : base((char*)p), size(n) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += align_up(bytes);
    return r;
  }
  bool ok() const { return off <= size && (base != nullptr || off == 0); }
};

// ---- side streams: work that only produces parameter gradients (weight-gradient GEMMs, bias column sums) is off the
// data-gradient critical path, and the independent per-task projections do not depend on each other.  Those launches
// are forked onto library-owned non-blocking streams with event fork/join (legal under stream capture: the CUDA graph
// gets parallel branches), so they fill the SMs the skinny critical-path kernels leave idle.  EGOT2_STREAMS=0 disables.
struct Side {
  cudaStream_t s = nullptr;
  cudaEvent_t fork[8] = {};
  cudaEvent_t join = nullptr;
  // deferred joins (egot2_side_defer): instead of joining at the end of an encoder-layer backward, the layer records a
  // `pend` event per workspace it used; the next layer backward that is handed the SAME workspace waits for it first.
  // With the caller alternating two workspaces, a layer's parameter-gradient backlog overlaps the whole next layer (or
  // the embedding backward) instead of stalling the data-gradient chain at the layer boundary.
  cudaEvent_t pend[2] = {};
  const void* pend_ws[2] = {nullptr, nullptr};
  int pend_next = 0;
  bool dirty = false;
};
int g_side_defer[kMaxDev] = {};
// One SET of three side streams per calling stream (up to kCallers sets per device, least-recently-used reuse): callers on
// different streams - the three forward/backward branches of an EgoT2-g step - then do not meet on a shared side stream,
// which would chain every branch's joins behind the other branches' parameter-gradient work.  All streams and events of a
// device are created on first use, so nothing is created while a stream capture is in progress; sharing a set (more than
// kCallers calling streams) only over-synchronises, it is never incorrect.
constexpr int kCallers = 4;
struct SideSets {
  Side sides[kCallers][3];
  cudaStream_t owner[kCallers] = {};
  unsigned long long used[kCallers] = {};
  unsigned long long tick = 0;
  int state = 0;      // 0 untried, 1 ok, -1 disabled / failed
};
SideSets g_side_sets[kMaxDev];
Side* get_side(int idx, cudaStream_t caller) {
  // per device: streams and events belong to the device that was current when they were created
  SideSets& ss = g_side_sets[cur_dev()];
  if (ss.state == 0) {
    ss.state = 1;
    if (env_is("EGOT2_STREAMS", "0")) ss.state = -1;
    for (int c = 0; c < kCallers && ss.state == 1; ++c)
      for (int i = 0; i < 3 && ss.state == 1; ++i) {
        Side& sd = ss.sides[c][i];
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) ss.state = -1;
        for (int k = 0; k < 8 && ss.state == 1; ++k)
          if (cudaEventCreateWithFlags(&sd.fork[k], cudaEventDisableTiming) != cudaSuccess) ss.state = -1;
        if (ss.state == 1 && cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) ss.state = -1;
        for (int k = 0; k < 2 && ss.state == 1; ++k)
          if (cudaEventCreateWithFlags(&sd.pend[k], cudaEventDisableTiming) != cudaSuccess) ss.state = -1;
      }
  }
  if (ss.state != 1) return nullptr;
  int slot = -1, lru = 0;
  for (int c = 0; c < kCallers; ++c) {
    if (ss.used[c] && ss.owner[c] == caller) { slot = c; break; }
    if (ss.used[c] < ss.used[lru]) lru = c;
  }
  if (slot < 0) { slot = lru; ss.owner[slot] = caller; }
  ss.used[slot] = ++ss.tick;
  return &ss.sides[slot][idx];
}
}  // namespace
int launch_priority(cudaStream_t st) {
  static int state = 0, lo = 0, hi = 0;      // 0 untried, 1 on, -1 off
  if (state == 0) {
    state = 1;
    if (env_is("EGOT2_PRIO", "0") || cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess || lo == hi) state = -1;
  }
  if (state != 1) return INT_MIN;
  {
    const SideSets& ss = g_side_sets[cur_dev()];
    if (ss.state == 1)
      for (int c = 0; c < kCallers; ++c)
        for (int i = 0; i < 3; ++i)
          if (ss.sides[c][i].s == st) return lo;          // numerically largest = least urgent
  }
  return hi;
}
namespace {
// side stream `sd` may start once everything enqueued on `main` so far has finished; returns the stream to launch on
cudaStream_t side_fork(cudaStream_t main, Side* sd, int k) {
  if (!sd || prof_enabled()) return main;      // launcher profiling: one stream, so that every launcher is timed alone
  if (cudaEventRecord(sd->fork[k], main) != cudaSuccess || cudaStreamWaitEvent(sd->s, sd->fork[k], 0) != cudaSuccess) return main;
  return sd->s;
}
int side_join(cudaStream_t main, Side* sd) {
  if (!sd || prof_enabled()) return 0;
  EGOT2_CUDA(cudaEventRecord(sd->join, sd->s));
  EGOT2_CUDA(cudaStreamWaitEvent(main, sd->join, 0));
  sd->dirty = false;
  sd->pend_ws[0] = sd->pend_ws[1] = nullptr;
  return 0;
}
bool side_deferred() { return g_side_defer[cur_dev()] != 0 && !prof_enabled(); }
// end of a layer backward in deferred mode: remember that side work reading `ws` is in flight
int side_mark_pending(Side* sd, const void* ws) {
  if (!sd) return 0;
  int slot = sd->pend_ws[0] == ws ? 0 : sd->pend_ws[1] == ws ? 1 : (sd->pend_next++ & 1);
  EGOT2_CUDA(cudaEventRecord(sd->pend[slot], sd->s));
  sd->pend_ws[slot] = ws;
  sd->dirty = true;
  return 0;
}
// start of a layer backward in deferred mode: side work that still reads this workspace must finish before it is rewritten
int side_wait_pending(cudaStream_t main, Side* sd, const void* ws) {
  if (!sd) return 0;
  for (int k = 0; k < 2; ++k)
    if (sd->pend_ws[k] == ws) {
      EGOT2_CUDA(cudaStreamWaitEvent(main, sd->pend[k], 0));
      sd->pend_ws[k] = nullptr;
    }
  return 0;
}

// weight-gradient GEMM: dW[N_out, K_in] += dY^T . X     (dY: (rows, N_out), X: (rows, K_in))
// share: how many sibling weight-gradient GEMMs run side by side (the per-task projections): each takes 1/share of the
// resident CTA slots so that all of them fit ONE wave together instead of queueing behind each other
int wgrad(int dtype, int rows, int n_out, int k_in, const void* dY, int ld_dy, int dy_rpg, int dy_gs, const void* X,
          int ld_x, int x_rpg, int x_gs, float* dW, cudaStream_t st, int share = 1) {
  // Tall outputs with a narrow input side (dW1 of the FFN: 2048 x 128) are computed as the TRANSPOSE, dW^T = X^T . dY:
  // with the token dimension as K every 128-row tile of the output re-streams the other operand, so the long side
  // belongs in N (256-wide tiles) - 3/4 of the L2 -> SM bytes of the plain orientation, which bound these GEMMs.  The
  // bulk-reduction epilogue adds the tile into the row-major dW transposed, at no extra cost.
  static const bool swap_on = !env_is("EGOT2_WGRAD_SWAP", "0");
  if (swap_on && dtype == EGOT2_BF16 && gemm_impl_mode() == 0 && k_in <= 128 && n_out >= 512 && !dy_rpg && !x_rpg && share == 1) {
    GemmArgs t;
    t.M = k_in; t.N = n_out; t.K = rows;
    t.A = X; t.lda = ld_x; t.trans_a = 1;
    t.B = dY; t.ldb = ld_dy; t.trans_b = 0;
    t.C = dW; t.ldc = k_in; t.trans_c = 1; t.in_dtype = dtype; t.out_dtype = EGOT2_F32; t.accumulate = 1;
    t.split_k = suggest_split_k(t.M, t.N, t.K);
    const int rc = t.split_k >= 2 ? gemm_sm100(t, st) : -3;
    if (rc >= 0) return rc;
  }
  GemmArgs g;
  g.M = n_out; g.N = k_in; g.K = rows;
  g.A = dY; g.lda = ld_dy; g.trans_a = 1; g.a_rpg = dy_rpg; g.a_gstride = dy_gs;
  g.B = X; g.ldb = ld_x; g.trans_b = 0; g.b_rpg = x_rpg; g.b_gstride = x_gs;
  g.C = dW; g.ldc = k_in; g.in_dtype = dtype; g.out_dtype = EGOT2_F32; g.accumulate = 1;
  g.split_k = suggest_split_k(g.M, g.N, g.K);
  if (share > 1) g.split_k = g.split_k / share > 0 ? g.split_k / share : 1;
  return gemm(g, st);
}

}  // namespace
}  // namespace egot2

using namespace egot2;

extern "C" const char* egot2_version(void) { return "egot2-b200 0.1 (sm_100a)"; }
extern "C" const char* egot2_last_error(void) { return g_err; }
extern "C" int egot2_sm_count(void) { return sm_count(); }
extern "C" uint64_t egot2_launch_count(void) { return g_launch_count; }

// EGOT2_TIMELINE builds: device buffer of 1 + 2*4000 u64 ([0] = count, then (globaltimer ns, file*100000+line) pairs); NULL = off
extern "C" int egot2_timeline_set(void* dev_buf) {
#ifdef EGOT2_TIMELINE
  for (int i = 0; i < g_tl_n; ++i) g_tl_setters[i]((unsigned long long*)dev_buf);
  return 0;
#else
  (void)dev_buf;
  EGOT2_CHECK(false, "this build has no timeline support (compile with -DEGOT2_TIMELINE)");
#endif
}

extern "C" int egot2_dropout_epoch_enable(int on) {
  if (on && !g_epoch_slot) {
    EGOT2_CUDA(cudaMalloc(&g_epoch_slot, sizeof(unsigned long long)));
    EGOT2_CUDA(cudaMemset(g_epoch_slot, 0, sizeof(unsigned long long)));
  }
  const unsigned long long* p = on ? g_epoch_slot : nullptr;
  for (int i = 0; i < g_epoch_n; ++i) g_epoch_setters[i](p);       // cudaMemcpyToSymbol: not under stream capture
  EGOT2_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int egot2_dropout_epoch_set(uint64_t value, void* stream) {
  EGOT2_CHECK(g_epoch_slot != nullptr, "dropout epoch: call egot2_dropout_epoch_enable(1) first");
  EGOT2_CUDA(cudaMemsetAsync(g_epoch_slot, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  if (value) {
    launch(epoch_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, g_epoch_slot, (unsigned long long)value);
    EGOT2_LAUNCH_CHECK();
  }
  return 0;
}
extern "C" int egot2_dropout_epoch_advance(void* stream) {
  EGOT2_CHECK(g_epoch_slot != nullptr, "dropout epoch: call egot2_dropout_epoch_enable(1) first");
  launch(epoch_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, g_epoch_slot, 1ull);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
extern "C" int egot2_dropout_epoch_host(uint64_t value) { g_host_epoch = value; return 0; }

extern "C" int egot2_pnr_metrics(int32_t B, int32_t n, const float* logits, const int64_t* label_idx,
                                 const float* label_onehot, const int64_t* sc_label, const double* fps,
                                 const int64_t* start_frame, const int64_t* end_frame, const int64_t* pnr_frame,
                                 double* err_sec, int64_t* out_counts, double* out_dist_sum, void* stream) {
  return pnr_metrics(B, n, logits, label_idx, label_onehot, sc_label, fps, start_frame, end_frame, pnr_frame, err_sec,
                     (long long*)out_counts, out_dist_sum, (cudaStream_t)stream);
}

extern "C" int egot2_prof_enable(int on) {
  if (on && !g_prof) {
    g_prof = (ProfRec*)calloc(kProfMax, sizeof(ProfRec));
    EGOT2_CHECK(g_prof != nullptr, "prof: out of host memory");
  }
  g_prof_on = on != 0;
  if (on) g_prof_n = 0;
  return 0;
}
// Text report, one line per launcher tag: "<tag>\t<launches>\t<total_us>\n" (synchronises the recorded events).
extern "C" int egot2_prof_report(char* buf, size_t buf_bytes) {
  EGOT2_CHECK(buf && buf_bytes > 0, "prof_report: no buffer");
  struct Agg { const char* tag; int n; double us; };
  Agg* agg = (Agg*)calloc(g_prof_n > 0 ? g_prof_n : 1, sizeof(Agg));
  int na = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    EGOT2_CUDA(cudaEventSynchronize(g_prof[i].e1));
    EGOT2_CUDA(cudaEventElapsedTime(&ms, g_prof[i].e0, g_prof[i].e1));
    int j = 0;
    for (; j < na; ++j) if (strcmp(agg[j].tag, g_prof[i].tag) == 0) break;
    if (j == na) { agg[na].tag = g_prof[i].tag; agg[na].n = 0; agg[na].us = 0; ++na; }
    agg[j].n += 1; agg[j].us += ms * 1e3;
  }
  size_t off = 0;
  buf[0] = 0;
  for (int j = 0; j < na; ++j) {
    int w = snprintf(buf + off, buf_bytes - off, "%s\t%d\t%.3f\n", agg[j].tag, agg[j].n, agg[j].us);
    if (w < 0 || (size_t)w >= buf_bytes - off) break;
    off += w;
  }
  free(agg);
  return 0;
}

// =============================================================================== embed stage
namespace {
// Which of the three streams (0 = the caller's, 1 / 2 = the side streams) runs segment k's projection: longest first onto the
// least loaded stream, cost = in_dim x tokens.  (Round-robin put PNR's 256-wide LTA projection behind the 8192-wide OSCC one
// while the stream of the 2048-wide one idled: the embedding stage ended 10 us later than its longest GEMM.)
void seg_stream_plan(const egot2_embed_desc* d, int* idx) {
  long long load[3] = {0, 0, 0};
  bool done[EGOT2_MAX_SEG] = {};
  for (int it = 0; it < d->n_seg; ++it) {
    int best = -1;
    long long bc = -1;
    for (int k = 0; k < d->n_seg; ++k) {
      if (done[k]) continue;
      const long long c = d->seg_has_proj[k] ? (long long)d->seg_in_dim[k] * d->seg_tokens[k] : 1;
      if (c > bc) { bc = c; best = k; }
    }
    int sidx = 0;
    for (int j = 1; j < 3; ++j) if (load[j] < load[sidx]) sidx = j;
    idx[best] = sidx;
    load[sidx] += bc;
    done[best] = true;
  }
}

// Split-K projections into an fp32 accumulator + one finishing pass (embed_extra.cu) pay off when a projection is a long-K,
// few-tile GEMM (HOI PNR / OSCC: 8192 -> 128 over clips x 16 rows = 32 output tiles for 148 SMs).
bool embed_splitk(const egot2_embed_desc* d) {
  if (env_is("EGOT2_EMBED_SPLITK", "0")) return false;
  if (d->no_ln || !embed_finish_supported(d->dtype, d->H)) return false;
  if (env_is("EGOT2_EMBED_SPLITK", "1")) return true;
  for (int k = 0; k < d->n_seg; ++k) {
    if (!d->seg_has_proj[k] || d->seg_tokens[k] == 0) continue;
    const long long tiles = (((long long)d->B * d->seg_tokens[k] + 127) / 128) * ((d->H + 127) / 128);
    if (d->seg_in_dim[k] >= 2048 && tiles * 2 <= sm_count()) return true;
  }
  return false;
}
// splits of segment k's projection: half of the resident CTA slots each (two projections run side by side on two streams),
// at least 8 k-blocks of 64 per split
int embed_seg_splits(const egot2_embed_desc* d, int k) {
  if (!d->seg_has_proj[k] || d->seg_tokens[k] == 0) return 1;
  const long long tiles = (((long long)d->B * d->seg_tokens[k] + 127) / 128) * ((d->H + 127) / 128);
  long long s = sm_count() / tiles;
  if (s > d->seg_in_dim[k] / 512) s = d->seg_in_dim[k] / 512;
  if (s > 16) s = 16;
  return s < 1 ? 1 : (int)s;
}
size_t embed_slab_floats(const egot2_embed_desc* d) {
  size_t n = 0;
  for (int k = 0; k < d->n_seg; ++k)
    if (d->seg_has_proj[k]) n += (size_t)embed_seg_splits(d, k) * d->B * d->seg_tokens[k] * d->H;
  return n;
}
}  // namespace

extern "C" size_t egot2_embed_workspace_bytes(const egot2_embed_desc* d, int backward) {
  size_t n = 256;
  if (!backward && embed_splitk(d)) n += align_up(embed_slab_floats(d) * 4);
  if (!backward && d->feat_dtype != d->dtype) {
    size_t mx = 0;
    for (int k = 0; k < d->n_seg; ++k) {
      const size_t e = (size_t)d->B * d->seg_tokens[k] * d->seg_in_dim[k];
      if (e > mx) mx = e;
    }
    n += align_up(mx * dtype_size(d->dtype));
  }
  if (backward && d->feat_dtype != d->dtype) {
    size_t mx = 0;
    for (int k = 0; k < d->n_seg; ++k) {
      const size_t e = (size_t)d->B * d->seg_tokens[k] * d->seg_in_dim[k];
      if (e > mx) mx = e;
    }
    n += align_up(mx * dtype_size(d->dtype));
  }
  if (backward) n += align_up((size_t)d->B * d->T * d->H * dtype_size(d->dtype));   // LN-backward output (dx stays intact for the table gradient)
  return n;
}

static int embed_check(const egot2_embed_desc* d) {
  EGOT2_CHECK(d->n_seg >= 1 && d->n_seg <= EGOT2_MAX_SEG, "embed: n_seg=%d out of range", d->n_seg);
  EGOT2_CHECK(d->dtype == EGOT2_F32 || d->dtype == EGOT2_BF16, "embed: bad dtype %d", d->dtype);
  EGOT2_CHECK(d->feat_dtype == d->dtype || (d->feat_dtype == EGOT2_F32 && d->dtype == EGOT2_BF16),
              "embed: features must be in the compute dtype, or fp32 with bf16 compute");
  int tot = 0;
  for (int k = 0; k < d->n_seg; ++k) {
    EGOT2_CHECK(d->seg_offset[k] == tot, "embed: segment %d offset %d != running total %d", k, d->seg_offset[k], tot);
    EGOT2_CHECK(d->seg_has_proj[k] || d->seg_in_dim[k] == d->H, "embed: pass-through segment %d must be H wide", k);
    tot += d->seg_tokens[k];
  }
  EGOT2_CHECK(tot == d->T, "embed: segments cover %d tokens, T=%d", tot, d->T);
  EGOT2_CHECK(!d->no_ln || !d->training || (d->p_feat == 0.f && d->p_embed == 0.f),
              "embed: the LayerNorm-free variant has no dropout sites (p_feat=%f p_embed=%f)", d->p_feat, d->p_embed);
  return 0;
}

extern "C" int egot2_embed_fwd(const egot2_embed_desc* d, const egot2_embed_in* in, const egot2_embed_out* out,
                               void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(embed_check(d));
  if (d->B == 0 || d->T == 0) return 0;
  const size_t es = dtype_size(d->dtype);
  Carver ws(workspace, ws_bytes);
  void* cast_buf = nullptr;
  const bool splitk = embed_splitk(d);
  float* zf = nullptr;
  EmbedSrc src;
  src.n = d->n_seg;
  if (splitk) {
    zf = (float*)ws.take(embed_slab_floats(d) * 4);
    EGOT2_CHECK(ws.ok(), "embed_fwd: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  }
  if (d->feat_dtype != d->dtype) {
    size_t mx = 0;
    for (int k = 0; k < d->n_seg; ++k) {
      const size_t e = (size_t)d->B * d->seg_tokens[k] * d->seg_in_dim[k];
      if (e > mx) mx = e;
    }
    cast_buf = ws.take(mx * dtype_size(d->dtype));
    EGOT2_CHECK(ws.ok(), "embed_fwd: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  }
  // the per-task projections are independent: task 0 stays on `st`, the others alternate over two side streams
  // (not when the features need the shared fp32 -> bf16 staging buffer)
  const bool par = d->feat_dtype == d->dtype;
  Side* sides[2] = {par ? get_side(0, st) : nullptr, par ? get_side(1, st) : nullptr};
  bool used[2] = {false, false};
  // fork BOTH side streams before anything of this stage is enqueued on `st`: a fork event recorded after segment 0's
  // launch would make the other segments wait for it (seen in the in-graph timeline: they started when it had finished)
  cudaStream_t side_st[2] = {st, st};
  int seg_idx[EGOT2_MAX_SEG] = {};
  seg_stream_plan(d, seg_idx);
  for (int i = 0; i < 2 && i + 1 < d->n_seg; ++i)
    if (sides[i]) { side_st[i] = side_fork(st, sides[i], 0); used[i] = side_st[i] != st; }
  // LayerNorm + token table (+ embedding dropout) inside the projection GEMMs' epilogue (H = 128: a tile holds whole rows):
  // possible when nothing sits between projection and LayerNorm (no feature dropout in this call) and every segment is projected
  bool ln_fused = !splitk && !d->no_ln && d->dtype == EGOT2_BF16 && d->feat_dtype == EGOT2_BF16 && d->H == 128 &&
                  !(d->training && d->p_feat > 0.f) && gemm_impl_mode() == 0 && env_is("EGOT2_GEMM_LN", "1");
  for (int k = 0; k < d->n_seg && ln_fused; ++k)
    if (d->seg_tokens[k] > 0 && (!d->seg_has_proj[k] || d->seg_in_dim[k] % 8 != 0)) ln_fused = false;
  for (int k = 0; k < d->n_seg; ++k) {
    const int Dk = d->seg_tokens[k], Kk = d->seg_in_dim[k];
    if (Dk == 0) continue;
    cudaStream_t sk = seg_idx[k] > 0 ? side_st[seg_idx[k] - 1] : st;
    char* zk = (char*)out->z + (size_t)d->seg_offset[k] * d->H * es;
    const void* feat = in->feat[k];
    if (splitk && !d->seg_has_proj[k]) {      // pass-through segment (LTA action features): the finishing pass reads it in place
      src.direct[k] = feat; src.direct_f32[k] = d->feat_dtype == EGOT2_F32 && d->dtype != EGOT2_F32;
      continue;
    }
    if (d->feat_dtype != d->dtype) {
      EGOT2_TRY(cast_f32_to(d->dtype, (const float*)feat, cast_buf, (size_t)d->B * Dk * Kk, sk));
      feat = cast_buf;
    }
    if (d->seg_has_proj[k]) {
      GemmArgs g;
      g.M = d->B * Dk; g.N = d->H; g.K = Kk;
      g.A = feat; g.lda = Kk;
      g.B = in->proj_w[k]; g.ldb = Kk; g.trans_b = 1;
      g.C = zk; g.ldc = d->H; g.c_rpg = Dk; g.c_gstride = d->T;
      g.bias = in->proj_b[k];
      g.in_dtype = d->dtype; g.out_dtype = d->dtype;
      if (splitk) {       // deterministic split-K: split s writes slab s of this segment (B*Dk, H) fp32; the first adds the bias
        g.C = zf; g.ldc = d->H; g.c_rpg = 0; g.c_gstride = 0; g.out_dtype = EGOT2_F32;
        g.split_k = embed_seg_splits(d, k); g.split_stride = (long long)d->B * Dk * d->H;
        int used = 1;
        g.splits_out = &used;
        EGOT2_TRY(gemm(g, sk));
        src.slab[k] = zf; src.splits[k] = used;
        zf += (size_t)embed_seg_splits(d, k) * d->B * Dk * d->H;
      } else {
        if (ln_fused) {
          g.ln_g = in->ln_g; g.ln_b = in->ln_b; g.ln_eps = d->ln_eps;
          g.ln_out = (char*)out->x + (size_t)d->seg_offset[k] * d->H * es; g.ln_stat = out->stat;
          g.ln_table = in->tok_table; g.ln_table_rows = d->T; g.ln_row0 = d->seg_offset[k];
          if (d->training && d->p_embed > 0.f) { g.ln_p_drop = d->p_embed; g.ln_drop_key = site_key(d->seed, SITE_EMBED, 0); }
          if (!gemm_sm100_ln_ok(g)) {        // e.g. an unaligned feature tensor: this segment cannot fuse, so none does
            ln_fused = false;                // the separate LayerNorm pass below then covers ALL rows (also those already fused)
            g.ln_g = nullptr; g.ln_b = nullptr; g.ln_out = nullptr; g.ln_stat = nullptr; g.ln_table = nullptr; g.ln_p_drop = 0.f;
          }
        }
        EGOT2_TRY(gemm(g, sk));
      }
    } else {
      EGOT2_CUDA(cudaMemcpy2DAsync(zk, (size_t)d->T * d->H * es, feat, (size_t)Dk * d->H * es, (size_t)Dk * d->H * es,
                                   d->B, cudaMemcpyDeviceToDevice, sk));
    }
  }
  for (int i = 0; i < 2; ++i) if (used[i]) EGOT2_TRY(side_join(st, sides[i]));
  const size_t n = (size_t)d->B * d->T * d->H;
  if (splitk) {
    for (int k = 0; k < d->n_seg; ++k) { src.tok_begin[k] = d->seg_offset[k]; src.tokens[k] = d->seg_tokens[k]; }
    return embed_finish(d->B, d->T, d->H, src, (d->feat_drop_tokens > 0 && d->feat_drop_tokens < d->T) ? d->feat_drop_tokens : 0,
                        d->training ? d->p_feat : 0.f, site_key(d->seed, SITE_FEAT, 0), in->ln_g, in->ln_b, d->ln_eps, in->tok_table,
                        d->training ? d->p_embed : 0.f, site_key(d->seed, SITE_EMBED, 0), out->z, out->stat, out->x, st);
  }
  if (d->training && d->p_feat > 0.f) {
    if (d->feat_drop_tokens > 0 && d->feat_drop_tokens < d->T)
      EGOT2_TRY(dropout_prefix_inplace(d->dtype, out->z, n, (size_t)d->T * d->H, (size_t)d->feat_drop_tokens * d->H, d->p_feat,
                                       site_key(d->seed, SITE_FEAT, 0), st));
    else
      EGOT2_TRY(dropout_inplace(d->dtype, out->z, n, d->p_feat, site_key(d->seed, SITE_FEAT, 0), st));
  }
  if (d->no_ln) return add_table(d->dtype, n, (size_t)d->T * d->H, out->z, in->tok_table, out->x, st);
  if (ln_fused) return 0;                 // x and stat came out of the projection GEMMs
  LayerNormArgs l;
  l.rows = d->B * d->T; l.H = d->H; l.dtype = d->dtype; l.x = out->z; l.g = in->ln_g; l.b = in->ln_b; l.eps = d->ln_eps;
  l.y = out->x; l.stat = out->stat; l.table = in->tok_table; l.table_rows = d->T;
  if (d->training && d->p_embed > 0.f) { l.p_drop = d->p_embed; l.drop_key = site_key(d->seed, SITE_EMBED, 0); }
  return layernorm_fwd(l, st);
}

extern "C" int egot2_embed_bwd(const egot2_embed_desc* d, const egot2_embed_in* in, const egot2_embed_out* saved,
                               const void* dx_c, const egot2_embed_grads* g, void* workspace, size_t ws_bytes,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(embed_check(d));
  if (d->B == 0 || d->T == 0) return 0;
  void* dx = const_cast<void*>(dx_c);
  const size_t es = dtype_size(d->dtype);
  const size_t n = (size_t)d->B * d->T * d->H;
  Carver ws(workspace, ws_bytes);
  void* cast_buf = nullptr;
  if (d->feat_dtype != d->dtype) {
    size_t mx = 0;
    for (int k = 0; k < d->n_seg; ++k) {
      const size_t e = (size_t)d->B * d->seg_tokens[k] * d->seg_in_dim[k];
      if (e > mx) mx = e;
    }
    cast_buf = ws.take(mx * es);
  }
  void* dz = ws.take(n * es);
  EGOT2_CHECK(ws.ok(), "embed_bwd: workspace too small (%zu < %zu)", ws_bytes, ws.off);
  // the embedding dropout's mask (applied after the LN + table add in the forward) is regenerated by both consumers of
  // dx instead of a separate in-place pass
  const float pe = (d->training && d->p_embed > 0.f) ? d->p_embed : 0.f;
  const uint64_t ke = site_key(d->seed, SITE_EMBED, 0);
  // token-table gradient (column sums of dx per token / per task): a side stream, beside the LayerNorm backward, which
  // therefore writes to the workspace instead of in place
  SegOut so;
  so.n = d->n_seg;
  bool want_table = g->tok_table != nullptr;
  for (int k = 0; k < d->n_seg; ++k) { so.tokens[k] = d->seg_tokens[k]; so.out[k] = g->seg_embed[k]; want_table |= g->seg_embed[k] != nullptr; }
  Side* tside = want_table ? get_side(2, st) : nullptr;
  if (want_table) EGOT2_TRY(table_grad(d->dtype, d->B, d->T, d->H, dx, g->tok_table, so, pe, ke, side_fork(st, tside, 0)));
  if (d->no_ln) {
    dz = dx;                              // x = z + table: the token gradient IS the projection-output gradient
  } else {
    LayerNormBwdArgs l;
    l.rows = d->B * d->T; l.H = d->H; l.dtype = d->dtype; l.x = saved->z; l.stat = saved->stat; l.g = in->ln_g;
    l.dy = dx; l.dx = dz; l.dg = g->ln_g; l.db = g->ln_b; l.dy_p_drop = pe; l.dy_drop_key = ke;
    EGOT2_TRY(layernorm_bwd(l, st));
    if (d->training && d->p_feat > 0.f) {
      if (d->feat_drop_tokens > 0 && d->feat_drop_tokens < d->T)
        EGOT2_TRY(dropout_prefix_inplace(d->dtype, dz, n, (size_t)d->T * d->H, (size_t)d->feat_drop_tokens * d->H, d->p_feat,
                                         site_key(d->seed, SITE_FEAT, 0), st));
      else
        EGOT2_TRY(dropout_inplace(d->dtype, dz, n, d->p_feat, site_key(d->seed, SITE_FEAT, 0), st));
    }
  }
  const bool par = d->feat_dtype == d->dtype;
  Side* sides[2] = {par ? get_side(0, st) : nullptr, par ? get_side(1, st) : nullptr};
  bool used[2] = {false, false};
  // fork BOTH side streams before anything of this stage is enqueued on `st`: a fork event recorded after segment 0's
  // launch would make the other segments wait for it (seen in the in-graph timeline: they started when it had finished)
  cudaStream_t side_st[2] = {st, st};
  int seg_idx[EGOT2_MAX_SEG] = {};
  seg_stream_plan(d, seg_idx);
  for (int i = 0; i < 2 && i + 1 < d->n_seg; ++i)
    if (sides[i]) { side_st[i] = side_fork(st, sides[i], 0); used[i] = side_st[i] != st; }
  int n_proj = 0, n_big = 0;
  long long sum_big = 0;
  for (int k = 0; k < d->n_seg; ++k) {
    const bool on = d->seg_has_proj[k] && d->seg_tokens[k] > 0;
    n_proj += on ? 1 : 0;
    n_big += (on && d->seg_in_dim[k] >= 2048) ? 1 : 0;
    if (on && d->seg_in_dim[k] >= 2048) sum_big += (long long)d->seg_in_dim[k] * d->seg_tokens[k];
  }
  // the weight-gradient GEMMs of the projections run side by side and share the resident CTA slots - among the LONG-K ones
  // only when there are any (PNR / OSCC 8192-wide beside SlowFast's 2048 / 256: the two big ones stream 134 MB and used to get
  // 32 CTAs each because the split was divided by all four)
  const int share_all = n_big > 0 ? n_big : n_proj;
  // ... and among the long-K ones in proportion to the bytes each streams (in_dim x tokens): PNR's two 8192-wide projections
  // beside SlowFast's 2048-wide one take a split of 2 each (64 CTAs) instead of 1 (32 CTAs, half of the GPU idle)
  auto big_share = [&](int Kk, int Dk) {
    const long long w = (long long)Kk * Dk;
    const long long sh = w > 0 ? (sum_big + w / 2) / w : 1;
    return (int)(sh < 1 ? 1 : sh);
  };
  bool bias_on_main = false;
  SegOut sb_main;
  {
    // every projection's bias gradient (column sums of its segment's rows of dz) in ONE launch of the per-segment
    // column-sum kernel instead of one strided column-sum launch per task
    SegOut sb;
    sb.n = d->n_seg;
    bool any = false;
    for (int k = 0; k < d->n_seg; ++k) {
      sb.tokens[k] = d->seg_tokens[k];
      sb.out[k] = (d->seg_has_proj[k] && d->seg_tokens[k] > 0) ? g->proj_b[k] : nullptr;
      any |= sb.out[k] != nullptr;
    }
    // On side stream 2 - unless the token-table gradient already runs there (HHI: two of these ~10 us reductions one after
    // the other on one stream were the end of the step; the in-graph timeline shows Adam waiting for the second): then on
    // the caller's stream, behind its projection weight gradient (below).
    Side* bside = get_side(2, st);
    if (any && tside) {
      bias_on_main = true;
      sb_main = sb;
    } else if (any) {
      EGOT2_TRY(table_grad(d->dtype, d->B, d->T, d->H, dz, nullptr, sb, 0.f, 0, side_fork(st, bside, 1)));
      if (bside && !tside) tside = bside;       // joined below
    }
  }
  for (int k = 0; k < d->n_seg; ++k) {
    const int Dk = d->seg_tokens[k], Kk = d->seg_in_dim[k];
    if (Dk == 0) continue;
    cudaStream_t sk = seg_idx[k] > 0 ? side_st[seg_idx[k] - 1] : st;
    const char* dzk = (const char*)dz + (size_t)d->seg_offset[k] * d->H * es;
    if (d->seg_has_proj[k]) {
      const void* feat = in->feat[k];
      if (d->feat_dtype != d->dtype) {
        EGOT2_TRY(cast_f32_to(d->dtype, (const float*)feat, cast_buf, (size_t)d->B * Dk * Kk, sk));
        feat = cast_buf;
      }
      if (g->proj_w[k])
        EGOT2_TRY(wgrad(d->dtype, d->B * Dk, d->H, Kk, dzk, d->H, Dk, d->T, feat, Kk, 0, 0, g->proj_w[k], sk,
                        !par ? 1 : ((n_big > 0 && Kk < 2048) ? 8 : (n_big > 0 ? big_share(Kk, Dk) : share_all))));
      if (g->dfeat[k]) {   // dF = dZ . W   (only for a trainable backbone: HHI --nofreeze)
        GemmArgs m;
        m.M = d->B * Dk; m.N = Kk; m.K = d->H;
        m.A = dzk; m.lda = d->H; m.a_rpg = Dk; m.a_gstride = d->T;
        m.B = in->proj_w[k]; m.ldb = Kk; m.trans_b = 0;
        m.C = g->dfeat[k]; m.ldc = Kk; m.in_dtype = d->dtype; m.out_dtype = EGOT2_F32;
        EGOT2_TRY(gemm(m, sk));
      }
    } else if (g->dfeat[k]) {   // pass-through stream (LTA action features come from a trainable head)
      if (d->dtype == EGOT2_F32) {
        EGOT2_CUDA(cudaMemcpy2DAsync(g->dfeat[k], (size_t)Dk * d->H * 4, dzk, (size_t)d->T * d->H * 4,
                                     (size_t)Dk * d->H * 4, d->B, cudaMemcpyDeviceToDevice, sk));
      } else {
        for (int b = 0; b < d->B; ++b)
          EGOT2_TRY(cast_to_f32(d->dtype, dzk + (size_t)b * d->T * d->H * es, g->dfeat[k] + (size_t)b * Dk * d->H,
                                (size_t)Dk * d->H, sk));
      }
    }
  }
  if (bias_on_main) EGOT2_TRY(table_grad(d->dtype, d->B, d->T, d->H, dz, nullptr, sb_main, 0.f, 0, st));
  for (int i = 0; i < 2; ++i) if (used[i]) EGOT2_TRY(side_join(st, sides[i]));
  if (tside) EGOT2_TRY(side_join(st, tside));
  return 0;
}

// =============================================================================== encoder layer
namespace {
struct LayerWs {
  void *d1, *d2, *d3, *d4, *d5, *dhid, *dqkv, *attn_ws;
  size_t attn_ws_bytes;
};
size_t layer_ws_layout(const egot2_layer_desc* d, int backward, void* base, size_t bytes, LayerWs* out) {
  Carver ws(base, bytes);
  if (backward) {
    const size_t es = dtype_size(d->dtype);
    const size_t M = (size_t)d->B * d->T;
    LayerWs w;
    w.d1 = ws.take(M * d->H * es);
    w.d2 = ws.take(M * d->H * es);
    w.d3 = ws.take(M * d->H * es);
    w.d4 = ws.take(M * d->H * es);
    w.d5 = ws.take(M * d->H * es);
    w.dhid = ws.take(M * d->FF * es);
    w.dqkv = ws.take(M * 3 * d->H * es);
    w.attn_ws_bytes = attention_bwd_workspace(d->dtype, d->B, d->T, d->H, d->heads);
    w.attn_ws = ws.take(w.attn_ws_bytes);
    if (out) *out = w;
  }
  return ws.off + 256;
}
int layer_check(const egot2_layer_desc* d) {
  EGOT2_CHECK(d->dtype == EGOT2_F32 || d->dtype == EGOT2_BF16, "layer: bad dtype %d", d->dtype);
  EGOT2_CHECK(d->heads > 0 && d->H % d->heads == 0, "layer: H=%d not divisible by heads=%d", d->H, d->heads);
  EGOT2_CHECK(d->p_drop >= 0.f && d->p_drop < 1.f, "layer: dropout p=%f out of [0,1)", d->p_drop);
  return 0;
}
}  // namespace

extern "C" size_t egot2_ffn_scratch_bytes(int32_t M) { return M > 0 ? ffn_scratch_bytes(M) : 0; }

extern "C" size_t egot2_encoder_layer_workspace_bytes(const egot2_layer_desc* d, int backward) {
  return layer_ws_layout(d, backward, nullptr, 0, nullptr);
}

extern "C" int egot2_encoder_layer_fwd(const egot2_layer_desc* d, const egot2_layer_params* p, const void* x_in,
                                       void* x_out, const egot2_layer_saved* s, void* workspace, size_t ws_bytes,
                                       void* stream) {
  (void)workspace; (void)ws_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(layer_check(d));
  const int M = d->B * d->T, H = d->H, FF = d->FF;
  if (M == 0) return 0;
  const float pd = d->training ? d->p_drop : 0.f;
  const uint32_t L = (uint32_t)d->layer_index;
  bool ln1_fused = false;
  // 1. packed in-projection: qkv = x . Win^T + bin
  {
    GemmArgs g; g.M = M; g.N = 3 * H; g.K = H; g.A = x_in; g.lda = H; g.B = p->in_proj_w; g.ldb = H; g.trans_b = 1;
    g.C = s->qkv; g.ldc = 3 * H; g.bias = p->in_proj_b; g.in_dtype = d->dtype; g.out_dtype = d->dtype;
    EGOT2_TRY(gemm(g, st));
  }
  // 2. per-clip, per-head softmax(q k^T / sqrt(dh)) v
  EGOT2_TRY(attention_fwd(d->dtype, d->B, d->T, H, d->heads, s->qkv, s->attn, s->lse, pd, site_key(d->seed, SITE_ATTN, L), st));
  // 3. y1 = x + dropout1(attn . Wo^T + bo)
  {
    GemmArgs g; g.M = M; g.N = H; g.K = H; g.A = s->attn; g.lda = H; g.B = p->out_proj_w; g.ldb = H; g.trans_b = 1;
    g.C = s->y1; g.ldc = H; g.bias = p->out_proj_b; g.residual = x_in; g.ldr = H;
    g.p_drop = pd; g.drop_key = site_key(d->seed, SITE_DROP1, L); g.in_dtype = d->dtype; g.out_dtype = d->dtype;
    // 4. x1 = norm1(y1) inside the same epilogue when a tile holds whole rows (H = 128)
    g.ln_g = p->norm1_g; g.ln_b = p->norm1_b; g.ln_eps = d->ln_eps; g.ln_out = s->x1; g.ln_stat = s->stat1;
    ln1_fused = d->dtype == EGOT2_BF16 && gemm_sm100_ln_ok(g) && !env_is("EGOT2_GEMM", "simt");
    if (!ln1_fused) { g.ln_g = nullptr; g.ln_b = nullptr; g.ln_out = nullptr; g.ln_stat = nullptr; }
    EGOT2_TRY(gemm(g, st));
  }
  // 4. x1 = norm1(y1)
  if (!ln1_fused) {
    LayerNormArgs l; l.rows = M; l.H = H; l.dtype = d->dtype; l.x = s->y1; l.g = p->norm1_g; l.b = p->norm1_b;
    l.eps = d->ln_eps; l.y = s->x1; l.stat = s->stat1;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  // 5-7 fused (bf16, H = 128): the hidden activation stays on chip between the two GEMMs
  if (ffn_fused_supported(d->dtype, H, FF) && !env_is("EGOT2_FFN", "unfused"))
    return ffn_fused_fwd(M, FF, s->x1, p->lin1_w, p->lin1_b, p->lin2_w, p->lin2_b, p->norm2_g, p->norm2_b, d->ln_eps,
                         s->hid, s->hid_mask, s->y2, s->stat2, x_out, pd, site_key(d->seed, SITE_FFN, L),
                         site_key(d->seed, SITE_DROP2, L), s->ffn_scratch, st);
  // 5. hid = dropout(relu(x1 . W1^T + b1))
  {
    GemmArgs g; g.M = M; g.N = FF; g.K = H; g.A = s->x1; g.lda = H; g.B = p->lin1_w; g.ldb = H; g.trans_b = 1;
    g.C = s->hid; g.ldc = FF; g.bias = p->lin1_b; g.relu = 1;
    g.p_drop = pd; g.drop_key = site_key(d->seed, SITE_FFN, L); g.drop_bit_mode = 1; g.in_dtype = d->dtype; g.out_dtype = d->dtype;
    EGOT2_TRY(gemm(g, st));
  }
  // 6. y2 = x1 + dropout2(hid . W2^T + b2)
  {
    GemmArgs g; g.M = M; g.N = H; g.K = FF; g.A = s->hid; g.lda = FF; g.B = p->lin2_w; g.ldb = FF; g.trans_b = 1;
    g.C = s->y2; g.ldc = H; g.bias = p->lin2_b; g.residual = s->x1; g.ldr = H;
    g.p_drop = pd; g.drop_key = site_key(d->seed, SITE_DROP2, L); g.in_dtype = d->dtype; g.out_dtype = d->dtype;
    EGOT2_TRY(gemm(g, st));
  }
  // 7. x_out = norm2(y2)
  {
    LayerNormArgs l; l.rows = M; l.H = H; l.dtype = d->dtype; l.x = s->y2; l.g = p->norm2_g; l.b = p->norm2_b;
    l.eps = d->ln_eps; l.y = x_out; l.stat = s->stat2;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  return 0;
}

extern "C" int egot2_encoder_layer_bwd(const egot2_layer_desc* d, const egot2_layer_params* p, const void* x_in,
                                       const egot2_layer_saved* s, void* dx_out, void* dx_in,
                                       const egot2_layer_grads* g, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(layer_check(d));
  const int M = d->B * d->T, H = d->H, FF = d->FF;
  if (M == 0) return 0;
  LayerWs w;
  const size_t need = layer_ws_layout(d, 1, workspace, ws_bytes, &w);
  EGOT2_CHECK(workspace && ws_bytes + 256 >= need, "encoder_layer_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const size_t es = dtype_size(d->dtype);
  const float pd = d->training ? d->p_drop : 0.f;
  const float inv_keep = pd > 0.f ? 1.f / (1.f - pd) : 1.f;
  const uint32_t L = (uint32_t)d->layer_index;
  const int dt = d->dtype;

  // Parameter-gradient work (weight-gradient GEMMs, bias column sums) runs on a side stream, forked after each producer
  // and joined before returning; the data-gradient chain stays on `st`.  Every buffer a side launch reads is written
  // once per call (d1..d5, dhid, dqkv), so the main chain never overwrites what the side stream may still be reading.
  // Three side streams: the four parameter-gradient groups total more kernel time than the data-gradient chain they hang
  // off, so on one stream the join at the end of the layer stalled the main chain; spread out, every group finishes
  // before the chain does.
  Side* sd0 = get_side(0, st);
  Side* sd1 = get_side(1, st);
  Side* sd2 = get_side(2, st);
  const bool defer = side_deferred();
  if (defer) {
    EGOT2_TRY(side_wait_pending(st, sd0, workspace));
    EGOT2_TRY(side_wait_pending(st, sd1, workspace));
    EGOT2_TRY(side_wait_pending(st, sd2, workspace));
  }

  // 1. through norm2: d1 = dL/dy2, and (same kernel) d2 = dropout2 mask applied to d1 = dL/d(linear2 out)
  const void* d2 = w.d1;
  {
    LayerNormBwdArgs l; l.rows = M; l.H = H; l.dtype = dt; l.x = s->y2; l.stat = s->stat2; l.g = p->norm2_g;
    l.dy = dx_out; l.dx = w.d1; l.dg = g->norm2_g; l.db = g->norm2_b;
    if (pd > 0.f) { l.dx2 = w.d2; l.dx2_p_drop = pd; l.dx2_drop_key = site_key(d->seed, SITE_DROP2, L); d2 = w.d2; }
    l.dcol = g->lin2_b;            // db2 = colsum(d2) comes out of the same pass
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  //    linear2 (side): dW2 += d2^T . hid
  // EGOT2_LATE_WGRAD=1 forks the two big weight-gradient GEMMs only after the LayerNorm1 backward instead of right after
  // their producers (their long-lived CTAs cannot be pre-empted and slow the short chain kernels behind the fused FFN
  // backward).  Measured: 333.6 vs 330.5 us per step - the work only moves into the backlog before the join - so off.
  static const bool late_wgrad = env_is("EGOT2_LATE_WGRAD", "1");
  auto fork_dw2 = [&]() -> int {
    cudaStream_t ss = side_fork(st, sd0, 0);
    return wgrad(dt, M, H, FF, d2, H, 0, 0, s->hid, FF, 0, 0, g->lin2_w, ss);
  };
  if (!late_wgrad) EGOT2_TRY(fork_dw2());
  //    dhid = (d2 . W2) * relu'(hid) * ffn-dropout scale ; d3 = dhid . W1 + d1 (residual branch)
  const bool fused_dx = s->hid_mask && ffn_fused_supported(dt, H, FF) && !env_is("EGOT2_FFN", "unfused");
  if (fused_dx) {
    // one tcgen05 kernel: dhid = gate(d2 . W2) and d3 = dhid . W1 + d1, dhid never re-read for the second GEMM
    EGOT2_TRY(ffn_fused_bwd_dx(M, FF, d2, pd > 0.f ? w.d1 : nullptr, s->hid_mask, p->lin1_w, p->lin2_w, pd, w.dhid, w.d3, s->ffn_scratch,
                               g->lin1_b, st));          // + db1: the bias gradient comes out of the same pass over dhid
  } else {
    GemmArgs m; m.M = M; m.N = FF; m.K = H; m.A = d2; m.lda = H; m.B = p->lin2_w; m.ldb = FF; m.trans_b = 0;
    m.C = w.dhid; m.ldc = FF; m.mask = s->hid; m.ldm = FF; m.mask_scale = inv_keep; m.in_dtype = dt; m.out_dtype = dt;
    EGOT2_TRY(gemm(m, st));
    GemmArgs n; n.M = M; n.N = H; n.K = FF; n.A = w.dhid; n.lda = FF; n.B = p->lin1_w; n.ldb = H; n.trans_b = 0;
    n.C = w.d3; n.ldc = H; n.residual = w.d1; n.ldr = H; n.in_dtype = dt; n.out_dtype = dt;
    EGOT2_TRY(gemm(n, st));
  }
  // 3. linear1 (side): dW1 += dhid^T . x1 ; db1 += colsum(dhid)
  auto fork_dw1 = [&]() -> int {
    cudaStream_t ss = side_fork(st, sd1, 1);
    EGOT2_TRY(wgrad(dt, M, FF, H, w.dhid, FF, 0, 0, s->x1, H, 0, 0, g->lin1_w, ss));
    if (!fused_dx) {
      cudaStream_t s2 = side_fork(st, sd2, 1);    // the bias sum reads dhid beside the GEMM (both mostly from L2)
      EGOT2_TRY(colsum_accum(dt, M, FF, w.dhid, FF, 0, 0, g->lin1_b, s2));
    }
    return 0;
  };
  if (!late_wgrad) EGOT2_TRY(fork_dw1());
  // 4. through norm1: d4 = dL/dy1, and (same kernel) d5 = dropout1 mask applied to it = dL/d(out_proj out)
  const void* dyo = w.d4;
  {
    LayerNormBwdArgs l; l.rows = M; l.H = H; l.dtype = dt; l.x = s->y1; l.stat = s->stat1; l.g = p->norm1_g;
    l.dy = w.d3; l.dx = w.d4; l.dg = g->norm1_g; l.db = g->norm1_b;
    if (pd > 0.f) { l.dx2 = w.d5; l.dx2_p_drop = pd; l.dx2_drop_key = site_key(d->seed, SITE_DROP1, L); dyo = w.d5; }
    l.dcol = g->out_proj_b;        // dbo = colsum(dyo) comes out of the same pass
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  if (late_wgrad) { EGOT2_TRY(fork_dw2()); EGOT2_TRY(fork_dw1()); }
  // 5. out_proj (side): dWo += dyo^T . attn ; dbo += colsum(dyo)      main: dattn = dyo . Wo  (d3 is free again)
  {
    cudaStream_t ss = side_fork(st, sd2, 2);
    EGOT2_TRY(wgrad(dt, M, H, H, dyo, H, 0, 0, s->attn, H, 0, 0, g->out_proj_w, ss));
  }
  {
    GemmArgs m; m.M = M; m.N = H; m.K = H; m.A = dyo; m.lda = H; m.B = p->out_proj_w; m.ldb = H; m.trans_b = 0;
    m.C = w.d3; m.ldc = H; m.in_dtype = dt; m.out_dtype = dt;
    EGOT2_TRY(gemm(m, st));
  }
  // 6. attention backward -> dqkv
  EGOT2_TRY(attention_bwd(dt, d->B, d->T, H, d->heads, s->qkv, s->attn, s->lse, w.d3, w.dqkv, pd,
                          site_key(d->seed, SITE_ATTN, L), w.attn_ws, w.attn_ws_bytes, st));
  // 7. in_proj (side): dWin += dqkv^T . x ; dbin += colsum(dqkv)      main: dx = dqkv . Win + d4 (residual branch)
  {
    cudaStream_t ss = side_fork(st, sd0, 3);
    EGOT2_TRY(wgrad(dt, M, 3 * H, H, w.dqkv, 3 * H, 0, 0, x_in, H, 0, 0, g->in_proj_w, ss));
    cudaStream_t s2 = side_fork(st, sd2, 3);
    EGOT2_TRY(colsum_accum(dt, M, 3 * H, w.dqkv, 3 * H, 0, 0, g->in_proj_b, s2));
  }
  {
    GemmArgs m; m.M = M; m.N = H; m.K = 3 * H; m.A = w.dqkv; m.lda = 3 * H; m.B = p->in_proj_w; m.ldb = H; m.trans_b = 0;
    m.C = dx_in; m.ldc = H; m.residual = w.d4; m.ldr = H; m.in_dtype = dt; m.out_dtype = dt;
    EGOT2_TRY(gemm(m, st));
  }
  if (defer) {
    EGOT2_TRY(side_mark_pending(sd0, workspace));
    EGOT2_TRY(side_mark_pending(sd1, workspace));
    return side_mark_pending(sd2, workspace);
  }
  EGOT2_TRY(side_join(st, sd0));
  EGOT2_TRY(side_join(st, sd1));
  return side_join(st, sd2);
}

extern "C" int egot2_side_defer(int on) { g_side_defer[cur_dev()] = on ? 1 : 0; return 0; }

extern "C" int egot2_side_join_all(void* stream) {
  SideSets& ss = g_side_sets[cur_dev()];
  if (ss.state == 1)
    for (int c = 0; c < kCallers; ++c)
      for (int i = 0; i < 3; ++i)
        if (ss.sides[c][i].dirty) EGOT2_TRY(side_join((cudaStream_t)stream, &ss.sides[c][i]));
  return 0;
}

// =============================================================================== head + loss
extern "C" int egot2_head_rows(const egot2_head_desc* d) { return d->pool ? d->B : d->B * d->row_tokens; }

extern "C" size_t egot2_head_workspace_bytes(const egot2_head_desc* d) {
  const size_t rows = (size_t)egot2_head_rows(d), es = dtype_size(d->dtype);
  const size_t ldl = ((size_t)d->n_out + 7) / 8 * 8;      // low-precision dlogits rows are padded to 16 B for TMA
  return align_up(rows * ldl * es) + align_up(rows * d->H * es) + align_up(rows * d->H * 4) + 256;
}

extern "C" int egot2_head_loss_fwd(const egot2_head_desc* d, const egot2_head_in* in, const egot2_head_out* out,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = egot2_head_rows(d);
  if (rows == 0) return 0;
  EGOT2_CHECK(d->pool || (d->row_tokens > 0 && d->row_tokens <= d->T), "head: row_tokens=%d out of range", d->row_tokens);
  if (head_fused_supported(*d)) {
    if (d->loss != EGOT2_LOSS_NONE)
      EGOT2_CHECK(in->labels && out->row_loss && out->loss, "head_loss_fwd: labels/row_loss/loss buffers required");
    EGOT2_TRY(head_fused_fwd(*d, *in, *out, st));          // logits + per-row loss terms + argmax in one kernel
    return loss_reduce(*d, rows, out->row_loss, out->loss, st);
  }
  EGOT2_TRY(pool_fwd(d->dtype, d->B, d->T, d->H, d->pool, d->row_tokens, in->x, out->pooled, st));
  const float ph = d->training ? d->p_head : 0.f;
  if (d->use_ln) {
    LayerNormArgs l; l.rows = rows; l.H = d->H; l.dtype = d->dtype; l.x = out->pooled; l.x_is_f32 = 1; l.g = in->ln_g;
    l.b = in->ln_b; l.eps = d->ln_eps; l.y = out->g; l.stat = out->stat;
    l.p_drop = ph; l.drop_key = site_key(d->seed, SITE_HEAD, 0);
    EGOT2_TRY(layernorm_fwd(l, st));
  } else {
    EGOT2_TRY(cast_f32_to(d->dtype, out->pooled, out->g, (size_t)rows * d->H, st));
    EGOT2_TRY(dropout_inplace(d->dtype, out->g, (size_t)rows * d->H, ph, site_key(d->seed, SITE_HEAD, 0), st));
  }
  {
    GemmArgs g; g.M = rows; g.N = d->n_out; g.K = d->H; g.A = out->g; g.lda = d->H; g.B = in->w; g.ldb = d->H; g.trans_b = 1;
    g.C = out->logits; g.ldc = d->n_out; g.bias = in->b; g.in_dtype = d->dtype; g.out_dtype = EGOT2_F32;
    EGOT2_TRY(gemm(g, st));
  }
  return loss_fwd(*d, rows, out->logits, in->labels, in->class_weight, out->row_loss, out->loss, out->argmax, st);
}

extern "C" int egot2_head_loss_bwd(const egot2_head_desc* d, const egot2_head_in* in, const egot2_head_out* saved,
                                   float* dlogits, float dloss_scale, void* dx, const egot2_head_grads* g,
                                   void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = egot2_head_rows(d);
  if (rows == 0) return 0;
  const size_t es = dtype_size(d->dtype);
  if (head_fused_supported(*d))       // d(loss)/d(logits) is derived inside the kernel (and written to `dlogits`)
    return head_fused_bwd(*d, *in, *saved, dlogits, dloss_scale, dx, *g, st);
  EGOT2_CHECK(workspace && ws_bytes >= egot2_head_workspace_bytes(d) - 256, "head_loss_bwd: workspace too small");
  Carver ws(workspace, ws_bytes);
  const int ldl = d->dtype == EGOT2_F32 ? d->n_out : (d->n_out + 7) / 8 * 8;
  void* dl_lp = ws.take((size_t)rows * ldl * es);
  void* dg = ws.take((size_t)rows * d->H * es);
  float* dpooled = (float*)ws.take((size_t)rows * d->H * 4);
  if (d->loss != EGOT2_LOSS_NONE)
    EGOT2_TRY(loss_bwd(*d, rows, saved->logits, in->labels, in->class_weight, saved->loss, dloss_scale, dlogits, st));
  const void* dl = dlogits;
  if (d->dtype != EGOT2_F32) {
    // bf16 copy with rows padded to a multiple of 8 elements: n_out = 20 x 593 = 11860 would otherwise miss the TMA
    // 16-byte pitch rule and push both head-gradient GEMMs onto the CUDA-core path
    EGOT2_TRY(cast_rows_f32_to_bf16(dlogits, rows, d->n_out, dl_lp, ldl, st));
    dl = dl_lp;
  }
  if (g->w) EGOT2_TRY(wgrad(d->dtype, rows, d->n_out, d->H, dl, ldl, 0, 0, saved->g, d->H, 0, 0, g->w, st));
  if (g->b) EGOT2_TRY(colsum_accum(EGOT2_F32, rows, d->n_out, dlogits, d->n_out, 0, 0, g->b, st));
  const float ph = d->training ? d->p_head : 0.f;
  // Wide heads (LTA: 20 x 593 classes) make dG = dlogits . W a long-K GEMM with a small output - 8 tiles for 148 SMs.  Without a
  // head LayerNorm the next consumer wants fp32 anyway: split K and accumulate straight into `dpooled`.
  if (d->dtype == EGOT2_BF16 && !d->use_ln && d->n_out >= 2048 && !env_is("EGOT2_HEAD_SPLITK", "0")) {
    EGOT2_CUDA(cudaMemsetAsync(dpooled, 0, (size_t)rows * d->H * 4, st));
    GemmArgs m; m.M = rows; m.N = d->H; m.K = d->n_out; m.A = dl; m.lda = ldl; m.B = in->w; m.ldb = d->H; m.trans_b = 0;
    m.C = dpooled; m.ldc = d->H; m.in_dtype = d->dtype; m.out_dtype = EGOT2_F32; m.accumulate = 1;
    m.split_k = suggest_split_k(m.M, m.N, m.K);
    EGOT2_TRY(gemm(m, st));
    EGOT2_TRY(dropout_inplace(EGOT2_F32, dpooled, (size_t)rows * d->H, ph, site_key(d->seed, SITE_HEAD, 0), st));
    return pool_bwd(d->dtype, d->B, d->T, d->H, d->pool, d->row_tokens, dpooled, dx, st);
  }
  {
    GemmArgs m; m.M = rows; m.N = d->H; m.K = d->n_out; m.A = dl; m.lda = ldl; m.B = in->w; m.ldb = d->H; m.trans_b = 0;
    m.C = dg; m.ldc = d->H; m.in_dtype = d->dtype; m.out_dtype = d->dtype;
    EGOT2_TRY(gemm(m, st));
  }
  EGOT2_TRY(dropout_inplace(d->dtype, dg, (size_t)rows * d->H, ph, site_key(d->seed, SITE_HEAD, 0), st));
  if (d->use_ln) {
    LayerNormBwdArgs l; l.rows = rows; l.H = d->H; l.dtype = d->dtype; l.x = saved->pooled; l.x_is_f32 = 1;
    l.stat = saved->stat; l.g = in->ln_g; l.dy = dg; l.dy_is_f32 = (d->dtype == EGOT2_F32); l.dx = dpooled; l.dx_is_f32 = 1;
    l.dg = g->ln_g; l.db = g->ln_b;
    EGOT2_TRY(layernorm_bwd(l, st));
  } else {
    EGOT2_TRY(cast_to_f32(d->dtype, dg, dpooled, (size_t)rows * d->H, st));
  }
  return pool_bwd(d->dtype, d->B, d->T, d->H, d->pool, d->row_tokens, dpooled, dx, st);
}

// =============================================================================== op-level entry points
extern "C" int egot2_gemm(int32_t dtype, int32_t M, int32_t N, int32_t K, const void* A, int32_t trans_a, const void* B,
                          int32_t trans_b, const float* bias, int32_t relu, void* C, int32_t c_is_f32,
                          int32_t accumulate, void* stream) {
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.trans_a = trans_a; g.lda = trans_a ? M : K;
  g.B = B; g.trans_b = trans_b; g.ldb = trans_b ? K : N;
  g.C = C; g.ldc = N; g.bias = bias; g.relu = relu;
  g.in_dtype = dtype; g.out_dtype = c_is_f32 ? EGOT2_F32 : dtype; g.accumulate = accumulate;
  if (accumulate && !relu) g.split_k = suggest_split_k(M, N, K);
  return gemm(g, (cudaStream_t)stream);
}

extern "C" int egot2_sum_into_f32(float* dst, float* a, float* b, size_t n, void* stream) {
  return sum_into_clear(dst, a, b, n, (cudaStream_t)stream);
}

extern "C" const char* egot2_gemm_last_impl(void) { return gemm_last_impl(); }

extern "C" int egot2_attention_fwd(int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t heads, const void* qkv,
                                   void* out, float* lse, float p_drop, int32_t training, uint64_t seed, void* stream) {
  return attention_fwd(dtype, B, T, H, heads, qkv, out, lse, training ? p_drop : 0.f, site_key(seed, SITE_ATTN, 0),
                       (cudaStream_t)stream);
}
extern "C" int egot2_attention_bwd(int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t heads, const void* qkv,
                                   const void* out, const float* lse, const void* dout, void* dqkv, float p_drop,
                                   int32_t training, uint64_t seed, void* workspace, size_t ws_bytes, void* stream) {
  return attention_bwd(dtype, B, T, H, heads, qkv, out, lse, dout, dqkv, training ? p_drop : 0.f,
                       site_key(seed, SITE_ATTN, 0), workspace, ws_bytes, (cudaStream_t)stream);
}
extern "C" size_t egot2_attention_bwd_workspace_bytes(int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t heads) {
  return attention_bwd_workspace(dtype, B, T, H, heads);
}
extern "C" int egot2_pool_fwd(int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t pool, int32_t row_tokens,
                              const void* x, float* pooled, void* stream) {
  return pool_fwd(dtype, B, T, H, pool, row_tokens, x, pooled, (cudaStream_t)stream);
}
extern "C" int egot2_pool_bwd(int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t pool, int32_t row_tokens,
                              const float* dpooled, void* dx, void* stream) {
  return pool_bwd(dtype, B, T, H, pool, row_tokens, dpooled, dx, (cudaStream_t)stream);
}
