// decoder.cu — EgoT2-g task-prompt decoder stages of the C ABI (include/egot2.h): one nn.TransformerDecoderLayer
// forward / backward and the prompt-token embedding.  Reference: HHI/models/multitask/task_prompt_model.py:163-172
// (CustomDecoderLayer), :187-194 (decoder construction), :260-269 (decode()).
// The token-wise GEMMs go through the same dispatch as the encoder (tcgen05 when the shape qualifies); the prompt has
// only 2 tokens per row, so attention uses the small-query kernels of attention_small.cu.
#include <math.h>

#define EGOT2_FILE_ID 7
#include "ops.h"

namespace egot2 {
namespace {

struct Carver {
  char* base; size_t size, off = 0;
  Carver(void* p, size_t n) : base((char*)p), size(n) {}
  void* take(size_t bytes) { void* r = base ? base + off : nullptr; off += align_up(bytes); return r; }
};

struct DecWs {
  void *d1, *d2, *dm, *da, *dhid, *dq_lp, *dkv_lp, *dqkv_lp;
  float *dq32, *dkv32, *dqkv32;
};
size_t dec_ws_layout(const egot2_decoder_desc* d, void* base, size_t bytes, DecWs* w) {
  Carver c(base, bytes);
  const size_t es = dtype_size(d->dtype), R = (size_t)d->rows * d->S, H = d->H;
  DecWs t;
  t.d1 = c.take(R * H * es); t.d2 = c.take(R * H * es); t.dm = c.take(R * H * es); t.da = c.take(R * H * es);
  t.dhid = c.take(R * d->FF * es);
  t.dq32 = (float*)c.take(R * H * 4); t.dkv32 = (float*)c.take((size_t)d->mem_rows * 2 * H * 4);
  t.dqkv32 = (float*)c.take(R * 3 * H * 4);
  t.dq_lp = c.take(R * H * es); t.dkv_lp = c.take((size_t)d->mem_rows * 2 * H * es); t.dqkv_lp = c.take(R * 3 * H * es);
  if (w) *w = t;
  return c.off + 256;
}

int dec_check(const egot2_decoder_desc* d) {
  EGOT2_CHECK(d->dtype == EGOT2_F32 || d->dtype == EGOT2_BF16, "decoder: bad dtype %d", d->dtype);
  EGOT2_CHECK(d->heads > 0 && d->H % d->heads == 0, "decoder: H=%d not divisible by heads=%d", d->H, d->heads);
  EGOT2_CHECK(d->S >= 1 && d->M >= 1 && d->kv_inner >= 1, "decoder: bad geometry S=%d M=%d kv_inner=%d", d->S, d->M, d->kv_inner);
  EGOT2_CHECK(d->p_drop >= 0.f && d->p_drop < 1.f, "decoder: dropout p=%f out of [0,1)", d->p_drop);
  // last memory row any (row, key) pair can touch must exist
  const long long n = d->rows - 1;
  const long long last = (n / d->kv_inner) * d->kv_outer + (long long)(d->M - 1) * d->kv_jstride + (n % d->kv_inner) * d->kv_istride;
  EGOT2_CHECK(d->rows == 0 || last < d->mem_rows, "decoder: memory mapping reaches row %lld of %d", last, d->mem_rows);
  return 0;
}

int lin(int dt, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc,
        cudaStream_t st, int relu = 0, float p = 0.f, uint64_t key = 0, const void* res = nullptr, int bit_mode = 0) {
  GemmArgs g; g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.trans_b = 1; g.C = C; g.ldc = ldc;
  g.bias = bias; g.relu = relu; g.p_drop = p; g.drop_key = key; g.drop_bit_mode = bit_mode; g.residual = res; g.ldr = ldc;
  g.in_dtype = dt; g.out_dtype = dt;
  return gemm(g, st);
}
// dX = dY . W  (+ residual)      dY (M,N_out) ; W (N_out, K_in) row-major
int dgrad(int dt, int M, int n_out, int k_in, const void* dY, int ld_dy, const void* W, int ldw, void* dX, int ldx,
          cudaStream_t st, const void* res = nullptr, const void* mask = nullptr, float mask_scale = 1.f) {
  GemmArgs g; g.M = M; g.N = k_in; g.K = n_out; g.A = dY; g.lda = ld_dy; g.B = W; g.ldb = ldw; g.trans_b = 0; g.C = dX; g.ldc = ldx;
  g.residual = res; g.ldr = ldx; g.mask = mask; g.ldm = ldx; g.mask_scale = mask_scale; g.in_dtype = dt; g.out_dtype = dt;
  return gemm(g, st);
}
int wgrad2(int dt, int rows, int n_out, int k_in, const void* dY, int ld_dy, const void* X, int ld_x, float* dW,
           cudaStream_t st) {
  GemmArgs g; g.M = n_out; g.N = k_in; g.K = rows; g.A = dY; g.lda = ld_dy; g.trans_a = 1; g.B = X; g.ldb = ld_x; g.trans_b = 0;
  g.C = dW; g.ldc = k_in; g.in_dtype = dt; g.out_dtype = EGOT2_F32; g.accumulate = 1;
  g.split_k = suggest_split_k(g.M, g.N, g.K);
  return gemm(g, st);
}

template <typename T>
__global__ void prompt_embed_fwd_kernel(int rows, int S, int H, const int64_t* __restrict__ tok, const float* __restrict__ emb,
                                        const float* __restrict__ pe, float p, uint64_t key, T* __restrict__ y) {
  EGOT2_PDL_ENTER();
  const int r = blockIdx.x, s = r % S;
  const float sc = sqrtf((float)H), inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const float* e = emb + (size_t)tok[r] * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = e[c] * sc + pe[(size_t)s * H + c];
    if (p > 0.f) v *= drop_scale(key ^ egot2_ep, (uint64_t)r * H + c, p, inv_keep);
    y[(size_t)r * H + c] = from_f32<T>(v);
  }
}
template <typename T>
__global__ void prompt_embed_bwd_kernel(int rows, int S, int H, const int64_t* __restrict__ tok, const T* __restrict__ dy,
                                        float p, uint64_t key, float* __restrict__ demb) {
  EGOT2_PDL_ENTER();
  const int r = blockIdx.x;
  const float sc = sqrtf((float)H), inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  float* e = demb + (size_t)tok[r] * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = to_f32(dy[(size_t)r * H + c]) * sc;
    if (p > 0.f) v *= drop_scale(key ^ egot2_ep, (uint64_t)r * H + c, p, inv_keep);
    atomicAdd(e + c, v);
  }
}

}  // namespace
}  // namespace egot2

using namespace egot2;

extern "C" size_t egot2_decoder_layer_workspace_bytes(const egot2_decoder_desc* d) {
  return dec_ws_layout(d, nullptr, 0, nullptr);
}

extern "C" int egot2_decoder_layer_fwd(const egot2_decoder_desc* d, const egot2_decoder_params* p, const void* y_in,
                                       const void* mem, void* y_out, const egot2_decoder_saved* s, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(dec_check(d));
  const int R = d->rows * d->S, H = d->H, FF = d->FF, dt = d->dtype;
  if (R == 0) return 0;
  const size_t es = dtype_size(dt);
  const float pd = d->training ? d->p_drop : 0.f;
  const uint32_t L = (uint32_t)d->layer_index;
  // 1-2. masked self-attention over the prompt
  EGOT2_TRY(lin(dt, R, 3 * H, H, y_in, H, p->sa_in_w, H, p->sa_in_b, s->qkv, 3 * H, st));
  {
    SmallAttnArgs a; a.dtype = dt; a.rows = d->rows; a.S = d->S; a.M = d->S; a.H = H; a.heads = d->heads; a.causal = 1;
    a.q = s->qkv; a.ldq = 3 * H; a.k = (const char*)s->qkv + (size_t)H * es; a.v = (const char*)s->qkv + (size_t)2 * H * es;
    a.ldkv = 3 * H; a.kv_inner = 1; a.kv_outer = d->S; a.kv_jstride = 1; a.kv_istride = 0;
    a.out = s->a1; a.ldo = H; a.p_drop = pd; a.drop_key = site_key(d->seed, SITE_DEC_SELF, L);
    EGOT2_TRY(small_attn_fwd(a, st));
  }
  // 3. y1 = y + dropout1(a1 . Wo^T + bo) ; x1 = norm1(y1)
  EGOT2_TRY(lin(dt, R, H, H, s->a1, H, p->sa_out_w, H, p->sa_out_b, s->y1, H, st, 0, pd, site_key(d->seed, SITE_DEC_DROP1, L), y_in));
  {
    LayerNormArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y1; l.g = p->norm1_g; l.b = p->norm1_b; l.eps = d->ln_eps;
    l.y = s->x1; l.stat = s->stat1;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  // 4-5. cross-attention: queries from the prompt, keys/values from every encoder token (rows of in_proj: q | k | v)
  EGOT2_TRY(lin(dt, R, H, H, s->x1, H, p->ca_in_w, H, p->ca_in_b, s->qc, H, st));
  EGOT2_TRY(lin(dt, d->mem_rows, 2 * H, H, mem, H, (const char*)p->ca_in_w + (size_t)H * H * es, H, p->ca_in_b + H, s->kvc, 2 * H, st));
  {
    SmallAttnArgs a; a.dtype = dt; a.rows = d->rows; a.S = d->S; a.M = d->M; a.H = H; a.heads = d->heads; a.causal = 0;
    a.q = s->qc; a.ldq = H; a.k = s->kvc; a.v = (const char*)s->kvc + (size_t)H * es; a.ldkv = 2 * H;
    a.kv_inner = d->kv_inner; a.kv_outer = d->kv_outer; a.kv_jstride = d->kv_jstride; a.kv_istride = d->kv_istride;
    a.out = s->a2; a.ldo = H; a.p_drop = pd; a.drop_key = site_key(d->seed, SITE_DEC_CROSS, L);
    EGOT2_TRY(small_attn_fwd(a, st));
  }
  // 6. y2 = x1 + dropout2(a2 . Wo^T + bo) ; x2 = norm2(y2)
  EGOT2_TRY(lin(dt, R, H, H, s->a2, H, p->ca_out_w, H, p->ca_out_b, s->y2, H, st, 0, pd, site_key(d->seed, SITE_DEC_DROP2, L), s->x1));
  {
    LayerNormArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y2; l.g = p->norm2_g; l.b = p->norm2_b; l.eps = d->ln_eps;
    l.y = s->x2; l.stat = s->stat2;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  // 7. feed-forward: y3 = x2 + dropout3(dropout(relu(x2 W1^T + b1)) W2^T + b2) ; y_out = norm3(y3)
  EGOT2_TRY(lin(dt, R, FF, H, s->x2, H, p->lin1_w, H, p->lin1_b, s->hid, FF, st, 1, pd, site_key(d->seed, SITE_DEC_FFN, L), nullptr, 1));
  EGOT2_TRY(lin(dt, R, H, FF, s->hid, FF, p->lin2_w, FF, p->lin2_b, s->y3, H, st, 0, pd, site_key(d->seed, SITE_DEC_DROP3, L), s->x2));
  {
    LayerNormArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y3; l.g = p->norm3_g; l.b = p->norm3_b; l.eps = d->ln_eps;
    l.y = y_out; l.stat = s->stat3;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  return 0;
}

extern "C" int egot2_decoder_layer_bwd(const egot2_decoder_desc* d, const egot2_decoder_params* p, const void* y_in,
                                       const void* mem, const egot2_decoder_saved* s, void* dy_out, void* dy_in,
                                       float* dmem, const egot2_decoder_grads* g, void* workspace, size_t ws_bytes,
                                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(dec_check(d));
  const int R = d->rows * d->S, H = d->H, FF = d->FF, dt = d->dtype;
  if (R == 0) return 0;
  DecWs w;
  const size_t need = dec_ws_layout(d, workspace, ws_bytes, &w);
  EGOT2_CHECK(workspace && ws_bytes + 256 >= need, "decoder_layer_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const size_t es = dtype_size(dt);
  const float pd = d->training ? d->p_drop : 0.f;
  const float inv_keep = pd > 0.f ? 1.f / (1.f - pd) : 1.f;
  const uint32_t L = (uint32_t)d->layer_index;

  // ---- norm3 -> d1 = dL/dy3 ; dm = dropout3-masked copy
  const void* dm = w.d1;
  {
    LayerNormBwdArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y3; l.stat = s->stat3; l.g = p->norm3_g;
    l.dy = dy_out; l.dx = w.d1; l.dg = g->norm3_g; l.db = g->norm3_b;
    if (pd > 0.f) { l.dx2 = w.dm; l.dx2_p_drop = pd; l.dx2_drop_key = site_key(d->seed, SITE_DEC_DROP3, L); dm = w.dm; }
    l.dcol = g->lin2_b;            // db2 = colsum(dm) comes out of the same pass (as in the encoder layer)
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  // ---- feed-forward
  EGOT2_TRY(wgrad2(dt, R, H, FF, dm, H, s->hid, FF, g->lin2_w, st));
  EGOT2_TRY(dgrad(dt, R, H, FF, dm, H, p->lin2_w, FF, w.dhid, FF, st, nullptr, s->hid, inv_keep));
  EGOT2_TRY(wgrad2(dt, R, FF, H, w.dhid, FF, s->x2, H, g->lin1_w, st));
  EGOT2_TRY(colsum_accum(dt, R, FF, w.dhid, FF, 0, 0, g->lin1_b, st));
  EGOT2_TRY(dgrad(dt, R, FF, H, w.dhid, FF, p->lin1_w, H, w.d2, H, st, w.d1));          // d2 = dL/dx2
  // ---- norm2 -> d1 = dL/dy2 ; dm = dropout2-masked
  dm = w.d1;
  {
    LayerNormBwdArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y2; l.stat = s->stat2; l.g = p->norm2_g;
    l.dy = w.d2; l.dx = w.d1; l.dg = g->norm2_g; l.db = g->norm2_b;
    if (pd > 0.f) { l.dx2 = w.dm; l.dx2_p_drop = pd; l.dx2_drop_key = site_key(d->seed, SITE_DEC_DROP2, L); dm = w.dm; }
    l.dcol = g->ca_out_b;
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  // ---- cross-attention
  EGOT2_TRY(wgrad2(dt, R, H, H, dm, H, s->a2, H, g->ca_out_w, st));
  EGOT2_TRY(dgrad(dt, R, H, H, dm, H, p->ca_out_w, H, w.da, H, st));                    // da = dL/da2
  // the three fp32 attention-gradient buffers are consecutive in the workspace: ONE clear for all of them
  EGOT2_TRY(zero_f32(w.dq32, (size_t)((w.dqkv32 + (size_t)R * 3 * H) - w.dq32), st));
  {
    SmallAttnArgs a; a.dtype = dt; a.rows = d->rows; a.S = d->S; a.M = d->M; a.H = H; a.heads = d->heads; a.causal = 0;
    a.q = s->qc; a.ldq = H; a.k = s->kvc; a.v = (const char*)s->kvc + (size_t)H * es; a.ldkv = 2 * H;
    a.kv_inner = d->kv_inner; a.kv_outer = d->kv_outer; a.kv_jstride = d->kv_jstride; a.kv_istride = d->kv_istride;
    a.ldo = H; a.p_drop = pd; a.drop_key = site_key(d->seed, SITE_DEC_CROSS, L);
    a.dout = w.da; a.dq = w.dq32; a.ld_dq = H; a.dk = w.dkv32; a.dv = w.dkv32 + H; a.ld_dkv = 2 * H;
    EGOT2_TRY(small_attn_bwd(a, st));
  }
  const void* dq = w.dq32; const void* dkv = w.dkv32;
  if (dt != EGOT2_F32) {       // bf16 copies for the GEMMs and the in_proj bias gradients (column sums) in the same pass
    EGOT2_TRY(colsum_cast_bf16(R, H, w.dq32, H, w.dq_lp, H, g->ca_in_b, st));
    EGOT2_TRY(colsum_cast_bf16(d->mem_rows, 2 * H, w.dkv32, 2 * H, w.dkv_lp, 2 * H, g->ca_in_b + H, st));
    dq = w.dq_lp; dkv = w.dkv_lp;
  } else {
    EGOT2_TRY(colsum_accum(EGOT2_F32, R, H, w.dq32, H, 0, 0, g->ca_in_b, st));
    EGOT2_TRY(colsum_accum(EGOT2_F32, d->mem_rows, 2 * H, w.dkv32, 2 * H, 0, 0, g->ca_in_b + H, st));
  }
  //   in_proj rows [0,H): queries (input x1) ; rows [H,3H): keys|values (input mem)
  EGOT2_TRY(wgrad2(dt, R, H, H, dq, H, s->x1, H, g->ca_in_w, st));
  EGOT2_TRY(wgrad2(dt, d->mem_rows, 2 * H, H, dkv, 2 * H, mem, H, g->ca_in_w + (size_t)H * H, st));
  if (dmem) {   // dmem += dkv . Wkv   (fp32, accumulated over the decoder layers)
    GemmArgs m; m.M = d->mem_rows; m.N = H; m.K = 2 * H; m.A = dkv; m.lda = 2 * H;
    m.B = (const char*)p->ca_in_w + (size_t)H * H * es; m.ldb = H; m.trans_b = 0; m.C = dmem; m.ldc = H;
    m.in_dtype = dt; m.out_dtype = EGOT2_F32; m.accumulate = 1;
    EGOT2_TRY(gemm(m, st));
  }
  EGOT2_TRY(dgrad(dt, R, H, H, dq, H, p->ca_in_w, H, w.d2, H, st, w.d1));               // d2 = dL/dx1
  // ---- norm1 -> d1 = dL/dy1 ; dm = dropout1-masked
  dm = w.d1;
  {
    LayerNormBwdArgs l; l.rows = R; l.H = H; l.dtype = dt; l.x = s->y1; l.stat = s->stat1; l.g = p->norm1_g;
    l.dy = w.d2; l.dx = w.d1; l.dg = g->norm1_g; l.db = g->norm1_b;
    if (pd > 0.f) { l.dx2 = w.dm; l.dx2_p_drop = pd; l.dx2_drop_key = site_key(d->seed, SITE_DEC_DROP1, L); dm = w.dm; }
    l.dcol = g->sa_out_b;
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  // ---- self-attention
  EGOT2_TRY(wgrad2(dt, R, H, H, dm, H, s->a1, H, g->sa_out_w, st));
  EGOT2_TRY(dgrad(dt, R, H, H, dm, H, p->sa_out_w, H, w.da, H, st));
  {
    SmallAttnArgs a; a.dtype = dt; a.rows = d->rows; a.S = d->S; a.M = d->S; a.H = H; a.heads = d->heads; a.causal = 1;
    a.q = s->qkv; a.ldq = 3 * H; a.k = (const char*)s->qkv + (size_t)H * es; a.v = (const char*)s->qkv + (size_t)2 * H * es;
    a.ldkv = 3 * H; a.kv_inner = 1; a.kv_outer = d->S; a.kv_jstride = 1; a.kv_istride = 0;
    a.ldo = H; a.p_drop = pd; a.drop_key = site_key(d->seed, SITE_DEC_SELF, L);
    a.dout = w.da; a.dq = w.dqkv32; a.ld_dq = 3 * H; a.dk = w.dqkv32 + H; a.dv = w.dqkv32 + 2 * H; a.ld_dkv = 3 * H;
    EGOT2_TRY(small_attn_bwd(a, st));
  }
  const void* dqkv = w.dqkv32;
  if (dt != EGOT2_F32) { EGOT2_TRY(colsum_cast_bf16(R, 3 * H, w.dqkv32, 3 * H, w.dqkv_lp, 3 * H, g->sa_in_b, st)); dqkv = w.dqkv_lp; }
  else EGOT2_TRY(colsum_accum(EGOT2_F32, R, 3 * H, w.dqkv32, 3 * H, 0, 0, g->sa_in_b, st));
  EGOT2_TRY(wgrad2(dt, R, 3 * H, H, dqkv, 3 * H, y_in, H, g->sa_in_w, st));
  EGOT2_TRY(dgrad(dt, R, 3 * H, H, dqkv, 3 * H, p->sa_in_w, H, dy_in, H, st, w.d1));
  return 0;
}

extern "C" int egot2_prompt_embed_fwd(int32_t dtype, int32_t rows, int32_t S, int32_t H, const int64_t* tokens,
                                      const float* embedding, const float* pe, float p_drop, int32_t training,
                                      uint64_t seed, void* y, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows * S == 0) return 0;
  const float p = training ? p_drop : 0.f;
  const uint64_t key = site_key(seed, SITE_PROMPT, 0);
  ProfScope prof(st, "prompt_embed_fwd rows%d S%d H%d", rows, S, H);
  if (dtype == EGOT2_F32) launch(prompt_embed_fwd_kernel<float>, dim3(rows * S), dim3(128), 0, st, rows, S, H, tokens, embedding, pe, p, key, (float*)y);
  else launch(prompt_embed_fwd_kernel<bf16>, dim3(rows * S), dim3(128), 0, st, rows, S, H, tokens, embedding, pe, p, key, (bf16*)y);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

extern "C" int egot2_prompt_embed_bwd(int32_t dtype, int32_t rows, int32_t S, int32_t H, const int64_t* tokens,
                                      const void* dy, float p_drop, int32_t training, uint64_t seed, float* d_embedding,
                                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows * S == 0) return 0;
  const float p = training ? p_drop : 0.f;
  const uint64_t key = site_key(seed, SITE_PROMPT, 0);
  ProfScope prof(st, "prompt_embed_bwd rows%d S%d H%d", rows, S, H);
  if (dtype == EGOT2_F32) launch(prompt_embed_bwd_kernel<float>, dim3(rows * S), dim3(128), 0, st, rows, S, H, tokens, (const float*)dy, p, key, d_embedding);
  else launch(prompt_embed_bwd_kernel<bf16>, dim3(rows * S), dim3(128), 0, st, rows, S, H, tokens, (const bf16*)dy, p, key, d_embedding);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
