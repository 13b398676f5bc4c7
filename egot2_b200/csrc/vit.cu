// vit.cu — one layer of the simple_vit `Transformer` used by the HOI PNR "simple_vit" translator siblings
// (HOI/models/pnr/simple_vit.py:55-107; HOI/models/pnr/video_model_transfer_3task.py:128-164):
//     x1 = to_out( softmax(q k^T * dim_head^-0.5) v ) + x      with [q|k|v] = to_qkv(LayerNorm_a(x))   (no biases)
//     x2 = W2 gelu(W1 LayerNorm_f(x1) + b1) + b2 + x1                                                (exact erf GELU)
// PRE-norm, no dropout, no final norm, dim_head independent of the model width (8 x 128 on a 256-wide model).
// Composition of the library's token-wise GEMMs (tcgen05 when the shape qualifies), LayerNorm and attention launchers
// plus the GELU kernels below; correctness-first like the EgoT2-g decoder (eager launches, nothing fused yet).
#include <math.h>

#define EGOT2_FILE_ID 12
#include "ops.h"

namespace egot2 {
namespace {

struct Carver {
  char* base; size_t size, off = 0;
  Carver(void* p, size_t n) : base((char*)p), size(n) {}
  void* take(size_t bytes) { void* r = base ? base + off : nullptr; off += align_up(bytes); return r; }
};

struct VitWs {
  void *du, *dh, *d1, *dout, *dqkv, *attn_ws;
  size_t attn_ws_bytes;
};
size_t vit_ws_layout(const egot2_vit_desc* d, void* base, size_t bytes, VitWs* w) {
  Carver c(base, bytes);
  const size_t es = dtype_size(d->dtype), M = (size_t)d->B * d->T, D = d->D, inner = (size_t)d->heads * d->dim_head;
  VitWs t;
  t.du = c.take(M * d->mlp * es);
  t.dh = c.take(M * D * es);
  t.d1 = c.take(M * D * es);
  t.dout = c.take(M * inner * es);
  t.dqkv = c.take(M * 3 * inner * es);
  t.attn_ws_bytes = attention_bwd_workspace(d->dtype, d->B, d->T, (int)inner, d->heads);
  t.attn_ws = c.take(t.attn_ws_bytes);
  if (w) *w = t;
  return c.off + 256;
}

int vit_check(const egot2_vit_desc* d) {
  EGOT2_CHECK(d->dtype == EGOT2_F32 || d->dtype == EGOT2_BF16, "vit: bad dtype %d", d->dtype);
  EGOT2_CHECK(d->B >= 0 && d->T >= 1 && d->D >= 1 && d->mlp >= 1, "vit: bad geometry B=%d T=%d D=%d mlp=%d", d->B, d->T, d->D, d->mlp);
  EGOT2_CHECK(d->heads >= 1 && d->dim_head >= 1, "vit: bad heads=%d dim_head=%d", d->heads, d->dim_head);
  return 0;
}

// C = A . W^T (+ bias) (+ residual)
int lin(int dt, int M, int N, int K, const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc,
        cudaStream_t st, const void* res = nullptr) {
  GemmArgs g; g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.trans_b = 1; g.C = C; g.ldc = ldc;
  g.bias = bias; g.residual = res; g.ldr = ldc; g.in_dtype = dt; g.out_dtype = dt;
  return gemm(g, st);
}
// dX = dY . W      dY (M,n_out) ; W (n_out, k_in) row-major
int dgrad(int dt, int M, int n_out, int k_in, const void* dY, int ld_dy, const void* W, int ldw, void* dX, int ldx,
          cudaStream_t st) {
  GemmArgs g; g.M = M; g.N = k_in; g.K = n_out; g.A = dY; g.lda = ld_dy; g.B = W; g.ldb = ldw; g.trans_b = 0; g.C = dX; g.ldc = ldx;
  g.in_dtype = dt; g.out_dtype = dt;
  return gemm(g, st);
}
// dW (n_out, k_in) += dY^T . X
int wgrad2(int dt, int rows, int n_out, int k_in, const void* dY, int ld_dy, const void* X, int ld_x, float* dW,
           cudaStream_t st) {
  GemmArgs g; g.M = n_out; g.N = k_in; g.K = rows; g.A = dY; g.lda = ld_dy; g.trans_a = 1; g.B = X; g.ldb = ld_x; g.trans_b = 0;
  g.C = dW; g.ldc = k_in; g.in_dtype = dt; g.out_dtype = EGOT2_F32; g.accumulate = 1;
  g.split_k = suggest_split_k(g.M, g.N, g.K);
  return gemm(g, st);
}

// exact GELU (torch.nn.GELU default, approximate='none'):  gelu(u) = u * Phi(u),  Phi(u) = (1 + erf(u / sqrt 2)) / 2
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
// d/du = Phi(u) + u * phi(u),  phi(u) = exp(-u^2 / 2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_grad_f(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.39894228040143268f * expf(-0.5f * u * u);
}

template <typename T>
__global__ void gelu_fwd_kernel(const T* __restrict__ u, T* __restrict__ a, size_t n) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a[i] = from_f32<T>(gelu_f(to_f32(u[i])));
}
// du = da * gelu'(u), in place on the gradient buffer
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ u, T* __restrict__ d, size_t n) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = from_f32<T>(to_f32(d[i]) * gelu_grad_f(to_f32(u[i])));
}

int ew_grid(size_t n) {
  size_t ctas = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (ctas > cap) ctas = cap;
  return ctas ? (int)ctas : 1;
}
int gelu_fwd(int dt, const void* u, void* a, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ProfScope prof(st, "gelu_fwd n%zu", n);
  if (dt == EGOT2_F32) launch(gelu_fwd_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (const float*)u, (float*)a, n);
  else launch(gelu_fwd_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (const bf16*)u, (bf16*)a, n);
  EGOT2_LAUNCH_CHECK();
  return 0;
}
int gelu_bwd(int dt, const void* u, void* d, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  ProfScope prof(st, "gelu_bwd n%zu", n);
  if (dt == EGOT2_F32) launch(gelu_bwd_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (const float*)u, (float*)d, n);
  else launch(gelu_bwd_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (const bf16*)u, (bf16*)d, n);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace egot2

using namespace egot2;

extern "C" size_t egot2_vit_layer_workspace_bytes(const egot2_vit_desc* d) { return vit_ws_layout(d, nullptr, 0, nullptr); }

extern "C" int egot2_vit_layer_fwd(const egot2_vit_desc* d, const egot2_vit_params* p, const void* x_in, void* x_out,
                                   const egot2_vit_saved* s, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(vit_check(d));
  const int M = d->B * d->T, D = d->D, inner = d->heads * d->dim_head, mlp = d->mlp, dt = d->dtype;
  if (M == 0) return 0;
  // 1. h = LayerNorm_a(x)
  {
    LayerNormArgs l; l.rows = M; l.H = D; l.dtype = dt; l.x = x_in; l.g = p->norm_a_g; l.b = p->norm_a_b; l.eps = d->ln_eps;
    l.y = s->h; l.stat = s->stat_a;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  // 2. [q|k|v] = h . Wqkv^T  (bias-free; head j of q/k/v = columns j*dim_head.. of its third, 'b n (h d) -> b h n d')
  EGOT2_TRY(lin(dt, M, 3 * inner, D, s->h, D, p->qkv_w, D, nullptr, s->qkv, 3 * inner, st));
  // 3. per clip and head: softmax(q k^T * dim_head^-0.5) v  (the attention kernels scale by 1/sqrt(head dim))
  EGOT2_TRY(attention_fwd(dt, d->B, d->T, inner, d->heads, s->qkv, s->attn, s->lse, 0.f, 0, st));
  // 4. x1 = attn . Wout^T + x
  EGOT2_TRY(lin(dt, M, D, inner, s->attn, inner, p->out_w, inner, nullptr, s->x1, D, st, x_in));
  // 5. h2 = LayerNorm_f(x1)
  {
    LayerNormArgs l; l.rows = M; l.H = D; l.dtype = dt; l.x = s->x1; l.g = p->norm_f_g; l.b = p->norm_f_b; l.eps = d->ln_eps;
    l.y = s->h2; l.stat = s->stat_f;
    EGOT2_TRY(layernorm_fwd(l, st));
  }
  // 6. u = h2 . W1^T + b1 ; act = gelu(u)
  EGOT2_TRY(lin(dt, M, mlp, D, s->h2, D, p->ff1_w, D, p->ff1_b, s->u, mlp, st));
  EGOT2_TRY(gelu_fwd(dt, s->u, s->act, (size_t)M * mlp, st));
  // 7. x_out = act . W2^T + b2 + x1
  return lin(dt, M, D, mlp, s->act, mlp, p->ff2_w, mlp, p->ff2_b, x_out, D, st, s->x1);
}

extern "C" int egot2_vit_layer_bwd(const egot2_vit_desc* d, const egot2_vit_params* p, const void* x_in,
                                   const egot2_vit_saved* s, const void* dx_out, void* dx_in, const egot2_vit_grads* g,
                                   void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EGOT2_TRY(vit_check(d));
  const int M = d->B * d->T, D = d->D, inner = d->heads * d->dim_head, mlp = d->mlp, dt = d->dtype;
  if (M == 0) return 0;
  VitWs w;
  const size_t need = vit_ws_layout(d, workspace, ws_bytes, &w);
  EGOT2_CHECK(workspace && ws_bytes + 256 >= need, "vit_layer_bwd: workspace too small (%zu < %zu)", ws_bytes, need);

  // ---- feed-forward: x2 = act W2^T + b2 + x1
  EGOT2_TRY(wgrad2(dt, M, D, mlp, dx_out, D, s->act, mlp, g->ff2_w, st));
  EGOT2_TRY(colsum_accum(dt, M, D, dx_out, D, 0, 0, g->ff2_b, st));
  EGOT2_TRY(dgrad(dt, M, D, mlp, dx_out, D, p->ff2_w, mlp, w.du, mlp, st));           // d(act)
  EGOT2_TRY(gelu_bwd(dt, s->u, w.du, (size_t)M * mlp, st));                           // d(u)
  EGOT2_TRY(wgrad2(dt, M, mlp, D, w.du, mlp, s->h2, D, g->ff1_w, st));
  EGOT2_TRY(colsum_accum(dt, M, mlp, w.du, mlp, 0, 0, g->ff1_b, st));
  EGOT2_TRY(dgrad(dt, M, mlp, D, w.du, mlp, p->ff1_w, D, w.dh, D, st));               // d(h2)
  {   // d1 = dL/dx1 = LayerNorm_f'(d(h2)) + dx_out (the residual branch)
    LayerNormBwdArgs l; l.rows = M; l.H = D; l.dtype = dt; l.x = s->x1; l.stat = s->stat_f; l.g = p->norm_f_g;
    l.dy = w.dh; l.dx = w.d1; l.dres = dx_out; l.dg = g->norm_f_g; l.db = g->norm_f_b;
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  // ---- attention: x1 = attn Wout^T + x
  EGOT2_TRY(wgrad2(dt, M, D, inner, w.d1, D, s->attn, inner, g->out_w, st));
  EGOT2_TRY(dgrad(dt, M, D, inner, w.d1, D, p->out_w, inner, w.dout, inner, st));     // d(attn)
  EGOT2_TRY(attention_bwd(dt, d->B, d->T, inner, d->heads, s->qkv, s->attn, s->lse, w.dout, w.dqkv, 0.f, 0, w.attn_ws,
                          w.attn_ws_bytes, st));
  EGOT2_TRY(wgrad2(dt, M, 3 * inner, D, w.dqkv, 3 * inner, s->h, D, g->qkv_w, st));
  EGOT2_TRY(dgrad(dt, M, 3 * inner, D, w.dqkv, 3 * inner, p->qkv_w, D, w.dh, D, st)); // d(h)
  {   // dx_in = LayerNorm_a'(d(h)) + d1
    LayerNormBwdArgs l; l.rows = M; l.H = D; l.dtype = dt; l.x = x_in; l.stat = s->stat_a; l.g = p->norm_a_g;
    l.dy = w.dh; l.dx = dx_in; l.dres = w.d1; l.dg = g->norm_a_g; l.db = g->norm_a_b;
    EGOT2_TRY(layernorm_bwd(l, st));
  }
  return 0;
}
