// embed_extra.cu — elementwise variants of the embed stage used by sibling translators only:
//   add_table:               tokens = projected features + positional table, NO LayerNorm
//                            (2-task simple_vit translator, HOI/models/pnr/video_model_transfer.py:63)
//   dropout_prefix_inplace:  feature dropout on the leading tokens of every clip only
//                            (2-task PNR translator with FEAT_DROPOUT_MODE > 0 drops the PNR segment alone, :95-96)
#define EGOT2_FILE_ID 14
#include "ops.h"

namespace egot2 {
namespace {

template <typename T>
__global__ void add_table_kernel(const T* __restrict__ z, const float* __restrict__ table, T* __restrict__ x, size_t n,
                                 size_t table_elems) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = from_f32<T>(to_f32(z[i]) + table[i % table_elems]);
}

template <typename T>
__global__ void dropout_prefix_kernel(T* __restrict__ x, size_t n, size_t period, size_t prefix, float p, float inv_keep,
                                      uint64_t key) {
  EGOT2_PDL_ENTER();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (i % period < prefix) x[i] = from_f32<T>(to_f32(x[i]) * drop_scale(key ^ egot2_ep, i, p, inv_keep));
}

int ew_grid(size_t n) {
  size_t ctas = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (ctas > cap) ctas = cap;
  return ctas ? (int)ctas : 1;
}

}  // namespace

int add_table(int dt, size_t n, size_t table_elems, const void* z, const float* table, void* x, cudaStream_t st) {
  if (n == 0) return 0;
  EGOT2_CHECK(table_elems > 0 && table, "add_table: empty table");
  ProfScope prof(st, "add_table n%zu", n);
  if (dt == EGOT2_F32) launch(add_table_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (const float*)z, table, (float*)x, n, table_elems);
  else launch(add_table_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (const bf16*)z, table, (bf16*)x, n, table_elems);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

int dropout_prefix_inplace(int dt, void* x, size_t n, size_t period, size_t prefix, float p, uint64_t key, cudaStream_t st) {
  if (p <= 0.f || n == 0 || prefix == 0) return 0;
  EGOT2_CHECK(period > 0 && prefix <= period && p < 1.f, "dropout_prefix: bad geometry (period %zu prefix %zu p %f)", period, prefix, p);
  const float inv_keep = 1.f / (1.f - p);
  ProfScope prof(st, "dropout_prefix n%zu", n);
  if (dt == EGOT2_F32) launch(dropout_prefix_kernel<float>, dim3(ew_grid(n)), dim3(256), 0, st, (float*)x, n, period, prefix, p, inv_keep, key);
  else launch(dropout_prefix_kernel<bf16>, dim3(ew_grid(n)), dim3(256), 0, st, (bf16*)x, n, period, prefix, p, inv_keep, key);
  EGOT2_LAUNCH_CHECK();
  return 0;
}

}  // namespace egot2
