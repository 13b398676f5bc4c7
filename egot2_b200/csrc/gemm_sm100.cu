// gemm_sm100.cu — tcgen05/TMEM/TMA GEMM for bf16 operands (placeholder until the kernel lands).
#include "ops.h"
namespace egot2 {
int gemm_sm100(const GemmArgs& a, cudaStream_t st) { (void)a; (void)st; return -1; }
}  // namespace egot2
