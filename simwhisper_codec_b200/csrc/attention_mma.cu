// placeholder wiring: replaced by the tensor-core kernel (see attention_mma.cu history)
#include "kernels.cuh"
namespace swc {
int attention_mma(const bf16* qkv, bf16* out, const long long* lens, int nb, int T, int H, cudaStream_t s) {
  return attention_simt(qkv, 1, out, lens, nb, T, H, s);
}
}  // namespace swc
