// Float64 instantiations of the contraction kernels (see contract_kernel.cuh / contract.cu).
#include "contract_kernel.cuh"

namespace tnb {
int launch_tiles_f64(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, bool small, cudaStream_t st) {
  return launch_tiles<false>(h, p, ak, bk, va, vb, small, st);
}
int launch_smallk_f64(Handle* h, GemmParams& p, cudaStream_t st) { return launch_smallk<false>(h, p, st); }
}  // namespace tnb
