// ComplexF64 instantiations of the contraction kernels (see contract_kernel.cuh / contract.cu).
#include "contract_kernel.cuh"

namespace tnb {
int launch_tiles_c128(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, bool small, cudaStream_t st) {
  return launch_tiles<true>(h, p, ak, bk, va, vb, small, st);
}
int launch_smallk_c128(Handle* h, GemmParams& p, cudaStream_t st) { return launch_smallk<true>(h, p, st); }
}  // namespace tnb
