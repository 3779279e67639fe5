// Factorizations (svd / eigh / qr) and the bond-level entry points built on them.
#include "tnb_internal.h"

using namespace tnb;
#define H ((Handle*)h)
#define ST ((cudaStream_t)stream)

extern "C" {

int tnb_svd_trunc(tnb_handle_t h, int dtype, int64_t m, int64_t n, void* A, int64_t maxdim, int64_t mindim,
                  double cutoff, int flags, int do_truncate, void* U, double* S, void* V, int64_t* n_keep,
                  double* truncerr, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "svd_trunc: not built yet");
}
int tnb_eigh_trunc(tnb_handle_t h, int dtype, int64_t n, void* A, int64_t maxdim, int64_t mindim, double cutoff,
                   int flags, int do_truncate, double* D, void* U, int64_t* n_keep, double* truncerr,
                   void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "eigh_trunc: not built yet");
}
int tnb_qr(tnb_handle_t h, int dtype, int64_t m, int64_t n, const void* A, void* Q, void* R, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "qr: not built yet");
}
int tnb_factorize_bond(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, void* phi, int ortho,
                       int which_decomp, int64_t maxdim, int64_t mindim, double cutoff, const void* rho_pert,
                       int normalize, void* A, void* B, int64_t* n_keep, double* truncerr, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "factorize_bond: not built yet");
}
int tnb_dmrg_bond_step(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L, const void* W1,
                       const void* W2, const void* R, void* A1, void* A2, int ortho, int which_decomp,
                       int64_t maxdim, int64_t mindim, double cutoff, double noise, int krylovdim, int maxiter,
                       double* energy, int64_t* n_keep, double* truncerr, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "dmrg_bond_step: not built yet");
}
int tnb_tebd_apply_gate(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiM, int64_t chiR, int32_t d1,
                        int32_t d2, const void* G, void* A1, void* A2, int64_t maxdim, int64_t mindim,
                        double cutoff, int64_t* n_keep, double* truncerr, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return set_err(H, TNB_ERR_UNSUPPORTED, "tebd_apply_gate: not built yet");
}
}
