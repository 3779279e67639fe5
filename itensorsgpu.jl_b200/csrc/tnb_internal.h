// Internal declarations shared by the translation units of libtnb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/tnb200.h"

namespace tnb {

constexpr int MAXG = 12;  // modes per group (M / N / K) after merging
constexpr int TNB_MAX_PEERS = 8;   // GPUs of one NVSwitch node

// One mode group of a contraction, linearised in mixed radix (ext[0] fastest).
// sX / sY are the element strides of each mode in the two tensors that carry the group:
//   M group: X = A, Y = C      N group: X = B, Y = C      K group: X = A, Y = B
struct Group {
  int n;
  int ext[MAXG];
  long long sX[MAXG];
  long long sY[MAXG];
};

struct GemmParams {
  Group gm, gn, gk;
  int M, N, K;
  const void* A;
  const void* B;
  void* C;
  double alpha_re, alpha_im, beta_re, beta_im;
  int conjA, conjB;
  int tilesM, tilesN, groupM;
  // batched form: blockIdx.y = batch; element offsets added to A/B/C (device arrays or strides)
  int batch;
  const long long* boffA;
  const long long* boffB;
  const long long* boffC;
  long long bstrideA, bstrideB, bstrideC;   // used when the corresponding boff* is null
  // split output: columns n >= splitN of batch b go to C + boffC2[b] (column index n - splitN)
  int splitN;
  const long long* boffC2;
  // lowerOnly: C is Hermitian and only its lower triangle (m >= n) is needed: CTAs whose tile lies strictly above
  // the diagonal exit at once (syr2k-style trailing update of the tridiagonalisation)
  int lowerOnly;
  // fused all-gather: when npeer > 0 the epilogue stores to every peerC[g] (same element offsets) instead of C
  int npeer;
  void* peerC[TNB_MAX_PEERS];
};

// Peer group of one NVSwitch node (one process per GPU): flag arrays of every rank mapped through CUDA IPC.  The
// library numbers its own barrier epochs; every rank must issue the same sequence of collective calls.
struct Comm {
  bool on = false;
  int rank = 0, world = 1;
  unsigned long long* flags[TNB_MAX_PEERS] = {};
  unsigned long long epoch = 0;
};

// Sharded matvec context of the Lanczos solver: this rank owns the slab L[:, l'_shard, :] (clp columns) and the
// result of matvec number i lands, through the fused all-gather, in out[i & 1][g] of every rank g.
struct ShardCtx {
  int64_t clp;
  void* const* out[2];
};

struct Handle {
  int device = 0;
  int num_sms = 148;
  std::string err;
  char* ws = nullptr;       // workspace arena
  size_t ws_bytes = 0;
  size_t ws_off = 0;        // bump pointer (reset by each public entry point)
  size_t ws_base = 0;       // bytes at the front of the arena pinned by an outer entry point
  double* scal = nullptr;   // small device scalar pool (256 doubles)
  double* scal_host = nullptr;  // pinned mirror
  uint64_t launches = 0;
  cudaStream_t copy_stream = nullptr;
  double* partials = nullptr;   // per-block partial sums for reductions (8192 doubles)
  unsigned* counter = nullptr;  // last-block-done ticket (self-resetting)
  void* what = nullptr;         // combined two-site MPO matrix of the fused H_eff step 2+3 (64 KB)
  cudaEvent_t ev[40] = {};      // chunk hand-over events of the pipelined host-buffer H_eff (created on first use)
  Comm comm;                    // peer group (tnb_comm_init)
};
constexpr int RED_MAX_BLOCKS = 1024;

int set_err(Handle* h, int code, const char* fmt, ...);
int check_cuda(Handle* h, cudaError_t e, const char* what);
// Bump-allocate from the arena (256-byte aligned).  Grows the arena if needed (device sync).
int ws_alloc(Handle* h, size_t bytes, void** out);
// Ensure the arena holds at least `bytes` BEFORE a sequence of ws_alloc calls, so that
// pointers handed out earlier in the same entry point stay valid.
int ws_require(Handle* h, size_t bytes);
inline void ws_reset(Handle* h) { h->ws_off = h->ws_base; }

#define TNB_CUDA(h, call)                                   \
  do {                                                      \
    int _s = tnb::check_cuda((h), (call), #call);           \
    if (_s) return _s;                                      \
  } while (0)
#define TNB_TRY(call)        \
  do {                       \
    int _s = (call);         \
    if (_s) return _s;       \
  } while (0)

// ---- contraction (contract.cu)
int contract_impl(Handle* h, int dtype, int nA, const int64_t* extA, const int32_t* modeA,
                  const void* A, int nB, const int64_t* extB, const int32_t* modeB,
                  const void* B, int nC, const int64_t* extC, const int32_t* modeC, void* C,
                  const void* alpha, const void* beta, int flags, cudaStream_t st);
int contract_impl_ex(Handle* h, int dtype, int nA, const int64_t* extA, const int32_t* modeA,
                     const void* A, int nB, const int64_t* extB, const int32_t* modeB,
                     const void* B, int nC, const int64_t* extC, const int32_t* modeC, void* C,
                     const void* alpha, const void* beta, int flags, cudaStream_t st,
                     const int64_t* strideC, void* const* peerC, int npeer, const int64_t* strideA = nullptr,
                     const int64_t* strideB = nullptr);
// plain column-major GEMM helper built on the same kernel:
//   C[m x n] (ldc) <- alpha * op(A) * op(B) + beta * C ; op = N / T / C(onj-transpose)
int gemm_impl(Handle* h, int dtype, char opA, char opB, int64_t m, int64_t n, int64_t k,
              const void* alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
              const void* beta, void* C, int64_t ldc, cudaStream_t st, int lower_only = 0);

// batched column-major GEMM: for b in [0,batch): C_b = alpha*op(A_b)*op(B_b) + beta*C_b where
// X_b = X + (offX ? offX[b] : b*strideX) elements.  offX are DEVICE arrays.
int gemm_batched_impl(Handle* h, int dtype, char opA, char opB, int64_t m, int64_t n, int64_t k,
                      const void* alpha, const void* A, int64_t lda, const long long* offA, long long strideA,
                      const void* B, int64_t ldb, const long long* offB, long long strideB, const void* beta,
                      void* C, int64_t ldc, const long long* offC, long long strideC, int batch,
                      cudaStream_t st, int splitN = 0, const long long* offC2 = nullptr);

// ---- vector ops (vecops.cu)
int permute_axpby_impl(Handle* h, int dtype, int n, const int64_t* extA, const int32_t* modeA,
                       const void* A, const int32_t* modeB, void* B, const void* alpha,
                       const void* beta, cudaStream_t st);
int scale_impl(Handle* h, int dtype, int64_t n, void* x, const void* alpha, cudaStream_t st);
int diag_contract_impl(Handle* h, int dtype, int n, const int64_t* extA, const int32_t* modeA, const void* A,
                       int32_t scaled_mode, const void* diag, int diag_dtype, const int32_t* modeC, void* C,
                       cudaStream_t st);
int dot_impl(Handle* h, int dtype, int64_t n, const void* x, const void* y, void* result_dev,
             cudaStream_t st);
int nrm2_impl(Handle* h, int dtype, int64_t n, const void* x, double* result_dev,
              cudaStream_t st);
int truncate_impl(Handle* h, const double* P_dev, int64_t len, int64_t maxdim, int64_t mindim,
                  double cutoff, int flags, int64_t* n_keep, double* truncerr, double* docut,
                  cudaStream_t st);

// ---- fused DMRG pieces (heff.cu)
size_t heff_workspace_bytes(int dtype, const tnb_bond_dims* d);
int heff_core_pub(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                  const void* W2, const void* R, const void* phi, void* out, void* t0, void* t1,
                  cudaStream_t st);
int heff_apply_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                    const void* W2, const void* R, const void* phi, void* out, cudaStream_t st);
int env_update_impl(Handle* h, int dtype, bool left, int64_t cl, int64_t cr, int32_t d, int32_t wl,
                    int32_t wr, const void* E, const void* A, const void* W, void* Enew,
                    cudaStream_t st);
int lanczos_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                 const void* W2, const void* R, void* phi, int krylovdim, int maxiter, double tol,
                 double* energy, int* n_matvec, cudaStream_t st, const ShardCtx* sc = nullptr);
// sharded H_eff*phi with the all-gather fused into step 4 (peer stores), no barrier
int heff_shard_fused_core(Handle* h, int dtype, const tnb_bond_dims* d, int rank, int world, int64_t clp,
                          const void* Lslab, const void* W1, const void* W2, const void* R, const void* phi,
                          void* const* out_peers, void* t0, void* t1, cudaStream_t st);
size_t heff_shard_ws_bytes(int dtype, const tnb_bond_dims* d, int64_t clp);
// device-side barrier over the handle's peer group (stream-ordered, asynchronous)
int comm_barrier(Handle* h, cudaStream_t st);
// barrier; region [off + rank*bytes, +bytes) of bufs[rank] -> same region of every peer buffer; barrier
int comm_allgather(Handle* h, void* const* bufs, size_t off, size_t bytes_per_rank, cudaStream_t st);
int factorize_core_pub(Handle* h, int dtype, int64_t m, int64_t n, void* M, int ortho, int which, int64_t maxdim,
                       int64_t mindim, double cutoff, const void* rho_pert, int normalize, void* A, void* B,
                       int64_t* n_keep, double* truncerr, cudaStream_t st);
size_t factorize_ws_bytes_pub(int dtype, int64_t m, int64_t n);
int noise_term_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                    const void* W2, const void* R, const void* phi, int ortho, double noise,
                    int accumulate, void* rho, void* t0, void* t1, cudaStream_t st);

// y <- x / *scal (0 if *scal <= tol); x and y may alias (vecops.cu)
int scale_inv_dev_impl(Handle* h, int dtype, int64_t n, const void* x, void* y, const double* scal, double tol,
                       cudaStream_t st);

// ---- factorizations (jacobi.cu, qr.cu)
size_t svd_ws_bytes(int dtype, int64_t m, int64_t n);
size_t eigh_ws_bytes(int dtype, int64_t n);
size_t qr_ws_bytes(int dtype, int64_t m, int64_t n);
// U: m x kmax, V: n x kmax (A ~ U diag(S) V^T); S receives the first ks singular values (ks <= min(m,n))
int svd_impl(Handle* h, int dtype, int64_t m, int64_t n, const void* A, int64_t lda, int64_t kmax, int64_t ks,
             void* U, int64_t ldu, double* S, void* V, int64_t ldv, cudaStream_t st);
// eigenvalues descending; D receives the first ks, U (n x kmax) the first kmax eigenvectors; A destroyed
int eigh_impl(Handle* h, int dtype, int64_t n, void* A, int64_t kmax, int64_t ks, double* D, void* U, int64_t ldu,
              cudaStream_t st);
int qr_impl(Handle* h, int dtype, int64_t m, int64_t n, const void* A, void* Q, void* R, cudaStream_t st);

// cudaFuncSetAttribute applies to the CURRENT device only: remember per device what has been set, so that a process
// owning one handle per GPU (julia: CUDA.device!; python: tn.handle() after torch.cuda.set_device) works on all of them.
constexpr int TNB_MAX_DEVICES = 64;
#define TNB_ONCE_PER_DEVICE(h, stmt)                          \
  do {                                                        \
    static bool _done[tnb::TNB_MAX_DEVICES] = {};             \
    const int _dv = (h)->device & (tnb::TNB_MAX_DEVICES - 1); \
    if (!_done[_dv]) {                                        \
      stmt;                                                   \
      _done[_dv] = true;                                      \
    }                                                         \
  } while (0)

inline size_t elsize(int dtype) { return dtype == TNB_C128 ? 16 : 8; }

}  // namespace tnb
