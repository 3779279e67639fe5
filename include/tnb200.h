/*
 * tnb200.h -- C ABI of libtnb200.so: the B200-native (sm_100a) replacement for the
 * ITensorsGPU.jl hot path (dense ITensor contraction, H_eff*psi + eigensolver, truncated
 * svd/eigen in replacebond!/factorize, TEBD gate step).
 *
 * The reference has no FFI of its own: it is Julia methods dispatched on `CuDense` storage
 * that forward to cuTENSOR / cuBLAS / cuSOLVER (override surface = the import list at
 * /root/reference/src/ITensorsGPU.jl:32-43).  Every entry point below names the reference
 * method body it replaces (file:line relative to /root/reference).  INTEGRATION.md shows
 * the Julia `ccall` stub for each one.
 *
 * Conventions
 *   - Tensors are flat, dense, COLUMN-MAJOR device buffers over their mode order
 *     (src/tensor/cudense.jl:252-254); complex is interleaved (re,im) = Julia ComplexF64.
 *   - Modes are small non-negative ints; the same label in two operands of a contraction
 *     means "contract", exactly like the mode numbering at src/tensor/cudense.jl:258-283.
 *   - Caller owns every tensor buffer.  The library owns only the opaque handle
 *     (workspace arena, plan cache -- the analogue of `ContractionPlans`,
 *     src/ITensorsGPU.jl:54-55).  One handle per host thread AND per device (tnb_create binds to the device that is
 *     current when it is called); not thread-safe.  The workspace arena and the device scalar pool are reused in
 *     stream order: issue the calls of one handle on one stream at a time (a second stream needs a second handle),
 *     exactly like the reference's single task-local CUDA.jl stream.
 *   - All calls are asynchronous on `stream` unless they return a host scalar
 *     (documented per function).  `stream` is a cudaStream_t passed as void*.
 *   - Return value: 0 on success, a tnb_status otherwise; tnb_last_error() gives text.
 *     Nothing here aborts the process and nothing falls back to the CPU.
 */
#ifndef TNB200_H
#define TNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnb_handle_s* tnb_handle_t;

typedef enum {
  TNB_OK = 0,
  TNB_ERR_BAD_ARG = 1,        /* Julia side: ArgumentError      (src/mps/cumpo.jl:21)        */
  TNB_ERR_DIM_MISMATCH = 2,   /* Julia side: DimensionMismatch  (src/cuitensor.jl:55,72,78)  */
  TNB_ERR_UNSUPPORTED = 3,    /* Julia side: error(...)         (src/tensor/cudense.jl:165)  */
  TNB_ERR_CUDA = 4,
  TNB_ERR_NO_CONVERGENCE = 5,
  TNB_ERR_ALLOC = 6
} tnb_status;

typedef enum { TNB_F64 = 0, TNB_C128 = 1 } tnb_dtype;

/* flags for tnb_contract */
#define TNB_CONJ_A 1
#define TNB_CONJ_B 2
#define TNB_HERM_UPPER 4   /* the result, matricised as (A-free modes) x (B-free modes), is Hermitian (Gram matrix):
                              only its upper triangle is computed (tile granularity); the strictly lower part of C is
                              left untouched.  tnb_eigh_trunc reads the upper triangle only. */

/* flags for truncation (kwargs of truncate!, src/tensor/cutruncate.jl:3-7) */
#define TNB_TRUNC_ABSOLUTE_CUTOFF 1   /* absoluteCutoff / use_absolute_cutoff */
#define TNB_TRUNC_NO_RELATIVE     2   /* doRelCutoff=false / use_relative_cutoff=false */

/* which_decomp for tnb_factorize ([EXT] ITensors factorize) */
#define TNB_DECOMP_AUTO  0
#define TNB_DECOMP_SVD   1
#define TNB_DECOMP_EIGEN 2
#define TNB_DECOMP_QR    3

#define TNB_ORTHO_LEFT  0
#define TNB_ORTHO_RIGHT 1

/* ------------------------------------------------------------------ handle ------------ */
int tnb_create(tnb_handle_t* out);            /* binds to the current CUDA device */
int tnb_destroy(tnb_handle_t h);
const char* tnb_last_error(tnb_handle_t h);
int tnb_version(void);
/* Pre-size the workspace arena (bytes).  Optional; the arena grows on demand (growing
 * synchronises the device). */
int tnb_reserve(tnb_handle_t h, size_t bytes);
size_t tnb_workspace_bytes(tnb_handle_t h);
/* Plan cache of tnb_contract -- the analogue of the reference's `ContractionPlans` dictionary and cuTENSOR autotune
 * (src/ITensorsGPU.jl:54-55, src/tensor/cudense.jl:285-326).  Key: dtype, flags, (mode, extent, stride) of all three
 * operands, base-pointer alignment; value: the grouped GEMM parameters and the kernel variant.  A hit skips planning.
 * With autotune on (default) a NEW shape of at least 2e9 flop with beta = 0 is timed once with each tile configuration
 * on the caller's operands (the first call of such a shape synchronises); every candidate sums each output element in
 * the same k order, so the choice never changes a bit of the result.  mode: 0 = heuristic only, 1 = autotune. */
int tnb_plan_cache_stats(tnb_handle_t h, uint64_t* entries, uint64_t* hits, uint64_t* misses, uint64_t* autotuned);
int tnb_plan_cache_clear(tnb_handle_t h);
int tnb_set_autotune(tnb_handle_t h, int mode);
/* Contraction calls per kernel family since the library was loaded: out4[0] = LDGSTS tile kernel, out4[1] = small-K
 * streaming kernel, out4[2] = TMA-staged kernel (both operands K-major; contract_tma.cu), out4[3] = those TMA calls
 * that split K over thread-block clusters (problems smaller than one wave of CTA slots). */
int tnb_kernel_family_counts(tnb_handle_t h, uint64_t* out4);
/* Upper bound (bytes) on the PAIR of temporaries of one H_eff*phi / noise term / environment update (default 40 GB;
 * 0 restores the default).  Above it the work is cut into slabs of the output bond (H_eff: independent
 * full-efficiency slabs, L taken as a strided window, result written as a strided window) or of a summed bond with
 * beta = 1 accumulation (environment updates, noise term), so that C5 (chi = 8192, MPO bond 30: 2 x 64 GB unchunked)
 * runs in a fixed workspace.  Supersedes the reference's dead out-of-core attempt (src/tensor/dense.jl:50-193).
 * Process-wide setting. */
int tnb_set_workspace_limit(tnb_handle_t h, size_t bytes);
size_t tnb_get_workspace_limit(tnb_handle_t h);
/* Workspace queries -- the `*_bufferSize` calls of this library (the reference gets these implicitly from cuSOLVER /
 * cuTENSOR, e.g. inside CUSOLVER.svd!/syevd! at src/tensor/culinearalgebra.jl:42,85-87 and the cuTENSOR workspace of
 * src/tensor/cudense.jl:328).  Pure host arithmetic: no handle, no GPU.  They return the arena bytes the entry point
 * will require, so a caller can tnb_reserve() once up front (growing the arena later synchronises the device) and plan
 * HBM for C5-sized bonds; 0 = bad argument.  They honour tnb_set_workspace_limit (whose handle may be NULL).
 *   tnb_bond_workspace_bytes: op = TNB_WS_HEFF_APPLY (tnb_heff_apply, tnb_noise_term), TNB_WS_FACTORIZE_BOND
 *     (tnb_factorize_bond, tnb_tebd_apply_gate), TNB_WS_DMRG_BOND_STEP (tnb_dmrg_bond_step: ortho, with_noise and
 *     krylovdim matter only here); *n_slabs (optional) = number of output-bond slabs H_eff is cut into under the limit.
 *   tnb_matrix_workspace_bytes: op = TNB_WS_SVD (tnb_svd_trunc), TNB_WS_EIGH (tnb_eigh_trunc, m = n), TNB_WS_QR (tnb_qr). */
#define TNB_WS_HEFF_APPLY      0
#define TNB_WS_FACTORIZE_BOND  1
#define TNB_WS_DMRG_BOND_STEP  2
#define TNB_WS_SVD             3
#define TNB_WS_EIGH            4
#define TNB_WS_QR              5
struct tnb_bond_dims_s;
size_t tnb_bond_workspace_bytes(int op, int dtype, const struct tnb_bond_dims_s* dims, int ortho, int with_noise,
                                int krylovdim, int32_t* n_slabs);
size_t tnb_matrix_workspace_bytes(int op, int dtype, int64_t m, int64_t n);
/* Number of kernels this library has launched through the handle (bench `gpu_launches`). */
uint64_t tnb_launch_count(tnb_handle_t h);

/* Dry run of tnb_contract's planner: how the contraction is matricised and which kernel family / tile it maps to.
 * Pure host logic -- needs neither a handle nor a GPU (the `-m "not gpu"` tests drive the planner through it and
 * replay the address arithmetic below against the CPU oracle).  The analogue on the reference side is the host part
 * of _contract! (src/tensor/cudense.jl:255-294: mode numbering, the string key, the descriptor set-up).
 *   C[sum_i dm_i c_stride_m[i] + sum_j dn_j c_stride_n[j]]
 *       = sum_k A[sum_i dm_i a_stride_m[i] + sum_l dk_l a_stride_k[l]] * B[sum_j dn_j b_stride_n[j] + sum_l dk_l b_stride_k[l]]
 * where (dm_i) are the mixed-radix digits of m in [0, M) over ext_m (ext_m[0] fastest), likewise n and k.
 * Extent-1 modes are dropped and modes adjacent in both carrying tensors are fused.  Operands are taken as dense
 * column-major with 16-byte aligned base pointers.  num_sms <= 0 means 148 (B200). */
#define TNB_PLAN_MAX_MODES 12
typedef struct {
  int64_t M, N, K;
  int32_t n_m, n_n, n_k;          /* merged modes per group */
  int32_t family;                 /* 0 = LDGSTS tile kernel, 1 = small-K streaming kernel, 2 = TMA-staged tile kernel
                                     (shape eligible: both operands K-major, <= 5-D boxes; the run-time autotune may still
                                     prefer family 0 for a given shape) */
  int64_t ext_m[TNB_PLAN_MAX_MODES], a_stride_m[TNB_PLAN_MAX_MODES], c_stride_m[TNB_PLAN_MAX_MODES];
  int64_t ext_n[TNB_PLAN_MAX_MODES], b_stride_n[TNB_PLAN_MAX_MODES], c_stride_n[TNB_PLAN_MAX_MODES];
  int64_t ext_k[TNB_PLAN_MAX_MODES], a_stride_k[TNB_PLAN_MAX_MODES], b_stride_k[TNB_PLAN_MAX_MODES];
  int32_t a_k_major, b_k_major;   /* staging direction of each operand's tiles (1 = along K) */
  int32_t a_vec, b_vec;           /* elements per asynchronous copy (2 = 16-byte copies of f64) */
  int32_t tile_m, tile_n, tile_k;
  int32_t herm_upper;             /* 1 if TNB_HERM_UPPER is honoured (upper-triangle tiles only) */
  int64_t tiles;                  /* CTAs of the launch */
  double waves;                   /* tiles / (2 CTA slots x num_sms) */
} tnb_plan_desc;
int tnb_plan_describe(int dtype,
                      int nA, const int64_t* extA, const int32_t* modeA,
                      int nB, const int64_t* extB, const int32_t* modeB,
                      int nC, const int64_t* extC, const int32_t* modeC,
                      int flags, int num_sms, tnb_plan_desc* out, char* err, size_t errlen);
/* Same with explicit element strides per mode (NULL = dense column-major): the operand is then a strided WINDOW of a
 * larger tensor, which is how the fixed-workspace and multi-GPU entry points address slabs of L, R, A and of the output
 * vector without copying them (tnb_set_workspace_limit, tnb_heff_apply_shard_fused, tnb_env_update_*_shard). */
int tnb_plan_describe_strided(int dtype,
                              int nA, const int64_t* extA, const int32_t* modeA, const int64_t* strideA,
                              int nB, const int64_t* extB, const int32_t* modeB, const int64_t* strideB,
                              int nC, const int64_t* extC, const int32_t* modeC, const int64_t* strideC,
                              int flags, int num_sms, tnb_plan_desc* out, char* err, size_t errlen);

/* ------------------------------------------------------------------ tier 1: primitives - */

/* C[modeC] <- alpha * sum_{shared} A[modeA] * B[modeB] + beta * C[modeC]
 * Replaces _contract! -> CUTENSOR.contraction!  (src/tensor/cudense.jl:238-331), the
 * scalar/outer branches of contract!! (src/tensor/cudense.jl:83-110, 74-81, 129-168),
 * the cuBLAS fallback _gemm_contract! (src/tensor/cudense.jl:170-236) and the alpha/beta
 * form contract!!(...,alpha,beta) (src/tensor/dense.jl:1-48).
 * alpha, beta: host pointers to one element of `dtype` (NULL = 1 and 0).
 * A mode that appears in only one tensor, or in all three, is TNB_ERR_BAD_ARG.
 * The index permutation is fused into the tile loads: no permuted copy is ever made. */
int tnb_contract(tnb_handle_t h, int dtype,
                 int nA, const int64_t* extA, const int32_t* modeA, const void* A,
                 int nB, const int64_t* extB, const int32_t* modeB, const void* B,
                 int nC, const int64_t* extC, const int32_t* modeC, void* C,
                 const void* alpha, const void* beta, int flags, void* stream);

/* B[modeB] <- alpha * A[modeA] + beta * B[modeB]   (modeB is a permutation of modeA)
 * Replaces permute!/permutedims!! -> CUTENSOR.permutation! (src/tensor/cudense.jl:447-500,
 * 38-60, 112-127) and +/- -> CUTENSOR.elementwiseBinary! (src/tensor/cudense.jl:333-445):
 * one fused pass instead of zeros + elementwiseBinary + copyto!. */
int tnb_permute_axpby(tnb_handle_t h, int dtype, int n, const int64_t* extA,
                      const int32_t* modeA, const void* A, const int32_t* modeB, void* B,
                      const void* alpha, const void* beta, void* stream);

/* C[modeC] <- A[modeA] * diag[index of scaled_mode]   (modeC a permutation of modeA)
 * Contraction of a dense tensor with a Diag tensor over one of the Diag's two indices: a scale along that mode
 * (the caller relabels the mode to the Diag's other index; modeC carries the NDTensors output order).  diag:
 * device vector of extent(scaled_mode) elements of diag_dtype (TNB_F64 for the S of svd / D of eigen).
 * Replaces the Diag x Dense contractions of src/tensor/cudiag.jl:105-161, which densify the diagonal
 * (zero-fill + scatter, :147-157) and run a full dense contraction. */
int tnb_diag_contract(tnb_handle_t h, int dtype, int n, const int64_t* extA, const int32_t* modeA, const void* A,
                      int32_t scaled_mode, const void* diag, int diag_dtype, const int32_t* modeC, void* C,
                      void* stream);

/* x <- alpha * x       (scalar * and /: src/tensor/cudense.jl:22,502) */
int tnb_scale(tnb_handle_t h, int dtype, int64_t n, void* x, const void* alpha, void* stream);

/* result = sum conj(x_i) * y_i  -> device scalar `result_dev` (1 element of dtype); if
 * result_host != NULL the call synchronises and also writes it there.
 * Replaces dot = scalar(dag(A)*B) (contract to rank 0 + D2H: src/tensor/cudense.jl:25-26) */
int tnb_dot(tnb_handle_t h, int dtype, int64_t n, const void* x, const void* y,
            void* result_dev, void* result_host, void* stream);

/* ||x||_2 -> device double; optional host copy (sync).  Replaces norm (cudense.jl:27). */
int tnb_nrm2(tnb_handle_t h, int dtype, int64_t n, const void* x, double* result_dev,
             double* result_host, void* stream);

/* Spectrum truncation on a DEVICE vector of descending weights (CPU rule of [EXT] NDTensors
 * truncate!; replaces truncate!(P::CuVector) src/tensor/cutruncate.jl:1-93 -- one kernel
 * and one 24-byte readback instead of ~10 kernels and ~6 syncs).  Synchronises. */
int tnb_truncate(tnb_handle_t h, const double* P_dev, int64_t len, int64_t maxdim,
                 int64_t mindim, double cutoff, int flags, int64_t* n_keep,
                 double* truncerr, double* docut, void* stream);

/* Thin SVD with truncation of A (m x n, column-major, lda = m; destroyed).
 * U (m x kmax), S (kmax), V (n x kmax) with A ~= U diag(S) V^T (CPU `conj!(MV)` convention),
 * kmax = min(m,n,maxdim); *n_keep columns are valid.  Blocked one-sided Jacobi.
 * Replaces svd(::CuDenseTensor{_,2}) -> CUSOLVER.svd! (src/tensor/culinearalgebra.jl:33-72).
 * Synchronises (returns host scalars). */
int tnb_svd_trunc(tnb_handle_t h, int dtype, int64_t m, int64_t n, void* A,
                  int64_t maxdim, int64_t mindim, double cutoff, int flags, int do_truncate,
                  void* U, double* S, void* V, int64_t* n_keep, double* truncerr,
                  void* stream);

/* Hermitian eigendecomposition, eigenvalues DEscending, truncated.  A (n x n, upper triangle
 * referenced like syevd 'U'; destroyed).  D (kmax) device, U (n x kmax).
 * Replaces eigen(::Hermitian{CuDenseTensor}) -> syevd!/heevd! + reverse + truncate!
 * (src/tensor/culinearalgebra.jl:74-108).  Synchronises. */
int tnb_eigh_trunc(tnb_handle_t h, int dtype, int64_t n, void* A, int64_t maxdim,
                   int64_t mindim, double cutoff, int flags, int do_truncate, double* D,
                   void* U, int64_t* n_keep, double* truncerr, void* stream);

/* Thin QR with explicit Q: A (m x n) -> Q (m x k), R (k x n), k = min(m,n).  diag(R) >= 0.
 * Replaces qr(::CuDenseTensor{_,2}) (src/tensor/culinearalgebra.jl:110-121). */
int tnb_qr(tnb_handle_t h, int dtype, int64_t m, int64_t n, const void* A, void* Q, void* R,
           void* stream);

/* ------------------------------------------------------------------ tier 2: fused DMRG - */
/* Fixed layouts (column-major): phi[l,s1,s2,r], L[l,l',a], R[r,r',c], W1[a,s1,s1',b],
 * W2[b,s2,s2',c]; unprimed = ket side.  These replace sequences of primitive calls that
 * [EXT] ITensors issues through the override surface; anchors are the reference call
 * sites examples/dmrg.jl:25, test/dmrg.jl:27,75. */

typedef struct tnb_bond_dims_s {
  int64_t chiL, chiR;   /* MPS bond dims left/right of the two sites */
  int32_t d1, d2;       /* site dims */
  int32_t wL, wM, wR;   /* MPO bond dims: left of site b, between, right of site b+1 */
} tnb_bond_dims;

/* out <- H_eff * phi = (((phi*L)*W1)*W2)*R   ([EXT] product(::ProjMPO, ::ITensor)); four
 * contractions on persistent workspace, no allocation / zero-fill (replaces 4x
 * src/tensor/cudense.jl:238-331 + 3x src/tensor/cudense.jl:62-72). */
int tnb_heff_apply(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L,
                   const void* W1, const void* W2, const void* R, const void* phi,
                   void* out, void* stream);
/* Multi-GPU form: the OUTPUT bond l' is sharded.  This rank holds the slab
 * L_slab[l, l'_shard, a] (lp_extent columns of l'), full phi, W1, W2, R, and produces
 * out_slab[l'_shard, s1', s2', r'] -- 1/G of every contraction, no reduction; the caller
 * all-gathers the slabs (NCCL) when the next matvec needs the full vector.  (The reference's
 * only multi-GPU attempt is dead cuBLASMg code: src/tensor/dense.jl:195-265.) */
int tnb_heff_apply_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int64_t lp_extent,
                         const void* L_slab, const void* W1, const void* W2, const void* R,
                         const void* phi, void* out_slab, void* stream);
/* Sharded H_eff with the all-gather FUSED into the last contraction (one process per GPU, one NVSwitch node):
 * the epilogue of step 4 stores this rank's slab of H*phi straight into the full-vector buffer
 * out_peers[g] ([chiL,d1,d2,chiR], same layout as phi) of EVERY rank g through NVLink peer pointers, so the
 * transfer overlaps the GEMM tile by tile and no reassembly pass is needed; a device-side barrier over the
 * peer-mapped flag arrays flag_peers[g] (world uint64 each, zero-initialised) then makes all slabs visible to
 * work queued behind this call on `stream`.  chiL = world * lp_extent.  `epoch` must increase by one with
 * every call on the same buffers; alternate two out buffers if a peer may still read the previous result.
 * out_peers / flag_peers are HOST arrays of `world` device pointers (own buffer included, from tnb_peer_alloc /
 * tnb_peer_open).  Asynchronous.  (Replaces tnb_heff_apply_shard + ncclAllGather + the slab interleave.) */
int tnb_heff_apply_shard_fused(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int rank, int world,
                               int64_t lp_extent, const void* L_slab, const void* W1, const void* W2,
                               const void* R, const void* phi, void* const* out_peers,
                               void* const* flag_peers, uint64_t epoch, void* stream);
/* Peer-mapped device buffers (CUDA IPC).  tnb_peer_alloc: zero-filled cudaMalloc + its 64-byte IPC handle (send it
 * to the other ranks with any host-side transport); tnb_peer_open maps a peer's buffer into this process;
 * tnb_peer_close unmaps it; tnb_peer_free releases an own buffer.  tnb_peer_status synchronises and reports
 * whether a device-side peer barrier ever timed out (a rank that never arrived). */
int tnb_peer_alloc(tnb_handle_t h, size_t bytes, void** ptr, unsigned char* ipc_handle_out /* 64 bytes */);
int tnb_peer_open(tnb_handle_t h, const unsigned char* ipc_handle /* 64 bytes */, void** ptr);
int tnb_peer_close(tnb_handle_t h, void* ptr);
int tnb_peer_free(tnb_handle_t h, void* ptr);
int tnb_peer_status(tnb_handle_t h, void* stream);

/* ------------------------------------------------------------------ multi-GPU peer group -- */
/* One process per GPU on one NVSwitch node.  tnb_comm_init registers this rank's peer group on the handle: rank,
 * world (<= 8) and `world` peer-mapped flag arrays (>= 8 uint64 each, zero-initialised; own one included; from
 * tnb_peer_alloc / tnb_peer_open).  The library then numbers its own barrier epochs, so every rank must issue the
 * same sequence of collective calls (the *_shard entry points below, tnb_comm_barrier, tnb_comm_allgather).  There
 * is no library collective on the data path: results cross NVLink as peer stores from GEMM epilogues or as
 * copy-engine writes into peer-mapped buffers.  (The reference's only multi-GPU attempt is dead cuBLASMg code,
 * src/tensor/dense.jl:195-265.) */
int tnb_comm_init(tnb_handle_t h, int rank, int world, void* const* flag_peers);
int tnb_comm_finalize(tnb_handle_t h);
/* Device-side barrier over the group, stream-ordered and asynchronous. */
int tnb_comm_barrier(tnb_handle_t h, void* stream);
/* All-gather over peer memory: barrier; bytes [offset + rank*bytes_per_rank, + bytes_per_rank) of this rank's own
 * buffer bufs[rank] are copied to the same place in every peer buffer bufs[g]; barrier.  Asynchronous. */
int tnb_comm_allgather(tnb_handle_t h, void* const* bufs, size_t offset_bytes, size_t bytes_per_rank, void* stream);
/* Bytes each rank's staging buffer must hold for the sharded environment updates / noise term at bond dimension
 * <= chi, site dimension d, MPO bond <= w. */
size_t tnb_shard_stage_bytes(int dtype, int64_t chi, int32_t d, int32_t w, int world);

/* Sharded [EXT] makeL!: rank g holds L_slab[l, l'_g, a] (chiL/world columns of l') and produces
 * Lnew_slab[r, r'_g, b] (chiR/world columns of r').  Two GEMMs with 1/world of the flops each and one all-gather of
 * the small-K intermediate through the peer-mapped staging buffers stage_peers[g]. */
int tnb_env_update_left_shard(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL, int32_t wR,
                              const void* L_slab, const void* A, const void* W, void* const* stage_peers,
                              void* Lnew_slab, void* stream);
/* Sharded [EXT] makeR!: R and Rnew are replicated (full) on every rank; the flops are split 1/world and the slabs
 * of the result are all-gathered through the staging buffers. */
int tnb_env_update_right_shard(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL, int32_t wR,
                               const void* R, const void* A, const void* W, void* const* stage_peers, void* Rnew,
                               void* stream);
/* tnb_eigsolve_lanczos with the matvec sharded over the output bond: every H*v is tnb_heff_apply_shard_fused into
 * the alternating peer-mapped full-vector buffers out_a_peers / out_b_peers ([chiL,d1,d2,chiR] each, one per rank);
 * Krylov vectors and all vector operations are replicated, so no scalar ever crosses GPUs. */
int tnb_eigsolve_lanczos_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L_slab, const void* W1,
                               const void* W2, const void* R, void* phi, void* const* out_a_peers,
                               void* const* out_b_peers, int krylovdim, int maxiter, double tol, double* energy,
                               int* n_matvec, void* stream);
/* tnb_dmrg_bond_step on a peer group: phi = A1*A2 (replicated), sharded Lanczos, noise term with its two large
 * contractions sharded (stage_peers; may be NULL when noise == 0), truncated factorization replicated (same
 * deterministic kernels on identical inputs on every rank -> bit-identical A1, A2, n_keep; nothing is broadcast).
 * Buffer sizes as for tnb_dmrg_bond_step. */
int tnb_dmrg_bond_step_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int64_t chiM, const void* L_slab,
                             const void* W1, const void* W2, const void* R, void* A1, void* A2, int ortho,
                             int which_decomp, int64_t maxdim, int64_t mindim, double cutoff, double noise,
                             int krylovdim, int maxiter, void* const* out_a_peers, void* const* out_b_peers,
                             void* const* stage_peers, double* energy, int64_t* n_keep, double* truncerr,
                             void* stream);
/* End-to-end sharded matvec with HOST buffers (bench.py `e2e` at N > 1): this rank uploads only its r-chunk of
 * phi_host (1/world of the vector; chunk g = r in [g*ceil(chiR/world), ...)), forwards it to the peer-mapped phi
 * buffers of all ranks over NVLink, runs its l' slab with the all-gather fused into step 4 (every rank then holds the
 * full H*phi on the device), and downloads only its own l' slab of H*phi into out_host (a strided window: phi_host /
 * out_host have phi's layout), piece by piece over r' while the next piece is being computed.  Synchronous. */
int tnb_heff_apply_shard_host(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L_slab, const void* W1,
                              const void* W2, const void* R, const void* phi_host, void* const* phi_peers,
                              void* const* out_peers, void* out_host, void* stream);

/* Same, phi and out in HOST memory (pinned or pageable): H2D + apply + D2H, synchronous.
 * This is the end-to-end form timed by bench.py `e2e`. */
int tnb_heff_apply_host(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L,
                        const void* W1, const void* W2, const void* R, const void* phi_host,
                        void* out_host, void* stream);

/* Lnew[r,r',b] <- ((L*A)*W)*conj(A')  with A[l,s,r]  ([EXT] makeL!) */
int tnb_env_update_left(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d,
                        int32_t wL, int32_t wR, const void* L, const void* A, const void* W,
                        void* Lnew, void* stream);
/* Rnew[l,l',a] <- ((R*A)*W)*conj(A')  ([EXT] makeR!) */
int tnb_env_update_right(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d,
                         int32_t wL, int32_t wR, const void* R, const void* A, const void* W,
                         void* Rnew, void* stream);

/* Lowest eigenpair of H_eff by one Lanczos cycle with full re-orthogonalisation
 * ([EXT] KrylovKit.eigsolve(PH, phi, 1, :SR; krylovdim, maxiter, tol)).  phi: in = start
 * vector, out = normalised Ritz vector.  Krylov vectors and all scalars stay on the device;
 * one readback at the end.  Synchronises. */
int tnb_eigsolve_lanczos(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L,
                         const void* W1, const void* W2, const void* R, void* phi,
                         int krylovdim, int maxiter, double tol, double* energy,
                         int* n_matvec, void* stream);

/* rho_pert <- noise * nt*nt^dagger  ([EXT] noiseterm(::ProjMPO, phi, ortho)); rho_pert is
 * (chiL*d1)^2 for ortho left, (d2*chiR)^2 for ortho right; accumulate=1 adds into it.
 * Hermitian: only the UPPER triangle is computed (what tnb_factorize_bond / tnb_eigh_trunc read). */
int tnb_noise_term(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L,
                   const void* W1, const void* W2, const void* R, const void* phi, int ortho,
                   double noise, int accumulate, void* rho, void* stream);

/* Split phi[l,s1,s2,r] -> A[l,s1,k] * B[k,s2,r] with truncation, orthogonality side and
 * normalisation ([EXT] replacebond! -> factorize; svd/eigen bodies at
 * src/tensor/culinearalgebra.jl:33-108).  Output capacity: A holds chiL*d1*kmax elements and B kmax*d2*chiR with
 *   kmax = min(r, maxdim),  r = min(chiL*d1, d2*chiR) on the svd / qr branches,
 *                           r = chiL*d1 (ortho left) or d2*chiR (ortho right) on the EIGEN branch
 * (which_decomp = EIGEN, or AUTO with rho_pert != NULL or cutoff > 1e-12): a perturbed density matrix can have
 * more non-zero eigenvalues than min(m, n), and like [EXT] factorize_eigen the call keeps up to maxdim of them.
 * maxdim <= 0 means "no limit" (kmax = r).  Matrix sizes up to 32768 (eigensolver workspace 4 n^2 doubles).
 * rho_pert may be NULL.  Synchronises. */
int tnb_factorize_bond(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, void* phi,
                       int ortho, int which_decomp, int64_t maxdim, int64_t mindim,
                       double cutoff, const void* rho_pert, int normalize, void* A, void* B,
                       int64_t* n_keep, double* truncerr, void* stream);

/* One full two-site DMRG bond update: phi = A1*A2; Lanczos; optional noise; factorize.
 * In: A1[chiL,d1,chiM], A2[chiM,d2,chiR].  Out (overwritten): A1[chiL,d1,k], A2[k,d2,chiR] with
 * k = *n_keep; both buffers must hold max(input, chiL*d1*kmax / kmax*d2*chiR) elements, kmax as documented at
 * tnb_factorize_bond (the eigen branch -- noise > 0 or cutoff > 1e-12 -- is bounded by the ORTHO side's dimension,
 * not by min(m, n)). */
int tnb_dmrg_bond_step(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int64_t chiM, const void* L,
                       const void* W1, const void* W2, const void* R, void* A1, void* A2,
                       int ortho, int which_decomp, int64_t maxdim, int64_t mindim,
                       double cutoff, double noise, int krylovdim, int maxiter,
                       double* energy, int64_t* n_keep, double* truncerr, void* stream);

/* One full two-site DMRG sweep (left-to-right, then right-to-left: 2(nsites-1) bond steps) with the host control flow
 * inside the library ([EXT] the body of ITensors' sweep loop in src/mps/dmrg.jl: sweepnext / position! / eigsolve /
 * replacebond!; reference call sites examples/dmrg.jl:25, test/dmrg.jl:27,75).  The orthogonality centre must be at
 * site 0 on entry and is there again on return.
 *   chi[0..nsites]   bond dimensions (chi[0] = chi[nsites] = 1), UPDATED in place
 *   d[j], w[t]       site dimensions, MPO bond dimensions (w[0] = w[nsites] = 1)
 *   A[j], capA[j]    site tensors A[chi[j], d[j], chi[j+1]] in buffers of capA[j] elements; a bond step needs room for
 *                    kmax = min(r, maxdim) (tnb_factorize_bond's capacity rule) on both of its sites
 *   W[j]             MPO tensors W[w[j], d[j], d[j], w[j+1]]
 *   env[t], capE[t]  one environment buffer per boundary t = 0..nsites (boundary t is left of site t), capE[t] >=
 *                    chi_max[t]^2 * w[t] elements: it holds the LEFT environment of the boundary while the centre is
 *                    right of it and the RIGHT environment otherwise.  build_right_envs != 0 builds the right
 *                    environments of boundaries 2..nsites-1 first ([EXT] position!(PH, psi, 1)); with 0 they must be
 *                    the ones a previous call left behind.
 *   bond_energies / bond_truncerrs (optional, 2(nsites-1) doubles each): per bond step, in sweep order.
 * Synchronises. */
int tnb_dmrg_sweep(tnb_handle_t h, int dtype, int32_t nsites, int64_t* chi, const int32_t* d, const int32_t* w,
                   void* const* A, const int64_t* capA, const void* const* W, void* const* env, const int64_t* capE,
                   int build_right_envs, int64_t maxdim, int64_t mindim, double cutoff, double noise, int which_decomp,
                   int krylovdim, int maxiter, double* energy, double* maxerr, double* bond_energies,
                   double* bond_truncerrs, void* stream);

/* theta[l,s1',s2',r] <- sum G[s1',s2',s1,s2] A1[l,s1,k] A2[k,s2,r]  then left-orthogonal
 * split with truncation ([EXT] apply / product(o, psi); examples/gate_evolution.jl:46).  A1 / A2 capacities as for
 * tnb_dmrg_bond_step with ortho left (cutoff > 1e-12 selects the eigen branch: kmax = min(chiL*d1, maxdim)). */
int tnb_tebd_apply_gate(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiM, int64_t chiR,
                        int32_t d1, int32_t d2, const void* G, void* A1, void* A2,
                        int64_t maxdim, int64_t mindim, double cutoff, int64_t* n_keep,
                        double* truncerr, void* stream);

/* Same gate in B form: B1[chiL,d1,chiM], B2[chiM,d2,chiR] right-canonical in the Schmidt bases, lamL[chiL] the
 * Schmidt values of the bond to the LEFT of the first site (device).  Out: B1[chiL,d1,k], B2[k,d2,chiR] (buffers
 * sized like tnb_tebd_apply_gate), lam_out[k] = Schmidt values of the updated bond (device, normalised).
 * In this form the gates of one even/odd TEBD layer share no data, which is what lets a layer be spread
 * over GPUs (itensorsgpu.jl_b200/tebd.py).  No division by Schmidt values.  Synchronises. */
int tnb_tebd_gate_bform(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiM, int64_t chiR,
                        int32_t d1, int32_t d2, const void* G, const double* lamL, void* B1, void* B2,
                        int64_t maxdim, int64_t mindim, double cutoff, double* lam_out, int64_t* n_keep,
                        double* truncerr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNB200_H */
