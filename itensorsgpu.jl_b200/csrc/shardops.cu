// Multi-GPU forms of the DMRG tier (one process per GPU on one NVSwitch node, no library collective on the data
// path): peer group of the handle, all-gather over peer memory, sharded environment updates, sharded Lanczos and
// the sharded bond step used by dmrg(..., comm=...), and the end-to-end host-buffer matvec of bench.py at N > 1.
//
// Partitioning ("load-balanced alternative"): rank g owns the slab L[:, l'_g, :] of every LEFT
// environment (it never moves); right environments, MPO tensors, the MPS and all Krylov vectors are replicated.
// Every GEMM of the matvec and of both environment updates does 1/world of the flops; what crosses NVLink is
// the result slab of step 4 (peer stores from the GEMM epilogue, heff.cu) and, per environment update, one
// all-gather of the small-K intermediate (copy engines writing into peer-mapped staging buffers).  The truncated
// factorization is replicated: every rank runs the same deterministic kernels on identical inputs (no atomics
// anywhere), so A, B and n_keep are bit-identical on all ranks and nothing has to be broadcast.
//
// The reference's only multi-GPU attempt is dead cuBLASMg code (/root/reference/src/tensor/dense.jl:195-265).
#include "tnb_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace tnb {

int heff_shard_fused_host_tail(Handle* h, int dtype, const tnb_bond_dims* d, int rank, int world, int64_t clp,
                               const void* Lslab, const void* W1, const void* W2, const void* R, const void* phi,
                               void* const* out_peers, void* t0, void* t1, void* out_host, cudaStream_t st);

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static int need_comm(Handle* h, const char* who) {
  if (!h->comm.on) return set_err(h, TNB_ERR_BAD_ARG, "%s: no peer group on this handle (tnb_comm_init)", who);
  return TNB_OK;
}

// labels shared by the contractions below
enum { mL = 0, mS1, mS2, mR, mLp, mA, mS1p, mB, mS2p, mC, mRp, mG, mLpp, mS1pp, mS2pp, mRpp, mG2 };

// ------------------------------------------------------------------------------------
// Lnew[r, r'_g, b] = sum_{l,l',a,s,s'} L[l,l',a] A[l,s,r] W[a,s,s',b] conj(A)[l',s',r'_g]          ([EXT] makeL!)
// in : Lslab[l, l'_g, a]   out: the r' slab of the new environment
// ------------------------------------------------------------------------------------
static int env_left_shard(Handle* h, int dtype, int64_t cl, int64_t cr, int64_t d, int64_t wl, int64_t wr,
                          const void* Lslab, const void* A, const void* W, void* const* stage, void* Lnew,
                          cudaStream_t st) {
  const int rank = h->comm.rank, world = h->comm.world;
  const size_t es = elsize(dtype);
  const int64_t clp = cl / world, crp = cr / world;
  const size_t slab = (size_t)clp * cr * d * wr;            // elements of T2_g
  ws_reset(h);
  const size_t n1 = al256((size_t)clp * wl * d * cr * es);
  TNB_TRY(ws_require(h, n1));
  void* t0;
  TNB_TRY(ws_alloc(h, n1, &t0));
  {  // T1[l'_g,a,s,r] = L[l,l'_g,a] A[l,s,r]
    int64_t ea[] = {cl, clp, wl}; int32_t ma[] = {mL, mLp, mA};
    int64_t eb[] = {cl, d, cr};   int32_t mb[] = {mL, mS1, mR};
    int64_t ec[] = {clp, wl, d, cr}; int32_t mc[] = {mLp, mA, mS1, mR};
    TNB_TRY(contract_impl(h, dtype, 3, ea, ma, Lslab, 3, eb, mb, A, 4, ec, mc, t0, nullptr, nullptr, 0, st));
  }
  {  // T2_g[l'_g,r,s',b] = T1 W[a,s,s',b]  -> this rank's region of its own staging buffer
    int64_t ea[] = {clp, wl, d, cr}; int32_t ma[] = {mLp, mA, mS1, mR};
    int64_t eb[] = {wl, d, d, wr};   int32_t mb[] = {mA, mS1, mS1p, mB};
    int64_t ec[] = {clp, cr, d, wr}; int32_t mc[] = {mLp, mR, mS1p, mB};
    TNB_TRY(contract_impl(h, dtype, 4, ea, ma, t0, 4, eb, mb, W, 4, ec, mc, (char*)stage[rank] + (size_t)rank * slab * es,
                          nullptr, nullptr, 0, st));
  }
  TNB_TRY(comm_allgather(h, stage, 0, slab * es, st));
  {  // Lnew[r,r'_g,b] = T2all[l'_s,r,s',b,g] conj(A)[(l'_s,g),s',r'_g]
    int64_t ea[] = {clp, cr, d, wr, world}; int32_t ma[] = {mLp, mR, mS1p, mB, mG};
    int64_t eb[] = {clp, world, d, crp};    int32_t mb[] = {mLp, mG, mS1p, mRp};
    int64_t ec[] = {cr, crp, wr};           int32_t mc[] = {mR, mRp, mB};
    TNB_TRY(contract_impl(h, dtype, 5, ea, ma, stage[rank], 4, eb, mb, (const char*)A + (size_t)rank * crp * cl * d * es, 3, ec, mc,
                          Lnew, nullptr, nullptr, TNB_CONJ_B, st));
  }
  return TNB_OK;
}

// ------------------------------------------------------------------------------------
// Rnew[l,l',a] = sum R[r,r',c] A[l,s,r] W[a,s,s',c] conj(A)[l',s',r']                                ([EXT] makeR!)
// R is replicated (full) and so is the result: two sharded GEMMs + two all-gathers.
// ------------------------------------------------------------------------------------
static int env_right_shard(Handle* h, int dtype, int64_t cl, int64_t cr, int64_t d, int64_t wl, int64_t wr,
                           const void* R, const void* A, const void* W, void* const* stage, void* Rnew,
                           cudaStream_t st) {
  const int rank = h->comm.rank, world = h->comm.world;
  const size_t es = elsize(dtype);
  const int64_t clp = cl / world, crp = cr / world;
  const size_t slabA = (size_t)crp * cl * wl * d;           // T2_g
  const size_t slabB = (size_t)cl * clp * wl;               // Rnew_g
  const size_t offB = al256((size_t)world * slabA * es);
  ws_reset(h);
  const size_t n1 = al256((size_t)crp * wr * cl * d * es);
  TNB_TRY(ws_require(h, n1));
  void* t0;
  TNB_TRY(ws_alloc(h, n1, &t0));
  {  // T1[r'_g,c,l,s] = R[r,r'_g,c] A[l,s,r]      (R window: r' restricted to this rank's slab)
    int64_t ea[] = {cr, crp, wr}; int32_t ma[] = {mR, mRp, mC};
    int64_t sa[] = {1, cr, cr * cr};
    int64_t eb[] = {cl, d, cr};   int32_t mb[] = {mL, mS1, mR};
    int64_t ec[] = {crp, wr, cl, d}; int32_t mc[] = {mRp, mC, mL, mS1};
    TNB_TRY(contract_impl_ex(h, dtype, 3, ea, ma, (const char*)R + (size_t)rank * crp * cr * es, 3, eb, mb, A, 4, ec, mc, t0,
                             nullptr, nullptr, 0, st, nullptr, nullptr, 0, sa, nullptr));
  }
  {  // T2_g[r'_g,l,a,s'] = T1 W[a,s,s',c]
    int64_t ea[] = {crp, wr, cl, d}; int32_t ma[] = {mRp, mC, mL, mS1};
    int64_t eb[] = {wl, d, d, wr};   int32_t mb[] = {mA, mS1, mS1p, mC};
    int64_t ec[] = {crp, cl, wl, d}; int32_t mc[] = {mRp, mL, mA, mS1p};
    TNB_TRY(contract_impl(h, dtype, 4, ea, ma, t0, 4, eb, mb, W, 4, ec, mc, (char*)stage[rank] + (size_t)rank * slabA * es,
                          nullptr, nullptr, 0, st));
  }
  TNB_TRY(comm_allgather(h, stage, 0, slabA * es, st));
  {  // Rnew_g[l,l'_g,a] = T2all[r'_s,l,a,s',g] conj(A)[l'_g,s',(r'_s,g)]   (A window: l' restricted to the slab)
    int64_t ea[] = {crp, cl, wl, d, world}; int32_t ma[] = {mRp, mL, mA, mS1p, mG};
    int64_t eb[] = {clp, d, crp, world};    int32_t mb[] = {mLp, mS1p, mRp, mG};
    int64_t sb[] = {1, cl, cl * d, cl * d * crp};
    int64_t ec[] = {cl, clp, wl};           int32_t mc[] = {mL, mLp, mA};
    TNB_TRY(contract_impl_ex(h, dtype, 5, ea, ma, stage[rank], 4, eb, mb, (const char*)A + (size_t)rank * clp * es, 3, ec, mc,
                             (char*)stage[rank] + offB + (size_t)rank * slabB * es, nullptr, nullptr, TNB_CONJ_B, st, nullptr,
                             nullptr, 0, nullptr, sb));
  }
  TNB_TRY(comm_allgather(h, stage, offB, slabB * es, st));
  {  // [l, l'_s, a, g] -> [l, (l'_s, g), a]
    int64_t ea[] = {cl, clp, wl, world}; int32_t ma[] = {mL, mLp, mA, mG};
    int32_t mb[] = {mL, mLp, mG, mA};
    TNB_TRY(permute_axpby_impl(h, dtype, 4, ea, ma, (const char*)stage[rank] + offB, mb, Rnew, nullptr, nullptr, st));
  }
  return TNB_OK;
}

// ------------------------------------------------------------------------------------
// noise term with the first (big) contraction and the small-K step sharded; Gram replicated
// ------------------------------------------------------------------------------------
static int noise_term_shard(Handle* h, int dtype, const tnb_bond_dims* d, const void* Lslab, const void* W1, const void* W2,
                            const void* R, const void* phi, int ortho, double noise, void* rho, void* const* stage, void* t0,
                            void* t1, cudaStream_t st) {
  const int rank = h->comm.rank, world = h->comm.world;
  const size_t es = elsize(dtype);
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2, wl = d->wL, wm = d->wM, wr = d->wR;
  const int64_t clp = cl / world, crp = cr / world;
  double alpha[2] = {noise, 0.0};
  if (ortho == TNB_ORTHO_LEFT) {
    const size_t slab = (size_t)d2 * cr * clp * d1 * wm;
    {  // T1[s1,s2,r,l'_g,a] = phi L_slab
      int64_t ea[] = {cl, d1, d2, cr}; int32_t ma[] = {mL, mS1, mS2, mR};
      int64_t eb[] = {cl, clp, wl};    int32_t mb[] = {mL, mLp, mA};
      int64_t ec[] = {d1, d2, cr, clp, wl}; int32_t mc[] = {mS1, mS2, mR, mLp, mA};
      TNB_TRY(contract_impl(h, dtype, 4, ea, ma, phi, 3, eb, mb, Lslab, 5, ec, mc, t0, nullptr, nullptr, 0, st));
    }
    {  // nt_g[s2,r,l'_g,s1',b] = T1 W1
      int64_t ea[] = {d1, d2, cr, clp, wl}; int32_t ma[] = {mS1, mS2, mR, mLp, mA};
      int64_t eb[] = {wl, d1, d1, wm};      int32_t mb[] = {mA, mS1, mS1p, mB};
      int64_t ec[] = {d2, cr, clp, d1, wm}; int32_t mc[] = {mS2, mR, mLp, mS1p, mB};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W1, 5, ec, mc, (char*)stage[rank] + (size_t)rank * slab * es,
                            nullptr, nullptr, 0, st));
    }
    TNB_TRY(comm_allgather(h, stage, 0, slab * es, st));
    {  // [s2,r,l'_s,s1',b,g] -> nt[s2,r,(l'_s,g),s1',b]
      int64_t ea[] = {d2, cr, clp, d1, wm, world}; int32_t ma[] = {mS2, mR, mLp, mS1p, mB, mG};
      int32_t mb[] = {mS2, mR, mLp, mG, mS1p, mB};
      TNB_TRY(permute_axpby_impl(h, dtype, 6, ea, ma, stage[rank], mb, t1, nullptr, nullptr, st));
    }
    {  // rho[l',s1',l'',s1''] = noise * nt conj(nt)
      int64_t ea[] = {d2, cr, cl, d1, wm}; int32_t ma[] = {mS2, mR, mLp, mS1p, mB};
      int32_t mb[] = {mS2, mR, mLpp, mS1pp, mB};
      int64_t ec[] = {cl, d1, cl, d1};     int32_t mc[] = {mLp, mS1p, mLpp, mS1pp};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 5, ea, mb, t1, 4, ec, mc, rho, alpha, nullptr, TNB_CONJ_B | TNB_HERM_UPPER, st));
    }
  } else {
    const size_t slab = (size_t)cl * d1 * d2 * crp * wm;
    {  // T1[l,s1,s2,r'_g,c] = phi R[r, r'_g, c]
      int64_t ea[] = {cl, d1, d2, cr}; int32_t ma[] = {mL, mS1, mS2, mR};
      int64_t eb[] = {cr, crp, wr};    int32_t mb[] = {mR, mRp, mC};
      int64_t sb[] = {1, cr, cr * cr};
      int64_t ec[] = {cl, d1, d2, crp, wr}; int32_t mc[] = {mL, mS1, mS2, mRp, mC};
      TNB_TRY(contract_impl_ex(h, dtype, 4, ea, ma, phi, 3, eb, mb, (const char*)R + (size_t)rank * crp * cr * es, 5, ec, mc, t0,
                               nullptr, nullptr, 0, st, nullptr, nullptr, 0, nullptr, sb));
    }
    {  // nt_g[l,s1,s2',r'_g,b] = T1 W2[b,s2,s2',c]
      int64_t ea[] = {cl, d1, d2, crp, wr}; int32_t ma[] = {mL, mS1, mS2, mRp, mC};
      int64_t eb[] = {wm, d2, d2, wr};      int32_t mb[] = {mB, mS2, mS2p, mC};
      int64_t ec[] = {cl, d1, d2, crp, wm}; int32_t mc[] = {mL, mS1, mS2p, mRp, mB};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W2, 5, ec, mc, (char*)stage[rank] + (size_t)rank * slab * es,
                            nullptr, nullptr, 0, st));
    }
    TNB_TRY(comm_allgather(h, stage, 0, slab * es, st));
    {  // [l,s1,s2',r'_s,b,g] -> nt[l,s1,s2',(r'_s,g),b]
      int64_t ea[] = {cl, d1, d2, crp, wm, world}; int32_t ma[] = {mL, mS1, mS2p, mRp, mB, mG};
      int32_t mb[] = {mL, mS1, mS2p, mRp, mG, mB};
      TNB_TRY(permute_axpby_impl(h, dtype, 6, ea, ma, stage[rank], mb, t1, nullptr, nullptr, st));
    }
    {  // rho[s2',r',s2'',r''] = noise * nt conj(nt)
      int64_t ea[] = {cl, d1, d2, cr, wm}; int32_t ma[] = {mL, mS1, mS2p, mRp, mB};
      int32_t mb[] = {mL, mS1, mS2pp, mRpp, mB};
      int64_t ec[] = {d2, cr, d2, cr};     int32_t mc[] = {mS2p, mRp, mS2pp, mRpp};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 5, ea, mb, t1, 4, ec, mc, rho, alpha, nullptr, TNB_CONJ_B | TNB_HERM_UPPER, st));
    }
  }
  return TNB_OK;
}

static int check_shard_dims(Handle* h, const tnb_bond_dims* d, const char* who) {
  if (!d) return set_err(h, TNB_ERR_BAD_ARG, "%s: null dims", who);
  if (d->chiL < 1 || d->chiR < 1 || d->d1 < 1 || d->d2 < 1 || d->wL < 1 || d->wM < 1 || d->wR < 1)
    return set_err(h, TNB_ERR_BAD_ARG, "%s: bond dims must be >= 1", who);
  if (d->chiL % h->comm.world) return set_err(h, TNB_ERR_BAD_ARG, "%s: chiL = %lld is not divisible by %d ranks", who, (long long)d->chiL, h->comm.world);
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;
#define H ((Handle*)h)
#define ST ((cudaStream_t)stream)

extern "C" {

int tnb_comm_init(tnb_handle_t h, int rank, int world, void* const* flag_peers) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (world < 1 || world > TNB_MAX_PEERS || rank < 0 || rank >= world) return set_err(H, TNB_ERR_BAD_ARG, "comm_init: rank %d of %d", rank, world);
  if (!flag_peers) return set_err(H, TNB_ERR_BAD_ARG, "comm_init: null flag table");
  for (int g = 0; g < world; ++g)
    if (!flag_peers[g]) return set_err(H, TNB_ERR_BAD_ARG, "comm_init: null flag pointer %d", g);
  H->comm = Comm();
  H->comm.on = true;
  H->comm.rank = rank;
  H->comm.world = world;
  for (int g = 0; g < world; ++g) H->comm.flags[g] = (unsigned long long*)flag_peers[g];
  return TNB_OK;
}

int tnb_comm_finalize(tnb_handle_t h) {
  if (!h) return TNB_ERR_BAD_ARG;
  H->comm = Comm();
  return TNB_OK;
}

int tnb_comm_barrier(tnb_handle_t h, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return comm_barrier(H, ST);
}

int tnb_comm_allgather(tnb_handle_t h, void* const* bufs, size_t offset_bytes, size_t bytes_per_rank, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!bufs) return set_err(H, TNB_ERR_BAD_ARG, "comm_allgather: null buffer table");
  return comm_allgather(H, bufs, offset_bytes, bytes_per_rank, ST);
}

int tnb_env_update_left_shard(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL, int32_t wR,
                              const void* L_slab, const void* A, const void* W, void* const* stage_peers,
                              void* Lnew_slab, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_TRY(need_comm(H, "env_update_left_shard"));
  if (!L_slab || !A || !W || !stage_peers || !Lnew_slab) return set_err(H, TNB_ERR_BAD_ARG, "env_update_left_shard: null pointer");
  if (chiL < 1 || chiR < 1 || d < 1 || wL < 1 || wR < 1) return set_err(H, TNB_ERR_BAD_ARG, "env_update_left_shard: dims");
  if (chiL % H->comm.world || chiR % H->comm.world)
    return set_err(H, TNB_ERR_BAD_ARG, "env_update_left_shard: bond dims %lld, %lld must be divisible by %d ranks", (long long)chiL, (long long)chiR, H->comm.world);
  return env_left_shard(H, dtype, chiL, chiR, d, wL, wR, L_slab, A, W, stage_peers, Lnew_slab, ST);
}

int tnb_env_update_right_shard(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL, int32_t wR,
                               const void* R, const void* A, const void* W, void* const* stage_peers, void* Rnew,
                               void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_TRY(need_comm(H, "env_update_right_shard"));
  if (!R || !A || !W || !stage_peers || !Rnew) return set_err(H, TNB_ERR_BAD_ARG, "env_update_right_shard: null pointer");
  if (chiL < 1 || chiR < 1 || d < 1 || wL < 1 || wR < 1) return set_err(H, TNB_ERR_BAD_ARG, "env_update_right_shard: dims");
  if (chiL % H->comm.world || chiR % H->comm.world)
    return set_err(H, TNB_ERR_BAD_ARG, "env_update_right_shard: bond dims %lld, %lld must be divisible by %d ranks", (long long)chiL, (long long)chiR, H->comm.world);
  return env_right_shard(H, dtype, chiL, chiR, d, wL, wR, R, A, W, stage_peers, Rnew, ST);
}

size_t tnb_shard_stage_bytes(int dtype, int64_t chi, int32_t d, int32_t w, int world) {
  const size_t es = elsize(dtype);
  const size_t env = al256((size_t)chi * chi * d * w * es) + al256((size_t)chi * chi * w * es);   // T2 + Rnew regions
  const size_t nt = al256((size_t)chi * chi * d * d * w * es);                                    // noise term
  (void)world;
  return std::max(env, nt) + 4096;
}

int tnb_eigsolve_lanczos_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L_slab, const void* W1,
                               const void* W2, const void* R, void* phi, void* const* out_a_peers,
                               void* const* out_b_peers, int krylovdim, int maxiter, double tol, double* energy,
                               int* n_matvec, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_TRY(need_comm(H, "eigsolve_lanczos_shard"));
  if (!L_slab || !W1 || !W2 || !R || !phi || !out_a_peers || !out_b_peers) return set_err(H, TNB_ERR_BAD_ARG, "eigsolve_lanczos_shard: null pointer");
  TNB_TRY(check_shard_dims(H, dims, "eigsolve_lanczos_shard"));
  ShardCtx sc;
  sc.clp = dims->chiL / H->comm.world;
  sc.out[0] = out_a_peers;
  sc.out[1] = out_b_peers;
  return lanczos_impl(H, dtype, dims, L_slab, W1, W2, R, phi, krylovdim, maxiter, tol, energy, n_matvec, ST, &sc);
}

int tnb_dmrg_bond_step_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* d, int64_t chiM, const void* L_slab,
                             const void* W1, const void* W2, const void* R, void* A1, void* A2, int ortho,
                             int which_decomp, int64_t maxdim, int64_t mindim, double cutoff, double noise,
                             int krylovdim, int maxiter, void* const* out_a_peers, void* const* out_b_peers,
                             void* const* stage_peers, double* energy, int64_t* n_keep, double* truncerr,
                             void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_TRY(need_comm(H, "dmrg_bond_step_shard"));
  if (!d || !L_slab || !W1 || !W2 || !R || !A1 || !A2 || !out_a_peers || !out_b_peers)
    return set_err(H, TNB_ERR_BAD_ARG, "dmrg_bond_step_shard: null pointer");
  if (chiM < 1) return set_err(H, TNB_ERR_BAD_ARG, "dmrg_bond_step_shard: chiM < 1");
  TNB_TRY(check_shard_dims(H, d, "dmrg_bond_step_shard"));
  const int world = H->comm.world;
  if (noise > 0) {
    if (!stage_peers) return set_err(H, TNB_ERR_BAD_ARG, "dmrg_bond_step_shard: the noise term needs the staging buffers");
    if (ortho == TNB_ORTHO_RIGHT && d->chiR % world)
      return set_err(H, TNB_ERR_BAD_ARG, "dmrg_bond_step_shard: chiR = %lld is not divisible by %d ranks", (long long)d->chiR, world);
  }
  const size_t es = elsize(dtype);
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2;
  const int64_t m = cl * d1, n = d2 * cr;
  const size_t phib = al256((size_t)m * n * es);
  const int64_t r = (ortho == TNB_ORTHO_LEFT) ? m : n;
  ShardCtx sc;
  sc.clp = cl / world;
  sc.out[0] = out_a_peers;
  sc.out[1] = out_b_peers;
  const size_t hw_slab = heff_shard_ws_bytes(dtype, d, sc.clp);
  // the gathered noise tensor is full size whatever the single-GPU chunking of H_eff would be
  const size_t hw_full = 2 * al256((size_t)cl * cr * d1 * d2 * std::max({d->wL, d->wM, d->wR}) * es);
  const size_t lan = hw_slab + (size_t)(krylovdim + 1) * phib + (1 << 16);
  const size_t nz = hw_slab / 2 + hw_full / 2 + (1 << 16);              // noise: slab-size T1 + full-size nt
  size_t need = phib + (noise > 0 ? al256((size_t)r * r * es) : 0);
  size_t stage = std::max(lan, noise > 0 ? nz : (size_t)0);
  stage = std::max(stage, factorize_ws_bytes_pub(dtype, m, n));
  ws_reset(H);
  TNB_TRY(ws_require(H, need + stage + (1 << 16)));
  void *phi, *rho = nullptr;
  TNB_TRY(ws_alloc(H, (size_t)m * n * es, &phi));
  if (noise > 0) TNB_TRY(ws_alloc(H, (size_t)r * r * es, &rho));
  const size_t mark = H->ws_off;
  TNB_TRY(gemm_impl(H, dtype, 'N', 'N', m, n, chiM, nullptr, A1, m, A2, chiM, nullptr, phi, m, ST));
  int nmv = 0;
  H->ws_base = mark;
  int rc = lanczos_impl(H, dtype, d, L_slab, W1, W2, R, phi, krylovdim, maxiter, 1e-14, energy, &nmv, ST, &sc);
  if (!rc && noise > 0) {
    ws_reset(H);
    void *t0, *t1;
    rc = ws_alloc(H, hw_slab / 2, &t0);
    if (!rc) rc = ws_alloc(H, hw_full / 2, &t1);
    if (!rc) rc = noise_term_shard(H, dtype, d, L_slab, W1, W2, R, phi, ortho, noise, rho, stage_peers, t0, t1, ST);
  }
  if (!rc) {
    ws_reset(H);
    rc = factorize_core_pub(H, dtype, m, n, phi, ortho, which_decomp, maxdim, mindim, cutoff, rho, 1, A1, A2, n_keep, truncerr, ST);
  }
  H->ws_base = 0;
  ws_reset(H);
  if (rc) return rc;
  return check_cuda(H, cudaStreamSynchronize(ST), "dmrg_bond_step_shard sync");
}

// End-to-end sharded matvec with HOST buffers (bench.py `e2e` at N > 1).  Every rank uploads only ITS r-chunk of phi
// (1/world of the vector over PCIe), forwards it to the peers over NVLink, runs its l' slab of the matvec with the
// all-gather fused into step 4 (after which every rank holds the full H*phi on the device), and downloads only ITS OWN
// l' slab of H*phi -- a strided window of out_host (phi's layout), piece by piece over r' while the next piece is still
// being computed (heff.cu: heff_shard_fused_host_tail).  Whole job: one vector up, one vector down, everything else over
// NVLink.  Synchronous.
int tnb_heff_apply_shard_host(tnb_handle_t h, int dtype, const tnb_bond_dims* d, const void* L_slab, const void* W1,
                              const void* W2, const void* R, const void* phi_host, void* const* phi_peers,
                              void* const* out_peers, void* out_host, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_TRY(need_comm(H, "heff_apply_shard_host"));
  if (!L_slab || !W1 || !W2 || !R || !phi_host || !phi_peers || !out_peers || !out_host)
    return set_err(H, TNB_ERR_BAD_ARG, "heff_apply_shard_host: null pointer");
  TNB_TRY(check_shard_dims(H, d, "heff_apply_shard_host"));
  const int rank = H->comm.rank, world = H->comm.world;
  const size_t es = elsize(dtype);
  const int64_t cl = d->chiL, cr = d->chiR, clp = cl / world;
  (void)es;
  const size_t col = (size_t)cl * d->d1 * d->d2 * es;                  // bytes of phi per unit of r
  const int64_t rc = (cr + world - 1) / world;
  const int64_t r0 = std::min<int64_t>(cr, rank * rc), r1 = std::min<int64_t>(cr, r0 + rc);
  ws_reset(H);
  const size_t hw = heff_shard_ws_bytes(dtype, d, clp);
  TNB_TRY(ws_require(H, hw));
  void *t0, *t1;
  TNB_TRY(ws_alloc(H, hw / 2, &t0));
  TNB_TRY(ws_alloc(H, hw / 2, &t1));
  // TNB_E2E_TRACE=1: per-phase CUDA-event times of every call on stderr (diagnostics; adds event overhead)
  static const bool trace = getenv("TNB_E2E_TRACE") != nullptr;
  cudaEvent_t te[6] = {};
  auto mark = [&](int i) { if (trace) { if (!te[i]) cudaEventCreate(&te[i]); cudaEventRecord(te[i], ST); } };
  mark(0);
  TNB_TRY(comm_barrier(H, ST));                                       // peers are done with the previous phi / result
  mark(1);
  if (r1 > r0) {
    char* own = (char*)phi_peers[rank] + r0 * col;
    TNB_CUDA(H, cudaMemcpyAsync(own, (const char*)phi_host + r0 * col, (r1 - r0) * col, cudaMemcpyHostToDevice, ST));
    for (int i = 1; i < world; ++i) {
      const int g = (rank + i) % world;
      TNB_CUDA(H, cudaMemcpyAsync((char*)phi_peers[g] + r0 * col, own, (r1 - r0) * col, cudaMemcpyDefault, ST));
    }
  }
  mark(2);
  TNB_TRY(comm_barrier(H, ST));                                       // the full phi is on every rank
  mark(3);
  // the copy stream must not start before work already queued on `stream`
  TNB_TRY(heff_shard_fused_host_tail(H, dtype, d, rank, world, clp, L_slab, W1, W2, R, phi_peers[rank], out_peers, t0, t1, out_host, ST));
  mark(4);
  TNB_TRY(comm_barrier(H, ST));        // every rank's slab has landed everywhere: the device copies are complete vectors
  mark(5);
  TNB_CUDA(H, cudaStreamSynchronize(H->copy_stream));
  const int status = check_cuda(H, cudaStreamSynchronize(ST), "heff_apply_shard_host sync");
  if (trace && !status) {
    float t[5];
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], te[i], te[i + 1]);
    fprintf(stderr, "[tnb e2e rank %d] barrier0 %.3f  h2d+forward %.3f  barrier1 %.3f  matvec (d2h overlapped) %.3f  barrier2 %.3f ms\n", rank, t[0], t[1],
            t[2], t[3], t[4]);
    for (auto& e : te) if (e) cudaEventDestroy(e);
  }
  return status;
}

}  // extern "C"
