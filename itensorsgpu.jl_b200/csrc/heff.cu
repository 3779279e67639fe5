// Tier 2: fused DMRG entry points built on the contraction kernel -- H_eff*phi, environment
// updates, device-resident Lanczos, noise term.
//
// These replace the *sequences* of primitive calls that [EXT] ITensors.jl issues through the
// reference's override surface (/root/reference/src/ITensorsGPU.jl:32-43) for
// `product(::ProjMPO, ::ITensor)`, `makeL!/makeR!`, `KrylovKit.eigsolve` and `noiseterm`;
// reference call sites: examples/dmrg.jl:25, test/dmrg.jl:27,75.  Compared with the
// reference path there is no per-contraction allocation + zero fill
// (src/tensor/cudense.jl:62-72), no descriptor/plan rebuild per call (cudense.jl:255-294)
// and no blocking D2H per dot/norm (cudense.jl:25-27): temporaries live in the handle's
// arena and every Lanczos scalar stays on the device until one final readback.
#include "tnb_arith.cuh"
#include "tnb_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace tnb {

// ------------------------------------------------------------------------------------
// H_eff steps 2+3 fused: T3[r,l',s1',s2',c] = sum_{a,s1,s2} T1[s1,s2,r,l',a] * What[(s1,s2,a),(s1',s2',c)],
// What = sum_b W1[a,s1,s1',b] W2[b,s2,s2',c].  K = w d^2 = 20: pure HBM streaming (read T1 once, write T3 once,
// 2 * 8 * d^2 w chi^2 bytes), one "pixel" (r,l') per thread, What broadcast from shared memory.
// The reference issues these as two cuTENSOR contractions with a full-size intermediate
// (src/tensor/cudense.jl:238-331 twice + the allocation/zero-fill at :62-72).
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void what_kernel(const T* __restrict__ W1, const T* __restrict__ W2, T* What, int wl, int wm, int wr, int d1, int d2) {
  const int NI = wl * d1 * d2, NO = d1 * d2 * wr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NI * NO; e += gridDim.x * blockDim.x) {
    const int i = e / NO, o = e % NO;
    const int s1 = i % d1, s2 = (i / d1) % d2, a = i / (d1 * d2);
    const int s1p = o % d1, s2p = (o / d1) % d2, c = o / (d1 * d2);
    T acc = a_zero<T>();
    for (int b = 0; b < wm; ++b)
      acc = a_add(acc, a_mul(W1[a + wl * (s1 + d1 * (s1p + d1 * b))], W2[b + wm * (s2 + d2 * (s2p + d2 * c))]));
    What[e] = acc;
  }
}

template <bool CPLX, int D, int W>
__global__ void __launch_bounds__(256) heff23_kernel(const typename ElemT<CPLX>::T* __restrict__ T1,
                                                      const typename ElemT<CPLX>::T* __restrict__ What,
                                                      typename ElemT<CPLX>::T* __restrict__ T3, long long npix) {
  using T = typename ElemT<CPLX>::T;
  constexpr int Q = D * D, NI = W * Q, NO = Q * W;
  __shared__ __align__(16) T Ws[NI * NO];
  for (int e = threadIdx.x; e < NI * NO; e += 256) Ws[e] = What[e];
  __syncthreads();
  // exactly one pixel per thread: with a pixel loop the compiler hoists all NI*NO What loads out of it and spills
  const long long x = (long long)blockIdx.x * 256 + threadIdx.x;
  if (x < npix) {
    T in[NI];
#pragma unroll
    for (int a = 0; a < W; ++a) {
      const T* src = T1 + (size_t)Q * (x + npix * a);
      if constexpr (!CPLX && (Q % 2 == 0)) {
#pragma unroll
        for (int q = 0; q < Q; q += 2) {
          const double2 v = *reinterpret_cast<const double2*>(src + q);
          in[a * Q + q] = v.x; in[a * Q + q + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) in[a * Q + q] = src[q];
      }
    }
    T acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = a_zero<T>();
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        if constexpr (CPLX) acc[o] = a_add(acc[o], a_mul(in[i], Ws[i * NO + o]));
        else acc[o] = fma(in[i], Ws[i * NO + o], acc[o]);
      }
#pragma unroll
    for (int o = 0; o < NO; ++o) T3[x + npix * o] = acc[o];
  }
}

// returns 1 if the fused kernel handled steps 2+3 (t_in -> t_out), 0 if the shape has no instantiation
static int heff23_fused(Handle* h, int dtype, const tnb_bond_dims* d, int64_t clp, const void* W1, const void* W2,
                        const void* t_in, void* t_out, void* what, cudaStream_t st) {
  static const bool off = getenv("TNB_HEFF23") && !strcmp(getenv("TNB_HEFF23"), "off");
  if (off || d->d1 != d->d2 || d->wL != d->wR) return 0;
  const long long npix = (long long)d->chiR * clp;
  if ((npix + 255) / 256 > 2147483647LL) return 0;
  const int grid = (int)((npix + 255) / 256);
  const bool c = dtype == TNB_C128;
  const int D = d->d1, W = d->wL;
  if (!((D == 2 && (W == 5 || W == 3)) || (D == 3 && W == 5 && !c))) return 0;
  const int ne = (W * D * D) * (D * D * W);
  if (c) what_kernel<double2><<<(ne + 255) / 256, 256, 0, st>>>((const double2*)W1, (const double2*)W2, (double2*)what, d->wL, d->wM, d->wR, D, D);
  else what_kernel<double><<<(ne + 255) / 256, 256, 0, st>>>((const double*)W1, (const double*)W2, (double*)what, d->wL, d->wM, d->wR, D, D);
#define TNB_H23(C, DD, WW) heff23_kernel<C, DD, WW><<<grid, 256, 0, st>>>((const typename ElemT<C>::T*)t_in, (const typename ElemT<C>::T*)what, (typename ElemT<C>::T*)t_out, npix)
  if (D == 2 && W == 5) { if (c) TNB_H23(true, 2, 5); else TNB_H23(false, 2, 5); }
  else if (D == 2 && W == 3) { if (c) TNB_H23(true, 2, 3); else TNB_H23(false, 2, 3); }
  else TNB_H23(false, 3, 5);
#undef TNB_H23
  h->launches += 2;
  return 1;
}

enum { mL = 0, mS1, mS2, mR, mLp, mA, mS1p, mB, mS2p, mC, mRp, mLpp, mS1pp, mS2pp, mRpp };

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// Temporaries of one matvec / noise term / environment update are 2 x (base * w) elements: 2 x 2.7 GB at C3, but
// 2 x 64 GB at C5 (chi = 8192, w = 30).  Above `g_ws_limit` bytes the work is cut into G slabs of an index that is a
// FREE (output) index of the whole chain of contractions -- the output bond l' for H_eff, so that every slab is an
// independent, full-efficiency H_eff with L taken as a strided window and the result written as a strided window of
// the output vector (the same code path as the multi-GPU shard) -- or over a contracted index with beta = 1
// accumulation in the last step (environment updates, noise term).  This supersedes the reference's dead out-of-core
// attempt (/root/reference/src/tensor/dense.jl:50-193).
static size_t g_ws_limit = (size_t)40 << 30;

static int pick_chunks(size_t pair_bytes, int64_t extent) {
  int G = 1;
  while (pair_bytes / G > g_ws_limit && extent % (2 * G) == 0 && extent / (2 * G) >= 32) G *= 2;
  return G;
}

static size_t heff_pair_bytes(int dtype, const tnb_bond_dims* d) {
  const size_t base = (size_t)d->chiL * d->chiR * d->d1 * d->d2;
  const size_t w = std::max({d->wL, d->wM, d->wR});
  return 2 * base * w * elsize(dtype);
}

static int heff_nchunks(int dtype, const tnb_bond_dims* d) { return pick_chunks(heff_pair_bytes(dtype, d), d->chiL); }

static size_t heff_ws_bytes(int dtype, const tnb_bond_dims* d) {
  const size_t base = (size_t)d->chiL * d->chiR * d->d1 * d->d2;
  const size_t w = std::max({d->wL, d->wM, d->wR});
  const int G = heff_nchunks(dtype, d);
  return 2 * al256(base / G * w * elsize(dtype));
}

// out <- (((phi*L)*W1)*W2)*R using two ping-pong temporaries t0,t1 (each base*max(w) elements)
// `clp` = extent of the output bond l' held by this call: chiL normally, chiL/G when the
// output bond is sharded over G GPUs (L is then the slab L[:, l'_shard, :]).
struct PeerOut {           // fused all-gather of the sharded result (step 4 epilogue stores to every GPU)
  int npeer = 0;
  void* base[TNB_MAX_PEERS];   // full-vector buffers [chiL,d1,d2,chiR] of every rank, already offset to this rank's l' slab
};

static int heff_core(Handle* h, int dtype, const tnb_bond_dims* d, int64_t clp, const void* L, const void* W1,
                     const void* W2, const void* R, const void* phi, void* out, void* t0, void* t1,
                     cudaStream_t st, const PeerOut* peers = nullptr, int64_t lp_stored = 0, const void** t3_only = nullptr) {
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2, wl = d->wL, wm = d->wM, wr = d->wR;
  // step 4 writes out[l'_slab, s1', s2', r'] -- dense, or (fused gather) a strided window of the full vector
  auto step4 = [&](const void* T3) -> int {
    int64_t ea[] = {cr, clp, d1, d2, wr}; int32_t ma[] = {mR, mLp, mS1p, mS2p, mC};
    int64_t eb[] = {cr, cr, wr};          int32_t mb[] = {mR, mRp, mC};
    int64_t ec[] = {clp, d1, d2, cr};     int32_t mc[] = {mLp, mS1p, mS2p, mRp};
    if (peers && peers->npeer > 0) {
      int64_t sc[] = {1, cl, cl * d1, cl * d1 * d2};
      return contract_impl_ex(h, dtype, 5, ea, ma, T3, 3, eb, mb, R, 4, ec, mc, peers->base[0], nullptr, nullptr, 0, st, sc,
                              peers->base, peers->npeer);
    }
    return contract_impl(h, dtype, 5, ea, ma, T3, 3, eb, mb, R, 4, ec, mc, out, nullptr, nullptr, 0, st);
  };
  {  // 1. T1[s1,s2,r,l',a] = phi[l,s1,s2,r] L[l,l',a]
    int64_t ea[] = {cl, d1, d2, cr}; int32_t ma[] = {mL, mS1, mS2, mR};
    int64_t eb[] = {cl, clp, wl};    int32_t mb[] = {mL, mLp, mA};
    int64_t ec[] = {d1, d2, cr, clp, wl}; int32_t mc[] = {mS1, mS2, mR, mLp, mA};
    if (lp_stored > 0 && lp_stored != clp) {      // L is the window [:, l'_0 : l'_0 + clp, :] of a stored L[cl, lp_stored, wl]
      int64_t sb[] = {1, cl, cl * lp_stored};
      TNB_TRY(contract_impl_ex(h, dtype, 4, ea, ma, phi, 3, eb, mb, L, 5, ec, mc, t0, nullptr, nullptr, 0, st, nullptr, nullptr, 0,
                               nullptr, sb));
    } else {
      TNB_TRY(contract_impl(h, dtype, 4, ea, ma, phi, 3, eb, mb, L, 5, ec, mc, t0, nullptr, nullptr, 0, st));
    }
  }
  // 2+3 fused (one streaming pass) when the shape has an instantiation
  if (heff23_fused(h, dtype, d, clp, W1, W2, t0, t1, h->what, st)) {
    TNB_TRY(check_cuda(h, cudaGetLastError(), "heff23"));
    if (t3_only) { *t3_only = t1; return TNB_OK; }
    return step4(t1);
  }
  {  // 2. T2[s2,r,l',s1',b] = T1 W1[a,s1,s1',b]
    int64_t ea[] = {d1, d2, cr, clp, wl}; int32_t ma[] = {mS1, mS2, mR, mLp, mA};
    int64_t eb[] = {wl, d1, d1, wm};      int32_t mb[] = {mA, mS1, mS1p, mB};
    int64_t ec[] = {d2, cr, clp, d1, wm}; int32_t mc[] = {mS2, mR, mLp, mS1p, mB};
    TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W1, 5, ec, mc, t1, nullptr, nullptr, 0, st));
  }
  {  // 3. T3[r,l',s1',s2',c] = T2 W2[b,s2,s2',c]
    int64_t ea[] = {d2, cr, clp, d1, wm}; int32_t ma[] = {mS2, mR, mLp, mS1p, mB};
    int64_t eb[] = {wm, d2, d2, wr};      int32_t mb[] = {mB, mS2, mS2p, mC};
    int64_t ec[] = {cr, clp, d1, d2, wr}; int32_t mc[] = {mR, mLp, mS1p, mS2p, mC};
    TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 4, eb, mb, W2, 5, ec, mc, t0, nullptr, nullptr, 0, st));
  }
  // 4. out[l',s1',s2',r'] = T3 R[r,r',c]
  if (t3_only) { *t3_only = t0; return TNB_OK; }
  return step4(t0);
}

// full H_eff*phi on one GPU; cut over the output bond when the temporaries would exceed the workspace limit
static int heff_apply_any(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1, const void* W2,
                          const void* R, const void* phi, void* out, void* t0, void* t1, cudaStream_t st) {
  const int G = heff_nchunks(dtype, d);
  if (G == 1) return heff_core(h, dtype, d, d->chiL, L, W1, W2, R, phi, out, t0, t1, st);
  const int64_t cl = d->chiL, clp = cl / G;
  const size_t es = elsize(dtype);
  for (int j = 0; j < G; ++j) {
    PeerOut po;
    po.npeer = 1;
    po.base[0] = (char*)out + (size_t)j * clp * es;
    TNB_TRY(heff_core(h, dtype, d, clp, (const char*)L + (size_t)j * clp * cl * es, W1, W2, R, phi, nullptr, t0, t1, st, &po, cl));
  }
  return TNB_OK;
}

static int check_dims(Handle* h, const tnb_bond_dims* d) {
  if (!d) return set_err(h, TNB_ERR_BAD_ARG, "null dims");
  if (d->chiL < 1 || d->chiR < 1 || d->d1 < 1 || d->d2 < 1 || d->wL < 1 || d->wM < 1 || d->wR < 1)
    return set_err(h, TNB_ERR_BAD_ARG, "bond dims must be >= 1");
  return TNB_OK;
}

int heff_apply_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                    const void* W2, const void* R, const void* phi, void* out, cudaStream_t st) {
  TNB_TRY(check_dims(h, d));
  ws_reset(h);
  const size_t need = heff_ws_bytes(dtype, d);
  TNB_TRY(ws_require(h, need));
  void *t0, *t1;
  TNB_TRY(ws_alloc(h, need / 2, &t0));
  TNB_TRY(ws_alloc(h, need / 2, &t1));
  return heff_apply_any(h, dtype, d, L, W1, W2, R, phi, out, t0, t1, st);
}

int heff_apply_shard_impl(Handle* h, int dtype, const tnb_bond_dims* d, int64_t clp, const void* Lslab,
                          const void* W1, const void* W2, const void* R, const void* phi, void* out,
                          cudaStream_t st) {
  TNB_TRY(check_dims(h, d));
  if (clp < 1 || clp > d->chiL) return set_err(h, TNB_ERR_BAD_ARG, "heff_apply_shard: bad shard extent");
  ws_reset(h);
  const size_t base = (size_t)clp * d->chiR * d->d1 * d->d2;
  const size_t w = std::max({d->wL, d->wM, d->wR});
  const size_t half = al256(base * w * elsize(dtype));
  TNB_TRY(ws_require(h, 2 * half));
  void *t0, *t1;
  TNB_TRY(ws_alloc(h, half, &t0));
  TNB_TRY(ws_alloc(h, half, &t1));
  return heff_core(h, dtype, d, clp, Lslab, W1, W2, R, phi, out, t0, t1, st);
}

// Device-side barrier over peer-mapped flag arrays: lane g publishes `epoch` into rank g's flags[rank] (system-scope
// release) and then waits until its own flags[g] has reached `epoch` (acquire).  Runs after the preceding work in
// stream order, so every peer store of this rank is performed before its flag.  A bounded spin (about 60 s)
// protects against a rank that never arrives: the kernel then records a failure in err[0] instead of hanging the GPU.
struct FlagPtrs { unsigned long long* p[TNB_MAX_PEERS]; };

__global__ void peer_barrier_kernel(FlagPtrs flags, int rank, int world, unsigned long long epoch, double* err) {
  const int g = threadIdx.x;
  if (g >= world) return;
  __threadfence_system();
  unsigned long long* remote = flags.p[g] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
  const unsigned long long* mine = flags.p[rank] + g;
  const long long t0 = clock64();
  unsigned long long v = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > 120000000000LL) { err[0] = 1.0; break; }      // ~60 s at 1.9 GHz
    __nanosleep(200);
  }
}

static int peer_barrier(Handle* h, void* const* flag_peers, int rank, int world, unsigned long long epoch, cudaStream_t st) {
  FlagPtrs fp;
  for (int g = 0; g < TNB_MAX_PEERS; ++g) fp.p[g] = g < world ? (unsigned long long*)flag_peers[g] : nullptr;
  peer_barrier_kernel<<<1, 32, 0, st>>>(fp, rank, world, epoch, h->scal + 200);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "peer barrier");
}

int comm_barrier(Handle* h, cudaStream_t st) {
  if (!h->comm.on) return set_err(h, TNB_ERR_BAD_ARG, "no peer group on this handle (tnb_comm_init)");
  if (h->comm.world == 1) return TNB_OK;
  h->comm.epoch++;
  return peer_barrier(h, (void* const*)h->comm.flags, h->comm.rank, h->comm.world, h->comm.epoch, st);
}

int comm_allgather(Handle* h, void* const* bufs, size_t off, size_t bytes, cudaStream_t st) {
  if (!h->comm.on) return set_err(h, TNB_ERR_BAD_ARG, "no peer group on this handle (tnb_comm_init)");
  const int rank = h->comm.rank, world = h->comm.world;
  if (world == 1) return TNB_OK;
  TNB_TRY(comm_barrier(h, st));      // every rank is done with the previous contents of the peer buffers
  const char* src = (const char*)bufs[rank] + off + (size_t)rank * bytes;
  for (int i = 1; i < world; ++i) {  // staggered destinations: no two ranks target the same peer at once
    const int g = (rank + i) % world;
    TNB_CUDA(h, cudaMemcpyAsync((char*)bufs[g] + off + (size_t)rank * bytes, src, bytes, cudaMemcpyDefault, st));
  }
  return comm_barrier(h, st);        // all slabs have landed everywhere
}

// Host-buffer form of the sharded matvec, device part: steps 1-3 once, then step 4 cut over r' into NC pieces.  Every
// piece stores this rank's l' slab of H*phi into ALL ranks' full-vector buffers (peer stores) and, as soon as it is
// done, its own copy of the piece starts downloading on the copy stream (a strided window of out_host, which has
// phi's layout) while the next piece is computed: the download needs no cross-rank synchronisation, because a rank
// downloads exactly what it computed itself.  Measured need: with 8 GPUs pulling on one host's memory a 64 MB
// download takes 3.7-5.8 ms, a quarter of the 21 ms matvec.
int heff_shard_fused_host_tail(Handle* h, int dtype, const tnb_bond_dims* d, int rank, int world, int64_t clp,
                               const void* Lslab, const void* W1, const void* W2, const void* R, const void* phi,
                               void* const* out_peers, void* t0, void* t1, void* out_host, cudaStream_t st) {
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2, wr = d->wR;
  const size_t es = elsize(dtype);
  const void* T3 = nullptr;
  TNB_TRY(heff_core(h, dtype, d, clp, Lslab, W1, W2, R, phi, nullptr, t0, t1, st, nullptr, 0, &T3));
  const int NC = cr >= 1024 ? 4 : 1;
  for (int i = 0; i < 2 * NC + 1 && i < 40; ++i)
    if (!h->ev[i]) TNB_CUDA(h, cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
  cudaStream_t cs = h->copy_stream;
  const int64_t rc = (cr + NC - 1) / NC;
  const size_t col = (size_t)cl * d1 * d2;                             // elements of the full vector per unit of r'
  for (int c = 0; c < NC; ++c) {
    const int64_t q0 = c * rc, q1 = std::min<int64_t>(cr, q0 + rc);
    if (q1 <= q0) continue;
    int64_t ea[] = {cr, clp, d1, d2, wr}; int32_t ma[] = {mR, mLp, mS1p, mS2p, mC};
    int64_t eb[] = {cr, q1 - q0, wr};     int32_t mb[] = {mR, mRp, mC};
    int64_t sb[] = {1, cr, cr * cr};
    int64_t ec[] = {clp, d1, d2, q1 - q0}; int32_t mc[] = {mLp, mS1p, mS2p, mRp};
    int64_t sc[] = {1, cl, cl * d1, cl * d1 * d2};
    void* bases[TNB_MAX_PEERS];
    for (int g = 0; g < world; ++g) bases[g] = (char*)out_peers[g] + ((size_t)rank * clp + (size_t)q0 * col) * es;
    TNB_TRY(contract_impl_ex(h, dtype, 5, ea, ma, T3, 3, eb, mb, (const char*)R + (size_t)q0 * cr * es, 4, ec, mc, bases[0], nullptr,
                             nullptr, 0, st, sc, bases, world, nullptr, sb));
    TNB_CUDA(h, cudaEventRecord(h->ev[c], st));
    TNB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[c], 0));
    TNB_CUDA(h, cudaMemcpy2DAsync((char*)out_host + ((size_t)rank * clp + (size_t)q0 * col) * es, (size_t)cl * es,
                                  (const char*)out_peers[rank] + ((size_t)rank * clp + (size_t)q0 * col) * es, (size_t)cl * es,
                                  (size_t)clp * es, (size_t)d1 * d2 * (q1 - q0), cudaMemcpyDeviceToHost, cs));
  }
  return TNB_OK;
}

size_t heff_shard_ws_bytes(int dtype, const tnb_bond_dims* d, int64_t clp) {
  const size_t base = (size_t)clp * d->chiR * d->d1 * d->d2;
  const size_t w = std::max({d->wL, d->wM, d->wR});
  return 2 * al256(base * w * elsize(dtype));
}

int heff_shard_fused_core(Handle* h, int dtype, const tnb_bond_dims* d, int rank, int world, int64_t clp,
                          const void* Lslab, const void* W1, const void* W2, const void* R, const void* phi,
                          void* const* out_peers, void* t0, void* t1, cudaStream_t st) {
  PeerOut po;
  po.npeer = world;
  for (int g = 0; g < world; ++g) po.base[g] = (char*)out_peers[g] + (size_t)rank * clp * elsize(dtype);
  return heff_core(h, dtype, d, clp, Lslab, W1, W2, R, phi, nullptr, t0, t1, st, &po);
}

int heff_apply_shard_fused_impl(Handle* h, int dtype, const tnb_bond_dims* d, int rank, int world, int64_t clp,
                                const void* Lslab, const void* W1, const void* W2, const void* R, const void* phi,
                                void* const* out_peers, void* const* flag_peers, unsigned long long epoch, cudaStream_t st) {
  TNB_TRY(check_dims(h, d));
  if (world < 1 || world > TNB_MAX_PEERS || rank < 0 || rank >= world) return set_err(h, TNB_ERR_BAD_ARG, "heff_apply_shard_fused: rank/world");
  if (clp < 1 || clp * world != d->chiL) return set_err(h, TNB_ERR_BAD_ARG, "heff_apply_shard_fused: chiL must be world * lp_extent");
  ws_reset(h);
  const size_t hw = heff_shard_ws_bytes(dtype, d, clp);
  TNB_TRY(ws_require(h, hw));
  void *t0, *t1;
  TNB_TRY(ws_alloc(h, hw / 2, &t0));
  TNB_TRY(ws_alloc(h, hw / 2, &t1));
  TNB_TRY(heff_shard_fused_core(h, dtype, d, rank, world, clp, Lslab, W1, W2, R, phi, out_peers, t0, t1, st));
  return peer_barrier(h, flag_peers, rank, world, epoch, st);
}

// ------------------------------------------------------------------------------------
// environment updates
// ------------------------------------------------------------------------------------
int env_update_impl(Handle* h, int dtype, bool left, int64_t cl, int64_t cr, int32_t d_, int32_t wl_,
                    int32_t wr_, const void* E, const void* A, const void* W, void* Enew, cudaStream_t st) {
  const int64_t d = d_, wl = wl_, wr = wr_;
  if (cl < 1 || cr < 1 || d < 1 || wl < 1 || wr < 1) return set_err(h, TNB_ERR_BAD_ARG, "env_update: dims");
  ws_reset(h);
  const size_t es = elsize(dtype);
  // the two temporaries are cut over the bra bond of the OLD environment (l' for makeL!, r' for makeR!), which the
  // last contraction sums over: chunk j contributes with beta = 1 (C5: 2 x 32 GB otherwise)
  const int64_t ce = left ? cl : cr;
  const int G = pick_chunks(2 * (size_t)cl * cr * d * std::max(wl, wr) * es, ce);
  const int64_t cc = ce / G;
  const size_t n1 = al256((size_t)(left ? cc * cr : cl * cc) * d * std::max(wl, wr) * es);
  TNB_TRY(ws_require(h, 2 * n1));
  void *t0, *t1;
  TNB_TRY(ws_alloc(h, n1, &t0));
  TNB_TRY(ws_alloc(h, n1, &t1));
  double one[2] = {1.0, 0.0};
  // labels: l, l', a, s, r, s', b, r'
  enum { l = 0, lp, a, s, r, sp, b, rp };
  for (int j = 0; j < G; ++j) {
    const void* beta = j ? one : nullptr;
    if (left) {
      {  // T1[l'_c,a,s,r] = L[l,l'_c,a] A[l,s,r]      (L window over l')
        int64_t ea[] = {cl, cc, wl}; int32_t ma[] = {l, lp, a};
        int64_t sa[] = {1, cl, cl * cl};
        int64_t eb[] = {cl, d, cr};  int32_t mb[] = {l, s, r};
        int64_t ec[] = {cc, wl, d, cr}; int32_t mc[] = {lp, a, s, r};
        TNB_TRY(contract_impl_ex(h, dtype, 3, ea, ma, (const char*)E + (size_t)j * cc * cl * es, 3, eb, mb, A, 4, ec, mc, t0, nullptr,
                                 nullptr, 0, st, nullptr, nullptr, 0, sa, nullptr));
      }
      {  // T2[l'_c,r,s',b] = T1 W[a,s,s',b]
        int64_t ea[] = {cc, wl, d, cr}; int32_t ma[] = {lp, a, s, r};
        int64_t eb[] = {wl, d, d, wr};  int32_t mb[] = {a, s, sp, b};
        int64_t ec[] = {cc, cr, d, wr}; int32_t mc[] = {lp, r, sp, b};
        TNB_TRY(contract_impl(h, dtype, 4, ea, ma, t0, 4, eb, mb, W, 4, ec, mc, t1, nullptr, nullptr, 0, st));
      }
      {  // Lnew[r,r',b] (+)= T2 conj(A)[l'_c,s',r']   (A window over its first mode)
        int64_t ea[] = {cc, cr, d, wr}; int32_t ma[] = {lp, r, sp, b};
        int64_t eb[] = {cc, d, cr};     int32_t mb[] = {lp, sp, rp};
        int64_t sb[] = {1, cl, cl * d};
        int64_t ec[] = {cr, cr, wr};    int32_t mc[] = {r, rp, b};
        TNB_TRY(contract_impl_ex(h, dtype, 4, ea, ma, t1, 3, eb, mb, (const char*)A + (size_t)j * cc * es, 3, ec, mc, Enew, nullptr,
                                 beta, TNB_CONJ_B, st, nullptr, nullptr, 0, nullptr, sb));
      }
    } else {
      // here E = R[r,r',c] with c = right MPO bond (wr), A[l,s,r], W[a,s,s',c]; labels b == c
      {  // T1[r'_c,c,l,s] = R[r,r'_c,c] A[l,s,r]      (R window over r')
        int64_t ea[] = {cr, cc, wr}; int32_t ma[] = {r, rp, b};
        int64_t sa[] = {1, cr, cr * cr};
        int64_t eb[] = {cl, d, cr};  int32_t mb[] = {l, s, r};
        int64_t ec[] = {cc, wr, cl, d}; int32_t mc[] = {rp, b, l, s};
        TNB_TRY(contract_impl_ex(h, dtype, 3, ea, ma, (const char*)E + (size_t)j * cc * cr * es, 3, eb, mb, A, 4, ec, mc, t0, nullptr,
                                 nullptr, 0, st, nullptr, nullptr, 0, sa, nullptr));
      }
      {  // T2[r'_c,l,a,s'] = T1 W[a,s,s',c]
        int64_t ea[] = {cc, wr, cl, d}; int32_t ma[] = {rp, b, l, s};
        int64_t eb[] = {wl, d, d, wr};  int32_t mb[] = {a, s, sp, b};
        int64_t ec[] = {cc, cl, wl, d}; int32_t mc[] = {rp, l, a, sp};
        TNB_TRY(contract_impl(h, dtype, 4, ea, ma, t0, 4, eb, mb, W, 4, ec, mc, t1, nullptr, nullptr, 0, st));
      }
      {  // Rnew[l,l',a] (+)= T2 conj(A)[l',s',r'_c]   (A slab over its last mode: contiguous)
        int64_t ea[] = {cc, cl, wl, d}; int32_t ma[] = {rp, l, a, sp};
        int64_t eb[] = {cl, d, cc};     int32_t mb[] = {lp, sp, rp};
        int64_t ec[] = {cl, cl, wl};    int32_t mc[] = {l, lp, a};
        TNB_TRY(contract_impl(h, dtype, 4, ea, ma, t1, 3, eb, mb, (const char*)A + (size_t)j * cc * cl * d * es, 3, ec, mc, Enew,
                              nullptr, beta, TNB_CONJ_B, st));
      }
    }
  }
  return TNB_OK;
}

// ------------------------------------------------------------------------------------
// device-resident Lanczos
// ------------------------------------------------------------------------------------
// Scalar pool layout (doubles, in h->scal): [16+j] alpha_j, [32+j] beta_j (residual norm after
// step j), [48 + 2*i] complex overlap scratch, [64] lambda, [65] k_eff, [66] last-coefficient
// residual |beta_k y_k|, [72+j] Ritz coefficients y_j, [90] nrm scratch.
constexpr int S_ALPHA = 16, S_BETA = 32, S_OVL = 48, S_LAM = 64, S_KEFF = 65, S_RES = 66, S_Y = 72, S_NRM = 90;
constexpr int KRYLOV_MAX = 12;

// y <- y - c*x with c = scal[ci] (+ i*scal[ci+1] if complex)
template <bool CPLX>
__global__ void __launch_bounds__(256) axmy_dev_kernel(double2* y, const double2* __restrict__ x, long long n2,
                                                       const double* scal, int ci, int tail) {
  const double cr = scal[ci], cim = CPLX ? scal[ci + 1] : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    double2 a = x[i], b = y[i];
    if (CPLX) { b.x -= cr * a.x - cim * a.y; b.y -= cr * a.y + cim * a.x; }
    else { b.x -= cr * a.x; b.y -= cr * a.y; }
    y[i] = b;
  }
  if (!CPLX && tail && blockIdx.x == 0 && threadIdx.x == 0) {
    double* yt = (double*)(y + n2); const double* xt = (const double*)(x + n2);
    yt[0] -= cr * xt[0];
  }
}

// y <- x * (scal[ci] > tol ? 1/scal[ci] : 0)
__global__ void __launch_bounds__(256) scale_inv_dev_kernel(double2* y, const double2* __restrict__ x, long long n2,
                                                            const double* scal, int ci, double tol, int tail) {
  const double b = scal[ci];
  const double f = b > tol ? 1.0 / b : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    double2 a = x[i];
    a.x *= f; a.y *= f;
    y[i] = a;
  }
  if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
    double* yt = (double*)(y + n2); const double* xt = (const double*)(x + n2);
    yt[0] = xt[0] * f;
  }
}

// out <- sum_j scal[S_Y+j] * V_j   (Ritz vector), V_j = V + j*stride2 (in double2 units)
__global__ void __launch_bounds__(256) ritz_combine_kernel(double2* out, const double2* __restrict__ V,
                                                           long long stride2, long long n2, int k,
                                                           const double* scal, int tail) {
  double y[KRYLOV_MAX];
  for (int j = 0; j < k; ++j) y[j] = scal[S_Y + j];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    double2 acc = make_double2(0.0, 0.0);
    for (int j = 0; j < k; ++j) { const double2 v = V[j * stride2 + i]; acc.x += y[j] * v.x; acc.y += y[j] * v.y; }
    out[i] = acc;
  }
  if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
    double acc = 0;
    for (int j = 0; j < k; ++j) acc += y[j] * ((const double*)(V + j * stride2 + n2))[0];
    ((double*)(out + n2))[0] = acc;
  }
}

// Lowest eigenpair of the k x k symmetric tridiagonal (alpha, beta) by cyclic Jacobi -- one
// thread; k <= 12.  k_eff = number of Krylov vectors actually spanned (first beta <= tol cuts).
__global__ void ritz_kernel(double* scal, int k, double tol) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int ke = k;
  for (int j = 0; j < k - 1; ++j) if (!(scal[S_BETA + j] > tol)) { ke = j + 1; break; }
  double T[KRYLOV_MAX][KRYLOV_MAX], U[KRYLOV_MAX][KRYLOV_MAX];
  for (int i = 0; i < ke; ++i) for (int j = 0; j < ke; ++j) { T[i][j] = 0.0; U[i][j] = i == j ? 1.0 : 0.0; }
  for (int i = 0; i < ke; ++i) T[i][i] = scal[S_ALPHA + i];
  for (int i = 0; i < ke - 1; ++i) T[i][i + 1] = T[i + 1][i] = scal[S_BETA + i];
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < ke; ++i) for (int j = i + 1; j < ke; ++j) off += T[i][j] * T[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < ke; ++p) for (int q = p + 1; q < ke; ++q) {
      const double apq = T[p][q];
      if (apq == 0.0) continue;
      const double tau = (T[q][q] - T[p][p]) / (2.0 * apq);
      const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
      for (int i = 0; i < ke; ++i) { const double a = T[i][p], b = T[i][q]; T[i][p] = c * a - s * b; T[i][q] = s * a + c * b; }
      for (int i = 0; i < ke; ++i) { const double a = T[p][i], b = T[q][i]; T[p][i] = c * a - s * b; T[q][i] = s * a + c * b; }
      for (int i = 0; i < ke; ++i) { const double a = U[i][p], b = U[i][q]; U[i][p] = c * a - s * b; U[i][q] = s * a + c * b; }
    }
  }
  int best = 0;
  for (int i = 1; i < ke; ++i) if (T[i][i] < T[best][best]) best = i;
  double nrm = 0;
  for (int i = 0; i < ke; ++i) nrm += U[i][best] * U[i][best];
  nrm = sqrt(nrm);
  for (int i = 0; i < k; ++i) scal[S_Y + i] = i < ke ? U[i][best] / nrm : 0.0;
  scal[S_LAM] = T[best][best];
  scal[S_KEFF] = (double)ke;
  scal[S_RES] = fabs(scal[S_BETA + ke - 1] * U[ke - 1][best] / nrm);
}

static int vec_grid(Handle* h, long long n2) {
  return (int)std::max<long long>(1, std::min<long long>((n2 + 1023) / 1024, (long long)h->num_sms * 8));
}

int lanczos_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                 const void* W2, const void* R, void* phi, int krylovdim, int maxiter, double tol,
                 double* energy, int* n_matvec, cudaStream_t st, const ShardCtx* sc) {
  TNB_TRY(check_dims(h, d));
  if (krylovdim < 1 || krylovdim > KRYLOV_MAX) return set_err(h, TNB_ERR_BAD_ARG, "lanczos: krylovdim must be in 1..%d", KRYLOV_MAX);
  if (maxiter < 1) return set_err(h, TNB_ERR_BAD_ARG, "lanczos: maxiter < 1");
  const bool cplx = dtype == TNB_C128;
  const size_t es = elsize(dtype);
  const long long n = (long long)d->chiL * d->chiR * d->d1 * d->d2;
  const long long nd = cplx ? 2 * n : n;      // length in doubles
  const long long n2 = nd / 2;
  const int tail = (int)(nd & 1);
  const size_t vbytes = al256((size_t)n * es);
  ws_reset(h);
  const size_t hw = sc ? heff_shard_ws_bytes(dtype, d, sc->clp) : heff_ws_bytes(dtype, d);
  TNB_TRY(ws_require(h, hw + (size_t)(krylovdim + 1) * vbytes + 1024));
  void *t0, *t1, *Vb, *w = nullptr;
  TNB_TRY(ws_alloc(h, hw / 2, &t0));
  TNB_TRY(ws_alloc(h, hw / 2, &t1));
  TNB_TRY(ws_alloc(h, (size_t)krylovdim * vbytes, &Vb));
  if (!sc) TNB_TRY(ws_alloc(h, vbytes, &w));
  // sharded matvec: every rank must have finished with the peer-visible result buffers of earlier calls
  if (sc) TNB_TRY(comm_barrier(h, st));
  const long long stride2 = (long long)(vbytes / 16);
  auto V = [&](int j) { return (void*)((char*)Vb + (size_t)j * vbytes); };
  const int grid = vec_grid(h, n2);
  double* scal = h->scal;
  int nmv = 0;
  for (int it = 0; it < maxiter; ++it) {
    // v1 = phi / ||phi||
    TNB_TRY(nrm2_impl(h, dtype, n, phi, scal + S_NRM, st));
    scale_inv_dev_kernel<<<grid, 256, 0, st>>>((double2*)V(0), (const double2*)phi, n2, scal, S_NRM, 0.0, tail);
    h->launches++;
    for (int j = 0; j < krylovdim; ++j) {
      if (sc) {
        // slab of H*v_j computed here, stored by the step-4 epilogue into EVERY rank's buffer (NVLink peer stores);
        // buffers alternate so that a fast rank's next matvec cannot overwrite a vector a slow rank still works on
        void* const* ob = sc->out[nmv & 1];
        TNB_TRY(heff_shard_fused_core(h, dtype, d, h->comm.rank, h->comm.world, sc->clp, L, W1, W2, R, V(j), ob, t0, t1, st));
        TNB_TRY(comm_barrier(h, st));
        w = ob[h->comm.rank];
      } else {
        TNB_TRY(heff_apply_any(h, dtype, d, L, W1, W2, R, V(j), w, t0, t1, st));
      }
      ++nmv;
      // alpha_j = Re <v_j, w>
      TNB_TRY(dot_impl(h, dtype, n, V(j), w, scal + S_OVL, st));
      TNB_CUDA(h, cudaMemcpyAsync(scal + S_ALPHA + j, scal + S_OVL, sizeof(double), cudaMemcpyDeviceToDevice, st));
      // w -= <v_i, w> v_i for all i <= j (modified Gram-Schmidt; covers the alpha_j v_j and
      // beta_{j-1} v_{j-1} terms of the three-term recurrence, then a second pass = full reorth.)
      for (int pass = 0; pass < 2; ++pass) {
        for (int i = j; i >= 0; --i) {
          if (!(pass == 0 && i == j)) TNB_TRY(dot_impl(h, dtype, n, V(i), w, scal + S_OVL, st));
          if (cplx) axmy_dev_kernel<true><<<grid, 256, 0, st>>>((double2*)w, (const double2*)V(i), n2, scal, S_OVL, 0);
          else axmy_dev_kernel<false><<<grid, 256, 0, st>>>((double2*)w, (const double2*)V(i), n2, scal, S_OVL, tail);
          h->launches++;
        }
      }
      TNB_TRY(nrm2_impl(h, dtype, n, w, scal + S_BETA + j, st));
      if (j + 1 < krylovdim) {
        scale_inv_dev_kernel<<<grid, 256, 0, st>>>((double2*)V(j + 1), (const double2*)w, n2, scal, S_BETA + j, tol, tail);
        h->launches++;
      }
    }
    ritz_kernel<<<1, 32, 0, st>>>(scal, krylovdim, tol);
    ritz_combine_kernel<<<grid, 256, 0, st>>>((double2*)phi, (const double2*)Vb, stride2, n2, krylovdim, scal, tail);
    h->launches += 2;
    TNB_CUDA(h, cudaGetLastError());
    if (it + 1 < maxiter) {
      TNB_CUDA(h, cudaMemcpyAsync(h->scal_host + S_LAM, scal + S_LAM, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      TNB_CUDA(h, cudaStreamSynchronize(st));
      if (h->scal_host[S_RES] < tol) break;
    }
  }
  // normalise (y is unit and V orthonormal, so this only removes rounding drift)
  TNB_TRY(nrm2_impl(h, dtype, n, phi, scal + S_NRM, st));
  scale_inv_dev_kernel<<<grid, 256, 0, st>>>((double2*)phi, (const double2*)phi, n2, scal, S_NRM, 0.0, tail);
  h->launches++;
  TNB_CUDA(h, cudaMemcpyAsync(h->scal_host + S_LAM, scal + S_LAM, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TNB_CUDA(h, cudaStreamSynchronize(st));
  if (energy) *energy = h->scal_host[S_LAM];
  if (n_matvec) *n_matvec = nmv;
  return TNB_OK;
}

// ------------------------------------------------------------------------------------
// noise term
// ------------------------------------------------------------------------------------
int noise_term_impl(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                    const void* W2, const void* R, const void* phi, int ortho, double noise,
                    int accumulate, void* rho, void* t0, void* t1, cudaStream_t st) {
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2, wl = d->wL, wm = d->wM, wr = d->wR;
  const size_t es = elsize(dtype);
  double alpha[2] = {noise, 0.0}, one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
  // t0 / t1 hold heff_workspace_bytes / 2 each: when H_eff is cut into G slabs, the noise term is cut into G chunks of
  // an index the Gram product sums over (r for ortho left, l for ortho right) and accumulates with beta = 1
  int G = heff_nchunks(dtype, d);
  const int64_t ce = (ortho == TNB_ORTHO_LEFT) ? cr : cl;
  while (G > 1 && ce % G) G /= 2;
  if (G < heff_nchunks(dtype, d)) return set_err(h, TNB_ERR_UNSUPPORTED, "noise_term: cannot cut bond %lld into %d chunks", (long long)ce, heff_nchunks(dtype, d));
  const int64_t cc = ce / G;
  for (int j = 0; j < G; ++j) {
    const void* beta = (j || accumulate) ? one : zero;
    if (ortho == TNB_ORTHO_LEFT) {
      {  // T1[s1,s2,r_c,l',a] = phi[l,s1,s2,r_c] L          (re-ordered: (phi*L)*W1 instead of (L*W1)*phi)
        int64_t ea[] = {cl, d1, d2, cc}; int32_t ma[] = {mL, mS1, mS2, mR};
        int64_t eb[] = {cl, cl, wl};     int32_t mb[] = {mL, mLp, mA};
        int64_t ec[] = {d1, d2, cc, cl, wl}; int32_t mc[] = {mS1, mS2, mR, mLp, mA};
        TNB_TRY(contract_impl(h, dtype, 4, ea, ma, (const char*)phi + (size_t)j * cc * cl * d1 * d2 * es, 3, eb, mb, L, 5, ec, mc, t0,
                              nullptr, nullptr, 0, st));
      }
      {  // nt[s2,r_c,l',s1',b] = T1 W1
        int64_t ea[] = {d1, d2, cc, cl, wl}; int32_t ma[] = {mS1, mS2, mR, mLp, mA};
        int64_t eb[] = {wl, d1, d1, wm};     int32_t mb[] = {mA, mS1, mS1p, mB};
        int64_t ec[] = {d2, cc, cl, d1, wm}; int32_t mc[] = {mS2, mR, mLp, mS1p, mB};
        TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W1, 5, ec, mc, t1, nullptr, nullptr, 0, st));
      }
      {  // rho[l',s1',l'',s1''] (+)= noise * nt conj(nt)
        int64_t ea[] = {d2, cc, cl, d1, wm}; int32_t ma[] = {mS2, mR, mLp, mS1p, mB};
        int32_t mb[] = {mS2, mR, mLpp, mS1pp, mB};
        int64_t ec[] = {cl, d1, cl, d1};     int32_t mc[] = {mLp, mS1p, mLpp, mS1pp};
        TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 5, ea, mb, t1, 4, ec, mc, rho, alpha, beta, TNB_CONJ_B | TNB_HERM_UPPER, st));
      }
    } else {
      {  // T1[l_c,s1,s2,r',c] = phi[l_c,s1,s2,r] R     (phi window over its first mode)
        int64_t ea[] = {cc, d1, d2, cr}; int32_t ma[] = {mL, mS1, mS2, mR};
        int64_t sa[] = {1, cl, cl * d1, cl * d1 * d2};
        int64_t eb[] = {cr, cr, wr};     int32_t mb[] = {mR, mRp, mC};
        int64_t ec[] = {cc, d1, d2, cr, wr}; int32_t mc[] = {mL, mS1, mS2, mRp, mC};
        TNB_TRY(contract_impl_ex(h, dtype, 4, ea, ma, (const char*)phi + (size_t)j * cc * es, 3, eb, mb, R, 5, ec, mc, t0, nullptr,
                                 nullptr, 0, st, nullptr, nullptr, 0, sa, nullptr));
      }
      {  // nt[l_c,s1,s2',r',b] = T1 W2[b,s2,s2',c]   ((s2',r') in rho's own order, so that the Gram below may skip
         //                                              its strictly-lower tiles)
        int64_t ea[] = {cc, d1, d2, cr, wr}; int32_t ma[] = {mL, mS1, mS2, mRp, mC};
        int64_t eb[] = {wm, d2, d2, wr};     int32_t mb[] = {mB, mS2, mS2p, mC};
        int64_t ec[] = {cc, d1, d2, cr, wm}; int32_t mc[] = {mL, mS1, mS2p, mRp, mB};
        TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W2, 5, ec, mc, t1, nullptr, nullptr, 0, st));
      }
      {  // rho[s2',r',s2'',r''] (+)= noise * nt conj(nt)
        int64_t ea[] = {cc, d1, d2, cr, wm}; int32_t ma[] = {mL, mS1, mS2p, mRp, mB};
        int32_t mb[] = {mL, mS1, mS2pp, mRpp, mB};
        int64_t ec[] = {d2, cr, d2, cr};     int32_t mc[] = {mS2p, mRp, mS2pp, mRpp};
        TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 5, ea, mb, t1, 4, ec, mc, rho, alpha, beta, TNB_CONJ_B | TNB_HERM_UPPER, st));
      }
    }
  }
  return TNB_OK;
}

size_t heff_workspace_bytes(int dtype, const tnb_bond_dims* d) { return heff_ws_bytes(dtype, d); }
// Host-buffer H_eff*phi with the PCIe copies hidden behind the two big GEMMs: phi is uploaded in NC chunks over r
// (its slowest mode) on the copy stream while step 1 already contracts the chunks that have arrived (each chunk's
// T1 slice is a strided window, r being a middle mode of T1); steps 2+3 run once; step 4 is cut over r' (a
// strided window of R, a contiguous chunk of the result) and every finished chunk starts its download while the
// next one is being computed.  Exposed copy time: one chunk each way instead of the whole vector twice.
static int heff_host_pipelined(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1,
                               const void* W2, const void* R, const void* phi_host, void* out_host, void* dphi,
                               void* dout, void* t0, void* t1, cudaStream_t st) {
  const int64_t cl = d->chiL, cr = d->chiR, d1 = d->d1, d2 = d->d2, wl = d->wL, wm = d->wM, wr = d->wR;
  const size_t es = elsize(dtype);
  const size_t nb = (size_t)cl * cr * d1 * d2 * es;
  // chunks over r / r': the exposed copy time is one chunk each way, but smaller GEMM chunks lose efficiency -- measured at
  // chi = 4096 (e2e TFLOP/s): 4 chunks 33.59, 8 chunks 32.77, 16 chunks 31.38 (device-resident: 34.87).  TNB_HOST_CHUNKS
  // overrides the default of 4.
  static const int nc_env = getenv("TNB_HOST_CHUNKS") ? std::max(1, std::min(16, atoi(getenv("TNB_HOST_CHUNKS")))) : 4;
  const int NC = (cr >= 1024 && heff_nchunks(dtype, d) == 1) ? nc_env : 1;
  if (NC == 1) {
    TNB_CUDA(h, cudaMemcpyAsync(dphi, phi_host, nb, cudaMemcpyHostToDevice, st));
    TNB_TRY(heff_apply_any(h, dtype, d, L, W1, W2, R, dphi, dout, t0, t1, st));
    TNB_CUDA(h, cudaMemcpyAsync(out_host, dout, nb, cudaMemcpyDeviceToHost, st));
    return check_cuda(h, cudaStreamSynchronize(st), "heff_apply_host sync");
  }
  for (int i = 0; i < 2 * NC + 1; ++i)
    if (!h->ev[i]) TNB_CUDA(h, cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
  cudaStream_t cs = h->copy_stream;
  // the copy stream must not run ahead of work already queued on `st` that still uses dphi / dout
  TNB_CUDA(h, cudaEventRecord(h->ev[2 * NC], st));
  TNB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[2 * NC], 0));
  const int64_t rc = (cr + NC - 1) / NC;
  const size_t slab = (size_t)cl * d1 * d2;                 // elements of phi per unit of r
  for (int c = 0; c < NC; ++c) {
    const int64_t r0 = c * rc, r1 = std::min<int64_t>(cr, r0 + rc);
    if (r1 <= r0) continue;
    TNB_CUDA(h, cudaMemcpyAsync((char*)dphi + r0 * slab * es, (const char*)phi_host + r0 * slab * es, (r1 - r0) * slab * es,
                                cudaMemcpyHostToDevice, cs));
    TNB_CUDA(h, cudaEventRecord(h->ev[c], cs));
  }
  for (int c = 0; c < NC; ++c) {       // 1. T1[s1,s2,r in chunk,l',a] = phi[l,s1,s2,r in chunk] L[l,l',a]
    const int64_t r0 = c * rc, r1 = std::min<int64_t>(cr, r0 + rc);
    if (r1 <= r0) continue;
    TNB_CUDA(h, cudaStreamWaitEvent(st, h->ev[c], 0));
    int64_t ea[] = {cl, d1, d2, r1 - r0}; int32_t ma[] = {mL, mS1, mS2, mR};
    int64_t eb[] = {cl, cl, wl};          int32_t mb[] = {mL, mLp, mA};
    int64_t ec[] = {d1, d2, r1 - r0, cl, wl}; int32_t mc[] = {mS1, mS2, mR, mLp, mA};
    int64_t sc[] = {1, d1, d1 * d2, d1 * d2 * cr, d1 * d2 * cr * cl};
    TNB_TRY(contract_impl_ex(h, dtype, 4, ea, ma, (const char*)dphi + r0 * slab * es, 3, eb, mb, L, 5, ec, mc,
                             (char*)t0 + (size_t)r0 * d1 * d2 * es, nullptr, nullptr, 0, st, sc, nullptr, 0));
  }
  const void* T3 = t0;
  if (heff23_fused(h, dtype, d, cl, W1, W2, t0, t1, h->what, st)) {
    TNB_TRY(check_cuda(h, cudaGetLastError(), "heff23"));
    T3 = t1;
  } else {
    {
      int64_t ea[] = {d1, d2, cr, cl, wl}; int32_t ma[] = {mS1, mS2, mR, mLp, mA};
      int64_t eb[] = {wl, d1, d1, wm};     int32_t mb[] = {mA, mS1, mS1p, mB};
      int64_t ec[] = {d2, cr, cl, d1, wm}; int32_t mc[] = {mS2, mR, mLp, mS1p, mB};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t0, 4, eb, mb, W1, 5, ec, mc, t1, nullptr, nullptr, 0, st));
    }
    {
      int64_t ea[] = {d2, cr, cl, d1, wm}; int32_t ma[] = {mS2, mR, mLp, mS1p, mB};
      int64_t eb[] = {wm, d2, d2, wr};     int32_t mb[] = {mB, mS2, mS2p, mC};
      int64_t ec[] = {cr, cl, d1, d2, wr}; int32_t mc[] = {mR, mLp, mS1p, mS2p, mC};
      TNB_TRY(contract_impl(h, dtype, 5, ea, ma, t1, 4, eb, mb, W2, 5, ec, mc, t0, nullptr, nullptr, 0, st));
    }
  }
  for (int c = 0; c < NC; ++c) {       // 4. out[l',s1',s2',r' in chunk] = T3 R[r, r' in chunk, c]
    const int64_t r0 = c * rc, r1 = std::min<int64_t>(cr, r0 + rc);
    if (r1 <= r0) continue;
    int64_t ea[] = {cr, cl, d1, d2, wr}; int32_t ma[] = {mR, mLp, mS1p, mS2p, mC};
    int64_t eb[] = {cr, r1 - r0, wr};    int32_t mb[] = {mR, mRp, mC};
    int64_t sb[] = {1, cr, cr * cr};
    int64_t ec[] = {cl, d1, d2, r1 - r0}; int32_t mc[] = {mLp, mS1p, mS2p, mRp};
    TNB_TRY(contract_impl_ex(h, dtype, 5, ea, ma, T3, 3, eb, mb, (const char*)R + (size_t)r0 * cr * es, 4, ec, mc,
                             (char*)dout + r0 * slab * es, nullptr, nullptr, 0, st, nullptr, nullptr, 0, nullptr, sb));
    TNB_CUDA(h, cudaEventRecord(h->ev[NC + c], st));
    TNB_CUDA(h, cudaStreamWaitEvent(cs, h->ev[NC + c], 0));
    TNB_CUDA(h, cudaMemcpyAsync((char*)out_host + r0 * slab * es, (const char*)dout + r0 * slab * es, (r1 - r0) * slab * es,
                                cudaMemcpyDeviceToHost, cs));
  }
  TNB_CUDA(h, cudaStreamSynchronize(cs));
  return check_cuda(h, cudaStreamSynchronize(st), "heff_apply_host sync");
}

int heff_core_pub(Handle* h, int dtype, const tnb_bond_dims* d, const void* L, const void* W1, const void* W2,
                  const void* R, const void* phi, void* out, void* t0, void* t1, cudaStream_t st) {
  return heff_apply_any(h, dtype, d, L, W1, W2, R, phi, out, t0, t1, st);
}

int heff_chunks_pub(int dtype, const tnb_bond_dims* d) { return heff_nchunks(dtype, d); }
void set_ws_limit(size_t bytes) { g_ws_limit = bytes ? bytes : ((size_t)40 << 30); }
size_t get_ws_limit() { return g_ws_limit; }

}  // namespace tnb

using namespace tnb;
#define H ((Handle*)h)
#define ST ((cudaStream_t)stream)

extern "C" {

int tnb_set_workspace_limit(tnb_handle_t h, size_t bytes) {
  (void)h;                        // process-wide setting: a null handle is accepted (host-only planning queries)
  set_ws_limit(bytes);
  return TNB_OK;
}

size_t tnb_get_workspace_limit(tnb_handle_t h) {
  (void)h;
  return get_ws_limit();
}

int tnb_heff_apply(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L, const void* W1,
                   const void* W2, const void* R, const void* phi, void* out, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L || !W1 || !W2 || !R || !phi || !out) return set_err(H, TNB_ERR_BAD_ARG, "heff_apply: null pointer");
  return heff_apply_impl(H, dtype, dims, L, W1, W2, R, phi, out, ST);
}

int tnb_heff_apply_shard(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int64_t lp_extent,
                         const void* L_slab, const void* W1, const void* W2, const void* R, const void* phi,
                         void* out_slab, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L_slab || !W1 || !W2 || !R || !phi || !out_slab) return set_err(H, TNB_ERR_BAD_ARG, "heff_apply_shard: null pointer");
  return heff_apply_shard_impl(H, dtype, dims, lp_extent, L_slab, W1, W2, R, phi, out_slab, ST);
}

int tnb_heff_apply_shard_fused(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, int rank, int world,
                               int64_t lp_extent, const void* L_slab, const void* W1, const void* W2, const void* R,
                               const void* phi, void* const* out_peers, void* const* flag_peers, uint64_t epoch,
                               void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L_slab || !W1 || !W2 || !R || !phi || !out_peers || !flag_peers)
    return set_err(H, TNB_ERR_BAD_ARG, "heff_apply_shard_fused: null pointer");
  for (int g = 0; g < world && g < TNB_MAX_PEERS; ++g)
    if (!out_peers[g] || !flag_peers[g]) return set_err(H, TNB_ERR_BAD_ARG, "heff_apply_shard_fused: null peer pointer %d", g);
  return heff_apply_shard_fused_impl(H, dtype, dims, rank, world, lp_extent, L_slab, W1, W2, R, phi, out_peers, flag_peers,
                                     (unsigned long long)epoch, ST);
}

// ---- peer-mapped buffers (CUDA IPC): one process per GPU on one NVSwitch node
int tnb_peer_alloc(tnb_handle_t h, size_t bytes, void** ptr, unsigned char* ipc_handle_out) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!ptr || !ipc_handle_out || bytes == 0) return set_err(H, TNB_ERR_BAD_ARG, "peer_alloc: bad argument");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return set_err(H, TNB_ERR_ALLOC, "peer_alloc: cudaMalloc(%zu) failed", bytes); }
  cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return check_cuda(H, e, "cudaIpcGetMemHandle"); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_out, &hd, 64);
  *ptr = p;
  return TNB_OK;
}

int tnb_peer_open(tnb_handle_t h, const unsigned char* ipc_handle, void** ptr) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!ipc_handle || !ptr) return set_err(H, TNB_ERR_BAD_ARG, "peer_open: null pointer");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, 64);
  return check_cuda(H, cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int tnb_peer_close(tnb_handle_t h, void* ptr) {
  if (!h) return TNB_ERR_BAD_ARG;
  return check_cuda(H, cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
}

int tnb_peer_free(tnb_handle_t h, void* ptr) {
  if (!h) return TNB_ERR_BAD_ARG;
  return check_cuda(H, cudaFree(ptr), "cudaFree(peer buffer)");
}

// 0 = every device-side peer barrier so far completed; TNB_ERR_NO_CONVERGENCE = one timed out.  Synchronises.
int tnb_peer_status(tnb_handle_t h, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  TNB_CUDA(H, cudaMemcpyAsync(H->scal_host + 101, H->scal + 200, sizeof(double), cudaMemcpyDeviceToHost, ST));
  TNB_CUDA(H, cudaStreamSynchronize(ST));
  if (H->scal_host[101] != 0.0) return set_err(H, TNB_ERR_NO_CONVERGENCE, "peer barrier timed out (a rank never arrived)");
  return TNB_OK;
}

int tnb_heff_apply_host(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L, const void* W1,
                        const void* W2, const void* R, const void* phi_host, void* out_host, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L || !W1 || !W2 || !R || !phi_host || !out_host) return set_err(H, TNB_ERR_BAD_ARG, "heff_apply_host: null pointer");
  TNB_TRY(check_dims(H, dims));
  const size_t nb = (size_t)dims->chiL * dims->chiR * dims->d1 * dims->d2 * elsize(dtype);
  ws_reset(H);
  const size_t hw = heff_workspace_bytes(dtype, dims);
  TNB_TRY(ws_require(H, hw + 2 * al256(nb)));
  void *t0, *t1, *dphi, *dout;
  TNB_TRY(ws_alloc(H, hw / 2, &t0));
  TNB_TRY(ws_alloc(H, hw / 2, &t1));
  TNB_TRY(ws_alloc(H, nb, &dphi));
  TNB_TRY(ws_alloc(H, nb, &dout));
  return heff_host_pipelined(H, dtype, dims, L, W1, W2, R, phi_host, out_host, dphi, dout, t0, t1, ST);
}

int tnb_env_update_left(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL,
                        int32_t wR, const void* L, const void* A, const void* W, void* Lnew, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L || !A || !W || !Lnew) return set_err(H, TNB_ERR_BAD_ARG, "env_update_left: null pointer");
  return env_update_impl(H, dtype, true, chiL, chiR, d, wL, wR, L, A, W, Lnew, ST);
}

int tnb_env_update_right(tnb_handle_t h, int dtype, int64_t chiL, int64_t chiR, int32_t d, int32_t wL,
                         int32_t wR, const void* R, const void* A, const void* W, void* Rnew, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!R || !A || !W || !Rnew) return set_err(H, TNB_ERR_BAD_ARG, "env_update_right: null pointer");
  return env_update_impl(H, dtype, false, chiL, chiR, d, wL, wR, R, A, W, Rnew, ST);
}

int tnb_eigsolve_lanczos(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L, const void* W1,
                         const void* W2, const void* R, void* phi, int krylovdim, int maxiter, double tol,
                         double* energy, int* n_matvec, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L || !W1 || !W2 || !R || !phi) return set_err(H, TNB_ERR_BAD_ARG, "eigsolve_lanczos: null pointer");
  return lanczos_impl(H, dtype, dims, L, W1, W2, R, phi, krylovdim, maxiter, tol, energy, n_matvec, ST);
}

int tnb_noise_term(tnb_handle_t h, int dtype, const tnb_bond_dims* dims, const void* L, const void* W1,
                   const void* W2, const void* R, const void* phi, int ortho, double noise, int accumulate,
                   void* rho, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!L || !W1 || !W2 || !R || !phi || !rho) return set_err(H, TNB_ERR_BAD_ARG, "noise_term: null pointer");
  TNB_TRY(check_dims(H, dims));
  ws_reset(H);
  const size_t hw = heff_workspace_bytes(dtype, dims);
  TNB_TRY(ws_require(H, hw));
  void *t0, *t1;
  TNB_TRY(ws_alloc(H, hw / 2, &t0));
  TNB_TRY(ws_alloc(H, hw / 2, &t1));
  return noise_term_impl(H, dtype, dims, L, W1, W2, R, phi, ortho, noise, accumulate, rho, t0, t1, ST);
}

}  // extern "C"
