// Hermitian -> real symmetric tridiagonal reduction and the matching back-transformation.
//
// Stage 1 and 3 of the fast eigensolver behind `eigen(::Hermitian{CuDenseTensor})`
// (/root/reference/src/tensor/culinearalgebra.jl:74-108, there: cuSOLVER syevd!/heevd!).
//
// Reduction: blocked Householder (latrd/sytrd structure) on FULL Hermitian storage.  Per column i
//   K1  row-parallel: finalise the previous w, update column i with the panel so far, ||x||, larfg
//   K2  column-parallel: y = A_trail^H v as one dot product per trailing column (HBM-bound: this is
//       the n^3/3 * 8 bytes of traffic that bounds the stage), plus p = W^H v, q = V^H v
// and per panel of 64 columns one DMMA GEMM  A_trail -= [V W] [W V]^H.  No atomics anywhere: every
// sum has a fixed order, so results are bit-reproducible run to run.
// The reflectors are left in the columns of A in explicit form (1 at the pivot, 0 above), so the
// back-transformation X <- H_0 ... H_{n-2} X is three DMMA GEMMs per block of 256 reflectors
// (compact WY, T factors from one batched Gram GEMM).
#include "tnb_arith.cuh"
#include "tnb_internal.h"

#include <algorithm>
#include <cstdlib>

namespace tnb {

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// Programmatic dependent launch (sm_90+): the K1 -> K2 -> K1 ... chain of one column is latency-bound for n <= 4096,
// so each kernel is launched with programmaticStreamSerialization: its blocks are scheduled while the previous
// kernel drains, run whatever does not depend on it, and block in pdl_wait() until it has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

constexpr int TD_NB = 64;        // panel width of the reduction
constexpr int TD_NBT = 256;      // reflectors per block in the back-transformation

// scal layout (elements of T): [0] = 1/(alpha-beta) (v = x * scal[0], v[pivot] = 1), [1] = alpha slot
//
// K1: 256 threads = 32 rows x 8 k-groups (warp w owns panel columns k = w, w+8, ...), so a thread has at most
// 8 (V,W) pairs to load and all of them are in flight at once; the k-groups are summed through shared memory
// in a fixed order.  The prologue (y^H v from K2's per-CTA partials, p^H q, w at the pivot row) is
// warp-parallel for the same reason: every dependent global-load round trip here is paid once per column.
template <bool CPLX>
__global__ void __launch_bounds__(256) td_k1_kernel(typename ElemT<CPLX>::T* __restrict__ A, long long n, long long i,
                                                     int jp, int do_column, typename ElemT<CPLX>::T* ZL,
                                                     typename ElemT<CPLX>::T* ZR, long long ldz,
                                                     const typename ElemT<CPLX>::T* __restrict__ ybuf,
                                                     const typename ElemT<CPLX>::T* __restrict__ yparts, int nparts,
                                                     const typename ElemT<CPLX>::T* __restrict__ P, int nbt,
                                                     typename ElemT<CPLX>::T* tau, double* d, double* e,
                                                     typename ElemT<CPLX>::T* scal, double* partials, unsigned* counter) {
  using T = typename ElemT<CPLX>::T;
  constexpr int nb = TD_NB;
  __shared__ T p_s[nb], q_s[nb], rowW[nb], rowV[nb];
  __shared__ T redT[8];
  __shared__ T accS[8][32], colS[8][32], yS[8][32];
  constexpr int TS = 128;     // = TD_TS: slot length of the symmetric matvec's partial buffer
  __shared__ double redD[8];
  __shared__ T bc[2];
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();            // the K2 of this column may be scheduled; it waits for this grid to complete
  const long long nrows = n - i;
  T alpha2 = a_zero<T>(), tau_p = a_zero<T>();
  const T* vprev = ZL + (size_t)(jp < 0 ? 0 : jp) * ldz;
  const long long r = i + (long long)blockIdx.x * 32 + lane;
  const bool active = r < n;
  // (a) issue every global load that depends on nothing else first: one round trip instead of five
  T vk[nb / 8], wk[nb / 8];
#pragma unroll
  for (int u = 0; u < nb / 8; ++u) {
    const int k = warp + 8 * u;
    const bool on = active && k < jp;
    vk[u] = on ? ZL[r + (size_t)k * ldz] : a_zero<T>();
    wk[u] = on ? ZL[r + (size_t)(nb + k) * ldz] : a_zero<T>();
  }
  const bool fin = (warp == 0 && active);
  const T a_in = (fin && do_column) ? A[r + (size_t)i * n] : a_zero<T>();
  T zi_v = a_zero<T>(), zi_w = a_zero<T>(), zi_v2 = a_zero<T>(), zi_w2 = a_zero<T>();   // row i of the panel (warp 0)
  if (warp == 0 && jp > 0) {
    if (lane < jp) { zi_v = ZL[i + (size_t)lane * ldz]; zi_w = ZL[i + (size_t)(nb + lane) * ldz]; }
    if (lane + 32 < jp) { zi_v2 = ZL[i + (size_t)(lane + 32) * ldz]; zi_w2 = ZL[i + (size_t)(nb + lane + 32) * ldz]; }
  }
  if (jp >= 0) tau_p = tau[i - 1];
  // everything above was written before the previous column's K2 started; what follows is K2's output
  pdl_wait();
  const T vp = (fin && jp >= 0) ? vprev[r] : a_zero<T>();
  // y of the previous column: either final values (K2) or the per-tile slots of the symmetric matvec (K2S),
  // which are summed here in a fixed order (warp w takes slots w, w+8, ...)
  T yr = (fin && jp >= 0 && nbt == 0) ? ybuf[r - i] : a_zero<T>();
  T ysl = a_zero<T>(), y0 = a_zero<T>();
  if (jp >= 0 && nbt > 0) {
    if (active) {
      const long long t = r - i;
      const T* src = P + ((size_t)(t / TS) * nbt) * TS + (t % TS);
      for (int k = warp; k < nbt; k += 8) ysl = a_add(ysl, src[(size_t)k * TS]);
    }
    if (warp == 0) {                      // y at the pivot row (t = 0), needed by lane 0 below
      for (int k = lane; k < nbt; k += 32) y0 = a_add(y0, P[(size_t)k * TS]);
    }
  }
  if (jp >= 0) {
    // y^H v: fixed-order sum of K2's per-CTA partials
    T part = a_zero<T>();
    for (int b = tid; b < nparts; b += 256) part = a_add(part, yparts[b]);
    if (tid < jp) { q_s[tid] = ybuf[nrows + tid]; p_s[tid] = ybuf[nrows + jp + tid]; }
    part = a_warp_sum(part);
    if (lane == 0) redT[warp] = part;
    __syncthreads();
    if (warp == 0) {
      // p^H q + q^H p and the panel part of w at the pivot row i, two panel columns per lane
      double pq = 0.0;
      T wacc = a_zero<T>();
      if (lane < jp) {
        pq += 2.0 * a_re(a_cmul(p_s[lane], q_s[lane]));
        wacc = a_add(wacc, a_add(a_mul(zi_v, p_s[lane]), a_mul(zi_w, q_s[lane])));
        if (do_column) { rowW[lane] = a_conj(zi_w); rowV[lane] = a_conj(zi_v); }
      }
      if (lane + 32 < jp) {
        pq += 2.0 * a_re(a_cmul(p_s[lane + 32], q_s[lane + 32]));
        wacc = a_add(wacc, a_add(a_mul(zi_v2, p_s[lane + 32]), a_mul(zi_w2, q_s[lane + 32])));
        if (do_column) { rowW[lane + 32] = a_conj(zi_w2); rowV[lane + 32] = a_conj(zi_v2); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pq += __shfl_xor_sync(0xffffffffu, pq, o);
      wacc = a_warp_sum(wacc);
      if (nbt > 0) y0 = a_warp_sum(y0);
      if (lane == 0) {
        T yhv = a_zero<T>();
        for (int w = 0; w < 8; ++w) yhv = a_add(yhv, redT[w]);
        const T wphv = a_cmul(tau_p, a_sub(yhv, a_real<T>(pq)));     // conj(tau) * (y^H v - p^H q - q^H p)
        const T al2 = a_scale(a_mul(tau_p, wphv), -0.5);
        bc[0] = al2;
        const T vpi = vprev[i];
        const T ypiv = (nbt > 0) ? y0 : ybuf[0];
        const T wi = a_add(a_mul(tau_p, a_sub(ypiv, wacc)), a_mul(al2, vpi));
        rowW[jp] = a_conj(wi);
        rowV[jp] = a_conj(vpi);
      }
    }
  }
  __syncthreads();
  if (jp >= 0) alpha2 = bc[0];
  // partial sums of this k-group
  T acc = a_zero<T>(), ca = a_zero<T>();
#pragma unroll
  for (int u = 0; u < nb / 8; ++u) {
    const int k = warp + 8 * u;
    if (k < jp) {
      acc = a_add(acc, a_add(a_mul(vk[u], p_s[k]), a_mul(wk[u], q_s[k])));
      if (do_column) ca = a_add(ca, a_add(a_mul(vk[u], rowW[k]), a_mul(wk[u], rowV[k])));
    }
  }
  accS[warp][lane] = acc;
  colS[warp][lane] = ca;
  yS[warp][lane] = ysl;
  __syncthreads();
  double part2 = 0.0;
  if (fin) {
    T accsum = a_zero<T>(), casum = a_zero<T>();
#pragma unroll
    for (int w = 0; w < 8; ++w) { accsum = a_add(accsum, accS[w][lane]); casum = a_add(casum, colS[w][lane]); }
    if (nbt > 0) {
#pragma unroll
      for (int w = 0; w < 8; ++w) yr = a_add(yr, yS[w][lane]);
    }
    T a = do_column ? a_sub(a_in, casum) : a_zero<T>();
    if (jp >= 0) {
      const T wr = a_add(a_mul(tau_p, a_sub(yr, accsum)), a_mul(alpha2, vp));
      ZL[r + (size_t)(nb + jp) * ldz] = wr;
      ZR[r + (size_t)jp * ldz] = wr;
      A[r + (size_t)(i - 1) * n] = vp;      // reflector i-1 stored in place (explicit form)
      if (do_column) a = a_sub(a, a_add(a_mul(vp, rowW[jp]), a_mul(wr, rowV[jp])));
    }
    if (do_column) {
      if (r == i) {
        d[i] = a_re(a);
        A[r + (size_t)i * n] = a_real<T>(a_re(a));
      } else {
        A[r + (size_t)i * n] = a;
        if (r == i + 1) scal[1] = a;
        else part2 = a_abs2(a);
      }
    }
  }
  if (!do_column || nrows < 2) return;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part2 += __shfl_xor_sync(0xffffffffu, part2, o);
    if (lane == 0) {
      partials[blockIdx.x] = part2;
      __threadfence();
      const unsigned t = atomicInc(counter, gridDim.x - 1);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double xs = 0.0;
  for (unsigned b = tid; b < gridDim.x; b += 256) xs += ((volatile double*)partials)[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) xs += __shfl_xor_sync(0xffffffffu, xs, o);
  if (lane == 0) redD[warp] = xs;
  __syncthreads();
  if (tid != 0) return;
  double xn2 = 0.0;
  for (int w = 0; w < 8; ++w) xn2 += redD[w];
  T alpha;
  if constexpr (CPLX) { alpha.x = ((volatile double*)scal)[2]; alpha.y = ((volatile double*)scal)[3]; }
  else alpha = ((volatile double*)scal)[1];
  const double ar = a_re(alpha), ai = a_im(alpha);
  T t, sc;
  double beta;
  if (xn2 == 0.0 && ai == 0.0) {
    t = a_zero<T>(); sc = a_zero<T>(); beta = ar;
  } else {
    const double nrm = sqrt(ar * ar + ai * ai + xn2);
    beta = ar >= 0 ? -nrm : nrm;
    if constexpr (CPLX) {
      t = make_double2((beta - ar) / beta, -ai / beta);
      const double dr = ar - beta, dd = dr * dr + ai * ai;
      sc = make_double2(dr / dd, -ai / dd);
    } else {
      t = (beta - ar) / beta;
      sc = 1.0 / (ar - beta);
    }
  }
  tau[i] = t;
  e[i] = beta;
  scal[0] = sc;
}

__device__ __forceinline__ double2 ld_stream16(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

constexpr int TD_K2T = 512;     // threads per K2 block (16 warps share one copy of v in shared memory)

// One warp per CW "extended columns": the nt trailing columns of A, then the j panel columns of V, then of W.
// out[c] = sum_t conj(col_c[t]) v[t] over the nt rows below the diagonal of column i.
// HBM-bound: every lane keeps 8 independent 16-byte loads in flight (L1 no-allocate so that v stays cached),
// 2 CTAs x 16 warps per SM.  Also emits this CTA's share of y^H v (fixed summation order).
template <bool CPLX, int CW, bool XS>
__global__ void __launch_bounds__(TD_K2T, 2) td_k2_kernel(const typename ElemT<CPLX>::T* __restrict__ A, long long n, long long i, int j,
                                                          typename ElemT<CPLX>::T* ZL, typename ElemT<CPLX>::T* ZR, long long ldz,
                                                          const typename ElemT<CPLX>::T* __restrict__ scal,
                                                          typename ElemT<CPLX>::T* __restrict__ ybuf,
                                                          typename ElemT<CPLX>::T* __restrict__ yparts) {
  using T = typename ElemT<CPLX>::T;
  constexpr int nb = TD_NB;
  constexpr int NW = TD_K2T / 32;
  constexpr int UNR = 8 / CW;
  extern __shared__ __align__(16) unsigned char k2_smem[];
  __shared__ T wsum[NW];
  T* xs = (T*)k2_smem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();               // K1 of this column must have completed (scal, the updated column, the finalised w)
  pdl_trigger();            // ... after which the next column's K1 may start its independent loads
  const long long r0 = i + 1;
  const int nt = (int)(n - r0);
  const T sc = scal[0];
  const T* x = A + r0 + (size_t)i * n;
  if (XS) {
#pragma unroll 8
    for (int t = tid; t < nt; t += TD_K2T) xs[t] = (t == 0) ? a_one<T>() : a_mul(x[t], sc);
  }
  // every CTA stores its share of the normalised reflector into the panel buffers
  for (long long t = (long long)blockIdx.x * TD_K2T + tid; t < nt; t += (long long)gridDim.x * TD_K2T) {
    const T v = (t == 0) ? a_one<T>() : a_mul(x[t], sc);
    ZL[r0 + t + (size_t)j * ldz] = v;
    ZR[r0 + t + (size_t)(nb + j) * ldz] = v;
  }
  if (XS) __syncthreads();
  const int Ct = nt + 2 * j;
  const int tasks = (Ct + CW - 1) / CW;
  T yv = a_zero<T>();      // this warp's share of y^H v (lane 0 only)
  for (int task = blockIdx.x * NW + warp; task < tasks; task += gridDim.x * NW) {
    const T* cp[CW];
#pragma unroll
    for (int u = 0; u < CW; ++u) {
      int c = task * CW + u;
      if (c >= Ct) c = task * CW;      // duplicate a valid column; result discarded
      if (c < nt) cp[u] = A + r0 + (size_t)(r0 + c) * n;
      else if (c < nt + j) cp[u] = ZL + r0 + (size_t)(c - nt) * ldz;
      else cp[u] = ZL + r0 + (size_t)(nb + c - nt - j) * ldz;
    }
    T acc[CW];
#pragma unroll
    for (int u = 0; u < CW; ++u) acc[u] = a_zero<T>();
    if constexpr (XS && !CPLX) {
      // 16-byte loads: peel one row where the column start is not 16-byte aligned
      int st[CW], nv[CW];
      int nvmax = 0;
#pragma unroll
      for (int u = 0; u < CW; ++u) {
        st[u] = (int)(((uintptr_t)cp[u] >> 3) & 1);
        nv[u] = (nt - st[u]) >> 1;
        nvmax = max(nvmax, nv[u]);
        if (lane == 0) {
          if (st[u]) acc[u] += cp[u][0] * xs[0];
          if ((nt - st[u]) & 1) acc[u] += cp[u][nt - 1] * xs[nt - 1];
        }
      }
      for (int m0 = lane; m0 < nvmax; m0 += 32 * UNR) {
        double2 av[CW][UNR];
#pragma unroll
        for (int u = 0; u < CW; ++u)
#pragma unroll
          for (int w = 0; w < UNR; ++w) {
            const int m = m0 + 32 * w;
            av[u][w] = (m < nv[u]) ? ld_stream16((const double2*)(cp[u] + st[u]) + m) : make_double2(0.0, 0.0);
          }
#pragma unroll
        for (int u = 0; u < CW; ++u)
#pragma unroll
          for (int w = 0; w < UNR; ++w) {
            const int m = m0 + 32 * w;
            const int t = (m < nv[u]) ? st[u] + 2 * m : 0;
            acc[u] += av[u][w].x * xs[t] + av[u][w].y * xs[t + 1];
          }
      }
    } else if constexpr (XS && CPLX) {
      for (int t0 = lane; t0 < nt; t0 += 32 * UNR) {
        double2 av[CW][UNR];
#pragma unroll
        for (int u = 0; u < CW; ++u)
#pragma unroll
          for (int w = 0; w < UNR; ++w) {
            const int t = t0 + 32 * w;
            av[u][w] = (t < nt) ? ld_stream16((const double2*)cp[u] + t) : make_double2(0.0, 0.0);
          }
#pragma unroll
        for (int u = 0; u < CW; ++u)
#pragma unroll
          for (int w = 0; w < UNR; ++w) {
            const int t = t0 + 32 * w;
            if (t < nt) acc[u] = a_add(acc[u], a_cmul(av[u][w], xs[t]));
          }
      }
    } else {
      for (int t = lane; t < nt; t += 32) {
        const T v = (t == 0) ? a_one<T>() : a_mul(__ldg(x + t), sc);
#pragma unroll
        for (int u = 0; u < CW; ++u) acc[u] = a_add(acc[u], a_cmul(cp[u][t], v));
      }
    }
#pragma unroll
    for (int u = 0; u < CW; ++u) {
      const T s = a_warp_sum(acc[u]);
      const int c = task * CW + u;
      if (lane == 0 && c < Ct) {
        ybuf[c] = s;
        if (c < nt) {
          T vc;
          if (XS) vc = xs[c];
          else vc = (c == 0) ? a_one<T>() : a_mul(__ldg(x + c), sc);
          yv = a_add(yv, a_cmul(s, vc));
        }
      }
    }
  }
  if (lane == 0) wsum[warp] = yv;
  __syncthreads();
  if (tid == 0) {
    T t = a_zero<T>();
    for (int w = 0; w < NW; ++w) t = a_add(t, wsum[w]);
    yparts[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------
// Symmetric form of K2 for large trailing matrices: read only the LOWER triangle (half the HBM traffic).
// The trailing matrix is cut into 128 x 128 tiles; tile (I,J), I >= J, is read once by one CTA and serves both
// y_I += A_IJ v_J (row part, accumulated per lane, then across the 8 warps in a fixed order) and
// y_J += A_IJ^H v_I (column dots).  Each (block, tile) pair owns one slot of P, so there are no atomics and the
// result is bit-reproducible; K2R sums the slots of every block in a fixed order.
// ------------------------------------------------------------------------------------
constexpr int TD_TS = 128;

__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) { return ld_stream16(p); }

template <bool CPLX>
__global__ void __launch_bounds__(256, 2) td_k2s_kernel(const typename ElemT<CPLX>::T* __restrict__ A, long long n, long long i, int j,
                                                      typename ElemT<CPLX>::T* ZL, typename ElemT<CPLX>::T* ZR, long long ldz,
                                                      const typename ElemT<CPLX>::T* __restrict__ scal,
                                                      typename ElemT<CPLX>::T* __restrict__ P, int nbt, int ntiles,
                                                      typename ElemT<CPLX>::T* __restrict__ ybuf,
                                                      typename ElemT<CPLX>::T* __restrict__ yparts) {
  using T = typename ElemT<CPLX>::T;
  constexpr int nb = TD_NB, TS = TD_TS;
  constexpr int CB = CPLX ? 2 : 4;      // columns per load batch: 4*CB independent loads in flight per lane
  __shared__ T vI[TS], vJ[TS], colres[TS];
  __shared__ T yvred[4];
  __shared__ T rowacc[8][TS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_trigger();
  const long long r0 = i + 1;
  const int nt = (int)(n - r0);
  const T sc = scal[0];
  const T* x = A + r0 + (size_t)i * n;
  auto vat = [&](int g) -> T { return g >= nt ? a_zero<T>() : (g == 0 ? a_one<T>() : a_mul(__ldg(x + g), sc)); };
  if ((int)blockIdx.x >= ntiles) {
    // panel dots q = V^H v, p = W^H v: one warp per panel column
    const int pc = ((int)blockIdx.x - ntiles) * 8 + warp;
    if (pc >= 2 * j) return;
    const T* col = (pc < j) ? ZL + r0 + (size_t)pc * ldz : ZL + r0 + (size_t)(nb + pc - j) * ldz;
    T acc = a_zero<T>();
    int t = lane;
    for (; t + 96 < nt; t += 128) {
      T cv[4], vv[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) { cv[w] = col[t + 32 * w]; vv[w] = __ldg(x + t + 32 * w); }
#pragma unroll
      for (int w = 0; w < 4; ++w) acc = a_add(acc, a_cmul(cv[w], (t + 32 * w == 0) ? a_one<T>() : a_mul(vv[w], sc)));
    }
    for (; t < nt; t += 32) acc = a_add(acc, a_cmul(col[t], vat(t)));
    acc = a_warp_sum(acc);
    if (lane == 0) ybuf[nt + pc] = acc;
    return;
  }
  int I = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((I + 1) * (I + 2) / 2 <= (int)blockIdx.x) ++I;
  while (I * (I + 1) / 2 > (int)blockIdx.x) --I;
  const int J = (int)blockIdx.x - I * (I + 1) / 2;
  const bool diag = (I == J);
  T vr[4], yr[4];
  bool rok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { yr[q] = a_zero<T>(); rok[q] = (I * TS + lane + 32 * q) < nt; }
  const T* base = A + (r0 + (size_t)I * TS + lane) + (size_t)(r0 + (size_t)J * TS) * n;
  // two-stage software pipeline: the loads of column batch b+1 are in flight while batch b is reduced
  constexpr int NBATCH = (TS / 8) / CB;
  T av[2][CB][4];
  auto load_batch = [&](int cb, T (*dst)[4]) {
#pragma unroll
    for (int u = 0; u < CB; ++u) {
      const int c = warp + 8 * (cb + u);
      const bool cok = (J * TS + c) < nt;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bool ok = cok && rok[q];
        if (diag) ok = ok && (lane + 32 * q >= c);
        dst[u][q] = ok ? ld_stream(base + 32 * q + (size_t)c * n) : a_zero<T>();
      }
    }
  };
  load_batch(0, av[0]);      // the matrix stream starts before the (dependent-latency) fetch of the two v segments
  if (tid < TS) {
    const T v = vat(I * TS + tid);
    vI[tid] = v;
    if (diag && I * TS + tid < nt) {
      ZL[r0 + I * TS + tid + (size_t)j * ldz] = v;
      ZR[r0 + I * TS + tid + (size_t)(nb + j) * ldz] = v;
    }
  } else {
    vJ[tid - TS] = vat(J * TS + tid - TS);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) vr[q] = vI[lane + 32 * q];
#pragma unroll
  for (int bt = 0; bt < NBATCH; ++bt) {
    if (bt + 1 < NBATCH) load_batch((bt + 1) * CB, av[(bt + 1) & 1]);
#pragma unroll
    for (int u = 0; u < CB; ++u) {
      const int c = warp + 8 * (bt * CB + u);
      const T vc = vJ[c];
      T s = a_zero<T>();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        s = a_add(s, a_cmul(av[bt & 1][u][q], vr[q]));
        if (!diag || (lane + 32 * q > c)) yr[q] = a_add(yr[q], a_mul(av[bt & 1][u][q], vc));
      }
      s = a_warp_sum(s);
      if (lane == 0) colres[c] = s;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) rowacc[warp][lane + 32 * q] = yr[q];
  __syncthreads();
  if (tid < TS) {
    T rs = a_zero<T>();
#pragma unroll
    for (int w = 0; w < 8; ++w) rs = a_add(rs, rowacc[w][tid]);
    T yv;                                 // this tile's share of y^H v
    if (diag) {
      const T y = a_add(rs, colres[tid]);
      P[((size_t)I * nbt + I) * TS + tid] = y;
      yv = a_cmul(y, vI[tid]);
    } else {
      P[((size_t)I * nbt + J) * TS + tid] = rs;
      P[((size_t)J * nbt + I) * TS + tid] = colres[tid];
      yv = a_add(a_cmul(rs, vI[tid]), a_cmul(colres[tid], vJ[tid]));
    }
    yv = a_warp_sum(yv);
    if (lane == 0) yvred[warp] = yv;
  }
  __syncthreads();
  if (tid == 0) yparts[blockIdx.x] = a_add(a_add(yvred[0], yvred[1]), a_add(yvred[2], yvred[3]));
}

// lower triangle -> upper triangle (conjugated) of the trailing block A[lo:, lo:], used once when the
// reduction switches from the lower-only (symmetric) kernels to the full-column ones
template <typename T>
__global__ void td_mirror_kernel(T* A, long long n, long long lo) {
  const long long m = n - lo;
  for (long long eidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; eidx < m * m; eidx += (long long)gridDim.x * blockDim.x) {
    const long long r = eidx % m, c = eidx / m;
    if (r > c) A[(lo + c) + (size_t)(lo + r) * n] = a_conj(A[(lo + r) + (size_t)(lo + c) * n]);
  }
}

// zero the part of A above the reflectors: column c keeps rows > c (pivot c+1 holds the explicit 1);
// the last column holds no reflector.
template <typename T>
__global__ void td_clean_kernel(T* A, long long n) {
  for (long long eidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; eidx < n * n; eidx += (long long)gridDim.x * blockDim.x) {
    const long long r = eidx % n, c = eidx / n;
    if (r <= c || c == n - 1) A[eidx] = a_zero<T>();
  }
}

size_t tridiag_ws_bytes(int dtype, int64_t n) {
  const size_t es = elsize(dtype);
  const size_t nbt = (size_t)(n + TD_TS - 1) / TD_TS;
  return 2 * al256((size_t)n * 2 * TD_NB * es) + al256((size_t)(n + 2 * TD_NB) * es) + al256(64) + al256(1024 * es) +
         al256(nbt * nbt * TD_TS * es) + al256((n / 256 + 2) * es) + 4096;
}

template <bool CPLX, int CW>
static int launch_k2(Handle* h, const typename ElemT<CPLX>::T* A, int64_t n, int64_t i, int j, typename ElemT<CPLX>::T* ZL,
                     typename ElemT<CPLX>::T* ZR, int64_t ldz, const typename ElemT<CPLX>::T* scal,
                     typename ElemT<CPLX>::T* ybuf, typename ElemT<CPLX>::T* yparts, cudaStream_t st) {
  using T = typename ElemT<CPLX>::T;
  const int nt = (int)(n - i - 1);
  const int Ct = nt + 2 * j;
  const int tasks = (Ct + CW - 1) / CW;
  constexpr int NW = TD_K2T / 32;
  int grid = std::min((tasks + NW - 1) / NW, h->num_sms * 2);
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)(nt + 2) * sizeof(T);
  if (smem <= 96 * 1024)
    launch_pdl(td_k2_kernel<CPLX, CW, true>, dim3(grid), dim3(TD_K2T), smem, st, true, A, (long long)n, (long long)i, j, ZL, ZR, (long long)ldz, scal, ybuf, yparts);
  else
    launch_pdl(td_k2_kernel<CPLX, CW, false>, dim3(grid), dim3(TD_K2T), 0, st, true, A, (long long)n, (long long)i, j, ZL, ZR, (long long)ldz, scal, ybuf, yparts);
  h->launches++;
  return grid;
}

// A (n x n, full Hermitian storage, ld n) -> d (n), e (n-1), tau (n); reflectors left in A (explicit form).
template <bool CPLX>
static int tridiag_core(Handle* h, int64_t n, void* Av, double* d, double* e, void* tauv, cudaStream_t st) {
  using T = typename ElemT<CPLX>::T;
  const int dtype = CPLX ? TNB_C128 : TNB_F64;
  constexpr int nb = TD_NB;
  T* A = (T*)Av;
  T* tau = (T*)tauv;
  void *ZLv, *ZRv, *yv, *scv, *ypv;
  TNB_TRY(ws_alloc(h, (size_t)n * 2 * nb * sizeof(T), &ZLv));
  TNB_TRY(ws_alloc(h, (size_t)n * 2 * nb * sizeof(T), &ZRv));
  TNB_TRY(ws_alloc(h, (size_t)(n + 2 * nb) * sizeof(T), &yv));
  TNB_TRY(ws_alloc(h, 64, &scv));
  const size_t nbt_max = (size_t)(n + TD_TS - 1) / TD_TS;
  TNB_TRY(ws_alloc(h, std::max<size_t>((size_t)h->num_sms * 4, nbt_max * (nbt_max + 1) / 2 + 8) * sizeof(T), &ypv));
  T *ZL = (T*)ZLv, *ZR = (T*)ZRv, *ybuf = (T*)yv, *scal = (T*)scv, *yparts = (T*)ypv;
  int nparts = 0;
  // symmetric (half-traffic) matvec for trailing sizes >= sym_min; TNB_TD_SYM_MIN overrides, 0 disables
  const char* sm_env = getenv("TNB_TD_SYM_MIN");
  const int64_t sym_min = sm_env ? atoll(sm_env) : 1536;
  void* Pv = nullptr;
  if (sym_min > 0 && n - 1 >= sym_min) {
    const size_t nbt = (size_t)(n + TD_TS - 1) / TD_TS;
    TNB_TRY(ws_alloc(h, nbt * nbt * TD_TS * sizeof(T), &Pv));
  }
  T* P = (T*)Pv;
  int nbt_prev = 0;            // > 0: the previous column's y lives in P slots (K2S), else in ybuf (K2)
  bool lower_valid_only = false;
  TNB_ONCE_PER_DEVICE(h, TNB_CUDA(h, cudaFuncSetAttribute(td_k2_kernel<CPLX, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                      TNB_CUDA(h, cudaFuncSetAttribute(td_k2_kernel<CPLX, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)));
  TNB_CUDA(h, cudaMemsetAsync(tau, 0, (size_t)n * sizeof(T), st));
  const double one[2] = {1.0, 0.0}, mone[2] = {-1.0, 0.0};
  auto k1 = [&](int64_t i, int jp, int do_column) {
    const int grid = (int)((n - i + 31) / 32);
    launch_pdl(td_k1_kernel<CPLX>, dim3(grid), dim3(256), 0, st, true, A, (long long)n, (long long)i, jp, do_column, ZL, ZR, (long long)n,
               (const T*)ybuf, (const T*)yparts, nparts, (const T*)P, nbt_prev, tau, d, e, scal, h->partials, h->counter);
    h->launches++;
  };
  for (int64_t p0 = 0; p0 < n - 1; p0 += nb) {
    const int jb = (int)std::min<int64_t>(nb, n - 1 - p0);
    if (jb < nb) {
      TNB_CUDA(h, cudaMemsetAsync(ZL, 0, (size_t)n * 2 * nb * sizeof(T), st));
      TNB_CUDA(h, cudaMemsetAsync(ZR, 0, (size_t)n * 2 * nb * sizeof(T), st));
    }
    // a panel takes the symmetric (lower-triangle-only) kernels iff every one of its columns is large enough
    const bool panel_sym = P && (n - (p0 + jb)) >= sym_min;
    if (!panel_sym && lower_valid_only) {      // switching to full-column kernels: rebuild the upper triangle once
      td_mirror_kernel<T><<<h->num_sms * 4, 256, 0, st>>>(A, n, p0);
      h->launches++;
      lower_valid_only = false;
    }
    for (int j = 0; j < jb; ++j) {
      const int64_t i = p0 + j;
      k1(i, j - 1, 1);
      const int Ct = (int)(n - i - 1) + 2 * j;
      const int nt = (int)(n - i - 1);
      if (panel_sym) {
        const int nbt = (nt + TD_TS - 1) / TD_TS;
        const int ntiles = nbt * (nbt + 1) / 2;
        launch_pdl(td_k2s_kernel<CPLX>, dim3(ntiles + (2 * j + 7) / 8), dim3(256), 0, st, true, (const T*)A, (long long)n, (long long)i, j, ZL, ZR,
                   (long long)n, (const T*)scal, P, nbt, ntiles, ybuf, yparts);
        nparts = ntiles;
        nbt_prev = nbt;
        h->launches++;
        continue;
      }
      nbt_prev = 0;
      if (Ct >= 2 * 32 * h->num_sms) nparts = launch_k2<CPLX, 2>(h, A, n, i, j, ZL, ZR, n, scal, ybuf, yparts, st);
      else nparts = launch_k2<CPLX, 1>(h, A, n, i, j, ZL, ZR, n, scal, ybuf, yparts, st);
    }
    const int64_t lo = p0 + jb;
    k1(lo, jb - 1, 0);
    const int64_t ntr = n - lo;
    // after a symmetric panel only the lower triangle is kept up to date (syr2k-style: half the update)
    TNB_TRY(gemm_impl(h, dtype, 'N', 'C', ntr, ntr, 2 * nb, mone, ZL + lo, n, ZR + lo, n, one, A + lo + (size_t)lo * n, n, st,
                      panel_sym ? 1 : 0));
    if (panel_sym) lower_valid_only = true;
  }
  nbt_prev = 0;
  k1(n - 1, -1, 1);
  td_clean_kernel<T><<<h->num_sms * 4, 256, 0, st>>>(A, n);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "tridiag");
}

int tridiag_impl(Handle* h, int dtype, int64_t n, void* A, double* d, double* e, void* tau, cudaStream_t st) {
  if (dtype == TNB_F64) return tridiag_core<false>(h, n, A, d, e, tau, st);
  if (dtype == TNB_C128) return tridiag_core<true>(h, n, A, d, e, tau, st);
  return set_err(h, TNB_ERR_UNSUPPORTED, "tridiag: dtype %d", dtype);
}

// ------------------------------------------------------------------------------------
// back-transformation  X <- H_0 H_1 ... H_{n-2} X   (X: n x kx, ld ldx)
// ------------------------------------------------------------------------------------
// T factor of one block of cb reflectors from G = V^H V and tau (larft, forward / columnwise)
template <typename T>
__global__ void __launch_bounds__(256) td_larft_kernel(const T* __restrict__ Gall, const T* __restrict__ tau_all, T* Tall,
                                                        int nbt, int cb_last, int nblk) {
  const int b = blockIdx.x;
  const int cb = (b == nblk - 1) ? cb_last : nbt;
  const T* G = Gall + (size_t)b * nbt * nbt;
  const T* tau = tau_all + (size_t)b * nbt;
  T* Tm = Tall + (size_t)b * nbt * nbt;
  const int tid = threadIdx.x;
  for (int c = 0; c < cb; ++c) {
    const T tc = tau[c];
    for (int t = tid; t <= c; t += blockDim.x) {
      if (t == c) Tm[c + (size_t)c * nbt] = tc;
      else {
        T acc = a_zero<T>();
        for (int l = t; l < c; ++l) acc = a_add(acc, a_mul(Tm[t + (size_t)l * nbt], G[l + (size_t)c * nbt]));
        Tm[t + (size_t)c * nbt] = a_neg(a_mul(tc, acc));
      }
    }
    __syncthreads();
  }
  for (int eidx = tid; eidx < nbt * nbt; eidx += blockDim.x) {
    const int r = eidx % nbt, c = eidx / nbt;
    if (r > c || r >= cb || c >= cb) Tm[eidx] = a_zero<T>();
  }
}

size_t backtransform_ws_bytes(int dtype, int64_t n, int64_t kx) {
  const size_t es = elsize(dtype);
  const int64_t nblk = std::max<int64_t>(1, (n - 1 + TD_NBT - 1) / TD_NBT);
  return 2 * al256((size_t)nblk * TD_NBT * TD_NBT * es) + 2 * al256((size_t)TD_NBT * kx * es) + 4096;
}

int backtransform_impl(Handle* h, int dtype, int64_t n, const void* Vst, const void* tau, void* X, int64_t ldx, int64_t kx,
                       cudaStream_t st) {
  const size_t es = elsize(dtype);
  const int64_t nref = n - 1;
  if (nref < 1 || kx < 1) return TNB_OK;
  const int nbt = TD_NBT;
  const int nblk = (int)((nref + nbt - 1) / nbt);
  const int cb_last = (int)(nref - (int64_t)(nblk - 1) * nbt);
  void *G, *Tm, *W1, *W2;
  TNB_TRY(ws_alloc(h, (size_t)nblk * nbt * nbt * es, &G));
  TNB_TRY(ws_alloc(h, (size_t)nblk * nbt * nbt * es, &Tm));
  TNB_TRY(ws_alloc(h, (size_t)nbt * kx * es, &W1));
  TNB_TRY(ws_alloc(h, (size_t)nbt * kx * es, &W2));
  const char* V = (const char*)Vst;
  // Gram of every block (rows 0..n-1: the explicit zeros above the pivots make the full range exact)
  if (nblk > 1)
    TNB_TRY(gemm_batched_impl(h, dtype, 'C', 'N', nbt, nbt, n, nullptr, V, n, nullptr, (long long)nbt * n, V, n, nullptr,
                              (long long)nbt * n, nullptr, G, nbt, nullptr, (long long)nbt * nbt, nblk - 1, st));
  {
    const int64_t c0 = (int64_t)(nblk - 1) * nbt;
    TNB_TRY(gemm_impl(h, dtype, 'C', 'N', cb_last, cb_last, n, nullptr, V + (size_t)c0 * n * es, n, V + (size_t)c0 * n * es, n,
                      nullptr, (char*)G + (size_t)(nblk - 1) * nbt * nbt * es, nbt, st));
  }
  if (dtype == TNB_C128) td_larft_kernel<double2><<<nblk, 256, 0, st>>>((const double2*)G, (const double2*)tau, (double2*)Tm, nbt, cb_last, nblk);
  else td_larft_kernel<double><<<nblk, 256, 0, st>>>((const double*)G, (const double*)tau, (double*)Tm, nbt, cb_last, nblk);
  h->launches++;
  const double one[2] = {1.0, 0.0}, mone[2] = {-1.0, 0.0};
  for (int b = nblk - 1; b >= 0; --b) {
    const int64_t c0 = (int64_t)b * nbt;
    const int cb = (b == nblk - 1) ? cb_last : nbt;
    const int64_t r0 = c0 + 1, rows = n - r0;
    const char* Vb = V + ((size_t)r0 + (size_t)c0 * n) * es;
    char* Xb = (char*)X + (size_t)r0 * es;
    const char* Tb = (const char*)Tm + (size_t)b * nbt * nbt * es;
    TNB_TRY(gemm_impl(h, dtype, 'C', 'N', cb, kx, rows, nullptr, Vb, n, Xb, ldx, nullptr, W1, cb, st));
    TNB_TRY(gemm_impl(h, dtype, 'N', 'N', cb, kx, cb, nullptr, Tb, nbt, W1, cb, nullptr, W2, cb, st));
    TNB_TRY(gemm_impl(h, dtype, 'N', 'N', rows, kx, cb, mone, Vb, n, W2, cb, one, Xb, ldx, st));
  }
  return check_cuda(h, cudaGetLastError(), "backtransform");
}

}  // namespace tnb
