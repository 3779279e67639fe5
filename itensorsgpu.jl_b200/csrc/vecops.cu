// Bandwidth-bound primitives: fused permute+axpby, scale, dot, nrm2, spectrum truncation.
//
// Replaces (file:line in /root/reference/src/tensor/): permute!/permutedims!! ->
// CUTENSOR.permutation! (cudense.jl:447-500), +/- -> CUDA.zeros + elementwiseBinary! +
// copyto! (cudense.jl:333-445), scalar */ (cudense.jl:22,502), norm (cudense.jl:27),
// truncate! (cutruncate.jl:1-93).  All are HBM-bound: coalesced, 16-byte vectorised,
// grids sized in multiples of the SM count, results of reductions stay on the device.
#include "tnb_internal.h"

#include <algorithm>
#include <array>
#include <cstring>

namespace tnb {

// ------------------------------------------------------------------------------------
// permute + axpby
// ------------------------------------------------------------------------------------
constexpr int MAXP = 24;
struct PermParams {
  int n;                 // modes after merging, in B order (mode 0 has unit stride in B)
  int ext[MAXP];
  long long sa[MAXP];    // stride in A
  long long sb[MAXP];    // stride in B
  long long total;
  int ja;                // index of the mode with unit stride in A (transpose kernel)
  long long tiles0, tilesA;
  double ar, ai, br, bi;
};

template <bool CPLX>
struct El { using T = double; };
template <>
struct El<true> { using T = double2; };

__device__ __forceinline__ double axpby1(double a, double x, double b, double y, bool hb) {
  return hb ? a * x + b * y : a * x;
}

template <bool CPLX>
__device__ __forceinline__ typename El<CPLX>::T combine(const PermParams& p, typename El<CPLX>::T x,
                                                       typename El<CPLX>::T* yp, bool hb);
template <>
__device__ __forceinline__ double combine<false>(const PermParams& p, double x, double* yp, bool hb) {
  return hb ? p.ar * x + p.br * (*yp) : p.ar * x;
}
template <>
__device__ __forceinline__ double2 combine<true>(const PermParams& p, double2 x, double2* yp, bool hb) {
  double2 v;
  v.x = p.ar * x.x - p.ai * x.y;
  v.y = p.ar * x.y + p.ai * x.x;
  if (hb) {
    double2 y = *yp;
    v.x += p.br * y.x - p.bi * y.y;
    v.y += p.br * y.y + p.bi * y.x;
  }
  return v;
}

// Same fastest mode in A and B: walk B linearly, gather A rows.  Each thread takes UNR
// consecutive elements of mode 0 so the mixed-radix decode is amortised.
template <bool CPLX>
__global__ void __launch_bounds__(256) permute_rows_kernel(const __grid_constant__ PermParams p,
                                                           const typename El<CPLX>::T* __restrict__ A,
                                                           typename El<CPLX>::T* B) {
  using T = typename El<CPLX>::T;
  constexpr int UNR = 4;
  const bool hb = (p.br != 0.0) || (p.bi != 0.0);
  const int e0 = p.ext[0];
  const long long chunks0 = (e0 + UNR - 1) / UNR;
  const long long nchunks = chunks0 * (p.total / e0);
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks;
       c += (long long)gridDim.x * blockDim.x) {
    long long rest = c / chunks0;
    const int i0 = (int)(c - rest * chunks0) * UNR;
    long long oa = (long long)i0 * p.sa[0], ob = i0;
    for (int m = 1; m < p.n; ++m) {
      const long long q = rest / p.ext[m];
      const int r = (int)(rest - q * p.ext[m]);
      oa += r * p.sa[m];
      ob += r * p.sb[m];
      rest = q;
    }
    T x[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (i0 + u < e0) x[u] = A[oa + u * p.sa[0]];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (i0 + u < e0) B[ob + u] = combine<CPLX>(p, x[u], &B[ob + u], hb);
  }
}

// Different fastest modes: 32x32 shared-memory tile over (B's fastest mode 0, A's fastest
// mode ja); reads coalesced along ja, writes coalesced along 0.
template <bool CPLX>
__global__ void __launch_bounds__(256) permute_tile_kernel(const __grid_constant__ PermParams p,
                                                           const typename El<CPLX>::T* __restrict__ A,
                                                           typename El<CPLX>::T* B) {
  using T = typename El<CPLX>::T;
  __shared__ T tile[32][33];
  const bool hb = (p.br != 0.0) || (p.bi != 0.0);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const long long ntile = p.tiles0 * p.tilesA * (p.total / p.ext[0] / p.ext[p.ja]);
  for (long long t = blockIdx.x; t < ntile; t += gridDim.x) {
    long long rest = t;
    const int t0 = (int)(rest % p.tiles0); rest /= p.tiles0;
    const int ta = (int)(rest % p.tilesA); rest /= p.tilesA;
    long long oa = 0, ob = 0;
    for (int m = 1; m < p.n; ++m) {
      if (m == p.ja) continue;
      const long long q = rest / p.ext[m];
      const int r = (int)(rest - q * p.ext[m]);
      oa += r * p.sa[m];
      ob += r * p.sb[m];
      rest = q;
    }
    const int i0 = t0 * 32, j0 = ta * 32;  // i along mode 0, j along mode ja
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < 32; yy += 8) {
      const int i = i0 + ty + yy, j = j0 + tx;
      if (i < p.ext[0] && j < p.ext[p.ja]) tile[ty + yy][tx] = A[oa + (long long)i * p.sa[0] + j];
    }
    __syncthreads();
#pragma unroll
    for (int yy = 0; yy < 32; yy += 8) {
      const int j = j0 + ty + yy, i = i0 + tx;
      if (i < p.ext[0] && j < p.ext[p.ja]) {
        T* dst = &B[ob + i + (long long)j * p.sb[p.ja]];
        *dst = combine<CPLX>(p, tile[tx][ty + yy], dst, hb);
      }
    }
  }
}

// identical layout: y <- a*x + b*y, 2 doubles per thread-iteration (16-byte accesses)
template <bool CPLX>
__global__ void __launch_bounds__(256) axpby_kernel(const __grid_constant__ PermParams p,
                                                    const double2* A, double2* B,
                                                    long long n2, const double* At, double* Bt, int tail) {
  const bool hb = (p.br != 0.0) || (p.bi != 0.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    double2 x = A[i];
    if (CPLX) {
      B[i] = combine<true>(p, x, &B[i], hb);
    } else {
      double2 v;
      if (hb) { double2 y = B[i]; v.x = p.ar * x.x + p.br * y.x; v.y = p.ar * x.y + p.br * y.y; }
      else { v.x = p.ar * x.x; v.y = p.ar * x.y; }
      B[i] = v;
    }
  }
  if (!CPLX && tail && blockIdx.x == 0 && threadIdx.x == 0) {
    double v = p.ar * At[0];
    if (hb) v += p.br * Bt[0];
    Bt[0] = v;
  }
}

static int grid_for(Handle* h, long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)h->num_sms * 8;
  return (int)std::max<long long>(1, std::min(b, cap));
}

int permute_axpby_impl(Handle* h, int dtype, int n, const int64_t* extA, const int32_t* modeA,
                       const void* A, const int32_t* modeB, void* B, const void* alpha,
                       const void* beta, cudaStream_t st) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return set_err(h, TNB_ERR_UNSUPPORTED, "permute: dtype %d", dtype);
  if (n < 0 || n > 64) return set_err(h, TNB_ERR_BAD_ARG, "permute: bad rank %d", n);
  if (!A || !B) return set_err(h, TNB_ERR_BAD_ARG, "permute: null pointer");
  const bool cplx = dtype == TNB_C128;
  // strides of A; order modes as in B
  std::vector<long long> sa(n), ext(n);
  long long s = 1;
  for (int i = 0; i < n; ++i) { sa[i] = s; s *= extA[i]; if (extA[i] < 1) return set_err(h, TNB_ERR_BAD_ARG, "permute: extent < 1"); }
  const long long total = s;
  PermParams p;
  memset(&p, 0, sizeof(p));
  std::vector<std::array<long long, 3>> v;  // ext, sa, sb
  long long sb = 1;
  for (int j = 0; j < n; ++j) {
    int ia = -1;
    for (int i = 0; i < n; ++i) if (modeA[i] == modeB[j]) { if (ia >= 0) return set_err(h, TNB_ERR_BAD_ARG, "permute: repeated mode"); ia = i; }
    if (ia < 0) return set_err(h, TNB_ERR_BAD_ARG, "permute: mode %d of B not in A", modeB[j]);
    for (int j2 = 0; j2 < j; ++j2) if (modeB[j2] == modeB[j]) return set_err(h, TNB_ERR_BAD_ARG, "permute: repeated mode in B");
    const long long e = extA[ia];
    if (e > 1) {
      if (!v.empty() && v.back()[1] * v.back()[0] == sa[ia] && v.back()[2] * v.back()[0] == sb &&
          v.back()[0] * e < 2147483647LL) v.back()[0] *= e;
      else v.push_back({e, sa[ia], sb});
    }
    sb *= e;
  }
  if (v.empty()) v.push_back({1, 1, 1});
  if ((int)v.size() > MAXP) return set_err(h, TNB_ERR_UNSUPPORTED, "permute: more than %d unmergeable modes", MAXP);
  p.n = (int)v.size();
  for (int i = 0; i < p.n; ++i) {
    if (v[i][0] > 2147483647LL) return set_err(h, TNB_ERR_UNSUPPORTED, "permute: extent too large");
    p.ext[i] = (int)v[i][0]; p.sa[i] = v[i][1]; p.sb[i] = v[i][2];
  }
  p.total = total;
  p.ar = 1; p.ai = 0; p.br = 0; p.bi = 0;
  if (alpha) { p.ar = ((const double*)alpha)[0]; if (cplx) p.ai = ((const double*)alpha)[1]; }
  if (beta) { p.br = ((const double*)beta)[0]; if (cplx) p.bi = ((const double*)beta)[1]; }
  if (total == 0) return TNB_OK;

  if (p.n == 1 && p.sa[0] == 1 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0)) {
    const long long nd = cplx ? total : total / 2;
    const int tail = cplx ? 0 : (int)(total & 1);
    const int grid = grid_for(h, nd, 256 * 4);
    if (cplx) axpby_kernel<true><<<grid, 256, 0, st>>>(p, (const double2*)A, (double2*)B, nd, nullptr, nullptr, 0);
    else axpby_kernel<false><<<grid, 256, 0, st>>>(p, (const double2*)A, (double2*)B, nd, (const double*)A + 2 * nd, (double*)B + 2 * nd, tail);
  } else if (p.sa[0] == 1 || p.n == 1) {
    const long long chunks = ((p.ext[0] + 3) / 4) * (total / p.ext[0]);
    const int grid = grid_for(h, chunks, 256);
    if (cplx) permute_rows_kernel<true><<<grid, 256, 0, st>>>(p, (const double2*)A, (double2*)B);
    else permute_rows_kernel<false><<<grid, 256, 0, st>>>(p, (const double*)A, (double*)B);
  } else {
    p.ja = -1;
    for (int i = 1; i < p.n; ++i) if (p.sa[i] == 1) p.ja = i;
    if (p.ja < 0) return set_err(h, TNB_ERR_BAD_ARG, "permute: internal (no unit-stride mode in A)");
    p.tiles0 = (p.ext[0] + 31) / 32;
    p.tilesA = (p.ext[p.ja] + 31) / 32;
    const long long ntile = p.tiles0 * p.tilesA * (total / p.ext[0] / p.ext[p.ja]);
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntile, (long long)h->num_sms * 16));
    if (cplx) permute_tile_kernel<true><<<grid, 256, 0, st>>>(p, (const double2*)A, (double2*)B);
    else permute_tile_kernel<false><<<grid, 256, 0, st>>>(p, (const double*)A, (double*)B);
  }
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "permute_axpby launch");
}

// ------------------------------------------------------------------------------------
// Diag x Dense: out[e] = A[e] * d[(e / inner) % ext]  -- the contraction of a dense tensor with a Diag tensor over
// one of the Diag's two indices is a scale along that mode (plus a relabel, and a permutation when the
// NDTensors output order moves the mode).  The reference densifies the diagonal and runs a full contraction
// (src/tensor/cudiag.jl:147-161): an n x n zero-fill + scatter + GEMM for what is one streaming pass.
// ------------------------------------------------------------------------------------
template <typename T, typename DT>
__global__ void __launch_bounds__(256) scale_mode_kernel(T* __restrict__ out, const T* __restrict__ in, long long total,
                                                          long long inner, long long ext, const DT* __restrict__ d) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const DT f = d[(e / inner) % ext];
    const T v = in[e];
    if constexpr (sizeof(T) == 16 && sizeof(DT) == 16) out[e] = make_double2(v.x * f.x - v.y * f.y, v.x * f.y + v.y * f.x);
    else if constexpr (sizeof(T) == 16) out[e] = make_double2(v.x * f, v.y * f);
    else out[e] = v * f;
  }
}

int diag_contract_impl(Handle* h, int dtype, int n, const int64_t* extA, const int32_t* modeA, const void* A,
                       int32_t scaled_mode, const void* diag, int diag_dtype, const int32_t* modeC, void* C,
                       cudaStream_t st) {
  if (dtype != TNB_F64 && dtype != TNB_C128) return set_err(h, TNB_ERR_UNSUPPORTED, "diag_contract: dtype %d", dtype);
  if (diag_dtype == TNB_C128 && dtype != TNB_C128) return set_err(h, TNB_ERR_UNSUPPORTED, "diag_contract: complex diagonal with a real tensor");
  if (n < 1 || n > 64 || !A || !C || !diag) return set_err(h, TNB_ERR_BAD_ARG, "diag_contract: bad argument");
  long long inner = 1, ext = -1, total = 1;
  bool same = true;
  for (int i = 0; i < n; ++i) {
    if (extA[i] < 1) return set_err(h, TNB_ERR_BAD_ARG, "diag_contract: extent < 1");
    if (modeA[i] == scaled_mode) { ext = extA[i]; inner = total; }
    total *= extA[i];
    if (modeC[i] != modeA[i]) same = false;
  }
  if (ext < 0) return set_err(h, TNB_ERR_BAD_ARG, "diag_contract: mode %d is not a mode of A", scaled_mode);
  const bool cplx = dtype == TNB_C128;
  void* dst = C;
  if (!same) {
    ws_reset(h);
    TNB_TRY(ws_alloc(h, (size_t)total * elsize(dtype), &dst));
  }
  const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16));
  if (!cplx) scale_mode_kernel<double, double><<<grid, 256, 0, st>>>((double*)dst, (const double*)A, total, inner, ext, (const double*)diag);
  else if (diag_dtype == TNB_C128) scale_mode_kernel<double2, double2><<<grid, 256, 0, st>>>((double2*)dst, (const double2*)A, total, inner, ext, (const double2*)diag);
  else scale_mode_kernel<double2, double><<<grid, 256, 0, st>>>((double2*)dst, (const double2*)A, total, inner, ext, (const double*)diag);
  h->launches++;
  TNB_TRY(check_cuda(h, cudaGetLastError(), "scale_mode launch"));
  if (!same) TNB_TRY(permute_axpby_impl(h, dtype, n, extA, modeA, dst, modeC, C, nullptr, nullptr, st));
  return TNB_OK;
}

int scale_impl(Handle* h, int dtype, int64_t n, void* x, const void* alpha, cudaStream_t st) {
  if (!alpha) return TNB_OK;
  int32_t mode = 0;
  int64_t ext = n;
  double zero[2] = {0.0, 0.0};
  if (n <= 0) return n == 0 ? TNB_OK : set_err(h, TNB_ERR_BAD_ARG, "scale: n < 0");
  return permute_axpby_impl(h, dtype, 1, &ext, &mode, x, &mode, x, alpha, zero, st);
}

// ------------------------------------------------------------------------------------
// reductions: dot (conj(x).y) and nrm2.  Two-level, deterministic: fixed grid, fixed
// per-block tree, the last block to finish reduces the per-block partials in index order.
// ------------------------------------------------------------------------------------
template <int MODE>  // 0: real dot, 1: complex dot (conj x), 2: sum of squares (over doubles)
__global__ void __launch_bounds__(256) reduce_kernel(const double2* __restrict__ x, const double2* __restrict__ y,
                                                     long long n2, const double* xt, const double* yt, int tail,
                                                     double* partials, unsigned* counter, double* out) {
  double sr = 0.0, si = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x) {
    const double2 a = x[i];
    if (MODE == 2) {
      sr += a.x * a.x + a.y * a.y;
    } else {
      const double2 b = y[i];
      if (MODE == 0) sr += a.x * b.x + a.y * b.y;
      else { sr += a.x * b.x + a.y * b.y; si += a.x * b.y - a.y * b.x; }
    }
  }
  if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
    if (MODE == 2) sr += xt[0] * xt[0]; else sr += xt[0] * yt[0];
  }
  __shared__ double shr[8], shi[8];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    if (MODE == 1) si += __shfl_xor_sync(0xffffffffu, si, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { shr[w] = sr; shi[w] = si; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0, ti = 0;
    for (int i = 0; i < 8; ++i) { tr += shr[i]; ti += shi[i]; }
    partials[2 * blockIdx.x] = tr;
    partials[2 * blockIdx.x + 1] = ti;
    __threadfence();
    const unsigned t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double tr = 0, ti = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) { tr += partials[2 * i]; ti += partials[2 * i + 1]; }
    // fixed-order tree over 256 threads
    __shared__ double br[256], bi[256];
    br[threadIdx.x] = tr; bi[threadIdx.x] = ti;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
      if (threadIdx.x < sft) { br[threadIdx.x] += br[threadIdx.x + sft]; bi[threadIdx.x] += bi[threadIdx.x + sft]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (MODE == 2) out[0] = sqrt(br[0]);
      else { out[0] = br[0]; if (MODE == 1) out[1] = bi[0]; }
      *counter = 0;
    }
  }
}

static int reduce_launch(Handle* h, int mode, const void* x, const void* y, long long ndoubles, double* out,
                         cudaStream_t st) {
  // operate on pairs of doubles; complex elements are exactly one pair
  if (((uintptr_t)x % 16) || (y && ((uintptr_t)y % 16)))
    return set_err(h, TNB_ERR_BAD_ARG, "reduction: pointers must be 16-byte aligned");
  const long long n2 = ndoubles / 2;
  const int tail = (int)(ndoubles & 1);
  int grid = (int)std::max<long long>(1, std::min<long long>((n2 + 1023) / 1024, std::min(RED_MAX_BLOCKS, h->num_sms * 4)));
  const double* xt = (const double*)x + 2 * n2;
  const double* yt = y ? (const double*)y + 2 * n2 : nullptr;
  if (mode == 0) reduce_kernel<0><<<grid, 256, 0, st>>>((const double2*)x, (const double2*)y, n2, xt, yt, tail, h->partials, h->counter, out);
  else if (mode == 1) reduce_kernel<1><<<grid, 256, 0, st>>>((const double2*)x, (const double2*)y, n2, xt, yt, tail, h->partials, h->counter, out);
  else reduce_kernel<2><<<grid, 256, 0, st>>>((const double2*)x, nullptr, n2, xt, nullptr, tail, h->partials, h->counter, out);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "reduce launch");
}

int dot_impl(Handle* h, int dtype, int64_t n, const void* x, const void* y, void* result_dev, cudaStream_t st) {
  if (n < 0 || !x || !y || !result_dev) return set_err(h, TNB_ERR_BAD_ARG, "dot: bad argument");
  if (dtype == TNB_F64) return reduce_launch(h, 0, x, y, n, (double*)result_dev, st);
  if (dtype == TNB_C128) return reduce_launch(h, 1, x, y, 2 * n, (double*)result_dev, st);
  return set_err(h, TNB_ERR_UNSUPPORTED, "dot: dtype %d", dtype);
}

int nrm2_impl(Handle* h, int dtype, int64_t n, const void* x, double* result_dev, cudaStream_t st) {
  if (n < 0 || !x || !result_dev) return set_err(h, TNB_ERR_BAD_ARG, "nrm2: bad argument");
  if (dtype != TNB_F64 && dtype != TNB_C128) return set_err(h, TNB_ERR_UNSUPPORTED, "nrm2: dtype %d", dtype);
  return reduce_launch(h, 2, x, nullptr, dtype == TNB_C128 ? 2 * n : n, result_dev, st);
}

__global__ void __launch_bounds__(256) scale_inv_generic_kernel(double* y, const double* x, long long nd,
                                                               const double* scal, double tol) {
  const double b = scal[0];
  const double f = b > tol ? 1.0 / b : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * f;
}

int scale_inv_dev_impl(Handle* h, int dtype, int64_t n, const void* x, void* y, const double* scal, double tol,
                       cudaStream_t st) {
  const long long nd = dtype == TNB_C128 ? 2 * n : n;
  if (nd <= 0) return TNB_OK;
  const int grid = grid_for(h, nd, 256 * 4);
  scale_inv_generic_kernel<<<grid, 256, 0, st>>>((double*)y, (const double*)x, nd, scal, tol);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "scale_inv launch");
}

// ------------------------------------------------------------------------------------
// spectrum truncation -- CPU rule of [EXT] NDTensors truncate!, one kernel, one readback.
// out[0] = truncerr, out[1] = docut, out[2] = n_keep.
// The discard walk is inherently sequential (running sum from the tail); it runs on one
// thread over data staged by the whole block, after a block-wide sum for the scale.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) truncate_kernel(const double* __restrict__ P, int len, int maxdim,
                                                        int mindim, double cutoff, int flags, double* out) {
  __shared__ double red[32];
  __shared__ double scale_s;
  // scale = sum(P) (relative cutoff) -- block reduction
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) s += fmax(P[i], 0.0);  // negatives are zeroed first
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    scale_s = t;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int origm = len;
  double truncerr = 0.0, docut = 0.0;
  int n = origm;
  if (P[0] <= 0.0) { out[0] = 0.0; out[1] = 0.0; out[2] = 1.0; return; }
  if (origm == 1) { out[0] = 0.0; out[1] = P[0] / 2; out[2] = 1.0; return; }
  auto w = [&](int i) { double v = P[i]; return v < 0.0 ? 0.0 : v; };  // trailing negatives count as 0
  while (n > maxdim) { truncerr += w(n - 1); --n; }
  if (flags & TNB_TRUNC_ABSOLUTE_CUTOFF) {
    while (n > mindim && w(n - 1) <= cutoff) { truncerr += w(n - 1); --n; }
  } else {
    double scale = 1.0;
    if (!(flags & TNB_TRUNC_NO_RELATIVE)) { scale = scale_s; if (scale == 0.0) scale = 1.0; }
    const double thr = cutoff * scale;
    while (n > mindim && (truncerr + w(n - 1) <= thr)) { truncerr += w(n - 1); --n; }
    truncerr /= scale;
  }
  if (n < 1) n = 1;
  if (n < origm) {
    const double a = w(n - 1), b = w(n);
    docut = (a + b) / 2;
    if (fabs(a - b) < 1e-3 * a) docut += 1e-3 * a;
  }
  out[0] = truncerr; out[1] = docut; out[2] = (double)n;
}

int truncate_impl(Handle* h, const double* P_dev, int64_t len, int64_t maxdim, int64_t mindim, double cutoff,
                  int flags, int64_t* n_keep, double* truncerr, double* docut, cudaStream_t st) {
  if (!P_dev || len < 1 || len > 2147483647LL) return set_err(h, TNB_ERR_BAD_ARG, "truncate: bad length");
  if (maxdim < 1 || maxdim > len) maxdim = len;
  if (mindim > maxdim) mindim = maxdim;
  if (mindim < 1) mindim = 1;
  if (cutoff < 0) cutoff = 0;
  truncate_kernel<<<1, 1024, 0, st>>>(P_dev, (int)len, (int)maxdim, (int)mindim, cutoff, flags, h->scal + 8);
  h->launches++;
  TNB_CUDA(h, cudaGetLastError());
  TNB_CUDA(h, cudaMemcpyAsync(h->scal_host + 8, h->scal + 8, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TNB_CUDA(h, cudaStreamSynchronize(st));
  if (truncerr) *truncerr = h->scal_host[8];
  if (docut) *docut = h->scal_host[9];
  if (n_keep) *n_keep = (int64_t)h->scal_host[10];
  return TNB_OK;
}

}  // namespace tnb
