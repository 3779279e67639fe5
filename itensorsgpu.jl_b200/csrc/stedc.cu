// Real symmetric tridiagonal eigensolver: Cuppen divide & conquer, everything on the device.
//
// Stage 2 of the fast eigensolver behind `eigen(::Hermitian{CuDenseTensor})`
// (/root/reference/src/tensor/culinearalgebra.jl:74-108; the reference calls cuSOLVER syevd!/heevd!).
//
// A uniform binary tree of L = 2^q leaves (<= 64 rows each).  Leaves: two-sided Jacobi in shared memory,
// one CTA per leaf.  Every level merges ALL its node pairs with the same batched kernels:
//   m1  one CTA per merge: z vector, merged sort order, deflation scan (LAPACK laed2 rules), compaction
//   m2  apply the deflation Givens rotations to the eigenvector columns (row-parallel), gather the
//       columns into [non-deflated | deflated] order
//   m3  secular equation, one warp per root: origin shifted to the nearer pole, two-pole rational
//       model step safeguarded by bisection
//   m4  Loewner / Gu-Eisenstat z-hat (keeps the eigenvectors numerically orthogonal)
//   m5  eigenvectors of the rank-one update, normalised
//   m6  Q_parent = Q_children * U on the DMMA contraction kernel
//   m7  sorted order of the merged eigenvalues for the next level
// The only host round trip per level is the readback of the non-deflated counts (GEMM shapes).
#include "tnb_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <vector>

namespace tnb {

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

constexpr int DC_LEAF = 64;
constexpr double DC_EPS = 2.220446049250313e-16;

struct DcBufs {
  long long n;
  double *dcur, *dout;   // eigenvalue of each physical column (ping-pong)
  int *idx, *idxo;       // idx[s+t] = physical column of the t-th smallest eigenvalue of the node starting at s
  double* zph;           // z by physical column
  double *dsg, *zsg;     // global fallback for the sorted d / z when they do not fit shared memory
  int* col;              // physical column of each sorted position
  int* flag;             // 1 = deflated
  int* dlist;            // deflated positions in ascending order of value
  int* gord;             // gather order: new column slot -> old physical column
  double *dk, *zk;       // non-deflated poles / weights (ascending), at the node offset
  int* org;              // secular roots: lam_i = dk[org_i] + mu_i
  double *mu, *zhat;
  int *kcnt, *nrot;      // per merge of the current level
  double* rho;
  int* rot_pq;           // rotation column pairs (2 ints each), at 2*s
  double* rot_cs;        // rotation (c, s), at 2*s
  const int *m_s, *m_n1, *m_N;   // merge descriptors of the current level
};

// ------------------------------------------------------------------------------------ scaling
// The deflation tolerance 8 eps max(|d|, |z|) compares eigenvalue-scale quantities with the O(1) entries of z, so it
// is only meaningful for |T| ~ 1 (LAPACK dstedc scales for the same reason): T is scaled to unit max-norm on
// entry and the eigenvalues are scaled back at the end.  scal[0] = max(|d|, |e|), scal[1] = 1/scal[0] (1 if zero).
__global__ void __launch_bounds__(1024) dc_maxnorm_kernel(const double* __restrict__ d, const double* __restrict__ e,
                                                           long long n, double* scal) {
  __shared__ double red[32];
  double m = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    m = fmax(m, fabs(d[i]));
    if (i < n - 1) m = fmax(m, fabs(e[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 32; ++w) t = fmax(t, red[w]);
    scal[0] = t;
    scal[1] = (t > 0.0 && isfinite(t)) ? 1.0 / t : 1.0;
  }
}

// x[i] *= scal[which]
__global__ void dc_scale_kernel(double* x, long long n, const double* __restrict__ scal, int which) {
  const double f = scal[which];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= f;
}

// ------------------------------------------------------------------------------------ leaves
__global__ void dc_tear_kernel(double* d, const double* __restrict__ e, const int* __restrict__ splits, int nsplit) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsplit) return;
  const int m = splits[t];
  const double r = fabs(e[m - 1]);
  d[m - 1] -= r;
  d[m] -= r;
}

__global__ void __launch_bounds__(256) dc_leaf_kernel(const double* __restrict__ d, const double* __restrict__ e,
                                                       const int* __restrict__ bounds, double* Q, long long n, double* dcur,
                                                       int* idx) {
  constexpr int N = DC_LEAF, LD = N + 1, H = N / 2;
  extern __shared__ __align__(16) double leaf_smem[];
  double* S = leaf_smem;
  double* W = leaf_smem + N * LD;
  __shared__ double rot_c[H], rot_s[H];
  __shared__ double red[8];
  __shared__ double thr_s;
  __shared__ int done;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int s = bounds[blockIdx.x], nl = bounds[blockIdx.x + 1] - s;
  for (int x = tid; x < N * LD; x += nt) { S[x] = 0.0; W[x] = 0.0; }
  __syncthreads();
  for (int i = tid; i < N; i += nt) {
    W[i * LD + i] = 1.0;
    if (i < nl) S[i * LD + i] = d[s + i];
    if (i < nl - 1) { const double v = e[s + i]; S[i * LD + i + 1] = v; S[(i + 1) * LD + i] = v; }
  }
  __syncthreads();
  {
    double a = 0.0;
    for (int x = tid; x < N * N; x += nt) { const double v = S[(x / N) * LD + (x % N)]; a += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) red[tid >> 5] = a;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int w = 0; w < (nt >> 5); ++w) t += red[w]; thr_s = 1e-17 * sqrt(t); }
    __syncthreads();
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double mx = 0.0;
    for (int x = tid; x < N * N; x += nt) {
      const int i = x % N, j = x / N;
      if (i < j) mx = fmax(mx, fabs(S[j * LD + i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int w = 0; w < (nt >> 5); ++w) t = fmax(t, red[w]); done = (t <= thr_s); }
    __syncthreads();
    if (done) break;
    for (int round = 0; round < N - 1; ++round) {
      if (tid < H) {
        const int k = tid;
        int p = (k == 0) ? (N - 1) : (round + k) % (N - 1);
        int q = (round + (N - 1) - k) % (N - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        const double spq = S[q * LD + p];
        double c = 1.0, sn = 0.0;
        if (spq != 0.0) {
          const double app = S[p * LD + p], aqq = S[q * LD + q];
          const double tau = (aqq - app) / (2.0 * spq);
          const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          sn = t * c;
        }
        rot_c[k] = c; rot_s[k] = sn;
      }
      __syncthreads();
      for (int x = tid; x < H * N * 2; x += nt) {
        const int i = x % N, k = (x / N) % H, which = x / (N * H);
        const double c = rot_c[k], sn = rot_s[k];
        if (sn == 0.0) continue;
        int p = (k == 0) ? (N - 1) : (round + k) % (N - 1);
        int q = (round + (N - 1) - k) % (N - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double* X = which ? W : S;
        const double xp = X[p * LD + i], xq = X[q * LD + i];
        X[p * LD + i] = c * xp - sn * xq;
        X[q * LD + i] = sn * xp + c * xq;
      }
      __syncthreads();
      for (int x = tid; x < H * N; x += nt) {
        const int j = x % N, k = x / N;
        const double c = rot_c[k], sn = rot_s[k];
        if (sn == 0.0) continue;
        int p = (k == 0) ? (N - 1) : (round + k) % (N - 1);
        int q = (round + (N - 1) - k) % (N - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        const double xp = S[j * LD + p], xq = S[j * LD + q];
        S[j * LD + p] = c * xp - sn * xq;
        S[j * LD + q] = sn * xp + c * xq;
      }
      __syncthreads();
      if (tid < H && rot_s[tid] != 0.0) {
        const int k = tid;
        int p = (k == 0) ? (N - 1) : (round + k) % (N - 1);
        int q = (round + (N - 1) - k) % (N - 1);
        S[q * LD + p] = 0.0;
        S[p * LD + q] = 0.0;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // eigenvalues, ascending order, eigenvectors
  if (tid < nl) {
    const double v = S[tid * LD + tid];
    int rank = 0;
    for (int j = 0; j < nl; ++j) {
      const double u = S[j * LD + j];
      if (u < v || (u == v && j < tid)) ++rank;
    }
    dcur[s + tid] = v;
    idx[s + rank] = s + tid;
  }
  for (int x = tid; x < nl * nl; x += nt) {
    const int r = x % nl, c = x / nl;
    Q[(size_t)(s + r) + (size_t)(s + c) * n] = W[c * LD + r];
  }
}

// ------------------------------------------------------------------------------------ m1
__global__ void __launch_bounds__(1024) dc_m1_kernel(DcBufs B, const double* __restrict__ Qin, const double* __restrict__ e,
                                                      int use_smem) {
  extern __shared__ __align__(16) double m1_smem[];
  __shared__ double redA[32], redB[32];
  __shared__ int cnts[1024];
  __shared__ double tol_s, rho_s;
  __shared__ int k_s;
  const int m = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int s = B.m_s[m], n1 = B.m_n1[m], N = B.m_N[m];
  const long long n = B.n;
  double* ds = use_smem ? m1_smem : B.dsg + s;
  double* zs = use_smem ? m1_smem + N : B.zsg + s;
  const double ecoup = e[s + n1 - 1];
  const double sgn = ecoup >= 0 ? 1.0 : -1.0;
  const double isq2 = 0.70710678118654752440;
  // A: z by physical column
  for (int t = tid; t < N; t += nt) {
    const int c = s + t;
    const double v = (t < n1) ? Qin[(size_t)(s + n1 - 1) + (size_t)c * n] : sgn * Qin[(size_t)(s + n1) + (size_t)c * n];
    B.zph[c] = v * isq2;
  }
  __syncthreads();
  // B: merged order of the two sorted children
  const int n2 = N - n1;
  for (int t = tid; t < N; t += nt) {
    int c, rank;
    double v;
    if (t < n1) {
      c = B.idx[s + t]; v = B.dcur[c];
      int lo = 0, hi = n2;           // number of child-2 values strictly less than v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (B.dcur[B.idx[s + n1 + mid]] < v) lo = mid + 1; else hi = mid; }
      rank = t + lo;
    } else {
      const int b = t - n1;
      c = B.idx[s + t]; v = B.dcur[c];
      int lo = 0, hi = n1;           // number of child-1 values <= v
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (B.dcur[B.idx[s + mid]] <= v) lo = mid + 1; else hi = mid; }
      rank = b + lo;
    }
    ds[rank] = v; zs[rank] = B.zph[c]; B.col[s + rank] = c;
  }
  __syncthreads();
  // C: tolerance
  {
    double dm = 0.0, zm = 0.0;
    for (int t = tid; t < N; t += nt) { dm = fmax(dm, fabs(ds[t])); zm = fmax(zm, fabs(zs[t])); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o)); zm = fmax(zm, __shfl_xor_sync(0xffffffffu, zm, o)); }
    if ((tid & 31) == 0) { redA[tid >> 5] = dm; redB[tid >> 5] = zm; }
    __syncthreads();
    if (tid == 0) {
      double a = 0, b = 0;
      for (int w = 0; w < (nt >> 5); ++w) { a = fmax(a, redA[w]); b = fmax(b, redB[w]); }
      tol_s = 8.0 * DC_EPS * fmax(a, b);
      rho_s = 2.0 * fabs(ecoup);
    }
    __syncthreads();
  }
  // D: serial deflation scan
  if (tid == 0) {
    const double tol = tol_s, rho = rho_s;
    int* flag = B.flag + s;
    int* dlist = B.dlist + s;
    int* rpq = B.rot_pq + 2 * (size_t)s;
    double* rcs = B.rot_cs + 2 * (size_t)s;
    const int* col = B.col + s;
    int pj = -1, nr = 0, nd = 0;
    double lastmax = -1e308;
    auto push_defl = [&](int pos) {
      const double v = ds[pos];
      int u = nd++;
      if (v < lastmax) {
        while (u > 0 && ds[dlist[u - 1]] > v) { dlist[u] = dlist[u - 1]; --u; }
      } else lastmax = v;
      dlist[u] = pos;
    };
    for (int t = 0; t < N; ++t) {
      if (rho * fabs(zs[t]) <= tol) { flag[t] = 1; push_defl(t); continue; }
      flag[t] = 0;
      if (pj < 0) { pj = t; continue; }
      double sn = zs[pj], c = zs[t];
      const double tau = hypot(c, sn);
      const double tt = ds[t] - ds[pj];
      c /= tau; sn = -sn / tau;
      if (fabs(tt * c * sn) <= tol) {
        zs[t] = tau; zs[pj] = 0.0;
        rpq[2 * nr] = col[pj]; rpq[2 * nr + 1] = col[t];
        rcs[2 * nr] = c; rcs[2 * nr + 1] = sn;
        ++nr;
        const double tn = ds[pj] * c * c + ds[t] * sn * sn;
        ds[t] = ds[pj] * sn * sn + ds[t] * c * c;
        ds[pj] = tn;
        flag[pj] = 1;
        push_defl(pj);
      }
      pj = t;
    }
    B.nrot[m] = nr;
    B.rho[m] = rho;
    k_s = N - nd;
    B.kcnt[m] = N - nd;
    __threadfence_block();
  }
  __syncthreads();
  // E: compaction.  Non-deflated keep their (ascending) order; deflated follow in dlist order.
  const int k = k_s;
  {
    const int chunk = (N + nt - 1) / nt;
    const int t0 = tid * chunk, t1 = min(N, t0 + chunk);
    int c = 0;
    for (int t = t0; t < t1; ++t) c += (B.flag[s + t] == 0);
    cnts[tid] = c;
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int w = 0; w < nt; ++w) { const int v = cnts[w]; cnts[w] = acc; acc += v; } }
    __syncthreads();
    int u = cnts[tid];
    for (int t = t0; t < t1; ++t) {
      if (B.flag[s + t] == 0) {
        B.dk[s + u] = ds[t]; B.zk[s + u] = zs[t]; B.gord[s + u] = B.col[s + t];
        ++u;
      }
    }
    for (int u2 = tid; u2 < N - k; u2 += nt) {
      const int pos = B.dlist[s + u2];
      B.gord[s + k + u2] = B.col[s + pos];
      B.dout[s + k + u2] = ds[pos];
    }
  }
}

// ------------------------------------------------------------------------------------ m2
__global__ void __launch_bounds__(128) dc_rotate_kernel(DcBufs B, double* Qin) {
  const int m = blockIdx.y;
  const int nr = B.nrot[m];
  if (nr == 0) return;
  const int s = B.m_s[m], N = B.m_N[m];
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= N) return;
  const int* rpq = B.rot_pq + 2 * (size_t)s;
  const double* rcs = B.rot_cs + 2 * (size_t)s;
  double* row = Qin + (size_t)(s + r);
  for (int t = 0; t < nr; ++t) {
    const int cp = rpq[2 * t], cq = rpq[2 * t + 1];
    const double c = rcs[2 * t], sn = rcs[2 * t + 1];
    const double qp = row[(size_t)cp * B.n], qn = row[(size_t)cq * B.n];
    row[(size_t)cp * B.n] = c * qp + sn * qn;
    row[(size_t)cq * B.n] = -sn * qp + c * qn;
  }
}

// new column slot t of the node: non-deflated (t < k) -> Qtmp (GEMM operand), deflated -> Qout (final)
__global__ void __launch_bounds__(256) dc_gather_kernel(DcBufs B, const double* __restrict__ Qin, double* Qtmp, double* Qout) {
  const int m = blockIdx.z;
  const int s = B.m_s[m], N = B.m_N[m];
  const int t = blockIdx.y;
  if (t >= N) return;
  const int k = B.kcnt[m];
  const int src = B.gord[s + t];
  double* dst = (t < k ? Qtmp : Qout) + (size_t)s + (size_t)(s + t) * B.n;
  const double* sp = Qin + (size_t)s + (size_t)src * B.n;
  for (int r = blockIdx.x * 256 + threadIdx.x; r < N; r += gridDim.x * 256) dst[r] = sp[r];
}

// ------------------------------------------------------------------------------------ m3
struct SecSums { double psi, dpsi, phi, dphi; };

__device__ __forceinline__ SecSums sec_eval(const double* __restrict__ dk, const double* __restrict__ zk, int k, int i, int o,
                                            double x, double rho, bool last, int lane) {
  SecSums r = {0.0, 0.0, 0.0, 0.0};
  const double dorg = dk[o];
  for (int j = lane; j < k; j += 32) {
    const double D = (dk[j] - dorg) - x;
    const double z = zk[j];
    const double t = z * z / D;
    if (last || j <= i) { r.psi += t; r.dpsi += t / D; }
    else { r.phi += t; r.dphi += t / D; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    r.psi += __shfl_xor_sync(0xffffffffu, r.psi, off);
    r.dpsi += __shfl_xor_sync(0xffffffffu, r.dpsi, off);
    r.phi += __shfl_xor_sync(0xffffffffu, r.phi, off);
    r.dphi += __shfl_xor_sync(0xffffffffu, r.dphi, off);
  }
  r.psi *= rho; r.dpsi *= rho; r.phi *= rho; r.dphi *= rho;
  return r;
}

__global__ void __launch_bounds__(128) dc_secular_kernel(DcBufs B) {
  const int m = blockIdx.y;
  const int k = B.kcnt[m];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= k) return;
  const int s = B.m_s[m];
  const double* dk = B.dk + s;
  const double* zk = B.zk + s;
  const double rho = B.rho[m];
  const bool last = (i == k - 1);
  int o;
  double lo, hi;
  if (!last) {
    const double mid = 0.5 * (dk[i + 1] - dk[i]);
    const SecSums f = sec_eval(dk, zk, k, i, i, mid, rho, false, lane);
    if (1.0 + f.psi + f.phi > 0.0) { o = i; lo = 0.0; hi = mid; }
    else { o = i + 1; lo = -mid; hi = 0.0; }
  } else {
    double z2 = 0.0;
    for (int j = lane; j < k; j += 32) z2 += zk[j] * zk[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) z2 += __shfl_xor_sync(0xffffffffu, z2, off);
    o = i; lo = 0.0; hi = rho * z2;
  }
  double x = 0.5 * (lo + hi);
  const double dorg = dk[o];
  for (int it = 0; it < 100; ++it) {
    const SecSums f = sec_eval(dk, zk, k, i, o, x, rho, last, lane);
    const double fv = 1.0 + f.psi + f.phi;
    const double err = 8.0 * DC_EPS * (1.0 + fabs(f.psi) + fabs(f.phi)) + DC_EPS * fabs(x) * (f.dpsi + f.dphi);
    if (fv > 0.0) hi = fmin(hi, x); else lo = fmax(lo, x);
    if (fabs(fv) <= err || hi - lo <= 2.0 * DC_EPS * fmax(fabs(lo), fabs(hi))) break;
    double qa, qb, qc;
    if (!last) {
      const double Di = (dk[i] - dorg) - x, Dj = (dk[i + 1] - dorg) - x;
      const double a = f.dpsi * Di * Di, sp = f.psi - f.dpsi * Di;
      const double b = f.dphi * Dj * Dj, tp = f.phi - f.dphi * Dj;
      const double c = 1.0 + sp + tp;
      qa = c; qb = -(c * (Di + Dj) + a + b); qc = Di * Dj * fv;
    } else {
      const double Di = (dk[i] - dorg) - x;
      const double a = f.dpsi * Di * Di, sp = f.psi - f.dpsi * Di;
      const double c = 1.0 + sp;
      qa = 0.0; qb = -c; qc = c * Di + a;
    }
    double eta = 0.0;
    bool have = false;
    if (qa == 0.0) {
      if (qb != 0.0) { eta = -qc / qb; have = true; }
    } else {
      const double disc = qb * qb - 4.0 * qa * qc;
      if (disc >= 0.0) {
        const double sq = sqrt(disc);
        const double qq = -0.5 * (qb + (qb >= 0.0 ? sq : -sq));
        double e1 = 0.0, e2 = qq / qa;
        bool ok1 = false, ok2 = (x + e2 > lo && x + e2 < hi);
        if (qq != 0.0) { e1 = qc / qq; ok1 = (x + e1 > lo && x + e1 < hi); }
        if (ok1 && ok2) { eta = fabs(e1) <= fabs(e2) ? e1 : e2; have = true; }
        else if (ok1) { eta = e1; have = true; }
        else if (ok2) { eta = e2; have = true; }
      }
    }
    double xn = x + eta;
    if (!have || !(xn > lo && xn < hi) || !isfinite(xn)) xn = 0.5 * (lo + hi);
    if (xn == x) break;
    x = xn;
  }
  if (lane == 0) {
    B.org[s + i] = o;
    B.mu[s + i] = x;
    B.dout[s + i] = dorg + x;
  }
}

// ------------------------------------------------------------------------------------ m4
__global__ void __launch_bounds__(128) dc_zhat_kernel(DcBufs B) {
  const int m = blockIdx.y;
  const int k = B.kcnt[m];
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (j >= k) return;
  const int s = B.m_s[m];
  const double* dk = B.dk + s;
  const int* org = B.org + s;
  const double* mu = B.mu + s;
  const double dj = dk[j];
  double prod = 1.0;
  for (int i = lane; i < k; i += 32) {
    const double num = (dk[org[i]] - dj) + mu[i];     // lam_i - d_j
    if (i == j) prod *= num / B.rho[m];
    else prod *= num / (dk[i] - dj);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, off);
  if (lane == 0) B.zhat[s + j] = copysign(sqrt(fabs(prod)), B.zk[s + j]);
}

// ------------------------------------------------------------------------------------ m5
// U (k x k, ld k) at Ubuf + s*n: column i = zhat_j / (d_j - lam_i), normalised
__global__ void __launch_bounds__(128) dc_vectors_kernel(DcBufs B, double* Ubuf) {
  const int m = blockIdx.y;
  const int k = B.kcnt[m];
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= k) return;
  const int s = B.m_s[m];
  const double* dk = B.dk + s;
  const double* zh = B.zhat + s;
  const double dorg = dk[B.org[s + i]], mui = B.mu[s + i];
  double* U = Ubuf + (size_t)s * B.n + (size_t)i * k;
  double ss = 0.0;
  for (int j = lane; j < k; j += 32) {
    const double v = zh[j] / ((dk[j] - dorg) - mui);
    ss += v * v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const double inv = 1.0 / sqrt(ss);
  for (int j = lane; j < k; j += 32) U[j] = zh[j] / ((dk[j] - dorg) - mui) * inv;
}

// ------------------------------------------------------------------------------------ m7
__global__ void __launch_bounds__(256) dc_order_kernel(DcBufs B) {
  const int m = blockIdx.y;
  const int s = B.m_s[m], N = B.m_N[m];
  const int k = B.kcnt[m];
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= N) return;
  const double* dv = B.dout + s;
  const double v = dv[t];
  int rank;
  if (t < k) {        // a root: deflated values strictly below it
    int lo = 0, hi = N - k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (dv[k + mid] < v) lo = mid + 1; else hi = mid; }
    rank = t + lo;
  } else {            // deflated: roots <= it
    int lo = 0, hi = k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (dv[mid] <= v) lo = mid + 1; else hi = mid; }
    rank = (t - k) + lo;
  }
  B.idxo[s + rank] = s + t;
}

// ------------------------------------------------------------------------------------ driver
size_t stedc_ws_bytes(int64_t n) {
  size_t t = 4 * al256((size_t)n * n * sizeof(double));       // Qa, Qb, Qtmp, U
  t += 12 * al256((size_t)n * sizeof(double)) + 8 * al256((size_t)n * sizeof(int)) + 256;
  t += 2 * al256(2 * (size_t)n * sizeof(double)) + 8 * al256((size_t)(n / 16 + 64) * sizeof(int));
  return t + (1 << 16);
}

// d (n), e (n-1) on the device; both are destroyed.  On return *Qres points at the n x n eigenvector
// matrix (ld n, arena memory), *dres at the eigenvalue of each column and *idxres at the ascending order.
int stedc_impl(Handle* h, int64_t n, double* d, double* e, double** Qres, double** dres, int** idxres, cudaStream_t st) {
  // Index arithmetic is int / n*n fits size_t; merges larger than 13824 rows (216 KB of shared memory for d and z)
  // keep the deflation scan's working vectors in global memory instead.  32768 bounds the 4 n^2 workspace at 34 GB.
  if (n > 32768) return set_err(h, TNB_ERR_UNSUPPORTED, "stedc: n = %lld > 32768", (long long)n);
  int L = 1;
  while ((n + L - 1) / L > DC_LEAF) L *= 2;
  std::vector<int> bounds(L + 1);
  for (int t = 0; t <= L; ++t) bounds[t] = (int)((long long)t * n / L);
  int levels = 0;
  while ((1 << levels) < L) ++levels;
  // descriptors of every level, concatenated
  std::vector<int> hs, hn1, hN, splits;
  std::vector<int> lvl_off(levels + 1, 0);
  for (int l = 1; l <= levels; ++l) {
    const int step = 1 << l, nm = L / step;
    lvl_off[l - 1] = (int)hs.size();
    for (int m = 0; m < nm; ++m) {
      const int s = bounds[m * step], mid = bounds[m * step + step / 2], en = bounds[(m + 1) * step];
      hs.push_back(s); hn1.push_back(mid - s); hN.push_back(en - s);
      splits.push_back(mid);
    }
  }
  lvl_off[levels] = (int)hs.size();
  const int nmtot = (int)hs.size();

  void *Qa, *Qb, *Qt, *Ub;
  TNB_TRY(ws_alloc(h, (size_t)n * n * sizeof(double), &Qa));
  TNB_TRY(ws_alloc(h, (size_t)n * n * sizeof(double), &Qb));
  TNB_TRY(ws_alloc(h, (size_t)n * n * sizeof(double), &Qt));
  TNB_TRY(ws_alloc(h, (size_t)n * n * sizeof(double), &Ub));
  DcBufs B;
  B.n = n;
  void* p;
  auto dalloc = [&](double** out, size_t cnt) { int rc = ws_alloc(h, cnt * sizeof(double), &p); *out = (double*)p; return rc; };
  auto ialloc = [&](int** out, size_t cnt) { int rc = ws_alloc(h, cnt * sizeof(int), &p); *out = (int*)p; return rc; };
  TNB_TRY(dalloc(&B.dcur, n)); TNB_TRY(dalloc(&B.dout, n)); TNB_TRY(dalloc(&B.zph, n));
  TNB_TRY(dalloc(&B.dsg, n)); TNB_TRY(dalloc(&B.zsg, n)); TNB_TRY(dalloc(&B.dk, n)); TNB_TRY(dalloc(&B.zk, n));
  TNB_TRY(dalloc(&B.mu, n)); TNB_TRY(dalloc(&B.zhat, n)); TNB_TRY(dalloc(&B.rot_cs, 2 * n));
  TNB_TRY(ialloc(&B.idx, n)); TNB_TRY(ialloc(&B.idxo, n)); TNB_TRY(ialloc(&B.col, n)); TNB_TRY(ialloc(&B.flag, n));
  TNB_TRY(ialloc(&B.dlist, n)); TNB_TRY(ialloc(&B.gord, n)); TNB_TRY(ialloc(&B.org, n)); TNB_TRY(ialloc(&B.rot_pq, 2 * n));
  const int nmax = std::max(1, L / 2);
  TNB_TRY(ialloc(&B.kcnt, nmax)); TNB_TRY(ialloc(&B.nrot, nmax)); TNB_TRY(dalloc(&B.rho, nmax));
  int *d_s, *d_n1, *d_N, *d_splits, *d_bounds;
  TNB_TRY(ialloc(&d_s, nmtot + 1)); TNB_TRY(ialloc(&d_n1, nmtot + 1)); TNB_TRY(ialloc(&d_N, nmtot + 1));
  TNB_TRY(ialloc(&d_splits, nmtot + 1)); TNB_TRY(ialloc(&d_bounds, L + 1));
  if (nmtot) {
    TNB_CUDA(h, cudaMemcpyAsync(d_s, hs.data(), nmtot * sizeof(int), cudaMemcpyHostToDevice, st));
    TNB_CUDA(h, cudaMemcpyAsync(d_n1, hn1.data(), nmtot * sizeof(int), cudaMemcpyHostToDevice, st));
    TNB_CUDA(h, cudaMemcpyAsync(d_N, hN.data(), nmtot * sizeof(int), cudaMemcpyHostToDevice, st));
    TNB_CUDA(h, cudaMemcpyAsync(d_splits, splits.data(), nmtot * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  TNB_CUDA(h, cudaMemcpyAsync(d_bounds, bounds.data(), (L + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  // both ping-pong buffers start as zero: a node only ever writes its own diagonal block, and the parent's
  // gather reads the full parent row range of each child column
  TNB_CUDA(h, cudaMemsetAsync(Qa, 0, (size_t)n * n * sizeof(double), st));
  TNB_CUDA(h, cudaMemsetAsync(Qb, 0, (size_t)n * n * sizeof(double), st));
  double* nscal;
  TNB_TRY(dalloc(&nscal, 2));
  dc_maxnorm_kernel<<<1, 1024, 0, st>>>(d, e, n, nscal);
  dc_scale_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 64), 256, 0, st>>>(d, n, nscal, 1);
  if (n > 1) dc_scale_kernel<<<(int)std::min<int64_t>((n + 254) / 256, 64), 256, 0, st>>>(e, n - 1, nscal, 1);
  h->launches += 3;
  if (nmtot) {
    dc_tear_kernel<<<(nmtot + 127) / 128, 128, 0, st>>>(d, e, d_splits, nmtot);
    h->launches++;
  }
  constexpr int LEAF_SMEM = 2 * DC_LEAF * (DC_LEAF + 1) * (int)sizeof(double);
  TNB_ONCE_PER_DEVICE(h, TNB_CUDA(h, cudaFuncSetAttribute(dc_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM)));
  dc_leaf_kernel<<<L, 256, LEAF_SMEM, st>>>(d, e, d_bounds, (double*)Qa, n, B.dcur, B.idx);
  h->launches++;
  TNB_CUDA(h, cudaStreamSynchronize(st));     // host descriptor vectors are about to be reused / go out of scope

  TNB_ONCE_PER_DEVICE(h, TNB_CUDA(h, cudaFuncSetAttribute(dc_m1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024)));
  double *Qin = (double*)Qa, *Qout = (double*)Qb;
  std::vector<int> hk(nmax);
  for (int l = 1; l <= levels; ++l) {
    const int off = lvl_off[l - 1], nm = lvl_off[l] - off;
    B.m_s = d_s + off; B.m_n1 = d_n1 + off; B.m_N = d_N + off;
    int maxN = 0;
    for (int m = 0; m < nm; ++m) maxN = std::max(maxN, hN[off + m]);
    const size_t smem = (size_t)maxN * 16;
    static const bool force_global = getenv("TNB_DC_NOSMEM") != nullptr;    // test hook for the large-merge path
    const int use_smem = smem <= 216 * 1024 && !force_global;
    dc_m1_kernel<<<nm, 1024, use_smem ? smem : 0, st>>>(B, Qin, e, use_smem);
    h->launches++;
    TNB_CUDA(h, cudaMemcpyAsync(hk.data(), B.kcnt, nm * sizeof(int), cudaMemcpyDeviceToHost, st));
    dc_rotate_kernel<<<dim3((maxN + 127) / 128, nm), 128, 0, st>>>(B, Qin);
    dc_gather_kernel<<<dim3(std::max(1, std::min(8, (maxN + 255) / 256)), maxN, nm), 256, 0, st>>>(B, Qin, (double*)Qt, Qout);
    dc_secular_kernel<<<dim3((maxN + 3) / 4, nm), 128, 0, st>>>(B);
    dc_zhat_kernel<<<dim3((maxN + 3) / 4, nm), 128, 0, st>>>(B);
    dc_vectors_kernel<<<dim3((maxN + 3) / 4, nm), 128, 0, st>>>(B, (double*)Ub);
    dc_order_kernel<<<dim3((maxN + 255) / 256, nm), 256, 0, st>>>(B);
    h->launches += 6;
    TNB_CUDA(h, cudaStreamSynchronize(st));
    TNB_CUDA(h, cudaGetLastError());
    for (int m = 0; m < nm; ++m) {
      const int s = hs[off + m], N = hN[off + m], k = hk[m];
      if (k <= 0) continue;
      TNB_TRY(gemm_impl(h, TNB_F64, 'N', 'N', N, k, k, nullptr, (double*)Qt + (size_t)s + (size_t)s * n, n,
                        (double*)Ub + (size_t)s * n, k, nullptr, Qout + (size_t)s + (size_t)s * n, n, st));
    }
    std::swap(Qin, Qout);
    std::swap(B.dcur, B.dout);
    std::swap(B.idx, B.idxo);
  }
  dc_scale_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 64), 256, 0, st>>>(B.dcur, n, nscal, 0);      // eigenvalues back to T's scale
  h->launches++;
  *Qres = Qin; *dres = B.dcur; *idxres = B.idx;
  return check_cuda(h, cudaGetLastError(), "stedc");
}

}  // namespace tnb
