// Blocked Householder QR with explicit thin Q (compact-WY, trailing updates on the DMMA GEMM).
//
// Replaces `qr(::CuDenseTensor{_,2})` -> cuSOLVER geqrf + CuMatrix(Q) (orgqr)
// (/root/reference/src/tensor/culinearalgebra.jl:110-121).  diag(R) is made real >= 0.
#include "tnb_internal.h"

#include <algorithm>
#include <vector>

namespace tnb {

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

template <bool CPLX> struct QT { using T = double; };
template <> struct QT<true> { using T = double2; };

template <typename T> __device__ __forceinline__ T tzero();
template <> __device__ __forceinline__ double tzero<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 tzero<double2>() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double tadd(double a, double b) { return a + b; }
__device__ __forceinline__ double2 tadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double tsub(double a, double b) { return a - b; }
__device__ __forceinline__ double2 tsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double tmul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 tmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double tconj(double a) { return a; }
__device__ __forceinline__ double2 tconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double tabs2(double a) { return a * a; }
__device__ __forceinline__ double tabs2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double tshfl(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 tshfl(double2 v, int o) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}

constexpr int QR_NB = 32;

// Unblocked Householder factorisation of one panel P (rows x jb, jb <= 32, leading dim ld), one CTA of
// 1024 threads (32 warps).  On exit: R on and above the diagonal, the essential parts of the reflectors
// v_c below it (v_c[c] = 1 implicit), tau[c], and the jb x jb upper-triangular T of the compact-WY form
// H_0 H_1 ... = I - V T V^H  (LAPACK larft, forward/columnwise).
template <typename T>
__global__ void __launch_bounds__(1024, 1) qr_panel_kernel(T* P, long long ld, long long rows, int jb, T* tau, T* Tm) {
  __shared__ T sh_w[32];
  __shared__ T sh_red[32];
  __shared__ T sh_tau;
  __shared__ T sh_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = 0; c < jb; ++c) {
    T* col = P + (size_t)c * ld;
    // ---- larfg on col[c:rows)
    double xn = 0.0;
    for (long long i = c + 1 + tid; i < rows; i += 1024) xn += tabs2(col[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xn += __shfl_xor_sync(0xffffffffu, xn, o);
    if (lane == 0) ((double*)sh_red)[warp] = xn;
    __syncthreads();
    if (tid == 0) {
      double s = 0;
      for (int w = 0; w < 32; ++w) s += ((double*)sh_red)[w];
      const T alpha = col[c];
      T t, sc;
      double ar, ai = 0.0;
      if constexpr (sizeof(T) == 16) { ar = alpha.x; ai = alpha.y; } else ar = alpha;
      if (s == 0.0 && ai == 0.0) {
        t = tzero<T>();
        if constexpr (sizeof(T) == 16) sc = make_double2(0.0, 0.0); else sc = 0.0;
      } else {
        const double nrm = sqrt(ar * ar + ai * ai + s);
        const double beta = ar >= 0 ? -nrm : nrm;
        if constexpr (sizeof(T) == 16) {
          t = make_double2((beta - ar) / beta, -ai / beta);
          const double dr = ar - beta, di = ai, dd = dr * dr + di * di;
          sc = make_double2(dr / dd, -di / dd);
          col[c] = make_double2(beta, 0.0);
        } else {
          t = (beta - ar) / beta;
          sc = 1.0 / (ar - beta);
          col[c] = beta;
        }
      }
      sh_tau = t; sh_scale = sc;
      tau[c] = t;
    }
    __syncthreads();
    const T tc = sh_tau, sc = sh_scale;
    for (long long i = c + 1 + tid; i < rows; i += 1024) col[i] = tmul(col[i], sc);
    __syncthreads();
    // ---- apply H_c^H = I - conj(tau) v v^H to the remaining panel columns: one warp per column
    const T ctc = tconj(tc);
    for (int j = c + 1 + warp; j < jb; j += 32) {
      T* cj = P + (size_t)j * ld;
      T w = (lane == 0) ? cj[c] : tzero<T>();     // v[c] = 1
      for (long long i = c + 1 + lane; i < rows; i += 32) w = tadd(w, tmul(tconj(col[i]), cj[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w = tadd(w, tshfl(w, o));
      const T f = tmul(ctc, w);
      if (lane == 0) cj[c] = tsub(cj[c], f);
      for (long long i = c + 1 + lane; i < rows; i += 32) cj[i] = tsub(cj[i], tmul(f, col[i]));
    }
    __syncthreads();
  }
  // ---- T factor: T(c,c) = tau_c ; T(0:c,c) = -tau_c * T(0:c,0:c) * (V(:,0:c)^H v_c)
  for (int c = 0; c < jb; ++c) {
    const T* vc = P + (size_t)c * ld;
    if (warp < c) {   // warp j computes z_j = v_j^H v_c  (j < c); rows >= c only (v_c is zero above c)
      const int j = warp;
      const T* vj = P + (size_t)j * ld;
      T w = (lane == 0) ? tconj(vj[c]) : tzero<T>();   // row c: v_c[c] = 1
      for (long long i = c + 1 + lane; i < rows; i += 32) w = tadd(w, tmul(tconj(vj[i]), vc[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w = tadd(w, tshfl(w, o));
      if (lane == 0) sh_w[j] = w;
    }
    __syncthreads();
    if (tid <= c) {
      const T tcq = tau[c];
      if (tid == c) Tm[c + c * QR_NB] = tcq;
      else {
        T acc = tzero<T>();
        for (int l = tid; l < c; ++l) acc = tadd(acc, tmul(Tm[tid + l * QR_NB], sh_w[l]));
        T neg = tmul(tcq, acc);
        if constexpr (sizeof(T) == 16) neg = make_double2(-neg.x, -neg.y); else neg = -neg;
        Tm[tid + c * QR_NB] = neg;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < QR_NB * QR_NB; e += 1024) {
    const int i = e % QR_NB, j = e / QR_NB;
    if (i > j || i >= jb || j >= jb) Tm[e] = tzero<T>();
  }
}

// V (rows x jb, ldv) <- unit-lower-trapezoidal reflectors taken from the factored panel
template <typename T>
__global__ void extract_v_kernel(const T* __restrict__ P, long long ld, long long rows, int jb, T* V, long long ldv) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < rows * jb; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e % rows;
    const int j = (int)(e / rows);
    T v;
    if (i < j) v = tzero<T>();
    else if (i == j) { if constexpr (sizeof(T) == 16) v = make_double2(1.0, 0.0); else v = 1.0; }
    else v = P[i + (size_t)j * ld];
    V[i + (size_t)j * ldv] = v;
  }
}

// R (k x n) <- upper triangle of Aw's first k rows, rows scaled by conj(phase_i) so diag(R) >= 0;
// phase_i = R_ii/|R_ii| is stored for the matching column scaling of Q.
template <typename T>
__global__ void extract_r_kernel(const T* __restrict__ Aw, long long lda, long long k, long long n, T* R, T* phase) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < k * n; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e % k, j = e / k;
    const T d = Aw[i + i * lda];
    const double a = sqrt(tabs2(d));
    T ph;
    if constexpr (sizeof(T) == 16) ph = a > 0 ? make_double2(d.x / a, d.y / a) : make_double2(1.0, 0.0);
    else ph = (d < 0) ? -1.0 : 1.0;
    if (j == 0) phase[i] = ph;
    T v = (i <= j) ? tmul(tconj(ph), Aw[i + j * lda]) : tzero<T>();
    if constexpr (sizeof(T) == 16) { if (i == j) v.y = 0.0; }
    R[e] = v;
  }
}

template <typename T>
__global__ void init_q_kernel(T* Q, long long m, long long k) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m * k; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e % m, j = e / m;
    if constexpr (sizeof(T) == 16) Q[e] = make_double2(i == j ? 1.0 : 0.0, 0.0); else Q[e] = (i == j) ? 1.0 : 0.0;
  }
}

template <typename T>
__global__ void scale_cols_kernel(T* Q, long long m, long long k, const T* __restrict__ phase) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m * k; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / m;
    Q[e] = tmul(Q[e], phase[j]);
  }
}

size_t qr_ws_bytes(int dtype, int64_t m, int64_t n) {
  const size_t es = elsize(dtype);
  const int64_t k = std::min(m, n);
  const int64_t npan = (k + QR_NB - 1) / QR_NB;
  return al256((size_t)m * n * es) + al256((size_t)m * QR_NB * es) + al256((size_t)npan * QR_NB * QR_NB * es) +
         2 * al256((size_t)QR_NB * std::max(n, k) * es) + al256((size_t)k * es) * 2 + 4096;
}

template <bool CPLX>
static int qr_core(Handle* h, int64_t m, int64_t n, const void* A, void* Q, void* R, cudaStream_t st) {
  using T = typename QT<CPLX>::T;
  const int dtype = CPLX ? TNB_C128 : TNB_F64;
  const size_t es = sizeof(T);
  const int64_t k = std::min(m, n);
  const int64_t npan = (k + QR_NB - 1) / QR_NB;
  void *Aw, *Vp, *Ts, *W1, *W2, *tau, *phase;
  TNB_TRY(ws_alloc(h, (size_t)m * n * es, &Aw));
  TNB_TRY(ws_alloc(h, (size_t)m * QR_NB * es, &Vp));
  TNB_TRY(ws_alloc(h, (size_t)npan * QR_NB * QR_NB * es, &Ts));
  TNB_TRY(ws_alloc(h, (size_t)QR_NB * std::max(n, k) * es, &W1));
  TNB_TRY(ws_alloc(h, (size_t)QR_NB * std::max(n, k) * es, &W2));
  TNB_TRY(ws_alloc(h, (size_t)k * es, &tau));
  TNB_TRY(ws_alloc(h, (size_t)k * es, &phase));
  TNB_CUDA(h, cudaMemcpyAsync(Aw, A, (size_t)m * n * es, cudaMemcpyDeviceToDevice, st));
  double one[2] = {1.0, 0.0}, mone[2] = {-1.0, 0.0};
  const int g = h->num_sms * 4;
  for (int64_t p = 0; p < npan; ++p) {
    const int64_t j0 = p * QR_NB;
    const int jb = (int)std::min<int64_t>(QR_NB, k - j0);
    const int64_t rows = m - j0;
    T* P = (T*)Aw + j0 + j0 * m;
    T* Tp = (T*)Ts + p * QR_NB * QR_NB;
    qr_panel_kernel<T><<<1, 1024, 0, st>>>(P, m, rows, jb, (T*)tau + j0, Tp);
    h->launches++;
    const int64_t nt = n - (j0 + jb);
    if (nt > 0) {
      extract_v_kernel<T><<<g, 256, 0, st>>>(P, m, rows, jb, (T*)Vp, rows);
      h->launches++;
      T* At = (T*)Aw + j0 + (j0 + jb) * m;
      // W1 = V^H A_trail ; W2 = T^H W1 ; A_trail -= V W2      (apply H_jb^H ... H_1^H = (I - V T V^H)^H)
      TNB_TRY(gemm_impl(h, dtype, 'C', 'N', jb, nt, rows, nullptr, Vp, rows, At, m, nullptr, W1, jb, st));
      TNB_TRY(gemm_impl(h, dtype, 'C', 'N', jb, nt, jb, nullptr, Tp, QR_NB, W1, jb, nullptr, W2, jb, st));
      TNB_TRY(gemm_impl(h, dtype, 'N', 'N', rows, nt, jb, mone, Vp, rows, W2, jb, one, At, m, st));
    }
  }
  extract_r_kernel<T><<<g, 256, 0, st>>>((const T*)Aw, m, k, n, (T*)R, (T*)phase);
  init_q_kernel<T><<<g, 256, 0, st>>>((T*)Q, m, k);
  h->launches += 2;
  // Q = H_1 H_2 ... H_k [I;0]: apply block reflectors in reverse order, Q_sub = (I - V T V^H) Q_sub
  for (int64_t p = npan - 1; p >= 0; --p) {
    const int64_t j0 = p * QR_NB;
    const int jb = (int)std::min<int64_t>(QR_NB, k - j0);
    const int64_t rows = m - j0;
    const int64_t nc = k - j0;
    T* P = (T*)Aw + j0 + j0 * m;
    T* Tp = (T*)Ts + p * QR_NB * QR_NB;
    T* Qs = (T*)Q + j0 + j0 * m;
    extract_v_kernel<T><<<g, 256, 0, st>>>(P, m, rows, jb, (T*)Vp, rows);
    h->launches++;
    TNB_TRY(gemm_impl(h, dtype, 'C', 'N', jb, nc, rows, nullptr, Vp, rows, Qs, m, nullptr, W1, jb, st));
    TNB_TRY(gemm_impl(h, dtype, 'N', 'N', jb, nc, jb, nullptr, Tp, QR_NB, W1, jb, nullptr, W2, jb, st));
    TNB_TRY(gemm_impl(h, dtype, 'N', 'N', rows, nc, jb, mone, Vp, rows, W2, jb, one, Qs, m, st));
  }
  scale_cols_kernel<T><<<g, 256, 0, st>>>((T*)Q, m, k, (const T*)phase);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "qr");
}

int qr_impl(Handle* h, int dtype, int64_t m, int64_t n, const void* A, void* Q, void* R, cudaStream_t st) {
  if (m < 1 || n < 1) return set_err(h, TNB_ERR_BAD_ARG, "qr: empty matrix");
  if (dtype == TNB_F64) return qr_core<false>(h, m, n, A, Q, R, st);
  if (dtype == TNB_C128) return qr_core<true>(h, m, n, A, Q, R, st);
  return set_err(h, TNB_ERR_UNSUPPORTED, "qr: dtype %d", dtype);
}

}  // namespace tnb
