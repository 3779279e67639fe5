// Real / complex-FP64 scalar helpers shared by the factorization kernels (device side).
#pragma once
#include <cuda_runtime.h>

namespace tnb {

template <bool CPLX> struct ElemT { using T = double; };
template <> struct ElemT<true> { using T = double2; };

template <typename T> __host__ __device__ __forceinline__ T a_zero();
template <> __host__ __device__ __forceinline__ double a_zero<double>() { return 0.0; }
template <> __host__ __device__ __forceinline__ double2 a_zero<double2>() { return make_double2(0.0, 0.0); }
template <typename T> __host__ __device__ __forceinline__ T a_one();
template <> __host__ __device__ __forceinline__ double a_one<double>() { return 1.0; }
template <> __host__ __device__ __forceinline__ double2 a_one<double2>() { return make_double2(1.0, 0.0); }
template <typename T> __host__ __device__ __forceinline__ T a_real(double r);
template <> __host__ __device__ __forceinline__ double a_real<double>(double r) { return r; }
template <> __host__ __device__ __forceinline__ double2 a_real<double2>(double r) { return make_double2(r, 0.0); }

__host__ __device__ __forceinline__ double a_add(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ double2 a_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double a_sub(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ double2 a_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double a_mul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ double2 a_mul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// conj(a) * b
__host__ __device__ __forceinline__ double a_cmul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ double2 a_cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ double a_scale(double a, double s) { return a * s; }
__host__ __device__ __forceinline__ double2 a_scale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ double a_conj(double a) { return a; }
__host__ __device__ __forceinline__ double2 a_conj(double2 a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ double a_neg(double a) { return -a; }
__host__ __device__ __forceinline__ double2 a_neg(double2 a) { return make_double2(-a.x, -a.y); }
__host__ __device__ __forceinline__ double a_abs2(double a) { return a * a; }
__host__ __device__ __forceinline__ double a_abs2(double2 a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ double a_re(double a) { return a; }
__host__ __device__ __forceinline__ double a_re(double2 a) { return a.x; }
__host__ __device__ __forceinline__ double a_im(double) { return 0.0; }
__host__ __device__ __forceinline__ double a_im(double2 a) { return a.y; }

__device__ __forceinline__ double a_shfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 a_shfl_xor(double2 v, int o) {
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
template <typename T> __device__ __forceinline__ T a_warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = a_add(v, a_shfl_xor(v, o));
  return v;
}

}  // namespace tnb
