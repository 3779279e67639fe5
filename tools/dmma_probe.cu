// Probe: DMMA.8x8x4 throughput vs warps per SM and operand variety (tuning aid, not shipped).
#include <cuda_runtime.h>
#include <stdio.h>
template <int NA, int NB>
__global__ void k(double* out, int iters, double a0) {
  double c[NA][NB][2];
  double a[NA], b[NB];
#pragma unroll
  for (int i = 0; i < NA; ++i) a[i] = a0 + i + threadIdx.x;
#pragma unroll
  for (int j = 0; j < NB; ++j) b[j] = a0 * 1e-9 + j;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NA, int NB>
void run(int sms, double* out, int warps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 4000;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<NA, NB><<<sms, warps * 32>>>(out, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double flops = (double)sms * warps * iters * NA * NB * 512.0;
  printf("{\"NA\":%d,\"NB\":%d,\"warps_per_sm\":%d,\"tflops\":%.2f}\n", NA, NB, warps, flops / ms * 1e-9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
  for (int w : {4, 8, 12, 16, 24, 32}) run<4, 4>(p.multiProcessorCount, out, w);
  for (int w : {4, 8, 12, 16}) run<4, 8>(p.multiProcessorCount, out, w);
  for (int w : {4, 8, 16}) run<2, 2>(p.multiProcessorCount, out, w);
  for (int w : {4, 8, 16}) run<1, 1>(p.multiProcessorCount, out, w);
  return 0;
}
