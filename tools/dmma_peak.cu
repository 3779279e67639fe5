// Microbenchmark: issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) and of plain DFMA on
// one B200, to establish the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a0, double b0) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int bps = 1; bps <= 4; bps *= 2) {
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dmma_loop<<<sms * bps, 256>>>(out, iters, 1.0, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)sms * bps * 8 /*warps*/ * iters * 16.0 * 512.0;
    printf("{\"kernel\":\"dmma_8x8x4\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", bps, flops / ms * 1e-9, ms);
  }
  for (int bps = 2; bps <= 8; bps *= 2) {
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_loop<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)sms * bps * 256.0 * iters * 16.0 * 2.0;
    printf("{\"kernel\":\"dfma\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n", bps, flops / ms * 1e-9, ms);
  }
  printf("{\"sms\":%d,\"clock_khz\":%d}\n", sms, p.clockRate);
  return 0;
}
