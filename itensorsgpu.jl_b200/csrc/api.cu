// Handle, error reporting, workspace arena and the extern "C" entry points of tier 1.
#include <stdarg.h>

#include "tnb_internal.h"

namespace tnb {

// plan cache of the contraction engine (contract.cu)
void plan_cache_drop(Handle* h);
void plan_cache_clear(Handle* h);
void plan_cache_stats(Handle* h, uint64_t* entries, uint64_t* hits, uint64_t* misses, uint64_t* tuned);
void plan_cache_set_autotune(Handle* h, int mode);
void kernel_family_counts(uint64_t out[4]);

int set_err(Handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

int check_cuda(Handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TNB_OK;
  return set_err(h, TNB_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

int ws_require(Handle* h, size_t bytes) {
  bytes += h->ws_base;
  if (bytes <= h->ws_bytes) return TNB_OK;
  if (h->ws_off != 0 || h->ws_base != 0)
    return set_err(h, TNB_ERR_ALLOC, "internal: workspace growth requested while in use (%zu > %zu)", bytes,
                   h->ws_bytes);
  TNB_CUDA(h, cudaDeviceSynchronize());
  if (h->ws) cudaFree(h->ws);
  h->ws = nullptr;
  h->ws_bytes = 0;
  size_t want = bytes + (bytes >> 3) + (1u << 20);
  cudaError_t e = cudaMalloc((void**)&h->ws, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc((void**)&h->ws, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err(h, TNB_ERR_ALLOC, "workspace allocation of %zu bytes failed", bytes);
  }
  h->ws_bytes = want;
  return TNB_OK;
}

int ws_alloc(Handle* h, size_t bytes, void** out) {
  size_t off = (h->ws_off + 255) & ~(size_t)255;
  if (off + bytes > h->ws_bytes) {
    if (h->ws_off != 0 || h->ws_base != 0)
      return set_err(h, TNB_ERR_ALLOC, "internal: workspace exhausted (%zu + %zu > %zu); ws_require missing",
                     off, bytes, h->ws_bytes);
    TNB_TRY(ws_require(h, off + bytes));
  }
  *out = h->ws + off;
  h->ws_off = off + bytes;
  return TNB_OK;
}

}  // namespace tnb

using namespace tnb;

#define H ((Handle*)h)
#define ST ((cudaStream_t)stream)

extern "C" {

int tnb_version(void) { return 100; }

int tnb_create(tnb_handle_t* out) {
  if (!out) return TNB_ERR_BAD_ARG;
  Handle* hd = new Handle();
  cudaError_t e = cudaGetDevice(&hd->device);
  if (e != cudaSuccess) { delete hd; return TNB_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, hd->device);
  if (e != cudaSuccess) { delete hd; return TNB_ERR_CUDA; }
  if (prop.major != 10) {
    // sm_100a-only binary: refuse anything else loudly instead of failing at first launch
    delete hd;
    return TNB_ERR_UNSUPPORTED;
  }
  hd->num_sms = prop.multiProcessorCount;
  if (cudaMalloc((void**)&hd->scal, 256 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost((void**)&hd->scal_host, 256 * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&hd->partials, 8 * RED_MAX_BLOCKS * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&hd->counter, 64) != cudaSuccess ||
      cudaMalloc(&hd->what, 64 * 1024) != cudaSuccess ||
      cudaStreamCreateWithFlags(&hd->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    if (hd->scal) cudaFree(hd->scal);
    if (hd->scal_host) cudaFreeHost(hd->scal_host);
    if (hd->partials) cudaFree(hd->partials);
    if (hd->counter) cudaFree(hd->counter);
    if (hd->what) cudaFree(hd->what);
    delete hd;
    return TNB_ERR_ALLOC;
  }
  cudaMemset(hd->scal, 0, 256 * sizeof(double));
  cudaMemset(hd->counter, 0, 64);
  *out = (tnb_handle_t)hd;
  return TNB_OK;
}

int tnb_destroy(tnb_handle_t h) {
  if (!h) return TNB_ERR_BAD_ARG;
  cudaDeviceSynchronize();
  if (H->ws) cudaFree(H->ws);
  if (H->scal) cudaFree(H->scal);
  if (H->partials) cudaFree(H->partials);
  if (H->counter) cudaFree(H->counter);
  if (H->what) cudaFree(H->what);
  if (H->scal_host) cudaFreeHost(H->scal_host);
  for (auto& e : H->ev) if (e) cudaEventDestroy(e);
  if (H->copy_stream) cudaStreamDestroy(H->copy_stream);
  plan_cache_drop(H);
  delete H;
  return TNB_OK;
}

const char* tnb_last_error(tnb_handle_t h) { return h ? H->err.c_str() : "null handle"; }

int tnb_reserve(tnb_handle_t h, size_t bytes) {
  if (!h) return TNB_ERR_BAD_ARG;
  ws_reset(H);
  return ws_require(H, bytes);
}

size_t tnb_workspace_bytes(tnb_handle_t h) { return h ? H->ws_bytes : 0; }
int tnb_plan_cache_stats(tnb_handle_t h, uint64_t* entries, uint64_t* hits, uint64_t* misses, uint64_t* autotuned) {
  if (!h) return TNB_ERR_BAD_ARG;
  plan_cache_stats(H, entries, hits, misses, autotuned);
  return TNB_OK;
}

int tnb_kernel_family_counts(tnb_handle_t h, uint64_t* out4) {
  if (!h || !out4) return TNB_ERR_BAD_ARG;
  kernel_family_counts(out4);
  return TNB_OK;
}

int tnb_plan_cache_clear(tnb_handle_t h) {
  if (!h) return TNB_ERR_BAD_ARG;
  plan_cache_clear(H);
  return TNB_OK;
}

int tnb_set_autotune(tnb_handle_t h, int mode) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return set_err(H, TNB_ERR_BAD_ARG, "set_autotune: mode %d", mode);
  plan_cache_set_autotune(H, mode);
  return TNB_OK;
}

uint64_t tnb_launch_count(tnb_handle_t h) { return h ? H->launches : 0; }

int tnb_contract(tnb_handle_t h, int dtype, int nA, const int64_t* extA, const int32_t* modeA,
                 const void* A, int nB, const int64_t* extB, const int32_t* modeB, const void* B,
                 int nC, const int64_t* extC, const int32_t* modeC, void* C, const void* alpha,
                 const void* beta, int flags, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return contract_impl(H, dtype, nA, extA, modeA, A, nB, extB, modeB, B, nC, extC, modeC, C, alpha, beta,
                       flags, ST);
}

int tnb_permute_axpby(tnb_handle_t h, int dtype, int n, const int64_t* extA, const int32_t* modeA,
                      const void* A, const int32_t* modeB, void* B, const void* alpha,
                      const void* beta, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return permute_axpby_impl(H, dtype, n, extA, modeA, A, modeB, B, alpha, beta, ST);
}

int tnb_diag_contract(tnb_handle_t h, int dtype, int n, const int64_t* extA, const int32_t* modeA, const void* A,
                      int32_t scaled_mode, const void* diag, int diag_dtype, const int32_t* modeC, void* C, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  if (!extA || !modeA || !modeC) return set_err(H, TNB_ERR_BAD_ARG, "diag_contract: null pointer");
  return diag_contract_impl(H, dtype, n, extA, modeA, A, scaled_mode, diag, diag_dtype, modeC, C, (cudaStream_t)stream);
}

int tnb_scale(tnb_handle_t h, int dtype, int64_t n, void* x, const void* alpha, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return scale_impl(H, dtype, n, x, alpha, ST);
}

int tnb_dot(tnb_handle_t h, int dtype, int64_t n, const void* x, const void* y, void* result_dev,
            void* result_host, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  void* r = result_dev ? result_dev : (void*)H->scal;
  TNB_TRY(dot_impl(H, dtype, n, x, y, r, ST));
  if (result_host) {
    TNB_CUDA(H, cudaMemcpyAsync(result_host, r, elsize(dtype), cudaMemcpyDeviceToHost, ST));
    TNB_CUDA(H, cudaStreamSynchronize(ST));
  }
  return TNB_OK;
}

int tnb_nrm2(tnb_handle_t h, int dtype, int64_t n, const void* x, double* result_dev,
             double* result_host, void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  double* r = result_dev ? result_dev : H->scal;
  TNB_TRY(nrm2_impl(H, dtype, n, x, r, ST));
  if (result_host) {
    TNB_CUDA(H, cudaMemcpyAsync(result_host, r, sizeof(double), cudaMemcpyDeviceToHost, ST));
    TNB_CUDA(H, cudaStreamSynchronize(ST));
  }
  return TNB_OK;
}

int tnb_truncate(tnb_handle_t h, const double* P_dev, int64_t len, int64_t maxdim, int64_t mindim,
                 double cutoff, int flags, int64_t* n_keep, double* truncerr, double* docut,
                 void* stream) {
  if (!h) return TNB_ERR_BAD_ARG;
  return truncate_impl(H, P_dev, len, maxdim, mindim, cutoff, flags, n_keep, truncerr, docut, ST);
}

}  // extern "C"
