// Library-level entry points of the C ABI: version, error string, device info.
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace b200bo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int arg_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

int cuda_status(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

}  // namespace b200bo

namespace b200bo {
// FP64 FMA issue-rate probe: 8 independent chains per thread, `iters` x 8 FMAs each. Gives the roofline
// denominator for the FP64-bound kernels on the very GPU a benchmark runs on (DMMA shares this pipe:
// tools/microbench/fp64_peaks.cu).
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double* out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Shared-memory read-bandwidth probe: 1024 threads per CTA, one CTA per SM, 8 independent conflict-free LDS.64 per
// thread and iteration out of a 32 KB table. The roofline denominator of the persistent BO loop (csrc/bo_persist.cu),
// whose hot loop is one 8-byte kernel-table look-up per (candidate, observation) pair.
__global__ void __launch_bounds__(1024) smem_lds_probe_kernel(double* out, int iters) {
  __shared__ double tab[4096];
  for (int e = threadIdx.x; e < 4096; e += 1024) tab[e] = e * 1e-6;
  __syncthreads();
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  for (int it = 0; it < iters; ++it) {
    const int base = (threadIdx.x + it * 32) & 511;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += tab[base + k * 512];
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace b200bo

extern "C" {

long long b200bo_smem_lds_probe(double* out, int blocks, int iters, void* stream) {
  if (!out || blocks < 1 || iters < 1) return b200bo::arg_error("smem_lds_probe: bad arguments");
  b200bo::smem_lds_probe_kernel<<<blocks, 1024, 0, b200bo::as_stream(stream)>>>(out, iters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(long long)b200bo::cuda_status(e, "smem_lds_probe");
  return (long long)blocks * 1024 * iters * 8 * 8;   // bytes read from shared memory
}

long long b200bo_fp64_fma_probe(double* out, int blocks, int iters, void* stream) {
  if (!out || blocks < 1 || iters < 1) return b200bo::arg_error("fp64_fma_probe: bad arguments");
  b200bo::fp64_fma_probe_kernel<<<blocks, 256, 0, b200bo::as_stream(stream)>>>(out, iters, 1.0000001, 1e-9);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(long long)b200bo::cuda_status(e, "fp64_fma_probe");
  return (long long)blocks * 256 * iters * 8 * 2;   // flops issued
}

const char* b200bo_version(void) { return "b200bo 0.1.0 (sm_100a)"; }

const char* b200bo_last_error_string(void) { return b200bo::g_err; }

int b200bo_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200bo::cuda_status(e, "cudaGetDevice");
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return b200bo::cuda_status(e, "cudaGetDeviceProperties");
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

}  // extern "C"
