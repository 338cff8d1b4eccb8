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

extern "C" {

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
