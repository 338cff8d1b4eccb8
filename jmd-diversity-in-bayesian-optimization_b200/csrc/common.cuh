// Shared device/host helpers for the b200bo kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/b200bo.h"

namespace b200bo {

void set_error(const char* fmt, ...);
int arg_error(const char* fmt, ...);     // records message, returns -1
int cuda_status(cudaError_t e, const char* where);   // 0 or positive cudaError_t (records message)

// per-trial reduction of the per-CTA (value, index) partials of the arg-max kernels (grid_acq.cu)
void launch_argmax_finish(const double* part_val, const int32_t* part_idx, int chunks, int T, const int32_t* n,
                          int32_t* arg_idx, double* arg_val, int32_t* status, cudaStream_t st);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define B200BO_CHECK_LAUNCH(where) \
  do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return ::b200bo::cuda_status(e__, where); } while (0)

constexpr double kInvSqrt2Pi = 0.3989422804014327;
constexpr double kInvSqrt2 = 0.7071067811865476;
constexpr double kLog2Pi = 1.8378770664093453;

// Kernel parameters precomputed on the host side of a launch (or per trial on device).
struct KernParams {
  double s;        // outputscale
  double a_scale;  // RBF: -0.5/l^2 ; Matern-1/2: 1/l ; Matern-3/2: sqrt(3)/l ; Matern-5/2: sqrt(5)/l
  double inv_ell;  // 1/l
};

__host__ __device__ inline KernParams make_kern_params(double s, double ell, int kernel_id) {
  KernParams p;
  p.s = s;
  p.inv_ell = 1.0 / ell;
  if (kernel_id == B200BO_KERNEL_RBF) p.a_scale = -0.5 / (ell * ell);
  else if (kernel_id == B200BO_KERNEL_MATERN12) p.a_scale = 1.0;   // uses r / ell directly
  else if (kernel_id == B200BO_KERNEL_MATERN32) p.a_scale = sqrt(3.0) / ell;
  else p.a_scale = sqrt(5.0) / ell;
  return p;
}

// s * k(r / ell) from the squared distance; operation order == oracle/gp.py:kernel_from_r2.
template <int KID>
__device__ __forceinline__ double kernel_from_r2(double r2, double s, double ell, double a_scale) {
  if (KID == B200BO_KERNEL_RBF) {
    return s * exp(r2 * a_scale);
  } else if (KID == B200BO_KERNEL_MATERN12) {
    return s * exp(-(sqrt(r2) / ell));
  } else if (KID == B200BO_KERNEL_MATERN32) {
    double a = a_scale * sqrt(r2);
    return s * ((1.0 + a) * exp(-a));
  } else {
    double a = a_scale * sqrt(r2);
    return s * ((1.0 + a + (a * a) / 3.0) * exp(-a));
  }
}

__device__ __forceinline__ double kernel_from_r2_dyn(int kid, double r2, double s, double ell, double a_scale) {
  switch (kid) {
    case B200BO_KERNEL_RBF: return kernel_from_r2<B200BO_KERNEL_RBF>(r2, s, ell, a_scale);
    case B200BO_KERNEL_MATERN12: return kernel_from_r2<B200BO_KERNEL_MATERN12>(r2, s, ell, a_scale);
    case B200BO_KERNEL_MATERN32: return kernel_from_r2<B200BO_KERNEL_MATERN32>(r2, s, ell, a_scale);
    default: return kernel_from_r2<B200BO_KERNEL_MATERN52>(r2, s, ell, a_scale);
  }
}

// s * d k(r/ell) / d ell ; oracle/gp.py:dkernel_dell_from_r2
__device__ __forceinline__ double dkernel_dell_from_r2(int kid, double r2, double s, double ell, double a_scale) {
  double r = sqrt(r2);
  if (kid == B200BO_KERNEL_RBF) {
    double q = r2 / (ell * ell);
    return s * q * exp(-0.5 * q) / ell;
  } else if (kid == B200BO_KERNEL_MATERN12) {
    double d = r / ell;
    return s * d * exp(-d) / ell;
  } else if (kid == B200BO_KERNEL_MATERN32) {
    double a = a_scale * r;
    return s * (a * a) * exp(-a) / ell;
  } else {
    double a = a_scale * r;
    return s * ((a * a) / 3.0) * (1.0 + a) * exp(-a) / ell;
  }
}

__device__ __forceinline__ double norm_pdf(double u) { return exp(-0.5 * (u * u)) * kInvSqrt2Pi; }
__device__ __forceinline__ double norm_cdf(double u) { return 0.5 * erfc(-u * kInvSqrt2); }

// acquisition value; oracle/gp.py:acquisition
template <int ACQ>
__device__ __forceinline__ double acquisition(double mu, double var, double best_f, double acq_param, double var_floor) {
  double sigma = sqrt(fmax(var, var_floor));
  if (ACQ == B200BO_ACQ_UCB) return mu + acq_param * sigma;   // acq_param = sqrt(beta), precomputed
  double u = (mu - best_f) / sigma;
  if (ACQ == B200BO_ACQ_PI) return norm_cdf(u);
  return sigma * (norm_pdf(u) + u * norm_cdf(u));
}

// arg-max candidate ordering: larger value wins; equal values -> lower index wins;
// NaN never wins (callers replace NaN by -inf and idx by INT_MAX beforehand).
__device__ __forceinline__ bool better(double v, int i, double bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- 1-D bulk copy global -> shared through the TMA unit (cp.async.bulk + mbarrier transaction count) -------------
// One elected thread arms the barrier with the byte count and issues the copy; every thread waits on the barrier's
// phase before reading the staged bytes. dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE_%=;\n"
      "bra MBAR_WAIT_%=;\n"
      "MBAR_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace b200bo
