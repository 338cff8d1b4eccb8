// Bootstrap statistics of optimality-gap curves: replaces utils.bootstrap (Src/utils.py:39-115), called from
// results.py:172-184,787-802 on (trials x 102) curve matrices and on vectors of cumulative gaps.
//
// The reference draws, `trials` (1000) times, N rows with replacement (random.choices), averages them position by
// position with np.mean over a Python list (1-D pairwise summation), and finally takes np.mean / np.percentile(2.5) /
// np.percentile(97.5) over the replicates. Here the index stream stays on the host (CPython's generator decides
// which rows are drawn — parity of the replicates), and the B x N x C gathered sums, the per-column sort and the
// interpolated percentiles run on the GPU with numpy's exact summation order and interpolation formula, so the
// result is bit-identical to the reference's.
//
// Roofline: HBM/L2 gather — B*N*C*8 B of (L2-resident) reads, B*N*4 B of indices; ~1 add per 8 B.
#include "common.cuh"

namespace b200bo {

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum) over the sequence
// get(0..n-1): < 8 elements serial; <= 128 elements eight accumulators combined as a fixed tree plus a serial
// remainder; larger inputs split at n/2 rounded down to a multiple of 8.
template <typename Get>
__device__ __forceinline__ double np_pairwise_leaf(const Get& get, long long lo, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, get(lo + i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = get(lo + j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], get(lo + i + j));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, get(lo + i));
  return res;
}

template <typename Get>
__device__ double np_pairwise_sum(const Get& get, long long n) {
  constexpr int DEPTH = 40;
  long long lo_s[DEPTH], n_s[DEPTH];
  double left_s[DEPTH];
  int ph_s[DEPTH];
  int sp = 0;
  lo_s[0] = 0; n_s[0] = n; ph_s[0] = 0;
  double ret = 0.0;
  bool have = false;
  while (true) {
    if (!have) {
      if (n_s[sp] <= 128) {
        ret = np_pairwise_leaf(get, lo_s[sp], (int)n_s[sp]);
        have = true;
        --sp;
      } else {
        long long n2 = n_s[sp] / 2;
        n2 -= n2 % 8;
        ph_s[sp] = 0;
        lo_s[sp + 1] = lo_s[sp]; n_s[sp + 1] = n2; ph_s[sp + 1] = 0;
        ++sp;
      }
    } else {
      if (sp < 0) break;
      long long n2 = n_s[sp] / 2;
      n2 -= n2 % 8;
      if (ph_s[sp] == 0) {
        left_s[sp] = ret;
        ph_s[sp] = 1;
        have = false;
        lo_s[sp + 1] = lo_s[sp] + n2; n_s[sp + 1] = n_s[sp] - n2; ph_s[sp + 1] = 0;
        ++sp;
      } else {
        ret = __dadd_rn(left_s[sp], ret);
        --sp;
      }
    }
  }
  return ret;
}

// boot[b][c] = np.mean([data[idx[b][i]][c] for i in range(N)])
__global__ void __launch_bounds__(128)
bootstrap_means_kernel(const double* __restrict__ data, int N, int C, const int32_t* __restrict__ idx, int B,
                       double* __restrict__ boot) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * C) return;
  const int b = (int)(e / C), c = (int)(e - (long long)b * C);
  const int32_t* row = idx + (size_t)b * N;
  auto get = [&](long long i) { return data[(size_t)row[i] * C + c]; };
  boot[e] = np_pairwise_sum(get, N) / (double)N;
}

// one CTA per column: mean over the B replicates (pairwise, original order), then an in-shared-memory bitonic sort
// and numpy's "linear" percentiles (function_base.py: virtual index (n - 1) q, _get_gamma, _lerp).
__global__ void __launch_bounds__(256)
column_stats_kernel(const double* __restrict__ boot, int B, int C, int P2, const double* __restrict__ q, int nq,
                    double* __restrict__ out) {
  extern __shared__ double sv[];
  __shared__ int nan_seen;
  const int c = blockIdx.x;
  if (threadIdx.x == 0) nan_seen = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < P2; b += blockDim.x) {
    double v = b < B ? boot[(size_t)b * C + c] : INFINITY;
    if (v != v) nan_seen = 1;
    sv[b] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    auto get = [&](long long i) { return sv[i]; };
    out[c] = np_pairwise_sum(get, B) / (double)B;
  }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = sv[i], b = sv[l];
          const bool up = (i & k) == 0;
          if (up ? (a > b) : (a < b)) { sv[i] = b; sv[l] = a; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x < nq) {
    const double qq = q[threadIdx.x];                                   // q / 100, divided on the host as numpy does
    const double vi = __dmul_rn((double)(B - 1), qq);               // method "linear": (n - 1) * q
    double fl = floor(vi);
    const double gamma = __dadd_rn(vi, -fl);
    long long lo = (long long)fl;
    long long hi = lo + 1;
    lo = lo < 0 ? 0 : (lo > B - 1 ? B - 1 : lo);
    hi = hi < 0 ? 0 : (hi > B - 1 ? B - 1 : hi);
    const double a = sv[lo], b = sv[hi];
    const double diff = __dadd_rn(b, -a);
    double r = __dadd_rn(a, __dmul_rn(diff, gamma));
    if (gamma >= 0.5) r = __dadd_rn(b, -__dmul_rn(diff, __dadd_rn(1.0, -gamma)));
    if (nan_seen) r = nan("");
    out[(size_t)(1 + threadIdx.x) * C + c] = r;
  }
}

}  // namespace b200bo

extern "C" {

int b200bo_bootstrap_means(const double* data, int N, int C, const int32_t* idx, int B, double* boot, void* stream) {
  using namespace b200bo;
  if (N < 1 || C < 1 || B < 0) return arg_error("bootstrap_means: N=%d C=%d B=%d", N, C, B);
  if (B == 0) return 0;
  if (!data || !idx || !boot) return arg_error("bootstrap_means: null pointer");
  const long long total = (long long)B * C;
  const int threads = 128;
  const long long blocks = (total + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return arg_error("bootstrap_means: B*C too large");
  bootstrap_means_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(data, N, C, idx, B, boot);
  B200BO_CHECK_LAUNCH("bootstrap_means");
  return 0;
}

int b200bo_column_mean_percentiles(const double* boot, int B, int C, const double* q, int nq, double* out, void* stream) {
  using namespace b200bo;
  if (B < 1 || C < 1 || nq < 0 || nq > 256) return arg_error("column_mean_percentiles: B=%d C=%d nq=%d", B, C, nq);
  if (!boot || !out || (nq > 0 && !q)) return arg_error("column_mean_percentiles: null pointer");
  int P2 = 1;
  while (P2 < B) P2 <<= 1;
  const size_t smem = (size_t)P2 * sizeof(double);
  if (smem > 200 * 1024) return arg_error("column_mean_percentiles: B=%d exceeds the in-shared-memory sort (max 16384)", B);
  cudaError_t e = cudaFuncSetAttribute(column_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "column_mean_percentiles: smem attribute");
  column_stats_kernel<<<C, 256, smem, as_stream(stream)>>>(boot, B, C, P2, q, nq, out);
  B200BO_CHECK_LAUNCH("column_mean_percentiles");
  return 0;
}

}  // extern "C"
