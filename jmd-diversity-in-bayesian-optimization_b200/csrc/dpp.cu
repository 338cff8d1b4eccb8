// K-D: ranked-DPP set scoring — log det exp(-gamma * D) for N index sets.
//
// Replaces diverse_data_generator.twoD_score (Src/data_gen.py:484-497) looped over the
// superset (data_gen.py:403) and the mean / np.std / z-score block (data_gen.py:406-408).
//
// Roofline: FP64-pipe bound. Per set (k = 10): 45 pair kernels (sub, sub, mul, fma, sqrt,
// mul, mul, exp) + a 10x10 Cholesky (220 FMA, 10 sqrt, 45 div-equivalents) + 10 log;
// HBM traffic is 4k bytes in + 8 bytes out = 48 B per set, pool (160 KB) stays in L1/L2.
//
// k <= KREG_MAX: one set per thread, packed lower triangle held in registers (fully
// unrolled, no pivoting: S is symmetric positive definite in exact arithmetic).
// larger k: one set per CTA, matrix in shared memory.
#include "common.cuh"

namespace b200bo {

constexpr int KREG_MAX = 12;

// D = (sqrt(dx^2 + dy^2))^2 : scipy pdist('euclidean') then **2 (data_gen.py:489).
__device__ __forceinline__ double dpp_entry(double ax, double ay, double bx, double by, double neg_gamma) {
  double dx = ax - bx, dy = ay - by;
  double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // unfused, as numpy sums squares
  return exp(neg_gamma * (d * d));
}

// log |det| of the packed symmetric matrix by an LDL^T elimination in natural order: sum_j log |d_j|.
// The reference takes np.linalg.slogdet(S)[1] (data_gen.py:497), i.e. log|det| from an LU factorisation, which stays
// finite (and informative: the leading pivots dominate) on the nearly singular S of very small gamma where a Cholesky
// would stop at the first non-positive pivot — gammacalc (data_gen.py:758-842) ranks exactly those cases. For a
// positive definite S the pivots are the Cholesky pivots L_jj^2. A zero pivot gives -inf, as slogdet does.
// sum_j log |d_j| with ONE log per set where the pivots allow it: their product is formed directly (an identity-like S
// gives exactly 1 and log 1 = 0, as the sum of logs does), and only a product that left the comfortable range of a
// double (|det| < 1e-250: nearly singular S at very small gamma) is redone as the sum of the logs. The ten logs were
// 59 % of the FP64-pipe work of the lattice variant.
template <int K>
__device__ __forceinline__ double log_abs_prod(const double (&d)[K]) {
  double p = 1.0;
#pragma unroll
  for (int j = 0; j < K; ++j) p *= d[j];
  p = fabs(p);
  if (p > 1e-250 && p < 1e250) return log(p);
  double acc = 0.0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) acc += log(fabs(d[j]));
  return acc;
}

template <int K>
__device__ __forceinline__ double ldl_logabsdet(double (&A)[K * (K + 1) / 2]) {
  double piv[K];
  bool singular = false;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double d = A[j * (j + 1) / 2 + j];
    singular = singular || d == 0.0;
    piv[j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int c = j + 1; c < K; ++c) {
      const double lc = A[c * (c + 1) / 2 + j] * inv;      // L_cj
#pragma unroll
      for (int r = c; r < K; ++r) A[r * (r + 1) / 2 + c] = fma(-A[r * (r + 1) / 2 + j], lc, A[r * (r + 1) / 2 + c]);
    }
  }
  return singular ? -INFINITY : log_abs_prod<K>(piv);
}

template <int K>
__global__ void __launch_bounds__(128)
dpp_logdet_reg_kernel(const double2* __restrict__ pool, const int32_t* __restrict__ idx, long long N, double neg_gamma,
                      double* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  double px[K], py[K];
  const int32_t* row = idx + i * K;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    double2 p = __ldg(&pool[row[a]]);
    px[a] = p.x;
    py[a] = p.y;
  }
  // packed lower triangle, A[r][c] (c <= r) at r*(r+1)/2 + c
  double A[K * (K + 1) / 2];
#pragma unroll
  for (int r = 0; r < K; ++r) {
#pragma unroll
    for (int c = 0; c < r; ++c) A[r * (r + 1) / 2 + c] = dpp_entry(px[r], py[r], px[c], py[c], neg_gamma);
    A[r * (r + 1) / 2 + r] = 1.0;   // exp(-gamma * 0)
  }
  out[i] = ldl_logabsdet<K>(A);
}

// ---- lattice variant ------------------------------------------------------------------------------------
// The reference's pool is always the constructX lattice (data_gen.py:294-303): an entry of the L-ensemble depends
// on (|dx|, |dy|) only, so it is looked up in a (Gy, Gx) table (80 KB for 100 x 100: L1-resident) built once per
// gamma by the same dpp_entry() code — bit-identical values, ~75 FP64-pipe instructions (sqrt + exp) replaced by
// one gather per pair. What is left per set is the 10 x 10 Cholesky + 10 logs.
__global__ void __launch_bounds__(256)
dpp_table_kernel(int Gx, int Gy, double h, double neg_gamma, double* __restrict__ table) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Gx * Gy) return;
  const int dy = e / Gx, dx = e - dy * Gx;
  table[e] = (dx | dy) ? dpp_entry(h * dx, h * dy, 0.0, 0.0, neg_gamma) : 1.0;
}

template <int K>
__global__ void __launch_bounds__(128)
dpp_logdet_lattice_kernel(const double* __restrict__ table, int Gx, int G, const int32_t* __restrict__ idx, long long N,
                          double* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int ix[K], iy[K];
  const int32_t* row = idx + i * K;
  bool ok = true;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const int g = row[a];
    ok = ok && g >= 0 && g < G;
    const int gg = ok ? g : 0;
    iy[a] = gg / Gx;
    ix[a] = gg - iy[a] * Gx;
  }
  double A[K * (K + 1) / 2];
#pragma unroll
  for (int r = 0; r < K; ++r) {
#pragma unroll
    for (int c = 0; c < r; ++c) A[r * (r + 1) / 2 + c] = __ldg(&table[abs(iy[r] - iy[c]) * Gx + abs(ix[r] - ix[c])]);
    A[r * (r + 1) / 2 + r] = 1.0;
  }
  const double acc = ldl_logabsdet<K>(A);
  out[i] = ok ? acc : nan("");
}

template <int K>
static void launch_lattice(const double* table, int Gx, int G, const int32_t* idx, long long N, double* out, cudaStream_t st) {
  const int threads = 128;
  const long long blocks = (N + threads - 1) / threads;
  dpp_logdet_lattice_kernel<K><<<(unsigned)blocks, threads, 0, st>>>(table, Gx, G, idx, N, out);
}

// generic k: one set per CTA, full matrix in shared memory (k*k doubles).
__global__ void __launch_bounds__(128)
dpp_logdet_smem_kernel(const double2* __restrict__ pool, const int32_t* __restrict__ idx, long long N, int k,
                       double neg_gamma, double* __restrict__ out) {
  extern __shared__ double sm[];
  double* A = sm;                 // k*k, row-major, lower part used
  double* px = sm + (size_t)k * k;
  double* py = px + k;
  __shared__ double s_acc;
  __shared__ int s_sing;
  for (long long set = blockIdx.x; set < N; set += gridDim.x) {
    const int32_t* row = idx + set * k;
    for (int a = threadIdx.x; a < k; a += blockDim.x) {
      double2 p = __ldg(&pool[row[a]]);
      px[a] = p.x;
      py[a] = p.y;
    }
    if (threadIdx.x == 0) { s_acc = 0.0; s_sing = 0; }
    __syncthreads();
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
      int r = e / k, c = e - r * k;
      if (c <= r) A[e] = (c == r) ? 1.0 : dpp_entry(px[r], py[r], px[c], py[c], neg_gamma);
    }
    __syncthreads();
    for (int j = 0; j < k; ++j) {
      double d = A[j * k + j];
      __syncthreads();   // everyone has read the pivot before column j is scaled
      if (threadIdx.x == 0) {                                // log|det| as slogdet: see ldl_logabsdet
        s_acc += log(fabs(d));
        if (d == 0.0) s_sing = 1;
      }
      const double inv = 1.0 / d;
      // trailing update of the lower triangle: rows r > j, cols j < c <= r  (A_rc -= A_rj * (A_cj / d))
      int m = k - j - 1;
      for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        int rr = e / m, cc = e - rr * m;
        if (cc <= rr) {
          int r = j + 1 + rr, c = j + 1 + cc;
          A[r * k + c] = fma(-A[r * k + j], A[c * k + j] * inv, A[r * k + c]);
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) out[set] = s_sing ? -INFINITY : s_acc;
    __syncthreads();
  }
}

// ---- statistics: mean, population sd (np.std, ddof 0), z = (x - mean) / sd -----------
// two-pass (mean, then centred sum of squares) like numpy; deterministic tree order:
// per-block partials in ws, final reduction by one block.
__global__ void __launch_bounds__(256) reduce_sum_kernel(const double* __restrict__ x, long long N, double shift,
                                                         int square, double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i] - shift;
    acc += square ? v * v : v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

// stage 0: stats[0] = sum(partial)/N ; stage 1: stats[1] = sqrt(sum(partial)/N)
__global__ void __launch_bounds__(256) finish_stat_kernel(const double* __restrict__ partial, int nblocks, long long N,
                                                          int stage, double* __restrict__ stats) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) acc += partial[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    if (stage == 0) stats[0] = t / (double)N;
    else stats[1] = sqrt(t / (double)N);
  }
}

__global__ void __launch_bounds__(256) centred_sq_kernel(const double* __restrict__ x, long long N,
                                                         const double* __restrict__ stats, double* __restrict__ partial) {
  __shared__ double sh[8];
  double mean = stats[0];
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    double v = x[i] - mean;
    acc = fma(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) zscore_kernel(const double* __restrict__ x, long long N,
                                                     const double* __restrict__ stats, double* __restrict__ z) {
  double mean = stats[0], sd = stats[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    z[i] = (x[i] - mean) / sd;
}

constexpr int STAT_BLOCKS = 592;   // 4 x 148 SMs

template <int K>
static void launch_reg(const double* pool, const int32_t* idx, long long N, double gamma, double* out, cudaStream_t st) {
  int threads = 128;
  long long blocks = (N + threads - 1) / threads;
  dpp_logdet_reg_kernel<K><<<(unsigned)blocks, threads, 0, st>>>(reinterpret_cast<const double2*>(pool), idx, N, -gamma, out);
}

}  // namespace b200bo

extern "C" {

int b200bo_dpp_table(int Gx, int Gy, double h, double gamma, double* table, void* stream) {
  using namespace b200bo;
  if (Gx < 1 || Gy < 1 || !(h > 0.0)) return arg_error("dpp_table: Gx=%d Gy=%d h=%g", Gx, Gy, h);
  if (!table) return arg_error("dpp_table: null pointer");
  const int total = Gx * Gy;
  dpp_table_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(Gx, Gy, h, -gamma, table);
  B200BO_CHECK_LAUNCH("dpp_table");
  return 0;
}

int b200bo_dpp_logdet_score_lattice(const double* table, int Gx, int Gy, const int32_t* idx, long long N, int k,
                                    double* logdet, void* stream) {
  using namespace b200bo;
  if (Gx < 1 || Gy < 1 || k < 1 || k > KREG_MAX || N < 0)
    return arg_error("dpp_logdet_score_lattice: Gx=%d Gy=%d k=%d (max %d) N=%lld", Gx, Gy, k, KREG_MAX, N);
  if (N == 0) return 0;
  if (!table || !idx || !logdet) return arg_error("dpp_logdet_score_lattice: null pointer");
  if (N > 2147483647LL * 128) return arg_error("dpp_logdet_score_lattice: N too large");
  cudaStream_t st = as_stream(stream);
  const int G = Gx * Gy;
  switch (k) {
    case 1: launch_lattice<1>(table, Gx, G, idx, N, logdet, st); break;
    case 2: launch_lattice<2>(table, Gx, G, idx, N, logdet, st); break;
    case 3: launch_lattice<3>(table, Gx, G, idx, N, logdet, st); break;
    case 4: launch_lattice<4>(table, Gx, G, idx, N, logdet, st); break;
    case 5: launch_lattice<5>(table, Gx, G, idx, N, logdet, st); break;
    case 6: launch_lattice<6>(table, Gx, G, idx, N, logdet, st); break;
    case 7: launch_lattice<7>(table, Gx, G, idx, N, logdet, st); break;
    case 8: launch_lattice<8>(table, Gx, G, idx, N, logdet, st); break;
    case 9: launch_lattice<9>(table, Gx, G, idx, N, logdet, st); break;
    case 10: launch_lattice<10>(table, Gx, G, idx, N, logdet, st); break;
    case 11: launch_lattice<11>(table, Gx, G, idx, N, logdet, st); break;
    default: launch_lattice<12>(table, Gx, G, idx, N, logdet, st); break;
  }
  B200BO_CHECK_LAUNCH("dpp_logdet_score_lattice");
  return 0;
}

int b200bo_dpp_logdet_score(const double* pool, int pool_size, const int32_t* idx, long long N, int k, double gamma,
                            double* logdet, void* stream) {
  using namespace b200bo;
  if (pool_size <= 0 || k < 1 || N < 0) return arg_error("dpp_logdet_score: bad sizes pool=%d k=%d N=%lld", pool_size, k, N);
  if (k > 160) return arg_error("dpp_logdet_score: k=%d > 160 unsupported", k);
  if (N == 0) return 0;
  if (!pool || !idx || !logdet) return arg_error("dpp_logdet_score: null pointer");
  if (N > 2147483647LL * 128) return arg_error("dpp_logdet_score: N too large");
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: launch_reg<1>(pool, idx, N, gamma, logdet, st); break;
    case 2: launch_reg<2>(pool, idx, N, gamma, logdet, st); break;
    case 3: launch_reg<3>(pool, idx, N, gamma, logdet, st); break;
    case 4: launch_reg<4>(pool, idx, N, gamma, logdet, st); break;
    case 5: launch_reg<5>(pool, idx, N, gamma, logdet, st); break;
    case 6: launch_reg<6>(pool, idx, N, gamma, logdet, st); break;
    case 7: launch_reg<7>(pool, idx, N, gamma, logdet, st); break;
    case 8: launch_reg<8>(pool, idx, N, gamma, logdet, st); break;
    case 9: launch_reg<9>(pool, idx, N, gamma, logdet, st); break;
    case 10: launch_reg<10>(pool, idx, N, gamma, logdet, st); break;
    case 11: launch_reg<11>(pool, idx, N, gamma, logdet, st); break;
    case 12: launch_reg<12>(pool, idx, N, gamma, logdet, st); break;
    default: {
      size_t smem = ((size_t)k * k + 2 * (size_t)k) * sizeof(double);
      cudaError_t e = cudaFuncSetAttribute(dpp_logdet_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e, "dpp_logdet_score: smem attribute");
      long long blocks = N < 148 * 8 ? N : 148 * 8;
      dpp_logdet_smem_kernel<<<(unsigned)blocks, 128, smem, st>>>(reinterpret_cast<const double2*>(pool), idx, N, k, -gamma, logdet);
    }
  }
  B200BO_CHECK_LAUNCH("dpp_logdet_score");
  return 0;
}

size_t b200bo_dpp_stats_workspace_bytes(long long N) { (void)N; return b200bo::STAT_BLOCKS * sizeof(double); }

int b200bo_dpp_zscore(const double* logdet, long long N, double* stats, double* z, void* ws, size_t ws_bytes, void* stream) {
  using namespace b200bo;
  if (!logdet || !stats || !ws) return arg_error("dpp_zscore: null pointer");
  if (N <= 0) return arg_error("dpp_zscore: N=%lld", N);
  if (ws_bytes < b200bo_dpp_stats_workspace_bytes(N)) return arg_error("dpp_zscore: workspace too small");
  cudaStream_t st = as_stream(stream);
  double* partial = reinterpret_cast<double*>(ws);
  reduce_sum_kernel<<<STAT_BLOCKS, 256, 0, st>>>(logdet, N, 0.0, 0, partial);
  finish_stat_kernel<<<1, 256, 0, st>>>(partial, STAT_BLOCKS, N, 0, stats);
  centred_sq_kernel<<<STAT_BLOCKS, 256, 0, st>>>(logdet, N, stats, partial);
  finish_stat_kernel<<<1, 256, 0, st>>>(partial, STAT_BLOCKS, N, 1, stats);
  if (z) zscore_kernel<<<STAT_BLOCKS, 256, 0, st>>>(logdet, N, stats, z);
  B200BO_CHECK_LAUNCH("dpp_zscore");
  return 0;
}

}  // extern "C"

// ---- combinadic un-ranking on device ------------------------------
// Replaces combination_finder.find_elm (Src/Combination.py:131-170), 93 % of the reference's
// generate() wall time (pure-Python big integers). 128-bit unsigned arithmetic covers
// C(10000,10) ~ 2.7e33 < 2^112; the binomials C(v,b), v <= n, b <= k come from a host-built
// table (b-major), the per-position search is a binary search for the largest v < a with
// C(v,b) <= x, exactly the quantity LargestV / LargestV_binary return.
namespace b200bo {

typedef unsigned __int128 u128;

__global__ void __launch_bounds__(128)
unrank_u128_kernel(const unsigned long long* __restrict__ binom_lo, const unsigned long long* __restrict__ binom_hi,
                   const unsigned long long* __restrict__ m_lo, const unsigned long long* __restrict__ m_hi,
                   unsigned long long tot_lo, unsigned long long tot_hi, int n, int k, long long N,
                   int32_t* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const u128 total = ((u128)tot_hi << 64) | tot_lo;       // C(n,k)
  const u128 m = ((u128)m_hi[i] << 64) | m_lo[i];
  if (m >= total) {                                        // m == C(n,k): the reference raises (Combination.py:150-153)
    for (int p = 0; p < k; ++p) out[i * k + p] = -1;
    return;
  }
  u128 x = (total - 1) - m;
  int a = n;
  for (int b = k; b >= 1; --b) {
    const unsigned long long* lo = binom_lo + (size_t)(b - 1) * (n + 1);
    const unsigned long long* hi = binom_hi + (size_t)(b - 1) * (n + 1);
    int l = b - 1, h = a - 1;                              // C(b-1,b) = 0 <= x
    while (l < h) {
      int mid = (l + h + 1) >> 1;
      u128 c = ((u128)__ldg(&hi[mid]) << 64) | __ldg(&lo[mid]);
      if (c <= x) l = mid; else h = mid - 1;
    }
    x -= ((u128)__ldg(&hi[l]) << 64) | __ldg(&lo[l]);
    out[i * k + (k - b)] = (n - 1) - l;
    a = l;
  }
}

}  // namespace b200bo

extern "C" int b200bo_unrank_u128(const uint64_t* binom_lo, const uint64_t* binom_hi, const uint64_t* m_lo,
                                  const uint64_t* m_hi, uint64_t total_lo, uint64_t total_hi, int n, int k, long long N,
                                  int32_t* out, void* stream) {
  using namespace b200bo;
  if (N < 0 || n < 1 || k < 1 || k > n) return arg_error("unrank_u128: bad sizes n=%d k=%d N=%lld", n, k, N);
  if (N == 0) return 0;
  if (!binom_lo || !binom_hi || !m_lo || !m_hi || !out) return arg_error("unrank_u128: null pointer");
  unrank_u128_kernel<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(binom_lo), reinterpret_cast<const unsigned long long*>(binom_hi),
      reinterpret_cast<const unsigned long long*>(m_lo), reinterpret_cast<const unsigned long long*>(m_hi),
      total_lo, total_hi, n, k, N, out);
  B200BO_CHECK_LAUNCH("unrank_u128");
  return 0;
}
