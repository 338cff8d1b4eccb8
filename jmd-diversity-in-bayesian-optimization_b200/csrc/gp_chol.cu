// K-A: batched kernel-matrix assembly + in-shared-memory Cholesky + L^-1 + alpha.  One trial per CTA.
// (K-C, the marginal likelihood and the fit, lives in gp_fit.cu.)
//
// Replaces what gpytorch builds behind SingleTaskGP for every model (re)creation in the
// BO loop (Src/BO.py:65,142,176,424): ScaleKernel(Matern|RBF)(X,X) + sigma^2 I,
// psd_safe_cholesky (jitter 1e-8,1e-7,1e-6 in fp64), the prediction-strategy caches
// mean_cache = Khat^-1 (y - c) and the root of Khat^-1.
//
// Roofline: FP64-pipe / latency bound, O(n^3) flops per trial on an n x n matrix that never
// leaves shared memory (n <= 128: 132 KB). HBM traffic per trial: 24 n B in,
// 8 n + 8*wfrag_doubles(n) B out — negligible beside K-B.
#include "common.cuh"

namespace b200bo {

constexpr int CH_THREADS = 256;

__host__ __device__ inline int wfrag_blocks(int n) { int nI = (n + 7) / 8; return nI * (nI + 1); }

struct CholSmem {
  double* A;    // n x ld (row-major), lower triangle + diagonal
  double* xs;   // n
  double* ys;   // n
  double* res;  // n : y - c, later alpha
  double* z;    // n : L^-1 (y - c)
  double* dg;   // n : scratch diagonal
};

__device__ __forceinline__ CholSmem carve(double* sm, int n, int ld) {
  CholSmem s;
  s.A = sm;
  s.xs = sm + (size_t)n * ld;
  s.ys = s.xs + n;
  s.res = s.ys + n;
  s.z = s.res + n;
  s.dg = s.z + n;
  return s;
}

__host__ inline size_t chol_smem_bytes(int n, int ld) { return ((size_t)n * ld + 5 * (size_t)n) * sizeof(double); }

// assemble lower triangle of Khat (+ jitter) into A
__device__ __forceinline__ void assemble_lower(const CholSmem& s, int n, int ld, int kid, double noise, double sc,
                                               double ell, double a_scale, double jitter) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    if (c > r) continue;
    double dx = s.xs[r] - s.xs[c], dy = s.ys[r] - s.ys[c];
    double r2 = dx * dx + dy * dy;
    double v = kernel_from_r2_dyn(kid, r2, sc, ell, a_scale);
    if (c == r) v += noise + jitter;
    s.A[r * ld + c] = v;
  }
}

// right-looking Cholesky of the lower triangle in place; returns false if a pivot is not > 0.
// 16x16 thread tile over the trailing block, no integer division in the inner loops.
__device__ __forceinline__ bool cholesky_lower(double* A, int n, int ld) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    double d = A[j * ld + j];
    if (!(d > 0.0)) {          // uniform: every thread reads the same d
      __syncthreads();
      return false;
    }
    double inv = 1.0 / sqrt(d);
    __syncthreads();
    if (threadIdx.x == 0) A[j * ld + j] = sqrt(d);
    for (int r = j + 1 + threadIdx.x; r < n; r += blockDim.x) A[r * ld + j] *= inv;
    __syncthreads();
    for (int r = j + 1 + ty; r < n; r += 16) {
      double lrj = A[r * ld + j];
      for (int c = j + 1 + tx; c <= r; c += 16) A[r * ld + c] = fma(-lrj, A[c * ld + j], A[r * ld + c]);
    }
  }
  __syncthreads();
  return true;
}

// in-place inverse of the lower-triangular L (column by column from the last; LAPACK dtrti2
// order).  Two threads per row split the dot product (even / odd k).
__device__ __forceinline__ void invert_lower(double* A, int n, int ld) {
  const int half = threadIdx.x & 1;
  const int row_in_pass = threadIdx.x >> 1;
  const int rows_per_pass = blockDim.x >> 1;
  for (int j = n - 1; j >= 0; --j) {
    __syncthreads();
    double wjj = 1.0 / A[j * ld + j];
    // t[i] = sum_{k=j+1..i} W[i][k] * L[k][j]
    double tacc[2] = {0.0, 0.0};   // rows of this thread in up to 2 passes (n <= 2 * rows_per_pass)
    int cnt = 0;
    // block-uniform loop bounds: every lane of a warp reaches the shuffle
    for (int ibase = j + 1; ibase < n; ibase += rows_per_pass) {
      const int i = ibase + row_in_pass;
      double acc = 0.0;
      if (i < n)
        for (int k = j + 1 + half; k <= i; k += 2) acc = fma(A[i * ld + k], A[k * ld + j], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      tacc[cnt++] = acc;
    }
    __syncthreads();
    cnt = 0;
    for (int ibase = j + 1; ibase < n; ibase += rows_per_pass) {
      const int i = ibase + row_in_pass;
      if (i < n && half == 0) A[i * ld + j] = -tacc[cnt] * wjj;
      ++cnt;
    }
    if (threadIdx.x == 0) A[j * ld + j] = wjj;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CH_THREADS)
gp_chol_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
               const double* __restrict__ theta, double* __restrict__ L_out, double* __restrict__ Linv_out,
               double* __restrict__ alpha_out, double* __restrict__ logdet_out, double* __restrict__ wfrag_out,
               int32_t* __restrict__ status,
               int n_max, int n_cap, int ld, int kernel_id, size_t wfrag_stride) {
  extern __shared__ double sm[];
  const int t = blockIdx.x;
  const int n = n_arr[t];
  if (n <= 0) return;                    // inactive / finished trial
  if (n > n_cap) {                       // caller's n_hint was too small for this trial
    if (threadIdx.x == 0 && status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    return;
  }
  CholSmem s = carve(sm, n_cap, ld);
  __shared__ double red[CH_THREADS / 32];

  const double* th = theta + (size_t)t * 4;
  const double noise = th[0], cst = th[1], sc = th[2], ell = th[3];
  const KernParams kp = make_kern_params(sc, ell, kernel_id);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s.xs[i] = X[((size_t)t * n_max + i) * 2];
    s.ys[i] = X[((size_t)t * n_max + i) * 2 + 1];
    s.res[i] = y[(size_t)t * n_max + i] - cst;
  }
  __syncthreads();

  int st = 0;
  bool ok = false;
  // psd_safe_cholesky: plain, then total diagonal jitter 1e-8, 1e-7, 1e-6
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6));
    if (attempt > 0) st |= B200BO_ST_JITTER;
    assemble_lower(s, n, ld, kernel_id, noise, sc, ell, kp.a_scale, jitter);
    ok = cholesky_lower(s.A, n, ld);
  }
  if (!ok) {
    st |= B200BO_ST_CHOL_FAIL;
    if (threadIdx.x == 0 && status) atomicOr(&status[t], st);
    const double qnan = nan("");
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (alpha_out) alpha_out[(size_t)t * n_max + i] = qnan;
    if (threadIdx.x == 0) {
      if (logdet_out) logdet_out[t] = qnan;
    }
    return;
  }
  if (threadIdx.x == 0 && status && st) atomicOr(&status[t], st);

  // log|Khat| = 2 sum log L_ii  (fixed order: warp tree + serial over warps)
  double ld_acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ld_acc += log(s.A[i * ld + i]);
  ld_acc = warp_sum(ld_acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ld_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) tsum += red[w];
    if (logdet_out) logdet_out[t] = 2.0 * tsum;
  }
  if (L_out) {
    double* Lo = L_out + (size_t)t * n_max * n_max;
    for (int e = threadIdx.x; e < n_max * n_max; e += blockDim.x) {
      int r = e / n_max, c = e - r * n_max;
      Lo[e] = (r < n && c <= r) ? s.A[r * ld + c] : 0.0;
    }
  }
  invert_lower(s.A, n, ld);     // A now holds W = L^-1 (lower)

  // z = W (y - c) ; alpha = W^T z
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k <= i; ++k) acc = fma(s.A[i * ld + k], s.res[k], acc);
    s.z[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = i; k < n; ++k) acc = fma(s.A[k * ld + i], s.z[k], acc);
    s.res[i] = acc;     // alpha
  }
  __syncthreads();
  if (alpha_out)
    for (int i = threadIdx.x; i < n_max; i += blockDim.x) alpha_out[(size_t)t * n_max + i] = i < n ? s.res[i] : 0.0;
  if (Linv_out) {
    double* Wo = Linv_out + (size_t)t * n_max * n_max;
    for (int e = threadIdx.x; e < n_max * n_max; e += blockDim.x) {
      int r = e / n_max, c = e - r * n_max;
      Wo[e] = (r < n && c <= r) ? s.A[r * ld + c] : 0.0;
    }
  }
  if (wfrag_out) {
    // fragment-major layout for mma.m8n8k4 A operands: block (I,J), J <= 2I+1, at I(I+1)+J;
    // lane l holds W[8I + l/4][4J + l%4]
    double* wf = wfrag_out + (size_t)t * wfrag_stride;
    const int nblk = wfrag_blocks(n);      // K-B reads exactly this prefix for the same n
    for (int e = threadIdx.x; e < nblk * 32; e += blockDim.x) {
      int b = e >> 5, lane = e & 31;
      // invert b = I(I+1) + J
      int I = (int)((sqrt(4.0 * b + 1.0) - 1.0) * 0.5);
      while (I * (I + 1) > b) --I;
      while ((I + 1) * (I + 2) <= b) ++I;
      int J = b - I * (I + 1);
      int r = 8 * I + (lane >> 2), c = 4 * J + (lane & 3);
      wf[e] = (r < n && c <= r) ? s.A[r * ld + c] : 0.0;
    }
  }

}

static int launch_chol(const double* X, const double* y, const int32_t* n, const double* th, double* L, double* Linv,
                       double* alpha, double* logdet, double* wfrag, int32_t* status, int T, int n_max, int n_hint,
                       int kernel_id, cudaStream_t st) {
  int n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;   // shared memory is sized by the largest live n
  int ld = n_cap | 1;   // odd leading dimension: 2-way (= optimal for fp64) bank pattern on column walks
  size_t smem = chol_smem_bytes(n_cap, ld);
  if (smem > 227 * 1024) return arg_error("gp_chol: n_max=%d needs %zu B shared memory", n_max, smem);
  size_t wstride = (size_t)wfrag_blocks(n_max) * 32;
  cudaError_t e = cudaFuncSetAttribute(gp_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "gp_assemble_chol: smem attribute");
  gp_chol_kernel<<<T, CH_THREADS, smem, st>>>(X, y, n, th, L, Linv, alpha, logdet, wfrag, status, n_max, n_cap, ld,
                                              kernel_id, wstride);
  B200BO_CHECK_LAUNCH("gp_assemble_chol");
  return 0;
}

}  // namespace b200bo

extern "C" {

size_t b200bo_wfrag_doubles(int n_max) { return (size_t)b200bo::wfrag_blocks(n_max) * 32; }

int b200bo_gp_assemble_chol(const double* X, const double* y, const int32_t* n, const double* theta, double* L,
                            double* Linv, double* alpha, double* logdet, double* wfrag, int32_t* status, int T,
                            int n_max, int n_hint, int kernel_id, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("gp_assemble_chol: T=%d n_max=%d (max %d)", T, n_max, B200BO_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("gp_assemble_chol: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !theta) return arg_error("gp_assemble_chol: null input pointer");
  return launch_chol(X, y, n, theta, L, Linv, alpha, logdet, wfrag, status, T, n_max, n_hint, kernel_id, as_stream(stream));
}

}  // extern "C"
