// K-A: batched kernel-matrix assembly + in-shared-memory Cholesky + L^-1 + alpha,
// and K-C: marginal log-likelihood value and analytic gradient.  One trial per CTA.
//
// Replaces what gpytorch builds behind SingleTaskGP for every model (re)creation in the
// BO loop (Src/BO.py:65,142,176,424): ScaleKernel(Matern|RBF)(X,X) + sigma^2 I,
// psd_safe_cholesky (jitter 1e-8,1e-7,1e-6 in fp64), the prediction-strategy caches
// mean_cache = Khat^-1 (y - c) and the root of Khat^-1, and the
// ExactMarginalLogLikelihood + autograd used by fit_gpytorch_model (BO.py:144,189).
//
// Roofline: FP64-pipe / latency bound, O(n^3) flops per trial on an n x n matrix that never
// leaves shared memory (n <= 128: 132 KB). HBM traffic per trial: 24 n B in,
// 8 n + 8*wfrag_doubles(n) B out — negligible beside K-B.
#include "lbfgs.cuh"

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

template <bool MLL>
__global__ void __launch_bounds__(CH_THREADS)
gp_chol_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
               const double* __restrict__ theta_or_raw, double* __restrict__ L_out, double* __restrict__ Linv_out,
               double* __restrict__ alpha_out, double* __restrict__ logdet_out, double* __restrict__ wfrag_out,
               double* __restrict__ mll_value, double* __restrict__ mll_grad, int32_t* __restrict__ status,
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
  __shared__ double sh_scalar[8];

  double noise, cst, sc, ell, raw_s = 0.0, raw_l = 0.0;
  {
    const double* th = theta_or_raw + (size_t)t * 4;
    noise = th[0];
    cst = th[1];
    if (MLL) {   // raw -> actual: softplus with torch's threshold 20 (oracle/gp.py:softplus)
      raw_s = th[2];
      raw_l = th[3];
      sc = raw_s > 20.0 ? raw_s : log1p(exp(raw_s));
      ell = raw_l > 20.0 ? raw_l : log1p(exp(raw_l));
    } else {
      sc = th[2];
      ell = th[3];
    }
  }
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
      if (MLL) {
        mll_value[t] = qnan;
        for (int q = 0; q < 4; ++q) mll_grad[(size_t)t * 4 + q] = qnan;
      }
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
    sh_scalar[0] = 2.0 * tsum;
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

  if (MLL) {
    // value = [ -1/2 z.z - 1/2 log|K| - n/2 log 2pi + log Gamma(noise; 1.1, 0.05) ] / n
    // grad wrt (noise, c, raw_s, raw_l); A = alpha alpha^T - Khat^-1
    const double a_prior = 1.1, rate = 0.05;
    // Khat^-1 = W^T W : strict upper triangle of the square array receives Kinv[r][c] (r < c),
    // the diagonal goes to s.dg.
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      int r = e / n, c = e - r * n;
      if (c < r) continue;
      double acc = 0.0;
      for (int k = c; k < n; ++k) acc = fma(s.A[k * ld + r], s.A[k * ld + c], acc);
      if (c == r) s.dg[r] = acc;
      else s.A[r * ld + c] = acc;   // strict upper part is free: W is lower
    }
    __syncthreads();
    // partial sums: q0 = z.z (recomputed from alpha: (y-c).alpha), q1 = sum alpha, q2 = tr(A),
    // q3 = sum A_ij Kl_ij, q4 = sum A_ij dK_ij
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double al = s.res[i];
      q0 = fma(y[(size_t)t * n_max + i] - cst, al, q0);
      q1 += al;
      q2 += al * al - s.dg[i];
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      int r = e / n, c = e - r * n;
      if (c < r) continue;
      double kinv = (c == r) ? s.dg[r] : s.A[r * ld + c];
      double a_rc = s.res[r] * s.res[c] - kinv;
      double dx = s.xs[r] - s.xs[c], dy = s.ys[r] - s.ys[c];
      double r2 = dx * dx + dy * dy;
      double kl = kernel_from_r2_dyn(kernel_id, r2, 1.0, ell, kp.a_scale);
      double dk = dkernel_dell_from_r2(kernel_id, r2, sc, ell, kp.a_scale);
      double w = (c == r) ? 1.0 : 2.0;     // symmetric: count (r,c) and (c,r)
      q3 = fma(w * a_rc, kl, q3);
      q4 = fma(w * a_rc, dk, q4);
    }
    double q[5] = {q0, q1, q2, q3, q4};
    for (int m = 0; m < 5; ++m) {
      double v = warp_sum(q[m]);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int w = 0; w < CH_THREADS / 32; ++w) tsum += red[w];
        sh_scalar[1 + m] = tsum;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const double nn = (double)n;
      const double logdet = sh_scalar[0];
      const double ll = -0.5 * sh_scalar[1] - 0.5 * logdet - 0.5 * nn * kLog2Pi;
      const double lp = a_prior * log(rate) + (a_prior - 1.0) * log(noise) - rate * noise - lgamma(a_prior);
      mll_value[t] = (ll + lp) / nn;
      const double sig_s = raw_s >= 0 ? 1.0 / (1.0 + exp(-raw_s)) : exp(raw_s) / (1.0 + exp(raw_s));
      const double sig_l = raw_l >= 0 ? 1.0 / (1.0 + exp(-raw_l)) : exp(raw_l) / (1.0 + exp(raw_l));
      double* g = mll_grad + (size_t)t * 4;
      g[0] = (0.5 * sh_scalar[3] + (a_prior - 1.0) / noise - rate) / nn;
      g[1] = sh_scalar[2] / nn;
      g[2] = 0.5 * sh_scalar[4] * sig_s / nn;
      g[3] = 0.5 * sh_scalar[5] * sig_l / nn;
    }
  }
}

// ---- fused hyper-parameter fit: the whole L-BFGS run of one trial inside one CTA ------------------
// One launch replaces the ~40-130 (K-C launch + optimiser-step launch) rounds of a fit: the CTA keeps the
// observations in shared memory, evaluates the marginal likelihood and its gradient (same arithmetic as
// gp_chol_kernel<true>), lets thread 0 advance the L-BFGS state machine (lbfgs.cuh) and repeats until the
// trial converges. Trials are independent, so nobody waits for the slowest trial of the batch.
__device__ __forceinline__ void mll_eval_block(const CholSmem& s, const double* __restrict__ yv, int n, int ld,
                                               int kernel_id, const double* raw, double* red, double* sh_scalar,
                                               int* sh_st) {
  // outputs in shared memory: sh_scalar[6] = value, sh_scalar[7..10] = gradient, *sh_st = status bits
  const double noise = raw[0], cst = raw[1], raw_s = raw[2], raw_l = raw[3];
  const double sc = raw_s > 20.0 ? raw_s : log1p(exp(raw_s));
  const double ell = raw_l > 20.0 ? raw_l : log1p(exp(raw_l));
  const KernParams kp = make_kern_params(sc, ell, kernel_id);
  const double a_prior = 1.1, rate = 0.05;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s.res[i] = yv[i] - cst;
  int st = 0;
  bool ok = false;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6));
    if (attempt > 0) st |= B200BO_ST_JITTER;
    __syncthreads();
    assemble_lower(s, n, ld, kernel_id, noise, sc, ell, kp.a_scale, jitter);
    ok = cholesky_lower(s.A, n, ld);
  }
  if (!ok) {
    if (threadIdx.x == 0) {
      *sh_st = st | B200BO_ST_CHOL_FAIL;
      for (int q = 6; q < 11; ++q) sh_scalar[q] = nan("");
    }
    __syncthreads();
    return;
  }
  double ld_acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ld_acc += log(s.A[i * ld + i]);
  ld_acc = warp_sum(ld_acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ld_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) tsum += red[w];
    sh_scalar[0] = 2.0 * tsum;
    *sh_st = st;
  }
  invert_lower(s.A, n, ld);
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
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    if (c < r) continue;
    double acc = 0.0;
    for (int k = c; k < n; ++k) acc = fma(s.A[k * ld + r], s.A[k * ld + c], acc);
    if (c == r) s.dg[r] = acc;
    else s.A[r * ld + c] = acc;
  }
  __syncthreads();
  double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double al = s.res[i];
    q0 = fma(yv[i] - cst, al, q0);
    q1 += al;
    q2 += al * al - s.dg[i];
  }
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    if (c < r) continue;
    double kinv = (c == r) ? s.dg[r] : s.A[r * ld + c];
    double a_rc = s.res[r] * s.res[c] - kinv;
    double dx = s.xs[r] - s.xs[c], dy = s.ys[r] - s.ys[c];
    double r2 = dx * dx + dy * dy;
    double kl = kernel_from_r2_dyn(kernel_id, r2, 1.0, ell, kp.a_scale);
    double dk = dkernel_dell_from_r2(kernel_id, r2, sc, ell, kp.a_scale);
    double w = (c == r) ? 1.0 : 2.0;
    q3 = fma(w * a_rc, kl, q3);
    q4 = fma(w * a_rc, dk, q4);
  }
  double q[5] = {q0, q1, q2, q3, q4};
  for (int m = 0; m < 5; ++m) {
    double v = warp_sum(q[m]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tsum = 0.0;
      for (int w = 0; w < CH_THREADS / 32; ++w) tsum += red[w];
      sh_scalar[1 + m] = tsum;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double nn = (double)n;
    const double ll = -0.5 * sh_scalar[1] - 0.5 * sh_scalar[0] - 0.5 * nn * kLog2Pi;
    const double lp = a_prior * log(rate) + (a_prior - 1.0) * log(noise) - rate * noise - lgamma(a_prior);
    const double sig_s = raw_s >= 0 ? 1.0 / (1.0 + exp(-raw_s)) : exp(raw_s) / (1.0 + exp(raw_s));
    const double sig_l = raw_l >= 0 ? 1.0 / (1.0 + exp(-raw_l)) : exp(raw_l) / (1.0 + exp(raw_l));
    sh_scalar[6] = (ll + lp) / nn;
    sh_scalar[7] = (0.5 * sh_scalar[3] + (a_prior - 1.0) / noise - rate) / nn;
    sh_scalar[8] = sh_scalar[2] / nn;
    sh_scalar[9] = 0.5 * sh_scalar[4] * sig_s / nn;
    sh_scalar[10] = 0.5 * sh_scalar[5] * sig_l / nn;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CH_THREADS)
gp_fit_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
              const double* __restrict__ raw0, int raw0_per_trial, double* __restrict__ theta_out,
              double* __restrict__ raw_out, double* __restrict__ mll_out, int32_t* __restrict__ iters_out,
              int32_t* __restrict__ status, int n_max, int n_cap, int ld, int kernel_id, LbfgsOpts o, int max_rounds) {
  extern __shared__ double sm[];
  const int t = blockIdx.x;
  const int n = n_arr[t];
  if (n <= 0) return;
  if (n > n_cap) {
    if (threadIdx.x == 0 && status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    return;
  }
  CholSmem s = carve(sm, n_cap, ld);
  double* yv = s.dg + n_cap;
  __shared__ double red[CH_THREADS / 32];
  __shared__ double sh_scalar[12];
  __shared__ LbfgsState lst;
  __shared__ double xt[4];
  __shared__ int sh_done, sh_st;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s.xs[i] = X[((size_t)t * n_max + i) * 2];
    s.ys[i] = X[((size_t)t * n_max + i) * 2 + 1];
    yv[i] = y[(size_t)t * n_max + i];
  }
  if (threadIdx.x == 0) {
    lst.hist_n = 0; lst.hist_head = 0; lst.iter = 0; lst.ls_iter = 0; lst.nfev = 0; lst.resets = 0; lst.phase = 0;
    for (int i = 0; i < 4; ++i) xt[i] = raw0[raw0_per_trial ? (size_t)t * 4 + i : i];
    xt[0] = fmax(xt[0], o.lb_noise);
    sh_done = 0;
    sh_st = 0;
  }
  __syncthreads();
  for (int round = 0; round <= max_rounds; ++round) {
    mll_eval_block(s, yv, n, ld, kernel_id, xt, red, sh_scalar, &sh_st);
    if (threadIdx.x == 0) {
      LbfgsOpts oo = o;
      if (round == max_rounds) oo.max_ls = -1;          // round budget exhausted: forced finish
      double gt[4];
      for (int i = 0; i < 4; ++i) gt[i] = -sh_scalar[7 + i];
      sh_done = lbfgs_advance(lst, oo, xt, -sh_scalar[6], gt) ? 1 : 0;
    }
    __syncthreads();
    if (sh_done) break;
  }
  if (threadIdx.x == 0) {
    if (raw_out) for (int i = 0; i < 4; ++i) raw_out[(size_t)t * 4 + i] = lst.x[i];
    if (theta_out) {
      theta_out[(size_t)t * 4 + 0] = lst.x[0];
      theta_out[(size_t)t * 4 + 1] = lst.x[1];
      theta_out[(size_t)t * 4 + 2] = softplus20(lst.x[2]);
      theta_out[(size_t)t * 4 + 3] = softplus20(lst.x[3]);
    }
    if (mll_out) mll_out[t] = -lst.f;
    if (iters_out) { iters_out[(size_t)t * 2] = lst.iter; iters_out[(size_t)t * 2 + 1] = lst.nfev; }
  }
}

static int launch_chol(bool mll, const double* X, const double* y, const int32_t* n, const double* th, double* L,
                       double* Linv, double* alpha, double* logdet, double* wfrag, double* value, double* grad,
                       int32_t* status, int T, int n_max, int n_hint, int kernel_id, cudaStream_t st) {
  int n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;   // shared memory is sized by the largest live n
  int ld = n_cap | 1;   // odd leading dimension: 2-way (= optimal for fp64) bank pattern on column walks
  size_t smem = chol_smem_bytes(n_cap, ld);
  if (smem > 227 * 1024) return arg_error("gp_chol: n_max=%d needs %zu B shared memory", n_max, smem);
  size_t wstride = (size_t)wfrag_blocks(n_max) * 32;
  cudaError_t e;
  if (mll) {
    e = cudaFuncSetAttribute(gp_chol_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "mll_value_grad: smem attribute");
    gp_chol_kernel<true><<<T, CH_THREADS, smem, st>>>(X, y, n, th, L, Linv, alpha, logdet, wfrag, value, grad, status,
                                                      n_max, n_cap, ld, kernel_id, wstride);
  } else {
    e = cudaFuncSetAttribute(gp_chol_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "gp_assemble_chol: smem attribute");
    gp_chol_kernel<false><<<T, CH_THREADS, smem, st>>>(X, y, n, th, L, Linv, alpha, logdet, wfrag, nullptr, nullptr,
                                                       status, n_max, n_cap, ld, kernel_id, wstride);
  }
  B200BO_CHECK_LAUNCH(mll ? "mll_value_grad" : "gp_assemble_chol");
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
  return launch_chol(false, X, y, n, theta, L, Linv, alpha, logdet, wfrag, nullptr, nullptr, status, T, n_max, n_hint,
                     kernel_id, as_stream(stream));
}

int b200bo_gp_fit(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                  double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                  int n_max, int n_hint, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter,
                  int max_ls, int max_rounds, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("gp_fit: T=%d n_max=%d (max %d)", T, n_max, B200BO_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("gp_fit: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !raw0) return arg_error("gp_fit: null pointer");
  int n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;
  int ld = n_cap | 1;
  size_t smem = chol_smem_bytes(n_cap, ld) + (size_t)n_cap * sizeof(double);
  if (smem > 227 * 1024) return arg_error("gp_fit: n=%d needs %zu B shared memory", n_cap, smem);
  cudaError_t e = cudaFuncSetAttribute(gp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "gp_fit: smem attribute");
  LbfgsOpts o;
  o.lb_noise = lb_noise; o.gtol = gtol; o.ftol = ftol; o.c1 = 1e-4; o.max_iter = max_iter; o.max_ls = max_ls;
  gp_fit_kernel<<<T, CH_THREADS, smem, as_stream(stream)>>>(X, y, n, raw0, raw0_per_trial, theta_out, raw_out, mll_out,
                                                          iters_out, status, n_max, n_cap, ld, kernel_id, o, max_rounds);
  B200BO_CHECK_LAUNCH("gp_fit");
  return 0;
}

int b200bo_mll_value_grad(const double* X, const double* y, const int32_t* n, const double* raw, double* value,
                          double* grad, int32_t* status, int T, int n_max, int n_hint, int kernel_id, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("mll_value_grad: T=%d n_max=%d (max %d)", T, n_max, B200BO_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("mll_value_grad: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !raw || !value || !grad) return arg_error("mll_value_grad: null pointer");
  return launch_chol(true, X, y, n, raw, nullptr, nullptr, nullptr, nullptr, nullptr, value, grad, status, T, n_max,
                     n_hint, kernel_id, as_stream(stream));
}

}  // extern "C"
