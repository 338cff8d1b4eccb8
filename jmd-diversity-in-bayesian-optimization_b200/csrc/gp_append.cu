// K-A (append variant): rank-1 extension of L^-1 and alpha when one observation is appended and the
// hyper-parameters are unchanged — the fixed-theta loops (BOloop_fixparam, Src/BO.py:380-461) re-create
// the same model plus one point every iteration (BO.py:416-429). One warp per trial.
//
//   K_{m+1} = [K_m k; k^T kappa],  L_{m+1} = [L_m 0; l^T lambda],  l = L_m^-1 k,  lambda^2 = kappa - l.l
//   W_{m+1} = L_{m+1}^-1 = [W_m 0; w^T 1/lambda],  w = -(W_m^T l) / lambda
//   z_{m+1} = [z_m; (w.(y_m - c) + (y_new - c) / lambda)],  alpha = W_{m+1}^T z_{m+1}
//
// Three triangular matrix-vector products over the fragment-major L^-1 (O(n^2) flops, ~3 x 8 wfrag(n)
// bytes from L2/HBM) replace the O(n^3) refactorisation of gp_chol.cu. A non-positive lambda^2 (what
// psd_safe_cholesky would answer with jitter) is not repaired here: the trial is handed to the full
// kernel through n_redo[], which the caller launches right after (inactive trials cost nothing there).
#include "common.cuh"

namespace b200bo {

constexpr int AP_WARPS = 4;

// l[r] = sum_{c <= r, c < m} W[r][c] v[c] for r < m  (v in shared memory, result to shared memory)
__device__ __forceinline__ void tri_matvec(const double* __restrict__ wf, int m, const double* v, double* out, int lane) {
  const int nI = (m + 7) >> 3;
  const int rr = lane >> 2, cc = lane & 3;
  for (int I = 0; I < nI; ++I) {
    double acc = 0.0;
    const double* blk = wf + (size_t)(I * (I + 1)) * 32 + lane;
    for (int J = 0; J <= 2 * I + 1; ++J) {
      const int c = 4 * J + cc;
      const double vv = c < m ? v[c] : 0.0;
      acc = fma(blk[J * 32], vv, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const int r = 8 * I + rr;
    if (cc == 0 && r < m) out[r] = acc;
  }
}

// out[c] = sum_{r >= c, r < m} W[r][c] v[r] for c < m
__device__ __forceinline__ void tri_matvec_t(const double* __restrict__ wf, int m, const double* v, double* out, int lane) {
  const int nI = (m + 7) >> 3;
  const int nJ = 2 * nI;
  const int rr = lane >> 2, cc = lane & 3;
  for (int J = 0; J < nJ; ++J) {
    double acc = 0.0;
    for (int I = J >> 1; I < nI; ++I) {
      const int r = 8 * I + rr;
      const double vv = r < m ? v[r] : 0.0;
      acc = fma(wf[(size_t)(I * (I + 1) + J) * 32 + lane], vv, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    const int c = 4 * J + cc;
    if (rr == 0 && c < m) out[c] = acc;
  }
}

__global__ void __launch_bounds__(AP_WARPS * 32)
gp_append_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                 const double* __restrict__ theta, double* __restrict__ wfrag, double* __restrict__ zvec,
                 double* __restrict__ alpha, double* __restrict__ wrow_out, int32_t* __restrict__ n_redo,
                 int32_t* __restrict__ status, int T, int n_max, int kernel_id, size_t wstride) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * AP_WARPS + warp;
  if (t >= T) return;
  double* kv = sm + (size_t)warp * 3 * n_max;   // k, later w
  double* lv = kv + n_max;                      // l
  double* zz = lv + n_max;                      // z
  const int n = n_arr[t];
  if (n_redo) { if (lane == 0) n_redo[t] = 0; }
  if (n <= 0) return;
  const int m = n - 1;                          // size before the append
  const double noise = theta[(size_t)t * 4], cst = theta[(size_t)t * 4 + 1];
  const double sc = theta[(size_t)t * 4 + 2], ell = theta[(size_t)t * 4 + 3];
  const KernParams kp = make_kern_params(sc, ell, kernel_id);
  double* wf = wfrag + (size_t)t * wstride;
  const double xn = X[((size_t)t * n_max + m) * 2], yn = X[((size_t)t * n_max + m) * 2 + 1];
  for (int j = lane; j < m; j += 32) {
    const double dx = xn - X[((size_t)t * n_max + j) * 2], dy = yn - X[((size_t)t * n_max + j) * 2 + 1];
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    kv[j] = kernel_from_r2_dyn(kernel_id, r2, sc, ell, kp.a_scale);
    zz[j] = zvec[(size_t)t * n_max + j];
  }
  __syncwarp();
  tri_matvec(wf, m, kv, lv, lane);
  __syncwarp();
  double ll = 0.0;
  for (int j = lane; j < m; j += 32) ll = fma(lv[j], lv[j], ll);
  ll = warp_sum(ll);
  const double kappa = kernel_from_r2_dyn(kernel_id, 0.0, sc, ell, kp.a_scale) + noise;
  const double lam2 = kappa - ll;
  if (!(lam2 > 0.0)) {                          // uniform across the warp
    if (lane == 0) {
      if (n_redo) n_redo[t] = n;
      else if (status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    }
    if (wrow_out)                                 // no usable new row: the incremental posterior stays as it was
      for (int c = lane; c <= m; c += 32) wrow_out[(size_t)t * n_max + c] = 0.0;
    return;
  }
  const double lam = sqrt(lam2), inv_lam = 1.0 / lam;
  tri_matvec_t(wf, m, lv, kv, lane);            // kv <- W^T l
  __syncwarp();
  // new row of W in fragment layout: row m -> row block I = m/8, lanes (m%8)*4 + c%4 of block (I, c/4)
  const int I = m >> 3, rl = (m & 7) * 4;
  double zacc = 0.0;
  for (int c = lane; c <= m; c += 32) {
    double w;
    if (c < m) {
      w = -kv[c] * inv_lam;
      zacc = fma(w, y[(size_t)t * n_max + c] - cst, zacc);
      kv[c] = w;
    } else {
      w = inv_lam;
    }
    wf[(size_t)(I * (I + 1) + (c >> 2)) * 32 + rl + (c & 3)] = w;
    if (wrow_out) wrow_out[(size_t)t * n_max + c] = w;     // dense copy of the new row for the incremental posterior
  }
  zacc = warp_sum(zacc);
  const double znew = zacc + (y[(size_t)t * n_max + m] - cst) * inv_lam;
  if (lane == 0) {
    zvec[(size_t)t * n_max + m] = znew;
    zz[m] = znew;
  }
  __syncwarp();
  // alpha = W_{m+1}^T z : old rows from the fragments, new row from kv / inv_lam
  tri_matvec_t(wf, m, zz, lv, lane);
  __syncwarp();
  for (int c = lane; c <= m; c += 32) {
    const double a = c < m ? lv[c] + kv[c] * znew : inv_lam * znew;
    alpha[(size_t)t * n_max + c] = a;
  }
}

// z = W (y - c) for a freshly factorised state (after the full K-A): one warp per trial
__global__ void __launch_bounds__(AP_WARPS * 32)
gp_zvec_kernel(const double* __restrict__ y, const int32_t* __restrict__ n_arr, const double* __restrict__ theta,
               const double* __restrict__ wfrag, double* __restrict__ zvec, int T, int n_max, size_t wstride) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * AP_WARPS + warp;
  if (t >= T) return;
  double* res = sm + (size_t)warp * 2 * n_max;
  double* out = res + n_max;
  const int n = n_arr[t];
  if (n <= 0) return;
  const double cst = theta[(size_t)t * 4 + 1];
  for (int j = lane; j < n; j += 32) res[j] = y[(size_t)t * n_max + j] - cst;
  __syncwarp();
  tri_matvec(wfrag + (size_t)t * wstride, n, res, out, lane);
  __syncwarp();
  for (int j = lane; j < n; j += 32) zvec[(size_t)t * n_max + j] = out[j];
}

}  // namespace b200bo

extern "C" {

int b200bo_gp_chol_append(const double* X, const double* y, const int32_t* n, const double* theta, double* wfrag,
                          double* zvec, double* alpha, double* wrow_out, int32_t* n_redo, int32_t* status, int T,
                          int n_max, int kernel_id, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("gp_chol_append: T=%d n_max=%d", T, n_max);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("gp_chol_append: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !theta || !wfrag || !zvec || !alpha) return arg_error("gp_chol_append: null pointer");
  size_t smem = (size_t)AP_WARPS * 3 * n_max * sizeof(double);
  gp_append_kernel<<<(T + AP_WARPS - 1) / AP_WARPS, AP_WARPS * 32, smem, as_stream(stream)>>>(
      X, y, n, theta, wfrag, zvec, alpha, wrow_out, n_redo, status, T, n_max, kernel_id, b200bo_wfrag_doubles(n_max));
  B200BO_CHECK_LAUNCH("gp_chol_append");
  return 0;
}

int b200bo_gp_zvec(const double* y, const int32_t* n, const double* theta, const double* wfrag, double* zvec, int T,
                   int n_max, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("gp_zvec: T=%d n_max=%d", T, n_max);
  if (T == 0) return 0;
  if (!y || !n || !theta || !wfrag || !zvec) return arg_error("gp_zvec: null pointer");
  size_t smem = (size_t)AP_WARPS * 2 * n_max * sizeof(double);
  gp_zvec_kernel<<<(T + AP_WARPS - 1) / AP_WARPS, AP_WARPS * 32, smem, as_stream(stream)>>>(
      y, n, theta, wfrag, zvec, T, n_max, b200bo_wfrag_doubles(n_max));
  B200BO_CHECK_LAUNCH("gp_zvec");
  return 0;
}

}  // extern "C"
