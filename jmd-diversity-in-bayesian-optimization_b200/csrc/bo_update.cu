// Per-iteration bookkeeping of the batched BO loop (one thread per trial).
//
// Replaces the Python statements of BOloop / BOloop_fixparam between two acquisitions
// (Src/BO.py:104 objective look-up, :164-165 torch.cat append, :168-169 yopt / yall.append,
// :241 error = optimal_y - yopt, loop condition :152) and the post-loop padding of yall to
// max_iter + 2 entries (BO.py:262, 267-271).
#include "common.cuh"

namespace b200bo {

__global__ void __launch_bounds__(128)
bo_update_kernel(const double2* __restrict__ grid, int G, const float* __restrict__ obj, const int32_t* __restrict__ surf_id,
                 const int32_t* __restrict__ arg_idx, double* __restrict__ X, double* __restrict__ y,
                 int32_t* __restrict__ n, double* __restrict__ best_f, double* __restrict__ yall,
                 int32_t* __restrict__ picked, int32_t* __restrict__ nit, int T, int n_max, int max_iter, int iter,
                 double tol, double optimal_y) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int nn = n[t];
  if (nn <= 0) return;                 // finished earlier
  int idx = arg_idx[t];
  if (idx < 0 || nn >= n_max) {        // nothing selectable (status already says why): stop this trial
    n[t] = -nn;
    return;
  }
  double2 x = grid[idx];
  // torch.tensor(obj_func(new_x), dtype=torch.float32) (BO.py:104): the objective is fp32
  double yn = (double)obj[(size_t)surf_id[t] * G + idx];
  X[((size_t)t * n_max + nn) * 2] = x.x;
  X[((size_t)t * n_max + nn) * 2 + 1] = x.y;
  y[(size_t)t * n_max + nn] = yn;
  double b = fmax(best_f[t], yn);
  best_f[t] = b;
  picked[(size_t)t * max_iter + iter] = idx;
  yall[(size_t)t * (max_iter + 2) + iter + 1] = b;
  nit[t] = iter + 1;
  nn += 1;
  // while error > tol and iteration < max_iter (BO.py:152)
  bool stop = !((optimal_y - b) > tol) || (iter + 1 >= max_iter);
  n[t] = stop ? -nn : nn;             // negative = finished, |n| = number of observations
}

__global__ void __launch_bounds__(128)
bo_init_kernel(const double* __restrict__ y, int32_t* __restrict__ n, double* __restrict__ best_f,
               double* __restrict__ yall, int32_t* __restrict__ picked, int32_t* __restrict__ nit, int T, int n_max,
               int max_iter, double tol, double optimal_y) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int nn = n[t];
  double b = -INFINITY;
  for (int i = 0; i < nn; ++i) b = fmax(b, y[(size_t)t * n_max + i]);
  best_f[t] = b;
  yall[(size_t)t * (max_iter + 2)] = b;      // result.yall.append(result.yopt), BO.py:146
  for (int i = 0; i < max_iter; ++i) picked[(size_t)t * max_iter + i] = -1;
  nit[t] = 0;
  // error starts at 100 (BO.py:147): the first iteration always runs when max_iter > 0
  if (max_iter <= 0) n[t] = -nn;
}

__global__ void __launch_bounds__(128)
bo_finalize_kernel(double* __restrict__ yall, const int32_t* __restrict__ nit, int32_t* __restrict__ n, int T,
                   int max_iter) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int k = nit[t];
  double last = yall[(size_t)t * (max_iter + 2) + k];
  // yall.append(yopt) once more, then pad with the last value to max_iter + 2 (BO.py:262,267-271)
  for (int i = k + 1; i < max_iter + 2; ++i) yall[(size_t)t * (max_iter + 2) + i] = last;
  if (n[t] < 0) n[t] = -n[t];
}

// Which trials the hyper-parameter fit before acquisition `iter` covers: the live ones, plus the trials that stopped
// in iteration iter-1 — the reference refits after EVERY append, also the last one, and records that model
// (BO.py:176-189 then res_hp.addhp, :252), so a finished trial is fitted once more on its final data.
__global__ void __launch_bounds__(128)
bo_fit_counts_kernel(const int32_t* __restrict__ n, const int32_t* __restrict__ nit, int32_t* __restrict__ n_fit, int T,
                     int iter) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int nn = n[t];
  n_fit[t] = nn > 0 ? nn : (nit[t] == iter && iter > 0 ? -nn : 0);
}

}  // namespace b200bo

extern "C" {

int b200bo_bo_fit_counts(const int32_t* n, const int32_t* nit, int32_t* n_fit, int T, int iter, void* stream) {
  using namespace b200bo;
  if (T < 0 || iter < 0) return arg_error("bo_fit_counts: bad sizes");
  if (T == 0) return 0;
  if (!n || !nit || !n_fit) return arg_error("bo_fit_counts: null pointer");
  bo_fit_counts_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(n, nit, n_fit, T, iter);
  B200BO_CHECK_LAUNCH("bo_fit_counts");
  return 0;
}

int b200bo_bo_init(const double* y, int32_t* n, double* best_f, double* yall, int32_t* picked, int32_t* nit, int T,
                   int n_max, int max_iter, double tol, double optimal_y, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || max_iter < 0) return arg_error("bo_init: bad sizes");
  if (T == 0) return 0;
  if (!y || !n || !best_f || !yall || !picked || !nit) return arg_error("bo_init: null pointer");
  bo_init_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(y, n, best_f, yall, picked, nit, T, n_max, max_iter, tol, optimal_y);
  B200BO_CHECK_LAUNCH("bo_init");
  return 0;
}

int b200bo_bo_update(const double* grid, int G, const float* obj, const int32_t* surf_id, const int32_t* arg_idx,
                     double* X, double* y, int32_t* n, double* best_f, double* yall, int32_t* picked, int32_t* nit,
                     int T, int n_max, int max_iter, int iter, double tol, double optimal_y, void* stream) {
  using namespace b200bo;
  if (T < 0 || G < 1 || n_max < 1 || iter < 0 || iter >= max_iter) return arg_error("bo_update: bad sizes");
  if (T == 0) return 0;
  if (!grid || !obj || !surf_id || !arg_idx || !X || !y || !n || !best_f || !yall || !picked || !nit)
    return arg_error("bo_update: null pointer");
  bo_update_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(
      reinterpret_cast<const double2*>(grid), G, obj, surf_id, arg_idx, X, y, n, best_f, yall, picked, nit, T, n_max,
      max_iter, iter, tol, optimal_y);
  B200BO_CHECK_LAUNCH("bo_update");
  return 0;
}

int b200bo_bo_finalize(double* yall, const int32_t* nit, int32_t* n, int T, int max_iter, void* stream) {
  using namespace b200bo;
  if (T == 0) return 0;
  if (!yall || !nit || !n) return arg_error("bo_finalize: null pointer");
  bo_finalize_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(yall, nit, n, T, max_iter);
  B200BO_CHECK_LAUNCH("bo_finalize");
  return 0;
}

}  // extern "C"
