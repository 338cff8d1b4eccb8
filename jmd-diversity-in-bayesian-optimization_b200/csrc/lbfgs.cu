// Batched hyper-parameter fit driver: one projected L-BFGS state machine per trial.
//
// Replaces botorch.fit.fit_gpytorch_model -> scipy L-BFGS-B over the 4 raw parameters
// (noise >= 1e-4 bound, constant, raw_outputscale, raw_lengthscale) called before every acquisition
// in BOloop (Src/BO.py:144,189). Objective and gradient come from K-C (gp_chol.cu); this kernel only
// advances each trial's optimiser by one function evaluation: Armijo backtracking along the L-BFGS
// direction (m = 10 pairs, as scipy's default), the bound handled by projection with the active
// coordinate removed from the two-loop recursion. scipy's Fortran trajectory is not reproduced
// (SURVEY.md §7 hard parts): parity is on the objective/gradient and on the optimum reached.
#include "lbfgs.cuh"

namespace b200bo {

__global__ void __launch_bounds__(128)
lbfgs_step_kernel(LbfgsState* __restrict__ st, double* __restrict__ xt_all, const double* __restrict__ val,
                  const double* __restrict__ grad, int32_t* __restrict__ n_eval, const int32_t* __restrict__ n_obs,
                  double* __restrict__ theta_out, double* __restrict__ raw_out, double* __restrict__ mll_out,
                  int32_t* __restrict__ iters_out, int T, LbfgsOpts o) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  LbfgsState& s = st[t];
  if (s.phase == 2) return;
  double* xt = xt_all + (size_t)t * 4;
  double gt[4];
  for (int i = 0; i < 4; ++i) gt[i] = -grad[(size_t)t * 4 + i];     // the engine minimises F = -MLL
  if (lbfgs_advance(s, o, xt, -val[t], gt)) {
    n_eval[t] = 0;                             // K-C skips this trial from now on
    if (raw_out) for (int i = 0; i < 4; ++i) raw_out[(size_t)t * 4 + i] = s.x[i];
    if (theta_out) {
      theta_out[(size_t)t * 4 + 0] = s.x[0];
      theta_out[(size_t)t * 4 + 1] = s.x[1];
      theta_out[(size_t)t * 4 + 2] = softplus20(s.x[2]);
      theta_out[(size_t)t * 4 + 3] = softplus20(s.x[3]);
    }
    if (mll_out) mll_out[t] = -s.f;
    if (iters_out) { iters_out[(size_t)t * 2] = s.iter; iters_out[(size_t)t * 2 + 1] = s.nfev; }
  }
}

__global__ void __launch_bounds__(128)
lbfgs_init_kernel(LbfgsState* __restrict__ st, double* __restrict__ xt_all, const double* __restrict__ raw0,
                  int raw0_per_trial, int32_t* __restrict__ n_eval, const int32_t* __restrict__ n_obs, int T, double lb) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  LbfgsState& s = st[t];
  s.hist_n = 0; s.hist_head = 0; s.iter = 0; s.ls_iter = 0; s.nfev = 0; s.resets = 0;
  int nn = n_obs[t];
  bool active = nn > 0;
  s.phase = active ? 0 : 2;
  n_eval[t] = active ? nn : 0;
  for (int i = 0; i < 4; ++i) xt_all[(size_t)t * 4 + i] = raw0[raw0_per_trial ? (size_t)t * 4 + i : i];
  xt_all[(size_t)t * 4] = fmax(xt_all[(size_t)t * 4], lb);
}

}  // namespace b200bo

extern "C" {

size_t b200bo_lbfgs_state_bytes(int T) { return (size_t)(T > 0 ? T : 0) * sizeof(b200bo::LbfgsState); }

int b200bo_lbfgs_init(void* state, double* xt, const double* raw0, int raw0_per_trial, int32_t* n_eval,
                      const int32_t* n_obs, int T, double lb_noise, void* stream) {
  using namespace b200bo;
  if (T < 0) return arg_error("lbfgs_init: T=%d", T);
  if (T == 0) return 0;
  if (!state || !xt || !raw0 || !n_eval || !n_obs) return arg_error("lbfgs_init: null pointer");
  lbfgs_init_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(reinterpret_cast<LbfgsState*>(state), xt, raw0,
                                                                   raw0_per_trial, n_eval, n_obs, T, lb_noise);
  B200BO_CHECK_LAUNCH("lbfgs_init");
  return 0;
}

int b200bo_lbfgs_step(void* state, double* xt, const double* value, const double* grad, int32_t* n_eval,
                      const int32_t* n_obs, double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out,
                      int T, double lb_noise, double gtol, double ftol, int max_iter, int max_ls, void* stream) {
  using namespace b200bo;
  if (T < 0) return arg_error("lbfgs_step: T=%d", T);
  if (T == 0) return 0;
  if (!state || !xt || !value || !grad || !n_eval || !n_obs) return arg_error("lbfgs_step: null pointer");
  LbfgsOpts o;
  o.lb_noise = lb_noise; o.gtol = gtol; o.ftol = ftol; o.c1 = 1e-4; o.max_iter = max_iter; o.max_ls = max_ls;
  lbfgs_step_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(reinterpret_cast<LbfgsState*>(state), xt, value, grad,
                                                                   n_eval, n_obs, theta_out, raw_out, mll_out, iters_out,
                                                                   T, o);
  B200BO_CHECK_LAUNCH("lbfgs_step");
  return 0;
}

}  // extern "C"
