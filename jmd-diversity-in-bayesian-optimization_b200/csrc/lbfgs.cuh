// Projected L-BFGS state machine shared by the step kernel (lbfgs.cu) and the fused fit kernel (gp_chol.cu).
#pragma once
#include "common.cuh"

namespace b200bo {

constexpr int LB_M = 10;

struct LbfgsState {
  double x[4], f, g[4];
  double d[4], t, gd;
  double S[LB_M][4], Y[LB_M][4], rho[LB_M];
  double f_prev;
  int hist_n, hist_head, iter, ls_iter, phase, nfev, resets, pad;
};

struct LbfgsOpts {
  double lb_noise;   // 1e-4
  double gtol;       // 1e-5  (projected-gradient max norm)
  double ftol;       // 2.22e-9 (relative decrease)
  double c1;         // 1e-4  Armijo
  int max_iter;
  int max_ls;
};

__device__ __forceinline__ double softplus20(double v) { return v > 20.0 ? v : log1p(exp(v)); }

__device__ __forceinline__ void project(double x[4], double lb) { x[0] = fmax(x[0], lb); }

// two-loop recursion on the free coordinates; writes d = -H g
__device__ inline void lbfgs_direction(LbfgsState& s, const bool free_[4]) {
  double q[4], al[LB_M];
  for (int i = 0; i < 4; ++i) q[i] = free_[i] ? s.g[i] : 0.0;
  for (int k = 0; k < s.hist_n; ++k) {
    int idx = (s.hist_head - 1 - k + 2 * LB_M) % LB_M;
    double a = 0.0;
    for (int i = 0; i < 4; ++i) a += (free_[i] ? s.S[idx][i] : 0.0) * q[i];
    a *= s.rho[idx];
    al[k] = a;
    for (int i = 0; i < 4; ++i) q[i] -= a * (free_[i] ? s.Y[idx][i] : 0.0);
  }
  double gamma = 1.0;
  if (s.hist_n > 0) {
    int idx = (s.hist_head - 1 + LB_M) % LB_M;
    double sy = 0.0, yy = 0.0;
    for (int i = 0; i < 4; ++i) { sy += s.S[idx][i] * s.Y[idx][i]; yy += s.Y[idx][i] * s.Y[idx][i]; }
    if (yy > 0.0 && sy > 0.0) gamma = sy / yy;
  }
  for (int i = 0; i < 4; ++i) q[i] *= gamma;
  for (int k = s.hist_n - 1; k >= 0; --k) {
    int idx = (s.hist_head - 1 - k + 2 * LB_M) % LB_M;
    double b = 0.0;
    for (int i = 0; i < 4; ++i) b += (free_[i] ? s.Y[idx][i] : 0.0) * q[i];
    b *= s.rho[idx];
    for (int i = 0; i < 4; ++i) q[i] += (al[k] - b) * (free_[i] ? s.S[idx][i] : 0.0);
  }
  for (int i = 0; i < 4; ++i) s.d[i] = free_[i] ? -q[i] : 0.0;
}

__device__ inline void new_search(LbfgsState& s, const LbfgsOpts& o, double* xt) {
  bool free_[4] = {true, true, true, true};
  if (s.x[0] <= o.lb_noise && s.g[0] > 0.0) free_[0] = false;   // at the bound, descent points outside
  lbfgs_direction(s, free_);
  double gd = 0.0;
  for (int i = 0; i < 4; ++i) gd += s.g[i] * s.d[i];
  if (!(gd < 0.0)) {                       // not a descent direction: restart from steepest descent
    s.hist_n = 0;
    s.resets += 1;
    gd = 0.0;
    for (int i = 0; i < 4; ++i) { s.d[i] = free_[i] ? -s.g[i] : 0.0; gd += s.g[i] * s.d[i]; }
  }
  s.gd = gd;
  double t = 1.0;
  if (s.hist_n == 0) {                     // first step / after a restart: unit-length step at most
    double nrm = 0.0;
    for (int i = 0; i < 4; ++i) nrm += s.d[i] * s.d[i];
    nrm = sqrt(nrm);
    t = nrm > 1.0 ? 1.0 / nrm : 1.0;
  }
  s.t = t;
  s.ls_iter = 0;
  for (int i = 0; i < 4; ++i) xt[i] = s.x[i] + t * s.d[i];
  project(xt, o.lb_noise);
}

__device__ inline bool converged(const LbfgsState& s, const LbfgsOpts& o) {
  double pg = 0.0;
  for (int i = 0; i < 4; ++i) {
    double gi = s.g[i];
    if (i == 0 && s.x[0] <= o.lb_noise && gi > 0.0) gi = 0.0;
    pg = fmax(pg, fabs(gi));
  }
  if (pg <= o.gtol) return true;
  if (s.iter > 0) {
    double den = fmax(fmax(fabs(s.f_prev), fabs(s.f)), 1.0);
    if ((s.f_prev - s.f) <= o.ftol * den) return true;
  }
  return s.iter >= o.max_iter;
}


// Advance one trial's optimiser with the objective F = -MLL and its gradient evaluated at xt.
// phase: 0 = waiting for the first evaluation, 1 = line search in progress, 2 = finished.
// Returns true when the trial is finished (xt then holds the accepted point).
__device__ inline bool lbfgs_advance(LbfgsState& s, const LbfgsOpts& o, double* xt, double ft, const double gt[4]) {
  bool finite = isfinite(ft) && isfinite(gt[0]) && isfinite(gt[1]) && isfinite(gt[2]) && isfinite(gt[3]);
  s.nfev += 1;
  bool done = false;
  if (s.phase == 0) {
    for (int i = 0; i < 4; ++i) { s.x[i] = xt[i]; s.g[i] = gt[i]; }
    s.f = ft;
    s.f_prev = ft;
    if (!finite) done = true;                 // cannot even evaluate the start: keep the start
    else if (converged(s, o)) done = true;
    else { new_search(s, o, xt); s.phase = 1; }
  } else {
    // Armijo on the projected step
    double dec = 0.0;
    for (int i = 0; i < 4; ++i) dec += s.g[i] * (xt[i] - s.x[i]);
    if (finite && ft <= s.f + o.c1 * dec) {
      double sv[4], yv[4], sy = 0.0, ss = 0.0, yy = 0.0;
      for (int i = 0; i < 4; ++i) {
        sv[i] = xt[i] - s.x[i];
        yv[i] = gt[i] - s.g[i];
        sy += sv[i] * yv[i]; ss += sv[i] * sv[i]; yy += yv[i] * yv[i];
      }
      if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
        int h = s.hist_head;
        for (int i = 0; i < 4; ++i) { s.S[h][i] = sv[i]; s.Y[h][i] = yv[i]; }
        s.rho[h] = 1.0 / sy;
        s.hist_head = (h + 1) % LB_M;
        if (s.hist_n < LB_M) s.hist_n += 1;
      }
      s.f_prev = s.f;
      s.f = ft;
      for (int i = 0; i < 4; ++i) { s.x[i] = xt[i]; s.g[i] = gt[i]; }
      s.iter += 1;
      if (converged(s, o)) done = true;
      else new_search(s, o, xt);
    } else {
      s.ls_iter += 1;
      if (s.ls_iter >= o.max_ls) {
        if (s.hist_n > 0 && s.resets < 3) {    // stale curvature: one more try from steepest descent
          s.hist_n = 0;
          s.resets += 1;
          new_search(s, o, xt);
        } else {
          done = true;
        }
      } else {
        s.t *= 0.5;
        for (int i = 0; i < 4; ++i) xt[i] = s.x[i] + s.t * s.d[i];
        project(xt, o.lb_noise);
      }
    }
  }
  if (!done && o.max_ls < 0) done = true;      // forced finish: keep the last accepted point
  if (done) {
    s.phase = 2;
    for (int i = 0; i < 4; ++i) xt[i] = s.x[i];
  }
  return done;
}

}  // namespace b200bo
