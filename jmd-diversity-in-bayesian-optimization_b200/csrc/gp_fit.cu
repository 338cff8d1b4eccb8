// K-C: marginal log-likelihood value + analytic gradient, and the fused per-trial hyper-parameter fit.
//
// Replaces ExactMarginalLogLikelihood + autograd + scipy L-BFGS-B inside botorch's fit_gpytorch_model
// (Src/BO.py:144,189,300).  One trial per CTA; the threads form an OW x OW owner grid and
// thread (tr, tc) owns the matrix entries (r, c) = (tr + OW a, tc + OW b) of the lower triangle IN REGISTERS
// for the whole evaluation (block-cyclic, so the triangular work stays balanced while the sweep advances).
//
// Arithmetic: the symmetric sweep operator (Dempster / Goodnight) applied to the augmented matrix
//        [ Khat   r ]   sweep pivots 0..n-1    [ -Khat^-1        alpha       ]
//        [ r^T    0 ]  -------------------->   [  alpha^T   -r^T Khat^-1 r   ]        r = y - c
// One pivot step is a rank-1 update of every owned entry from the pivot column, which is the only thing
// exchanged through shared memory (n doubles + 1/pivot per step, double buffered: one barrier per step).
// The pivots are the Cholesky pivots L_kk^2, so log|Khat| = sum log d_k and "pivot <= 0" is exactly the
// psd_safe_cholesky failure test (jitter 1e-8 / 1e-7 / 1e-6 retries as gpytorch).  Khat^-1, alpha and the
// quadratic form come out of the same n^3/2 FMAs that Cholesky + L^-1 + L^-T L^-1 would need, without the
// three dependent triangular phases.  Pairwise distances are cached in shared memory once per fit (they do
// not depend on the hyper-parameters), so an evaluation costs two exp per entry and no sqrt.
//
// Roofline: FP64 pipe / latency; HBM traffic is 24 n B per trial per fit.
#include "lbfgs.cuh"

namespace b200bo {

// Owner grid OW x OW threads: 8 x 8 (two warps, so every trial of a 1 000-trial batch is resident at once) while
// the augmented matrix fits 64 rows, 16 x 16 beyond (up to 16 * 9 - 1 = 143 >= B200BO_NMAX observations).

__host__ __device__ constexpr int slot_of(int a, int b) { return a * (a + 1) / 2 + b; }

// Kernel value and derivative from the cached distance dc (RBF caches r^2). Every kind is poly(a) * exp(-a), so the
// unrolled per-entry code holds ONE inlined exp whatever the kind (code size is what bounds this kernel's
// instruction fetch). Values are bit-identical to kernel_from_r2 (common.cuh), i.e. to the matrix K-A factorises.
__device__ __forceinline__ double kern_arg(int kid, double dc, double ell, double a_scale) {
  if (kid == B200BO_KERNEL_RBF) return -(dc * a_scale);
  if (kid == B200BO_KERNEL_MATERN12) return dc / ell;
  return a_scale * dc;
}

__device__ __forceinline__ double kern_from_exp(int kid, double a, double e, double s) {
  if (kid == B200BO_KERNEL_MATERN32) return s * ((1.0 + a) * e);
  if (kid == B200BO_KERNEL_MATERN52) return s * ((1.0 + a + (a * a) / 3.0) * e);
  return s * e;
}

__device__ __forceinline__ double kern_only(int kid, double dc, double s, double ell, double a_scale) {
  const double a = kern_arg(kid, dc, ell, a_scale);
  return kern_from_exp(kid, a, exp(-a), s);
}

// kl = k / s ; dk = (d k / d ell) * ell / s   (the common factor s / ell is applied once to the reduced sum)
__device__ __forceinline__ void kern_pair_from_exp(int kid, double a, double e, double& kl, double& dk) {
  if (kid == B200BO_KERNEL_RBF) {
    kl = e;
    dk = (2.0 * a) * e;                      // q = r^2 / l^2 = 2 a
  } else if (kid == B200BO_KERNEL_MATERN12) {
    kl = e;
    dk = a * e;
  } else if (kid == B200BO_KERNEL_MATERN32) {
    kl = (1.0 + a) * e;
    dk = (a * a) * e;
  } else {
    kl = (1.0 + a + (a * a) / 3.0) * e;
    dk = ((a * a) / 3.0) * (1.0 + a) * e;
  }
}

__device__ __forceinline__ void kern_pair(int kid, double dc, double ell, double a_scale, double& kl, double& dk) {
  const double a = kern_arg(kid, dc, ell, a_scale);
  kern_pair_from_exp(kid, a, exp(-a), kl, dk);
}

struct FitSmem {
  double* dist;    // NS * FT   cached distances, [slot][thread]
  double* kbuf;    // NS * FT   per-entry staging between the runtime transcendental loops and the register tile
  double* ebuf;    // NS * FT   exp(-a) of the current parameters (variants with room: the gradient traces reuse it)
  double* vbuf;    // 2 * (NP + 2) pivot column (double buffered), [NP] = 1 / pivot
  double* alpha;   // NP  (alpha[NP-1] = -r^T Khat^-1 r)
  double* dpiv;    // NP  pivots
  double* xs;      // NP
  double* ys;      // NP
  double* yv;      // NP
};

// keep exp(-a) per entry between the assembly and the gradient traces where it does not cost occupancy
template <int NB, int OW>
__host__ __device__ constexpr bool fit_ecache() { return OW == 8 ? NB <= 6 : NB == 8; }

// the staging buffer is walked in NCH chunks: two for the 16 x 16 grid at NB = 6, 7 (n = 80 ... 111, the costliest fits
// of a 100-iteration run), which brings the CTA to 68 / 90 KB of shared memory so that three / two of them share an SM
template <int NB, int OW>
__host__ __device__ constexpr int fit_kchunks() { return (OW == 16 && (NB == 6 || NB == 7)) ? 2 : 1; }
template <int NB, int OW>
__host__ __device__ constexpr int fit_kchunk_slots() { return (NB * (NB + 1) / 2 + fit_kchunks<NB, OW>() - 1) / fit_kchunks<NB, OW>(); }

template <int NB, int OW>
__host__ __device__ constexpr size_t fit_smem_doubles() {
  return (size_t)(NB * (NB + 1) / 2 * (fit_ecache<NB, OW>() ? 2 : 1) + fit_kchunk_slots<NB, OW>()) * (OW * OW) +
         2 * (OW * NB + 2) + 5 * (OW * NB);
}

template <int NB, int OW>
__device__ __forceinline__ FitSmem fit_carve(double* sm) {
  constexpr int NS = NB * (NB + 1) / 2, NP = OW * NB, FT = OW * OW;
  FitSmem s;
  s.dist = sm;
  s.kbuf = s.dist + (size_t)NS * FT;
  s.ebuf = s.kbuf + (size_t)fit_kchunk_slots<NB, OW>() * FT;
  s.vbuf = s.ebuf + (fit_ecache<NB, OW>() ? (size_t)NS * FT : 0);
  s.alpha = s.vbuf + 2 * (NP + 2);
  s.dpiv = s.alpha + NP;
  s.xs = s.dpiv + NP;
  s.ys = s.xs + NP;
  s.yv = s.ys + NP;
  return s;
}

// shared scalars of one evaluation
struct FitPar {
  double noise, cst, sc, ell, a_scale, sig_s, sig_l, lp_const;
};

__device__ __forceinline__ void fit_params_from_raw(const double* raw, int kernel_id, double lp_const, FitPar& p) {
  const double raw_s = raw[2], raw_l = raw[3];
  p.noise = raw[0];
  p.cst = raw[1];
  p.sc = raw_s > 20.0 ? raw_s : log1p(exp(raw_s));        // softplus with torch's threshold (oracle/gp.py:softplus)
  p.ell = raw_l > 20.0 ? raw_l : log1p(exp(raw_l));
  p.a_scale = make_kern_params(p.sc, p.ell, kernel_id).a_scale;
  p.sig_s = raw_s >= 0 ? 1.0 / (1.0 + exp(-raw_s)) : exp(raw_s) / (1.0 + exp(raw_s));
  p.sig_l = raw_l >= 0 ? 1.0 / (1.0 + exp(-raw_l)) : exp(raw_l) / (1.0 + exp(raw_l));
  p.lp_const = lp_const;
}

// cache pairwise distances of the owned entries (sqrt(r2); r2 itself for RBF). Runtime loop over the slots: the
// transcendental code exists once, not once per slot.
template <int NB, int OW>
__device__ __forceinline__ void fit_cache_dist(const FitSmem& s, int n, int kernel_id) {
  constexpr int NS = NB * (NB + 1) / 2, FT = OW * OW;
  const int tid = threadIdx.x, tr = tid / OW, tc = tid % OW;
  int a = 0, b = 0;
#pragma unroll 1
  for (int sl = 0; sl < NS; ++sl) {
    const int r = tr + OW * a, c = tc + OW * b;
    double v = 0.0;
    if (r < n && c <= r) {
      double dx = s.xs[r] - s.xs[c], dy = s.ys[r] - s.ys[c];
      double r2 = dx * dx + dy * dy;
      v = kernel_id == B200BO_KERNEL_RBF ? r2 : sqrt(r2);
    }
    s.dist[(size_t)sl * FT + tid] = v;
    if (++b > a) { b = 0; ++a; }
  }
}

// One evaluation of value and gradient at the parameters in `par` (shared memory, read by every thread).
// Results: out[0] = value, out[1..4] = gradient wrt (noise, c, raw_s, raw_l), *st_out = status bits. Written by
// thread 0; a trailing barrier makes them visible. red: (FT/32) * 5 doubles of scratch.
template <int NB, int OW>
__device__ __forceinline__ void mll_eval_sweep(const FitSmem& s, int n, int kernel_id, const FitPar* par, double* red,
                                               double* out, int* st_out) {
  constexpr int NS = NB * (NB + 1) / 2, NP = OW * NB, FT = OW * OW;
  const int tid = threadIdx.x, tr = tid / OW, tc = tid % OW;
  const double noise = par->noise, cst = par->cst, sc = par->sc, ell = par->ell, a_scale = par->a_scale;
  double A[NS];
  int st = 0;
  bool ok = false;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    const double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6));
    if (attempt > 0) st |= B200BO_ST_JITTER;
    __syncthreads();                       // previous users of vbuf / alpha are done
    // ---- assemble the owned entries: Khat (lower), augmented row NP-1 = y - c
    {
      constexpr int NCH = fit_kchunks<NB, OW>(), CH = fit_kchunk_slots<NB, OW>();
      int a = 0, b = 0;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int lo = ch * CH, hi = (ch + 1) * CH < NS ? (ch + 1) * CH : NS;
#pragma unroll 1
        for (int sl = lo; sl < hi; ++sl) {     // runtime loop: one copy of the exp code
          const int r = tr + OW * a, c = tc + OW * b;
          double v = 0.0;
          if (r < n && c <= r) {
            if (fit_ecache<NB, OW>()) {
              const double ka = kern_arg(kernel_id, s.dist[(size_t)sl * FT + tid], ell, a_scale);
              double e;
              if (attempt == 0) {
                e = exp(-ka);
                s.ebuf[(size_t)sl * FT + tid] = e;
              } else {
                e = s.ebuf[(size_t)sl * FT + tid];
              }
              v = kern_from_exp(kernel_id, ka, e, sc);
            } else {
              v = kern_only(kernel_id, s.dist[(size_t)sl * FT + tid], sc, ell, a_scale);
            }
            if (r == c) v += noise + jitter;
          } else if (r == NP - 1 && c < n) {
            v = s.yv[c] - cst;
          }
          s.kbuf[(size_t)(sl - lo) * FT + tid] = v;
          if (++b > a) { b = 0; ++a; }
        }
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)          // own writes: no barrier needed
          if (sl >= lo && sl < hi) A[sl] = s.kbuf[(size_t)(sl - lo) * FT + tid];
      }
    }
    // ---- sweep pivots 0..n-1
    ok = true;
    int buf = 0;
#pragma unroll
    for (int kb = 0; kb < NB; ++kb) {
      const int kt_end = min(OW, n - OW * kb);
#pragma unroll 1
      for (int kt = 0; kt < kt_end; ++kt) {
        const int k = OW * kb + kt;
        double* v = s.vbuf + buf * (NP + 2);
        if (tc == kt) {                    // column k: entries (r, k), r >= k
#pragma unroll
          for (int a = kb; a < NB; ++a)
            if (a > kb || tr >= kt) v[tr + OW * a] = A[slot_of(a, kb)];
        }
        if (tr == kt) {                    // row k: entries (k, c), c < k
#pragma unroll
          for (int b = 0; b <= kb; ++b)
            if (b < kb || tc < kt) v[tc + OW * b] = A[slot_of(kb, b)];
          if (tc == kt) {
            const double d = A[slot_of(kb, kb)];
            v[NP] = 1.0 / d;
            s.dpiv[k] = d;
          }
        }
        __syncthreads();
        const double d = v[k];
        if (!(d > 0.0)) { ok = false; break; }      // uniform: every thread reads the same pivot
        const double inv = v[NP];
        double vr[NB], u[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) vr[a] = v[tr + OW * a];
#pragma unroll
        for (int b = 0; b < NB; ++b) u[b] = v[tc + OW * b] * inv;
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) A[slot_of(a, b)] = fma(-vr[a], u[b], A[slot_of(a, b)]);
        if (tc == kt) {                    // pivot column becomes v / d
#pragma unroll
          for (int a = kb; a < NB; ++a) A[slot_of(a, kb)] = vr[a] * inv;
        }
        if (tr == kt) {                    // pivot row becomes v / d, pivot itself -1 / d
#pragma unroll
          for (int b = 0; b <= kb; ++b) A[slot_of(kb, b)] = u[b];
          if (tc == kt) A[slot_of(kb, kb)] = -inv;
        }
        buf ^= 1;
      }
      if (!ok) break;
    }
  }
  if (!ok) {
    if (tid == 0) {
      *st_out = st | B200BO_ST_CHOL_FAIL;
      for (int q = 0; q < 5; ++q) out[q] = nan("");
    }
    __syncthreads();
    return;
  }
  // ---- alpha (augmented row) to shared memory
  if (tr == OW - 1) {
#pragma unroll
    for (int b = 0; b < NB; ++b) s.alpha[tc + OW * b] = A[slot_of(NB - 1, b)];
  }
  __syncthreads();
  // ---- partial sums: q1 = sum alpha, q2 = tr(alpha alpha^T - Khat^-1), q3 = sum A_ij k_ij, q4 = sum A_ij dk_ij,
  //      q5 = sum log pivots.  A = alpha alpha^T - Khat^-1; the registers hold -Khat^-1.
  double q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0, q5 = 0.0;
  if (tid < n) {
    q1 = s.alpha[tid];
    q5 = log(s.dpiv[tid]);
  }
  double ar[NB], ac[NB];
#pragma unroll
  for (int a = 0; a < NB; ++a) ar[a] = s.alpha[tr + OW * a];
#pragma unroll
  for (int b = 0; b < NB; ++b) ac[b] = s.alpha[tc + OW * b];
  {
    constexpr int NCH = fit_kchunks<NB, OW>(), CH = fit_kchunk_slots<NB, OW>();
    int ra = 0, rb = 0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int lo = ch * CH, hi = (ch + 1) * CH < NS ? (ch + 1) * CH : NS;
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          if (slot_of(a, b) < lo || slot_of(a, b) >= hi) continue;
          const int r = tr + OW * a, c = tc + OW * b;
          double wa = 0.0;
          if (r < n && c <= r) {
            const double a_rc = fma(ar[a], ac[b], A[slot_of(a, b)]);
            if (c == r) q2 += a_rc;
            wa = (c == r) ? a_rc : 2.0 * a_rc;            // symmetric: (r,c) and (c,r)
          }
          s.kbuf[(size_t)(slot_of(a, b) - lo) * FT + tid] = wa;
        }
#pragma unroll 1
      for (int sl = lo; sl < hi; ++sl) {         // runtime loop: one copy of the exp code
        const int r = tr + OW * ra, c = tc + OW * rb;
        if (r < n && c <= r) {
          double kl, dk;
          if (fit_ecache<NB, OW>()) {
            const double ka = kern_arg(kernel_id, s.dist[(size_t)sl * FT + tid], ell, a_scale);
            kern_pair_from_exp(kernel_id, ka, s.ebuf[(size_t)sl * FT + tid], kl, dk);
          } else {
            kern_pair(kernel_id, s.dist[(size_t)sl * FT + tid], ell, a_scale, kl, dk);
          }
          const double wa = s.kbuf[(size_t)(sl - lo) * FT + tid];
          q3 = fma(wa, kl, q3);
          q4 = fma(wa, dk, q4);
        }
        if (++rb > ra) { rb = 0; ++ra; }
      }
    }
  }
  double q[5] = {q1, q2, q3, q4, q5};
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    const double v = warp_sum(q[m]);
    if ((tid & 31) == 0) red[(tid >> 5) * 5 + m] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double tsum[5];
    for (int m = 0; m < 5; ++m) {
      double acc = 0.0;
      for (int w = 0; w < FT / 32; ++w) acc += red[w * 5 + m];
      tsum[m] = acc;
    }
    const double nn = (double)n;
    const double quad = -s.alpha[NP - 1];                                   // r^T Khat^-1 r
    const double ll = -0.5 * quad - 0.5 * tsum[4] - 0.5 * nn * kLog2Pi;
    const double a_prior = 1.1, rate = 0.05;
    const double lp = par->lp_const + (a_prior - 1.0) * log(noise) - rate * noise;
    out[0] = (ll + lp) / nn;
    out[1] = (0.5 * tsum[1] + (a_prior - 1.0) / noise - rate) / nn;
    out[2] = tsum[0] / nn;
    out[3] = 0.5 * tsum[2] * par->sig_s / nn;
    out[4] = 0.5 * (tsum[3] * (sc / ell)) * par->sig_l / nn;
    *st_out = st;
  }
  __syncthreads();
}

__device__ __forceinline__ double gamma_prior_const() {
  const double a_prior = 1.1, rate = 0.05;
  return a_prior * log(rate) - lgamma(a_prior);
}

template <int NB, int OW>
__device__ __forceinline__ void fit_load_trial(const FitSmem& s, const double* __restrict__ X, const double* __restrict__ y,
                                               int t, int n, int n_max, int kernel_id) {
  for (int i = threadIdx.x; i < OW * NB; i += OW * OW) {
    const bool live = i < n;
    s.xs[i] = live ? X[((size_t)t * n_max + i) * 2] : 0.0;
    s.ys[i] = live ? X[((size_t)t * n_max + i) * 2 + 1] : 0.0;
    s.yv[i] = live ? y[(size_t)t * n_max + i] : 0.0;
  }
  __syncthreads();
  fit_cache_dist<NB, OW>(s, n, kernel_id);
}

// ---- K-C: one evaluation per trial ------------------------------------------------------------------------
template <int NB, int OW>
__global__ void __launch_bounds__(OW * OW, (OW == 8 ? 1 : (NB <= 6 ? 3 : (NB <= 7 ? 2 : 1))))
mll_sweep_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                 const double* __restrict__ raw, double* __restrict__ value, double* __restrict__ grad,
                 int32_t* __restrict__ status, int n_max, int kernel_id) {
  extern __shared__ double sm[];
  const int t = blockIdx.x;
  const int n = n_arr[t];
  if (n <= 0) return;
  if (n > OW * NB - 1) {
    if (threadIdx.x == 0 && status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    return;
  }
  const FitSmem s = fit_carve<NB, OW>(sm);
  __shared__ double red[(OW * OW / 32) * 5];
  __shared__ double out[5];
  __shared__ FitPar par;
  __shared__ int sh_st;
  fit_load_trial<NB, OW>(s, X, y, t, n, n_max, kernel_id);
  if (threadIdx.x == 0) fit_params_from_raw(raw + (size_t)t * 4, kernel_id, gamma_prior_const(), par);
  __syncthreads();
  mll_eval_sweep<NB, OW>(s, n, kernel_id, &par, red, out, &sh_st);
  if (threadIdx.x == 0) {
    value[t] = out[0];
    for (int q = 0; q < 4; ++q) grad[(size_t)t * 4 + q] = out[1 + q];
    if (status && sh_st) atomicOr(&status[t], sh_st);
  }
}

// ---- fused fit: the whole L-BFGS run of one trial inside one CTA -------------------------------------------
// The CTA keeps the observations and their pairwise distances in shared memory, evaluates the marginal likelihood
// and its gradient, lets thread 0 advance the L-BFGS state machine (lbfgs.cuh) and repeats until the trial
// converges: one launch per fit, and nobody waits for the slowest trial of the batch between evaluations.
template <int NB, int OW>
__global__ void __launch_bounds__(OW * OW, (OW == 8 ? 1 : (NB <= 6 ? 3 : (NB <= 7 ? 2 : 1))))
gp_fit_sweep_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                    const double* __restrict__ raw0, int raw0_per_trial, double* __restrict__ theta_out,
                    double* __restrict__ raw_out, double* __restrict__ mll_out, int32_t* __restrict__ iters_out,
                    int32_t* __restrict__ status, int n_max, int kernel_id, LbfgsOpts o, int max_rounds,
                    const int32_t* __restrict__ order) {
  extern __shared__ double sm[];
  // CTAs are handed out in blockIdx order: with `order` (trials sorted by expected evaluation count, longest first) the
  // launch ends with its shortest fits instead of waiting for a long one that started late
  const int t = order ? order[blockIdx.x] : blockIdx.x;
  const int n = n_arr[t];
  if (n <= 0) return;
  if (n > OW * NB - 1) {
    if (threadIdx.x == 0 && status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    return;
  }
  const FitSmem s = fit_carve<NB, OW>(sm);
  __shared__ double red[(OW * OW / 32) * 5];
  __shared__ double out[5];
  __shared__ FitPar par;
  __shared__ LbfgsState lst;
  __shared__ double xt[4];
  __shared__ int sh_done, sh_st;
  fit_load_trial<NB, OW>(s, X, y, t, n, n_max, kernel_id);
  if (threadIdx.x == 0) {
    lst.hist_n = 0; lst.hist_head = 0; lst.iter = 0; lst.ls_iter = 0; lst.nfev = 0; lst.resets = 0; lst.phase = 0;
    for (int i = 0; i < 4; ++i) xt[i] = raw0[raw0_per_trial ? (size_t)t * 4 + i : i];
    xt[0] = fmax(xt[0], o.lb_noise);
    fit_params_from_raw(xt, kernel_id, gamma_prior_const(), par);
    sh_done = 0;
    sh_st = 0;
  }
  __syncthreads();
  for (int round = 0; round <= max_rounds; ++round) {
    mll_eval_sweep<NB, OW>(s, n, kernel_id, &par, red, out, &sh_st);
    if (threadIdx.x == 0) {
      LbfgsOpts oo = o;
      if (round == max_rounds) oo.max_ls = -1;          // round budget exhausted: forced finish
      double gt[4];
      for (int i = 0; i < 4; ++i) gt[i] = -out[1 + i];   // the optimiser minimises F = -MLL
      sh_done = lbfgs_advance(lst, oo, xt, -out[0], gt) ? 1 : 0;
      if (!sh_done) fit_params_from_raw(xt, kernel_id, par.lp_const, par);
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

template <int NB, int OW>
static int launch_mll(const double* X, const double* y, const int32_t* n, const double* raw, double* value, double* grad,
                      int32_t* status, int T, int n_max, int kernel_id, cudaStream_t st) {
  const size_t smem = fit_smem_doubles<NB, OW>() * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(mll_sweep_kernel<NB, OW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "mll_value_grad: smem attribute");
  mll_sweep_kernel<NB, OW><<<T, OW * OW, smem, st>>>(X, y, n, raw, value, grad, status, n_max, kernel_id);
  B200BO_CHECK_LAUNCH("mll_value_grad");
  return 0;
}

template <int NB, int OW>
static int launch_fit(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                      double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                      int n_max, int kernel_id, const LbfgsOpts& o, int max_rounds, const int32_t* order, cudaStream_t st) {
  const size_t smem = fit_smem_doubles<NB, OW>() * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(gp_fit_sweep_kernel<NB, OW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "gp_fit: smem attribute");
  gp_fit_sweep_kernel<NB, OW><<<T, OW * OW, smem, st>>>(X, y, n, raw0, raw0_per_trial, theta_out, raw_out, mll_out, iters_out,
                                               status, n_max, kernel_id, o, max_rounds, order);
  B200BO_CHECK_LAUNCH("gp_fit");
  return 0;
}

// n observations + the augmented row: variant id = NB for the 8 x 8 grid (<= 64 rows), 100 + NB for 16 x 16
static int variant_for(int n_cap) { return n_cap + 1 <= 64 ? (n_cap + 1 + 7) / 8 : 100 + (n_cap + 1 + 15) / 16; }

}  // namespace b200bo

extern "C" {

int b200bo_gp_fit(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                  double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                  int n_max, int n_hint, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter,
                  int max_ls, int max_rounds, void* stream) {
  return b200bo_gp_fit_ordered(X, y, n, raw0, raw0_per_trial, theta_out, raw_out, mll_out, iters_out, status, T, n_max,
                               n_hint, kernel_id, lb_noise, gtol, ftol, max_iter, max_ls, max_rounds, nullptr, stream);
}

int b200bo_gp_fit_ordered(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                          double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                          int n_max, int n_hint, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter,
                          int max_ls, int max_rounds, const int32_t* order, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("gp_fit: T=%d n_max=%d (max %d)", T, n_max, B200BO_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("gp_fit: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !raw0) return arg_error("gp_fit: null pointer");
  const int n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;
  LbfgsOpts o;
  o.lb_noise = lb_noise; o.gtol = gtol; o.ftol = ftol; o.c1 = 1e-4; o.max_iter = max_iter; o.max_ls = max_ls;
  cudaStream_t st = as_stream(stream);
#define B200BO_FIT_CASE(ID, NB, OW) \
  case ID: return launch_fit<NB, OW>(X, y, n, raw0, raw0_per_trial, theta_out, raw_out, mll_out, iters_out, status, T, \
                                     n_max, kernel_id, o, max_rounds, order, st);
  switch (variant_for(n_cap)) {
    B200BO_FIT_CASE(1, 1, 8) B200BO_FIT_CASE(2, 2, 8) B200BO_FIT_CASE(3, 3, 8) B200BO_FIT_CASE(4, 4, 8)
    B200BO_FIT_CASE(5, 5, 8) B200BO_FIT_CASE(6, 6, 8) B200BO_FIT_CASE(7, 7, 8) B200BO_FIT_CASE(8, 8, 8)
    B200BO_FIT_CASE(105, 5, 16) B200BO_FIT_CASE(106, 6, 16) B200BO_FIT_CASE(107, 7, 16) B200BO_FIT_CASE(108, 8, 16)
    B200BO_FIT_CASE(109, 9, 16)
  }
#undef B200BO_FIT_CASE
  return arg_error("gp_fit: n=%d unsupported", n_cap);
}

int b200bo_mll_value_grad(const double* X, const double* y, const int32_t* n, const double* raw, double* value,
                          double* grad, int32_t* status, int T, int n_max, int n_hint, int kernel_id, void* stream) {
  using namespace b200bo;
  if (T < 0 || n_max < 1 || n_max > B200BO_NMAX) return arg_error("mll_value_grad: T=%d n_max=%d (max %d)", T, n_max, B200BO_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("mll_value_grad: kernel_id=%d", kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !raw || !value || !grad) return arg_error("mll_value_grad: null pointer");
  const int n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;
  cudaStream_t st = as_stream(stream);
#define B200BO_MLL_CASE(ID, NB, OW) \
  case ID: return launch_mll<NB, OW>(X, y, n, raw, value, grad, status, T, n_max, kernel_id, st);
  switch (variant_for(n_cap)) {
    B200BO_MLL_CASE(1, 1, 8) B200BO_MLL_CASE(2, 2, 8) B200BO_MLL_CASE(3, 3, 8) B200BO_MLL_CASE(4, 4, 8)
    B200BO_MLL_CASE(5, 5, 8) B200BO_MLL_CASE(6, 6, 8) B200BO_MLL_CASE(7, 7, 8) B200BO_MLL_CASE(8, 8, 8)
    B200BO_MLL_CASE(105, 5, 16) B200BO_MLL_CASE(106, 6, 16) B200BO_MLL_CASE(107, 7, 16) B200BO_MLL_CASE(108, 8, 16)
    B200BO_MLL_CASE(109, 9, 16)
  }
#undef B200BO_MLL_CASE
  return arg_error("mll_value_grad: n=%d unsupported", n_cap);
}

}  // extern "C"
