// K-C for large training sets (n up to 4096): marginal log-likelihood value + gradient and the fused L-BFGS fit,
// one fit per CTA with the n x n matrix in global memory (L2 / HBM) instead of registers.
//
// Replaces fit_gpytorch_model on the 1 010 ... 1 209-point models of the hyper-parameter extraction experiment
// (Scripts/Experiment3-2-1.py:63-87,296-320: HPextract -> BO.initialize_model(nu=1.5) -> fit_gpytorch_model), which
// produces the "fixed" hyper-parameters BOloop_fixparam consumes. 200 such fits per
// (family, seed) are independent: one CTA each, ~148 in flight.
//
// Arithmetic: the same symmetric sweep as gp_fit.cu, blocked by 64 pivots. For block K (pivots k0 .. k0+63):
//     S    = -D^-1                      D = A[K,K], swept in registers (16 x 16 owners x 4 x 4 entries)
//     P_i  = A[i,K] D^-1  (= -V_i S)    for every row i outside K           (panel, 64-wide)
//     A_ij = A_ij - P_i V_j^T           for every i, j outside K            (rank-64 update of the whole triangle)
//     A[i,K] = P_i,  A[K,K] = S
// which is what 64 consecutive single-pivot sweeps compose to. After all blocks A = -Khat^-1; pivots give log|Khat|.
// Rows n .. npad-1 are padded with the identity (pivot 1, no coupling), so every block is a full block.
// The rank-64 update runs on the FP64 MMA path (mma.sync.m8n8k4.f64, as K-B): 64 x 64 output tile per CTA pass, a warp
// owns 8 rows x 64 columns (eight independent DMMA chains), operands staged in shared memory (9 LDS.64 per 8 DMMA).
// (A first version with a 4 x 4 DFMA register tile was shared-memory bound at 21 % of the FP64 pipe.)
//
// Roofline: FP64 pipe (n^3/2 FMA per evaluation; 0.89 GFMA at n = 1210); matrix traffic 2 * 8 * npad^2/2 B per block
// step, i.e. npad/64 sweeps over an 5.9 MB triangle per evaluation (mostly L2-resident per SM-local fit).
#include "lbfgs.cuh"

namespace b200bo {

constexpr int LGB = 64;            // pivot block / tile edge
constexpr int LGT = 256;           // threads: 16 x 16 owners, 4 x 4 entries of a tile each
constexpr int LG_NMAX = 8192;

struct LargePar {
  double noise, cst, sc, ell, a_scale, sig_s, sig_l, lp_const;
};

__device__ __forceinline__ void large_params_from_raw(const double* raw, int kernel_id, double lp_const, LargePar& p) {
  const double raw_s = raw[2], raw_l = raw[3];
  p.noise = raw[0];
  p.cst = raw[1];
  p.sc = raw_s > 20.0 ? raw_s : log1p(exp(raw_s));
  p.ell = raw_l > 20.0 ? raw_l : log1p(exp(raw_l));
  p.a_scale = make_kern_params(p.sc, p.ell, kernel_id).a_scale;
  p.sig_s = raw_s >= 0 ? 1.0 / (1.0 + exp(-raw_s)) : exp(raw_s) / (1.0 + exp(raw_s));
  p.sig_l = raw_l >= 0 ? 1.0 / (1.0 + exp(-raw_l)) : exp(raw_l) / (1.0 + exp(raw_l));
  p.lp_const = lp_const;
}

// kl = k / s, dk = (dk/dell) * ell / s from the squared distance (same forms as gp_fit.cu:kern_pair)
__device__ __forceinline__ void large_kern_pair(int kid, double r2, double ell, double a_scale, double& kl, double& dk) {
  double a;
  if (kid == B200BO_KERNEL_RBF) a = -(r2 * a_scale);
  else if (kid == B200BO_KERNEL_MATERN12) a = sqrt(r2) / ell;
  else a = a_scale * sqrt(r2);
  const double e = exp(-a);
  if (kid == B200BO_KERNEL_RBF) { kl = e; dk = (2.0 * a) * e; }
  else if (kid == B200BO_KERNEL_MATERN12) { kl = e; dk = a * e; }
  else if (kid == B200BO_KERNEL_MATERN32) { kl = (1.0 + a) * e; dk = (a * a) * e; }
  else { kl = (1.0 + a + (a * a) / 3.0) * e; dk = ((a * a) / 3.0) * (1.0 + a) * e; }
}

struct LargeCtx {
  double* A;        // (npad, ld) lower triangle, row-major
  double* Vt;       // (nblk, 64, 64) panel before the update, one dense [row][k] tile per 64-row block
  double* Pt;       // (nblk, 64, 64) panel after
  double* dpiv;     // (npad) pivots
  double* B0;       // shared: 64 x 64 operand tile, row-major [row][k], leading dimension LDT
  double* B1;
  double* Ss;       // shared: S = -D^-1, full symmetric, same layout
  double* xs;       // workspace (npad): L1/L2-cached, keeps shared memory at three tiles so two fits share an SM
  double* ys;
  double* res;      // y - c
  double* alpha;
  double* vbuf;     // shared 2 x (64 + 2)
  int n, npad, nblk, ld;
};

constexpr int LDT = 68;            // operand tile leading dimension: each half-warp of a fragment load hits 16 banks pairs

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// acc -= P V^T on the FP64 MMA path. Warp w owns rows 8w .. 8w+7 of the 64 x 64 tile and all eight 8-column tiles
// (eight independent DMMA chains); lane l holds C[8w + l/4][8j + 2(l%4) + {0,1}]. Per 4-wide k-block: one A fragment
// (P[row][k], negated) feeds eight DMMAs, each with its own B fragment (V[col][k]) — 9 LDS.64 per 8 DMMA.
__device__ __forceinline__ void tile_mma(double (&acc)[8][2], const double* __restrict__ Ps, const double* __restrict__ Vs,
                                         int warp, int lane) {
  const double* pa = Ps + (8 * warp + (lane >> 2)) * LDT + (lane & 3);
  const double* pb = Vs + (lane >> 2) * LDT + (lane & 3);
#pragma unroll 4
  for (int kb = 0; kb < LGB / 4; ++kb) {
    const double a = -pa[4 * kb];
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[j][0], acc[j][1], a, pb[j * 8 * LDT + 4 * kb]);
  }
}

__device__ __forceinline__ void tile_copy_in(double* dst, const double* __restrict__ src) {   // dense 64 x 64 -> padded
  for (int e = threadIdx.x; e < LGB * LGB / 2; e += LGT) {
    const int r = e >> 5, k = (e & 31) * 2;
    const double2 v = *reinterpret_cast<const double2*>(src + r * LGB + k);
    *reinterpret_cast<double2*>(dst + r * LDT + k) = v;
  }
}

// same copy through cp.async (16 B per request, L2 -> shared memory without staging in registers): the next operand
// tile streams in behind the DMMAs of the current one
__device__ __forceinline__ void tile_copy_in_async(double* dst, const double* __restrict__ src) {
  for (int e = threadIdx.x; e < LGB * LGB / 2; e += LGT) {
    const int r = e >> 5, k = (e & 31) * 2;
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + r * LDT + k);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src + r * LGB + k));
  }
  asm volatile("cp.async.commit_group;\n" ::);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// full symmetric sweep of a 64 x 64 tile held as T[a][b] <-> (tr + 16 a, tc + 16 b); returns false on pivot <= 0
__device__ __forceinline__ bool sweep_tile(double (&T)[4][4], double* vbuf, double* dpiv_out, int tr, int tc) {
  int buf = 0;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
#pragma unroll 1
    for (int kt = 0; kt < 16; ++kt) {
      const int k = 16 * kb + kt;
      double* v = vbuf + buf * (LGB + 2);
      if (tc == kt) {
#pragma unroll
        for (int a = 0; a < 4; ++a) v[tr + 16 * a] = T[a][kb];
        if (tr == kt) {
          const double d = T[kb][kb];
          v[LGB] = 1.0 / d;
          dpiv_out[k] = d;
        }
      }
      __syncthreads();
      const double d = v[k];
      if (!(d > 0.0)) return false;           // uniform
      const double inv = v[LGB];
      double vr[4], u[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) vr[a] = v[tr + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) u[b] = v[tc + 16 * b] * inv;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) T[a][b] = fma(-vr[a], u[b], T[a][b]);
      if (tc == kt) {
#pragma unroll
        for (int a = 0; a < 4; ++a) T[a][kb] = vr[a] * inv;
      }
      if (tr == kt) {
#pragma unroll
        for (int b = 0; b < 4; ++b) T[kb][b] = u[b];
        if (tc == kt) T[kb][kb] = -inv;
      }
      buf ^= 1;
    }
  }
  return true;
}

// One evaluation. out[0] = value, out[1..4] = gradient, *st_out = status bits (thread 0 writes, barrier at the end).
__device__ void large_eval(const LargeCtx& cx, int kernel_id, const LargePar* par, double* red, double* out, int* st_out) {
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int n = cx.n, ld = cx.ld, nblk = cx.nblk;
  const double noise = par->noise, sc = par->sc, ell = par->ell, a_scale = par->a_scale;
  int st = 0;
  bool ok = false;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    const double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6));
    if (attempt > 0) st |= B200BO_ST_JITTER;
    __syncthreads();
    // ---- assemble the lower triangle (identity on the padding)
    for (int I = 0; I < nblk; ++I)
      for (int J = 0; J <= I; ++J) {
#pragma unroll 1
        for (int e = 0; e < 16; ++e) {
          const int r = I * LGB + tr + 16 * (e >> 2), c = J * LGB + tc + 16 * (e & 3);
          if (c > r) continue;
          double v;
          if (r < n) {
            const double dx = cx.xs[r] - cx.xs[c], dy = cx.ys[r] - cx.ys[c];
            double kl, dk;
            large_kern_pair(kernel_id, dx * dx + dy * dy, ell, a_scale, kl, dk);
            v = sc * kl;
            if (r == c) v += noise + jitter;
          } else {
            v = r == c ? 1.0 : 0.0;
          }
          cx.A[(size_t)r * ld + c] = v;
        }
      }
    __syncthreads();
    // ---- blocked sweep
    ok = true;
    for (int K = 0; K < nblk && ok; ++K) {
      const int k0 = K * LGB;
      double T[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int r = k0 + tr + 16 * a, c = k0 + tc + 16 * b;
          T[a][b] = r >= c ? cx.A[(size_t)r * ld + c] : cx.A[(size_t)c * ld + r];
        }
      ok = sweep_tile(T, cx.vbuf, cx.dpiv + k0, tr, tc);
      if (!ok) break;
      // S to shared memory (k-pair-major) and back to the matrix (lower part)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int r = tr + 16 * a, c = tc + 16 * b;          // S[r][c]; as operand V2[k = r][col = c]
          cx.Ss[c * LDT + r] = T[a][b];                       // B operand V[col = c][k = r] = S[r][c]
          if (c <= r) cx.A[(size_t)(k0 + r) * ld + k0 + c] = T[a][b];
        }
      __syncthreads();
      // ---- panel: V_I (old) and P_I = V_I D^-1 for every other row block
      const int warp = tid >> 5, lane = tid & 31;
      const int mr = 8 * warp + (lane >> 2), mc = 2 * (lane & 3);        // this lane's row / first column in a tile
      for (int I = 0; I < nblk; ++I) {
        if (I == K) continue;
        if (I > K) {
          for (int e = tid; e < LGB * LGB; e += LGT) {
            const int r = e >> 6, k = e & 63;
            cx.B0[r * LDT + k] = cx.A[(size_t)(I * LGB + r) * ld + k0 + k];
          }
        } else {
          for (int e = tid; e < LGB * LGB; e += LGT) {
            const int k = e >> 6, r = e & 63;
            cx.B0[r * LDT + k] = cx.A[(size_t)(k0 + k) * ld + I * LGB + r];
          }
        }
        __syncthreads();
        {
          double* vt = cx.Vt + (size_t)I * LGB * LGB;
          for (int e = tid; e < LGB * LGB; e += LGT) vt[e] = cx.B0[(e >> 6) * LDT + (e & 63)];
        }
        double acc[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
        tile_mma(acc, cx.B0, cx.Ss, warp, lane);                   // -V S = V D^-1
        double* pt = cx.Pt + (size_t)I * LGB * LGB;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * j + mc;
          *reinterpret_cast<double2*>(pt + mr * LGB + c) = make_double2(acc[j][0], acc[j][1]);
          if (I > K) {
            *reinterpret_cast<double2*>(cx.A + (size_t)(I * LGB + mr) * ld + k0 + c) = make_double2(acc[j][0], acc[j][1]);
          } else {
            cx.A[(size_t)(k0 + c) * ld + I * LGB + mr] = acc[j][0];
            cx.A[(size_t)(k0 + c + 1) * ld + I * LGB + mr] = acc[j][1];
          }
        }
        __syncthreads();
      }
      // ---- rank-64 update of every tile outside row / column block K. The V_J operand tiles are double-buffered
      //      (B1 and the S buffer, which is idle in this phase) and prefetched with cp.async.
      for (int I = 0; I < nblk; ++I) {
        if (I == K) continue;
        __syncthreads();                                   // everyone is done with B0 / the operand buffers
        tile_copy_in_async(cx.B0, cx.Pt + (size_t)I * LGB * LGB);
        int J = (K == 0) ? 1 : 0;                          // first J != K (J <= I always has one: J = I)
        tile_copy_in_async(cx.B1, cx.Vt + (size_t)J * LGB * LGB);
        int cur = 0;
        while (J <= I) {
          int Jn = J + 1;
          if (Jn == K) ++Jn;
          cp_async_wait_all();
          __syncthreads();                                 // tile J landed; the other buffer is free again
          double* vcur = cur ? cx.Ss : cx.B1;
          if (Jn <= I) tile_copy_in_async(cur ? cx.B1 : cx.Ss, cx.Vt + (size_t)Jn * LGB * LGB);
          double acc[8][2];
          double* arow = cx.A + (size_t)(I * LGB + mr) * ld + J * LGB + mc;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(arow + 8 * j);    // above the diagonal: discarded below
            acc[j][0] = v.x; acc[j][1] = v.y;
          }
          tile_mma(acc, cx.B0, vcur, warp, lane);
          if (I != J) {
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<double2*>(arow + 8 * j) = make_double2(acc[j][0], acc[j][1]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = 8 * j + mc;
              if (c <= mr) arow[8 * j] = acc[j][0];
              if (c + 1 <= mr) arow[8 * j + 1] = acc[j][1];
            }
          }
          J = Jn;
          cur ^= 1;
        }
      }
      __syncthreads();
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
  // ---- alpha = Khat^-1 (y - c) = -(A_lower + A_lower^T - diag) res : row part per warp, column part per thread
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < n; i += LGT / 32) {
    double acc = 0.0;
    for (int j = lane; j <= i; j += 32) acc = fma(cx.A[(size_t)i * ld + j], cx.res[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) cx.alpha[i] = acc;
  }
  __syncthreads();
  for (int j = tid; j < n; j += LGT) {
    double acc = 0.0;
    for (int i = j + 1; i < n; ++i) acc = fma(cx.A[(size_t)i * ld + j], cx.res[i], acc);
    cx.alpha[j] = -(cx.alpha[j] + acc);
  }
  __syncthreads();
  // ---- sums: q0 = res . alpha, q1 = sum alpha, q2 = tr(alpha alpha^T - Kinv), q3, q4 kernel traces, q5 = sum log pivots
  double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0, q5 = 0.0;
  for (int i = tid; i < n; i += LGT) {
    const double al = cx.alpha[i];
    q0 = fma(cx.res[i], al, q0);
    q1 += al;
    q5 += log(cx.dpiv[i]);
  }
  for (int I = 0; I < nblk; ++I)
    for (int J = 0; J <= I; ++J) {
#pragma unroll 1
      for (int e = 0; e < 16; ++e) {
        const int r = I * LGB + tr + 16 * (e >> 2), c = J * LGB + tc + 16 * (e & 3);
        if (r >= n || c > r) continue;
        const double a_rc = fma(cx.alpha[r], cx.alpha[c], cx.A[(size_t)r * ld + c]);
        if (r == c) q2 += a_rc;
        const double dx = cx.xs[r] - cx.xs[c], dy = cx.ys[r] - cx.ys[c];
        double kl, dk;
        large_kern_pair(kernel_id, dx * dx + dy * dy, ell, a_scale, kl, dk);
        const double wa = r == c ? a_rc : 2.0 * a_rc;
        q3 = fma(wa, kl, q3);
        q4 = fma(wa, dk, q4);
      }
    }
  double q[6] = {q0, q1, q2, q3, q4, q5};
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const double v = warp_sum(q[m]);
    if (lane == 0) red[warp * 6 + m] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double ts[6];
    for (int m = 0; m < 6; ++m) {
      double acc = 0.0;
      for (int w = 0; w < LGT / 32; ++w) acc += red[w * 6 + m];
      ts[m] = acc;
    }
    const double nn = (double)n;
    const double a_prior = 1.1, rate = 0.05;
    const double ll = -0.5 * ts[0] - 0.5 * ts[5] - 0.5 * nn * kLog2Pi;
    const double lp = par->lp_const + (a_prior - 1.0) * log(noise) - rate * noise;
    out[0] = (ll + lp) / nn;
    out[1] = (0.5 * ts[2] + (a_prior - 1.0) / noise - rate) / nn;
    out[2] = ts[1] / nn;
    out[3] = 0.5 * ts[3] * par->sig_s / nn;
    out[4] = 0.5 * (ts[4] * (sc / ell)) * par->sig_l / nn;
    *st_out = st;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LGT, 2)
gp_fit_large_kernel(const double* __restrict__ X, const double* __restrict__ y, const int32_t* __restrict__ n_arr,
                    const double* __restrict__ raw0, int raw0_per_trial, int eval_only, double* __restrict__ theta_out,
                    double* __restrict__ raw_out, double* __restrict__ mll_out, double* __restrict__ grad_out,
                    int32_t* __restrict__ iters_out, int32_t* __restrict__ status, int n_max, int npad_max, int kernel_id,
                    LbfgsOpts o, int max_rounds, double* __restrict__ ws, size_t ws_per_fit) {
  extern __shared__ double2 sm2[];
  const int t = blockIdx.x;
  const int n = n_arr[t];
  if (n <= 0) return;
  if (n > n_max) {
    if (threadIdx.x == 0 && status) atomicOr(&status[t], B200BO_ST_CHOL_FAIL);
    return;
  }
  LargeCtx cx;
  cx.n = n;
  cx.nblk = (n + LGB - 1) / LGB;
  cx.npad = cx.nblk * LGB;
  cx.ld = npad_max;
  double* w = ws + (size_t)t * ws_per_fit;
  cx.A = w;
  cx.Vt = cx.A + (size_t)npad_max * npad_max;
  cx.Pt = cx.Vt + (size_t)npad_max * LGB;
  cx.dpiv = cx.Pt + (size_t)npad_max * LGB;
  cx.B0 = reinterpret_cast<double*>(sm2);
  cx.B1 = cx.B0 + LGB * LDT;
  cx.Ss = cx.B1 + LGB * LDT;
  cx.vbuf = cx.Ss + LGB * LDT;
  cx.xs = cx.dpiv + npad_max;
  cx.ys = cx.xs + npad_max;
  cx.res = cx.ys + npad_max;
  cx.alpha = cx.res + npad_max;
  __shared__ double red[(LGT / 32) * 6];
  __shared__ double out[5];
  __shared__ LargePar par;
  __shared__ LbfgsState lst;
  __shared__ double xt[4];
  __shared__ int sh_done, sh_st;
  for (int i = threadIdx.x; i < cx.npad; i += LGT) {
    const bool live = i < n;
    cx.xs[i] = live ? X[((size_t)t * n_max + i) * 2] : 0.0;
    cx.ys[i] = live ? X[((size_t)t * n_max + i) * 2 + 1] : 0.0;
  }
  if (threadIdx.x == 0) {
    lst.hist_n = 0; lst.hist_head = 0; lst.iter = 0; lst.ls_iter = 0; lst.nfev = 0; lst.resets = 0; lst.phase = 0;
    for (int i = 0; i < 4; ++i) xt[i] = raw0[raw0_per_trial ? (size_t)t * 4 + i : i];
    if (!eval_only) xt[0] = fmax(xt[0], o.lb_noise);
    const double a_prior = 1.1, rate = 0.05;
    large_params_from_raw(xt, kernel_id, a_prior * log(rate) - lgamma(a_prior), par);
    sh_done = 0;
    sh_st = 0;
  }
  __syncthreads();
  for (int round = 0; round <= max_rounds; ++round) {
    for (int i = threadIdx.x; i < cx.npad; i += LGT) cx.res[i] = i < n ? y[(size_t)t * n_max + i] - par.cst : 0.0;
    __syncthreads();
    large_eval(cx, kernel_id, &par, red, out, &sh_st);
    if (eval_only) break;
    if (threadIdx.x == 0) {
      LbfgsOpts oo = o;
      if (round == max_rounds) oo.max_ls = -1;
      double gt[4];
      for (int i = 0; i < 4; ++i) gt[i] = -out[1 + i];
      sh_done = lbfgs_advance(lst, oo, xt, -out[0], gt) ? 1 : 0;
      if (!sh_done) large_params_from_raw(xt, kernel_id, par.lp_const, par);
    }
    __syncthreads();
    if (sh_done) break;
  }
  if (threadIdx.x == 0) {
    if (eval_only) {
      if (mll_out) mll_out[t] = out[0];
      if (grad_out) for (int i = 0; i < 4; ++i) grad_out[(size_t)t * 4 + i] = out[1 + i];
      if (status && sh_st) atomicOr(&status[t], sh_st);
    } else {
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
}

static size_t large_ws_doubles_per_fit(int n_max) {
  const size_t npad = (size_t)((n_max + LGB - 1) / LGB) * LGB;
  return npad * npad + 2 * npad * LGB + 5 * npad;
}

static int launch_large(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                        int eval_only, double* theta_out, double* raw_out, double* mll_out, double* grad_out,
                        int32_t* iters_out, int32_t* status, int T, int n_max, int kernel_id, const LbfgsOpts& o,
                        int max_rounds, void* ws, size_t ws_bytes, cudaStream_t st, const char* what) {
  if (T < 0 || n_max < 1 || n_max > LG_NMAX) return arg_error("%s: T=%d n_max=%d (max %d)", what, T, n_max, LG_NMAX);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("%s: kernel_id=%d", what, kernel_id);
  if (T == 0) return 0;
  if (!X || !y || !n || !raw0 || !ws) return arg_error("%s: null pointer", what);
  const size_t per_fit = large_ws_doubles_per_fit(n_max);
  if (ws_bytes < per_fit * sizeof(double) * (size_t)T) return arg_error("%s: workspace too small (%zu < %zu)", what, ws_bytes,
                                                                           per_fit * sizeof(double) * (size_t)T);
  const int npad = ((n_max + LGB - 1) / LGB) * LGB;
  const size_t smem = 3 * (size_t)LGB * LDT * sizeof(double) + 2 * (LGB + 2) * sizeof(double);
  if (smem > 227 * 1024) return arg_error("%s: n_max=%d needs %zu B shared memory", what, n_max, smem);
  cudaError_t e = cudaFuncSetAttribute(gp_fit_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, what);
  gp_fit_large_kernel<<<T, LGT, smem, st>>>(X, y, n, raw0, raw0_per_trial, eval_only, theta_out, raw_out, mll_out, grad_out,
                                             iters_out, status, n_max, npad, kernel_id, o, max_rounds,
                                             static_cast<double*>(ws), per_fit);
  B200BO_CHECK_LAUNCH(what);
  return 0;
}

}  // namespace b200bo

extern "C" {

size_t b200bo_gp_fit_large_workspace_bytes(int T, int n_max) {
  if (T < 0 || n_max < 1) return 0;
  return b200bo::large_ws_doubles_per_fit(n_max) * sizeof(double) * (size_t)T;
}

int b200bo_gp_fit_large(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                        double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                        int n_max, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter, int max_ls,
                        int max_rounds, void* ws, size_t ws_bytes, void* stream) {
  using namespace b200bo;
  LbfgsOpts o;
  o.lb_noise = lb_noise; o.gtol = gtol; o.ftol = ftol; o.c1 = 1e-4; o.max_iter = max_iter; o.max_ls = max_ls;
  return launch_large(X, y, n, raw0, raw0_per_trial, 0, theta_out, raw_out, mll_out, nullptr, iters_out, status, T, n_max,
                      kernel_id, o, max_rounds, ws, ws_bytes, as_stream(stream), "gp_fit_large");
}

int b200bo_mll_value_grad_large(const double* X, const double* y, const int32_t* n, const double* raw, double* value,
                                double* grad, int32_t* status, int T, int n_max, int kernel_id, void* ws, size_t ws_bytes,
                                void* stream) {
  using namespace b200bo;
  if (!value || !grad) return arg_error("mll_value_grad_large: null pointer");
  LbfgsOpts o;
  o.lb_noise = 0.0; o.gtol = 0.0; o.ftol = 0.0; o.c1 = 1e-4; o.max_iter = 0; o.max_ls = 0;
  return launch_large(X, y, n, raw, 1, 1, nullptr, nullptr, value, grad, nullptr, status, T, n_max, kernel_id, o, 0, ws,
                      ws_bytes, as_stream(stream), "mll_value_grad_large");
}

}  // extern "C"
