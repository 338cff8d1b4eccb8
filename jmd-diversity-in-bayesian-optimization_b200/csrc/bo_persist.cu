// K-F: the WHOLE fixed-hyper-parameter BO loop of a trial inside one persistent CTA.
//
// Replaces, for BOloop_fixparam-style runs (Src/BO.py:380-461: EI -> optimize_acqf -> append -> re-created model with
// the same hyper-parameters, every iteration) on a candidate lattice, the per-iteration launch sequence
// gp_chol_append -> grid_posterior_(update_)acq_argmax -> argmax_finish -> bo_update of the engine by ONE launch:
// a CTA takes a trial from a work queue and runs all of its iterations before taking the next one.
//
// Arithmetic (same posterior as the full O(n^2) kernel, csrc/grid_acq.cu; same recurrences as csrc/gp_append.cu and
// csrc/grid_inc.cu): with W = L^-1 kept row by row, appending observation m gives
//     l = W k(x_m, X_<m),  lambda^2 = k(0) + noise - l.l,  W[m][:] = (-(W^T l) / lambda, 1 / lambda),
//     z_m = W[m][:] . (y - c)
// and every candidate x keeps (mu(x) - c, s - var(x)) = (sum_i v_i z_i, sum_i v_i^2), v_i(x) = W[i][:] . k(X, x):
//     v_m(x) = sum_{j<=m} W[m][j] k(x_j, x)   ->   mu += v_m z_m,  ss += v_m^2        (one length-m dot product)
// then EI / UCB / PI, observed-candidate mask, arg-max (lowest index on ties, NaN never wins), objective look-up,
// append, best-so-far / stop test (BO.py:104,152,164-169,241) and the final padding of yall (BO.py:262-271).
//
// Why one CTA per trial: the hot loop is n x G kernel-table look-ups per iteration (G = 10^4 candidates, n = 10..110),
// 8 bytes each, and NOTHING else scales with n x G. The table (y-signed, (2Gy-1) x Gx doubles = 159 KB for the
// 100 x 100 lattice of data_gen.constructX) lives in SHARED MEMORY for the whole life of the CTA, so the loop runs at
// the shared-memory bandwidth (128 B/clk/SM = 16 look-ups/clk/SM: the roofline of this kernel) instead of L1/L2
// gather latency; the per-candidate state (160 KB per trial) is private to the CTA and stays in L2 across the
// iterations of the trial — HBM sees the trial's inputs and results only; and there are no launches or grid-wide
// barriers between iterations.
// Thread mapping: a thread owns C = 10 candidates of one lattice column, rows p, p + P, ..., p + 9P (P = ceil(Gy / 10)
// row phases); the lanes of a warp are consecutive columns of one row phase, so a warp reads a contiguous (or mirrored:
// same-address broadcast) window of one table row — conflict-free — and the row index (gy - y_j + Gy-1) is linear in the
// candidate's row, so the C look-ups of one (thread, observation) pair share ONE address computation and differ by a
// constant stride. The Gx % 32 left-over columns of all row phases are packed into the last warps (4 columns x 8 phases
// per warp on the 100-wide lattice: rows of consecutive phases are 100 doubles apart = 4 bank pairs, so those warps are
// conflict-free too).
//
// Acquisition (EI): every thread first picks the most promising of its candidates by the fp32 proxy u = (mu - best) /
// sigma and evaluates it exactly; the CTA-wide maximum of those ~1000 exact values is a threshold that some candidate
// ATTAINS, so every candidate whose upper bound  EI <= sigma phi(u) / (1 + u^2)  (u <= 0, Mills-ratio bound) lies below
// it cannot be the arg-max and is skipped; the survivors (a neighbourhood of the peak) are compacted on a per-warp
// stack and evaluated with full lanes. The arg-max is exact (ties: lowest flat index).
#include "common.cuh"
#include <limits.h>

namespace b200bo {

constexpr int BP_THREADS = 1024;
constexpr int BP_WARPS = BP_THREADS / 32;
constexpr int BP_C = 10;                 // candidate rows per thread
constexpr int BP_STACK = 64;             // survivor stack entries per warp

struct __align__(16) BpObs {             // one LDS.128 broadcast per (thread, observation)
  double w;                              // W[m][j]: the new row of L^-1
  int rowoff;                            // -cy_j * Gx * 8: byte offset of the observation's table row
  int cx;                                // lattice column of the observation
};

struct BoPersistParams {
  int Gx, Gy, G;
  double x0, y0, h;
  const double* ktabY;       // (H, 2Gy-1, Gx) y-signed kernel table
  const int32_t* ktab_id;    // (T) or null
  const int32_t* order;      // (T) processing order (sorted by table) or null
  int32_t* counter;          // work-queue head (zero before the launch)
  const float* obj;          // (S, G)
  const int32_t* surf_id;    // (T)
  double* X;                 // (T, n_max, 2)
  double* y;                 // (T, n_max)
  int32_t* n;                // (T) in: initial observations; out: observations at the end
  const double* theta;       // (T, 4)
  double* state;             // (gridDim.x, 2, Gs) per-CTA scratch: mu - c, s - var
  double* best_f;            // (T)
  double* yall;              // (T, max_iter + 2)
  int32_t* picked;           // (T, max_iter)
  int32_t* nit;              // (T)
  int32_t* status;           // (T)
  unsigned long long* stats; // (8) optional counters: [0] acquisitions, [1] exact evaluations, [2] batches
  int T, n_max, max_iter, acq_id, prune, Gs, rows_pad, P;
  double acq_param, var_floor, tol, optimal_y;
  // two-trials-per-CTA kernel (bo_fixed_loop2_kernel)
  const int32_t* slot_start; // (H + 1) first position in `order` of every table slot, or null (one slot, all trials)
  int32_t* slot_cnt;         // (H) per-slot work-queue heads (zero before the launch)
  int H;
};

__host__ __device__ inline size_t bp_align16(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline int bp_phases(int Gy) { return (Gy + BP_C - 1) / BP_C; }

struct BpSmem {
  size_t table, W, obs, kv, lv, yv, mask, stack, red_val, red_idx, ctl, thr, total;
};

__host__ __device__ inline BpSmem bp_layout(int Gx, int Gy, int n_max) {
  BpSmem L;
  const int rows_pad = Gy - 1 + bp_phases(Gy) * BP_C;
  size_t o = 0;
  L.table = o; o = bp_align16(o + (size_t)rows_pad * Gx * 8);
  L.W = o;     o = bp_align16(o + (size_t)n_max * (n_max + 1) / 2 * 8);
  L.obs = o;   o = bp_align16(o + (size_t)n_max * sizeof(BpObs));
  L.kv = o;    o = bp_align16(o + (size_t)n_max * 8);
  L.lv = o;    o = bp_align16(o + (size_t)n_max * 8);
  L.yv = o;    o = bp_align16(o + (size_t)n_max * 8);
  L.mask = o;  o = bp_align16(o + (size_t)((Gx * Gy + 31) / 32) * 4);
  L.stack = o; o = bp_align16(o + (size_t)BP_WARPS * BP_STACK * 4);
  L.red_val = o; o = bp_align16(o + BP_WARPS * 8);
  L.red_idx = o; o = bp_align16(o + BP_WARPS * 4);
  L.ctl = o;   o = bp_align16(o + 16 * 4);
  L.thr = o;   o = bp_align16(o + 16);
  L.total = o;
  return L;
}

__device__ __forceinline__ double bp_acq(int acq_id, double mean, double var, double best, double acq_param, double var_floor) {
  if (acq_id == B200BO_ACQ_UCB) return acquisition<B200BO_ACQ_UCB>(mean, var, best, acq_param, var_floor);
  if (acq_id == B200BO_ACQ_PI) return acquisition<B200BO_ACQ_PI>(mean, var, best, acq_param, var_floor);
  return acquisition<B200BO_ACQ_EI>(mean, var, best, acq_param, var_floor);
}

template <int GXC, int PC>
__global__ void __launch_bounds__(BP_THREADS, 1)
bo_fixed_loop_kernel(const BoPersistParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const BpSmem L = bp_layout(p.Gx, p.Gy, p.n_max);
  double* table = reinterpret_cast<double*>(smraw + L.table);
  double* W = reinterpret_cast<double*>(smraw + L.W);
  BpObs* obs = reinterpret_cast<BpObs*>(smraw + L.obs);
  double* kv = reinterpret_cast<double*>(smraw + L.kv);
  double* lv = reinterpret_cast<double*>(smraw + L.lv);
  double* yv = reinterpret_cast<double*>(smraw + L.yv);            // y_j - c
  unsigned* omask = reinterpret_cast<unsigned*>(smraw + L.mask);
  double* red_val = reinterpret_cast<double*>(smraw + L.red_val);
  int* red_idx = reinterpret_cast<int*>(smraw + L.red_idx);
  int* ctl = reinterpret_cast<int*>(smraw + L.ctl);                 // [0] trial, [1] picked idx, [2] flags, [3] stop
  unsigned long long* thr_bits = reinterpret_cast<unsigned long long*>(smraw + L.thr);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* stk = reinterpret_cast<int*>(smraw + L.stack) + warp * BP_STACK;
  const int Gx = GXC ? GXC : p.Gx;
  const int P = PC ? PC : p.P;
  const int Gy = p.Gy, G = p.G;
  const int rowbytes = Gx * 8;
  const int cstride = P * rowbytes;                                 // table / lattice rows of one thread are P apart
  const int tab_rows = 2 * Gy - 1;
  const int mask_words = (G + 31) >> 5;
  double* mu_s = p.state + (size_t)blockIdx.x * 2 * p.Gs;
  double* ss_s = mu_s + p.Gs;
  // zero the padding rows of the table once (read by the dead rows of the last row phase only)
  for (int e = tab_rows * Gx + tid; e < p.rows_pad * Gx; e += BP_THREADS) table[e] = 0.0;
  // candidate ownership, fixed for the life of the CTA: (row phase ph, column gx); full 32-column groups first, the
  // Gx % 32 left-over columns of all phases packed behind them
  int ph, gx;
  bool active;
  {
    const int FG = Gx >> 5, LC = Gx & 31;
    const int nfull = P * FG * 32;
    if (tid < nfull) {
      ph = warp / FG;
      gx = (warp - ph * FG) * 32 + lane;
      active = true;
    } else {
      const int q = tid - nfull;
      active = LC > 0 && q < P * LC;
      ph = active ? q / LC : 0;
      gx = active ? FG * 32 + (q - ph * LC) : 0;
    }
  }
  const int g0 = ph * Gx + gx;                                      // flat index of the thread's first candidate
  const int gstep = P * Gx;
  int loaded_slot = -1;
  // optional phase timers (B200BO_KF_STATS=1): thread 0's clock between phase boundaries, summed into stats[3..9]
  long long tk = 0;
#define BP_TICK(k) do { if (p.stats && tid == 0) { const long long now__ = clock64(); atomicAdd(&p.stats[k], (unsigned long long)(now__ - tk)); tk = now__; } } while (0)

  for (;;) {
    if (tid == 0) ctl[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int q = ctl[0];
    if (q >= p.T) break;
    const int t = p.order ? p.order[q] : q;
    const int n0 = p.n[t];
    const int slot = p.ktab_id ? p.ktab_id[t] : 0;
    if (slot != loaded_slot) {                                       // CTA-uniform
      const double* src1 = p.ktabY + (size_t)slot * tab_rows * Gx;
      if ((tab_rows * Gx) & 1) {                                     // odd table: slots are only 8-byte aligned
        for (int e = tid; e < tab_rows * Gx; e += BP_THREADS) table[e] = __ldg(src1 + e);
      } else {
        const double2* src = reinterpret_cast<const double2*>(src1);
        double2* dst = reinterpret_cast<double2*>(table);
        const int n2 = (tab_rows * Gx) >> 1;
        for (int e = tid; e < n2; e += BP_THREADS) dst[e] = __ldg(src + e);
      }
      loaded_slot = slot;
    }
    const double noise = p.theta[(size_t)t * 4], cst = p.theta[(size_t)t * 4 + 1], sc = p.theta[(size_t)t * 4 + 2];
    double* Xt = p.X + (size_t)t * p.n_max * 2;
    double* yt = p.y + (size_t)t * p.n_max;
    double* yall = p.yall + (size_t)t * (p.max_iter + 2);
    int32_t* picked = p.picked + (size_t)t * (p.max_iter > 0 ? p.max_iter : 1);
    const float* objt = p.obj + (size_t)p.surf_id[t] * G;
    for (int w = tid; w < mask_words; w += BP_THREADS) omask[w] = 0u;
    if (tid == 0) { ctl[2] = 0; ctl[3] = 0; }
    __syncthreads();
    // stage the initial set: lattice coordinates, y - c; anything off the lattice hands the trial back to the host
    for (int j = tid; j < n0 && j < p.n_max; j += BP_THREADS) {
      const double x = Xt[2 * j], yy = Xt[2 * j + 1];
      const double fx = rint((x - p.x0) / p.h), fy = rint((yy - p.y0) / p.h);
      const bool on = (__dadd_rn(p.x0, __dmul_rn(p.h, fx)) == x) && (__dadd_rn(p.y0, __dmul_rn(p.h, fy)) == yy) &&
                      fx >= 0.0 && fy >= 0.0 && fx < (double)Gx && fy < (double)Gy;
      const int xi = on ? (int)fx : 0, yi = on ? (int)fy : 0;
      obs[j].cx = xi;
      obs[j].rowoff = -yi * rowbytes;
      yv[j] = yt[j] - cst;
      if (!on) atomicOr(&ctl[2], B200BO_ST_OFF_LATTICE | B200BO_ST_FALLBACK);
    }
    if (warp == 0) {
      double b = -INFINITY;
      for (int j = lane; j < n0 && j < p.n_max; j += 32) b = fmax(b, yt[j]);
      b = warp_max(b);
      if (lane == 0) {
        p.best_f[t] = b;
        yall[0] = b;                                                 // result.yall.append(result.yopt), BO.py:146
        p.nit[t] = 0;
        red_val[0] = b;
      }
      for (int i = lane; i < p.max_iter; i += 32) picked[i] = -1;
    }
    __syncthreads();
    double best = red_val[0];
    int flags_cta = ctl[2];
    __syncthreads();
    if (n0 > p.n_max || n0 < 1) flags_cta |= B200BO_ST_FALLBACK;
    const bool runnable = p.max_iter > 0 && !(flags_cta & B200BO_ST_FALLBACK);
    int n_cur = n0;                                                  // observations known (appended or waiting)
    int it_done = 0;

    if (runnable) {
      for (int m = 0; m < n_cur; ++m) {
        // ---- append observation m: k vector against the earlier observations (table look-ups)
        if (p.stats && tid == 0) tk = clock64();
        const int cxm = obs[m].cx, rom = obs[m].rowoff;
        for (int j = tid; j < m; j += BP_THREADS) {
          // (y_m - y_j + Gy-1) * rowbytes + |x_m - x_j| * 8, with rowoff = -cy * rowbytes
          const int off = (Gy - 1) * rowbytes - rom + obs[j].rowoff + abs(cxm - obs[j].cx) * 8;
          kv[j] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(table) + off);
        }
        __syncthreads();
        // l = W k : one warp per row
        for (int i = warp; i < m; i += BP_WARPS) {
          const double* wr = W + (size_t)i * (i + 1) / 2;
          double acc = 0.0;
          for (int j = lane; j <= i; j += 32) acc = fma(wr[j], kv[j], acc);
          acc = warp_sum(acc);
          if (lane == 0) lv[i] = acc;
        }
        __syncthreads();
        // lambda^2 = k(0) + noise - l.l   (every warp computes the same bits)
        double ll = 0.0;
        for (int j = lane; j < m; j += 32) ll = fma(lv[j], lv[j], ll);
        ll = warp_sum(ll);
        const double kappa = table[(size_t)(Gy - 1) * Gx] + noise;
        const double lam2 = kappa - ll;
        if (!(lam2 > 0.0)) {                                         // what psd_safe_cholesky answers with jitter:
          flags_cta |= B200BO_ST_FALLBACK;                           // the host re-runs the trial on the classic path
          break;
        }
        const double inv_lam = 1.0 / sqrt(lam2);
        // new row: w_c = -(sum_{i>=c} W[i][c] l_i) / lambda : one thread per column (consecutive columns: conflict-free)
        if (tid < m) {
          // every thread walks the same rows (row i: consecutive columns = consecutive addresses), two chains
          double acc = 0.0, acc1 = 0.0;
          int i = tid;
          if ((m - i) & 1) { acc1 = W[(size_t)i * (i + 1) / 2 + tid] * lv[i]; ++i; }
          for (; i < m; i += 2) {
            acc = fma(W[(size_t)i * (i + 1) / 2 + tid], lv[i], acc);
            acc1 = fma(W[(size_t)(i + 1) * (i + 2) / 2 + tid], lv[i + 1], acc1);
          }
          acc += acc1;
          const double w = -acc * inv_lam;
          W[(size_t)m * (m + 1) / 2 + tid] = w;
          obs[tid].w = w;
        }
        if (tid == BP_THREADS - 1) {
          W[(size_t)m * (m + 1) / 2 + m] = inv_lam;
          obs[m].w = inv_lam;
          // the appended observation is no longer selectable
          const int gflat = (-rom / rowbytes) * Gx + cxm;
          omask[gflat >> 5] |= 1u << (gflat & 31);
          thr_bits[0] = 0ull;
        }
        __syncthreads();
        // z_m = W[m][:] . (y - c)   (every warp computes the same bits)
        double zn = 0.0;
        for (int j = lane; j <= m; j += 32) zn = fma(obs[j].w, yv[j], zn);
        zn = warp_sum(zn);
        const bool acquire = m >= n0 - 1;                             // the last initial point and every later one
        BP_TICK(3);

        // ---- fold row m into the state of the thread's candidates
        double v[BP_C];
#pragma unroll
        for (int c = 0; c < BP_C; ++c) v[c] = 0.0;
        if (active) {
          const char* tb = reinterpret_cast<const char*>(table) + (size_t)(ph + Gy - 1) * rowbytes;
#pragma unroll 2
          for (int j = 0; j <= m; ++j) {
            const int4 o = *reinterpret_cast<const int4*>(&obs[j]);   // {w.lo, w.hi, rowoff, cx}: one LDS.128
            const double ow = __hiloint2double(o.y, o.x);
            const char* a = tb + (o.z + abs(gx - o.w) * 8);
#pragma unroll
            for (int c = 0; c < BP_C; ++c)
              v[c] = fma(ow, *reinterpret_cast<const double*>(a + c * cstride), v[c]);
          }
        }
        BP_TICK(4);
        const bool ei_prune = p.acq_id == B200BO_ACQ_EI && p.prune;
        unsigned surv = 0;                                            // bit c: candidate c still needs an exact value
        float seed_u = -INFINITY;
        int seed_c = -1;
        double seed_mu = 0.0, seed_ss = 0.0;
#pragma unroll
        for (int c = 0; c < BP_C; ++c) {
          const bool live = active && (ph + c * P) < Gy;
          const int g = g0 + c * gstep;
          double mu = 0.0, ss = 0.0;
          if (live && m > 0) { mu = mu_s[g]; ss = ss_s[g]; }
          mu = fma(v[c], zn, mu);
          ss = fma(v[c], v[c], ss);
          if (live) { mu_s[g] = mu; ss_s[g] = ss; }
          if (acquire && live) {
            const bool masked = (omask[g >> 5] >> (g & 31)) & 1u;
            if (!masked) {
              surv |= 1u << c;
              if (ei_prune) {
                const float d = (float)((cst + mu) - best);
                const float u = d * rsqrtf(fmaxf((float)(sc - ss), (float)p.var_floor));
                if (u > seed_u) { seed_u = u; seed_c = c; seed_mu = mu; seed_ss = ss; }
              }
            }
          }
        }
        if (!acquire) { __syncthreads(); continue; }
        BP_TICK(5);

        double best_v = -INFINITY;
        int best_i = INT_MAX;
        int flags = 0;
        unsigned long long n_eval = 0, n_batch = 0;
        if (ei_prune) {
          // exact value of every thread's most promising candidate -> an ATTAINED threshold for the whole CTA
          if (seed_c >= 0) {
            const double aq = acquisition<B200BO_ACQ_EI>(cst + seed_mu, sc - seed_ss, best, p.acq_param, p.var_floor);
            const int g = g0 + seed_c * gstep;
            if (isnan(aq) || aq == INFINITY) flags |= B200BO_ST_NONFINITE;
            if (!isnan(aq)) { best_v = aq; best_i = g; }
            surv &= ~(1u << seed_c);
          }
          if (p.stats) { n_eval += __popc(__ballot_sync(0xffffffffu, seed_c >= 0)); n_batch += 1; }
          const double wb = warp_max(best_v);
          if (lane == 0 && wb > 0.0 && wb < INFINITY) atomicMax(thr_bits, (unsigned long long)__double_as_longlong(wb));
          __syncthreads();
          BP_TICK(6);
          const double thr = __longlong_as_double((long long)thr_bits[0]);
          if (thr > 0.0 && surv) {
            // skip x when  sigma phi(u) / (1 + u^2) < thr  (u < 0):  u^2/2 > log(phi(0) / thr) + log(sigma) - log(1 + u^2);
            // the two candidate-dependent logs in fp32 with an absolute margin (they enter a sum of magnitude >= 1)
            const double lbase = log(kInvSqrt2Pi / thr) * (1.0 + 1e-12) + 1e-12;
#pragma unroll
            for (int c = 0; c < BP_C; ++c) {
              if ((surv >> c) & 1u) {
                const int g = g0 + c * gstep;
                const double d = (cst + mu_s[g]) - best;
                const double varf = fmax(sc - ss_s[g], p.var_floor);
                const float vf = (float)varf, df = (float)d;
                const float u2 = df * df * __frcp_rn(vf);
                float corr = 0.5f * __logf(vf);
                if (u2 < 1e30f) corr -= __logf(1.0f + u2);
                corr += 1e-3f + 1e-5f * fabsf(corr);
                if (d < 0.0 && 0.5 * d * d > varf * (lbase + (double)corr)) surv &= ~(1u << c);
              }
            }
          }
        }
        BP_TICK(7);
        // survivors: compact on the warp's stack, evaluate 32 at a time with full lanes
        {
          int cnt = 0;
          auto eval_batch = [&](int first, int count) {
            if (lane < count) {
              const int g = stk[first + lane];
              const double aq = bp_acq(p.acq_id, cst + mu_s[g], sc - ss_s[g], best, p.acq_param, p.var_floor);
              if (isnan(aq) || aq == INFINITY) flags |= B200BO_ST_NONFINITE;
              if (!isnan(aq) && better(aq, g, best_v, best_i)) { best_v = aq; best_i = g; }
            }
            n_eval += count; n_batch += 1;
          };
          const unsigned any = __reduce_or_sync(0xffffffffu, surv);
#pragma unroll 1
          for (int c = 0; c < BP_C; ++c) {
            if (!((any >> c) & 1u)) continue;                         // warp-uniform
            const bool push = (surv >> c) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, push);
            if (push) stk[cnt + __popc(bal & ((1u << lane) - 1u))] = g0 + c * gstep;
            cnt += __popc(bal);
            __syncwarp();
            if (cnt >= 32) {
              eval_batch(cnt - 32, 32);
              cnt -= 32;
              __syncwarp();
            }
          }
          if (cnt > 0) eval_batch(0, cnt);
        }
        warp_argmax(best_v, best_i);
        flags = __reduce_or_sync(0xffffffffu, flags);
        if (lane == 0) {
          red_val[warp] = best_v;
          red_idx[warp] = best_i;
          if (flags) atomicOr(&ctl[2], flags);
          if (p.stats) {
            atomicAdd(&p.stats[1], n_eval);
            atomicAdd(&p.stats[2], n_batch);
            if (warp == 0) atomicAdd(&p.stats[0], 1ull);
          }
        }
        __syncthreads();
        BP_TICK(8);
        // ---- bookkeeping (BO.py:104,164-169,152,241): warp 0 finishes the arg-max, its lane 0 appends
        if (warp == 0) {
          double bv = red_val[lane];
          int bi = red_idx[lane];
          warp_argmax(bv, bi);
          if (lane == 0) {
            const int it = m - (n0 - 1);
            int stop = 0;
            if (bi == INT_MAX) {                                     // nothing selectable
              atomicOr(&ctl[2], B200BO_ST_ALL_MASKED);
              bi = -1;
              stop = 1;
            } else if (m + 1 >= p.n_max) {
              bi = -1;
              stop = 1;
            } else {
              const int gy = bi / Gx, gxx = bi - gy * Gx;
              const double xn = __dadd_rn(p.x0, __dmul_rn(p.h, (double)gxx)), yn = __dadd_rn(p.y0, __dmul_rn(p.h, (double)gy));
              const double val = (double)objt[bi];                    // torch.tensor(obj_func(new_x), dtype=float32)
              Xt[2 * (m + 1)] = xn;
              Xt[2 * (m + 1) + 1] = yn;
              yt[m + 1] = val;
              obs[m + 1].cx = gxx;
              obs[m + 1].rowoff = -gy * rowbytes;
              yv[m + 1] = val - cst;
              const double b = fmax(best, val);                       // best so far
              picked[it] = bi;
              yall[it + 1] = b;
              p.nit[t] = it + 1;
              red_val[0] = b;
              stop = (!((p.optimal_y - b) > p.tol) || (it + 1 >= p.max_iter)) ? 2 : 0;
            }
            ctl[1] = bi;
            ctl[3] = stop;
          }
        }
        __syncthreads();
        const int stop = ctl[3];
        if (ctl[1] >= 0) {
          best = red_val[0];
          n_cur = m + 2;                                             // the picked point joins the observations
          it_done = m - (n0 - 1) + 1;
        }
        __syncthreads();
        BP_TICK(9);
        if (stop) break;
      }
    }
    // ---- finalize (BO.py:262,267-271): repeat the last value to max_iter + 2 entries; publish counts and flags
    if (tid == 0) {
      flags_cta |= ctl[2];
      const bool ok = runnable && !(flags_cta & B200BO_ST_FALLBACK);
      const int k = ok ? it_done : 0;
      const double last = yall[k];
      for (int i = k + 1; i < p.max_iter + 2; ++i) yall[i] = last;
      if (ok) p.best_f[t] = best;
      p.n[t] = ok ? (n0 + it_done) : n0;
      if (!ok) p.nit[t] = 0;
      if (flags_cta) atomicOr(&p.status[t], flags_cta);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two trials per CTA. The one-trial kernel above spends about half of an iteration outside the look-up loop — the
// row append (a serial dependency chain over a 110 x 110 triangle), the state pass through L2, the exact evaluations,
// the arg-max and the bookkeeping — all latency-bound phases during which the shared-memory pipe idles. Here the CTA's
// 1024 threads form two halves of 512 that run two DIFFERENT trials with the SAME hyper-parameters out of ONE table,
// each half on its own named barrier, so the latency-bound phases of one trial overlap the look-up loop of the other.
// To make room for the second trial's L^-1 the table is the unsigned one (Gy x Gx: |dy|, |dx|; 80 KB instead of
// 159 KB): the row offset of candidate c of a thread becomes |e - c P rowbytes| + a0, ONE integer instruction
// (__sad -> VABSDIFF) per look-up. A thread owns 2 x BP_C = 20 rows of one lattice column (rows ph + c P2), folded in
// two passes of 10 so the accumulators stay in registers.
// Work distribution: trials sorted by table slot; a CTA visits the slots round-robin from a staggered start, both
// halves pull trials from the slot's queue until it is empty, then the CTA (a CTA-wide barrier) loads the next table.
// The EI prune test reads a per-candidate fp32 key kept in registers from the state pass (no second trip to L2), and
// every warp fetches the objective value of its own best candidate before the CTA-level arg-max (the winner's value is
// then already on chip).
#ifndef KF2_UNROLL
#define KF2_UNROLL 4          // observations per trip of the look-up loop
#endif
#ifndef KF2_SEED_REREAD
#define KF2_SEED_REREAD 1     // 1: the seed candidate's state is read back from L2 instead of living in four registers
#endif
constexpr int kKf2Unroll = KF2_UNROLL;
constexpr int BP2_HALF = 512;
constexpr int BP2_WARPS = BP2_HALF / 32;
#ifndef KF2_PASSES
#define KF2_PASSES 2          // the 20 rows of a thread are folded in this many passes (rows per pass in registers)
#endif
constexpr int BP2_PASSES = KF2_PASSES;
constexpr int BP2_RP = (2 * BP_C) / BP2_PASSES;   // rows per pass
constexpr int BP2_C = BP2_RP * BP2_PASSES;        // candidate rows per thread (20)

struct Bp2Smem {
  size_t table, half0, W, obs, kv, lv, yv, mask, stack, red_val, red_idx, red_obj, ctl, thr, half_bytes, total;
};

__host__ __device__ inline int bp2_phases(int Gy) { return (Gy + BP2_C - 1) / BP2_C; }

__host__ __device__ inline Bp2Smem bp2_layout(int Gx, int Gy, int n_max) {
  Bp2Smem L;
  // rows of the last row phase past Gy are read (and discarded): |row - cy| <= (P2 * BP2_C - 1)
  const int rows = bp2_phases(Gy) * BP2_C;
  size_t o = 0;
  L.table = o; o = bp_align16(o + (size_t)rows * Gx * 8);
  L.half0 = o;
  size_t h = 0;
  L.W = h;     h = bp_align16(h + (size_t)n_max * (n_max + 1) / 2 * 8);
  L.obs = h;   h = bp_align16(h + (size_t)n_max * sizeof(BpObs));
  L.kv = h;    h = bp_align16(h + (size_t)n_max * 8);
  L.lv = h;    h = bp_align16(h + (size_t)n_max * 8);
  L.yv = h;    h = bp_align16(h + (size_t)n_max * 8);
  L.mask = h;  h = bp_align16(h + (size_t)((Gx * Gy + 31) / 32) * 4);
  L.stack = h; h = bp_align16(h + (size_t)BP2_WARPS * BP_STACK * 4);
  L.red_val = h; h = bp_align16(h + BP2_WARPS * 8);
  L.red_idx = h; h = bp_align16(h + BP2_WARPS * 4);
  L.red_obj = h; h = bp_align16(h + BP2_WARPS * 4);
  L.ctl = h;   h = bp_align16(h + 16 * 4);
  L.thr = h;   h = bp_align16(h + 16);
  L.half_bytes = h;
  L.total = o + 2 * h;
  return L;
}

__device__ __forceinline__ void bp2_sync(int half) { asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "n"(BP2_HALF) : "memory"); }

template <int GXC, int PC>
__global__ void __launch_bounds__(BP_THREADS, 1)
bo_fixed_loop2_kernel(const BoPersistParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const Bp2Smem L = bp2_layout(p.Gx, p.Gy, p.n_max);
  const int half = threadIdx.x >> 9;
  const int tid = threadIdx.x & (BP2_HALF - 1), lane = tid & 31, warp = tid >> 5;
  unsigned char* hb = smraw + L.half0 + (size_t)half * L.half_bytes;
  double* table = reinterpret_cast<double*>(smraw + L.table);
  double* W = reinterpret_cast<double*>(hb + L.W);
  BpObs* obs = reinterpret_cast<BpObs*>(hb + L.obs);       // here: rowoff = cy * rowbytes (>= 0)
  double* kv = reinterpret_cast<double*>(hb + L.kv);
  double* lv = reinterpret_cast<double*>(hb + L.lv);
  double* yv = reinterpret_cast<double*>(hb + L.yv);
  unsigned* omask = reinterpret_cast<unsigned*>(hb + L.mask);
  int* stk = reinterpret_cast<int*>(hb + L.stack) + warp * BP_STACK;
  double* red_val = reinterpret_cast<double*>(hb + L.red_val);
  int* red_idx = reinterpret_cast<int*>(hb + L.red_idx);
  float* red_obj = reinterpret_cast<float*>(hb + L.red_obj);
  int* ctl = reinterpret_cast<int*>(hb + L.ctl);           // [0] queue position, [1] picked idx, [2] flags, [3] stop
  unsigned long long* thr_bits = reinterpret_cast<unsigned long long*>(hb + L.thr);
  const int Gx = GXC ? GXC : p.Gx;
  const int P = PC ? PC : p.P;                             // row phases of the 20-row ownership
  const int Gy = p.Gy, G = p.G;
  const int rowbytes = Gx * 8;
  const int cstride = P * rowbytes;
  const int mask_words = (G + 31) >> 5;
  double* mu_s = p.state + ((size_t)blockIdx.x * 2 + half) * 2 * p.Gs;
  double* ss_s = mu_s + p.Gs;
  // zero the padding rows once (rows >= Gy are only reached by the dead rows of the last row phase)
  for (int e = Gy * Gx + (int)threadIdx.x; e < P * BP2_C * Gx; e += BP_THREADS) table[e] = 0.0;
  int ph, gx;
  bool active;
  {
    const int FG = Gx >> 5, LC = Gx & 31;
    const int nfull = P * FG * 32;
    if (tid < nfull) {
      ph = warp / FG;
      gx = (warp - ph * FG) * 32 + lane;
      active = true;
    } else {
      const int q = tid - nfull;
      active = LC > 0 && q < P * LC;
      ph = active ? q / LC : 0;
      gx = active ? FG * 32 + (q - ph * LC) : 0;
    }
  }
  const int g0 = ph * Gx + gx;
  const int gstep = P * Gx;
  const int tab_rows = 2 * Gy - 1;
  long long tk = 0;
#define BP2_TICK(k) do { if (p.stats && threadIdx.x == 0) { const long long now__ = clock64(); atomicAdd(&p.stats[k], (unsigned long long)(now__ - tk)); tk = now__; } } while (0)

  const int H = p.H > 0 ? p.H : 1;
  const int slot_first = (int)(((long long)blockIdx.x * H) / gridDim.x);
  for (int si = 0; si < H; ++si) {
    int slot = slot_first + si;
    if (slot >= H) slot -= H;
    const int q_lo = p.slot_start ? p.slot_start[slot] : 0;
    const int q_hi = p.slot_start ? p.slot_start[slot + 1] : p.T;
    if (q_hi <= q_lo) continue;                                        // CTA-uniform
    __syncthreads();                                                   // both halves are done with the previous table
    {
      // the dy >= 0 half of the y-signed table: rows Gy-1 .. 2Gy-2 of the slot
      const double* src1 = p.ktabY + ((size_t)slot * tab_rows + (Gy - 1)) * Gx;
      const int cnt = Gy * Gx;
      if ((cnt & 1) || ((((size_t)slot * tab_rows + (Gy - 1)) * Gx) & 1)) {
        for (int e = threadIdx.x; e < cnt; e += BP_THREADS) table[e] = __ldg(src1 + e);
      } else {
        const double2* src = reinterpret_cast<const double2*>(src1);
        double2* dst = reinterpret_cast<double2*>(table);
        for (int e = threadIdx.x; e < (cnt >> 1); e += BP_THREADS) dst[e] = __ldg(src + e);
      }
    }
    __syncthreads();

    for (;;) {                                                         // this half's trials of the slot
      if (tid == 0) ctl[0] = atomicAdd(&p.slot_cnt[slot], 1);
      bp2_sync(half);
      const int q = q_lo + ctl[0];
      if (q >= q_hi) break;                                            // half-uniform
      const int t = p.order ? p.order[q] : q;
      const int n0 = p.n[t];
      const double noise = p.theta[(size_t)t * 4], cst = p.theta[(size_t)t * 4 + 1], sc = p.theta[(size_t)t * 4 + 2];
      double* Xt = p.X + (size_t)t * p.n_max * 2;
      double* yt = p.y + (size_t)t * p.n_max;
      double* yall = p.yall + (size_t)t * (p.max_iter + 2);
      int32_t* picked = p.picked + (size_t)t * (p.max_iter > 0 ? p.max_iter : 1);
      const float* objt = p.obj + (size_t)p.surf_id[t] * G;
      for (int w = tid; w < mask_words; w += BP2_HALF) omask[w] = 0u;
      if (tid == 0) { ctl[2] = 0; ctl[3] = 0; }
      bp2_sync(half);
      for (int j = tid; j < n0 && j < p.n_max; j += BP2_HALF) {
        const double x = Xt[2 * j], yy = Xt[2 * j + 1];
        const double fx = rint((x - p.x0) / p.h), fy = rint((yy - p.y0) / p.h);
        const bool on = (__dadd_rn(p.x0, __dmul_rn(p.h, fx)) == x) && (__dadd_rn(p.y0, __dmul_rn(p.h, fy)) == yy) &&
                        fx >= 0.0 && fy >= 0.0 && fx < (double)Gx && fy < (double)Gy;
        const int xi = on ? (int)fx : 0, yi = on ? (int)fy : 0;
        obs[j].cx = xi;
        obs[j].rowoff = yi * rowbytes;
        yv[j] = yt[j] - cst;
        if (!on) atomicOr(&ctl[2], B200BO_ST_OFF_LATTICE | B200BO_ST_FALLBACK);
      }
      if (warp == 0) {
        double b = -INFINITY;
        for (int j = lane; j < n0 && j < p.n_max; j += 32) b = fmax(b, yt[j]);
        b = warp_max(b);
        if (lane == 0) {
          p.best_f[t] = b;
          yall[0] = b;                                                 // result.yall.append(result.yopt), BO.py:146
          p.nit[t] = 0;
          red_val[0] = b;
        }
        for (int i = lane; i < p.max_iter; i += 32) picked[i] = -1;
      }
      bp2_sync(half);
      double best = red_val[0];
      int flags_cta = ctl[2];
      bp2_sync(half);
      if (n0 > p.n_max || n0 < 1) flags_cta |= B200BO_ST_FALLBACK;
      const bool runnable = p.max_iter > 0 && !(flags_cta & B200BO_ST_FALLBACK);
      int n_cur = n0;
      int it_done = 0;

      if (runnable) {
        for (int m = 0; m < n_cur; ++m) {
          if (p.stats && threadIdx.x == 0) tk = clock64();
          // ---- append observation m
          const int cxm = obs[m].cx, rom = obs[m].rowoff;
          for (int j = tid; j < m; j += BP2_HALF) {
            const int off = abs(rom - obs[j].rowoff) + abs(cxm - obs[j].cx) * 8;
            kv[j] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(table) + off);
          }
          bp2_sync(half);
          for (int i = warp; i < m; i += BP2_WARPS) {                  // l = W k : one warp per row
            const double* wr = W + (size_t)i * (i + 1) / 2;
            double acc = 0.0;
            for (int j = lane; j <= i; j += 32) acc = fma(wr[j], kv[j], acc);
            acc = warp_sum(acc);
            if (lane == 0) lv[i] = acc;
          }
          bp2_sync(half);
          double ll = 0.0;                                             // every warp computes the same bits
          for (int j = lane; j < m; j += 32) ll = fma(lv[j], lv[j], ll);
          ll = warp_sum(ll);
          const double kappa = table[0] + noise;
          const double lam2 = kappa - ll;
          if (!(lam2 > 0.0)) {                                         // psd_safe_cholesky territory: hand the trial back
            flags_cta |= B200BO_ST_FALLBACK;
            break;
          }
          const double inv_lam = 1.0 / sqrt(lam2);
          if (tid < m) {
            double acc = 0.0, acc1 = 0.0;
            int i = tid;
            if ((m - i) & 1) { acc1 = W[(size_t)i * (i + 1) / 2 + tid] * lv[i]; ++i; }
            for (; i < m; i += 2) {
              acc = fma(W[(size_t)i * (i + 1) / 2 + tid], lv[i], acc);
              acc1 = fma(W[(size_t)(i + 1) * (i + 2) / 2 + tid], lv[i + 1], acc1);
            }
            acc += acc1;
            const double w = -acc * inv_lam;
            W[(size_t)m * (m + 1) / 2 + tid] = w;
            obs[tid].w = w;
          }
          if (tid == BP2_HALF - 1) {
            W[(size_t)m * (m + 1) / 2 + m] = inv_lam;
            obs[m].w = inv_lam;
            const int gflat = (rom / rowbytes) * Gx + cxm;             // the appended observation is no longer selectable
            omask[gflat >> 5] |= 1u << (gflat & 31);
            thr_bits[0] = 0ull;
          }
          bp2_sync(half);
          double zn = 0.0;                                             // z_m = W[m][:] . (y - c)
          for (int j = lane; j <= m; j += 32) zn = fma(obs[j].w, yv[j], zn);
          zn = warp_sum(zn);
          const bool acquire = m >= n0 - 1;
          const bool ei_prune = p.acq_id == B200BO_ACQ_EI && p.prune;
          BP2_TICK(3);

          // ---- fold row m into the state of the thread's 20 candidates, two passes of 10
          unsigned surv = 0;                                           // bit c: candidate c still needs an exact value
          float key[BP2_C];                                            // EI prune keys (registers)
          float seed_u = -INFINITY;
          int seed_c = -1;
#if !KF2_SEED_REREAD
          double seed_mu = 0.0, seed_ss = 0.0;
#endif
#pragma unroll
          for (int pass = 0; pass < BP2_PASSES; ++pass) {
            double v[BP2_RP];
#pragma unroll
            for (int c = 0; c < BP2_RP; ++c) v[c] = 0.0;
            if (active) {
              const unsigned tb32 = smem_u32(table);                   // shared-window address of the table
              const int erow = (ph + pass * BP2_RP * P) * rowbytes;    // byte offset of the pass's first row
#pragma unroll kKf2Unroll
              for (int j = 0; j <= m; ++j) {
                const int4 o = *reinterpret_cast<const int4*>(&obs[j]);   // {w.lo, w.hi, cy * rowbytes, cx}
                const double ow = __hiloint2double(o.y, o.x);
                const unsigned a0 = tb32 + (unsigned)(abs(gx - o.w) * 8);
                const int e = o.z - erow;                              // (cy - first row) * rowbytes
#pragma unroll
                for (int c = 0; c < BP2_RP; ++c) {
                  double kval;                                         // table[|row_c - cy|][|gx - cx|]: VABSDIFF + LDS.64
                  asm("ld.shared.f64 %0, [%1];" : "=d"(kval) : "r"(__sad(e, c * cstride, a0)));
                  v[c] = fma(ow, kval, v[c]);
                }
              }
            }
#pragma unroll
            for (int c = 0; c < BP2_RP; ++c) {
              const int cc = pass * BP2_RP + c;
              const bool live = active && (ph + cc * P) < Gy;
              const int g = g0 + cc * gstep;
              double mu = 0.0, ss = 0.0;
              if (live && m > 0) { mu = mu_s[g]; ss = ss_s[g]; }
              mu = fma(v[c], zn, mu);
              ss = fma(v[c], v[c], ss);
              if (live) { mu_s[g] = mu; ss_s[g] = ss; }
              key[cc] = -INFINITY;
              if (acquire && live) {
                const bool masked = (omask[g >> 5] >> (g & 31)) & 1u;
                if (!masked) {
                  surv |= 1u << cc;
                  if (ei_prune) {
                    // skip x when sigma phi(u) / (1 + u^2) < thr (u < 0): u^2 / 2 - log(sigma) + log(1 + u^2) >
                    // log(phi(0) / thr); the left side in fp32 with relative + absolute margins (a lower bound)
                    const float df = (float)((cst + mu) - best);
                    const float vf = fmaxf((float)(sc - ss), (float)p.var_floor);
                    const float rv = __frcp_rn(vf);
                    const float u = df * sqrtf(rv);
#if KF2_SEED_REREAD
                    if (u > seed_u) { seed_u = u; seed_c = cc; }
#else
                    if (u > seed_u) { seed_u = u; seed_c = cc; seed_mu = mu; seed_ss = ss; }
#endif
                    if (df < 0.0f) {
                      const float u2 = df * df * rv;
                      float corr = 0.5f * __logf(vf);
                      if (u2 < 1e30f) corr -= __logf(1.0f + u2);
                      corr += 1e-3f + 1e-5f * fabsf(corr);
                      key[cc] = 0.5f * u2 * (1.0f - 1e-5f) - corr;
                    }
                  }
                }
              }
            }
          }
          BP2_TICK(4);
          if (!acquire) { bp2_sync(half); continue; }

          double best_v = -INFINITY;
          int best_i = INT_MAX;
          int flags = 0;
          unsigned long long n_eval = 0, n_batch = 0;
          if (ei_prune) {
            if (seed_c >= 0) {                                         // an ATTAINED threshold for the whole trial
              const int g = g0 + seed_c * gstep;
#if KF2_SEED_REREAD
              const double seed_mu = mu_s[g], seed_ss = ss_s[g];
#endif
              const double aq = acquisition<B200BO_ACQ_EI>(cst + seed_mu, sc - seed_ss, best, p.acq_param, p.var_floor);
              if (isnan(aq) || aq == INFINITY) flags |= B200BO_ST_NONFINITE;
              if (!isnan(aq)) { best_v = aq; best_i = g; }
              surv &= ~(1u << seed_c);
            }
            if (p.stats) { n_eval += __popc(__ballot_sync(0xffffffffu, seed_c >= 0)); n_batch += 1; }
            const double wb = warp_max(best_v);
            if (lane == 0 && wb > 0.0 && wb < INFINITY) atomicMax(thr_bits, (unsigned long long)__double_as_longlong(wb));
            bp2_sync(half);
            BP2_TICK(6);
            const double thr = __longlong_as_double((long long)thr_bits[0]);
            if (thr > 0.0 && surv) {
              // log(phi(0) / thr), rounded up into fp32
              const double lbase = log(kInvSqrt2Pi / thr) * (1.0 + 1e-12) + 1e-12;
              const float lb = __double2float_ru(lbase);
#pragma unroll
              for (int cc = 0; cc < BP2_C; ++cc)
                if (key[cc] > lb) surv &= ~(1u << cc);
            }
          }
          BP2_TICK(7);
          {
            int cnt = 0;
            auto eval_batch = [&](int first, int count) {
              if (lane < count) {
                const int g = stk[first + lane];
                const double aq = bp_acq(p.acq_id, cst + mu_s[g], sc - ss_s[g], best, p.acq_param, p.var_floor);
                if (isnan(aq) || aq == INFINITY) flags |= B200BO_ST_NONFINITE;
                if (!isnan(aq) && better(aq, g, best_v, best_i)) { best_v = aq; best_i = g; }
              }
              n_eval += count; n_batch += 1;
            };
            const unsigned any = __reduce_or_sync(0xffffffffu, surv);
#pragma unroll 1
            for (int cc = 0; cc < BP2_C; ++cc) {
              if (!((any >> cc) & 1u)) continue;                       // warp-uniform
              const bool push = (surv >> cc) & 1u;
              const unsigned bal = __ballot_sync(0xffffffffu, push);
              if (push) stk[cnt + __popc(bal & ((1u << lane) - 1u))] = g0 + cc * gstep;
              cnt += __popc(bal);
              __syncwarp();
              if (cnt >= 32) {
                eval_batch(cnt - 32, 32);
                cnt -= 32;
                __syncwarp();
              }
            }
            if (cnt > 0) eval_batch(0, cnt);
          }
          warp_argmax(best_v, best_i);
          flags = __reduce_or_sync(0xffffffffu, flags);
          if (lane == 0) {
            red_val[warp] = best_v;
            red_idx[warp] = best_i;
            red_obj[warp] = best_i != INT_MAX ? __ldg(&objt[best_i]) : 0.0f;   // fetched while the other warps finish
            if (flags) atomicOr(&ctl[2], flags);
            if (p.stats) {
              atomicAdd(&p.stats[1], n_eval);
              atomicAdd(&p.stats[2], n_batch);
              if (warp == 0) atomicAdd(&p.stats[0], 1ull);
            }
          }
          bp2_sync(half);
          BP2_TICK(8);
          // ---- bookkeeping (BO.py:104,164-169,152,241): warp 0 finishes the arg-max, its lane 0 appends
          if (warp == 0) {
            const int my_idx = lane < BP2_WARPS ? red_idx[lane] : INT_MAX;
            double bv = lane < BP2_WARPS ? red_val[lane] : -INFINITY;
            int bi = my_idx;
            warp_argmax(bv, bi);
            const unsigned who = __ballot_sync(0xffffffffu, my_idx == bi && lane < BP2_WARPS);
            if (lane == 0) {
              const int it = m - (n0 - 1);
              int stop = 0;
              if (bi == INT_MAX) {                                     // nothing selectable
                atomicOr(&ctl[2], B200BO_ST_ALL_MASKED);
                bi = -1;
                stop = 1;
              } else if (m + 1 >= p.n_max) {
                bi = -1;
                stop = 1;
              } else {
                const int gy = bi / Gx, gxx = bi - gy * Gx;
                const double xn = __dadd_rn(p.x0, __dmul_rn(p.h, (double)gxx)), yn = __dadd_rn(p.y0, __dmul_rn(p.h, (double)gy));
                const double val = (double)red_obj[__ffs(who) - 1];    // torch.tensor(obj_func(new_x), dtype=float32)
                Xt[2 * (m + 1)] = xn;
                Xt[2 * (m + 1) + 1] = yn;
                yt[m + 1] = val;
                obs[m + 1].cx = gxx;
                obs[m + 1].rowoff = gy * rowbytes;
                yv[m + 1] = val - cst;
                const double b = fmax(best, val);
                picked[it] = bi;
                yall[it + 1] = b;
                p.nit[t] = it + 1;
                red_val[0] = b;
                stop = (!((p.optimal_y - b) > p.tol) || (it + 1 >= p.max_iter)) ? 2 : 0;
              }
              ctl[1] = bi;
              ctl[3] = stop;
            }
          }
          bp2_sync(half);
          const int stop = ctl[3];
          if (ctl[1] >= 0) {
            best = red_val[0];
            n_cur = m + 2;
            it_done = m - (n0 - 1) + 1;
          }
          bp2_sync(half);
          BP2_TICK(9);
          if (stop) break;
        }
      }
      if (tid == 0) {
        flags_cta |= ctl[2];
        const bool ok = runnable && !(flags_cta & B200BO_ST_FALLBACK);
        const int k = ok ? it_done : 0;
        const double last = yall[k];
        for (int i = k + 1; i < p.max_iter + 2; ++i) yall[i] = last;
        if (ok) p.best_f[t] = best;
        p.n[t] = ok ? (n0 + it_done) : n0;
        if (!ok) p.nit[t] = 0;
        if (flags_cta) atomicOr(&p.status[t], flags_cta);
      }
      bp2_sync(half);
    }
  }
}

// y-signed table: ktabY[slot][(dy + Gy-1) Gx + dx] = s k(h sqrt(dx^2 + dy^2) / l), dx >= 0; same device function and
// operands as kernel_table_kernel / the generic path (bit-identical values)
__global__ void __launch_bounds__(256)
kernel_table_ysigned_kernel(const double* __restrict__ theta, int kernel_id, double h, int Gx, int Gy,
                            double* __restrict__ ktabY) {
  const int slot = blockIdx.y;                                       // H <= 65535 (checked by the entry)
  const double sc = theta[(size_t)slot * 4 + 2], ell = theta[(size_t)slot * 4 + 3];
  const KernParams kp = make_kern_params(sc, ell, kernel_id);
  const int H2 = 2 * Gy - 1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Gx * H2; e += gridDim.x * blockDim.x) {
    const int r = e / Gx, c = e - r * Gx;
    const double dx = h * (double)c, dy = h * (double)abs(r - (Gy - 1));
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    ktabY[(size_t)slot * Gx * H2 + e] = kernel_from_r2_dyn(kernel_id, r2, sc, ell, kp.a_scale);
  }
}

static int bp_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  return sms;
}

}  // namespace b200bo

extern "C" {

int b200bo_kernel_table_ysigned(const double* theta, int H, int kernel_id, double h, int Gx, int Gy, double* ktabY,
                                void* stream) {
  using namespace b200bo;
  if (!theta || !ktabY) return arg_error("kernel_table_ysigned: null pointer");
  if (H < 1 || H > 65535 || Gx < 1 || Gy < 1 || !(h > 0.0)) return arg_error("kernel_table_ysigned: bad sizes");
  if (kernel_id < 0 || kernel_id > 3) return arg_error("kernel_table_ysigned: kernel_id=%d", kernel_id);
  dim3 grid(64, H);
  kernel_table_ysigned_kernel<<<grid, 256, 0, as_stream(stream)>>>(theta, kernel_id, h, Gx, Gy, ktabY);
  B200BO_CHECK_LAUNCH("kernel_table_ysigned");
  return 0;
}

static int bp2_fits(int Gx, int Gy, int n_max) {
  using namespace b200bo;
  if ((long long)bp2_phases(Gy) * Gx > BP2_HALF) return 0;        // one thread of a half per (row phase, column)
  return bp2_layout(Gx, Gy, n_max).total <= 232448 ? 1 : 0;
}

int b200bo_bo_fixed_loop_supported(int Gx, int Gy, int n_max) {
  using namespace b200bo;
  if (Gx < 1 || Gy < 1 || n_max < 1 || n_max > B200BO_NMAX) return 0;
  if ((long long)Gx * Gy > (1 << 24)) return 0;
  if (bp2_fits(Gx, Gy, n_max)) return 1;
  if ((long long)bp_phases(Gy) * Gx > BP_THREADS) return 0;        // one thread per (row phase, column)
  return bp_layout(Gx, Gy, n_max).total <= 232448 ? 1 : 0;       // 227 KB of dynamic shared memory per CTA on sm_100
}

int b200bo_bo_fixed_loop_ctas(int T) {
  const int sms = b200bo::bp_sm_count();
  return T < sms ? (T < 1 ? 1 : T) : sms;
}

size_t b200bo_bo_fixed_loop_workspace_bytes(int T, int Gx, int Gy) {
  if (T <= 0 || Gx <= 0 || Gy <= 0) return 0;
  const size_t Gs = ((size_t)Gx * Gy + 15) & ~(size_t)15;
  // header | per-CTA state (two trials per CTA in the two-halves kernel) | per-slot queue heads (<= T slots)
  return 256 + (size_t)b200bo_bo_fixed_loop_ctas(T) * 4 * Gs * sizeof(double) + (size_t)T * sizeof(int32_t);
}

int b200bo_bo_fixed_loop(int Gx, int Gy, double x0, double y0, double h, const double* ktabY, const int32_t* ktab_id,
                         const int32_t* order, const int32_t* slot_start, int H, const float* obj, const int32_t* surf_id,
                         double* X, double* y, int32_t* n, const double* theta, double* best_f, double* yall, int32_t* picked,
                         int32_t* nit, int32_t* status, int T, int n_max, int max_iter, int acq_id, double acq_param,
                         double var_floor, double tol, double optimal_y, void* ws, size_t ws_bytes, void* stream) {
  using namespace b200bo;
  const char* who = "bo_fixed_loop";
  if (T < 0 || max_iter < 0 || !(h > 0.0)) return arg_error("%s: bad sizes", who);
  if (!b200bo_bo_fixed_loop_supported(Gx, Gy, n_max))
    return arg_error("%s: lattice %d x %d with n_max %d does not fit the shared-memory table", who, Gx, Gy, n_max);
  if (T == 0) return 0;
  if (!ktabY || !obj || !surf_id || !X || !y || !n || !theta || !best_f || !yall || !picked || !nit || !status || !ws)
    return arg_error("%s: null pointer", who);
  if (acq_id < 0 || acq_id > 2) return arg_error("%s: acq_id=%d", who, acq_id);
  if (acq_id == B200BO_ACQ_UCB && !(acq_param >= 0.0)) return arg_error("%s: UCB beta < 0", who);
  if (H < 1 || H > T) return arg_error("%s: H=%d table slots for T=%d trials", who, H, T);
  if (H > 1 && (!ktab_id || !order || !slot_start)) return arg_error("%s: H > 1 needs ktab_id, order and slot_start", who);
  if (ws_bytes < b200bo_bo_fixed_loop_workspace_bytes(T, Gx, Gy)) return arg_error("%s: workspace too small", who);
  cudaStream_t st = as_stream(stream);
  BoPersistParams p;
  p.Gx = Gx; p.Gy = Gy; p.G = Gx * Gy; p.x0 = x0; p.y0 = y0; p.h = h;
  p.ktabY = ktabY; p.ktab_id = ktab_id; p.order = order;
  p.counter = reinterpret_cast<int32_t*>(ws);
  static const int stats_env = [] { const char* e = getenv("B200BO_KF_STATS"); return e ? atoi(e) : 0; }();
  p.stats = stats_env ? reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + 64) : nullptr;
  p.state = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
  p.obj = obj; p.surf_id = surf_id; p.X = X; p.y = y; p.n = n; p.theta = theta;
  p.best_f = best_f; p.yall = yall; p.picked = picked; p.nit = nit; p.status = status;
  p.T = T; p.n_max = n_max; p.max_iter = max_iter; p.acq_id = acq_id;
  p.acq_param = acq_id == B200BO_ACQ_UCB ? sqrt(acq_param) : acq_param;
  p.var_floor = var_floor; p.tol = tol; p.optimal_y = optimal_y;
  static const int prune_env = [] { const char* e = getenv("B200BO_KB_PRUNE"); return e ? atoi(e) : 1; }();
  p.prune = prune_env;
  p.Gs = (int)((((size_t)Gx * Gy) + 15) & ~(size_t)15);
  const int ctas1 = b200bo_bo_fixed_loop_ctas(T);
  const size_t state_bytes = (size_t)ctas1 * 4 * p.Gs * sizeof(double);
  p.slot_start = H > 1 ? slot_start : nullptr;
  p.slot_cnt = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + 256 + state_bytes);
  p.H = H;
  cudaError_t e = cudaMemsetAsync(ws, 0, 256, st);
  if (e != cudaSuccess) return cuda_status(e, who);
  static const int dual_env = [] { const char* e = getenv("B200BO_KF_DUAL"); return e ? atoi(e) : 1; }();
  if (dual_env && bp2_fits(Gx, Gy, n_max)) {
    // two trials per CTA, unsigned table (see bo_fixed_loop2_kernel)
    e = cudaMemsetAsync(p.slot_cnt, 0, (size_t)H * sizeof(int32_t), st);
    if (e != cudaSuccess) return cuda_status(e, who);
    p.P = bp2_phases(Gy);
    p.rows_pad = p.P * BP2_C;
    const size_t smem = bp2_layout(Gx, Gy, n_max).total;
    const int sms = bp_sm_count();
    const int ctas = (T + 1) / 2 < sms ? (T + 1) / 2 : sms;
    if (Gx == 100 && p.P == 5) {
      e = cudaFuncSetAttribute(bo_fixed_loop2_kernel<100, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e, who);
      bo_fixed_loop2_kernel<100, 5><<<ctas, BP_THREADS, smem, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(bo_fixed_loop2_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e, who);
      bo_fixed_loop2_kernel<0, 0><<<ctas, BP_THREADS, smem, st>>>(p);
    }
    B200BO_CHECK_LAUNCH(who);
    return 0;
  }
  if ((long long)bp_phases(Gy) * Gx > BP_THREADS || bp_layout(Gx, Gy, n_max).total > 232448)
    return arg_error("%s: lattice %d x %d with n_max %d fits neither kernel variant", who, Gx, Gy, n_max);
  p.P = bp_phases(Gy);
  p.rows_pad = Gy - 1 + p.P * BP_C;
  const size_t smem = bp_layout(Gx, Gy, n_max).total;
  const int ctas = ctas1;
  if (Gx == 100 && p.P == 10) {
    e = cudaFuncSetAttribute(bo_fixed_loop_kernel<100, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, who);
    bo_fixed_loop_kernel<100, 10><<<ctas, BP_THREADS, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(bo_fixed_loop_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, who);
    bo_fixed_loop_kernel<0, 0><<<ctas, BP_THREADS, smem, st>>>(p);
  }
  B200BO_CHECK_LAUNCH(who);
  return 0;
}

}  // extern "C"
