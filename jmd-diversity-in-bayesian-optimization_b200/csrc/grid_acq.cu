// K-B: fused grid posterior + acquisition + arg-max (the roofline kernel of this repo).
//
// Replaces ExpectedImprovement(model, best_f) -> SingleTaskGP.posterior and the continuous
// optimize_acqf of the reference (Src/BO.py:158-161, 72-83) by a dense
// evaluation + arg-max over the candidate grid. For trial t and candidate x:
//     k*_j = s k(|x - x_j| / l),  mu = c + k*^T alpha,  var = s - | L^-1 k* |^2,
//     acq  = EI / UCB / PI (mu, max(var, var_floor), best_f)
// Nothing of size n x G ever exists in HBM: the cross-covariance tile lives in the
// B-fragment registers of mma.sync.m8n8k4.f64 (DMMA), L^-1 (lower triangular, fragment-major)
// is staged once per CTA in shared memory, and only (value, index) pairs leave the SM.
//
// Mapping: one CTA = one (trial, column chunk); a warp walks 32-column super-tiles as four
// 8-column MMA tiles. For tile rows I (8 rows) only k-blocks J <= 2I+1 (lower triangle) are
// issued. Each lane evaluates the kernel for exactly the (j, column) pairs its B fragment
// needs, so there is no redundancy and no shared-memory staging of k*.
//
// Three kernels share that mapping:
//   grid_acq_kernel<KID, ...>            generic candidates: sqrt + exp per (observation, candidate) pair;
//   grid_acq_lattice_kernel<NI, THREADS> candidates and observations on one lattice: kernel values from a table of
//                                        lattice offsets (any lattice shape);
//   grid_acq_lattice8_kernel<NI, THREADS> the production path for lattices with 8 <= Gx <= 255, Gy <= 256 (the
//                                        reference's 100 x 100 and 200 x 200): byte-packed coordinates, 4-instruction
//                                        look-ups, observed-point bit mask, EI survivors compacted per warp, ragged
//                                        last row block (tail rows by DFMA / posterior mean in the spare row).
// All three return the same arg-max; mu / var agree bit for bit except where the lattice8 comments say otherwise
// (<= 1e-13 relative, tests/test_gpu_gp.py).
//
// Roofline: FP64 pipe (DMMA and DFMA share it on B200: tools/microbench/fp64_peaks.cu).
// Algorithmic flops per (trial, launch): F_B(n, G) = G (n^2 + 17 n + 30);
// algorithmic bytes: 24 n + 8 wfrag_doubles(n) + 40 read, 16 written per CTA + 16 G grid
// bytes (L2-resident, shared by all trials).
#include "common.cuh"
#include <limits.h>
#include <stdlib.h>

// tuning switches of the lattice kernel (A/B-tested on the GPU; defaults = what measured best)
#ifndef KB_MINB_SMALL
#define KB_MINB_SMALL 5      // resident 128-thread CTAs per SM the register allocator targets while NI <= 8
#endif
#ifndef KB_MU2
#define KB_MU2 0             // 1: two interleaved accumulators for the posterior-mean dot product
#endif
#ifndef KB8_MINB_MID
#define KB8_MINB_MID 5       // byte-coordinate tile, 128-thread CTAs, 6 <= NI <= 8: CTAs/SM target (measured: 5 beats 4 by 1-3 %)
#endif
#ifndef KB8_RS_MAX_NI
#define KB8_RS_MAX_NI 4      // byte-coordinate tile: reduce-scatter epilogue over four tiles up to this many row blocks
#endif
#ifndef KB8_MROW
#define KB8_MROW 1           // byte-coordinate tile: posterior mean from a spare row of the last row block (MMA path)
#endif
#ifndef KB8_TAIL
#define KB8_TAIL 1           // byte-coordinate tile: 1..2 live rows of the last row block by DFMA instead of DMMA
#endif
#ifndef KB_ABLATE
#define KB_ABLATE 0          // timing experiments ONLY (results are wrong when non-zero): bit 0 = no table look-ups,
#endif                       // bit 1 = L^-1 fragment from a register instead of shared memory, bit 2 = no acquisition

namespace b200bo {

constexpr int GA_THREADS = 128;
static int env_int(const char* name, int dflt);
constexpr int GA_WARPS = GA_THREADS / 32;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}


// V = L^-1 K* on the FP64 MMA path. Row blocks (8 rows of L^-1) are processed in groups of up to CH:
// one accumulator chain per row block, so a warp has CH independent DMMA chains in flight (a single
// chain is latency bound: 29 cycles per dependent DMMA against a 16-cycle issue interval). Row block
// I needs the k-blocks J <= 2I+1 (lower triangle); the group issues J ascending and a chain simply
// stops when its staircase ends. I and J are compile-time so kval[] stays in registers.
template <int I, int NB, int NT, int NKV>
__device__ __forceinline__ void mma_group(const double* __restrict__ wsm, int lane, const double (&kval)[NT][NKV],
                                          double (&ss)[NT][2]) {
  double acc[NB][NT][2];
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) { acc[b][tt][0] = 0.0; acc[b][tt][1] = 0.0; }
#pragma unroll
  for (int J = 0; J <= 2 * (I + NB - 1) + 1; ++J) {
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (J <= 2 * (I + b) + 1) {
        // one shared-memory fragment of L^-1 feeds NT column tiles (halves LSU traffic per DMMA at NT = 2)
        const double a = wsm[(size_t)((I + b) * (I + b + 1) + J) * 32 + lane];
#pragma unroll
        for (int tt = 0; tt < NT; ++tt) dmma_m8n8k4(acc[b][tt][0], acc[b][tt][1], a, kval[tt][J]);
      }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b)
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) {
      ss[tt][0] = fma(acc[b][tt][0], acc[b][tt][0], ss[tt][0]);
      ss[tt][1] = fma(acc[b][tt][1], acc[b][tt][1], ss[tt][1]);
    }
}

template <int I, int CH, int NI_MAX, bool END = (I >= NI_MAX)>
struct MmaGroups {
  template <int NT, int NKV>
  static __device__ __forceinline__ void run(const double* __restrict__ wsm, int lane, int nI,
                                             const double (&kval)[NT][NKV], double (&ss)[NT][2]) {
    if (I < nI) {                       // uniform
      const int nb = nI - I;
      if (nb >= CH) mma_group<I, CH>(wsm, lane, kval, ss);
      else if (CH > 3 && nb == 3) mma_group<I, (CH > 3 ? 3 : 1)>(wsm, lane, kval, ss);
      else if (CH > 2 && nb == 2) mma_group<I, (CH > 2 ? 2 : 1)>(wsm, lane, kval, ss);
      else mma_group<I, 1>(wsm, lane, kval, ss);
      MmaGroups<I + CH, CH, NI_MAX>::run(wsm, lane, nI, kval, ss);
    }
  }
};

template <int I, int CH, int NI_MAX>
struct MmaGroups<I, CH, NI_MAX, true> {
  template <int NT, int NKV>
  static __device__ __forceinline__ void run(const double* __restrict__, int, int, const double (&)[NT][NKV],
                                             double (&)[NT][2]) {}
};

struct GridAcqParams {
  const double2* grid;      // (G) candidates
  int G;
  const double* X;          // (T, n_max, 2)
  const double* wfrag;      // (T, wstride)
  const double* alpha;      // (T, n_max)
  const int32_t* n;         // (T)
  const double* theta;      // (T, 4)
  const double* best_f;     // (T)
  const uint8_t* extra_mask;  // (G) or null
  int mask_observed;
  int acq_id;
  double acq_param;         // UCB: sqrt(beta)
  double var_floor;
  double* part_val;         // (T, chunks)
  int32_t* part_idx;        // (T, chunks)
  double* mu_out;           // (T, G) or null
  double* var_out;
  double* acq_out;
  int32_t* status;          // (T)
  int n_max;
  int n_cap;                // upper bound of n[t] over the launch (template / smem selection)
  size_t wstride;
  int chunks;
  int prune;                // EI pruning allowed (exact; see grid_acq_lattice_kernel)
  int cols_per_cta;         // multiple of 32
  // lattice variant: candidates g <-> (g % Gx, g / Gx), x = x0 + h * ix; kernel values come from
  // ktab[slot][|dy| * Gx + |dx|] (built by kernel_table_kernel with the same device function)
  int Gx, Gy;
  double x0, y0, h;
  const double* ktab;       // (H, Gy, Gx)
  const int32_t* ktab_id;   // (T) table slot per trial, or null = slot 0
};

template <int KID, int NI_MAX, int CH, int MINB, int NT>
__global__ void __launch_bounds__(GA_THREADS, MINB)
grid_acq_kernel(const GridAcqParams p) {
  extern __shared__ __align__(16) double sm[];
  const int t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - t * p.chunks;
  const int n = p.n[t];
  if (n <= 0) return;                     // finished / inactive trial (uniform per CTA)
  if (n > p.n_cap || n > 8 * NI_MAX) {    // n_hint too small: report, do not touch memory
    if (threadIdx.x == 0) {
      p.part_val[(size_t)t * p.chunks + chunk] = -INFINITY;
      p.part_idx[(size_t)t * p.chunks + chunk] = INT_MAX;
      if (p.status) atomicOr(&p.status[t], B200BO_ST_ALL_MASKED);
    }
    return;
  }
  const int nI = (n + 7) >> 3;
  const int nJ = 2 * nI;
  const int npad = (8 * nI + 31) & ~31;   // look-ups run in chunks of 8 k-blocks = 32 observations
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = nI * (nI + 1);

  double* wsm = sm;                       // nblk * 32
  double* xs = sm + (size_t)nblk * 32;    // npad
  double* ys = xs + npad;
  double* al = ys + npad;
  __shared__ double red_val[GA_WARPS];
  __shared__ int red_idx[GA_WARPS];
  __shared__ int red_flag;

  {   // stage L^-1 fragments (contiguous prefix of the trial's wfrag) and the observations
    const double2* src = reinterpret_cast<const double2*>(p.wfrag + (size_t)t * p.wstride);
    double2* dst = reinterpret_cast<double2*>(wsm);
    for (int e = threadIdx.x; e < nblk * 16; e += GA_THREADS) dst[e] = src[e];
    for (int j = threadIdx.x; j < npad; j += GA_THREADS) {
      bool v = j < n;
      xs[j] = v ? p.X[((size_t)t * p.n_max + j) * 2] : 0.0;
      ys[j] = v ? p.X[((size_t)t * p.n_max + j) * 2 + 1] : 0.0;
      al[j] = v ? p.alpha[(size_t)t * p.n_max + j] : 0.0;
    }
    if (threadIdx.x == 0) red_flag = 0;
  }
  const double cst = p.theta[(size_t)t * 4 + 1];
  const double sc = p.theta[(size_t)t * 4 + 2];
  const double ell = p.theta[(size_t)t * 4 + 3];
  const KernParams kp = make_kern_params(sc, ell, KID);
  const double best_f = p.best_f[t];
  __syncthreads();

  const int col0 = chunk * p.cols_per_cta;
  const int col_end = min(p.G, col0 + p.cols_per_cta);
  const int jrow = lane & 3;        // k-row inside a 4-block / C column pair
  const int bcol = lane >> 2;       // B column / C row

  double best_v = -INFINITY;
  int best_i = INT_MAX;
  int flags = 0;

  for (int base = col0 + warp * 32; base < col_end; base += GA_WARPS * 32) {
    double my_ss = 0.0, my_mu = 0.0;
    int my_masked = 0;
#pragma unroll 1
    for (int q = 0; q < 4; q += NT) {
      double kval[NT][2 * NI_MAX];
      double mu_part[NT];
      int masked[NT];
#pragma unroll
      for (int tt = 0; tt < NT; ++tt) {
        const int gb = min(base + 8 * (q + tt) + bcol, p.G - 1);
        mu_part[tt] = 0.0;
        masked[tt] = 0;
        const double2 gp = __ldg(&p.grid[gb]);
#pragma unroll
        for (int J = 0; J < 2 * NI_MAX; ++J) {
          kval[tt][J] = 0.0;
          if (J < nJ) {
            const int j = 4 * J + jrow;
            const double dx = gp.x - xs[j], dy = gp.y - ys[j];
            const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            double kv = kernel_from_r2<KID>(r2, sc, ell, kp.a_scale);
            const bool valid = j < n;
            kv = valid ? kv : 0.0;
            masked[tt] |= (valid && r2 == 0.0) ? 1 : 0;
            kval[tt][J] = kv;
            mu_part[tt] = fma(kv, al[j], mu_part[tt]);
          }
        }
      }
      // V = L^-1 K* on the FP64 MMA path; accumulate squared entries per C column
      double ss[NT][2];
#pragma unroll
      for (int tt = 0; tt < NT; ++tt) { ss[tt][0] = 0.0; ss[tt][1] = 0.0; }
      MmaGroups<0, CH, NI_MAX>::run(wsm, lane, nI, kval, ss);
#pragma unroll
      for (int tt = 0; tt < NT; ++tt) {
        double ss0 = ss[tt][0], ss1 = ss[tt][1], mp = mu_part[tt];
        int mk = masked[tt];
        // rows live in lane>>2: reduce over xor 4, 8, 16
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
          ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
        }
        // k-rows live in lane&3: reduce over xor 1, 2
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          mp += __shfl_xor_sync(0xffffffffu, mp, o);
          mk |= __shfl_xor_sync(0xffffffffu, mk, o);
        }
        // hand column cc = lane&7 of this tile to lane 8(q+tt) + cc
        const int cc = lane & 7;
        const double t0 = __shfl_sync(0xffffffffu, ss0, cc >> 1);
        const double t1 = __shfl_sync(0xffffffffu, ss1, cc >> 1);
        const double tm = __shfl_sync(0xffffffffu, mp, cc * 4);
        const int tk = __shfl_sync(0xffffffffu, mk, cc * 4);
        if ((lane >> 3) == q + tt) {
          my_ss = (cc & 1) ? t1 : t0;
          my_mu = tm;
          my_masked = tk;
        }
      }
    }
    const int g = base + lane;
    if (g < col_end) {
      const double mu = cst + my_mu;
      const double var = sc - my_ss;
      double a;
      if (p.acq_id == B200BO_ACQ_UCB) a = acquisition<B200BO_ACQ_UCB>(mu, var, best_f, p.acq_param, p.var_floor);
      else if (p.acq_id == B200BO_ACQ_PI) a = acquisition<B200BO_ACQ_PI>(mu, var, best_f, p.acq_param, p.var_floor);
      else a = acquisition<B200BO_ACQ_EI>(mu, var, best_f, p.acq_param, p.var_floor);
      if (p.mu_out) p.mu_out[(size_t)t * p.G + g] = mu;
      if (p.var_out) p.var_out[(size_t)t * p.G + g] = var;
      if (p.acq_out) p.acq_out[(size_t)t * p.G + g] = a;
      bool m = (p.mask_observed && my_masked) || (p.extra_mask && p.extra_mask[g]);
      if (!m) {
        if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
        if (!isnan(a) && better(a, g, best_v, best_i)) { best_v = a; best_i = g; }
      }
    }
  }
  warp_argmax(best_v, best_i);
  flags = __reduce_or_sync(0xffffffffu, flags);
  if (lane == 0) {
    red_val[warp] = best_v;
    red_idx[warp] = best_i;
    if (flags) atomicOr(&red_flag, flags);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bv = red_val[0];
    int bi = red_idx[0];
    for (int w = 1; w < GA_WARPS; ++w)
      if (better(red_val[w], red_idx[w], bv, bi)) { bv = red_val[w]; bi = red_idx[w]; }
    p.part_val[(size_t)t * p.chunks + chunk] = bv;
    p.part_idx[(size_t)t * p.chunks + chunk] = bi;
    if (red_flag && p.status) atomicOr(&p.status[t], red_flag);
  }
}

// ---------------------------------------------------------------------------------------------
// Lattice K-B, k-outer schedule (the production path for lattice candidates).
//
// Same math and fragment layout as grid_acq_kernel, different loop order: the k-blocks J run in the
// outer loop and every row block I >= J/2 accumulates into its own register pair, so
//   * only ONE cross-covariance fragment value is live at a time (no kval[] array): ~100 registers
//     instead of ~165, i.e. 16-20 resident warps per SM instead of 12;
//   * for a fixed J the DMMAs of different row blocks are independent: up to NI chains in flight;
//   * the table look-up of k-block J+2 is issued while k-blocks J, J+1 run on the MMA pipe.
// NI = ceil(n_cap / 8) is a template parameter so that all accumulators and the triangular loop
// bounds are compile-time; trials with fewer row blocks (ragged batches) take the predicated body.
template <int NI, bool EXACT>
__device__ __forceinline__ void lattice_tile(const double* __restrict__ wsm, const int* __restrict__ xyi,
                                             const double* __restrict__ al, const double* __restrict__ tab, int Gx,
                                             int gxi, int gyi, int jrow, int lane, int n, int nI, double& ss0,
                                             double& ss1, double& mu_part, int& masked) {
  constexpr int NJ = 2 * NI;
  double acc[NI][2];
#pragma unroll
  for (int I = 0; I < NI; ++I) { acc[I][0] = 0.0; acc[I][1] = 0.0; }
  auto index_of = [&](int J) {
    if (KB_ABLATE & 1) return J;
    const int pk = xyi[4 * J + jrow];
    const int dxi = abs(gxi - (pk & 0xffff)), dyi = abs(gyi - (pk >> 16));
    return dyi * Gx + dxi;
  };
  // software pipeline: table values two k-blocks ahead
  int ix0 = index_of(0), ix1 = index_of(1);
  auto look_up = [&](int ix) { return (KB_ABLATE & 1) ? (double)(ix + gxi) : __ldg(tab + ix); };
  const double a_reg = (KB_ABLATE & 2) ? wsm[lane] : 0.0;
  double kv0 = look_up(ix0), kv1 = look_up(ix1);
  mu_part = 0.0;
  double mu_odd = 0.0;
  masked = 0;
#pragma unroll
  for (int J = 0; J < NJ; ++J) {
    if (!EXACT && J >= 2 * nI) break;                      // uniform
    const double kv = kv0;
    const int ix = ix0;
    kv0 = kv1; ix0 = ix1;
    if (J + 2 < NJ && (EXACT || J + 2 < 2 * nI)) {
      ix1 = index_of(J + 2);
      kv1 = look_up(ix1);
    }
    // padded observations (j >= n) meet zero columns of L^-1 and alpha = 0: any finite value is harmless
    masked |= (ix == 0 && 4 * J + jrow < n) ? 1 : 0;       // dxi < Gx, so index 0 <=> the observed point itself
    if (KB_MU2 && (J & 1)) mu_odd = fma(kv, al[4 * J + jrow], mu_odd);
    else mu_part = fma(kv, al[4 * J + jrow], mu_part);
#pragma unroll
    for (int I = J / 2; I < NI; ++I) {
      if (EXACT || I < nI)
        dmma_m8n8k4(acc[I][0], acc[I][1], (KB_ABLATE & 2) ? __hiloint2double(__double2hiint(a_reg) ^ (I + 1), __double2loint(a_reg)) : wsm[(size_t)(I * (I + 1) + J) * 32 + lane], kv);
    }
  }
  if (KB_MU2) mu_part += mu_odd;
  ss0 = 0.0;
  ss1 = 0.0;
#pragma unroll
  for (int I = 0; I < NI; ++I) {
    ss0 = fma(acc[I][0], acc[I][0], ss0);
    ss1 = fma(acc[I][1], acc[I][1], ss1);
  }
}

template <int NI, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 128 ? (NI <= 8 ? KB_MINB_SMALL : 4) : (THREADS == 256 ? 2 : 1)))
grid_acq_lattice_kernel(const GridAcqParams p) {
  constexpr int NWARPS = THREADS / 32;
  extern __shared__ __align__(16) double sm[];
  const int t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - t * p.chunks;
  const int n = p.n[t];
  if (n <= 0) return;                     // finished / inactive trial (uniform per CTA)
  if (n > p.n_cap || n > 8 * NI) {        // n_hint too small: report, do not touch memory
    if (threadIdx.x == 0) {
      p.part_val[(size_t)t * p.chunks + chunk] = -INFINITY;
      p.part_idx[(size_t)t * p.chunks + chunk] = INT_MAX;
      if (p.status) atomicOr(&p.status[t], B200BO_ST_ALL_MASKED);
    }
    return;
  }
  const int nI = (n + 7) >> 3;
  const int npad = 8 * NI;                // look-ups may run one k-block pair ahead of nI: stage the full template size
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = nI * (nI + 1);
  double* wsm = sm;                       // NI (NI + 1) * 32
  double* al = sm + (size_t)NI * (NI + 1) * 32;
  int* xyi = reinterpret_cast<int*>(al + npad);
  __shared__ double red_val[NWARPS];
  __shared__ int red_idx[NWARPS];
  __shared__ int red_flag;
  __shared__ __align__(8) unsigned long long stage_bar;
  {
    // L^-1 fragments of this trial: one contiguous, 256-byte granular range (prefix property of the fragment layout)
    // -> ONE bulk copy through the TMA unit (cp.async.bulk + mbarrier), issued by thread 0 while the other threads
    // decode the observation coordinates
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) tma_bulk_g2s(wsm, p.wfrag + (size_t)t * p.wstride, (unsigned)nblk * 256u, &stage_bar);
    for (int j = threadIdx.x; j < npad; j += THREADS) {
      const bool v = j < n;
      const double x = v ? p.X[((size_t)t * p.n_max + j) * 2] : p.x0;
      const double y = v ? p.X[((size_t)t * p.n_max + j) * 2 + 1] : p.y0;
      al[j] = v ? p.alpha[(size_t)t * p.n_max + j] : 0.0;
      const double fx = rint((x - p.x0) / p.h), fy = rint((y - p.y0) / p.h);
      const bool on = (__dadd_rn(p.x0, __dmul_rn(p.h, fx)) == x) && (__dadd_rn(p.y0, __dmul_rn(p.h, fy)) == y) && fx >= 0.0 && fy >= 0.0 &&
                      fx < (double)p.Gx && fy < (double)p.Gy;
      xyi[j] = on ? ((int)fx | ((int)fy << 16)) : 0;
      if (v && !on && p.status) atomicOr(&p.status[t], B200BO_ST_OFF_LATTICE);
    }
    if (threadIdx.x == 0) red_flag = 0;
  }
  const double* tab = p.ktab + (size_t)(p.ktab_id ? p.ktab_id[t] : 0) * p.Gx * p.Gy;
  const double cst = p.theta[(size_t)t * 4 + 1];
  const double sc = p.theta[(size_t)t * 4 + 2];
  const double best_f = p.best_f[t];
  __syncthreads();
  mbar_wait(&stage_bar, 0);               // the staged fragments have landed

  const int col0 = chunk * p.cols_per_cta;
  const int col_end = min(p.G, col0 + p.cols_per_cta);
  const int jrow = lane & 3, bcol = lane >> 2;
  double best_v = -INFINITY;
  int best_i = INT_MAX;
  int flags = 0;
  const bool prune = p.acq_id == B200BO_ACQ_EI && p.acq_out == nullptr && p.prune;
  double wbest = 0.0, lthr = INFINITY;

  for (int base = col0 + warp * 32; base < col_end; base += NWARPS * 32) {
    double my_ss = 0.0, my_mu = 0.0;
    int my_masked = 0;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int gb = min(base + 8 * q + bcol, p.G - 1);
      const int gyi = gb / p.Gx, gxi = gb - gyi * p.Gx;
      double ss0, ss1, mp;
      int mk;
      if (nI == NI) lattice_tile<NI, true>(wsm, xyi, al, tab, p.Gx, gxi, gyi, jrow, lane, n, nI, ss0, ss1, mp, mk);
      else lattice_tile<NI, false>(wsm, xyi, al, tab, p.Gx, gxi, gyi, jrow, lane, n, nI, ss0, ss1, mp, mk);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {                  // rows live in lane>>2
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
      }
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {                   // k-rows live in lane&3
        mp += __shfl_xor_sync(0xffffffffu, mp, o);
        mk |= __shfl_xor_sync(0xffffffffu, mk, o);
      }
      const int cc = lane & 7;                            // hand column cc of this tile to lane 8q + cc
      const double t0 = __shfl_sync(0xffffffffu, ss0, cc >> 1);
      const double t1 = __shfl_sync(0xffffffffu, ss1, cc >> 1);
      const double tm = __shfl_sync(0xffffffffu, mp, cc * 4);
      const int tk = __shfl_sync(0xffffffffu, mk, cc * 4);
      if ((lane >> 3) == q) {
        my_ss = (cc & 1) ? t1 : t0;
        my_mu = tm;
        my_masked = tk;
      }
    }
    const int g = base + lane;
    // EI pruning (exact): for mu < best_f, EI = sigma phi(u) + (mu - best_f) Phi(u) <= 0.39894 sqrt(s) exp(-u^2/2)
    // because the posterior sigma never exceeds the prior sqrt(s). A candidate whose bound is below the warp's
    // running maximum can neither win nor tie, so its sqrt / exp / erfc (~160 FP64-pipe slots) are skipped;
    // the test itself is two multiplies. Off when the caller asked for the acquisition field.
    if (prune) {
      const double nb = warp_max(best_v);
      if (nb > wbest) {                                   // uniform
        wbest = nb;
        lthr = log(kInvSqrt2Pi * sqrt(sc) / nb) * (1.0 + 1e-12) + 1e-12;
      }
    }
    if (g < col_end) {
      const double mu = cst + my_mu;
      const double var = sc - my_ss;
      const bool m = (p.mask_observed && my_masked) || (p.extra_mask && p.extra_mask[g]);
      if (p.mu_out) p.mu_out[(size_t)t * p.G + g] = mu;
      if (p.var_out) p.var_out[(size_t)t * p.G + g] = var;
      const double d = mu - best_f;
      const bool skip = (KB_ABLATE & 4) ? (var > -1e300) : (prune && wbest > 0.0 && d < 0.0 && 0.5 * d * d > fmax(var, p.var_floor) * lthr);
      if (!skip) {
        double a;
        if (p.acq_id == B200BO_ACQ_UCB) a = acquisition<B200BO_ACQ_UCB>(mu, var, best_f, p.acq_param, p.var_floor);
        else if (p.acq_id == B200BO_ACQ_PI) a = acquisition<B200BO_ACQ_PI>(mu, var, best_f, p.acq_param, p.var_floor);
        else a = acquisition<B200BO_ACQ_EI>(mu, var, best_f, p.acq_param, p.var_floor);
        if (p.acq_out) p.acq_out[(size_t)t * p.G + g] = a;
        if (!m) {
          if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
          if (!isnan(a) && better(a, g, best_v, best_i)) { best_v = a; best_i = g; }
        }
      }
    }
  }
  warp_argmax(best_v, best_i);
  flags = __reduce_or_sync(0xffffffffu, flags);
  if (lane == 0) {
    red_val[warp] = best_v;
    red_idx[warp] = best_i;
    if (flags) atomicOr(&red_flag, flags);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bv = red_val[0];
    int bi = red_idx[0];
    for (int w = 1; w < NWARPS; ++w)
      if (better(red_val[w], red_idx[w], bv, bi)) { bv = red_val[w]; bi = red_idx[w]; }
    p.part_val[(size_t)t * p.chunks + chunk] = bv;
    p.part_idx[(size_t)t * p.chunks + chunk] = bi;
    if (red_flag && p.status) atomicOr(&p.status[t], red_flag);
  }
}

// ---------------------------------------------------------------------------------------------
// Lattice K-B, byte-coordinate variant (production path when 8 <= Gx <= 255 and Gy <= 256, i.e. every lattice the
// reference uses: 100x100 and 200x200).
//
// Ablation on the GPU (profiles/README.md, "K-B ablation") showed that the time of a column tile is, to within 1-2 %,
//     16 cycles x #DMMA  +  1 cycle x #other warp-instructions        (per SM sub-partition)
// i.e. instruction issue does not overlap with the FP64 MMA pipe, shared-memory bandwidth is NOT a limiter, and every
// integer instruction around the MMAs costs pipe time. The generic-lattice tile above spends ~18 instructions per table
// look-up (unpack, two subtracts, two IABS, IMAD, 4-instruction 64-bit address, mask compares, alpha LDS + DFMA).
// Here:
//   * coordinates are packed one byte each (x | y << 8): |dx|, |dy| come from ONE VABSDIFF4 and the table index
//     |dy| Gx + |dx| from ONE IDP.4A against the byte weights (1, Gx) — then IMAD.WIDE + LDG: 4 instructions;
//   * observations and alpha are stored per k-row (lane & 3), so one LDS.128 fetches the coordinates of four k-blocks
//     and one LDS.128 two alpha values;
//   * "candidate is an observed point" comes from a per-CTA bit mask built while staging (one broadcast LDS per 32
//     candidates) instead of two compares per look-up and a shuffle reduction per tile;
//   * the candidate's (gx, gy) is divided out once per 32 candidates and stepped by 8 per tile.
// Same table, same indices, same summation order as lattice_tile -> bit-identical results.
// TAIL (2, only with EXACT): the last row block holds r = n - 8 (NI - 1) <= TAIL live rows. A DMMA spends the
// same 16 cycles on it as on a full block (2 NI of the tile's NI (NI + 1) MMAs for one or two rows), so those rows are
// taken off the MMA path: row rho of the block is a per-lane DFMA chain over the lane's own k-row (the L^-1 entry is a
// 4-address broadcast LDS from the same fragment block), summed over the four k-row lanes at the end of the tile.
// 2.75 cycles per (row, k-block) instead of 16 per k-block; the last k-block (columns >= 8 NI - 4 >= n) is not
// needed at all. Rows r..TAIL-1 are zero in the fragment block and add nothing.
// MROW (only with EXACT, TAIL == 0, 3..7 live rows in the last row block): the posterior mean rides on the MMA path.
// The kernel has written alpha into row 7 of the last row block of the STAGED copy of L^-1 (a spare row), so
// C[7][col] of that block is sum_j alpha_j k*_j(col): the lanes with lane >> 2 == 7 end the tile holding the means of the
// tile's eight candidates (returned in mu_part / mu_hi for columns 2 (lane & 3) and + 1) and the per-look-up DFMA and the
// alpha loads disappear. Those lanes keep the spare row out of the sum of squares.
template <int NI, bool EXACT, int TAIL, bool MROW = false>
__device__ __forceinline__ void lattice_tile8(const double* __restrict__ wsm, const uint4* __restrict__ xyq,
                                              const double2* __restrict__ alq, const double* __restrict__ tab,
                                              unsigned wpk, unsigned cpk, int lane, int nI, double& ss0, double& ss1,
                                              double& mu_part, double& tail_ss, double* mu_hi = nullptr) {
  static_assert(TAIL == 0 || EXACT, "tail rows only for full-size trials");
  static_assert(!MROW || (EXACT && TAIL == 0), "mean row only for full-size trials without tail rows");
  constexpr int ND = TAIL ? NI - 1 : NI;                   // row blocks on the MMA path
  constexpr int NJ = TAIL ? 2 * NI - 1 : 2 * NI;           // k-blocks that can hold a non-zero column
  double acc[ND > 0 ? ND : 1][2];
  double tl[TAIL > 0 ? TAIL : 1];
#pragma unroll
  for (int I = 0; I < ND; ++I) { acc[I][0] = 0.0; acc[I][1] = 0.0; }
#pragma unroll
  for (int r = 0; r < TAIL; ++r) tl[r] = 0.0;
  auto look_up = [&](unsigned pk) { return __ldg(tab + __dp4a(__vabsdiffu4(cpk, pk), wpk, 0u)); };
  uint4 pq = xyq[0];                                       // observations of k-blocks 0..3 (this lane's k-row)
  double kv0 = look_up(pq.x), kv1 = look_up(pq.y);         // software pipeline: table values two k-blocks ahead
  double2 a2 = make_double2(0.0, 0.0);
  mu_part = 0.0;
  const double* wtail = wsm + (size_t)(NI - 1) * NI * 32 + (lane & 3);
#pragma unroll
  for (int J = 0; J < NJ; ++J) {
    if (!EXACT && J >= 2 * nI) break;                      // uniform
    const double kv = kv0;
    kv0 = kv1;
    if (J + 2 < NJ && (EXACT || J + 2 < 2 * nI)) {
      if ((J + 2) % 4 == 0) pq = xyq[(J + 2) / 4];
      kv1 = look_up((J + 2) % 4 == 0 ? pq.x : (J + 2) % 4 == 1 ? pq.y : (J + 2) % 4 == 2 ? pq.z : pq.w);
    }
    if (!MROW) {
      if (J % 2 == 0) a2 = alq[J / 2];
      // padded observations (j >= n) meet zero columns of L^-1 and alpha = 0: any finite value is harmless
      mu_part = fma(kv, (J % 2 == 0) ? a2.x : a2.y, mu_part);
    }
#pragma unroll
    for (int I = J / 2; I < ND; ++I) {
      if (EXACT || I < nI) dmma_m8n8k4(acc[I][0], acc[I][1], wsm[(size_t)(I * (I + 1) + J) * 32 + lane], kv);
    }
#pragma unroll
    for (int r = 0; r < TAIL; ++r) tl[r] = fma(wtail[J * 32 + 4 * r], kv, tl[r]);
  }
  if (MROW) {
    mu_part = acc[ND - 1][0];
    *mu_hi = acc[ND - 1][1];
    if ((lane >> 2) == 7) { acc[ND - 1][0] = 0.0; acc[ND - 1][1] = 0.0; }
  }
  ss0 = 0.0;
  ss1 = 0.0;
#pragma unroll
  for (int I = 0; I < ND; ++I) {
    ss0 = fma(acc[I][0], acc[I][0], ss0);
    ss1 = fma(acc[I][1], acc[I][1], ss1);
  }
  tail_ss = 0.0;
#pragma unroll
  for (int r = 0; r < TAIL; ++r) {                         // v of row 8 (NI - 1) + r for candidate lane >> 2
    double v = tl[r];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    tail_ss = fma(v, v, tail_ss);
  }
}

// per-k-row strides of the alpha / coordinate arrays in 16-byte units, made ODD so that the four k-rows of a warp's
// LDS.128 fall into distinct bank groups (a 128-byte stride - NI = 8 - would be a 4-way conflict)
template <int NI> __host__ __device__ constexpr int lattice8_al_stride() { return NI | 1; }                  // double2 units (>= NJ / 2)
template <int NI> __host__ __device__ constexpr int lattice8_xy_stride() { return ((2 * NI + 3) / 4) | 1; }  // uint4 units (>= NJ / 4)

template <int NI>
__host__ __device__ constexpr size_t lattice8_fixed_smem() {
  return (size_t)NI * (NI + 1) * 32 * sizeof(double) + 4 * (size_t)(lattice8_al_stride<NI>() + lattice8_xy_stride<NI>()) * 16;
}

// One super-tile = four column tiles = 32 candidates, for one tile variant (chosen once per CTA: the variant depends only
// on the trial's n). Returns this lane's candidate offset `off` within the super-tile with its sum of squares and mean.
//
// Reduction of the four tiles' partial sums. NI <= KB8_RS_MAX_NI: a reduce-scatter — the shuffles share the LSU data
// pipe with the fragment loads and the table gathers, and that pipe is 85-89 % busy at n <= 32 (ncu). A butterfly per
// tile moves 44 words per tile through it; here every exchange halves the number of values a lane keeps (8 column
// sums over the 8 row lanes: xor 4, 8, 16; 4 mean sums over the 4 k-row lanes: xor 1, 2), 22-24 words per FOUR tiles,
// and each lane ends with exactly one candidate. The pairing order equals the butterfly's (same partner at every
// level, fp add is commutative), so the sums are bit-identical to it. Measured: +4..13 % at n <= 24, nothing at
// n = 32..48, -1 % from n = 64 on (the six values kept across tiles cost registers) — larger NI keep the per-tile butterfly.
template <int NI, bool EXACT, int TAIL, bool MROW>
__device__ __forceinline__ void lattice8_supertile(const double* __restrict__ wsm, const uint4* __restrict__ xyq,
                                                   const double2* __restrict__ alq, const double* __restrict__ tab,
                                                   unsigned wpk, int Gx, int Gy, int first, int lane, int nI,
                                                   double& my_ss, double& my_mu, int& off) {
  constexpr bool RS = NI <= KB8_RS_MAX_NI;
  static_assert(!(RS && MROW), "the mean row is handed over by the butterfly epilogue only");
  const int j0 = lane & 1, j1 = (lane >> 1) & 1, b0 = (lane >> 2) & 1, b1 = (lane >> 3) & 1, b2 = lane >> 4;
  // lattice coordinates of this lane's candidate in tile 0 (first = base + lane / 4); later tiles step x by 8 (Gx >= 8).
  // Candidates past G - 1 stay inside the table (their results are discarded by the caller).
  int gyi = first / Gx, gxi = first - gyi * Gx;
  gyi = min(gyi, Gy - 1);
  double w_even = 0.0, m_even = 0.0, t_even = 0.0, u0 = 0.0, mu0 = 0.0, tt0 = 0.0;
  double ssF = 0.0, muF = 0.0, ttF = 0.0;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const unsigned cpk = (unsigned)gxi | ((unsigned)gyi << 8);
    double ss0, ss1, mp, tq, mh = 0.0;
    lattice_tile8<NI, EXACT, TAIL, MROW>(wsm, xyq, alq, tab, wpk, cpk, lane, nI, ss0, ss1, mp, tq, &mh);
    gxi += 8;
    if (gxi >= Gx) { gxi -= Gx; gyi = min(gyi + 1, Gy - 1); }
    if constexpr (RS) {
      // level 1 (row lanes xor 4): keep column 2 jrow + b0 of this tile
      const double w = (b0 ? ss1 : ss0) + __shfl_xor_sync(0xffffffffu, b0 ? ss0 : ss1, 4);
      if (!(q & 1)) {                                     // uniform
        w_even = w; m_even = mp; t_even = tq;
      } else {
        // level 2 (xor 8): keep tile 2 (q / 2) + b1; mean level 1 (k-row lanes xor 1): keep tile 2 (q / 2) + j0
        const double u = (b1 ? w : w_even) + __shfl_xor_sync(0xffffffffu, b1 ? w_even : w, 8);
        const double mm = (j0 ? mp : m_even) + __shfl_xor_sync(0xffffffffu, j0 ? m_even : mp, 1);
        const double tt = j0 ? tq : t_even;
        if (q == 1) {
          u0 = u; mu0 = mm; tt0 = tt;
        } else {
          // level 3 (xor 16): keep tile pair b2; mean level 2 (xor 2): keep tile pair j1
          ssF = (b2 ? u : u0) + __shfl_xor_sync(0xffffffffu, b2 ? u0 : u, 16);
          muF = (j1 ? mm : mu0) + __shfl_xor_sync(0xffffffffu, j1 ? mu0 : mm, 2);
          ttF = j1 ? tt : tt0;
        }
      }
    } else {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {                  // rows live in lane>>2
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, o);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, o);
      }
      const int cc = lane & 7;                            // hand column cc of this tile to lane 8q + cc
      double tm;
      if constexpr (MROW) {                               // the means sit complete on lanes 28..31
        const double m0 = __shfl_sync(0xffffffffu, mp, 28 + (cc >> 1));
        const double m1 = __shfl_sync(0xffffffffu, mh, 28 + (cc >> 1));
        tm = (cc & 1) ? m1 : m0;
      } else {
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) mp += __shfl_xor_sync(0xffffffffu, mp, o);   // k-rows live in lane&3
        tm = __shfl_sync(0xffffffffu, mp, cc * 4);
      }
      const double t0 = __shfl_sync(0xffffffffu, ss0, cc >> 1);
      const double t1 = __shfl_sync(0xffffffffu, ss1, cc >> 1);
      const double tt = TAIL ? __shfl_sync(0xffffffffu, tq, cc * 4) : 0.0;
      if ((lane >> 3) == q) {
        ssF = ((cc & 1) ? t1 : t0) + tt;
        muF = tm;
      }
    }
  }
  if constexpr (RS) {
    // this lane holds the column sum of candidate off = 8 (2 b2 + b1) + 4 j1 + 2 j0 + b0 of the super-tile; its mean (and
    // tail sum) sit on the lane with column bits (j1, j0, b0) as row lane and tile bits (b2, b1) as k-row lane
    const int src = 4 * (4 * j1 + 2 * j0 + b0) + 2 * b2 + b1;
    my_mu = __shfl_sync(0xffffffffu, muF, src);
    my_ss = ssF + (TAIL ? __shfl_sync(0xffffffffu, ttF, src) : 0.0);
    off = 8 * (2 * b2 + b1) + 4 * j1 + 2 * j0 + b0;
  } else {                                                // butterfly: lane l holds candidate l
    my_mu = muF;
    my_ss = ssF;
    off = lane;
  }
}

template <int NI, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS == 128 ? (NI <= 5 ? KB_MINB_SMALL : (NI <= 8 ? KB8_MINB_MID : 4)) : (THREADS == 256 ? 2 : 1)))
grid_acq_lattice8_kernel(const GridAcqParams p) {
  constexpr int NWARPS = THREADS / 32;
  constexpr int ALS = 2 * lattice8_al_stride<NI>();   // alpha row stride in doubles
  constexpr int XYS = 4 * lattice8_xy_stride<NI>();   // coordinate row stride in 32-bit words
  extern __shared__ __align__(16) double sm[];
  const int t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - t * p.chunks;
  const int n = p.n[t];
  if (n <= 0) return;                     // finished / inactive trial (uniform per CTA)
  if (n > p.n_cap || n > 8 * NI) {        // n_hint too small: report, do not touch memory
    if (threadIdx.x == 0) {
      p.part_val[(size_t)t * p.chunks + chunk] = -INFINITY;
      p.part_idx[(size_t)t * p.chunks + chunk] = INT_MAX;
      if (p.status) atomicOr(&p.status[t], B200BO_ST_ALL_MASKED);
    }
    return;
  }
  const int nI = (n + 7) >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = nI * (nI + 1);
  const int col0 = chunk * p.cols_per_cta;
  const int col_end = min(p.G, col0 + p.cols_per_cta);
  double* wsm = sm;                                                    // NI (NI + 1) * 32
  double* alT = sm + (size_t)NI * (NI + 1) * 32;                       // [4][ALS]: alpha of observation 4 J + k-row
  unsigned* xyT = reinterpret_cast<unsigned*>(alT + 4 * ALS);          // [4][XYS]: x | y << 8 of the same observation
  unsigned* obsbits = xyT + 4 * XYS;                                   // bit (g - col0): candidate g is an observed point
  __shared__ double red_val[NWARPS];
  __shared__ int red_idx[NWARPS];
  __shared__ int red_flag;
  __shared__ __align__(8) unsigned long long stage_bar;
  {
    if (threadIdx.x == 0) mbar_init(&stage_bar, 1);
    for (int w = threadIdx.x; w < (p.cols_per_cta >> 5); w += THREADS) obsbits[w] = 0u;
    __syncthreads();
    // L^-1 fragments: one contiguous, 256-byte granular range -> ONE bulk copy through the TMA unit
    if (threadIdx.x == 0) tma_bulk_g2s(wsm, p.wfrag + (size_t)t * p.wstride, (unsigned)nblk * 256u, &stage_bar);
    for (int j = threadIdx.x; j < 4 * XYS; j += THREADS) {
      const bool v = j < n;
      const double x = v ? p.X[((size_t)t * p.n_max + j) * 2] : p.x0;
      const double y = v ? p.X[((size_t)t * p.n_max + j) * 2 + 1] : p.y0;
      const double fx = rint((x - p.x0) / p.h), fy = rint((y - p.y0) / p.h);
      const bool on = (__dadd_rn(p.x0, __dmul_rn(p.h, fx)) == x) && (__dadd_rn(p.y0, __dmul_rn(p.h, fy)) == y) && fx >= 0.0 && fy >= 0.0 &&
                      fx < (double)p.Gx && fy < (double)p.Gy;
      const int xi = on ? (int)fx : 0, yi = on ? (int)fy : 0;
      xyT[(j & 3) * XYS + (j >> 2)] = (unsigned)xi | ((unsigned)yi << 8);
      if (j < 4 * ALS) alT[(j & 3) * ALS + (j >> 2)] = v ? p.alpha[(size_t)t * p.n_max + j] : 0.0;
      if (v && !on && p.status) atomicOr(&p.status[t], B200BO_ST_OFF_LATTICE);
      if (v && on && p.mask_observed) {
        const int gflat = yi * p.Gx + xi;
        if (gflat >= col0 && gflat < col_end) atomicOr(&obsbits[(gflat - col0) >> 5], 1u << ((gflat - col0) & 31));
      }
    }
    if (threadIdx.x == 0) red_flag = 0;
  }
  const double* tab = p.ktab + (size_t)(p.ktab_id ? p.ktab_id[t] : 0) * p.Gx * p.Gy;
  // keep base + slot offset as ONE 64-bit register pair: the look-up address is then a single IMAD.WIDE (left to
  // itself the compiler re-adds the uniform slot offset per look-up: IADD3 + IMAD.X + LEA + LEA.HI.X)
  asm volatile("" : "+l"(tab));
  const double cst = p.theta[(size_t)t * 4 + 1];
  const double sc = p.theta[(size_t)t * 4 + 2];
  const double best_f = p.best_f[t];
  __syncthreads();
  mbar_wait(&stage_bar, 0);               // the staged fragments have landed

  const int jrow = lane & 3, bcol = lane >> 2;
  const uint4* xyq = reinterpret_cast<const uint4*>(xyT + jrow * XYS);
  const double2* alq = reinterpret_cast<const double2*>(alT + jrow * ALS);
  const unsigned wpk = 1u | ((unsigned)p.Gx << 8);
  double best_v = -INFINITY;
  int best_i = INT_MAX;
  int flags = 0;
  const bool prune = p.acq_id == B200BO_ACQ_EI && p.acq_out == nullptr && p.prune;
  double wbest = 0.0, lthr = INFINITY;
  // per-warp stack of candidates that survived the pruning test (64 entries: at most 31 left + 32 pushed)
  double* q_mu = reinterpret_cast<double*>(obsbits + (p.cols_per_cta >> 5)) + warp * 128;
  double* q_var = q_mu + 64;
  int* q_g = reinterpret_cast<int*>(reinterpret_cast<double*>(obsbits + (p.cols_per_cta >> 5)) + NWARPS * 128) + warp * 64;
  int qn = 0;
  // one or two live rows in the last row block leave the MMA path (lattice_tile8, TAIL = 2): measured -5..-7 % per
  // launch at n = 8k + 1, 8k + 2. Four DFMA rows (n = 8k + 3, 8k + 4) were measured too and LOSE 2..20 %: not built.
  static_assert(KB8_TAIL == 0 || KB8_TAIL == 1, "KB8_TAIL is a switch");
  const int tail_rows = (KB8_TAIL && nI == NI && n - 8 * (NI - 1) <= 2) ? 2 : 0;
  // 3..7 live rows in the last row block: row 7 is spare and carries alpha (lattice_tile8, MROW); butterfly epilogue only
  const bool mrow = KB8_MROW && NI > KB8_RS_MAX_NI && nI == NI && !tail_rows && n - 8 * (NI - 1) <= 7;
  if (mrow) {                             // uniform
    for (int j = threadIdx.x; j < 8 * NI; j += THREADS)
      wsm[(size_t)((NI - 1) * NI + (j >> 2)) * 32 + 28 + (j & 3)] = alT[(j & 3) * ALS + (j >> 2)];
    __syncthreads();
  }
  const int variant = tail_rows == 2 ? 0 : (mrow ? 1 : (nI == NI ? 2 : 3));     // tile variant of this trial

  for (int base = col0 + warp * 32; base < col_end; base += NWARPS * 32) {
    double my_ss, my_mu;
    int off;
    switch (variant) {                                    // uniform
      case 0: lattice8_supertile<NI, true, 2, false>(wsm, xyq, alq, tab, wpk, p.Gx, p.Gy, base + bcol, lane, nI, my_ss, my_mu, off); break;
      case 1: lattice8_supertile<NI, true, 0, (NI > KB8_RS_MAX_NI)>(wsm, xyq, alq, tab, wpk, p.Gx, p.Gy, base + bcol, lane, nI, my_ss, my_mu, off); break;
      case 2: lattice8_supertile<NI, true, 0, false>(wsm, xyq, alq, tab, wpk, p.Gx, p.Gy, base + bcol, lane, nI, my_ss, my_mu, off); break;
      default: lattice8_supertile<NI, false, 0, false>(wsm, xyq, alq, tab, wpk, p.Gx, p.Gy, base + bcol, lane, nI, my_ss, my_mu, off); break;
    }
    const int g = base + off;
    const bool my_masked = (obsbits[(base - col0) >> 5] >> off) & 1u;
    const double mu = cst + my_mu;
    const double var = sc - my_ss;
    const bool live = g < col_end;
    const bool m = my_masked || (p.extra_mask && live && p.extra_mask[g]);
    if (live) {
      if (p.mu_out) p.mu_out[(size_t)t * p.G + g] = mu;
      if (p.var_out) p.var_out[(size_t)t * p.G + g] = var;
    }
    if (prune) {
      // EI with exact pruning and warp-level compaction. For mu < best_f, EI = sigma phi(u) + (mu - best_f) Phi(u)
      // <= 0.39894 sqrt(s) exp(-u^2 / 2) because the posterior sigma never exceeds the prior sqrt(s): a candidate whose
      // bound is below the warp's running maximum can neither win nor tie, so its sqrt / exp / erfc (~160 FP64-pipe
      // slots) are skipped; the test is two multiplies. Survivors are not evaluated in place (one surviving lane would
      // make the whole warp walk the EI code) but pushed on a per-warp stack in shared memory and evaluated 32 at a
      // time with full lanes; the threshold is refreshed after each batch (a stale threshold only prunes less).
      const double d = mu - best_f;
      const bool skip = wbest > 0.0 && d < 0.0 && 0.5 * d * d > fmax(var, p.var_floor) * lthr;
      const bool need = live && !m && !skip;
      const unsigned bal = __ballot_sync(0xffffffffu, need);
      if (bal) {                                          // uniform
        if (need) {
          const int pos = qn + __popc(bal & ((1u << lane) - 1u));
          q_mu[pos] = mu; q_var[pos] = var; q_g[pos] = g;
        }
        qn += __popc(bal);
        if (qn >= 32) {
          __syncwarp();
          qn -= 32;
          const double a = acquisition<B200BO_ACQ_EI>(q_mu[qn + lane], q_var[qn + lane], best_f, p.acq_param, p.var_floor);
          const int ga = q_g[qn + lane];
          if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
          if (!isnan(a) && better(a, ga, best_v, best_i)) { best_v = a; best_i = ga; }
          __syncwarp();
          const double nb = warp_max(best_v);
          if (nb > wbest) {                               // uniform
            wbest = nb;
            lthr = log(kInvSqrt2Pi * sqrt(sc) / nb) * (1.0 + 1e-12) + 1e-12;
          }
        }
      }
    } else if (live) {
      double a;
      if (p.acq_id == B200BO_ACQ_UCB) a = acquisition<B200BO_ACQ_UCB>(mu, var, best_f, p.acq_param, p.var_floor);
      else if (p.acq_id == B200BO_ACQ_PI) a = acquisition<B200BO_ACQ_PI>(mu, var, best_f, p.acq_param, p.var_floor);
      else a = acquisition<B200BO_ACQ_EI>(mu, var, best_f, p.acq_param, p.var_floor);
      if (p.acq_out) p.acq_out[(size_t)t * p.G + g] = a;
      if (!m) {
        if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
        if (!isnan(a) && better(a, g, best_v, best_i)) { best_v = a; best_i = g; }
      }
    }
  }
  if (prune && qn > 0) {                                  // the rest of the stack (qn < 32)
    __syncwarp();
    if (lane < qn) {
      const double a = acquisition<B200BO_ACQ_EI>(q_mu[lane], q_var[lane], best_f, p.acq_param, p.var_floor);
      const int ga = q_g[lane];
      if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
      if (!isnan(a) && better(a, ga, best_v, best_i)) { best_v = a; best_i = ga; }
    }
  }
  warp_argmax(best_v, best_i);
  flags = __reduce_or_sync(0xffffffffu, flags);
  if (lane == 0) {
    red_val[warp] = best_v;
    red_idx[warp] = best_i;
    if (flags) atomicOr(&red_flag, flags);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bv = red_val[0];
    int bi = red_idx[0];
    for (int w = 1; w < NWARPS; ++w)
      if (better(red_val[w], red_idx[w], bv, bi)) { bv = red_val[w]; bi = red_idx[w]; }
    p.part_val[(size_t)t * p.chunks + chunk] = bv;
    p.part_idx[(size_t)t * p.chunks + chunk] = bi;
    if (red_flag && p.status) atomicOr(&p.status[t], red_flag);
  }
}

template <int NI, int THREADS>
static int launch_lattice8_t(const GridAcqParams& p, int T, cudaStream_t st) {
  constexpr size_t stack_bytes = (size_t)(THREADS / 32) * 64 * (2 * sizeof(double) + sizeof(int));   // survivors' stacks
  const size_t smem = lattice8_fixed_smem<NI>() + (size_t)(p.cols_per_cta >> 5) * sizeof(unsigned) + stack_bytes;
  // the mask grows with the column chunk (5 KB for a whole 200x200 lattice): set the limit once, to the largest case
  cudaError_t e = cudaFuncSetAttribute(grid_acq_lattice8_kernel<NI, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(lattice8_fixed_smem<NI>() + 8320 + stack_bytes));
  if (e != cudaSuccess) return cuda_status(e, "grid_posterior_acq_argmax_lattice: smem attribute");
  grid_acq_lattice8_kernel<NI, THREADS><<<(unsigned)((size_t)T * p.chunks), THREADS, smem, st>>>(p);
  return 0;
}

template <int NI, int THREADS>
static int launch_lattice_t(const GridAcqParams& p, int T, cudaStream_t st) {
  size_t smem = ((size_t)NI * (NI + 1) * 32 + 8 * (size_t)NI) * sizeof(double) + 8 * (size_t)NI * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(grid_acq_lattice_kernel<NI, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "grid_posterior_acq_argmax_lattice: smem attribute");
  grid_acq_lattice_kernel<NI, THREADS><<<(unsigned)((size_t)T * p.chunks), THREADS, smem, st>>>(p);
  return 0;
}

template <int NI>
static int launch_lattice(const GridAcqParams& p, int T, cudaStream_t st) {
  // CTA size: with >= 9 row blocks the L^-1 fragments of four 128-thread CTAs leave the SM's L1 too small for the
  // kernel table (look-ups then pay L2 latency); two 256-thread CTAs keep the same 16 warps/SM and the table in L1.
  // Measured (tools/perf_probe.py): n = 110: 0.80 -> 0.85 of FP64 peak; n <= 64: 128 threads are better.
  static const int thr = env_int("B200BO_KB_THREADS", 0);
  const int use = thr ? thr : (NI >= 9 ? 256 : 128);
  // byte-coordinate tile (VABSDIFF4 + IDP.4A look-ups, observed-point bit mask) whenever the lattice fits one byte per
  // axis and the mask fits the reserved shared memory; B200BO_KB_BYTE=0 forces the general tile (A/B runs)
  static const int byte_ok = env_int("B200BO_KB_BYTE", 1);
  if (byte_ok && p.Gx >= 8 && p.Gx <= 255 && p.Gy <= 256 && (p.cols_per_cta >> 5) * sizeof(unsigned) <= 8320) {
    if (use == 256) return launch_lattice8_t<NI, 256>(p, T, st);
    return launch_lattice8_t<NI, 128>(p, T, st);
  }
  if (use == 256) return launch_lattice_t<NI, 256>(p, T, st);
  return launch_lattice_t<NI, 128>(p, T, st);
}

static int dispatch_lattice(const GridAcqParams& p, int T, cudaStream_t st) {
  switch ((p.n_cap + 7) / 8) {
    case 1: return launch_lattice<1>(p, T, st);
    case 2: return launch_lattice<2>(p, T, st);
    case 3: return launch_lattice<3>(p, T, st);
    case 4: return launch_lattice<4>(p, T, st);
    case 5: return launch_lattice<5>(p, T, st);
    case 6: return launch_lattice<6>(p, T, st);
    case 7: return launch_lattice<7>(p, T, st);
    case 8: return launch_lattice<8>(p, T, st);
    case 9: return launch_lattice<9>(p, T, st);
    case 10: return launch_lattice<10>(p, T, st);
    case 11: return launch_lattice<11>(p, T, st);
    case 12: return launch_lattice<12>(p, T, st);
    case 13: return launch_lattice<13>(p, T, st);
    case 14: return launch_lattice<14>(p, T, st);
    case 15: return launch_lattice<15>(p, T, st);
    default: return launch_lattice<16>(p, T, st);
  }
}

__global__ void __launch_bounds__(128)
argmax_finish_kernel(const double* __restrict__ part_val, const int32_t* __restrict__ part_idx, int chunks, int T,
                     const int32_t* __restrict__ n, int32_t* __restrict__ arg_idx, double* __restrict__ arg_val,
                     int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  if (n[t] <= 0) {                         // inactive trial: nothing was evaluated
    arg_idx[t] = -1;
    if (arg_val) arg_val[t] = -INFINITY;
    return;
  }
  double bv = -INFINITY;
  int bi = INT_MAX;
  for (int c = 0; c < chunks; ++c) {
    double v = part_val[(size_t)t * chunks + c];
    int i = part_idx[(size_t)t * chunks + c];
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
  if (bi == INT_MAX) {
    bi = -1;
    if (status) atomicOr(&status[t], B200BO_ST_ALL_MASKED);
  }
  arg_idx[t] = bi;
  if (arg_val) arg_val[t] = bv;
}

void launch_argmax_finish(const double* part_val, const int32_t* part_idx, int chunks, int T, const int32_t* n,
                          int32_t* arg_idx, double* arg_val, int32_t* status, cudaStream_t st) {
  argmax_finish_kernel<<<(T + 127) / 128, 128, 0, st>>>(part_val, part_idx, chunks, T, n, arg_idx, arg_val, status);
}

// Column chunks per trial. A launch has T x chunks CTAs for `slots` resident CTAs (148 SMs x CTAs per SM of the
// instantiation): first enough chunks to fill the GPU once, then up to 15 more when that shortens the last, partially
// filled wave — e.g. 1 000 trials on 740 slots: 1 chunk = 2 waves for 1.35 waves of work, 5 chunks = 7 waves for 6.76.
// Every extra chunk re-stages L^-1 and restarts the pruning threshold (~1 % each, measured with B200BO_KB_WAVES); 2 % per
// chunk is charged in the cost, so large batches (T >> slots) keep one CTA per trial.
inline int max_chunks_of(int G) { return (G + GA_THREADS - 1) / GA_THREADS; }

inline int pick_chunks(int T, int G, int slots) {
  static const int waves = env_int("B200BO_KB_WAVES", 1);   // experiments: force at least this many waves
  const int max_chunks = max_chunks_of(G);
  int lo = (int)(((long long)waves * slots + T - 1) / T);
  if (lo < 1) lo = 1;
  if (lo > max_chunks) lo = max_chunks;
  int best = lo;
  double best_cost = 1e300;
  for (int c = lo; c <= max_chunks && c < lo + 16; ++c) {
    const long long ctas = (long long)T * c;
    const double cost = (double)((ctas + slots - 1) / slots) / c * (1.0 + 0.02 * (c - lo));
    if (cost < best_cost - 1e-12) { best_cost = cost; best = c; }
  }
  return best;
}

// upper bound of pick_chunks over every instantiation (workspace sizing; at most 6 CTAs per SM)
inline int max_pick_chunks(int T, int G) {
  static const int waves = env_int("B200BO_KB_WAVES", 1);
  long long lo = ((long long)waves * 148 * 6 + T - 1) / T;
  long long hi = lo + 16;
  const int max_chunks = max_chunks_of(G);
  return (int)(hi > max_chunks ? max_chunks : hi);
}

// resident CTAs per SM of the kernel a launch with n_cap observations will take (launch bounds / shared memory)
inline int kb_ctas_per_sm(bool lattice, int n_cap) {
  const int nI = (n_cap + 7) / 8;
  if (!lattice) return nI <= 8 ? 4 : 3;
  if (nI >= 9) return 2;                 // 256-thread CTAs
  return nI <= 5 ? KB_MINB_SMALL : KB8_MINB_MID;
}

inline size_t grid_acq_smem(int nI) {
  size_t npad = ((size_t)8 * nI + 31) & ~(size_t)31;
  return ((size_t)nI * (nI + 1) * 32 + 3 * npad) * sizeof(double);
}

template <int KID, int NI_MAX, int CH, int MINB, int NT>
static int launch_grid_acq_v(const GridAcqParams& p, int T, cudaStream_t st) {
  size_t smem = grid_acq_smem(NI_MAX);
  cudaError_t e = cudaFuncSetAttribute(grid_acq_kernel<KID, NI_MAX, CH, MINB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "grid_posterior_acq_argmax: smem attribute");
  // dynamic smem sized by the largest n actually possible in this instantiation would waste
  // occupancy for small n; size it by n_max rounded to the tile instead.
  int nIa = (p.n_cap + 7) / 8;
  if (nIa > NI_MAX) nIa = NI_MAX;
  smem = grid_acq_smem(nIa);
  grid_acq_kernel<KID, NI_MAX, CH, MINB, NT><<<(unsigned)((size_t)T * p.chunks), GA_THREADS, smem, st>>>(p);
  return 0;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// generic kernel: two accumulator chains per row-block group, one column tile per MMA group (measured best)
template <int KID, int NI_MAX>
static int launch_grid_acq(const GridAcqParams& p, int T, cudaStream_t st) {
  return launch_grid_acq_v<KID, NI_MAX, 2, (NI_MAX <= 8 ? 4 : 3), 1>(p, T, st);
}

template <int KID>
static int dispatch_ni(const GridAcqParams& p, int T, cudaStream_t st) {
  if (p.n_cap <= 32) return launch_grid_acq<KID, 4>(p, T, st);
  if (p.n_cap <= 64) return launch_grid_acq<KID, 8>(p, T, st);
  return launch_grid_acq<KID, 16>(p, T, st);
}

// ktab[slot][dy * TW + dx] = s k(h sqrt(dx^2 + dy^2) / l): the same device function and the same
// operand values the generic path sees for lattice points, so both variants agree bit for bit.
__global__ void __launch_bounds__(256)
kernel_table_kernel(const double* __restrict__ theta, int H, int kernel_id, double h, int TW, int TH,
                    double* __restrict__ ktab) {
  for (int slot = blockIdx.y; slot < H; slot += gridDim.y) {         // gridDim.y <= 65535: slots beyond that wrap
    const double sc = theta[(size_t)slot * 4 + 2], ell = theta[(size_t)slot * 4 + 3];
    const KernParams kp = make_kern_params(sc, ell, kernel_id);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < TW * TH; e += gridDim.x * blockDim.x) {
      const int dyi = e / TW, dxi = e - dyi * TW;
      const double dx = h * (double)dxi, dy = h * (double)dyi;
      const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      ktab[(size_t)slot * TW * TH + e] = kernel_from_r2_dyn(kernel_id, r2, sc, ell, kp.a_scale);
    }
  }
}

}  // namespace b200bo

extern "C" {

size_t b200bo_grid_acq_workspace_bytes(int T, int G) {
  if (T <= 0 || G <= 0) return 0;
  int chunks = b200bo::max_pick_chunks(T, G);
  return (size_t)T * chunks * (sizeof(double) + sizeof(int32_t)) + 16;
}

int b200bo_grid_acq_chunks(int T, int G, int n_cap, int lattice) {
  if (T <= 0 || G <= 0 || n_cap < 1) return 0;
  return b200bo::pick_chunks(T, G, 148 * b200bo::kb_ctas_per_sm(lattice != 0, n_cap));
}

static int grid_acq_common(bool lattice, const double* grid, int G, int Gx, int Gy, double x0, double y0, double h,
                           const double* ktab, const int32_t* ktab_id, const double* X, const double* wfrag,
                           const double* alpha, const int32_t* n, const double* theta, const double* best_f,
                           const uint8_t* extra_mask, int mask_observed, int acq_id, double acq_param,
                           double var_floor, int32_t* arg_idx, double* arg_val, double* mu, double* var, double* acq,
                           int32_t* status, int T, int n_max, int n_hint, int kernel_id, void* ws, size_t ws_bytes,
                           void* stream) {
  using namespace b200bo;
  const char* who = lattice ? "grid_posterior_acq_argmax_lattice" : "grid_posterior_acq_argmax";
  if (T < 0 || G < 1 || n_max < 1 || n_max > B200BO_NMAX)
    return arg_error("%s: T=%d G=%d n_max=%d (max %d)", who, T, G, n_max, B200BO_NMAX);
  if (T == 0) return 0;
  if ((!lattice && !grid) || !X || !wfrag || !alpha || !n || !theta || !best_f || !arg_idx || !ws)
    return arg_error("%s: null pointer", who);
  if (lattice && (!ktab || Gx < 1 || Gy < 1 || (long long)Gx * Gy != G || !(h > 0.0)))
    return arg_error("%s: bad lattice Gx=%d Gy=%d h=%g", who, Gx, Gy, h);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("%s: kernel_id=%d", who, kernel_id);
  if (acq_id < 0 || acq_id > 2) return arg_error("%s: acq_id=%d", who, acq_id);
  if (acq_id == B200BO_ACQ_UCB && !(acq_param >= 0.0)) return arg_error("%s: UCB beta < 0", who);
  if (ws_bytes < b200bo_grid_acq_workspace_bytes(T, G)) return arg_error("%s: workspace too small", who);
  cudaStream_t st = as_stream(stream);
  GridAcqParams p;
  p.grid = reinterpret_cast<const double2*>(grid);
  p.G = G;
  p.X = X; p.wfrag = wfrag; p.alpha = alpha; p.n = n; p.theta = theta; p.best_f = best_f;
  p.extra_mask = extra_mask;
  p.mask_observed = mask_observed;
  p.acq_id = acq_id;
  p.acq_param = acq_id == B200BO_ACQ_UCB ? sqrt(acq_param) : acq_param;
  p.var_floor = var_floor;
  p.n_cap = (n_hint > 0 && n_hint < n_max) ? n_hint : n_max;
  p.chunks = pick_chunks(T, G, 148 * kb_ctas_per_sm(lattice, p.n_cap));
  static const int prune_env = env_int("B200BO_KB_PRUNE", 1);   // read once per process
  p.prune = prune_env;
  p.cols_per_cta = (((G + p.chunks - 1) / p.chunks + GA_THREADS - 1) / GA_THREADS) * GA_THREADS;
  p.part_val = reinterpret_cast<double*>(ws);
  p.part_idx = reinterpret_cast<int32_t*>(p.part_val + (size_t)T * p.chunks);
  p.mu_out = mu; p.var_out = var; p.acq_out = acq;
  p.status = status;
  p.n_max = n_max;
  p.wstride = b200bo_wfrag_doubles(n_max);
  p.Gx = Gx; p.Gy = Gy; p.x0 = x0; p.y0 = y0; p.h = h; p.ktab = ktab; p.ktab_id = ktab_id;
  int rc;
  if (lattice) {
    // table look-ups do not depend on the kernel family: one set of instantiations serves all four
    rc = dispatch_lattice(p, T, st);
  } else {
    switch (kernel_id) {
      case B200BO_KERNEL_RBF: rc = dispatch_ni<B200BO_KERNEL_RBF>(p, T, st); break;
      case B200BO_KERNEL_MATERN12: rc = dispatch_ni<B200BO_KERNEL_MATERN12>(p, T, st); break;
      case B200BO_KERNEL_MATERN32: rc = dispatch_ni<B200BO_KERNEL_MATERN32>(p, T, st); break;
      default: rc = dispatch_ni<B200BO_KERNEL_MATERN52>(p, T, st); break;
    }
  }
  if (rc) return rc;
  B200BO_CHECK_LAUNCH(who);
  argmax_finish_kernel<<<(T + 127) / 128, 128, 0, st>>>(p.part_val, p.part_idx, p.chunks, T, n, arg_idx, arg_val, status);
  B200BO_CHECK_LAUNCH(who);
  return 0;
}

int b200bo_grid_posterior_acq_argmax(const double* grid, int G, const double* X, const double* wfrag,
                                     const double* alpha, const int32_t* n, const double* theta,
                                     const double* best_f, const uint8_t* extra_mask, int mask_observed, int acq_id,
                                     double acq_param, double var_floor, int32_t* arg_idx, double* arg_val, double* mu,
                                     double* var, double* acq, int32_t* status, int T, int n_max, int n_hint,
                                     int kernel_id, void* ws, size_t ws_bytes, void* stream) {
  return grid_acq_common(false, grid, G, 0, 0, 0.0, 0.0, 1.0, nullptr, nullptr, X, wfrag, alpha, n, theta, best_f,
                         extra_mask, mask_observed, acq_id, acq_param, var_floor, arg_idx, arg_val, mu, var, acq, status,
                         T, n_max, n_hint, kernel_id, ws, ws_bytes, stream);
}

int b200bo_grid_posterior_acq_argmax_lattice(int Gx, int Gy, double x0, double y0, double h, const double* ktab,
                                             const int32_t* ktab_id, const double* X, const double* wfrag,
                                             const double* alpha, const int32_t* n, const double* theta,
                                             const double* best_f, const uint8_t* extra_mask, int mask_observed,
                                             int acq_id, double acq_param, double var_floor, int32_t* arg_idx,
                                             double* arg_val, double* mu, double* var, double* acq, int32_t* status,
                                             int T, int n_max, int n_hint, int kernel_id, void* ws, size_t ws_bytes,
                                             void* stream) {
  return grid_acq_common(true, nullptr, Gx * Gy, Gx, Gy, x0, y0, h, ktab, ktab_id, X, wfrag, alpha, n, theta, best_f,
                         extra_mask, mask_observed, acq_id, acq_param, var_floor, arg_idx, arg_val, mu, var, acq, status,
                         T, n_max, n_hint, kernel_id, ws, ws_bytes, stream);
}

int b200bo_kernel_table(const double* theta, int H, int kernel_id, double h, int TW, int TH, double* ktab, void* stream) {
  using namespace b200bo;
  if (!theta || !ktab) return arg_error("kernel_table: null pointer");
  if (H < 1 || TW < 1 || TH < 1 || !(h > 0.0)) return arg_error("kernel_table: bad sizes H=%d TW=%d TH=%d", H, TW, TH);
  if (kernel_id < 0 || kernel_id > 3) return arg_error("kernel_table: kernel_id=%d", kernel_id);
  dim3 grid((TW * TH + 255) / 256 > 64 ? 64 : (TW * TH + 255) / 256, H > 65535 ? 65535 : H);
  kernel_table_kernel<<<grid, 256, 0, as_stream(stream)>>>(theta, H, kernel_id, h, TW, TH, ktab);
  B200BO_CHECK_LAUNCH("kernel_table");
  return 0;
}

}  // extern "C"
