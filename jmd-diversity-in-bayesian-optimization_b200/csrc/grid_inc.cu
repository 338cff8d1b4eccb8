// K-B (incremental variant): posterior update by ONE new row of L^-1 + acquisition + arg-max.
//
// With fixed hyper-parameters an appended observation only adds row n of L^-1 (gp_append.cu), so for every
// candidate x, with v_i(x) = (L^-1 k*(x))_i and z = L^-1 (y - c):
//     mu_n(x) - c = sum_i v_i z_i      ->  mu  += v_n z_n
//     s - var_n(x) = sum_i v_i^2       ->  ss  += v_n^2
// i.e. the O(n^2) product of grid_acq.cu collapses to one length-n dot product per candidate and iteration;
// (mu - c, ss) per (trial, candidate) persist in HBM (16 B each way per candidate and iteration). The posterior
// is the same function of the data as the full recomputation (same v_i, summed row by row); this is the
// batched analogue of re-using gpytorch's prediction caches instead of re-creating the model (BO.py:424-429).
//
// Roofline: HBM 32 B per candidate-iteration (read + write of the two state words) against
// ~n (look-up + FMA) + acquisition on the issue / FP64 pipes: issue-bound for n >~ 20, HBM-bound below.
// One thread per candidate; a warp's 32 candidates are consecutive lattice nodes, so the kernel-table
// look-up for one observation is a contiguous 256 B read of the signed-offset table.
#include "common.cuh"
#include <limits.h>

namespace b200bo {

constexpr int GI_THREADS = 256;
constexpr int GI_WARPS = GI_THREADS / 32;

struct GridIncParams {
  int Gx, Gy, G;
  double x0, y0, h;
  const double* ktab2;       // (H, 2Gy-1, 2Gx-1) signed-offset kernel table
  const int32_t* ktab_id;    // (T) or null
  const double* X;           // (T, n_max, 2)
  const double* wrow;        // (T, n_max): row n-1 of L^-1
  const double* zvec;        // (T, n_max)
  const int32_t* n;          // (T)
  const double* theta;       // (T, 4)
  const double* best_f;      // (T)
  double* mu_s;              // (T, G)  mu - c
  double* ss_s;              // (T, G)  s - var
  uint32_t* obs_mask;        // (T, ceil(G/32)) observed-candidate bits
  const uint8_t* extra_mask; // (G) or null
  int acquire;               // 0: state update only (building up the initial set)
  int first;                 // 1: this is the first row (state starts from zero, nothing is read)
  int acq_id;
  double acq_param, var_floor;
  double* part_val;
  int32_t* part_idx;
  double* acq_out;           // (T, G) or null
  int32_t* status;
  int n_max, chunks, cols_per_cta, prune;
};

__global__ void __launch_bounds__(GI_THREADS, 4)
grid_inc_kernel(const GridIncParams p) {
  extern __shared__ double sm[];
  const int t = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - t * p.chunks;
  const int n = p.n[t];
  if (n <= 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* wv = sm;                                   // n weights of the new row
  int* ob = reinterpret_cast<int*>(wv + p.n_max);    // n signed-table offsets
  __shared__ double red_val[GI_WARPS];
  __shared__ int red_idx[GI_WARPS];
  __shared__ int red_flag;
  const int W2 = 2 * p.Gx - 1;
  const int C0 = (p.Gy - 1) * W2 + (p.Gx - 1);       // offset of (dx, dy) = (0, 0)
  const int col0 = chunk * p.cols_per_cta;
  const int col_end = min(p.G, col0 + p.cols_per_cta);
  const int words = (p.G + 31) >> 5;
  for (int j = threadIdx.x; j < n; j += GI_THREADS) {
    const double x = p.X[((size_t)t * p.n_max + j) * 2], y = p.X[((size_t)t * p.n_max + j) * 2 + 1];
    const double fx = rint((x - p.x0) / p.h), fy = rint((y - p.y0) / p.h);
    const bool on = (__dadd_rn(p.x0, __dmul_rn(p.h, fx)) == x) && (__dadd_rn(p.y0, __dmul_rn(p.h, fy)) == y) && fx >= 0.0 && fy >= 0.0 &&
                    fx < (double)p.Gx && fy < (double)p.Gy;
    const int xi = on ? (int)fx : 0, yi = on ? (int)fy : 0;
    ob[j] = (yi * W2 + xi - C0) * 8;                 // table BYTE offset = 8 (gy W2 + gx) - ob[j]
    wv[j] = p.wrow[(size_t)t * p.n_max + j];
    if (!on && p.status) atomicOr(&p.status[t], B200BO_ST_OFF_LATTICE);
    if (j == n - 1 && on) {                          // the new observation: mark its candidate as observed
      const int gflat = yi * p.Gx + xi;
      if (gflat >= col0 && gflat < col_end) atomicOr(&p.obs_mask[(size_t)t * words + (gflat >> 5)], 1u << (gflat & 31));
    }
  }
  if (threadIdx.x == 0) red_flag = 0;
  // CTA-uniform table base + 32-bit unsigned byte offsets: one integer subtract per (candidate, observation) pair
  // instead of a 64-bit address computation (5 instructions per pair in the first version, ncu source view)
  // (the base stays the kernel parameter, i.e. a uniform register; the slot offset rides in the 32-bit part)
  const char* tab = reinterpret_cast<const char*>(p.ktab2);
  const unsigned slot_off = (unsigned)(p.ktab_id ? p.ktab_id[t] : 0) * (unsigned)(W2 * (2 * p.Gy - 1) * 8);
  auto look = [tab](unsigned byte_off) { return __ldg(reinterpret_cast<const double*>(tab + byte_off)); };
  const double cst = p.theta[(size_t)t * 4 + 1], sc = p.theta[(size_t)t * 4 + 2];
  const double best_f = p.best_f[t];
  const double zn = p.zvec[(size_t)t * p.n_max + n - 1];
  __syncthreads();

  double best_v = -INFINITY;
  int best_i = INT_MAX;
  int flags = 0;
  const bool prune = p.acq_id == B200BO_ACQ_EI && p.acq_out == nullptr && p.prune;
  double wbest = 0.0, lthr = INFINITY;
  double* mu_s = p.mu_s + (size_t)t * p.G;
  double* ss_s = p.ss_s + (size_t)t * p.G;

  for (int base = col0 + warp * 32; base < col_end; base += GI_WARPS * 32) {
    const int g = base + lane;
    const bool live = g < col_end;
    const int gc = live ? g : col_end - 1;
    const int gy = gc / p.Gx, gx = gc - gy * p.Gx;
    const unsigned gb = slot_off + (unsigned)((gy * W2 + gx) * 8);
    double mu = 0.0, ss = 0.0;
    if (!p.first && live) { mu = mu_s[g]; ss = ss_s[g]; }
    // v = sum_j w_j k(x_j, x): four independent partial sums keep the table loads in flight
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int j = 0;
    for (; j + 4 <= n; j += 4) {
      const double k0 = look(gb - (unsigned)ob[j]), k1 = look(gb - (unsigned)ob[j + 1]);
      const double k2 = look(gb - (unsigned)ob[j + 2]), k3 = look(gb - (unsigned)ob[j + 3]);
      v0 = fma(wv[j], k0, v0);
      v1 = fma(wv[j + 1], k1, v1);
      v2 = fma(wv[j + 2], k2, v2);
      v3 = fma(wv[j + 3], k3, v3);
    }
    for (; j < n; ++j) v0 = fma(wv[j], look(gb - (unsigned)ob[j]), v0);
    const double v = (v0 + v1) + (v2 + v3);
    mu = fma(v, zn, mu);
    ss = fma(v, v, ss);
    if (live) { mu_s[g] = mu; ss_s[g] = ss; }
    if (!p.acquire) continue;                        // uniform
    if (prune) {
      const double nb = warp_max(best_v);
      if (nb > wbest) {
        wbest = nb;
        lthr = log(kInvSqrt2Pi * sqrt(sc) / nb) * (1.0 + 1e-12) + 1e-12;
      }
    }
    if (live) {
      const double m = cst + mu;
      const double var = sc - ss;
      const bool masked = ((p.obs_mask[(size_t)t * words + (g >> 5)] >> (g & 31)) & 1u) || (p.extra_mask && p.extra_mask[g]);
      const double d = m - best_f;
      const bool skip = prune && wbest > 0.0 && d < 0.0 && 0.5 * d * d > fmax(var, p.var_floor) * lthr;
      if (!skip) {
        double a;
        if (p.acq_id == B200BO_ACQ_UCB) a = acquisition<B200BO_ACQ_UCB>(m, var, best_f, p.acq_param, p.var_floor);
        else if (p.acq_id == B200BO_ACQ_PI) a = acquisition<B200BO_ACQ_PI>(m, var, best_f, p.acq_param, p.var_floor);
        else a = acquisition<B200BO_ACQ_EI>(m, var, best_f, p.acq_param, p.var_floor);
        if (p.acq_out) p.acq_out[(size_t)t * p.G + g] = a;
        if (!masked) {
          if (isnan(a) || a == INFINITY) flags |= B200BO_ST_NONFINITE;
          if (!isnan(a) && better(a, g, best_v, best_i)) { best_v = a; best_i = g; }
        }
      }
    }
  }
  if (!p.acquire) return;
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
    for (int w = 1; w < GI_WARPS; ++w)
      if (better(red_val[w], red_idx[w], bv, bi)) { bv = red_val[w]; bi = red_idx[w]; }
    p.part_val[(size_t)t * p.chunks + chunk] = bv;
    p.part_idx[(size_t)t * p.chunks + chunk] = bi;
    if (red_flag && p.status) atomicOr(&p.status[t], red_flag);
  }
}

// signed-offset table: ktab2[slot][(dy + Gy-1) (2Gx-1) + dx + Gx-1] = s k(h sqrt(dx^2 + dy^2) / l), same device
// function and operands as kernel_table_kernel / the generic path (bit-identical values)
__global__ void __launch_bounds__(256)
kernel_table_signed_kernel(const double* __restrict__ theta, int kernel_id, double h, int Gx, int Gy,
                           double* __restrict__ ktab2) {
  const int slot = blockIdx.y;
  const double sc = theta[(size_t)slot * 4 + 2], ell = theta[(size_t)slot * 4 + 3];
  const KernParams kp = make_kern_params(sc, ell, kernel_id);
  const int W2 = 2 * Gx - 1, H2 = 2 * Gy - 1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < W2 * H2; e += gridDim.x * blockDim.x) {
    const int r = e / W2, c = e - r * W2;
    const double dx = h * (double)abs(c - (Gx - 1)), dy = h * (double)abs(r - (Gy - 1));
    const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    ktab2[(size_t)slot * W2 * H2 + e] = kernel_from_r2_dyn(kernel_id, r2, sc, ell, kp.a_scale);
  }
}

inline int inc_chunks(int T, int G) {
  int max_chunks = (G + GI_THREADS - 1) / GI_THREADS;
  int want = (8 * 148 + T - 1) / T;
  if (want < 1) want = 1;
  return want > max_chunks ? max_chunks : want;
}

}  // namespace b200bo

extern "C" {

int b200bo_kernel_table_signed(const double* theta, int H, int kernel_id, double h, int Gx, int Gy, double* ktab2,
                               void* stream) {
  using namespace b200bo;
  if (!theta || !ktab2) return arg_error("kernel_table_signed: null pointer");
  if (H < 1 || H > 65535 || Gx < 1 || Gy < 1 || !(h > 0.0)) return arg_error("kernel_table_signed: bad sizes");
  if (kernel_id < 0 || kernel_id > 3) return arg_error("kernel_table_signed: kernel_id=%d", kernel_id);
  dim3 grid(128, H);
  kernel_table_signed_kernel<<<grid, 256, 0, as_stream(stream)>>>(theta, kernel_id, h, Gx, Gy, ktab2);
  B200BO_CHECK_LAUNCH("kernel_table_signed");
  return 0;
}

size_t b200bo_grid_inc_workspace_bytes(int T, int G) {
  if (T <= 0 || G <= 0) return 0;
  return (size_t)T * b200bo::inc_chunks(T, G) * (sizeof(double) + sizeof(int32_t)) + 16;
}

int b200bo_grid_posterior_update_acq_argmax(int Gx, int Gy, double x0, double y0, double h, const double* ktab2,
                                            const int32_t* ktab_id, const double* X, const double* wrow,
                                            const double* zvec, const int32_t* n, const double* theta,
                                            const double* best_f, double* mu_state, double* ss_state,
                                            uint32_t* obs_mask, const uint8_t* extra_mask, int first, int acquire,
                                            int acq_id, double acq_param, double var_floor, int32_t* arg_idx,
                                            double* arg_val, double* acq, int32_t* status, int T, int n_max, void* ws,
                                            size_t ws_bytes, void* stream) {
  using namespace b200bo;
  const char* who = "grid_posterior_update_acq_argmax";
  if (T < 0 || Gx < 1 || Gy < 1 || n_max < 1 || n_max > B200BO_NMAX || !(h > 0.0)) return arg_error("%s: bad sizes", who);
  if (T == 0) return 0;
  if (!ktab2 || !X || !wrow || !zvec || !n || !theta || !best_f || !mu_state || !ss_state || !obs_mask)
    return arg_error("%s: null pointer", who);
  if (acquire && (!arg_idx || !ws)) return arg_error("%s: arg_idx / workspace missing", who);
  if (acq_id < 0 || acq_id > 2) return arg_error("%s: acq_id=%d", who, acq_id);
  if (acq_id == B200BO_ACQ_UCB && !(acq_param >= 0.0)) return arg_error("%s: UCB beta < 0", who);
  const int G = Gx * Gy;
  if (acquire && ws_bytes < b200bo_grid_inc_workspace_bytes(T, G)) return arg_error("%s: workspace too small", who);
  cudaStream_t st = as_stream(stream);
  GridIncParams p;
  p.Gx = Gx; p.Gy = Gy; p.G = G; p.x0 = x0; p.y0 = y0; p.h = h;
  p.ktab2 = ktab2; p.ktab_id = ktab_id; p.X = X; p.wrow = wrow; p.zvec = zvec; p.n = n; p.theta = theta;
  p.best_f = best_f; p.mu_s = mu_state; p.ss_s = ss_state; p.obs_mask = obs_mask; p.extra_mask = extra_mask;
  p.acquire = acquire; p.first = first; p.acq_id = acq_id;
  p.acq_param = acq_id == B200BO_ACQ_UCB ? sqrt(acq_param) : acq_param;
  p.var_floor = var_floor;
  p.chunks = inc_chunks(T, G);
  p.cols_per_cta = (((G + p.chunks - 1) / p.chunks + GI_THREADS - 1) / GI_THREADS) * GI_THREADS;
  p.part_val = reinterpret_cast<double*>(ws);
  p.part_idx = ws ? reinterpret_cast<int32_t*>(p.part_val + (size_t)T * p.chunks) : nullptr;
  p.acq_out = acq; p.status = status; p.n_max = n_max;
  static const int prune_env = [] { const char* pr = getenv("B200BO_KB_PRUNE"); return pr ? atoi(pr) : 1; }();
  p.prune = prune_env;
  size_t smem = (size_t)n_max * (sizeof(double) + sizeof(int));
  grid_inc_kernel<<<(unsigned)((size_t)T * p.chunks), GI_THREADS, smem, st>>>(p);
  B200BO_CHECK_LAUNCH(who);
  if (acquire) {
    launch_argmax_finish(p.part_val, p.part_idx, p.chunks, T, n, arg_idx, arg_val, status, st);
    B200BO_CHECK_LAUNCH(who);
  }
  return 0;
}

}  // extern "C"
