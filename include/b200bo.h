/*
 * b200bo — C ABI of the B200-native engine for the data-parallel hot path of
 * IDEALLab/JMD-Diversity-in-Bayesian-Optimization.
 *
 * The reference is pure Python and has no FFI of its own; every entry point below
 * names the reference call site it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _h (host);
 *  - the caller owns every buffer; the library allocates nothing persistent;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream), re-entrant and CUDA-graph capturable;
 *  - return 0 = OK, < 0 = argument error (nothing launched), > 0 = cudaError_t;
 *    b200bo_last_error_string() describes the last non-zero return of this thread;
 *  - numerical failures are per trial in status[] (B200BO_ST_* bits), never a return code.
 *
 *  theta = (noise sigma^2, mean constant c, outputscale s, lengthscale l), ACTUAL
 *  values (the Python layer applies gpytorch's softplus to raw_outputscale /
 *  raw_lengthscale; noise is untransformed, BO.py:33-68 + SURVEY.md A.1).
 */
#ifndef B200BO_H
#define B200BO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200BO_KERNEL_RBF 0
#define B200BO_KERNEL_MATERN12 1
#define B200BO_KERNEL_MATERN32 2
#define B200BO_KERNEL_MATERN52 3

#define B200BO_ACQ_EI 0
#define B200BO_ACQ_UCB 1
#define B200BO_ACQ_PI 2

#define B200BO_ST_JITTER 1      /* Cholesky needed diagonal jitter (psd_safe_cholesky) */
#define B200BO_ST_CHOL_FAIL 2   /* Cholesky failed after 3 jitter levels              */
#define B200BO_ST_NONFINITE 4   /* NaN / +inf acquisition value seen                   */
#define B200BO_ST_ALL_MASKED 8  /* no selectable candidate                             */
#define B200BO_ST_OFF_LATTICE 16 /* lattice path given an off-lattice observation      */
#define B200BO_ST_FALLBACK 32   /* b200bo_bo_fixed_loop handed the trial back untouched (non-positive pivot or
                                   off-lattice start): run it through the per-iteration kernels */

#define B200BO_NMAX 128         /* max observations per trial in the GP kernels        */

const char* b200bo_version(void);
const char* b200bo_last_error_string(void);
/* sm_count / cc of the current device; returns cudaError_t */
int b200bo_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* FP64 FMA issue-rate probe (roofline denominator measured on the GPU the benchmark runs on):
 * launches blocks x 256 threads x iters x 8 dependent-chain FMAs, writes blocks*256 doubles to out,
 * returns the flop count issued (> 0) or -(error code). */
long long b200bo_fp64_fma_probe(double* out, int blocks, int iters, void* stream);
/* shared-memory read bandwidth (conflict-free LDS.64, 1024 threads per CTA): out (blocks*1024) f64 scratch; returns the
 * bytes read (< 0: error). Roofline denominator of b200bo_bo_fixed_loop. */
long long b200bo_smem_lds_probe(double* out, int blocks, int iters, void* stream);

/* ---- K-D  ranked-DPP set score -------------------------------------------------
 * replaces diverse_data_generator.twoD_score looped over the superset,
 * Src/data_gen.py:484-497 called from :403.
 * pool (P,2) f64; idx (N,k) int32 rows index pool; logdet (N) f64 out.
 * logdet[i] = log |det exp(-gamma * D)|, D_ab = (sqrt(|x_a-x_b|^2))^2, from the pivots of an LDL^T elimination:
 * finite on nearly singular S like numpy's slogdet (which the reference calls), -inf when a pivot is exactly 0. */
int b200bo_dpp_logdet_score(const double* pool, int pool_size, const int32_t* idx, long long N, int k,
                            double gamma, double* logdet, void* stream);

/* Lattice variant of K-D: the pool is the constructX lattice (Src/data_gen.py:294-303), node g <-> (g % Gx, g / Gx),
 * x = x0 + h ix with coordinates on which differences are exact (integers or dyadic steps), so an entry of the
 * L-ensemble depends on (|dx|, |dy|) only and is read from table (Gy,Gx) f64 built by b200bo_dpp_table for this gamma
 * (same arithmetic as the generic kernel: values bit-identical). k <= 12; idx (N,k) int32 flat node indices. */
int b200bo_dpp_table(int Gx, int Gy, double h, double gamma, double* table, void* stream);
int b200bo_dpp_logdet_score_lattice(const double* table, int Gx, int Gy, const int32_t* idx, long long N, int k,
                                    double* logdet, void* stream);

/* ---- ranked-DPP ordering on the device (csrc/dpp_rank.cu) ---------------------------------------
 * b200bo_dpp_rank replaces ind = np.argsort(scores); dist = superset[ind]; scores[ind] (Src/data_gen.py:409-415), made
 * deterministic: ascending (score, set index) — a stable LSD radix sort on the order-preserving 64-bit image of the
 * f64 scores (NaN last, -0 == +0). z (N) f64 in; pos (N,k) int32 pool indices of the sets; pool (pool_size,2) f64.
 * Out (each may be NULL): z_sorted (N), order (N) int32 = the permutation, pos_sorted (N,k), dist (N,k,2) f64 = the
 * sets' points in rank order. ws: >= b200bo_dpp_rank_workspace_bytes(N). */
size_t b200bo_dpp_rank_workspace_bytes(long long N);
int b200bo_dpp_rank(const double* z, long long N, const int32_t* pos, int k, const double* pool, int pool_size,
                    double* z_sorted, int32_t* order, int32_t* pos_sorted, double* dist, void* ws, size_t ws_bytes,
                    void* stream);
/* compare_subDPP's rank, np.sum(scores < z) (data_gen.py:481): count (1) uint64 out. */
int b200bo_dpp_count_less(const double* scores, long long N, double z, uint64_t* count, void* stream);
/* sub_dpp_sample's np.setdiff1d(range(lo, hi), acquiredsets)[r] (data_gen.py:445-466): the r-th (0-based) rank of the
 * window [lo, hi) not yet acquired; acquired_mask ((N+31)/32 words, bit = rank) is updated; loc (1) int32 out (-1: fewer
 * than r+1 ranks left); set_out (k,2) f64 = that set's points (NULL to skip). r is the host Generator's draw. */
int b200bo_dpp_window_pick(uint32_t* acquired_mask, long long N, int lo, int hi, int r, const int32_t* pos_sorted, int k,
                           const double* pool, int pool_size, int32_t* loc, double* set_out, void* stream);
/* gammacalc (data_gen.py:758-842) scores the same N sets under every gamma of its range: one launch, the squared
 * distances computed once per set. gammas (n_gamma) f64 on the device; logdet (n_gamma, N) f64 out; k <= 12. Values are
 * bit-identical to b200bo_dpp_logdet_score called once per gamma. */
int b200bo_dpp_logdet_score_multigamma(const double* pool, int pool_size, const int32_t* idx, long long N, int k,
                                       const double* gammas, int n_gamma, double* logdet, void* stream);

/* mean / population-sd / z-scores of the scoring block of sub_dpp, data_gen.py:406-408.
 * stats (2) f64 out = {mean, sd}; z (N) f64 out; ws: >= b200bo_dpp_stats_workspace_bytes(N). */
size_t b200bo_dpp_stats_workspace_bytes(long long N);
int b200bo_dpp_zscore(const double* logdet, long long N, double* stats, double* z, void* ws, size_t ws_bytes,
                      void* stream);

/* combinadic un-ranking: m-th k-subset of range(n) in lexicographic order, 128-bit ranks.
 * replaces combination_finder.find_elm (Src/Combination.py:131-170) looped at data_gen.py:392.
 * binom_lo/hi (k, n+1) uint64: C(v,b) for b = 1..k (row b-1), v = 0..n; m_lo/hi (N) ranks;
 * total = C(n,k) (must be < 2^128). out (N,k) int32 ascending indices; rank >= total -> row of -1
 * (the reference raises there). */
int b200bo_unrank_u128(const uint64_t* binom_lo, const uint64_t* binom_hi, const uint64_t* m_lo, const uint64_t* m_hi,
                       uint64_t total_lo, uint64_t total_hi, int n, int k, long long N, int32_t* out, void* stream);

/* element ranks of the ranked-DPP sampler, drawn as CPython draws them: for every seeds[i] (int64, the values
 * np.random.randint hands to random.seed) out[i] = the value of  random.seed(seeds[i]); random.randint(0, n - 1)
 * under CPython 3.9 rules (non-int seed -> (size_t) hash; MT19937 init_by_array; _randbelow by rejection on
 * getrandbits(n.bit_length())). Replaces the loop at Src/data_gen.py:868-871. n = (n_hi, n_lo) < 2^128, > 0.
 * out_lo / out_hi (N) uint64; fail (1) int32 is incremented for elements that needed more than 56 rejection
 * rounds (probability < 2^-56 each; redo those on the host). Bit-exact. */
int b200bo_mt_seeded_randbelow(const int64_t* seeds, long long N, uint64_t n_lo, uint64_t n_hi, uint64_t* out_lo,
                               uint64_t* out_hi, int32_t* fail, void* stream);

/* ---- K-E  Wildcat Wells --------------------------------------------------------
 * replaces objectives.wildcatwells (Src/objectives.py:107-218, N = 1 peak, d = 2) and the
 * 10 201 x 3 SimplexNoise.simplexNoise calls inside it (Src/SimpleNoise.py:85-172).
 * Per surface s: hseed[s] = randHash(seed) (SimpleNoise.py:20), vecs (S,4,2) int8 = the
 * gradient table shuffled by random.Random(hseed) (SimpleNoise.py:81-83; 4 entries, host),
 * peak_xy (S,2) f64 (objectives.py:139-149), smooth / rug_amp (S) f64.
 * surf (S,101,101) f32 out, row-major [iy][ix], max == 100. */
int b200bo_wildcat_surface(const uint64_t* hseed, const int8_t* vecs, const double* peak_xy, const double* smooth,
                           const double* rug_amp, float* surf, int S, void* stream);
/* objective evaluation f(x) of generate_cont() (Src/objectives.py:322-351, scipy LinearNDInterpolator over the
 * 101x101 nodes, :349-350): surf (S,101,101) f32, surf_id (M) int32 or NULL (surface 0), xy (M,2) f64 -> out (M) f64.
 * diag_mask (100*100) uint8, [iy*100 + jx] = 1 when Qhull's Delaunay triangulation cuts cell
 * [jx,jx+1]x[iy,iy+1] along (jx,iy)-(jx+1,iy+1), 0 along the other diagonal (exported once by
 * tools/make_diag_mask.py — the split depends on the node coordinates only); NULL = main diagonal everywhere.
 * Node values are exact, half-node values bit-identical to the interpolator's; NaN outside [0,100]^2. */
int b200bo_wildcat_eval(const float* surf, const int32_t* surf_id, const uint8_t* diag_mask, const double* xy,
                        double* out, long long M, void* stream);
/* raw simplex noise at M points (M,2) f64 for one (hseed, vecs): SimpleNoise.py:85-172 */
int b200bo_simplex_noise(uint64_t hseed, const int8_t* vecs_h, const double* xy, double* out, long long M,
                         void* stream);

/* ---- K-A  kernel matrix + Cholesky --------------------------------------------
 * replaces gpytorch ScaleKernel(Matern|RBF) + psd_safe_cholesky + the prediction-strategy
 * caches built by SingleTaskGP (BO.py:65,142,176).
 * X (T,n_max,2) f64, y (T,n_max) f64, n (T) int32 (n[t] <= n_max <= B200BO_NMAX),
 * theta (T,4) f64.  Outputs (any may be NULL): L (T,n_max,n_max) lower, row-major;
 * Linv (T,n_max,n_max) = L^-1; alpha (T,n_max) = Khat^-1 (y - c); logdet (T) = log|Khat|;
 * wfrag (T, b200bo_wfrag_doubles(n_max)) = L^-1 in the MMA-fragment layout K-B consumes.
 * status (T) int32 is OR-ed into.  n[t] <= 0 marks an inactive trial (skipped).
 * n_hint: upper bound of n[t] over this launch (0 = n_max); sizes shared memory / picks the
 * kernel variant, so a loop whose n grows by one per iteration does not pay for n_max early. */
size_t b200bo_wfrag_doubles(int n_max);
int b200bo_gp_assemble_chol(const double* X, const double* y, const int32_t* n, const double* theta, double* L,
                            double* Linv, double* alpha, double* logdet, double* wfrag, int32_t* status, int T,
                            int n_max, int n_hint, int kernel_id, void* stream);

/* K-A append variant: one observation was appended (it is row n[t]-1) and theta is unchanged —
 * the fixed-hyper-parameter loops re-create the same model plus one point per iteration
 * (Src/BO.py:416-429). Extends L^-1 (wfrag, in/out), z = L^-1 (y - c) (zvec (T,n_max), in/out) and
 * recomputes alpha in O(n^2). wfrag rows beyond the old n must be zero (zero-initialise the buffer).
 * n_redo (T) int32 out: 0, or n[t] for trials whose new pivot was not positive — pass it as `n` to
 * b200bo_gp_assemble_chol (+ b200bo_gp_zvec) right after to refactorise those with the jitter rule.
 * wrow_out (T,n_max) optional: dense copy of the new row of L^-1 (input of the incremental K-B).
 * b200bo_gp_zvec: z = L^-1 (y - c) from a fresh factorisation. */
int b200bo_gp_chol_append(const double* X, const double* y, const int32_t* n, const double* theta, double* wfrag,
                          double* zvec, double* alpha, double* wrow_out, int32_t* n_redo, int32_t* status, int T,
                          int n_max, int kernel_id, void* stream);
int b200bo_gp_zvec(const double* y, const int32_t* n, const double* theta, const double* wfrag, double* zvec, int T,
                   int n_max, void* stream);

/* ---- K-B  fused grid posterior + acquisition + arg-max --------------------------
 * replaces ExpectedImprovement(model,best_f) + optimize_acqf over the box (BO.py:158-161,
 * 72-83) by a dense arg-max over the candidate grid (the north-star redesign).
 * grid (G,2) f64 candidates; for trial t: mu = c + k*^T alpha, var = s - |L^-1 k*|^2,
 * acq = EI/UCB/PI(mu, max(var,var_floor), best_f[t]); candidates with r^2 == 0 to an
 * observation are masked when mask_observed != 0; extra_mask (G) uint8 optional (shared).
 * arg_idx (T) int32 / arg_val (T) f64 out: first index of the maximum, -1 if none.
 * Optional debug fields mu/var/acq (T,G) f64 (NULL to skip).
 * ws: >= b200bo_grid_acq_workspace_bytes(T, G). */
size_t b200bo_grid_acq_workspace_bytes(int T, int G);
/* Column chunks (CTAs) per trial a K-B launch of T trials over G candidates with at most n_cap observations will use
 * (lattice != 0: the lattice entry). Host-only, no CUDA call: for tests and capacity planning. Never more than the
 * workspace above is sized for. */
int b200bo_grid_acq_chunks(int T, int G, int n_cap, int lattice);
int b200bo_grid_posterior_acq_argmax(const double* grid, int G, const double* X, const double* wfrag,
                                     const double* alpha, const int32_t* n, const double* theta,
                                     const double* best_f, const uint8_t* extra_mask, int mask_observed,
                                     int acq_id, double acq_param, double var_floor, int32_t* arg_idx,
                                     double* arg_val, double* mu, double* var, double* acq, int32_t* status,
                                     int T, int n_max, int n_hint, int kernel_id, void* ws, size_t ws_bytes,
                                     void* stream);

/* Lattice variant of K-B: candidates are the nodes of a regular Gx x Gy lattice, candidate
 * g <-> (ix = g % Gx, iy = g / Gx), x = x0 + h ix, y = y0 + h iy (data_gen.constructX,
 * Src/data_gen.py:294-303), and every observation is a node of the same lattice (the ranked-DPP
 * initial sets and all arg-max queries are). The kernel value then depends on (|dx|,|dy|) only and is
 * read from ktab[slot][|dy| * Gx + |dx|] instead of being recomputed G*n times; values are
 * bit-identical to the generic path. ktab (H,Gy,Gx) f64 from b200bo_kernel_table; ktab_id (T) int32
 * slot per trial or NULL (slot 0). An off-lattice observation sets B200BO_ST_OFF_LATTICE. */
int b200bo_kernel_table(const double* theta, int H, int kernel_id, double h, int TW, int TH, double* ktab, void* stream);
int b200bo_grid_posterior_acq_argmax_lattice(int Gx, int Gy, double x0, double y0, double h, const double* ktab,
                                             const int32_t* ktab_id, const double* X, const double* wfrag,
                                             const double* alpha, const int32_t* n, const double* theta,
                                             const double* best_f, const uint8_t* extra_mask, int mask_observed,
                                             int acq_id, double acq_param, double var_floor, int32_t* arg_idx,
                                             double* arg_val, double* mu, double* var, double* acq, int32_t* status,
                                             int T, int n_max, int n_hint, int kernel_id, void* ws, size_t ws_bytes,
                                             void* stream);

/* Incremental variant of K-B for fixed hyper-parameters on a lattice: the appended observation adds one
 * row w of L^-1 (b200bo_gp_chol_append, wrow_out), so per candidate  v = sum_j w_j k(x_j, x),
 * mu_state += v * z_n,  ss_state += v^2  (mu = c + mu_state, var = s - ss_state) — one length-n dot product
 * instead of the O(n^2) product; the two state words per (trial, candidate) live in HBM (T,G) f64 each.
 * ktab2 (H, 2Gy-1, 2Gx-1) from b200bo_kernel_table_signed; obs_mask (T, ceil(G/32)) uint32 observed-candidate
 * bits (zero-initialise; the kernel sets the bit of observation n-1). first != 0: state starts from zero.
 * acquire == 0: state update only (used while the initial set is fed row by row).
 * Same posterior as the full kernels (rows summed in order); replaces the same reference lines. */
int b200bo_kernel_table_signed(const double* theta, int H, int kernel_id, double h, int Gx, int Gy, double* ktab2,
                               void* stream);
size_t b200bo_grid_inc_workspace_bytes(int T, int G);
int b200bo_grid_posterior_update_acq_argmax(int Gx, int Gy, double x0, double y0, double h, const double* ktab2,
                                            const int32_t* ktab_id, const double* X, const double* wrow,
                                            const double* zvec, const int32_t* n, const double* theta,
                                            const double* best_f, double* mu_state, double* ss_state,
                                            uint32_t* obs_mask, const uint8_t* extra_mask, int first, int acquire,
                                            int acq_id, double acq_param, double var_floor, int32_t* arg_idx,
                                            double* arg_val, double* acq, int32_t* status, int T, int n_max, void* ws,
                                            size_t ws_bytes, void* stream);

/* ---- K-F  the whole fixed-hyper-parameter BO loop of a trial in one persistent CTA -----------------
 * replaces BOloop_fixparam's loop body (Src/BO.py:404-446: ExpectedImprovement -> optimize_acqf -> objective ->
 * torch.cat append -> yopt / yall.append -> re-created model with the same hyper-parameters) for ALL iterations of a
 * trial: one launch instead of b200bo_gp_chol_append + b200bo_grid_posterior_update_acq_argmax + b200bo_bo_update per
 * iteration. Same posterior recurrences as those entries (rank-1 row of L^-1, per-candidate (mu - c, s - var) folded
 * row by row); the kernel table lives in shared memory, the per-candidate state in a per-CTA scratch (L2-resident).
 * ktabY (H, 2Gy-1, Gx) f64 from b200bo_kernel_table_ysigned: ktabY[slot][(dy + Gy-1) Gx + |dx|]; ktab_id (T) or NULL;
 * H = number of table slots; for H > 1: order (T) int32 = the trials sorted by slot and slot_start (H+1) int32 = first
 * position of every slot in order (all three may be NULL for H == 1). Two kernel variants, chosen by what fits: two
 * trials per CTA (two 512-thread halves on ONE unsigned table, 80 KB for 100 x 100: the latency-bound phases of one
 * trial overlap the look-up loop of the other; B200BO_KF_DUAL=0 disables it) or one trial per CTA on the y-signed table.
 * In: X (T,n_max,2), y (T,n_max), n (T) = initial observations (>= 1). Out: X, y, n (observations at the end, positive),
 * best_f (T), yall (T,max_iter+2) padded as BO.py:262-271, picked (T,max_iter) (-1 = none), nit (T), status (T).
 * A trial whose start is off the lattice or whose appended pivot is not positive comes back with
 * B200BO_ST_FALLBACK set, n = its initial count, nit = 0: the caller runs it through the per-iteration entries
 * (psd_safe_cholesky's jitter rule lives in b200bo_gp_assemble_chol).
 * b200bo_bo_fixed_loop_supported: does the lattice's table (+ n_max rows of L^-1) fit the 227 KB of shared memory
 * (100 x 100 with n_max <= 116 does; 200 x 200 does not). ws: >= b200bo_bo_fixed_loop_workspace_bytes(T, Gx, Gy). */
int b200bo_kernel_table_ysigned(const double* theta, int H, int kernel_id, double h, int Gx, int Gy, double* ktabY,
                                void* stream);
int b200bo_bo_fixed_loop_supported(int Gx, int Gy, int n_max);
int b200bo_bo_fixed_loop_ctas(int T);
size_t b200bo_bo_fixed_loop_workspace_bytes(int T, int Gx, int Gy);
int b200bo_bo_fixed_loop(int Gx, int Gy, double x0, double y0, double h, const double* ktabY, const int32_t* ktab_id,
                         const int32_t* order, const int32_t* slot_start, int H, const float* obj, const int32_t* surf_id,
                         double* X, double* y, int32_t* n, const double* theta, double* best_f, double* yall, int32_t* picked,
                         int32_t* nit, int32_t* status, int T, int n_max, int max_iter, int acq_id, double acq_param,
                         double var_floor, double tol, double optimal_y, void* ws, size_t ws_bytes, void* stream);

/* ---- K-C  marginal log-likelihood value and gradient ----------------------------
 * replaces ExactMarginalLogLikelihood + autograd inside fit_gpytorch_model (BO.py:144,189).
 * raw (T,4) f64 = (noise, c, raw_outputscale, raw_lengthscale); value (T) f64 =
 * [log N(y; c, Khat) + log Gamma(noise; 1.1, 0.05)] / n ; grad (T,4) f64. */
int b200bo_mll_value_grad(const double* X, const double* y, const int32_t* n, const double* raw, double* value,
                          double* grad, int32_t* status, int T, int n_max, int n_hint, int kernel_id, void* stream);

/* ---- batched hyper-parameter fit driver -------------------------------------------------
 * replaces botorch.fit.fit_gpytorch_model (scipy L-BFGS-B over the 4 raw parameters with the
 * bound noise >= 1e-4), Src/BO.py:144,189. One optimiser state machine per trial; the caller
 * alternates b200bo_mll_value_grad(raw = xt, n = n_eval) and b200bo_lbfgs_step until n_eval is 0
 * everywhere. state: b200bo_lbfgs_state_bytes(T) bytes; xt (T,4) current evaluation points (in: start);
 * raw0 (4) shared or (T,4) per-trial start; n_obs (T) observations per trial (<= 0: trial skipped);
 * outputs on convergence: theta_out (T,4) actual values, raw_out (T,4), mll_out (T), iters_out (T,2)
 * = {iterations, evaluations}. */
size_t b200bo_lbfgs_state_bytes(int T);
int b200bo_lbfgs_init(void* state, double* xt, const double* raw0, int raw0_per_trial, int32_t* n_eval,
                      const int32_t* n_obs, int T, double lb_noise, void* stream);
int b200bo_lbfgs_step(void* state, double* xt, const double* value, const double* grad, int32_t* n_eval,
                      const int32_t* n_obs, double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out,
                      int T, double lb_noise, double gtol, double ftol, int max_iter, int max_ls, void* stream);

/* Fused fit: the whole L-BFGS run of every trial in ONE launch (one CTA per trial keeps the observations in
 * shared memory and alternates marginal-likelihood evaluation and optimiser step until convergence) — same
 * arithmetic and optimiser as the b200bo_mll_value_grad + b200bo_lbfgs_step rounds, without the per-round
 * launches and without waiting for the slowest trial. max_rounds bounds the evaluations per trial. */
int b200bo_gp_fit(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                  double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                  int n_max, int n_hint, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter,
                  int max_ls, int max_rounds, void* stream);
/* the same with a launch order: order (T) int32 = a permutation of the trials, CTA b fits trial order[b] (NULL = identity).
 * CTAs start in that order, so sorting by expected evaluation count, longest first (e.g. the counts of the previous
 * iteration's fit), ends the launch with its shortest fits. Results do not depend on it. */
int b200bo_gp_fit_ordered(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                          double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                          int n_max, int n_hint, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter,
                          int max_ls, int max_rounds, const int32_t* order, void* stream);

/* ---- K-C for large training sets -------------------------------------
 * replaces fit_gpytorch_model on the 1 010 ... 1 209-point models of the hyper-parameter extraction experiment
 * (Scripts/Experiment3-2-1.py:63-87: HPextract -> BO.initialize_model(nu=1.5) -> fit_gpytorch_model).
 * Same objective, gradient and optimiser as b200bo_gp_fit / b200bo_mll_value_grad, for n[t] <= n_max <= 8192: one
 * fit per CTA, the matrix lives in the caller's workspace (b200bo_gp_fit_large_workspace_bytes(T, n_max) bytes),
 * blocked symmetric sweep with a register-tiled rank-64 update. X (T,n_max,2), y (T,n_max), n (T). */
size_t b200bo_gp_fit_large_workspace_bytes(int T, int n_max);
int b200bo_gp_fit_large(const double* X, const double* y, const int32_t* n, const double* raw0, int raw0_per_trial,
                        double* theta_out, double* raw_out, double* mll_out, int32_t* iters_out, int32_t* status, int T,
                        int n_max, int kernel_id, double lb_noise, double gtol, double ftol, int max_iter, int max_ls,
                        int max_rounds, void* ws, size_t ws_bytes, void* stream);
int b200bo_mll_value_grad_large(const double* X, const double* y, const int32_t* n, const double* raw, double* value,
                                double* grad, int32_t* status, int T, int n_max, int kernel_id, void* ws, size_t ws_bytes,
                                void* stream);

/* ---- bootstrap statistics of gap curves ------------------------------
 * replaces utils.bootstrap (Src/utils.py:39-115; callers results.py:172-184,787-802).
 * data (N,C) f64 row-major = N trials x C positions (C = 1 for a vector of cumulative gaps);
 * idx (B,N) int32 = for replicate b, the rows drawn by random.choices (the host draws them from CPython's
 * generator, so the replicates are the reference's); boot (B,C) out: boot[b][c] = np.mean of the drawn rows at
 * position c with numpy's 1-D pairwise summation order (bit-identical to the reference's list means). */
int b200bo_bootstrap_means(const double* data, int N, int C, const int32_t* idx, int B, double* boot, void* stream);
/* per column of boot (B,C): out (1+nq, C) = { np.mean over the B replicates (pairwise),
 * np.percentile(100 q[0]), ... } with numpy's default "linear" interpolation; q (nq) f64 DEVICE = percent / 100.
 * B <= 16384 (one CTA sorts a column in shared memory); a column containing NaN yields NaN percentiles. */
int b200bo_column_mean_percentiles(const double* boot, int B, int C, const double* q, int nq, double* out, void* stream);

/* ---- batched BO loop bookkeeping (one thread per trial) ---------------------------
 * replaces the Python statements between two acquisitions in BOloop / BOloop_fixparam:
 * objective look-up (Src/BO.py:104), append (:164-165), yopt / yall.append (:168-169),
 * stop test (:152,241) and the post-loop padding of yall to max_iter+2 (:262,267-271).
 * State: X (T,n_max,2), y (T,n_max), n (T) int32 (n < 0 = finished, |n| observations),
 * best_f (T), yall (T,max_iter+2), picked (T,max_iter) int32 (-1 = none), nit (T) int32.
 * obj (S,G) f32 = objective at every candidate for S surfaces; surf_id (T) int32. */
int b200bo_bo_init(const double* y, int32_t* n, double* best_f, double* yall, int32_t* picked, int32_t* nit, int T,
                   int n_max, int max_iter, double tol, double optimal_y, void* stream);
int b200bo_bo_update(const double* grid, int G, const float* obj, const int32_t* surf_id, const int32_t* arg_idx,
                     double* X, double* y, int32_t* n, double* best_f, double* yall, int32_t* picked, int32_t* nit,
                     int T, int n_max, int max_iter, int iter, double tol, double optimal_y, void* stream);
int b200bo_bo_finalize(double* yall, const int32_t* nit, int32_t* n, int T, int max_iter, void* stream);
/* observation counts for the hyper-parameter fit before acquisition `iter` (BOloop, Src/BO.py:176-189,252): live
 * trials, plus the trials that stopped in iteration iter-1 — the reference refits after every append, the last one
 * included, and records that model. n_fit (T) int32 out: |n| for those trials, 0 otherwise. */
int b200bo_bo_fit_counts(const int32_t* n, const int32_t* nit, int32_t* n_fit, int T, int iter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200BO_H */
