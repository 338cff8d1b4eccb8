// Ranked-DPP ordering on the device: the sort, gather, rank and window steps that
// follow the scoring kernel (csrc/dpp.cu) in sub_dpp / sub_dpp_sample / compare_subDPP, and the multi-gamma scorer
// behind gammacalc.
//
//   b200bo_dpp_rank            replaces  ind = np.argsort(scores); dist = superset[ind]; scores = scores[ind]
//                                        (Src/data_gen.py:409-415) — made deterministic: ascending (score, set index)
//   b200bo_dpp_count_less      replaces  np.sum(scores < z)                          (compare_subDPP, data_gen.py:481)
//   b200bo_dpp_window_pick     replaces  np.setdiff1d(range(lo, hi), acquiredsets)[r] + the look-up of that set
//                                        (sub_dpp_sample, data_gen.py:445-466; r is the host Generator's draw)
//   b200bo_dpp_logdet_score_multigamma   the scores of the same N sets under every gamma of gammacalc's range in one
//                                        launch (data_gen.py:758-842 re-scores them 256 times)
//
// The sort is a hand-written stable LSD radix sort of the 64-bit order-preserving image of the f64 scores with the set
// index as payload: 8 passes of 8 bits, per pass a per-tile digit histogram, a per-digit exclusive scan over the tiles and a
// stable scatter (per-warp digit counters + __match_any_sync ranks: a warp owns a contiguous run of the tile, rounds of
// 32 consecutive elements, so equal keys keep their order). HBM-bound: 2 x 8 passes x 12 B per set plus the gather.
#include "common.cuh"

namespace b200bo {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;      // 2048 elements per CTA
constexpr int RS_BINS = 256;

// ascending order of the doubles == ascending order of the keys; -0.0 == +0.0 (ties go by index), every NaN last
__device__ __forceinline__ unsigned long long f64_sort_key(double v) {
  if (v != v) return ~0ull;
  v += 0.0;                                          // -0.0 -> +0.0
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
rs_keys_kernel(const double* __restrict__ z, long long N, unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    keys[i] = f64_sort_key(z[i]);
    vals[i] = (int32_t)i;
  }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long* __restrict__ keys, long long N, int shift, int nblocks, unsigned* __restrict__ hist) {
  __shared__ unsigned h[RS_BINS];
  h[threadIdx.x] = 0u;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const long long e = base + r * RS_THREADS + threadIdx.x;
    if (e < N) atomicAdd(&h[(unsigned)(keys[e] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];          // digit-major: the scan order
}

// per digit (one CTA per digit row of the digit-major histogram): exclusive scan over the tiles in place, row total to
// tot[digit]. Coalesced; the scatter kernel adds the digits-below offset (a 256-wide scan of tot) itself.
__global__ void __launch_bounds__(1024) rs_rowscan_kernel(unsigned* __restrict__ hist, int nblocks, unsigned* __restrict__ tot) {
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry_s;
  unsigned* row = hist + (size_t)blockIdx.x * nblocks;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0u;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + tid;
    const unsigned v = i < nblocks ? row[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = wsum[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    if (i < nblocks) row[i] = carry + wsum[warp] + incl - v;
    __syncthreads();
    if (tid == 1023) carry_s = carry + wsum[31] + incl;
    __syncthreads();
  }
  if (tid == 0) tot[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long* __restrict__ keys_in, const int32_t* __restrict__ vals_in, long long N, int shift,
                  int nblocks, const unsigned* __restrict__ bases, const unsigned* __restrict__ tot,
                  unsigned long long* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ unsigned cnt[RS_WARPS][RS_BINS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0u;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE + (long long)warp * (32 * RS_ROUNDS);
  unsigned long long key[RS_ROUNDS];
  int32_t val[RS_ROUNDS];
  unsigned off[RS_ROUNDS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const long long e = base + r * 32 + lane;
    const bool valid = e < N;
    key[r] = valid ? keys_in[e] : 0ull;
    val[r] = valid ? vals_in[e] : 0;
    const unsigned d = (unsigned)(key[r] >> shift) & 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (256u + lane));
    const unsigned before = cnt[warp][d];
    __syncwarp();
    if (valid && (peers & lt) == 0u) cnt[warp][d] = before + __popc(peers);     // the lowest lane of each digit group
    __syncwarp();
    off[r] = valid ? before + __popc(peers & lt) : 0xffffffffu;
  }
  __syncthreads();
  {
    // per digit: keys with a smaller digit (256-wide exclusive scan of the digit totals), then this digit's keys in the
    // tiles before this one, then the warps of the tile in order
    __shared__ unsigned dsum[RS_WARPS];
    const unsigned tv = tot[tid];
    unsigned incl = tv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) dsum[warp] = incl;
    __syncthreads();
    unsigned below = incl - tv;
    for (int w = 0; w < warp; ++w) below += dsum[w];
    unsigned run = below + bases[(size_t)tid * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const unsigned t = cnt[w][tid];
      cnt[w][tid] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    if (off[r] != 0xffffffffu) {
      const unsigned d = (unsigned)(key[r] >> shift) & 255u;
      const unsigned p = cnt[warp][d] + off[r];
      keys_out[p] = key[r];
      vals_out[p] = val[r];
    }
  }
}

// hist tiles use the (round, thread) layout, the scatter the (warp, round, lane) layout: both cover the same tile
// [blockIdx.x * RS_TILE, +RS_TILE), so the per-tile digit counts agree.

__global__ void __launch_bounds__(256)
rank_gather_kernel(const int32_t* __restrict__ order, long long N, const double* __restrict__ z, const int32_t* __restrict__ pos,
                   int k, const double2* __restrict__ pool, int pool_size, double* __restrict__ z_sorted,
                   int32_t* __restrict__ pos_sorted, double2* __restrict__ dist, int32_t* __restrict__ order_out) {
  const long long total = N * k;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / k;
    const int a = (int)(e - i * k);
    const int src = order[i];
    const int g = pos ? pos[(long long)src * k + a] : 0;
    if (pos_sorted) pos_sorted[e] = g;
    if (dist) dist[e] = (g >= 0 && g < pool_size) ? __ldg(&pool[g]) : make_double2(nan(""), nan(""));
    if (a == 0) {
      if (z_sorted) z_sorted[i] = z[src];
      if (order_out) order_out[i] = src;
    }
  }
}

__global__ void __launch_bounds__(256)
count_less_kernel(const double* __restrict__ scores, long long N, double z, unsigned long long* __restrict__ out) {
  unsigned c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    c += scores[i] < z ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// the r-th (0-based) rank of [lo, hi) whose bit in `acquired` is clear; marks it; gathers its set. One CTA.
__global__ void __launch_bounds__(256)
window_pick_kernel(unsigned* __restrict__ acquired, int lo, int hi, int r, const int32_t* __restrict__ pos_sorted, int k,
                   const double2* __restrict__ pool, int pool_size, int32_t* __restrict__ loc_out, double2* __restrict__ set_out) {
  __shared__ int wsum[8];
  __shared__ int s_base, s_loc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_base = 0; s_loc = -1; }
  __syncthreads();
  const int w0 = lo >> 5, w1 = (hi + 31) >> 5;
  for (int wb = w0; wb < w1 && s_loc < 0; wb += 256) {
    const int w = wb + tid;
    unsigned freeb = 0u;
    if (w < w1) {
      freeb = ~acquired[w];
      if (w == w0) freeb &= ~0u << (lo & 31);
      if (w == (hi >> 5)) freeb &= (hi & 31) ? ((1u << (hi & 31)) - 1u) : 0u;
      if (w > (hi >> 5)) freeb = 0u;
    }
    const int c = __popc(freeb);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = s_base + incl - c;
    int tot = 0;
    for (int q = 0; q < 8; ++q) {
      if (q < warp) before += wsum[q];
      tot += wsum[q];
    }
    if (c > 0 && r >= before && r < before + c) {
      const int bit = __fns(freeb, 0, r - before + 1);
      acquired[w] |= 1u << bit;
      s_loc = (w << 5) + bit;
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  const int loc = s_loc;
  if (tid == 0) *loc_out = loc;
  if (loc >= 0 && set_out && tid < k) {
    const int g = pos_sorted[(long long)loc * k + tid];
    set_out[tid] = (g >= 0 && g < pool_size) ? __ldg(&pool[g]) : make_double2(nan(""), nan(""));
  }
}

// ---- multi-gamma scorer: squared distances once, exp + LDL^T per gamma (generic pool; k <= 12) -----------------------
template <int K>
__device__ __forceinline__ double ldl_logabsdet_unitdiag(const double (&D)[K * (K - 1) / 2 > 0 ? K * (K - 1) / 2 : 1], double neg_gamma) {
  // same elimination order and operations as ldl_logabsdet (csrc/dpp.cu) on A = exp(neg_gamma * D), unit diagonal
  double A[K * (K + 1) / 2];
#pragma unroll
  for (int r = 0; r < K; ++r) {
#pragma unroll
    for (int c = 0; c < r; ++c) A[r * (r + 1) / 2 + c] = exp(neg_gamma * D[r * (r - 1) / 2 + c]);
    A[r * (r + 1) / 2 + r] = 1.0;
  }
  double piv[K];                                      // log|det| exactly as ldl_logabsdet (csrc/dpp.cu) forms it
  bool singular = false;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double d = A[j * (j + 1) / 2 + j];
    singular = singular || d == 0.0;
    piv[j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int c = j + 1; c < K; ++c) {
      const double lc = A[c * (c + 1) / 2 + j] * inv;
#pragma unroll
      for (int r = c; r < K; ++r) A[r * (r + 1) / 2 + c] = fma(-A[r * (r + 1) / 2 + j], lc, A[r * (r + 1) / 2 + c]);
    }
  }
  if (singular) return -INFINITY;
  double p = 1.0;
#pragma unroll
  for (int j = 0; j < K; ++j) p *= piv[j];
  p = fabs(p);
  if (p > 1e-250 && p < 1e250) return log(p);
  double acc = 0.0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) acc += log(fabs(piv[j]));
  return acc;
}

template <int K>
__global__ void __launch_bounds__(128)
dpp_multigamma_kernel(const double2* __restrict__ pool, const int32_t* __restrict__ idx, long long N,
                      const double* __restrict__ gammas, int n_gamma, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  double px[K], py[K];
  const int32_t* row = idx + i * K;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const double2 p = __ldg(&pool[row[a]]);
    px[a] = p.x;
    py[a] = p.y;
  }
  double D[K * (K - 1) / 2 > 0 ? K * (K - 1) / 2 : 1];
#pragma unroll
  for (int r = 1; r < K; ++r) {
#pragma unroll
    for (int c = 0; c < r; ++c) {
      const double dx = px[r] - px[c], dy = py[r] - py[c];
      const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // as dpp_entry (csrc/dpp.cu)
      D[r * (r - 1) / 2 + c] = d * d;
    }
  }
  for (int g = 0; g < n_gamma; ++g) out[(long long)g * N + i] = ldl_logabsdet_unitdiag<K>(D, -__ldg(&gammas[g]));
}

template <int K>
static void launch_multigamma(const double* pool, const int32_t* idx, long long N, const double* gammas, int n_gamma,
                              double* out, cudaStream_t st) {
  dpp_multigamma_kernel<K><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(reinterpret_cast<const double2*>(pool), idx, N, gammas,
                                                                       n_gamma, out);
}

static inline int rs_tiles(long long N) { return (int)((N + RS_TILE - 1) / RS_TILE); }

}  // namespace b200bo

extern "C" {

size_t b200bo_dpp_rank_workspace_bytes(long long N) {
  using namespace b200bo;
  if (N <= 0) return 0;
  const size_t n = (size_t)N;
  const size_t keys = ((n * 8 + 255) / 256) * 256, vals = ((n * 4 + 255) / 256) * 256;
  return 2 * keys + 2 * vals + ((size_t)rs_tiles(N) + 1) * RS_BINS * sizeof(unsigned) + 256;
}

int b200bo_dpp_rank(const double* z, long long N, const int32_t* pos, int k, const double* pool, int pool_size,
                    double* z_sorted, int32_t* order, int32_t* pos_sorted, double* dist, void* ws, size_t ws_bytes,
                    void* stream) {
  using namespace b200bo;
  if (N < 0 || N > 2147483647LL) return arg_error("dpp_rank: N=%lld", N);
  if (N == 0) return 0;
  if (!z || !ws) return arg_error("dpp_rank: null pointer");
  if ((pos_sorted || dist) && (!pos || k < 1)) return arg_error("dpp_rank: pos / k needed for pos_sorted or dist");
  if (dist && (!pool || pool_size < 1)) return arg_error("dpp_rank: pool needed for dist");
  if (ws_bytes < b200bo_dpp_rank_workspace_bytes(N)) return arg_error("dpp_rank: workspace too small");
  cudaStream_t st = as_stream(stream);
  const size_t n = (size_t)N;
  const size_t keys_b = ((n * 8 + 255) / 256) * 256, vals_b = ((n * 4 + 255) / 256) * 256;
  char* w = reinterpret_cast<char*>(ws);
  unsigned long long* keys[2] = {reinterpret_cast<unsigned long long*>(w), reinterpret_cast<unsigned long long*>(w + keys_b)};
  int32_t* vals[2] = {reinterpret_cast<int32_t*>(w + 2 * keys_b), reinterpret_cast<int32_t*>(w + 2 * keys_b + vals_b)};
  unsigned* hist = reinterpret_cast<unsigned*>(w + 2 * keys_b + 2 * vals_b);
  const int tiles = rs_tiles(N);
  unsigned* tot = hist + (size_t)tiles * RS_BINS;
  const int blocks = (int)((N + 255) / 256 < 1184 ? (N + 255) / 256 : 1184);
  rs_keys_kernel<<<blocks, 256, 0, st>>>(z, N, keys[0], vals[0]);
  int cur = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = pass * 8;
    rs_hist_kernel<<<tiles, RS_THREADS, 0, st>>>(keys[cur], N, shift, tiles, hist);
    rs_rowscan_kernel<<<RS_BINS, 1024, 0, st>>>(hist, tiles, tot);
    rs_scatter_kernel<<<tiles, RS_THREADS, 0, st>>>(keys[cur], vals[cur], N, shift, tiles, hist, tot, keys[cur ^ 1], vals[cur ^ 1]);
    cur ^= 1;
  }
  const int kk = (pos_sorted || dist) ? k : 1;
  const long long total = N * kk;
  const int gblocks = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  rank_gather_kernel<<<gblocks, 256, 0, st>>>(vals[cur], N, z, pos, kk, reinterpret_cast<const double2*>(pool), pool_size,
                                              z_sorted, pos_sorted, reinterpret_cast<double2*>(dist), order);
  B200BO_CHECK_LAUNCH("dpp_rank");
  return 0;
}

int b200bo_dpp_count_less(const double* scores, long long N, double z, uint64_t* count, void* stream) {
  using namespace b200bo;
  if (N < 0 || !count || (N > 0 && !scores)) return arg_error("dpp_count_less: bad arguments");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint64_t), st);
  if (e != cudaSuccess) return cuda_status(e, "dpp_count_less");
  if (N == 0) return 0;
  const int blocks = (int)((N + 255) / 256 < 1184 ? (N + 255) / 256 : 1184);
  count_less_kernel<<<blocks, 256, 0, st>>>(scores, N, z, reinterpret_cast<unsigned long long*>(count));
  B200BO_CHECK_LAUNCH("dpp_count_less");
  return 0;
}

int b200bo_dpp_window_pick(uint32_t* acquired_mask, long long N, int lo, int hi, int r, const int32_t* pos_sorted, int k,
                           const double* pool, int pool_size, int32_t* loc, double* set_out, void* stream) {
  using namespace b200bo;
  if (!acquired_mask || !loc) return arg_error("dpp_window_pick: null pointer");
  if (lo < 0 || hi < lo || hi > N || r < 0) return arg_error("dpp_window_pick: window [%d, %d) of %lld, r=%d", lo, hi, N, r);
  if (set_out && (!pos_sorted || !pool || k < 1 || k > 256)) return arg_error("dpp_window_pick: pos_sorted / pool / k");
  window_pick_kernel<<<1, 256, 0, as_stream(stream)>>>(acquired_mask, lo, hi, r, pos_sorted, k,
                                                       reinterpret_cast<const double2*>(pool), pool_size, loc,
                                                       reinterpret_cast<double2*>(set_out));
  B200BO_CHECK_LAUNCH("dpp_window_pick");
  return 0;
}

int b200bo_dpp_logdet_score_multigamma(const double* pool, int pool_size, const int32_t* idx, long long N, int k,
                                       const double* gammas, int n_gamma, double* logdet, void* stream) {
  using namespace b200bo;
  if (pool_size <= 0 || k < 1 || k > 12 || N < 0 || n_gamma < 1)
    return arg_error("dpp_logdet_score_multigamma: pool=%d k=%d (max 12) N=%lld n_gamma=%d", pool_size, k, N, n_gamma);
  if (N == 0) return 0;
  if (!pool || !idx || !gammas || !logdet) return arg_error("dpp_logdet_score_multigamma: null pointer");
  cudaStream_t st = as_stream(stream);
  switch (k) {
    case 1: launch_multigamma<1>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 2: launch_multigamma<2>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 3: launch_multigamma<3>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 4: launch_multigamma<4>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 5: launch_multigamma<5>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 6: launch_multigamma<6>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 7: launch_multigamma<7>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 8: launch_multigamma<8>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 9: launch_multigamma<9>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 10: launch_multigamma<10>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    case 11: launch_multigamma<11>(pool, idx, N, gammas, n_gamma, logdet, st); break;
    default: launch_multigamma<12>(pool, idx, N, gammas, n_gamma, logdet, st); break;
  }
  B200BO_CHECK_LAUNCH("dpp_logdet_score_multigamma");
  return 0;
}

}  // extern "C"
