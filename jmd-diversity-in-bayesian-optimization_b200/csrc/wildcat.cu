// K-E: Wildcat Wells surface generation (one CTA per surface) and raw simplex noise.
//
// Replaces objectives.wildcatwells (Src/objectives.py:107-218; N = 1 peak, d = 2) and the
// pure-Python SimplexNoise.simplexNoise loop inside it (Src/SimpleNoise.py:85-172,205-220).
// All arithmetic is fp64 in the reference's operation order; the 4-entry gradient table
// (shuffled on the host by random.Random(hseed), SimpleNoise.py:81-83) is passed in.
//
// Roofline: FP64-pipe bound (~3 MFLOP per surface, 40.8 KB written). The two fp64 fields
// (Gaussian pdf and noise amplitude A) stay in shared memory (2 x 10201 x 8 B = 163 KB)
// between the three global max-reductions, so HBM sees only the final fp32 surface.
#include "common.cuh"

namespace b200bo {

constexpr int WW_SIDE = 101;
constexpr int WW_PTS = WW_SIDE * WW_SIDE;
constexpr int WW_THREADS = 512;

// SimpleNoise.py:205-220 — no masking; < 2^58 for the inputs that occur, uint64 is exact.
__host__ __device__ __forceinline__ unsigned long long rand_hash(unsigned long long v) {
  unsigned long long h = v & 255ull;
  h += h << 10; h ^= h >> 6;
  h += (v >> 8) & 255ull;
  h += h << 10; h ^= h >> 6;
  h += (v >> 16) & 255ull;
  h += h << 10; h ^= h >> 6;
  h += h << 3; h ^= h >> 11;
  return h + (h << 15);
}

// t**4 as CPython computes it: one rounding of the exact fourth power (C pow), obtained
// here from an error-free square: (t2 + e2)^2 = t2^2 + 2 t2 e2 + O(e2^2).
__device__ __forceinline__ double pow4_rn(double t) {
  double t2 = t * t;
  double e2 = fma(t, t, -t2);
  double t4 = t2 * t2;
  double e4 = fma(t2, t2, -t4);
  return t4 + fma(2.0 * t2, e2, e4);
}

struct SimplexConsts {
  double f, g, c2f, scaler;   // skew, unskew, cornerToFaceSquared, valueScaler (SimpleNoise.py:24,31,41,65-70)
};

__device__ __forceinline__ SimplexConsts simplex_consts() {
  SimplexConsts c;
  c.f = 0.3660254037844386;        // ((2+1)**.5 - 1)/2
  c.g = 0.2113248654051871;        // f/(1+2f)
  c.c2f = 0.5000000000000001;      // a**2
  c.scaler = 113.0;
  return c;
}

// SimpleNoise.py:85-172 for d = 2; vx/vy: gradient table components.
// Python evaluates every * and + with its own rounding; __dmul_rn/__dadd_rn keep nvcc from
// contracting them into FMAs, which is what makes the result bit-identical to the reference.
__device__ __forceinline__ double simplex_noise_2d(double x, double y, unsigned long long hseed,
                                                   const double vx[4], const double vy[4]) {
  const SimplexConsts k = simplex_consts();
  double s = __dmul_rn(__dadd_rn(__dadd_rn(0.0, x), y), k.f);     // sum(loc)*f, sum starts from int 0
  double fx = floor(__dadd_rn(x, s)), fy = floor(__dadd_rn(y, s));
  long long ix = (long long)fx, iy = (long long)fy;
  double t0 = __dmul_rn((double)(ix + iy), k.g);
  double cx = __dadd_rn(__dadd_rn(x, -fx), t0);                  // loc[i]-intSkewLoc[i]+t, left to right
  double cy = __dadd_rn(__dadd_rn(y, -fy), t0);
  bool x_first = cx > cy;                                        // sortWith + reverse: ties -> axis 1 first
  double n = 0.0;
  double skew = 0.0;
  long long ox = 0, oy = 0;
#pragma unroll
  for (int step = 0; step < 3; ++step) {
    if (step == 1) { if (x_first) ox += 1; else oy += 1; }
    if (step == 2) { if (x_first) oy += 1; else ox += 1; }
    double ux = __dadd_rn(__dadd_rn(cx, -(double)ox), skew);
    double uy = __dadd_rn(__dadd_rn(cy, -(double)oy), skew);
    double t = k.c2f;
    t = __dadd_rn(t, -__dmul_rn(ux, ux));
    t = __dadd_rn(t, -__dmul_rn(uy, uy));
    if (t > 0.0) {
      unsigned long long index = hseed + (unsigned long long)(ix + ox) + ((unsigned long long)(iy + oy) << 5);
      int vi = (int)(rand_hash(index) & 3ull);                   // % vecCount (= 4)
      double gr = __dadd_rn(0.0, __dmul_rn(vx[vi], ux));
      gr = __dadd_rn(gr, __dmul_rn(vy[vi], uy));
      n = __dadd_rn(n, __dmul_rn(gr, pow4_rn(t)));
    }
    skew = __dadd_rn(skew, k.g);
  }
  return __dmul_rn(n, k.scaler);
}

__device__ __forceinline__ double block_max(double v, double* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, sh[w]);
  return r;
}

__global__ void __launch_bounds__(WW_THREADS, 1)
wildcat_surface_kernel(const unsigned long long* __restrict__ hseed, const int8_t* __restrict__ vecs,
                       const double* __restrict__ peak_xy, const double* __restrict__ smooth,
                       const double* __restrict__ rug_amp, float* __restrict__ surf, double log_norm) {
  extern __shared__ double sm[];
  double* gaus = sm;
  double* amp = sm + WW_PTS;
  __shared__ double red[WW_THREADS / 32];
  const int sidx = blockIdx.x;
  const unsigned long long hs = hseed[sidx];
  double vx[4], vy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    vx[i] = (double)vecs[sidx * 8 + i * 2];
    vy[i] = (double)vecs[sidx * 8 + i * 2 + 1];
  }
  const double px = peak_xy[sidx * 2], py = peak_xy[sidx * 2 + 1];
  const double smoothness = smooth[sidx];
  const double om = 1.0 - smoothness;
  const double fs = __dmul_rn(100.0, __dadd_rn(1.0, -__dmul_rn(om, om)));        // objectives.py:191  (1-smoothness)**2 == om*om exactly rounded once
  const double fs4 = fs / 4.0, fs8 = fs / 8.0;
  const double nstep = 100.0 / 101.0;              // np.arange(0,100,100/101), objectives.py:101
  const double var = 100.0;                        // (0.2*50)**2, objectives.py:185

  double gmax = -INFINITY, amax = -INFINITY;
  for (int p = threadIdx.x; p < WW_PTS; p += WW_THREADS) {
    int iy = p / WW_SIDE, ix = p - iy * WW_SIDE;
    double dx = (double)ix - px, dy = (double)iy - py;
    double maha = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) / var;
    double g = exp(__dmul_rn(-0.5, __dadd_rn(log_norm, maha)));
    double nx = (double)ix * nstep, ny = (double)iy * nstep;
    double v = __dadd_rn(__dadd_rn(simplex_noise_2d(nx / fs, ny / fs, hs, vx, vy),
                                   simplex_noise_2d(nx / fs4, ny / fs4, hs, vx, vy) / 4.0),
                         simplex_noise_2d(nx / fs8, ny / fs8, hs, vx, vy) / 8.0);
    double a = __dadd_rn(v, 1.0);                  // (values+1)*N, N = 1
    gaus[p] = g;
    amp[p] = a;
    gmax = fmax(gmax, g);
    amax = fmax(amax, a);
  }
  gmax = block_max(gmax, red);
  amax = block_max(amax, red);
  // objectives.py:213-214
  const double power = -log(amax / gmax) / log(10.0);
  const double ra = 1.0 - rug_amp[sidx];
  const double scale = pow(10.0, power - pow(ra, 5.0));
  double smax = -INFINITY;
  for (int p = threadIdx.x; p < WW_PTS; p += WW_THREADS) {
    double sv = __dadd_rn(gaus[p], __dmul_rn(amp[p], scale));
    gaus[p] = sv;
    smax = fmax(smax, sv);
  }
  smax = block_max(smax, red);
  float* out = surf + (size_t)sidx * WW_PTS;
  for (int p = threadIdx.x; p < WW_PTS; p += WW_THREADS) out[p] = (float)((gaus[p] / smax) * 100.0);
}

__global__ void __launch_bounds__(256)
simplex_noise_kernel(unsigned long long hseed, double v0x, double v0y, double v1x, double v1y, double v2x, double v2y,
                     double v3x, double v3y, const double2* __restrict__ xy, double* __restrict__ out, long long M) {
  double vx[4] = {v0x, v1x, v2x, v3x};
  double vy[4] = {v0y, v1y, v2y, v3y};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    double2 p = xy[i];
    out[i] = simplex_noise_2d(p.x, p.y, hseed, vx, vy);
  }
}

// Piecewise-linear evaluation of a surface between lattice nodes = the LinearNDInterpolator that the
// reference's generate_cont() returns (objectives.py:349-350). At nodes the value is the fp32 node value
// exactly. Inside a unit cell the interpolant is linear on each of the two triangles of scipy's Delaunay
// (Qhull) triangulation of the 101 x 101 nodes; which diagonal Qhull cuts a cell along depends on the node
// coordinates only, so it is an input: diag_mask[iy*100 + jx] = 1 -> (jx,iy)-(jx+1,iy+1) ("main"),
// 0 -> (jx+1,iy)-(jx,iy+1) ("anti"); NULL = main everywhere. The value is the barycentric sum
// c0*v0 + c1*v1 + c2*v2 as scipy forms it (interpnd.pyx: _barycentric_coordinates + the weighted sum); with the
// weights of a unit right triangle being exact differences of the fractional coordinates, half-node queries
// (the 200 x 200 lattice of configs[3]) come out bit-identical, general queries within a few ulp.
// Outside [0,100]^2: NaN, the interpolator's fill value.
__global__ void __launch_bounds__(256)
wildcat_eval_kernel(const float* __restrict__ surf, const int32_t* __restrict__ surf_id,
                    const uint8_t* __restrict__ diag_mask, const double2* __restrict__ xy,
                    double* __restrict__ out, long long M) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const double2 p = xy[i];
    const float* s = surf + (size_t)(surf_id ? surf_id[i] : 0) * WW_PTS;
    double r;
    if (!(p.x >= 0.0 && p.x <= 100.0 && p.y >= 0.0 && p.y <= 100.0)) {
      r = nan("");
    } else {
      int jx = (int)floor(p.x), iy = (int)floor(p.y);
      if (jx > 99) jx = 99;
      if (iy > 99) iy = 99;
      const double u = p.x - jx, v = p.y - iy;
      const double v00 = s[iy * WW_SIDE + jx], v10 = s[iy * WW_SIDE + jx + 1];
      const double v01 = s[(iy + 1) * WW_SIDE + jx], v11 = s[(iy + 1) * WW_SIDE + jx + 1];
      const bool main_diag = diag_mask ? (diag_mask[iy * (WW_SIDE - 1) + jx] != 0) : true;
      double c0, c1, c2, a0, a1, a2;
      if (main_diag) {
        if (u >= v) { c0 = 1.0 - u; c1 = u - v; c2 = v; a0 = v00; a1 = v10; a2 = v11; }
        else        { c0 = 1.0 - v; c1 = v - u; c2 = u; a0 = v00; a1 = v01; a2 = v11; }
      } else {
        if (u + v <= 1.0) { c0 = (1.0 - u) - v; c1 = u; c2 = v; a0 = v00; a1 = v10; a2 = v01; }
        else              { c0 = (u - 1.0) + v; c1 = 1.0 - u; c2 = 1.0 - v; a0 = v11; a1 = v01; a2 = v10; }
      }
      r = __dadd_rn(__dadd_rn(__dmul_rn(c0, a0), __dmul_rn(c1, a1)), __dmul_rn(c2, a2));
    }
    out[i] = r;
  }
}

}  // namespace b200bo

extern "C" {

int b200bo_wildcat_eval(const float* surf, const int32_t* surf_id, const uint8_t* diag_mask, const double* xy, double* out,
                        long long M, void* stream) {
  using namespace b200bo;
  if (M < 0) return arg_error("wildcat_eval: M=%lld", M);
  if (M == 0) return 0;
  if (!surf || !xy || !out) return arg_error("wildcat_eval: null pointer");
  long long blocks = (M + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  wildcat_eval_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(surf, surf_id, diag_mask,
                                                                      reinterpret_cast<const double2*>(xy), out, M);
  B200BO_CHECK_LAUNCH("wildcat_eval");
  return 0;
}

int b200bo_wildcat_surface(const uint64_t* hseed, const int8_t* vecs, const double* peak_xy, const double* smooth,
                           const double* rug_amp, float* surf, int S, void* stream) {
  using namespace b200bo;
  if (S < 0) return arg_error("wildcat_surface: S=%d", S);
  if (S == 0) return 0;
  if (!hseed || !vecs || !peak_xy || !smooth || !rug_amp || !surf) return arg_error("wildcat_surface: null pointer");
  size_t smem = 2 * (size_t)WW_PTS * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(wildcat_surface_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_status(e, "wildcat_surface: smem attribute");
  // scipy multivariate_normal.logpdf: -0.5*(rank*log(2pi) + log|cov| + maha), cov = 100*I
  const double log_norm = 2.0 * log(2.0 * M_PI) + 2.0 * log(100.0);
  wildcat_surface_kernel<<<S, WW_THREADS, smem, as_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(hseed), vecs, peak_xy, smooth, rug_amp, surf, log_norm);
  B200BO_CHECK_LAUNCH("wildcat_surface");
  return 0;
}

int b200bo_simplex_noise(uint64_t hseed, const int8_t* vecs_h, const double* xy, double* out, long long M, void* stream) {
  using namespace b200bo;
  if (M < 0) return arg_error("simplex_noise: M=%lld", M);
  if (M == 0) return 0;
  if (!vecs_h || !xy || !out) return arg_error("simplex_noise: null pointer");
  long long blocks = (M + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  simplex_noise_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      hseed, vecs_h[0], vecs_h[1], vecs_h[2], vecs_h[3], vecs_h[4], vecs_h[5], vecs_h[6], vecs_h[7],
      reinterpret_cast<const double2*>(xy), out, M);
  B200BO_CHECK_LAUNCH("simplex_noise");
  return 0;
}

}  // extern "C"
