// Element ranks of the ranked-DPP sampler, drawn on the device bit for bit as CPython draws them.
//
// Replaces the loop of data_gen.iid_comb_locs (Src/data_gen.py:865-871, the C(n,k) > int64 branch):
//     for int_seed in seed_list:  random.seed(int_seed);  elements.append(random.randint(0, total))
// i.e. one freshly seeded Mersenne Twister per element — 97 % of generate()'s wall time once un-ranking and scoring
// run on the GPU (one MT19937 init_by_array = 1 247 dependent 32-bit steps per element in CPython's C code).
// Each thread restates CPython 3.9's chain for one element:
//   random.seed(np.int64 s)      -> key = (size_t) hash(s) as little-endian 32-bit words (hash(-1) == -2), _randommodule.c
//   init_by_array(key)           -> mt[624] from the constant init_genrand(19650218) state
//   randint(0, total)            -> randrange(0, total + 1) -> _randbelow_with_getrandbits(n):
//                                   k = n.bit_length(); r = getrandbits(k) until r < n   (Lib/random.py)
//   getrandbits(k)               -> ceil(k/32) tempered words, least significant first, the last one >> (32 - k % 32)
// Only the first outputs are needed, so the 624-word twist is evaluated lazily for the words actually consumed
// (word t < 227 depends on the un-twisted mt[t], mt[t+1], mt[t+397] only). Integer work: bit-exact.
//
// Roofline: issue / local-memory bound (2.5 KB of private state per element), ~10 k integer instructions per element.
#include "common.cuh"

namespace b200bo {

typedef unsigned __int128 u128;
constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_MAX_WORDS = 224;          // lazy twist limit (t + 397 < 624): 56 draws of 4 words

__constant__ uint32_t c_mt_genrand[MT_N];  // init_genrand(19650218u), uploaded stream-ordered by every call

struct MtGenrandTable {
  uint32_t mt[MT_N];
  MtGenrandTable() {
    mt[0] = 19650218u;
    for (int j = 1; j < MT_N; ++j) mt[j] = 1812433253u * (mt[j - 1] ^ (mt[j - 1] >> 30)) + (uint32_t)j;
  }
};

__device__ __forceinline__ uint32_t mt_word(const uint32_t* mt, int t) {
  uint32_t y = (mt[t] & 0x80000000u) | (mt[t + 1] & 0x7fffffffu);
  y = mt[t + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

__global__ void __launch_bounds__(128)
mt_seeded_randbelow_kernel(const long long* __restrict__ seeds, long long N, unsigned long long n_lo,
                           unsigned long long n_hi, int nbits, unsigned long long* __restrict__ out_lo,
                           unsigned long long* __restrict__ out_hi, int32_t* __restrict__ fail) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t mt[MT_N];
  for (int j = 0; j < MT_N; ++j) mt[j] = c_mt_genrand[j];
  long long s = seeds[i];
  if (s == -1) s = -2;                                   // hash(-1) == -2
  const unsigned long long v = (unsigned long long)s;    // (size_t) hash
  uint32_t key[2] = {(uint32_t)v, (uint32_t)(v >> 32)};
  const int klen = key[1] ? 2 : 1;
  {
    int a = 1, b = 0;
    for (int k = MT_N; k; --k) {                         // max(N, key_length) = 624
      mt[a] = (mt[a] ^ ((mt[a - 1] ^ (mt[a - 1] >> 30)) * 1664525u)) + key[b] + (uint32_t)b;
      ++a; ++b;
      if (a >= MT_N) { mt[0] = mt[MT_N - 1]; a = 1; }
      if (b >= klen) b = 0;
    }
    for (int k = MT_N - 1; k; --k) {
      mt[a] = (mt[a] ^ ((mt[a - 1] ^ (mt[a - 1] >> 30)) * 1566083941u)) - (uint32_t)a;
      ++a;
      if (a >= MT_N) { mt[0] = mt[MT_N - 1]; a = 1; }
    }
    mt[0] = 0x80000000u;
  }
  const u128 n = ((u128)n_hi << 64) | n_lo;
  const int words = (nbits - 1) / 32 + 1;
  int t = 0;
  bool done = false;
  u128 r = 0;
  while (!done && t + words <= MT_MAX_WORDS) {
    r = 0;
    int k = nbits;
    for (int w = 0; w < words; ++w, k -= 32) {
      uint32_t y = mt_word(mt, t++);
      if (k < 32) y >>= (32 - k);
      r |= (u128)y << (32 * w);
    }
    done = r < n;
  }
  out_lo[i] = (unsigned long long)r;
  out_hi[i] = (unsigned long long)(r >> 64);
  if (!done) atomicAdd(fail, 1);
}

}  // namespace b200bo

extern "C" int b200bo_mt_seeded_randbelow(const int64_t* seeds, long long N, uint64_t n_lo, uint64_t n_hi,
                                          uint64_t* out_lo, uint64_t* out_hi, int32_t* fail, void* stream) {
  using namespace b200bo;
  if (N < 0) return arg_error("mt_seeded_randbelow: N=%lld", N);
  if (!n_lo && !n_hi) return arg_error("mt_seeded_randbelow: n must be positive");
  if (N == 0) return 0;
  if (!seeds || !out_lo || !out_hi || !fail) return arg_error("mt_seeded_randbelow: null pointer");
  static const MtGenrandTable table;                   // immutable after construction (thread-safe static init)
  cudaError_t e = cudaMemcpyToSymbolAsync(c_mt_genrand, table.mt, sizeof(table.mt), 0, cudaMemcpyHostToDevice,
                                          as_stream(stream));
  if (e != cudaSuccess) return cuda_status(e, "mt_seeded_randbelow: constant upload");
  int nbits = 0;
  for (u128 t = ((u128)n_hi << 64) | n_lo; t; t >>= 1) ++nbits;
  mt_seeded_randbelow_kernel<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(seeds), N, n_lo, n_hi, nbits, reinterpret_cast<unsigned long long*>(out_lo),
      reinterpret_cast<unsigned long long*>(out_hi), fail);
  B200BO_CHECK_LAUNCH("mt_seeded_randbelow");
  return 0;
}
