// Micro-benchmarks that give the roofline denominators for the FP64-bound
// kernels of this repo: DFMA and DMMA
// (mma.sync.m8n8k4.f64) issue rate, their overlap, fp64 transcendental cost and
// shared-memory operand delivery.  Prints one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096;

__global__ void k_dfma(double* out, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, float a, float b) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same warp: 8 DMMA + NF DFMA per iteration
template <int NF>
__global__ void k_mix(double* out, double a, double b) {
  double c0[8], c1[8], f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int q = 0; q < NF / 8; ++q) f[i * (NF / 8) + q] = fma(f[i * (NF / 8) + q], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// different warps: even warps DMMA, odd warps DFMA
__global__ void k_split(double* out, double a, double b) {
  int w = threadIdx.x >> 5;
  double s = 0;
  if (w & 1) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
  } else {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c0[i], c1[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental throughput: OP 0 exp, 1 sqrt, 2 erfc, 3 div, 4 rsqrt, 5 log
template <int OP>
__global__ void k_trans(double* out, double a) {
  double acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = 0.5 + threadIdx.x * 1e-4 + i * 0.01;
  for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double x = acc[i];
      if (OP == 0) x = exp(-x) + a;
      if (OP == 1) x = sqrt(x) + a;
      if (OP == 2) x = erfc(x) + a;
      if (OP == 3) x = a / x + 0.25;
      if (OP == 4) x = rsqrt(x) * a + 0.1;
      if (OP == 5) x = log(x + 1.0) + a;
      acc[i] = x;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

// shared-memory operand delivery. MODE 0: LDS.64 broadcast, 1: LDS.64 lane-distinct,
// 2: LDS.128 broadcast, 3: LDS.128 lane-distinct, 4: LDS.32 broadcast
template <int MODE>
__global__ void k_lds(double* out, int stride) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int lane = threadIdx.x & 31;
  int base = (MODE == 1) ? lane : (MODE == 3 ? lane * 2 : 0);
  for (int it = 0; it < ITER / 4; ++it) {
    int o = (it * stride) & 1023;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int idx = base + o + u * 64;
      if (MODE == 0 || MODE == 1) { s0 += sm[idx]; }
      else if (MODE == 2 || MODE == 3) {
        double2 v = *reinterpret_cast<const double2*>(&sm[idx & ~1]);
        s1 += v.x; s2 += v.y;
      } else {
        s3 += reinterpret_cast<const float*>(sm)[idx];
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  int blocks = sms * 8, threads = 256;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  float* outf = reinterpret_cast<float*>(out);
  double nthreads = double(blocks) * threads;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);

  float ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"dfma_tflops\": %.3f", nthreads * ITER * 8 * 2 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>(outf, 1.0000001f, 1e-9f); });
  printf(", \"ffma_tflops\": %.3f", nthreads * ITER * 8 * 2 / (ms * 1e-3) / 1e12);

  double warps = nthreads / 32;
  ms = time_ms([&] { k_dmma<2><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"dmma_tflops_acc2\": %.3f", warps * ITER * 2 * 512 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dmma<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"dmma_tflops_acc4\": %.3f", warps * ITER * 4 * 512 / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  float ms_dmma8 = ms;
  printf(", \"dmma_tflops_acc8\": %.3f", warps * ITER * 8 * 512 / (ms * 1e-3) / 1e12);

  // mix: time relative to DMMA-only tells whether DFMA rides for free
  ms = time_ms([&] { k_mix<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"mix_8dmma_8dfma_ms\": %.4f, \"dmma8_only_ms\": %.4f", ms, ms_dmma8);
  printf(", \"mix8_total_tflops\": %.3f", (warps * ITER * 8 * 512 + nthreads * ITER * 8 * 2) / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_mix<16><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"mix_8dmma_16dfma_ms\": %.4f", ms);
  printf(", \"mix16_total_tflops\": %.3f", (warps * ITER * 8 * 512 + nthreads * ITER * 16 * 2) / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_mix<32><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"mix_8dmma_32dfma_ms\": %.4f", ms);
  printf(", \"mix32_total_tflops\": %.3f", (warps * ITER * 8 * 512 + nthreads * ITER * 32 * 2) / (ms * 1e-3) / 1e12);
  ms = time_ms([&] { k_split<<<blocks, threads>>>(out, 1.0000001, 1e-9); });
  printf(", \"split_warps_ms\": %.4f, \"split_total_tflops\": %.3f", ms,
         (warps / 2 * ITER * 8 * 512 + nthreads / 2 * ITER * 8 * 2) / (ms * 1e-3) / 1e12);

  // DMMA dependent-chain latency / ILP: W warps per SMSP (block = 128*W threads, 1 block per SM), A chains per warp
  {
    int cfgW[3] = {1, 2, 4};
    for (int wi = 0; wi < 3; ++wi) {
      int W = cfgW[wi];
      int thr = 128 * W;
      double wps = double(sms) * 4 * W;   // warps
      float m1 = time_ms([&] { k_dmma<1><<<sms, thr>>>(out, 1.0000001, 1e-9); });
      float m2 = time_ms([&] { k_dmma<2><<<sms, thr>>>(out, 1.0000001, 1e-9); });
      float m4 = time_ms([&] { k_dmma<4><<<sms, thr>>>(out, 1.0000001, 1e-9); });
      double clk = double(p.clockRate) * 1e3;
      printf(", \"dmma_cyc_per_instr_W%d_A1\": %.2f, \"dmma_cyc_per_instr_W%d_A2\": %.2f, \"dmma_cyc_per_instr_W%d_A4\": %.2f",
             W, m1 * 1e-3 * clk / (ITER * 1.0 * W), W, m2 * 1e-3 * clk / (ITER * 2.0 * W), W, m4 * 1e-3 * clk / (ITER * 4.0 * W));
      (void)wps;
    }
  }

  const char* names[6] = {"exp", "sqrt", "erfc", "div", "rsqrt", "log"};
  float tms[6];
  tms[0] = time_ms([&] { k_trans<0><<<blocks, threads>>>(out, 0.3); });
  tms[1] = time_ms([&] { k_trans<1><<<blocks, threads>>>(out, 0.3); });
  tms[2] = time_ms([&] { k_trans<2><<<blocks, threads>>>(out, 0.3); });
  tms[3] = time_ms([&] { k_trans<3><<<blocks, threads>>>(out, 0.3); });
  tms[4] = time_ms([&] { k_trans<4><<<blocks, threads>>>(out, 0.3); });
  tms[5] = time_ms([&] { k_trans<5><<<blocks, threads>>>(out, 0.3); });
  for (int i = 0; i < 6; ++i)
    printf(", \"%s_gops\": %.2f", names[i], nthreads * (ITER / 8) * 4 / (tms[i] * 1e-3) / 1e9);

  const char* lnames[5] = {"lds64_bcast", "lds64_distinct", "lds128_bcast", "lds128_distinct", "lds32_bcast"};
  float lms[5];
  size_t smb = 4096 * sizeof(double);
  lms[0] = time_ms([&] { k_lds<0><<<blocks, threads, smb>>>(out, 7); });
  lms[1] = time_ms([&] { k_lds<1><<<blocks, threads, smb>>>(out, 7); });
  lms[2] = time_ms([&] { k_lds<2><<<blocks, threads, smb>>>(out, 7); });
  lms[3] = time_ms([&] { k_lds<3><<<blocks, threads, smb>>>(out, 7); });
  lms[4] = time_ms([&] { k_lds<4><<<blocks, threads, smb>>>(out, 7); });
  for (int i = 0; i < 5; ++i)
    printf(", \"%s_warp_instr_per_clk_per_sm\": %.3f", lnames[i],
           warps * (ITER / 4) * 8 / (lms[i] * 1e-3) / (double(p.clockRate) * 1e3) / sms);
  printf("}\n");
  return 0;
}
