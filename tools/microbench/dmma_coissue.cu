// Does instruction issue overlap with the FP64 MMA pipe on B200? One warp stream: 8 independent DMMA chains per iteration
// with K other instructions per DMMA (KIND 0: IMAD, 1: LOP3/IADD3 (alu), 2: LDS.64 distinct, 3: SHFL, 4: FFMA).
// Prints time relative to the DMMA-only loop: if issue overlapped freely the ratio would stay 1.0 until K ~ 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_coissue dmma_coissue.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int ITER = 4096;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int KIND, int K>
__global__ void k_mix(double* out, double a, double b, int ia, int ib) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double c0[8], c1[8];
  int x[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; x[i] = threadIdx.x * 7 + i; f[i] = threadIdx.x + i; }
  double ls = 0;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma(c0[i], c1[i], a, b);
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int r = (i + q) & 7;
        if (KIND == 0) x[r] = x[r] * ia + ib;
        else if (KIND == 1) x[r] = (x[r] ^ ia) + ib;
        else if (KIND == 2) { ls += sm[(x[r] + it) & 1023]; }
        else if (KIND == 3) x[r] = __shfl_xor_sync(0xffffffffu, x[r], 1 + (q & 15));
        else f[r] = fmaf(f[r], (float)a, (float)b);
      }
    }
  }
  double s = ls;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i] + x[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

template <int KIND>
static void sweep(const char* name, double* out, int blocks, int threads, float base) {
  float m1 = time_ms([&] { k_mix<KIND, 1><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
  float m2 = time_ms([&] { k_mix<KIND, 2><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
  float m4 = time_ms([&] { k_mix<KIND, 4><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
  float m8 = time_ms([&] { k_mix<KIND, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
  float m16 = time_ms([&] { k_mix<KIND, 16><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
  printf(", \"%s_per_dmma_1_2_4_8_16_time_ratio\": [%.3f, %.3f, %.3f, %.3f, %.3f]", name, m1 / base, m2 / base, m4 / base,
         m8 / base, m16 / base);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out;
  printf("{\"gpu\": \"%s\"", p.name);
  // 4 and 8 warps per SM sub-partition
  for (int cfg = 0; cfg < 2; ++cfg) {
    int threads = 256, blocks = sms * (cfg == 0 ? 2 : 4);
    CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    float base = time_ms([&] { k_mix<0, 0><<<blocks, threads>>>(out, 1.0000001, 1e-9, 3, 5); });
    double warps = double(blocks) * threads / 32;
    printf(", \"warps_per_smsp\": %d, \"dmma_only_tflops\": %.2f", cfg == 0 ? 4 : 8, warps * ITER * 8 * 512 / (base * 1e-3) / 1e12);
    sweep<0>(cfg == 0 ? "w4_imad" : "w8_imad", out, blocks, threads, base);
    sweep<1>(cfg == 0 ? "w4_alu" : "w8_alu", out, blocks, threads, base);
    sweep<2>(cfg == 0 ? "w4_lds64" : "w8_lds64", out, blocks, threads, base);
    sweep<3>(cfg == 0 ? "w4_shfl" : "w8_shfl", out, blocks, threads, base);
    sweep<4>(cfg == 0 ? "w4_ffma" : "w8_ffma", out, blocks, threads, base);
    CK(cudaFree(out));
  }
  printf("}\n");
  return 0;
}
