// Micro-probe (measurement only, not part of the library): FP32 throughput and board
// power of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) chains on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// CHAINS independent dependency chains per thread; each loop iteration does CHAINS FMAs
// (scalar: 1 flop-pair each; packed: 2 each).
template <int CHAINS, bool PACKED>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (PACKED) {
    unsigned long long x[CHAINS];
    const unsigned long long b = pack(seed, seed * 0.5f), c = pack(0.25f, 0.125f);
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = pack(tid * 1e-6f + k, k * 0.5f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) x[k] = fma2(x[k], b, c);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s ^= x[k];
    if (s == 0x1234567ull) out[tid] = 1.0f;
  } else {
    float x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = tid * 1e-6f + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) x[k] = fma1(x[k], seed, 0.25f);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += x[k];
    if (s == 0.123f) out[tid] = s;
  }
}

template <int CHAINS, bool PACKED>
void run(const char* name, int warps_per_sm, float* out) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * warps_per_sm / 8;      // 256 threads = 8 warps per block
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<CHAINS, PACKED><<<blocks, 256>>>(out, iters, 1.0001f);
  cudaDeviceSynchronize();
  // ~2 s of back-to-back launches so nvidia-smi sees the load
  float ms_total = 0;
  int launches = 0;
  cudaEventRecord(e0);
  for (int r = 0; r < 200; ++r) {
    probe<CHAINS, PACKED><<<blocks, 256>>>(out, iters, 1.0001f);
    ++launches;
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms_total, e0, e1);
  const double fmas = (double)blocks * 256 * iters * CHAINS * (PACKED ? 2 : 1) * launches;
  char cmd[256];
  printf("%-28s warps/SM %2d chains %d: %7.2f TFMA/s  (%.1f ms)\n", name, warps_per_sm, CHAINS,
         fmas / (ms_total * 1e-3) / 1e12, ms_total);
  fflush(stdout);
  (void)cmd;
}

int main() {
  float* out;
  cudaMalloc(&out, 1 << 26);
  for (int w : {12, 32, 64}) {
    run<4, false>("scalar FFMA", w, out);
    run<4, true>("packed FFMA2", w, out);
    run<8, false>("scalar FFMA", w, out);
    run<8, true>("packed FFMA2", w, out);
  }
  return 0;
}
