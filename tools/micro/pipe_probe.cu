// Micro-probe (measurement only, not part of the library): what do the DATA-MOVEMENT
// pipelines that could feed the fused GD-loss kernel cost, in GB/s and in board power,
// when they carry the kernel's traffic mix and nothing else?
//
// Per box pair the loss kernel reads 28 B (pred) + 28 B (target) + 4 B (weight) and writes
// 28 B (grad): 88 B, 68 % reads.  DESIGN.md section 4 shows the kernel is bounded by the
// 1000 W board cap: with the math removed the shipped pipeline (bulk copies -> shared
// memory -> LDS.128 -> STS.128 -> bulk store) already draws ~994 W at 6.2 TB/s while a plain
// device copy draws 775 W at 6.5 TB/s.  This probe separates the pieces; each mode moves the
// same bytes and computes grad = pred + target * weight (one FFMA per element):
//
//   0 flat      coalesced LDG.128 / STG.128 over the arrays viewed as flat float4 (no AoS
//               transposition at all: the floor for this read/write mix)
//   1 tma_only  per-warp 2-stage ring of bulk copies into shared memory, the pred tile is
//               written back by a bulk store straight from the stage (no LDS / STS, no math)
//   2 tma_regs  the shipped pipeline: bulk copies -> LDS.128 (lane owns 4 consecutive rows)
//               -> FFMA -> STS.128 -> bulk store
//   3 direct    no shared memory: every lane loads its 4 consecutive rows (112 B per array)
//               with 7 + 7 + 1 LDG.128 at a 112 B lane stride and stores them with 7 STG.128
//   4 tma_stg   bulk copies -> LDS.128 -> FFMA -> STG.128 straight from registers (lane
//               stride 112 B); no staging of the output in shared memory
//
// Usage:  pipe_probe <mode> [seconds=2] [log2_rows=24]     (prints one JSON line)
// tools/pipe_probe.py builds it, samples nvidia-smi while each mode runs and adds the
// shipped kernel with and without math for reference.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));          \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all0() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr int kRowsPerLane = 4;
constexpr int kTileRows = 32 * kRowsPerLane;        // 128 rows per warp tile
constexpr int kTileFloats = kTileRows * 7;          // 896
constexpr int kTileBytes = kTileFloats * 4;         // 3584
constexpr int kWTileBytes = kTileRows * 4;          // 512
constexpr int kStageBytes = 2 * kTileBytes + kWTileBytes;

// element e (0..27) of a lane's 4 consecutive rows belongs to row e / 7
__device__ __forceinline__ float wsel(const float4& w, int e) {
  const int r = e / 7;
  return r == 0 ? w.x : (r == 1 ? w.y : (r == 2 ? w.z : w.w));
}
__device__ __forceinline__ float4 fma4(const float4& a, const float4& b, const float4& w, int j) {
  return make_float4(fmaf(b.x, wsel(w, 4 * j), a.x), fmaf(b.y, wsel(w, 4 * j + 1), a.y),
                     fmaf(b.z, wsel(w, 4 * j + 2), a.z), fmaf(b.w, wsel(w, 4 * j + 3), a.w));
}

// ---- mode 0: flat coalesced copy-like kernel ---------------------------------------------
__global__ void __launch_bounds__(256) k_flat(const float4* __restrict__ p,
                                              const float4* __restrict__ t,
                                              const float* __restrict__ w, float4* __restrict__ g,
                                              long long n4) {
  const long long stride = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a[4], b[4];
    float ww[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = __ldcs(p + i + u * stride);
      b[u] = __ldcs(t + i + u * stride);
#pragma unroll
      for (int c = 0; c < 4; ++c) ww[u][c] = __ldg(w + (4 * (i + u * stride) + c) / 7);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      __stcs(g + i + u * stride,
             make_float4(fmaf(b[u].x, ww[u][0], a[u].x), fmaf(b[u].y, ww[u][1], a[u].y),
                         fmaf(b[u].z, ww[u][2], a[u].z), fmaf(b[u].w, ww[u][3], a[u].w)));
  }
  for (; i < n4; i += stride) {
    const float4 a = __ldcs(p + i), b = __ldcs(t + i);
    float ww[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ww[c] = __ldg(w + (4 * i + c) / 7);
    __stcs(g + i, make_float4(fmaf(b.x, ww[0], a.x), fmaf(b.y, ww[1], a.y), fmaf(b.z, ww[2], a.z),
                              fmaf(b.w, ww[3], a.w)));
  }
}

// ---- modes 1, 2, 4: per-warp ring of bulk copies -------------------------------------------
// MODE 1: store the pred stage back as it is.  MODE 2: LDS -> FFMA -> STS -> bulk store.
// MODE 4: LDS -> FFMA -> STG.128 from registers.
template <int MODE>
__global__ void __launch_bounds__(384, 1) k_ring(const float* __restrict__ pred,
                                                 const float* __restrict__ target,
                                                 const float* __restrict__ weight,
                                                 float* __restrict__ grad, long long ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kPerWarp = 2 * kStageBytes + kTileBytes + 16;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  unsigned char* base = smem + (size_t)warp * kPerWarp;
  float* og = reinterpret_cast<float*>(base + 2 * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * kStageBytes + kTileBytes);
  const long long gwarp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int my_n = (int)((ntiles - gwarp + nwarps - 1) / nwarps);   // tiles gwarp + k nwarps
  const uint64_t policy = policy_evict_first();
  const bool leader = elect_one();

  auto issue = [&](int i) {
    const long long row0 = (gwarp + (long long)i * nwarps) * kTileRows;
    const int s = i & 1;
    unsigned char* st = base + s * kStageBytes;
    mbar_arrive_expect_tx(&bars[s], kStageBytes);
    bulk_load(st, pred + row0 * 7, kTileBytes, &bars[s], policy);
    bulk_load(st + kTileBytes, target + row0 * 7, kTileBytes, &bars[s], policy);
    bulk_load(st + 2 * kTileBytes, weight + row0, kWTileBytes, &bars[s], policy);
  };
  if (leader) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    for (int i = 0; i < (my_n < 2 ? my_n : 2); ++i) issue(i);
  }
  __syncwarp();
  for (int i = 0; i < my_n; ++i) {
    const int s = i & 1;
    const long long row0 = (gwarp + (long long)i * nwarps) * kTileRows;
    unsigned char* st = base + s * kStageBytes;
    mbar_wait(&bars[s], (uint32_t)((i >> 1) & 1));
    if (MODE == 1) {
      // the stage itself is the source of the store: it may only be refilled once the
      // store has read it
      if (leader) {
        bulk_store(grad + row0 * 7, st, kTileBytes);
        bulk_commit();
        bulk_wait_read0();
        if (i + 2 < my_n) issue(i + 2);
      }
      __syncwarp();
      continue;
    }
    const float4* sp = reinterpret_cast<const float4*>(st) + 7 * lane;
    const float4* stg = reinterpret_cast<const float4*>(st + kTileBytes) + 7 * lane;
    const float4 w = reinterpret_cast<const float4*>(st + 2 * kTileBytes)[lane];
    float4 a[7], b[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      a[j] = sp[j];
      b[j] = stg[j];
    }
    if (MODE == 2 && leader && i > 0) bulk_wait_read0();   // og free again
    __syncwarp();
    if (leader && i + 2 < my_n) issue(i + 2);
    if (MODE == 2) {
      float4* o4 = reinterpret_cast<float4*>(og) + 7 * lane;
#pragma unroll
      for (int j = 0; j < 7; ++j) o4[j] = fma4(a[j], b[j], w, j);
      fence_proxy_async_smem();
      __syncwarp();
      if (leader) {
        bulk_store(grad + row0 * 7, og, kTileBytes);
        bulk_commit();
      }
    } else {   // MODE 4
      float4* g4 = reinterpret_cast<float4*>(grad + row0 * 7) + 7 * lane;
#pragma unroll
      for (int j = 0; j < 7; ++j) __stcs(g4 + j, fma4(a[j], b[j], w, j));
    }
  }
  if (leader && MODE != 4) bulk_wait_all0();
}

// ---- mode 3: registers only ----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_direct(const float4* __restrict__ p4,
                                                const float4* __restrict__ t4,
                                                const float4* __restrict__ w4,
                                                float4* __restrict__ g4, long long ntiles) {
  const int lane = threadIdx.x & 31;
  const long long gwarp = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long tile = gwarp; tile < ntiles; tile += nwarps) {
    const long long o = tile * (kTileFloats / 4) + 7 * lane;
    float4 a[7], b[7];
    const float4 w = __ldcs(w4 + tile * 32 + lane);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      a[j] = __ldcs(p4 + o + j);
      b[j] = __ldcs(t4 + o + j);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) __stcs(g4 + o + j, fma4(a[j], b[j], w, j));
  }
}

double now() {
  timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

}  // namespace

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const double seconds = argc > 2 ? atof(argv[2]) : 2.0;
  const int lg = argc > 3 ? atoi(argv[3]) : 24;
  if (mode < 0 || mode > 4 || lg < 10 || lg > 27) {
    fprintf(stderr, "usage: pipe_probe <mode 0..4> [seconds] [log2_rows 10..27]\n");
    return 1;
  }
  const long long n = 1LL << lg;                     // rows, a multiple of 128
  const long long ntiles = n / kTileRows;
  const size_t box_bytes = (size_t)n * 28, w_bytes = (size_t)n * 4;
  float *pred, *target, *weight, *grad;
  CK(cudaMalloc(&pred, box_bytes));
  CK(cudaMalloc(&target, box_bytes));
  CK(cudaMalloc(&weight, w_bytes));
  CK(cudaMalloc(&grad, box_bytes));
  // deterministic contents (checked on a sample afterwards)
  std::vector<float> hp((size_t)n * 7), ht((size_t)n * 7), hw((size_t)n);
  uint32_t x = 12345u;
  auto rnd = [&]() {
    x = x * 1664525u + 1013904223u;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
  };
  for (auto& v : hp) v = rnd();
  for (auto& v : ht) v = rnd();
  for (auto& v : hw) v = 0.5f + rnd();
  CK(cudaMemcpy(pred, hp.data(), box_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(target, ht.data(), box_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(weight, hw.data(), w_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(grad, 0xff, box_bytes));

  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr int kRingWarps = 12;
  constexpr int kRingSmem = kRingWarps * (2 * kStageBytes + kTileBytes + 16);
  CK(cudaFuncSetAttribute(k_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));
  CK(cudaFuncSetAttribute(k_ring<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));
  CK(cudaFuncSetAttribute(k_ring<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));

  auto launch = [&]() {
    switch (mode) {
      case 0:
        k_flat<<<sms * 8, 256>>>((const float4*)pred, (const float4*)target, weight, (float4*)grad,
                                 n * 7 / 4);
        break;
      case 1: k_ring<1><<<sms, kRingWarps * 32, kRingSmem>>>(pred, target, weight, grad, ntiles); break;
      case 2: k_ring<2><<<sms, kRingWarps * 32, kRingSmem>>>(pred, target, weight, grad, ntiles); break;
      case 4: k_ring<4><<<sms, kRingWarps * 32, kRingSmem>>>(pred, target, weight, grad, ntiles); break;
      case 3:
        k_direct<<<sms * 8, 256>>>((const float4*)pred, (const float4*)target,
                                   (const float4*)weight, (float4*)grad, ntiles);
        break;
    }
  };
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  // result check on the first and last tile
  std::vector<float> hg(2 * kTileFloats);
  CK(cudaMemcpy(hg.data(), grad, kTileBytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hg.data() + kTileFloats, grad + (n - kTileRows) * 7, kTileBytes,
                cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int half = 0; half < 2; ++half) {
    const long long r0 = half ? n - kTileRows : 0;
    for (int e = 0; e < kTileFloats; ++e) {
      const long long idx = r0 * 7 + e;
      const float want = mode == 1 ? hp[idx] : fmaf(ht[idx], hw[idx / 7], hp[idx]);
      if (hg[half * kTileFloats + e] != want) ++bad;
    }
  }

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double t0 = now();
  long calls = 0;
  CK(cudaEventRecord(e0));
  while (now() - t0 < seconds) {
    for (int i = 0; i < 200; ++i) launch();
    calls += 200;
    CK(cudaDeviceSynchronize());
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  const double t1 = now();
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= calls;
  const char* names[] = {"flat", "tma_only", "tma_regs", "direct", "tma_stg"};
  // mode 1 does not write a function of target / weight but still reads them: same 88 B / row
  printf("{\"mode\": %d, \"name\": \"%s\", \"rows\": %lld, \"ms\": %.5f, \"GBps\": %.1f, "
         "\"t0\": %.3f, \"t1\": %.3f, \"calls\": %ld, \"mismatches\": %ld}\n",
         mode, names[mode], n, ms, 88.0 * n / ms / 1e6, t0, t1, calls, bad);
  return bad ? 3 : 0;
}
