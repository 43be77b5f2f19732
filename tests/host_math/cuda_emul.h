// TEST-ONLY: a minimal emulation of the CUDA execution model for g++, used to run the
// kernels' SOURCE (csrc/*.cuh) on a machine without a GPU.
//
// One OS thread stands for one CUDA thread of a CTA; CTAs of a grid run one after the other.
// `__shared__` becomes function-static storage, dynamic shared memory a static buffer,
// `__syncthreads` a CTA-wide barrier, warp collectives (`__shfl_*_sync`, `__reduce_min_sync`,
// `__ballot_sync`, `__syncwarp`) a 32-thread barrier around a scratch line, global atomics the
// GCC builtins.  Bulk copies and mbarriers are emulated in csrc/gd_common.cuh
// (GD_HOST_EMULATION branch) as synchronous memcpys.  What this checks is index, tiling,
// scheduling and reduction LOGIC; arithmetic runs as the host instantiation of gd_math.cuh
// (plain float: no MUFU, no FFMA2), and asynchronous-copy races cannot be seen.
// Include this header BEFORE any csrc header.  Never part of the shipped library.
#pragma once
#include <cuda_runtime.h>

#include <string.h>

#include <atomic>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define GD_HOST_EMULATION 1

struct EmuDim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local EmuDim3 threadIdx, blockIdx;
static EmuDim3 gridDim, blockDim;

struct EmuCta {
  explicit EmuCta(int nthreads) : all(nthreads) {
    for (int w = 0; w < (nthreads + 31) / 32; ++w) warps.emplace_back(new Warp());
  }
  struct Warp {
    std::barrier<> bar{32};
    unsigned long long scratch[32];
  };
  std::barrier<> all;
  std::vector<std::unique_ptr<Warp>> warps;
  int vote = 0;
};
static EmuCta* g_cta = nullptr;
alignas(128) static unsigned char g_emu_smem[228 * 1024];

static inline unsigned char* emu_dynamic_smem() { return g_emu_smem; }
static inline void emu_yield() { std::this_thread::yield(); }
static inline void __syncthreads() { g_cta->all.arrive_and_wait(); }
static inline int __syncthreads_or(int p) {
  if (p) __atomic_fetch_or(&g_cta->vote, 1, __ATOMIC_SEQ_CST);
  g_cta->all.arrive_and_wait();
  const int r = __atomic_load_n(&g_cta->vote, __ATOMIC_SEQ_CST);
  g_cta->all.arrive_and_wait();
  if (threadIdx.x == 0) g_cta->vote = 0;
  g_cta->all.arrive_and_wait();
  return r;
}
static inline int __syncthreads_and(int p) {
  if (!p) __atomic_fetch_or(&g_cta->vote, 1, __ATOMIC_SEQ_CST);
  g_cta->all.arrive_and_wait();
  const int r = __atomic_load_n(&g_cta->vote, __ATOMIC_SEQ_CST);
  g_cta->all.arrive_and_wait();
  if (threadIdx.x == 0) g_cta->vote = 0;
  g_cta->all.arrive_and_wait();
  return !r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline EmuCta::Warp& emu_warp() { return *g_cta->warps[threadIdx.x >> 5]; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp().bar.arrive_and_wait(); }

// every lane publishes `v`, then reads the line: the basis of all warp collectives
template <typename T, typename F>
static inline auto emu_exchange(T v, F&& pick) {
  static_assert(sizeof(T) <= 8, "32/64-bit values only");
  EmuCta::Warp& w = emu_warp();
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  w.scratch[threadIdx.x & 31] = bits;
  w.bar.arrive_and_wait();
  T line[32];
  for (int i = 0; i < 32; ++i) memcpy(&line[i], &w.scratch[i], sizeof(T));
  auto r = pick(line);
  w.bar.arrive_and_wait();
  return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) {
  return emu_exchange(v, [&](const T* line) { return line[src & 31]; });
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
  const int me = (int)(threadIdx.x & 31);
  return emu_exchange(v, [&](const T* line) { return line[(me ^ mask) & 31]; });
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  return emu_exchange(v, [&](const unsigned* line) {
    unsigned m = line[0];
    for (int i = 1; i < 32; ++i) m = line[i] < m ? line[i] : m;
    return m;
  });
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  return emu_exchange((unsigned)(p ? 1u : 0u), [&](const unsigned* line) {
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= line[i] << i;
    return m;
  });
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
template <typename T>
static inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
static inline T __ldcs(const T* p) { return *p; }
static inline unsigned long long __ldcg(const unsigned long long* p) {
  return __atomic_load_n(p, __ATOMIC_SEQ_CST);
}
static inline double __ldcg(const double* p) {
  unsigned long long b = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_SEQ_CST);
  double d;
  memcpy(&d, &b, 8);
  return d;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline unsigned atomicOr(unsigned* p, unsigned v) {
  return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned __ldcg(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline float __ldg(const float* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
using std::min;

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __device__
#define __device__
#undef __forceinline__
#define __forceinline__ inline

// run `kernel(args)` over a grid, one CTA at a time, `nthreads` OS threads per CTA
template <typename K, typename A>
static void emu_launch(K kernel, unsigned gx, unsigned gy, int nthreads, const A& args) {
  gridDim.x = gx;
  gridDim.y = gy;
  blockDim.x = (unsigned)nthreads;
  for (unsigned by = 0; by < gy; ++by) {
    for (unsigned bx = 0; bx < gx; ++bx) {
      EmuCta cta(nthreads);
      g_cta = &cta;
      std::vector<std::thread> ts;
      ts.reserve(nthreads);
      for (int t = 0; t < nthreads; ++t) {
        ts.emplace_back([=, &args]() {
          threadIdx.x = (unsigned)t;
          blockIdx.x = bx;
          blockIdx.y = by;
          kernel(args);
        });
      }
      for (auto& th : ts) th.join();
      g_cta = nullptr;
    }
  }
}
