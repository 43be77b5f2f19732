// Shared device/host helpers for the gd_loss_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/gd_loss_b200.h"
#include "gd_math.cuh"

namespace gdk {

constexpr int kThreads = 256;           // threads per CTA
constexpr int kTile = 256;              // rows per tile: one row per thread
constexpr int kRowBytes = 28;           // 7 fp32
constexpr int kTileBytes = kTile * kRowBytes;   // 7168, a multiple of 16

extern std::atomic<int64_t> g_launches;  // bench `gpu_launches`

constexpr int kMaxDevices = 64;
struct DeviceInfo {
  int sm_count;
};
const DeviceInfo& device_info();          // cached per current device
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}

inline gd::PairParams<float> make_pair_params(const gd_loss_config& c) {
  gd::PairParams<float> p;
  for (int i = 0; i < 3; ++i) p.off[i] = c.center_offset[i];
  const double a2 = (double)c.alpha * (double)c.alpha;       // ref:99 alpha*alpha
  p.alpha2 = (float)a2;
  p.inv_alpha2 = (float)(1.0 / a2);                          // ref:137,182
  p.tau = c.tau;
  p.tau_on = c.tau >= 1.0f ? 1 : 0;                          // ref:36
  p.fun = c.fun;
  p.flag = c.flag ? 1 : 0;
  return p;
}

inline bool config_ok(const gd_loss_config* c) {
  return c && c->loss_type >= 0 && c->loss_type < gd::kNumLossTypes &&
         c->fun >= GD_FUN_NONE && c->fun <= GD_FUN_NLOG;
}

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// one lane of the (converged) warp; the same lane every time for a full mask
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) --------------
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
// global -> shared, completion counted in bytes on `bar`; L2 evict-first (streamed once)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* gdst, const void* smem_src, uint32_t bytes,
                                                uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                   gdst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- system-scope accesses to peer memory (NVLink P2P) ----------------------
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

#elif defined(GD_HOST_EMULATION)
inline void st_relaxed_sys(double* p, double v) { __atomic_store(p, &v, __ATOMIC_RELAXED); }
inline double ld_relaxed_sys(const double* p) {
  double v;
  __atomic_load(p, &v, __ATOMIC_RELAXED);
  return v;
}
inline void st_release_sys(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned int ld_acquire_sys(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
// Host stand-ins for the copy-engine / mbarrier primitives (tests/host_math/loss_emul.cpp
// runs the kernels' SOURCE with one OS thread per CUDA thread).  A bulk copy is a memcpy
// done by the issuing thread; the mbarrier keeps its phase bit and pending byte count in the
// same 8 bytes of "shared memory".  The harness provides threadIdx, __shfl_*_sync,
// __syncwarp, ... before including this header.
struct EmuMbar {
  volatile uint32_t phase;
  volatile int32_t pending;
};
inline bool elect_one() { return (threadIdx.x & 31u) == 0u; }
inline void mbar_init(uint64_t* bar, uint32_t) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->phase = 0;
  b->pending = 0;
}
inline void fence_mbar_init() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void fence_proxy_async_smem() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  reinterpret_cast<EmuMbar*>(bar)->pending = (int32_t)bytes;
}
inline void emu_complete_tx(uint64_t* bar, uint32_t bytes) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->pending = b->pending - (int32_t)bytes;
  if (b->pending == 0) {
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    b->phase = b->phase ^ 1u;                  // phase complete
  }
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  while (__atomic_load_n(&b->phase, __ATOMIC_SEQ_CST) == parity) emu_yield();
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
}
inline void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t) {
  memcpy(smem_dst, gsrc, bytes);
  emu_complete_tx(bar, bytes);
}
inline void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  memcpy(gdst, smem_src, bytes);
}
inline void bulk_store_hint(void* gdst, const void* smem_src, uint32_t bytes, uint64_t) {
  memcpy(gdst, smem_src, bytes);
}
inline void bulk_commit() {}
template <int N>
inline void bulk_wait_read() {}
template <int N>
inline void bulk_wait_all() {}
inline uint64_t policy_evict_first() { return 0; }
inline uint64_t policy_evict_normal() { return 0; }
inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__ / GD_HOST_EMULATION

}  // namespace gdk
