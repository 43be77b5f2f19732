// Pairwise N x M Gaussian-distance kernels for B200 (sm_100a).
//
// New surface (the reference has only IoU matrices,
// core/bbox/assigners/sim_ota_3d_assigner.py:91-93): out[i,j] =
// postprocess(distance(boxes1[i], boxes2[j])), equal to the element-wise loss path of
// gaussian_distance_loss.py on the broadcast-expanded pairs -- plus the consumer
// fused in (row f2): per-row and per-column arg-minima for assigner use, so the
// matrix need not be written at all.
//
// Mapping.  A CTA owns a tile of kRowsPerCta rows of boxes1, converted ONCE into
// BoxGauss (centre, half extents, sin/cos yaw and every other per-box quantity:
// squares, a*b, A-B, reciprocals, the box's factor of the gwd normaliser) in shared
// memory.  Lanes map to columns: a lane keeps CPL (1 or 2) column boxes in registers
// (columns base + 32 q + lane, so the stores of one q stay coalesced), i.e. a warp covers
// 32 CPL columns and every shared-memory broadcast of a row Gaussian serves CPL pairs;
// the 8 warps are arranged wx x wy with wx = number of column groups in use (1, 2, 4 or
// 8) and wy = 8 / wx row phases, so small M (a handful of ground-truth boxes) still fills
// the CTA.  The inner loop evaluates the branch-free FAST cores (robust cores on a cold
// branch); sin/cos of the yaw difference come from the angle-difference identities; the
// log1p of the post map is the short pairwise version (Mth::log1p_lean, ~4e-7).  Stores
// are coalesced 4 B/pair streaming stores through a row pointer that is advanced, not
// recomputed.  FP32 CUDA-core math, no tensor cores (not a contraction).
// The kernel is issue bound, so the design goal is instructions per pair (round 1: ~105
// matrix / ~160 fused; profiles/r02*_pairwise.md for this version).
//
// Reductions.  Values map to order-preserving 32-bit keys (NaN lowest, as torch.min
// propagates NaN).  Row minimum: in-lane minimum over the CPL columns, one REDUX.MIN and
// CPL ballots per warp and row, lowest column wins ties.  Column minimum: a compare/select per pair in the owning lane,
// merged across CTAs with one 64-bit atomic per column on (key << 32 | row) -- lowest
// row wins ties -- and unpacked by the last CTA to finish (atomic ticket), which also
// restores the workspace.  Both reductions and the matrix come out of the SAME
// instruction sequence per pair, so indices derived from either are bit-identical.
#pragma once
#include "gd_common.cuh"
#include "gd_packed.cuh"

namespace gdk {

constexpr int kRowsPerCta = 64;
constexpr int kWarps = kThreads / 32;

struct PairwiseArgs {
  const float* b1;
  long long n;
  const float* b2;
  long long m;
  float* out;                      // nullable when reducing
  long long out_stride;
  int similarity;                  // write 1 - value (assigners' "larger is closer")
  int force_cpl1;                  // GD_PAIR_CPL1: one column per lane
  float* row_min;                  // [n]   REDUCE only
  int* row_argmin;                 // [n]
  unsigned long long* col_keys;    // [m] workspace holding ~key (so zero is the identity of the
                                   // atomicMax merge): zero on entry and on exit; nullable
  float* col_min;                  // [m]
  int* col_argmin;                 // [m]
  unsigned int* ticket;            // zero on entry and on exit
  gd::PairParams<float> pp;
};

// order-preserving float -> uint32 (NaN -> 0: lowest)
__device__ __forceinline__ unsigned int order_key(float v) {
  unsigned int b = __float_as_uint(v);
  b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;
  return (v != v) ? 0u : b;
}
__device__ __forceinline__ float key_value(unsigned int b) {
  if (b == 0u) return __uint_as_float(0x7fc00000u);
  b ^= (b >> 31) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(b);
}

// column groups (warps side by side) in use for m columns when one warp covers `warp_cols`
__host__ __device__ inline int pairwise_wx(long long m, int warp_cols) {
  return m <= warp_cols ? 1 : (m <= 2LL * warp_cols ? 2 : (m <= 4LL * warp_cols ? 4 : 8));
}

// SPEC < 0: fun / tau_on / flag at run time; SPEC >= 0: bits [1:0] fun, [2] tau_on,
// [3] flag compile-time (as gd_warp_kernel).  CPL: columns per lane (1 or 2).
template <int LOSS, int SPEC, bool REDUCE, int CPL>
__global__ void __launch_bounds__(kThreads) gd_pairwise_kernel(const PairwiseArgs a) {
  __shared__ gd::BoxGauss<float> s_rows[kRowsPerCta];
  __shared__ unsigned long long s_best[REDUCE ? kRowsPerCta : 1][kWarps];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gd::PairParams<float> pp = a.pp;
  if (SPEC >= 0) {
    pp.fun = SPEC & 3;
    pp.tau_on = (SPEC >> 2) & 1;
    pp.flag = (SPEC >> 3) & 1;
  }
  pp.lean = 1;                                  // short log1p of the pairwise value path
  constexpr int kWarpCols = 32 * CPL;           // columns one warp covers per pass
  const int wx = pairwise_wx(a.m, kWarpCols);
  const int wy = kWarps / wx;
  const int cgrp = warp % wx, ry = warp / wx;
  const long long chunk = (long long)kWarpCols * wx;
  const bool want_col = REDUCE && a.col_keys != nullptr;
  const bool one_chunk = a.m <= chunk;         // column minima can stay in registers across tiles
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  unsigned int cbest[CPL], crow[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    cbest[q] = 0xffffffffu;
    crow[q] = 0u;
  }

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kRowsPerCta;
    const int rows = (int)min((long long)kRowsPerCta, a.n - row0);
    __syncthreads();                           // previous tile fully consumed
    if (tid < rows) s_rows[tid] = gd::box_gauss(a.b1 + (row0 + tid) * 7, pp);
    if (REDUCE) {
      for (int i = tid; i < kRowsPerCta * kWarps; i += kThreads) (&s_best[0][0])[i] = ~0ull;
    }
    __syncthreads();
    for (long long c0 = (long long)blockIdx.y * chunk; c0 < a.m; c0 += (long long)gridDim.y * chunk) {
      const long long jb = c0 + (long long)kWarpCols * cgrp;     // first column of this warp
      if (jb >= a.m) continue;                 // this warp's columns are all past the end
      gd::BoxGauss<float> t[CPL];
      bool live[CPL];
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const long long j = jb + 32 * q + lane;
        live[q] = j < a.m;
        if (live[q]) t[q] = gd::box_gauss(a.b2 + j * 7, pp);
        else t[q] = s_rows[0];                 // any valid box: the result is discarded
        if (want_col && !one_chunk) cbest[q] = 0xffffffffu;
      }
      // row pointer of this lane's first column, advanced by wy rows per iteration
      float* orow = a.out != nullptr ? a.out + (row0 + ry) * a.out_stride + jb + lane : nullptr;
      const long long ostep = (long long)wy * a.out_stride;
#pragma unroll 2
      for (int r = ry; r < rows; r += wy) {
        float v[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) v[q] = gd::pair_value_auto<float, LOSS>(s_rows[r], t[q], pp);
        if (a.out != nullptr) {
#pragma unroll
          for (int q = 0; q < CPL; ++q)
            if (live[q]) __stcs(orow + 32 * q, a.similarity ? 1.0f - v[q] : v[q]);
          orow += ostep;
        }
        if (REDUCE) {
          unsigned int key[CPL];
#pragma unroll
          for (int q = 0; q < CPL; ++q) key[q] = live[q] ? order_key(v[q]) : 0xffffffffu;
          unsigned int kmin = key[0];
#pragma unroll
          for (int q = 1; q < CPL; ++q) kmin = key[q] < kmin ? key[q] : kmin;
          const unsigned int mn = __reduce_min_sync(0xffffffffu, kmin);
          // lowest column holding the minimum: the columns of q = 0 precede those of q = 1
          unsigned int col = 0u;
#pragma unroll
          for (int q = CPL - 1; q >= 0; --q) {
            const unsigned int who = __ballot_sync(0xffffffffu, key[q] == mn);
            if (who) col = 32u * q + (unsigned int)(__ffs(who) - 1);
          }
          if (lane == 0 && mn != 0xffffffffu) {
            const unsigned long long k64 =
                ((unsigned long long)mn << 32) | (unsigned int)(jb + col);
            // one chunk: this (row, warp) slot is visited once per tile -- plain store
            if (one_chunk || k64 < s_best[r][warp]) s_best[r][warp] = k64;
          }
          if (want_col) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              if (key[q] < cbest[q]) {         // rows ascend within a lane: first minimum kept
                cbest[q] = key[q];
                crow[q] = (unsigned int)(row0 + r);
              }
            }
          }
        }
      }
      if (want_col && !one_chunk) {
#pragma unroll
        for (int q = 0; q < CPL; ++q)
          if (live[q] && cbest[q] != 0xffffffffu)
            atomicMax(a.col_keys + jb + 32 * q + lane,
                      ~(((unsigned long long)cbest[q] << 32) | crow[q]));
      }
    }
    if (REDUCE) {
      __syncthreads();
      if (tid < rows) {
        unsigned long long k = s_best[tid][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) k = s_best[tid][w] < k ? s_best[tid][w] : k;
        a.row_min[row0 + tid] = key_value((unsigned int)(k >> 32));
        a.row_argmin[row0 + tid] = (int)(unsigned int)(k & 0xffffffffu);
      }
    }
  }
  if (want_col) {
    if (one_chunk) {
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const long long j = (long long)kWarpCols * cgrp + 32 * q + lane;
        if (j < a.m && cbest[q] != 0xffffffffu)
          atomicMax(a.col_keys + j, ~(((unsigned long long)cbest[q] << 32) | crow[q]));
      }
    }
    // last CTA to finish unpacks the column keys and restores the workspace
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (long long j = tid; j < a.m; j += kThreads) {
        const unsigned long long k = ~__ldcg(a.col_keys + j);
        a.col_min[j] = key_value((unsigned int)(k >> 32));
        a.col_argmin[j] = (int)(unsigned int)(k & 0xffffffffu);
        a.col_keys[j] = 0ull;
      }
      if (tid == 0) *a.ticket = 0u;
    }
  }
}

// Host-side launchers.  tests/host_math/pairwise_emul.cpp compiles THIS header with g++
// (GD_HOST_EMULATION: one OS thread per CUDA thread, barriers for __syncthreads and the
// warp collectives) to run the kernels' index / reduction logic on a machine without a GPU;
// it supplies its own launch loop, so the <<< >>> code is left out there.
#if !defined(GD_HOST_EMULATION)
// GD_B200_PAIR_CPL=1|2 pins the columns per lane (measurements); default: 2 once a warp's 64
// columns are at least half used.
inline int pairwise_cpl(long long m) {
  static const int forced = [] {
    const char* e = getenv("GD_B200_PAIR_CPL");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  return m > 32 ? 2 : 1;
}
// Two columns per lane pay for the two light distances only (matrix: gwd3d +2.5 %, kld3d
// +2.7 % at 200k x 256, fused reductions: no difference; bd3d -9 %: register pressure) --
// profiles/r02_pairwise.md.  The choice depends on the distance alone, so the matrix-only and
// the fused-reduction launches of one distance run the same mapping (their values are compared
// bit for bit by the tests: indices derived from either must agree).
template <int LOSS>
constexpr bool pairwise_cpl2_pays() {
  return LOSS == gd::kGwd || LOSS == gd::kKld;
}

template <int LOSS, int SPEC, bool REDUCE, int CPL>
int launch_pairwise_cpl(const PairwiseArgs& a, cudaStream_t st) {
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  if (ntiles > 2147483647LL || a.m > 0x7fffffffLL || a.n > 0xffffffffLL) return GD_ERR_BAD_ARG;
  const int wx = pairwise_wx(a.m, 32 * CPL);
  dim3 grid;
  if (REDUCE) {                              // every CTA walks all columns of its rows
    long long gx = ntiles;
    if (a.col_keys) {                        // persistent: bounds the column atomics
      const long long cap = (long long)device_info().sm_count * 6;
      if (gx > cap) gx = cap;
    }
    grid = dim3((unsigned)gx, 1);
  } else {
    long long gy = (a.m + 32LL * CPL * wx - 1) / (32LL * CPL * wx);
    if (gy > 65535) gy = 65535;
    grid = dim3((unsigned)ntiles, (unsigned)gy);
  }
  gd_pairwise_kernel<LOSS, SPEC, REDUCE, CPL><<<grid, kThreads, 0, st>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS, int SPEC, bool REDUCE>
int launch_pairwise_inst(const PairwiseArgs& a, cudaStream_t st) {
  if constexpr (pairwise_cpl2_pays<LOSS>()) {
    if (pairwise_cpl(a.m) == 2 && !a.force_cpl1)
      return launch_pairwise_cpl<LOSS, SPEC, REDUCE, 2>(a, st);
  }
  return launch_pairwise_cpl<LOSS, SPEC, REDUCE, 1>(a, st);
}

// compile-time specialisation for the shipped configurations of the three headline
// distances (fun in {none, log1p}, flag = default true, tau on/off), as the loss kernel
template <int LOSS, bool REDUCE>
int launch_pairwise_spec(const PairwiseArgs& a, cudaStream_t st) {
  constexpr bool kHasSpec = (LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd);
  if constexpr (kHasSpec) {
    const gd::PairParams<float>& pp = a.pp;
    if (pp.flag == 1 && (pp.fun == gd::kFunNone || pp.fun == gd::kFunLog1p)) {
      switch (pp.fun | (pp.tau_on << 2) | (1 << 3)) {
        case 8: return launch_pairwise_inst<LOSS, 8, REDUCE>(a, st);
        case 9: return launch_pairwise_inst<LOSS, 9, REDUCE>(a, st);
        case 12: return launch_pairwise_inst<LOSS, 12, REDUCE>(a, st);
        case 13: return launch_pairwise_inst<LOSS, 13, REDUCE>(a, st);
        default: break;
      }
    }
  }
  return launch_pairwise_inst<LOSS, -1, REDUCE>(a, st);
}

template <int LOSS>
int launch_pairwise(const PairwiseArgs& a, cudaStream_t st) {
  return a.row_min ? launch_pairwise_spec<LOSS, true>(a, st)
                   : launch_pairwise_spec<LOSS, false>(a, st);
}
#endif  // !GD_HOST_EMULATION

}  // namespace gdk
