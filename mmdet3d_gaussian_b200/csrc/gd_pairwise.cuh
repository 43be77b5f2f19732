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
// memory.  Lanes map to columns (boxes2, converted once into registers); the 8 warps
// are arranged wx x wy with wx = number of 32-column groups in use (1, 2, 4 or 8) and
// wy = 8 / wx row phases, so small M (a handful of ground-truth boxes) still fills the
// CTA.  The inner loop reads the row Gaussian as a shared-memory broadcast and
// evaluates the branch-free FAST cores (robust cores on a cold branch); sin/cos of the
// yaw difference come from the angle-difference identities.  Stores are coalesced
// 4 B/pair streaming stores.  FP32 CUDA-core math, no tensor cores (not a contraction).
//
// Reductions.  Values map to order-preserving 32-bit keys (NaN lowest, as torch.min
// propagates NaN).  Row minimum: one REDUX.MIN + one ballot per warp and row, lowest
// column wins ties.  Column minimum: a compare/select per pair in the owning lane,
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

// SPEC < 0: fun / tau_on / flag at run time; SPEC >= 0: bits [1:0] fun, [2] tau_on,
// [3] flag compile-time (as gd_warp_kernel).
template <int LOSS, int SPEC, bool REDUCE>
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
  const int wx = a.m <= 32 ? 1 : (a.m <= 64 ? 2 : (a.m <= 128 ? 4 : 8));
  const int wy = kWarps / wx;
  const int cgrp = warp % wx, ry = warp / wx;
  const long long chunk = 32LL * wx;
  const bool want_col = REDUCE && a.col_keys != nullptr;
  const bool one_chunk = a.m <= chunk;         // column minima can stay in registers across tiles
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  unsigned int cbest = 0xffffffffu, crow = 0u;

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
      if (c0 + 32LL * cgrp >= a.m) continue;   // this warp's 32 columns are all past the end
      const long long j = c0 + 32LL * cgrp + lane;
      const bool live = j < a.m;
      gd::BoxGauss<float> t;
      if (live) t = gd::box_gauss(a.b2 + j * 7, pp);
      else t = s_rows[0];                      // any valid box: the result is discarded
      if (want_col && !one_chunk) cbest = 0xffffffffu;
#pragma unroll 2
      for (int r = ry; r < rows; r += wy) {
        const float v = gd::pair_value_auto<float, LOSS>(s_rows[r], t, pp);
        if (a.out != nullptr && live)
          __stcs(a.out + (row0 + r) * a.out_stride + j, a.similarity ? 1.0f - v : v);
        if (REDUCE) {
          const unsigned int key = live ? order_key(v) : 0xffffffffu;
          const unsigned int mn = __reduce_min_sync(0xffffffffu, key);
          const unsigned int who = __ballot_sync(0xffffffffu, key == mn);
          if (lane == 0 && mn != 0xffffffffu) {
            const unsigned long long k64 =
                ((unsigned long long)mn << 32) |
                (unsigned int)(c0 + 32LL * cgrp + (__ffs(who) - 1));
            if (k64 < s_best[r][warp]) s_best[r][warp] = k64;
          }
          if (want_col && key < cbest) {       // rows ascend within a lane: first minimum kept
            cbest = key;
            crow = (unsigned int)(row0 + r);
          }
        }
      }
      if (want_col && !one_chunk && live && cbest != 0xffffffffu)
        atomicMax(a.col_keys + j, ~(((unsigned long long)cbest << 32) | crow));
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
      const long long j = 32LL * cgrp + lane;
      if (j < a.m && cbest != 0xffffffffu)
        atomicMax(a.col_keys + j, ~(((unsigned long long)cbest << 32) | crow));
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
template <int LOSS, int SPEC, bool REDUCE>
int launch_pairwise_inst(const PairwiseArgs& a, cudaStream_t st) {
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  if (ntiles > 2147483647LL || a.m > 0x7fffffffLL || a.n > 0xffffffffLL) return GD_ERR_BAD_ARG;
  const int wx = a.m <= 32 ? 1 : (a.m <= 64 ? 2 : (a.m <= 128 ? 4 : 8));
  dim3 grid;
  if (REDUCE) {                              // every CTA walks all columns of its rows
    long long gx = ntiles;
    if (a.col_keys) {                        // persistent: bounds the column atomics
      const long long cap = (long long)device_info().sm_count * 6;
      if (gx > cap) gx = cap;
    }
    grid = dim3((unsigned)gx, 1);
  } else {
    long long gy = (a.m + 32LL * wx - 1) / (32LL * wx);
    if (gy > 65535) gy = 65535;
    grid = dim3((unsigned)ntiles, (unsigned)gy);
  }
  gd_pairwise_kernel<LOSS, SPEC, REDUCE><<<grid, kThreads, 0, st>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
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

// ---------------------------------------------------------------------------
// Packed-FP32 variant (OPT-IN: GD_PAIR_PACKED / GD_B200_PAIRWISE_PACKED=1).  The scalar
// kernel above is issue bound (profiles/r01d_pairwise_ncu.md: 89 % of the issue slots, 48 %
// of the FMA pipe, ~105 instructions per pair in the matrix loop and ~50 more per pair for
// the fused reductions).  Three changes, same results contract:
//   * TWO rows per pass.  The row tile lives in shared memory as float2 {row 2k, row 2k+1}
//     per BoxGauss field; one 128-bit broadcast load yields two packed operands and the
//     FAST cores run as FFMA2 / FMUL2 / FADD2 (gd_packed.cuh) with the column box as the
//     scalar-broadcast operand: half the issue slots for the FP32 part.
//   * CPL columns per lane (1, 2, 4 or 8; column q of a lane is base + 32 q + lane, so the
//     stores of one q stay coalesced).  A warp covers 32 CPL columns of a row pair, every
//     shared-memory load serves 2 CPL pairs, and the row reduction is done in the lane first:
//     ONE pair of REDUX.MIN per row and warp instead of one REDUX + ballot per 32 pairs.
//   * All 8 warps of the CTA take different row pairs (pair p -> warp p mod 8): a row is
//     owned by one warp, so its running minimum needs no cross-warp merge.
// Reductions keep the contract of the scalar kernel: order-preserving keys, NaN lowest,
// ties -> lowest column / lowest row, matrix and minima from one instruction sequence.
// Values may differ from the scalar kernel in the last bit (FMA contraction).
// Only gwd3d / kld3d / bd3d with a compile-time SPEC; anything else runs the scalar kernel.
// ---------------------------------------------------------------------------
constexpr int kPairsPerCta = kRowsPerCta / 2;
constexpr int kPairStride = (gd::kGaussFields + 1) & ~1;   // float2 per pair, padded: 16-B rows

template <int LOSS, int SPEC, bool REDUCE, int CPL>
__global__ void __launch_bounds__(kThreads) gd_pairwise_packed_kernel(const PairwiseArgs a) {
  static_assert(SPEC >= 0, "packed pairwise kernels are compile-time specialised");
  __shared__ __align__(16) float2 s_pair[kPairsPerCta][kPairStride];
  __shared__ unsigned char s_nice[kRowsPerCta];
  __shared__ unsigned long long s_best[REDUCE ? kRowsPerCta : 1];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gd::PairParams<float> pp = a.pp;
  pp.fun = SPEC & 3;
  pp.tau_on = (SPEC >> 2) & 1;
  pp.flag = (SPEC >> 3) & 1;
  const gd::PairParams<gd::f2> pp2 = gd::broadcast_params(pp);
  constexpr long long kChunk = 32LL * CPL;     // columns one warp covers per pass
  const bool want_col = REDUCE && a.col_keys != nullptr;
  const bool one_chunk = a.m <= kChunk;        // column minima can stay in registers across tiles
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  unsigned int cbest[CPL], crow[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    cbest[q] = 0xffffffffu;
    crow[q] = 0u;
  }

  // scalar BoxGauss of one row of the tile, rebuilt from the packed tile (cold path only)
  auto row_gauss = [&](int r) {
    float f[gd::kGaussFields];
    const float* src = reinterpret_cast<const float*>(&s_pair[r >> 1][0]) + (r & 1);
#pragma unroll
    for (int k = 0; k < gd::kGaussFields; ++k) f[k] = src[2 * k];
    gd::BoxGauss<float> b = gd::gauss_from_fields<float>(f);
    b.nice = s_nice[r];
    return b;
  };
  auto flush_cols = [&](long long c0) {        // this lane's column minima -> global keys
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const long long j = c0 + 32LL * q + lane;
      if (j < a.m && cbest[q] != 0xffffffffu)
        atomicMax(a.col_keys + j, ~(((unsigned long long)cbest[q] << 32) | crow[q]));
    }
  };

  // this lane's CPL column boxes (scalar: they enter the packed math as broadcast operands)
  gd::BoxGauss<float> t[CPL];
  bool live[CPL];
  auto load_cols = [&](long long c0) {
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const long long j = c0 + 32LL * q + lane;
      live[q] = j < a.m;
      // a dead lane evaluates the last column again; its results are never used
      t[q] = gd::box_gauss(a.b2 + (live[q] ? j : a.m - 1) * 7, pp);
    }
    if (want_col && !one_chunk) {
#pragma unroll
      for (int q = 0; q < CPL; ++q) cbest[q] = 0xffffffffu;
    }
  };
  // rows of one tile -> packed shared-memory tile (+ the tile's running row minima)
  auto convert_rows = [&](long long row0, int rows) {
    __syncthreads();                           // previous tile fully consumed
    if (tid < rows) {
      const gd::BoxGauss<float> b = gd::box_gauss(a.b1 + (row0 + tid) * 7, pp);
      float f[gd::kGaussFields];
      gd::gauss_to_fields(b, f);
      float* dst = reinterpret_cast<float*>(&s_pair[tid >> 1][0]) + (tid & 1);
#pragma unroll
      for (int k = 0; k < gd::kGaussFields; ++k) dst[2 * k] = f[k];
      s_nice[tid] = (unsigned char)b.nice;
      if (tid == rows - 1 && (rows & 1)) {     // odd tile: the last pair's upper half is a copy
#pragma unroll
        for (int k = 0; k < gd::kGaussFields; ++k) dst[2 * k + 1] = f[k];
        s_nice[rows] = (unsigned char)b.nice;
      }
    }
    if (REDUCE && tid < kRowsPerCta) s_best[tid] = ~0ull;
    __syncthreads();
  };
  auto finish_rows = [&](long long row0, int rows) {
    if (REDUCE) {
      __syncthreads();
      if (tid < rows) {
        const unsigned long long k = s_best[tid];
        a.row_min[row0 + tid] = key_value((unsigned int)(k >> 32));
        a.row_argmin[row0 + tid] = (int)(unsigned int)(k & 0xffffffffu);
      }
    }
  };
  // all row pairs of the tile against this lane's columns of the chunk at c0
  auto process = [&](long long row0, int rows, long long c0) {
    const int npairs = (rows + 1) >> 1;
    float* optr = a.out != nullptr ? a.out + (row0 + 2 * warp) * a.out_stride + c0 + lane : nullptr;
    const long long pair_step = 2LL * kWarps * a.out_stride;
    for (int pr = warp; pr < npairs; pr += kWarps) {
      gd::f2 f2v[kPairStride];
      const float4* src4 = reinterpret_cast<const float4*>(&s_pair[pr][0]);
#pragma unroll
      for (int k = 0; k < kPairStride / 2; ++k) {
        const float4 v = src4[k];              // 128-bit broadcast load: two fields x {row 2 pr, 2 pr + 1}
        f2v[2 * k] = gd::mk2(v.x, v.y);
        f2v[2 * k + 1] = gd::mk2(v.z, v.w);
      }
      const gd::BoxGauss<gd::f2> p2 = gd::gauss_from_fields<gd::f2>(f2v);
      const bool nice0 = s_nice[2 * pr] != 0, nice1 = s_nice[2 * pr + 1] != 0;
      const bool has_hi = 2 * pr + 1 < rows;    // warp-uniform
      // in-lane minima over this lane's columns, per row of the pair: (key, q)
      unsigned int k0 = 0xffffffffu, k1 = 0xffffffffu, q0 = 0u, q1 = 0u;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        gd::BoxGauss<gd::f2> t2;
        {
          float f[gd::kGaussFields];
          gd::f2 fq[gd::kGaussFields];
          gd::gauss_to_fields(t[q], f);
#pragma unroll
          for (int k = 0; k < gd::kGaussFields; ++k) fq[k] = gd::mk2(f[k], f[k]);
          t2 = gd::gauss_from_fields<gd::f2>(fq);
        }
        gd::m2 rare{!(nice0 && t[q].nice), !(nice1 && t[q].nice)};
        const gd::f2 v2 = gd::pair_value_fast2<LOSS>(p2, t2, pp2, &rare);
        float v0 = gd::lo2(v2), v1 = gd::hi2(v2);
        if (rare.lo || rare.hi) {               // cold: one branch on the common path
          if (rare.lo) v0 = gd::pair_value<float, LOSS>(row_gauss(2 * pr), t[q], pp);
          if (rare.hi) v1 = gd::pair_value<float, LOSS>(row_gauss(2 * pr + 1), t[q], pp);
        }
        if (optr != nullptr && live[q]) {
          __stcs(optr + 32 * q, a.similarity ? 1.0f - v0 : v0);
          if (has_hi) __stcs(optr + a.out_stride + 32 * q, a.similarity ? 1.0f - v1 : v1);
        }
        if (REDUCE) {
          const unsigned int key0 = live[q] ? order_key(v0) : 0xffffffffu;
          const unsigned int key1 = (live[q] && has_hi) ? order_key(v1) : 0xffffffffu;
          if (key0 < k0) {                     // strict: the lowest q (lowest column) keeps ties
            k0 = key0;
            q0 = (unsigned int)q;
          }
          if (key1 < k1) {
            k1 = key1;
            q1 = (unsigned int)q;
          }
          if (want_col) {                      // rows ascend within a lane: first minimum kept
            if (key0 < cbest[q]) {
              cbest[q] = key0;
              crow[q] = (unsigned int)(row0 + 2 * pr);
            }
            if (key1 < cbest[q]) {
              cbest[q] = key1;
              crow[q] = (unsigned int)(row0 + 2 * pr + 1);
            }
          }
        }
      }
      if (optr != nullptr) optr += pair_step;
      if (REDUCE) {
        // row minimum over the warp's 32 CPL columns: min key, then min column among the
        // lanes that hold it; this warp owns the row, so the running best needs no atomics
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const unsigned int key = h ? k1 : k0;
          const unsigned int col = (unsigned int)(c0 + 32LL * (h ? q1 : q0) + lane);
          const unsigned int mn = __reduce_min_sync(0xffffffffu, key);
          const unsigned int cmin = __reduce_min_sync(0xffffffffu, key == mn ? col : 0xffffffffu);
          if (lane == 0 && mn != 0xffffffffu) {
            const unsigned long long k64 = ((unsigned long long)mn << 32) | cmin;
            if (k64 < s_best[2 * pr + h]) s_best[2 * pr + h] = k64;
          }
        }
      }
    }
  };

  // Loop order.  Column boxes are converted (sincos, reciprocals, ...) once per chunk and
  // CTA and reused for every tile the CTA walks -- possible whenever a row's minimum does not
  // have to be carried across chunks: matrix-only launches (each CTA owns one chunk column,
  // blockIdx.y) and reductions whose columns fit one chunk (m <= 32 CPL; up to 256 GT boxes).
  // Reductions over several chunks keep the tile-outer order so that s_best stays per tile.
  if (!REDUCE || one_chunk) {
    for (long long c0 = (long long)blockIdx.y * kChunk; c0 < a.m; c0 += (long long)gridDim.y * kChunk) {
      load_cols(c0);
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * kRowsPerCta;
        const int rows = (int)min((long long)kRowsPerCta, a.n - row0);
        convert_rows(row0, rows);
        process(row0, rows, c0);
        finish_rows(row0, rows);
      }
    }
  } else {
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long row0 = tile * kRowsPerCta;
      const int rows = (int)min((long long)kRowsPerCta, a.n - row0);
      convert_rows(row0, rows);
      for (long long c0 = 0; c0 < a.m; c0 += kChunk) {
        load_cols(c0);
        process(row0, rows, c0);
        if (want_col) flush_cols(c0);
      }
      finish_rows(row0, rows);
    }
  }
  if (want_col) {
    if (one_chunk) flush_cols(0);
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

#if !defined(GD_HOST_EMULATION)
template <int LOSS, int SPEC, bool REDUCE, int CPL>
int launch_pairwise_packed_cpl(const PairwiseArgs& a, cudaStream_t st) {
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  dim3 grid;
  if (REDUCE) {                              // every CTA walks all columns of its rows
    long long gx = ntiles;
    if (a.col_keys) {                        // persistent: bounds the column atomics
      const long long cap = (long long)device_info().sm_count * 6;
      if (gx > cap) gx = cap;
    }
    grid = dim3((unsigned)gx, 1);
  } else {                                   // one chunk column per blockIdx.y, persistent in x
    long long gy = (a.m + 32LL * CPL - 1) / (32LL * CPL);
    if (gy > 65535) gy = 65535;
    long long gx = ((long long)device_info().sm_count * 8 + gy - 1) / gy;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1) gx = 1;
    grid = dim3((unsigned)gx, (unsigned)gy);
  }
  gd_pairwise_packed_kernel<LOSS, SPEC, REDUCE, CPL><<<grid, kThreads, 0, st>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS, int SPEC, bool REDUCE>
int launch_pairwise_packed_inst(const PairwiseArgs& a, cudaStream_t st) {
  const long long ntiles = (a.n + kRowsPerCta - 1) / kRowsPerCta;
  if (ntiles > 2147483647LL || a.m > 0x7fffffffLL || a.n > 0xffffffffLL) return GD_ERR_BAD_ARG;
  if (a.m <= 32) return launch_pairwise_packed_cpl<LOSS, SPEC, REDUCE, 1>(a, st);
  if (a.m <= 64) return launch_pairwise_packed_cpl<LOSS, SPEC, REDUCE, 2>(a, st);
  if (a.m <= 128) return launch_pairwise_packed_cpl<LOSS, SPEC, REDUCE, 4>(a, st);
  return launch_pairwise_packed_cpl<LOSS, SPEC, REDUCE, 8>(a, st);
}

// returns kNoPackedKernel when the configuration has no packed instantiation; the caller
// then launches the scalar kernel
constexpr int kNoPackedKernel = -1000;

template <int LOSS>
int launch_pairwise_packed(const PairwiseArgs& a, cudaStream_t st) {
  static_assert(LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd, "gwd3d, kld3d, bd3d");
  const gd::PairParams<float>& pp = a.pp;
  if (!(pp.flag == 1 && (pp.fun == gd::kFunNone || pp.fun == gd::kFunLog1p))) return kNoPackedKernel;
#define GD_PACKED_CASE(S)                                                         \
  case S:                                                                         \
    return a.row_min ? launch_pairwise_packed_inst<LOSS, S, true>(a, st)          \
                     : launch_pairwise_packed_inst<LOSS, S, false>(a, st);
  switch (pp.fun | (pp.tau_on << 2) | (1 << 3)) {
    GD_PACKED_CASE(8) GD_PACKED_CASE(9) GD_PACKED_CASE(12) GD_PACKED_CASE(13)
    default: break;
  }
#undef GD_PACKED_CASE
  return kNoPackedKernel;
}
#endif  // !GD_HOST_EMULATION

}  // namespace gdk
