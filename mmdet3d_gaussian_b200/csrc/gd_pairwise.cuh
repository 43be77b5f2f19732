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
// The kernels are issue bound, so the design goal is instructions per pair (round 1: ~111
// matrix / ~160 fused; now 61 / 68: profiles/r03_pairwise.md).
//
// Kernels in this header:
//   gd_pairwise_kernel / gd_pairwise_kernel_m3 (pairwise_body)  the matrix, lanes on columns;
//       with REDUCE also the row / column minima of the same launch (column-lane reductions)
//   gd_pairwise_filter_kernel   the matrix loop with "append candidates below a bound" in place
//       of the store (column top-k of gd_simota.cu)
//   gd_pairwise_rowlane_kernel  row / column minima without the matrix, lanes on ROWS
// Every one of them evaluates a pair through gd::pair_value_auto / gd::pair_value_fast, whose
// FAST cores are written with explicitly rounded operations (gd_math.cuh, namespace pw): the
// value of a pair is the same bits in every kernel, whatever the mapping.
//
// Reductions of the column-lane kernel.  Values map to order-preserving 32-bit keys (NaN lowest, as torch.min
// propagates NaN).  Row minimum: in-lane minimum over the CPL columns, one REDUX.MIN and
// CPL ballots per warp and row, lowest column wins ties.  Column minimum: a compare/select per pair in the owning lane,
// merged across CTAs with one 64-bit atomic per column on (key << 32 | row) -- lowest
// row wins ties -- and unpacked by the last CTA to finish (atomic ticket), which also
// restores the workspace.
#pragma once
#include "gd_common.cuh"
#include "gd_packed.cuh"

namespace gdk {

constexpr int kRowsPerCta = 64;
constexpr int kRowsPerCtaBig = 128;     // persistent matrix / filter launches: half the barriers per row
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
  int idx64;                       // GD_PAIR_INDEX64: row_argmin / col_argmin are int64 arrays
  int tile_rows;                   // rows per tile of the column-lane kernel (0: kRowsPerCta);
                                   // the matrix launcher sizes it so the waves come out even
  float* row_min;                  // [n]   REDUCE only
  int* row_argmin;                 // [n]
  unsigned long long* col_keys;    // [m] workspace holding ~key (so zero is the identity of the
                                   // atomicMax merge): zero on entry and on exit; nullable
  float* col_min;                  // [m]
  int* col_argmin;                 // [m]
  unsigned int* ticket;            // zero on entry and on exit
  gd::PairParams<float> pp;
};

// index outputs: int32 (the C ABI's default) or int64 (torch's index type, GD_PAIR_INDEX64)
__device__ __forceinline__ void store_index(int* base, long long i, int v, int idx64) {
  if (idx64) reinterpret_cast<long long*>(base)[i] = (long long)v;
  else base[i] = v;
}

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
__device__ __forceinline__ void pairwise_body(const PairwiseArgs& a) {
  // the persistent matrix launch of the exact-form distances may ask for tiles up to
  // kRowsPerCtaBig rows (a.tile_rows); every other launch stays at kRowsPerCta
  constexpr int kTileCap = (!REDUCE && gd::PairwiseExact<LOSS>::value) ? kRowsPerCtaBig : kRowsPerCta;
  __shared__ gd::BoxGauss<float> s_rows[kTileCap];
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
  const int tile_rows = a.tile_rows > 0 ? a.tile_rows : kRowsPerCta;       // <= kTileCap
  const long long ntiles = (a.n + tile_rows - 1) / tile_rows;
  // One column chunk for the whole launch (matrix kernel of the exact-form distance): the lane's
  // column Gaussians are converted once and kept across the CTA's tiles.  Compile-time off
  // elsewhere: the longer live ranges cost kld3d / bd3d a resident CTA (-7 %, gpurun_out/r03c).
  constexpr bool kKeepCols = !REDUCE && gd::PairwiseExact<LOSS>::value;
  const bool keep_cols = kKeepCols && gridDim.y == 1 && one_chunk;
  bool cols_ready = false;
  gd::BoxGauss<float> tkeep[kKeepCols ? CPL : 1];
  bool lkeep[kKeepCols ? CPL : 1];
  unsigned int cbest[CPL], crow[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    cbest[q] = 0xffffffffu;
    crow[q] = 0u;
  }

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * tile_rows;
    const int rows = (int)min((long long)tile_rows, a.n - row0);
    __syncthreads();                           // previous tile fully consumed
    int row_nice = 1;
    if (tid < rows) {
      s_rows[tid] = gd::box_gauss(a.b1 + (row0 + tid) * 7, pp);
      row_nice = s_rows[tid].nice;
    }
    if (REDUCE) {
      for (int i = tid; i < kRowsPerCta * kWarps; i += kThreads) (&s_best[0][0])[i] = ~0ull;
    }
    const bool tile_nice = __syncthreads_and(row_nice) != 0;   // also publishes the tile
    for (long long c0 = (long long)blockIdx.y * chunk; c0 < a.m; c0 += (long long)gridDim.y * chunk) {
      const long long jb = c0 + (long long)kWarpCols * cgrp;     // first column of this warp
      if (jb >= a.m) continue;                 // this warp's columns are all past the end
      gd::BoxGauss<float> t[CPL];
      bool live[CPL];
      if (kKeepCols && keep_cols && cols_ready) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          t[q] = tkeep[kKeepCols ? q : 0];
          live[q] = lkeep[kKeepCols ? q : 0];
        }
      } else {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const long long j = jb + 32 * q + lane;
          live[q] = j < a.m;
          if (live[q]) t[q] = gd::box_gauss(a.b2 + j * 7, pp);
          else t[q] = s_rows[0];               // any valid box: the result is discarded
          if (want_col && !one_chunk) cbest[q] = 0xffffffffu;
          if (kKeepCols) {
            tkeep[kKeepCols ? q : 0] = t[q];
            lkeep[kKeepCols ? q : 0] = live[q];
          }
        }
        cols_ready = true;
      }
      // row pointer of this lane's first column, advanced by wy rows per iteration
      float* orow = a.out != nullptr ? a.out + (row0 + ry) * a.out_stride + jb + lane : nullptr;
      const long long ostep = (long long)wy * a.out_stride;
      if constexpr (!REDUCE && gd::PairwiseExact<LOSS>::value) {
        // Matrix only, every column of the warp in range, every box of the tile and of the
        // warp's columns nice (the common case): straight-line FAST cores and unconditional
        // coalesced stores, no per-pair screen, range check or branch.  A guard tripping inside
        // a FAST core (rare) is remembered and the lane's columns are rewritten afterwards
        // through pair_value_auto -- the same thread stored them, so program order decides.
        bool ok = tile_nice;
#pragma unroll
        for (int q = 0; q < CPL; ++q) ok = ok && live[q] && t[q].nice != 0;
        if (__all_sync(0xffffffffu, ok)) {
          // 1 - v and v as one FMA with hoisted constants: identical bits (single rounding)
          const float sgn = a.similarity ? -1.0f : 1.0f, off = a.similarity ? 1.0f : 0.0f;
          bool redo = false;
          float* o = orow;
#pragma unroll 2
          for (int r = ry; r < rows; r += wy) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              bool rare = false;
              const float f = gd::pair_value_fast<float, LOSS>(s_rows[r], t[q], pp, &rare);
              redo |= rare;
              __stcs(o + 32 * q, fmaf(f, sgn, off));
            }
            o += ostep;
          }
          if (redo) {
            o = orow;
            for (int r = ry; r < rows; r += wy) {
#pragma unroll
              for (int q = 0; q < CPL; ++q) {
                const float v = gd::pair_value_auto<float, LOSS>(s_rows[r], t[q], pp);
                __stcs(o + 32 * q, a.similarity ? 1.0f - v : v);
              }
              o += ostep;
            }
          }
          continue;
        }
      }
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
        store_index(a.row_argmin, row0 + tid, (int)(unsigned int)(k & 0xffffffffu), a.idx64);
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
        store_index(a.col_argmin, j, (int)(unsigned int)(k & 0xffffffffu), a.idx64);
        a.col_keys[j] = 0ull;
      }
      if (tid == 0) *a.ticket = 0u;
    }
  }
}

template <int LOSS, int SPEC, bool REDUCE, int CPL>
__global__ void __launch_bounds__(kThreads) gd_pairwise_kernel(const PairwiseArgs a) {
  pairwise_body<LOSS, SPEC, REDUCE, CPL>(a);
}
// the same body capped at 85 registers so that three CTAs stay resident (the persistent matrix
// launch of the exact-form distance, two columns per lane)
template <int LOSS, int SPEC, int CPL>
__global__ void __launch_bounds__(kThreads, 3) gd_pairwise_kernel_m3(const PairwiseArgs a) {
  pairwise_body<LOSS, SPEC, false, CPL>(a);
}

// ---------------------------------------------------------------------------
// Filter pass of the column top-k (gd_simota.cu): the matrix kernel's loop with the store
// replaced by "append (key, row) to column j's candidate buffer when it is <= thr[j]".  One
// column chunk (m <= 32 CPL wx), persistent CTAs, column Gaussians and limits kept in registers;
// straight-line FAST cores when every box in sight is nice, as in the matrix kernel.
// ---------------------------------------------------------------------------
struct FilterArgs {
  const unsigned long long* thr;             // [m] upper bound of the column's k-th (key, row)
  unsigned int* count;                       // [m] candidates appended so far (zero on entry)
  unsigned long long* cand;                  // [m][cap]
  unsigned int cap;
};

template <int LOSS, int SPEC, int CPL>
__global__ void __launch_bounds__(kThreads, 3) gd_pairwise_filter_kernel(const PairwiseArgs a,
                                                                         const FilterArgs f) {
  __shared__ gd::BoxGauss<float> s_rows[kRowsPerCtaBig];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gd::PairParams<float> pp = a.pp;
  if (SPEC >= 0) {
    pp.fun = SPEC & 3;
    pp.tau_on = (SPEC >> 2) & 1;
    pp.flag = (SPEC >> 3) & 1;
  }
  pp.lean = 1;
  constexpr int kWarpCols = 32 * CPL;
  const int wx = pairwise_wx(a.m, kWarpCols);
  const int wy = kWarps / wx;
  const int cgrp = warp % wx, ry = warp / wx;
  const long long jb = (long long)kWarpCols * cgrp;            // first column of this warp
  const bool warp_live = jb < a.m;
  const int tile_rows = a.tile_rows > 0 ? a.tile_rows : kRowsPerCta;
  const long long ntiles = (a.n + tile_rows - 1) / tile_rows;
  gd::BoxGauss<float> t[CPL];
  bool live[CPL];
  unsigned long long lim[CPL];
  bool cols_ok = warp_live;
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const long long j = jb + 32 * q + lane;
    live[q] = j < a.m;
    t[q] = gd::box_gauss(a.b2 + (live[q] ? j : 0) * 7, pp);     // dead lane: any valid box
    lim[q] = live[q] ? f.thr[j] : 0ull;                        // dead lane: nothing passes
    cols_ok = cols_ok && live[q] && t[q].nice != 0;
  }
  const bool fast_cols = gd::PairwiseExact<LOSS>::value && __all_sync(0xffffffffu, cols_ok);
  auto offer = [&](int q, long long row, float v) {
    const unsigned long long k64 = ((unsigned long long)order_key(v) << 32) | (unsigned int)row;
    if (k64 <= lim[q]) {
      const long long j = jb + 32 * q + lane;
      const unsigned int pos = atomicAdd(f.count + j, 1u);
      if (pos < f.cap) f.cand[j * f.cap + pos] = k64;
    }
  };
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * tile_rows;
    const int rows = (int)min((long long)tile_rows, a.n - row0);
    __syncthreads();                           // previous tile fully consumed
    int row_nice = 1;
    if (tid < rows) {
      s_rows[tid] = gd::box_gauss(a.b1 + (row0 + tid) * 7, pp);
      row_nice = s_rows[tid].nice;
    }
    const bool tile_nice = __syncthreads_and(row_nice) != 0;   // also publishes the tile
    if (!warp_live) continue;
    if constexpr (gd::PairwiseExact<LOSS>::value) {
      if (fast_cols && tile_nice) {
        bool redo = false;
#pragma unroll 2
        for (int r = ry; r < rows; r += wy) {
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            bool rare = false;
            const float v = gd::pair_value_fast<float, LOSS>(s_rows[r], t[q], pp, &rare);
            redo |= rare;
            if (!rare) offer(q, row0 + r, v);
          }
        }
        if (redo) {                            // only the pairs whose FAST core gave up
          for (int r = ry; r < rows; r += wy) {
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
              bool rare = false;
              (void)gd::pair_value_fast<float, LOSS>(s_rows[r], t[q], pp, &rare);
              if (rare) offer(q, row0 + r, gd::pair_value_auto<float, LOSS>(s_rows[r], t[q], pp));
            }
          }
        }
        continue;
      }
    }
    for (int r = ry; r < rows; r += wy) {
#pragma unroll
      for (int q = 0; q < CPL; ++q)
        if (live[q]) offer(q, row0 + r, gd::pair_value_auto<float, LOSS>(s_rows[r], t[q], pp));
    }
  }
}

// ---------------------------------------------------------------------------
// Fused reductions WITHOUT the matrix (row f2, the assigner's launch): lanes map to ROWS.
//
// With lanes on columns (the kernel above, needed for coalesced matrix stores) a row minimum
// costs a REDUX, ballots and a shared-memory update per warp and row, and every value goes
// through the key mapping -- the fused launch ran 35-40 % slower than writing the matrix it
// avoids (profiles/r02_pairwise.md).  Here a lane keeps RPL row Gaussians in registers and the
// column Gaussians (m <= kRowLaneCols) sit in shared memory, read as broadcasts:
//   * row minimum: NaN-sticky FMNMX + compare/select per pair, in the lane, no collective;
//     a NaN result (rare) is resolved on a cold pass (first NaN column, as torch.min);
//   * column minimum: two integer ops for the order key, one REDUX.MIN per warp and column for
//     the key and one for the lowest row holding it, merged into a WARP-PRIVATE (key, row) table
//     in shared memory by a compare and a predicated store -- no branch, no atomic, the same
//     cost whatever the history (a "did it improve?" pre-check does not pay: at C4 a warp sees
//     one or two units, so most columns improve every time); the 8 tables are merged once per
//     CTA at the end;
//   * when every box of the warp's rows and every column box is "nice" (the common case) the
//     loop carries no per-pair screen at all.
// Work is handed out per WARP in units of 32 RPL rows: the first unit of every warp statically
// (CTA-interleaved), further ones through a counter when the caller gave a workspace
// (ticket[1]), so the 4 schedulers of an SM stay evenly loaded at sizes that give each only a
// handful of units (C4: 3125 units over 592 schedulers).
// Values come out of the same gd::core_eval instruction sequence as the matrix kernel's.
// ---------------------------------------------------------------------------
constexpr int kRowLaneCols = 256;
struct alignas(16) ColBox {
  gd::BoxGauss<float> g;
};
struct CleanLoop { static constexpr bool value = true; };
struct GeneralLoop { static constexpr bool value = false; };

__device__ __forceinline__ float min_nan(float a, float b) {     // NaN propagates, and sticks
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  return (a != a || b != b) ? __uint_as_float(0x7fc00000u) : (b < a ? b : a);
#endif
}
// order-preserving float -> uint32 for non-NaN values (NaNs are handled on the cold pass)
__device__ __forceinline__ unsigned int order_key_fast(float v) {
  const unsigned int b = __float_as_uint(v);
  return b ^ ((unsigned int)((int)b >> 31) | 0x80000000u);
}

template <int LOSS, int SPEC, int RPL>
__global__ void __launch_bounds__(kThreads) gd_pairwise_rowlane_kernel(const PairwiseArgs a) {
  __shared__ ColBox s_cols[kRowLaneCols];
  __shared__ unsigned long long s_wbest[kWarps][kRowLaneCols];   // per warp: key << 32 | row; ~0: none
  __shared__ unsigned long long s_dummy[kThreads];               // lanes 1..31: sink of the table update
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gd::PairParams<float> pp = a.pp;
  if (SPEC >= 0) {
    pp.fun = SPEC & 3;
    pp.tau_on = (SPEC >> 2) & 1;
    pp.flag = (SPEC >> 3) & 1;
  }
  pp.lean = 1;
  const bool want_col = a.col_keys != nullptr;
  const int cnt = (int)a.m;                                    // <= kRowLaneCols (launcher)
  constexpr int kUnitRows = 32 * RPL;
  const long long nunits = (a.n + kUnitRows - 1) / kUnitRows;

  int nice = 1;
  for (int j = tid; j < cnt; j += kThreads) {
    s_cols[j].g = gd::box_gauss(a.b2 + (long long)j * 7, pp);
    nice &= s_cols[j].g.nice;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s_wbest[w][j] = ~0ull;
  }
  s_dummy[tid] = 0ull;                                         // nothing is ever below it: never written
  const bool cols_nice = __syncthreads_and(nice) != 0;         // also publishes the columns

  // unit schedule: the first unit of a warp is static and CTA-interleaved (consecutive units
  // go to different CTAs, then to the next warp slot); further units come from the counter when
  // there is one (else the static stride continues)
  const bool dynamic = a.ticket != nullptr;
  const long long gwarps = (long long)gridDim.x * kWarps;
  long long unit = (long long)blockIdx.x + (long long)gridDim.x * warp;
  unsigned long long* wbest = s_wbest[warp];
  unsigned long long* const slot0 = lane == 0 ? wbest : &s_dummy[tid];
  const int slot_step = lane == 0 ? 1 : 0;
  for (bool first = true;; first = false) {
    if (!first) {
      if (dynamic) {
        unsigned int u = 0;
        if (lane == 0) u = atomicAdd(a.ticket + 1, 1u);
        unit = gwarps + (long long)__shfl_sync(0xffffffffu, u, 0);
      } else {
        unit += gwarps;
      }
    }
    if (unit >= nunits) break;
    const long long row0 = unit * kUnitRows;
    gd::BoxGauss<float> rb[RPL];
    bool live[RPL];
    float best[RPL];
    int bj[RPL];
    bool clean = cols_nice;
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const long long r = row0 + 32 * q + lane;
      live[q] = r < a.n;
      rb[q] = gd::box_gauss(a.b1 + (live[q] ? r : a.n - 1) * 7, pp);   // dead lane: any valid box
      clean = clean && live[q] && rb[q].nice;
      best[q] = __uint_as_float(0x7f800000u);
      bj[q] = 0;
    }
    const bool warp_clean = gd::PairwiseExact<LOSS>::value && __all_sync(0xffffffffu, clean);
    const unsigned int rowbase = (unsigned int)row0;
    unsigned int rowid[RPL];
#pragma unroll
    for (int q = 0; q < RPL; ++q) rowid[q] = rowbase + 32u * q + (unsigned int)lane;

    bool redo = false;                           // CLEAN sweep: a guard of the FAST core tripped
    auto sweep = [&](auto clean_tag, auto col_tag) {
      constexpr bool CLEAN = decltype(clean_tag)::value;
      constexpr bool COLS = decltype(col_tag)::value;
#pragma unroll 2
      for (int j = 0; j < cnt; ++j) {
        const gd::BoxGauss<float>& t = s_cols[j].g;
        float v[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          if constexpr (CLEAN && gd::PairwiseExact<LOSS>::value) {
            // straight-line FAST core; a tripped guard turns the value into +inf (it cannot
            // lower any minimum) and the whole unit is redone below on the general path
            bool rare = false;
            const float f = gd::pair_value_fast<float, LOSS>(rb[q], t, pp, &rare);
            v[q] = rare ? __uint_as_float(0x7f800000u) : f;
            redo |= rare;
          } else {
            v[q] = gd::pair_value_auto<float, LOSS>(rb[q], t, pp);
          }
          const float old = best[q];
          best[q] = min_nan(old, v[q]);
          bj[q] = v[q] < old ? j : bj[q];        // strict: the lowest column keeps a tie
        }
        if constexpr (COLS) {
          unsigned int key[RPL];
#pragma unroll
          for (int q = 0; q < RPL; ++q) {
            key[q] = order_key_fast(v[q]);
            if (!CLEAN) key[q] = live[q] ? key[q] : 0xffffffffu;
          }
          unsigned int kmin = key[0];
#pragma unroll
          for (int q = 1; q < RPL; ++q) kmin = key[q] < kmin ? key[q] : kmin;
          const unsigned int mn = __reduce_min_sync(0xffffffffu, kmin);
          // lowest row holding the minimum (rowid[q] = this lane's rows, hoisted)
          unsigned int rr = 0xffffffffu;
#pragma unroll
          for (int q = RPL - 1; q >= 0; --q) rr = key[q] == mn ? rowid[q] : rr;
          const unsigned int rlow = __reduce_min_sync(0xffffffffu, rr);
          // warp-private table, touched by lane 0 only inside the sweep (no atomic, nothing
          // shared between lanes); the other lanes run the same predicated store on a private
          // dummy word, so the warp never diverges.  mn == ~0 (dead lanes only, or a NaN the cold pass overrides
          // with key 0) never gets in.
          const unsigned long long cand = ((unsigned long long)mn << 32) | rlow;
          unsigned long long* slot = slot0 + (long long)j * slot_step;   // lane 0: wbest[j]
          if (CLEAN) {
            if (cand < *slot) *slot = cand;
          } else {
            if (cand < *slot && mn != 0xffffffffu) *slot = cand;
          }
        }
      }
    };
    auto sweep_cols = [&](auto clean_tag) {
      if (want_col) sweep(clean_tag, CleanLoop{});
      else sweep(clean_tag, GeneralLoop{});
    };
    if (warp_clean) {
      sweep_cols(CleanLoop{});
      if (__any_sync(0xffffffffu, redo)) {       // start over: every pair through its own screen
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          best[q] = __uint_as_float(0x7f800000u);
          bj[q] = 0;
        }
        sweep_cols(GeneralLoop{});
      }
    } else {
      sweep_cols(GeneralLoop{});
    }

    // cold pass: a NaN in the row.  torch.min returns NaN with the first NaN column; every
    // column holding a NaN gets the lowest key (0) with the lowest such row.
    __syncwarp();                                // lane 0's table updates before other lanes' atomics
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      if (live[q] && best[q] != best[q]) {
        bool first = true;
        for (int j = 0; j < cnt; ++j) {
          const float v = gd::pair_value_auto<float, LOSS>(rb[q], s_cols[j].g, pp);
          if (v != v) {
            if (first) bj[q] = j;
            first = false;
            if (want_col) atomicMin(&wbest[j], (unsigned long long)(rowbase + 32u * q + lane));
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const long long r = row0 + 32 * q + lane;
      if (live[q]) {
        a.row_min[r] = best[q];
        store_index(a.row_argmin, r, bj[q], a.idx64);
      }
    }
    __syncwarp();                                // atomics of the cold pass before lane 0 goes on
  }

  if (want_col) {
    __syncthreads();
    for (int j = tid; j < cnt; j += kThreads) {
      unsigned long long k = s_wbest[0][j];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) k = s_wbest[w][j] < k ? s_wbest[w][j] : k;
      if (k != ~0ull) atomicMax(a.col_keys + j, ~k);
    }
    // last CTA to finish unpacks the column keys and restores the workspace
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (long long j = tid; j < a.m; j += kThreads) {
        const unsigned long long k = ~__ldcg(a.col_keys + j);
        a.col_min[j] = key_value((unsigned int)(k >> 32));
        store_index(a.col_argmin, j, (int)(unsigned int)(k & 0xffffffffu), a.idx64);
        a.col_keys[j] = 0ull;
      }
      if (tid == 0) {
        a.ticket[1] = 0u;
        *a.ticket = 0u;
      }
    }
  } else if (dynamic) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
        a.ticket[1] = 0u;
        __threadfence();
        *a.ticket = 0u;
      }
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
    if constexpr (gd::PairwiseExact<LOSS>::value) if (gy == 1) {
      // One column chunk: persistent CTAs (the lane's column Gaussians are converted once per
      // CTA, not once per tile) and a tile height chosen so that every CTA slot of the GPU gets
      // the same number of equal tiles -- 200k rows in 64-row tiles are 7.04 waves of 444 CTAs,
      // i.e. 8 wave times; in 57-row tiles they are 7.9 waves of shorter tiles.
      static int occ[kMaxDevices] = {};
      const int dev = current_device();
      if (occ[dev] == 0) {
        int per_sm = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, gd_pairwise_kernel_m3<LOSS, SPEC, CPL>, kThreads, 0);
        if (e != cudaSuccess) return (int)e;
        occ[dev] = per_sm > 0 ? per_sm : 1;
      }
      const long long slots = (long long)device_info().sm_count * occ[dev];
      const long long big_tiles = (a.n + kRowsPerCtaBig - 1) / kRowsPerCtaBig;
      const long long waves = (big_tiles + slots - 1) / slots;
      long long rows = (a.n + waves * slots - 1) / (waves * slots);     // <= kRowsPerCtaBig
      if (rows < 16) rows = 16;
      if (rows > kRowsPerCtaBig) rows = kRowsPerCtaBig;
      PairwiseArgs b = a;
      b.tile_rows = (int)rows;
      const long long nt = (a.n + rows - 1) / rows;
      grid = dim3((unsigned)(nt < slots ? nt : slots), 1);
      gd_pairwise_kernel_m3<LOSS, SPEC, CPL><<<grid, kThreads, 0, st>>>(b);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      return (int)cudaGetLastError();
    }
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

// GD_B200_PAIR_ROWLANE=0 keeps the column-lane kernel for the fused reductions (measurements)
inline bool pairwise_rowlane_enabled() {
  static const bool on = [] {
    const char* e = getenv("GD_B200_PAIR_ROWLANE");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

template <int LOSS, int SPEC>
int launch_rowlane_inst(const PairwiseArgs& a, cudaStream_t st) {
  // two rows per lane halve the shared-memory reads per pair; bd3d / the symmetric KLDs hold
  // too many per-box terms in registers for that
  constexpr int RPL = pairwise_cpl2_pays<LOSS>() ? 2 : 1;
  auto kern = gd_pairwise_rowlane_kernel<LOSS, SPEC, RPL>;
  static int occ[kMaxDevices] = {};
  const int dev = current_device();
  if (occ[dev] == 0) {
    int per_sm = 0;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
    if (e != cudaSuccess) return (int)e;
    occ[dev] = per_sm > 0 ? per_sm : 1;
  }
  const long long nunits = (a.n + 32 * RPL - 1) / (32 * RPL);
  if (nunits > 0xffffffffLL || a.n > 0xffffffffLL) return GD_ERR_BAD_ARG;
  long long grid = (nunits + kWarps - 1) / kWarps;
  const long long cap = (long long)device_info().sm_count * occ[dev];
  if (grid > cap) grid = cap;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS>
int launch_rowlane(const PairwiseArgs& a, cudaStream_t st) {
  constexpr bool kHasSpec = (LOSS == gd::kGwd || LOSS == gd::kKld || LOSS == gd::kBd);
  if constexpr (kHasSpec) {
    const gd::PairParams<float>& pp = a.pp;
    if (pp.flag == 1 && (pp.fun == gd::kFunNone || pp.fun == gd::kFunLog1p)) {
      switch (pp.fun | (pp.tau_on << 2) | (1 << 3)) {
        case 8: return launch_rowlane_inst<LOSS, 8>(a, st);
        case 9: return launch_rowlane_inst<LOSS, 9>(a, st);
        case 12: return launch_rowlane_inst<LOSS, 12>(a, st);
        case 13: return launch_rowlane_inst<LOSS, 13>(a, st);
        default: break;
      }
    }
  }
  return launch_rowlane_inst<LOSS, -1>(a, st);
}

// Filter pass: one column chunk only (m <= 256); returns GD_ERR_LAYOUT otherwise (the caller
// falls back to its own chunked kernel).  Persistent, even waves, like the matrix launch.
template <int LOSS, int SPEC, int CPL>
int launch_filter_inst(const PairwiseArgs& a, const FilterArgs& f, cudaStream_t st) {
  auto kern = gd_pairwise_filter_kernel<LOSS, SPEC, CPL>;
  static int occ[kMaxDevices] = {};
  const int dev = current_device();
  if (occ[dev] == 0) {
    int per_sm = 0;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
    if (e != cudaSuccess) return (int)e;
    occ[dev] = per_sm > 0 ? per_sm : 1;
  }
  const long long slots = (long long)device_info().sm_count * occ[dev];
  const long long ntiles = (a.n + kRowsPerCtaBig - 1) / kRowsPerCtaBig;
  const long long waves = (ntiles + slots - 1) / slots;
  long long rows = (a.n + waves * slots - 1) / (waves * slots);
  if (rows < 16) rows = 16;
  if (rows > kRowsPerCtaBig) rows = kRowsPerCtaBig;
  PairwiseArgs b = a;
  b.tile_rows = (int)rows;
  const long long nt = (a.n + rows - 1) / rows;
  kern<<<(unsigned)(nt < slots ? nt : slots), kThreads, 0, st>>>(b, f);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS>
int launch_filter(const PairwiseArgs& a, const FilterArgs& f, cudaStream_t st) {
  constexpr int CPL = pairwise_cpl2_pays<LOSS>() ? 2 : 1;
  const int cpl = (CPL == 2 && a.m > 32) ? 2 : 1;
  if (a.m > 32LL * cpl * kWarps || a.n <= 0) return GD_ERR_LAYOUT;
  const gd::PairParams<float>& pp = a.pp;
  const bool spec13 = pp.flag == 1 && pp.fun == gd::kFunLog1p && pp.tau_on == 1;   // the SimOTA setting
  if constexpr (CPL == 2) {
    if (cpl == 2)
      return spec13 ? launch_filter_inst<LOSS, 13, 2>(a, f, st) : launch_filter_inst<LOSS, -1, 2>(a, f, st);
  }
  return spec13 ? launch_filter_inst<LOSS, 13, 1>(a, f, st) : launch_filter_inst<LOSS, -1, 1>(a, f, st);
}

template <int LOSS>
int launch_pairwise(const PairwiseArgs& a, cudaStream_t st) {
  // reductions without the matrix: lanes on rows (no per-row collective); with the matrix (or
  // more columns than the shared-memory stage holds): lanes on columns, coalesced stores
  // (for the distances whose pair value is fixed to the last bit in the source,
  // gd::PairwiseExact: the two kernels must agree exactly)
  if constexpr (gd::PairwiseExact<LOSS>::value) {
    if (a.row_min && a.out == nullptr && a.m <= kRowLaneCols && !a.force_cpl1 &&
        pairwise_rowlane_enabled())
      return launch_rowlane<LOSS>(a, st);
  }
  return a.row_min ? launch_pairwise_spec<LOSS, true>(a, st)
                   : launch_pairwise_spec<LOSS, false>(a, st);
}
#endif  // !GD_HOST_EMULATION

}  // namespace gdk
