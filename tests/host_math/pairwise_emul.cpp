// TEST-ONLY: runs the pairwise CUDA kernel's SOURCE (csrc/gd_pairwise.cuh: gd_pairwise_kernel
// with one and with two columns per lane) on the host, so their index,
// tiling, tie-breaking and reduction logic can be checked on a machine without a GPU.
//
// How: the header is compiled by g++ under the execution-model emulation of cuda_emul.h (one
// OS thread per CUDA thread, barriers for __syncthreads and the warp collectives, CTAs of the
// grid one after the other).  The arithmetic is the host
// instantiation of gd_math.cuh / gd_packed.cuh (plain float, no MUFU, no FFMA2), so VALUES
// are not what the device computes to the last bit; what is exercised is everything around
// the arithmetic: which (row, column) lands where, partial tiles, odd row counts, dead
// lanes, chunked columns, the workspace protocol, NaN-first / lowest-index tie rules.
// It is never linked into the shipped library and is not a CPU fallback.
//
// Output: a small binary protocol on stdout is avoided; the harness is a shared library
// driven by tests/test_pairwise_emulation.py through ctypes.
#include "cuda_emul.h"

#include "../../mmdet3d_gaussian_b200/csrc/gd_pairwise.cuh"

namespace gdk {
std::atomic<int64_t> g_launches{0};
}

static gdk::PairwiseArgs make_args(const gd_loss_config* cfg, const float* b1, long long n,
                                   const float* b2, long long m, float* out, float* row_min,
                                   int* row_argmin, float* col_min, int* col_argmin,
                                   unsigned long long* col_keys, unsigned* ticket, int similarity) {
  gdk::PairwiseArgs a{};
  a.b1 = b1;
  a.n = n;
  a.b2 = b2;
  a.m = m;
  a.out = out;
  a.out_stride = m;
  a.similarity = similarity;
  a.row_min = row_min;
  a.row_argmin = row_argmin;
  a.col_keys = col_keys;
  a.col_min = col_min;
  a.col_argmin = col_argmin;
  a.ticket = ticket;
  a.pp = gdk::make_pair_params(*cfg);
  return a;
}

template <int LOSS, int SPEC, bool REDUCE, int CPL>
static void run_cpl(const gdk::PairwiseArgs& a, unsigned cap) {
  // the grid rules of launch_pairwise_cpl (csrc/gd_pairwise.cuh), with the SM count as input
  const long long ntiles = (a.n + gdk::kRowsPerCta - 1) / gdk::kRowsPerCta;
  const int wx = gdk::pairwise_wx(a.m, 32 * CPL);
  unsigned gx, gy = 1;
  if (REDUCE) {
    long long g = ntiles;
    if (a.col_keys && g > cap) g = cap;
    gx = (unsigned)g;
  } else {
    gx = (unsigned)ntiles;
    gy = (unsigned)((a.m + 32LL * CPL * wx - 1) / (32LL * CPL * wx));
    if (gy == 1 && gd::PairwiseExact<LOSS>::value) {   // persistent, even waves: launch_pairwise_cpl with `cap` CTA slots
      const long long slots = cap;
      const long long big_tiles = (a.n + gdk::kRowsPerCtaBig - 1) / gdk::kRowsPerCtaBig;
      const long long waves = (big_tiles + slots - 1) / slots;
      long long rows = (a.n + waves * slots - 1) / (waves * slots);
      if (rows < 16) rows = 16;
      if (rows > gdk::kRowsPerCtaBig) rows = gdk::kRowsPerCtaBig;
      gdk::PairwiseArgs b = a;
      b.tile_rows = (int)rows;
      const long long nt = (a.n + rows - 1) / rows;
      emu_launch(gdk::gd_pairwise_kernel<LOSS, SPEC, REDUCE, CPL>, (unsigned)(nt < slots ? nt : slots), 1,
                 gdk::kThreads, b);
      return;
    }
  }
  emu_launch(gdk::gd_pairwise_kernel<LOSS, SPEC, REDUCE, CPL>, gx, gy, gdk::kThreads, a);
}

// the row-lane kernel of the fused reductions (no matrix), with launch_rowlane_inst's grid rule
template <int LOSS, int SPEC, int RPL>
static void run_rowlane(const gdk::PairwiseArgs& a, unsigned cap) {
  const long long nunits = (a.n + 32 * RPL - 1) / (32 * RPL);
  long long grid = (nunits + gdk::kWarps - 1) / gdk::kWarps;
  if (grid > cap) grid = cap;
  emu_launch(gdk::gd_pairwise_rowlane_kernel<LOSS, SPEC, RPL>, (unsigned)grid, 1, gdk::kThreads, a);
}

// loss: 0 gwd3d, 1 kld3d, 5 bd3d; fun log1p, tau >= 1 (SPEC 13), flag on -- the C4 configuration.
// packed: 0 = one column per lane (CPL 1), 1 = two columns per lane (CPL 2; the default mapping
// for m > 32).  reduce: fused minima (+ the matrix when `out`).
// force_cpl: 3 / 4 = the ROW-lane kernel with two / one rows per lane (reduce only, no matrix,
// m <= 512; dynamic unit schedule), 5 = the same without column minima (static schedule, as
// gd_pairwise_row_argmin launches it); else unused.
// cap: CTA cap of the persistent reductions (the library uses 6 x SM count).
extern "C" int gd_emul_pairwise(int loss, int packed, int reduce, int force_cpl, unsigned cap,
                                const float* b1, long long n, const float* b2, long long m,
                                float* out, float* row_min, int* row_argmin, float* col_min,
                                int* col_argmin, int similarity) {
  gd_loss_config cfg{};
  cfg.loss_type = loss;
  cfg.fun = GD_FUN_LOG1P;
  cfg.flag = 1;
  cfg.tau = 1.0f;
  cfg.alpha = 1.0f;
  cfg.center_offset[0] = 0.0f;
  cfg.center_offset[1] = 0.0f;
  cfg.center_offset[2] = 0.5f;
  std::vector<unsigned long long> keys((size_t)m + 1, 0ull);
  unsigned ticket[8] = {};
  gdk::PairwiseArgs a =
      make_args(&cfg, b1, n, b2, m, out, reduce ? row_min : nullptr, reduce ? row_argmin : nullptr,
                col_min, col_argmin, reduce ? keys.data() : nullptr, ticket, similarity);
  constexpr int S = 13;
  if (force_cpl >= 3) {
    if (!reduce || m > gdk::kRowLaneCols) return -1;
    a.out = nullptr;
    if (force_cpl == 5) {
      a.col_keys = nullptr;
      a.ticket = nullptr;
    }
#define GD_EMU_ROWLANE(L)                                                       \
    if (loss == L) {                                                            \
      if (force_cpl == 4) run_rowlane<L, S, 1>(a, cap);                         \
      else run_rowlane<L, S, 2>(a, cap);                                        \
    }
    GD_EMU_ROWLANE(0) GD_EMU_ROWLANE(1) GD_EMU_ROWLANE(5)
#undef GD_EMU_ROWLANE
    int dirty = 0;
    for (unsigned t : ticket) dirty |= (t != 0u);
    for (auto k : keys) dirty |= (k != 0ull);
    return dirty;
  }
#define GD_EMU_CASE(L)                                                          \
  if (loss == L) {                                                              \
    if (packed) {                                                               \
      if (reduce) run_cpl<L, S, true, 2>(a, cap);                               \
      else run_cpl<L, S, false, 2>(a, cap);                                     \
    } else {                                                                    \
      if (reduce) run_cpl<L, S, true, 1>(a, cap);                               \
      else run_cpl<L, S, false, 1>(a, cap);                                     \
    }                                                                           \
  }
  GD_EMU_CASE(0) GD_EMU_CASE(1) GD_EMU_CASE(5)
#undef GD_EMU_CASE
  // workspace protocol: the kernels must leave keys and ticket zeroed again
  int dirty = 0;
  for (unsigned t : ticket) dirty |= (t != 0u);
  for (auto k : keys) dirty |= (k != 0ull);
  return dirty;
}
