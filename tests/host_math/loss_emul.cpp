// TEST-ONLY: runs the fused loss kernels' SOURCE (csrc/gd_loss_kernels.cuh: gd_warp_kernel
// with its per-warp copy rings, tile schedule, balanced last round, FAST / robust row
// handling, packed-math pairing, deterministic grid-wide sum; and gd_staged_kernel) on the
// host under the execution-model emulation of cuda_emul.h.  Bulk copies and mbarriers are
// synchronous stand-ins (csrc/gd_common.cuh, GD_HOST_EMULATION), so copy/compute races are
// invisible here; what is checked is the schedule (every row processed exactly once for any
// n, grid and warp count), the shared-memory layouts (128-bit and strided), tails, weights
// modes, masking, row-loss output and the reduction.  Compile-time variants come from
// -DGD_TUNE_DEFAULT=<bits> exactly as in tools/prepare_variants.sh.  Arithmetic is the host
// instantiation (plain float); scalar and packed FAST cores are bit-identical there.
// Never part of the shipped library; not a CPU fallback.
#include "cuda_emul.h"

#include "../../mmdet3d_gaussian_b200/csrc/gd_loss_kernels.cuh"

namespace gdk {
std::atomic<int64_t> g_launches{0};
}

template <int LOSS, int SPEC, int WM, bool PACK>
static void run_warp(const gdk::LossArgs& a, int grid, int warps) {
  const gdk::WarpLayout L = gdk::warp_layout(4, a.wmode, a.grad != nullptr, a.row_loss != nullptr);
  if ((size_t)warps * L.per_warp > sizeof(g_emu_smem)) return;
  if (a.grad != nullptr) {
    emu_launch(gdk::gd_warp_kernel<LOSS, true, 4, SPEC, WM, PACK>, (unsigned)grid, 1, warps * 32, a);
  } else {
    if constexpr (!PACK && SPEC < 0)
      emu_launch(gdk::gd_warp_kernel<LOSS, false, 4, SPEC, WM, PACK>, (unsigned)grid, 1, warps * 32, a);
  }
}

template <int LOSS>
static int run_loss(const gdk::LossArgs& a, int kind, int spec, int pack, int grid, int warps) {
  if (kind == 0) {                                   // staged kernel
    if (a.grad) emu_launch(gdk::gd_staged_kernel<LOSS, true>, (unsigned)grid, 1, gdk::kThreads, a);
    else emu_launch(gdk::gd_staged_kernel<LOSS, false>, (unsigned)grid, 1, gdk::kThreads, a);
    return 0;
  }
  // warp kernel: the instantiations launch_warp() selects
  if (spec < 0) {
    run_warp<LOSS, -1, -1, false>(a, grid, warps);
    return 0;
  }
  if (a.wmode == GD_WEIGHT_ROW7 || a.row_loss || !a.grad) return 1;   // no specialised kernel
#define GD_EMU_SPEC(S)                                                    \
  if (spec == S) {                                                        \
    if (a.wmode == GD_WEIGHT_ROW) {                                       \
      if (pack) run_warp<LOSS, S, 1, true>(a, grid, warps);               \
      else run_warp<LOSS, S, 1, false>(a, grid, warps);                   \
    } else {                                                              \
      if (pack) run_warp<LOSS, S, 0, true>(a, grid, warps);               \
      else run_warp<LOSS, S, 0, false>(a, grid, warps);                   \
    }                                                                     \
    return 0;                                                             \
  }
  GD_EMU_SPEC(8) GD_EMU_SPEC(9) GD_EMU_SPEC(13)
#undef GD_EMU_SPEC
  return 1;
}

// kind: 0 staged, 1 warp pipeline.  spec: -1 run-time parameters, else 8 (fun none), 9 (log1p),
// 13 (log1p + tau) with flag on.  loss: 0 gwd3d, 1 kld3d, 5 bd3d.  Returns 0, 1 = no such
// instantiation, 2 = the workspace was not left zeroed.
extern "C" int gd_emul_loss(int loss, int kind, int spec, int pack, int grid, int warps,
                            const float* pred, const float* target, const float* weight,
                            int wmode, long long n, float scale, float tau, int mask_zero,
                            float* loss_sum, float* row_loss, float* grad) {
  gd_loss_config cfg{};
  cfg.loss_type = loss;
  cfg.fun = spec < 0 ? GD_FUN_LOG1P : ((spec & 3) == 1 ? GD_FUN_LOG1P : GD_FUN_NONE);
  cfg.flag = 1;
  cfg.tau = tau;
  cfg.alpha = 1.0f;
  cfg.center_offset[0] = 0.0f;
  cfg.center_offset[1] = 0.0f;
  cfg.center_offset[2] = 0.5f;
  std::vector<double> partials((size_t)grid + 1, 0.0);
  unsigned ticket = 0;
  gdk::LossArgs a{};
  a.pred = pred;
  a.target = target;
  a.weight = weight;
  a.pstride = 7;
  a.tstride = 7;
  a.wstride = wmode == GD_WEIGHT_ROW7 ? 7 : 1;
  a.n = n;
  a.wmode = wmode;
  a.mask_zero_w = mask_zero;
  a.scale = scale;
  a.loss_sum = loss_sum;
  a.row_loss = row_loss;
  a.grad = grad;
  a.partials = partials.data();
  a.ticket = &ticket;
  a.pp = gdk::make_pair_params(cfg);
  a.tune = GD_TUNE_DEFAULT;
  int rc = 1;
  if (loss == 0) rc = run_loss<0>(a, kind, spec, pack, grid, warps);
  if (loss == 1) rc = run_loss<1>(a, kind, spec, pack, grid, warps);
  if (loss == 5) rc = run_loss<5>(a, kind, spec, pack, grid, warps);
  if (rc == 0 && ticket != 0) rc = 2;
  return rc;
}

extern "C" int gd_emul_tune_default() { return GD_TUNE_DEFAULT; }
