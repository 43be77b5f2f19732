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

// gd_warp_kernel<LOSS, GRAD, 4, -1, -1, false, ANY = true>: row-strided / unaligned inputs,
// plus the device-side scale divisor and the any-positive status word.  The plan (row_lo,
// n_bulk, shifts) is the library's own (gdk::plan_any).  ws = {ticket, any word}.
template <int LOSS>
static void run_any(const gdk::LossArgs& a, int grid, int warps) {
  const gdk::WarpLayout L = gdk::warp_layout(4, a.wmode, a.grad != nullptr, a.row_loss != nullptr,
                                             true, a.pstride, a.tstride, a.wstride);
  const int fit = (int)((long long)gdk::kSmemBudget / L.per_warp);   // as launch_warp_inst
  if (warps > fit) warps = fit;
  if (warps < 1) return;
  if (a.grad) emu_launch(gdk::gd_warp_kernel<LOSS, true, 4, -1, -1, false, true>, (unsigned)grid, 1, warps * 32, a);
  else emu_launch(gdk::gd_warp_kernel<LOSS, false, 4, -1, -1, false, true>, (unsigned)grid, 1, warps * 32, a);
}

extern "C" int gd_emul_loss_any(int loss, int kind, int grid, int warps, const float* pred,
                                long long pstride, const float* target, long long tstride,
                                const float* weight, int wmode, long long wstride, long long n,
                                float scale, const float* scale_div, float tau, float* status,
                                float* loss_sum, float* row_loss, float* grad, int early_return,
                                long long er_wrow, long long er_wcol) {
  gd_loss_config cfg{};
  cfg.loss_type = loss;
  cfg.fun = GD_FUN_LOG1P;
  cfg.flag = 1;
  cfg.tau = tau;
  cfg.alpha = 1.0f;
  cfg.center_offset[2] = 0.5f;
  std::vector<double> partials((size_t)grid + 1, 0.0);
  unsigned ws[4] = {0, 0, 0, 0};
  gdk::LossArgs a{};
  a.pred = pred;
  a.target = target;
  a.weight = weight;
  a.pstride = pstride;
  a.tstride = tstride;
  a.wstride = wmode == GD_WEIGHT_NONE ? 0 : wstride;
  a.n = n;
  a.wmode = wmode;
  a.scale = scale;
  a.scale_div = scale_div;
  a.status = status;
  a.early_return = early_return;
  a.er_wrow = er_wrow;
  a.er_wcol = er_wcol;
  a.loss_sum = loss_sum;
  a.row_loss = row_loss;
  a.grad = grad;
  a.partials = partials.data();
  a.ticket = ws;
  a.pp = gdk::make_pair_params(cfg);
  a.tune = GD_TUNE_DEFAULT;
  if (kind == 1) {
    if (n < 16) return 1;
    gdk::plan_any(&a);
    if (loss == 0) run_any<0>(a, grid, warps);
    else if (loss == 1) run_any<1>(a, grid, warps);
    else if (loss == 5) run_any<5>(a, grid, warps);
    else return 1;
  } else {
    if (loss == 0) run_loss<0>(a, 0, -1, 0, grid, warps);
    else if (loss == 1) run_loss<1>(a, 0, -1, 0, grid, warps);
    else if (loss == 5) run_loss<5>(a, 0, -1, 0, grid, warps);
    else return 1;
  }
  return (ws[0] != 0 || ws[1] != 0) ? 2 : 0;
}
