// Box decoders fused IN FRONT of the Gaussian-distance loss.
//
// In the reference the loss never sees raw network outputs: its two call sites
// first decode them into real-world boxes with a chain of eager torch ops,
//
//   * GDAnchor3DHead.loss_single (models/dense_heads/gd_anchor3d_head.py:107-141):
//       rows gathered at pos_inds, then bbox_coder.decode(anchors, deltas) for the
//       prediction AND the target (:133-136) -- upstream mmdet3d
//       DeltaXYZWLHRBBoxCoder.decode (not in the reference checkout; positional
//       form restated below), then GDLoss on the two decoded [P,7] tensors;
//   * CenterGDHead.loss (models/dense_heads/gd_centerpoint_head.py:413-434):
//       CenterPointBBoxYawCoder.decode(locs, preds, correct_yaw=False)[..., :7]
//       (core/bbox/coders/centerpoint_bbox_yaw_coders.py:18-56) for the prediction;
//       the target rows are already real-world boxes.
//
// Here the decode is the prologue and its Jacobian the epilogue of the same
// per-pair evaluation (gd_math.cuh), so the decoded boxes never exist in memory and
// the gradient arrives at the raw deltas / head outputs directly.
//
// Numerics: prediction and target of the anchor head are decoded against the SAME
// anchor, so everything the distance needs is formed from differences of the
// deltas -- centre difference (d_p - d_t) * diag, extent difference
// a * exp(d_t) * expm1(d_p - d_t), yaw difference d_p - d_t -- instead of from two
// separately rounded absolute boxes.  Against the fp64 oracle this is at least as
// accurate as the reference's fp32 decode-then-subtract.
#pragma once
#include "gd_math.cuh"

namespace gd {

enum Coder : int {
  kCoderDeltaXYZWLHR = 0,     // mmdet3d DeltaXYZWLHRBBoxCoder.decode        (upstream)
  kCoderCenterPointYaw = 1    // CenterPointBBoxYawCoder.decode, correct_yaw=False
};

// Constants of CenterPointBBoxCoderRev.__init__ (centerpoint_bbox_coders.py:9-20)
// folded on the host: sx = out_size_factor * voxel_size[0], x0 = pc_range[0], ...
struct CenterDecodeParams {
  double sx, sy, x0, y0;
  int norm_bbox;              // dims = exp(pred)            centerpoint_bbox_yaw_coders.py:37-38
};

template <typename T>
GD_HD T exp_t(T x);
template <>
GD_HD double exp_t<double>(double x) { return ::exp(x); }
template <>
GD_HD float exp_t<float>(float x) { return ::expf(x); }
template <typename T>
GD_HD T expm1_t(T x);
template <>
GD_HD double expm1_t<double>(double x) { return ::expm1(x); }
template <>
GD_HD float expm1_t<float>(float x) { return ::expm1f(x); }
template <typename T>
GD_HD T sqrt_ieee(T x);
template <>
GD_HD double sqrt_ieee<double>(double x) { return ::sqrt(x); }
template <>
GD_HD float sqrt_ieee<float>(float x) { return ::sqrtf(x); }

// Yaw part of PairGeom from (r_p, r_t) and their exactly known difference.
template <typename T, bool NEED_PRED_ROT>
GD_HD void geom_set_yaw(PairGeom<T>* g, T rp, T rt, T dr) {
  const T ybig = (T)16;
  if (rp >= -ybig && rp <= ybig && rt >= -ybig && rt <= ybig) {
    Mth<T>::sincos(dr, &g->sd, &g->cd);
    if (NEED_PRED_ROT) {
      Mth<T>::sincos(rp, &g->sp, &g->cp);
    } else {
      g->sp = (T)0;
      g->cp = (T)1;
    }
  } else {                      // huge yaws: same treatment as make_geom
    T st, ct;
    Mth<T>::sincos(rp, &g->sp, &g->cp);
    Mth<T>::sincos(rt, &st, &ct);
    g->sd = g->sp * ct - g->cp * st;
    g->cd = g->cp * ct + g->sp * st;
  }
}

// ---------------------------------------------------------------------------
// Anchor head: pred box = decode(anchor, dp), target box = decode(anchor, dt).
//   decode (positional, identical in every mmdet3d release that has the coder):
//     diag = sqrt(a3^2 + a4^2);  x = d0 diag + a0;  y = d1 diag + a1;
//     z = d2 a5 + (a2 + a5/2) - exp(d5) a5 / 2;
//     ext_k = exp(d_k) a_k, k = 3,4,5;  yaw = d6 + a6
// Returns the (unweighted) loss of the pair; gdelta[0..6] = gscale * d loss / d dp.
// ---------------------------------------------------------------------------
template <typename T, int LOSS, bool GRAD>
GD_HD T anchor_pair_eval(const T* a, const T* dp, const T* dt, const PairParams<T>& P, T gscale,
                         T* gdelta) {
  constexpr bool kNeedRot = !(LOSS == kGwd || LOSS == kKfiou);
  const T diag = sqrt_ieee(a[4] * a[4] + a[3] * a[3]);
  T pe[3], te[3], de[3];                       // decoded extents and their differences
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    te[k] = exp_t(dt[3 + k]) * a[3 + k];
    pe[k] = exp_t(dp[3 + k]) * a[3 + k];
    de[k] = te[k] * expm1_t(dp[3 + k] - dt[3 + k]);          // pe - te without cancellation
  }
  PairGeom<T> g;
  // centre = xyz + off * UNCLAMPED extents (ref:12); the anchor's own position cancels
  g.dx = (dp[0] - dt[0]) * diag + P.off[0] * de[0];
  g.dy = (dp[1] - dt[1]) * diag + P.off[1] * de[1];
  g.dz = (dp[2] - dt[2]) * a[5] + (P.off[2] - (T)0.5) * de[2];
  T dummy;
  g.ap = (T)0.5 * clamp_extent(pe[0], &g.ma);
  g.bp = (T)0.5 * clamp_extent(pe[1], &g.mb);
  g.ep = (T)0.5 * clamp_extent(pe[2], &g.me);
  g.at = (T)0.5 * clamp_extent(te[0], &dummy);
  g.bt = (T)0.5 * clamp_extent(te[1], &dummy);
  g.et = (T)0.5 * clamp_extent(te[2], &dummy);
  geom_set_yaw<T, kNeedRot>(&g, dp[6] + a[6], dt[6] + a[6], dp[6] - dt[6]);
  geom_derive(&g);
  T gb[7];
  bool unused = false;
  const T val = core_eval<T, LOSS, GRAD, false>(g, P, gscale, gb, &unused);
  if (GRAD) {
    gdelta[0] = gb[0] * diag;                  // d x / d d0 = diag
    gdelta[1] = gb[1] * diag;
    gdelta[2] = gb[2] * a[5];                  // d z / d d2 = a5
    gdelta[3] = gb[3] * pe[0];                 // d ext / d d = ext
    gdelta[4] = gb[4] * pe[1];
    gdelta[5] = (gb[5] - (T)0.5 * gb[2]) * pe[2];   // z also moves by -ext5/2
    gdelta[6] = gb[6];
  }
  return val;
}

// ---------------------------------------------------------------------------
// CenterPoint head: pred box = decode(loc, pr) with correct_yaw=False
//   x = (pr0 + loc_x) * out_size_factor * voxel_size[0] + pc_range[0]   (:30-31)
//   y likewise (:32-33);  z = pr2 (:34);  dims = exp(pr[3:6]) if norm_bbox (:35-37);
//   yaw = pr6 (:38).
// The target row is a real-world box.  x/y are formed in double: (pr + loc) carries
// an integer part up to the feature-map size, and the loss depends on x_p - x_t.
// gpred[0..6] = gscale * d loss / d pr[0..6].
// ---------------------------------------------------------------------------
template <typename T, int LOSS, bool GRAD>
GD_HD T center_pair_eval(const T* pr, long long loc_x, long long loc_y, const T* t,
                         const CenterDecodeParams& D, const PairParams<T>& P, T gscale,
                         T* gpred) {
  constexpr bool kNeedRot = !(LOSS == kGwd || LOSS == kKfiou);
  T pe[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) pe[k] = D.norm_bbox ? exp_t(pr[3 + k]) : pr[3 + k];
  PairGeom<T> g;
  const double xd = ((double)pr[0] + (double)loc_x) * D.sx + D.x0 - (double)t[0];
  const double yd = ((double)pr[1] + (double)loc_y) * D.sy + D.y0 - (double)t[1];
  g.dx = (T)xd + P.off[0] * (pe[0] - t[3]);
  g.dy = (T)yd + P.off[1] * (pe[1] - t[4]);
  g.dz = (pr[2] - t[2]) + P.off[2] * (pe[2] - t[5]);
  T dummy;
  g.ap = (T)0.5 * clamp_extent(pe[0], &g.ma);
  g.bp = (T)0.5 * clamp_extent(pe[1], &g.mb);
  g.ep = (T)0.5 * clamp_extent(pe[2], &g.me);
  g.at = (T)0.5 * clamp_extent(t[3], &dummy);
  g.bt = (T)0.5 * clamp_extent(t[4], &dummy);
  g.et = (T)0.5 * clamp_extent(t[5], &dummy);
  geom_set_yaw<T, kNeedRot>(&g, pr[6], t[6], pr[6] - t[6]);
  geom_derive(&g);
  T gb[7];
  bool unused = false;
  const T val = core_eval<T, LOSS, GRAD, false>(g, P, gscale, gb, &unused);
  if (GRAD) {
    gpred[0] = gb[0] * (T)D.sx;
    gpred[1] = gb[1] * (T)D.sy;
    gpred[2] = gb[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) gpred[3 + k] = D.norm_bbox ? gb[3 + k] * pe[k] : gb[3 + k];
    gpred[6] = gb[6];
  }
  return val;
}

}  // namespace gd
