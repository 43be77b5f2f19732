// Per-pair Gaussian-distance math: value and analytic d/d(pred) in registers.
//
// One call evaluates one (pred, target) box pair for one of the seven distance
// types of the reference (mmdet3d_gaussian/models/losses/gaussian_distance_loss.py,
// cited as ref:LINE) and, when GRAD, the hand-derived gradient w.r.t. the seven
// pred columns, replacing the reference's autograd chain of ~160-250 small ops.
//
// Formulation.  The reference builds R, S, Sigma = R S^2 R^T with batched 2x2
// matmuls (ref:86-88,116-117,149-150).  Here everything is written in the pred
// box's own frame: with half extents a=w/2, b=h/2, e=l/2 (A=a^2 ...; C,D,F for
// the target), yaw difference dl = r_p - r_t, s2 = sin^2 dl:
//
//   tr(Sigma_p Sigma_t)     = (AC+BD) cos^2 + (AD+BC) sin^2             (ref:88-90)
//   det(Sigma_p + Sigma_t)  = (A+C)(B+D) + (A-B)(C-D) s2                (ref:155-157,235-236)
//   Sigma_t in pred frame   = [[C c^2 + D s^2, (D-C) s c], [., C s^2 + D c^2]]
//
// and terms that are differences of nearly equal quantities when the boxes
// almost coincide are evaluated from differences/ratios instead (SURVEY.md
// section 7 "hard parts"), e.g. A+B+C+D-2 sqrt(U) = (a_p-a_t)^2+(b_p-b_t)^2+2 eta with
// eta = (A-B)(C-D) s2 / (V + sqrt(U)), V = a_p a_t + b_p b_t.  Against the fp64
// oracle this is at least as accurate as the reference's own fp32 evaluation.
//
// The file compiles for the device (nvcc, T=float) and, for formula
// verification only, on the host (tests/host_math builds it with g++ in float
// and double).  The shipped library contains no host compute path.
#pragma once
#include <type_traits>

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef GD_PRECISE_MATH
#define GD_PRECISE_MATH 0
#endif

#if defined(__CUDACC__)
#define GD_HD __host__ __device__ __forceinline__
#else
#define GD_HD inline
#endif

namespace gd {

enum LossType : int {
  kGwd = 0,      // 'gwd3d'         ref:42-106
  kKld = 1,      // 'kld3d'         ref:109-141
  kJd = 2,       // 'jd3d'          ref:189-198
  kSymMax = 3,   // 'kld3d_symmax'  ref:201-211
  kSymMin = 4,   // 'kld3d_symmin'  ref:214-224
  kBd = 5,       // 'bd3d'          ref:144-186
  kKfiou = 6,    // 'kfiou3d'       ref:227-248
  kNumLossTypes = 7
};

enum Fun : int { kFunNone = 0, kFunLog1p = 1, kFunExpm1 = 2, kFunNlog = 3 };  // ref:24-34

template <typename T>
struct PairParams {
  T off[3];      // center_offset                      ref:12
  T alpha2;      // alpha^2                            ref:99
  T inv_alpha2;  // 1/alpha^2                          ref:137,182
  T tau;         // only used when tau_on              ref:36-37
  int fun;       // Fun
  int tau_on;    // tau >= 1.0 decided on the host     ref:36
  int flag;      // 'normalize' (gwd) or 'sqrt' (others)   ref:43,110,145
  int lean = 0;  // pairwise value path: log1p of the post map by Mth::log1p_lean (~4e-7)
};

// Optional compile-time "diet" of the FAST cores (template parameter DIET, default 0 =
// the validated code, unchanged).  Both bits remove instructions only; see DESIGN.md.
//   kDietGuards: the "nice row" screen of make_geom uses 3-input NaN-propagating
//                min/max (FMNMX3.NAN) instead of 16 compares -- same decision.
//   kDietStd:    the caller promises alpha == 1 and center_offset == (0, 0, 0.5) (the
//                reference's defaults, ref:261-262, and what every shipped config uses):
//                the multiplications by alpha^2 / 1/alpha^2 and by the zero x/y offsets
//                disappear.  Values differ from DIET = 0 by FMA contraction only (<= 1 ulp
//                per operation).
enum Diet : int { kDietGuards = 1, kDietStd = 2 };

template <int DIET, typename T>
GD_HD T a2_times(const PairParams<T>& P, T x) {      // alpha^2 * x
  if constexpr ((DIET & kDietStd) != 0) return x;
  else return P.alpha2 * x;
}
template <int DIET, typename T>
GD_HD T times_ia2(T x, const PairParams<T>& P) {     // x / alpha^2
  if constexpr ((DIET & kDietStd) != 0) return x;
  else return x * P.inv_alpha2;
}

// ---------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------
template <typename T>
struct Mth;

template <>
struct Mth<double> {
  typedef bool mask;                         // result type of a comparison (one bit per value)
  static GD_HD double sel(bool c, double a, double b) { return c ? a : b; }
  static GD_HD double rcp(double x) { return 1.0 / x; }
  static GD_HD double sqrt(double x) { return ::sqrt(x); }
  static GD_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static GD_HD void sincos(double x, double* s, double* c) {
    *s = ::sin(x);
    *c = ::cos(x);
  }
  static GD_HD double log(double x) { return ::log(x); }
  static GD_HD double log1p(double x) { return ::log1p(x); }
  static GD_HD double expm1(double x) { return ::expm1(x); }
  static GD_HD double exp(double x) { return ::exp(x); }
  static GD_HD double rcbrt(double x) { return 1.0 / ::cbrt(x); }
  // S - log(r1 r2 r3), r_i = 1 + q_i, S = sum q_i; plain evaluation is accurate
  // enough in double.
  template <bool FAST>
  static GD_HD double sum_minus_log_ratios(double S, double pair, double r1, double r2,
                                           double r3, bool* rare) {
    (void)pair;
    (void)rare;
    return S - (::log(r1) + ::log(r2) + ::log(r3));
  }
  static GD_HD double inf() { return HUGE_VAL; }
  // FAST-path stand-ins (the double instantiation only checks formulas)
  static GD_HD void sincos_fast(double x, double* s, double* c) { sincos(x, s, c); }
  static GD_HD double log1p_pos(double x) { return ::log1p(x); }
  static GD_HD double rsixthroot(double x) { return ::pow(x, -1.0 / 6.0); }
  static GD_HD bool nice_row_minmax(double p3, double p4, double p5, double t3, double t4,
                                    double t5, double yp, double yt) {
    const double lo = 1e-4, hi = 1e4, ymax = 16;
    return p3 >= lo && p4 >= lo && p5 >= lo && t3 >= lo && t4 >= lo && t5 >= lo && p3 <= hi &&
           p4 <= hi && p5 <= hi && t3 <= hi && t4 <= hi && t5 <= hi && yp >= -ymax &&
           yp <= ymax && yt >= -ymax && yt <= ymax;
  }
};

template <>
struct Mth<float> {
  typedef bool mask;
  static GD_HD float sel(bool c, float a, float b) { return c ? a : b; }
  // Device arithmetic: MUFU approximations (rcp/sqrt/rsqrt.approx.ftz: <= 1-2 ulp)
  // instead of the IEEE-rounded sequences (each ~8 instructions + a slow-path
  // call).  The parity budget is 1e-5 relative; these cost ~1e-7 per operation
  // and keep the kernel on the HBM side of the roofline.  -DGD_PRECISE_MATH=1
  // restores the IEEE versions (used by the tests to bound the difference).
  static GD_HD float rcp(float x) {
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#elif defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
  }
  static GD_HD float sqrt(float x) {
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#elif defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return ::sqrtf(x);
#endif
  }
  static GD_HD float rsqrt(float x) {
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#elif defined(__CUDA_ARCH__)
    return __frcp_rn(__fsqrt_rn(x));
#else
    return 1.0f / ::sqrtf(x);
#endif
  }
  static GD_HD void sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    ::sincosf(x, s, c);
#else
    *s = ::sinf(x);
    *c = ::cosf(x);
#endif
  }
  static GD_HD float log(float x) { return ::logf(x); }
  static GD_HD float log1p(float x) { return ::log1pf(x); }
  static GD_HD float expm1(float x) { return ::expm1f(x); }
  static GD_HD float exp(float x) { return ::expf(x); }
  static GD_HD float rcbrt(float x) {
#if defined(__CUDA_ARCH__)
    return ::rcbrtf(x);
#else
    return 1.0f / ::cbrtf(x);
#endif
  }
  // S - log(y) with S = q1+q2+q3, y = r1 r2 r3, r_i = 1+q_i the extent ratios, and
  // pair = y - 1 - S (the mixed products of the q_i).  This is the log-det part of
  // the KLD shape term; together with sum q_i^2/2 it vanishes QUADRATICALLY as the
  // boxes coincide, so it must be accurate in a relative sense for |q| << 1, where
  // the naive S - log(y) loses everything -- and it must not cancel for |q| >> 1
  // (degenerate 1e-7 extents, ref:13-14) either.
  //   y = 2^k m, m in [sqrt(1/2), sqrt(2));  s = f/(2+f);  log(m) = 2s + 2 s^3 P(s^2)
  //   k == 0 (y near 1): f = x := S + pair (exact small quantity, not y - 1);
  //            x - log(1+x) = x s - 2 s^3 P(s^2), result = that - pair
  //   k != 0: result = S - k ln2 - 2s - tail directly (it is O(0.06) or larger)
  // y is formed as a product of ratios so it stays accurate when x is within
  // rounding of -1.
  template <bool FAST>
  static GD_HD float sum_minus_log_ratios(float S, float pair, float r1, float r2,
                                          float r3, bool* rare) {
    const float y = r1 * r2 * r3;
    if (FAST) {                                // caller re-runs the row on the robust path
      *rare |= !(y > 1.0e-30f && y < 1.0e30f);
    } else if (!(y > 1.0e-37f) || !(y < 3.0e38f)) {   // product under/overflowed, or nan
      return S - (::logf(r1) + ::logf(r2) + ::logf(r3));
    }
    uint32_t yb;
    memcpy(&yb, &y, 4);
    const int k = (int)((int32_t)(yb - 0x3f3504f3u) >> 23);
    const uint32_t mb = yb - ((uint32_t)k << 23);
    float m;
    memcpy(&m, &mb, 4);
    const float x = S + pair;
    const float f = (k == 0) ? x : (m - 1.0f);
    const float s = f * rcp(2.0f + f);
    const float z = s * s;
    // |s| <= 0.1716, z <= 0.02944: 6 terms -> truncation < 2e-9 relative to 1/3
    float p = 1.0f / 13.0f;
    p = fmaf(p, z, 1.0f / 11.0f);
    p = fmaf(p, z, 1.0f / 9.0f);
    p = fmaf(p, z, 1.0f / 7.0f);
    p = fmaf(p, z, 1.0f / 5.0f);
    p = fmaf(p, z, 1.0f / 3.0f);
    const float tail = 2.0f * s * z * p;
    const float kf = (float)k;
    float r = fmaf(-kf, 0.693145751953125f, S);          // ln2 hi (k*hi exact)
    r = fmaf(-kf, 1.428606765330187e-06f, r);            // ln2 lo
    const float direct = (r - 2.0f * s) - tail;
    return (k == 0) ? (fmaf(x, s, -tail) - pair) : direct;
  }
  static GD_HD float inf() { return HUGE_VALF; }

  // ---- branch-free helpers of the FAST path (rows pre-screened as "nice") ----
  // sin/cos for |x| <= ~1e4: 3-term Cody-Waite reduction by pi/2 + degree-7/8
  // minimax polynomials on [-pi/4, pi/4]; ~1 ulp, no slow path, no branch.
  static GD_HD void sincos_fast(float x, float* sn, float* cs) {
#if defined(__CUDA_ARCH__)
    const int q = __float2int_rn(x * 0.63661974668502807617f);
#else
    const int q = (int)::rintf(x * 0.63661974668502807617f);
#endif
    const float k = (float)q;
    float r = fmaf(k, -1.5707962512969970703f, x);
    r = fmaf(k, -7.5497894158615963534e-08f, r);
    r = fmaf(k, -5.3903029534742383927e-15f, r);
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3327032626e-3f);
    ps = fmaf(z, ps, -1.6666662693e-1f);
    const float sr = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.44331570e-5f, -1.38878601e-3f);
    pc = fmaf(z, pc, 4.16667275e-2f);
    pc = fmaf(z, pc, -4.99999970e-1f);
    const float cr = fmaf(z, pc, 1.0f);
    const float s0 = (q & 1) ? cr : sr;
    const float c0 = (q & 1) ? sr : cr;
    *sn = (q & 2) ? -s0 : s0;
    *cs = ((q + 1) & 2) ? -c0 : c0;
  }
  // log(1+x) for 0 <= x < ~1e30, branch free, ~1 ulp (same range reduction and
  // series as sum_minus_log_ratios).
  static GD_HD float log1p_pos(float x) {
    const float y = 1.0f + x;
    uint32_t yb;
    memcpy(&yb, &y, 4);
    const int k = (int)((int32_t)(yb - 0x3f3504f3u) >> 23);
    const uint32_t mb = yb - ((uint32_t)k << 23);
    float m;
    memcpy(&m, &mb, 4);
    const float f = (k == 0) ? x : (m - 1.0f);
    const float s = f * rcp(2.0f + f);
    const float z = s * s;
    float p = 1.0f / 13.0f;
    p = fmaf(p, z, 1.0f / 11.0f);
    p = fmaf(p, z, 1.0f / 9.0f);
    p = fmaf(p, z, 1.0f / 7.0f);
    p = fmaf(p, z, 1.0f / 5.0f);
    p = fmaf(p, z, 1.0f / 3.0f);
    const float kf = (float)k;
    const float lo = fmaf(kf, 1.428606765330187e-06f, 2.0f * s * z * p);
    return fmaf(kf, 0.693145751953125f, fmaf(2.0f, s, lo));
  }
  // log(1+x) for 0 <= x < ~1e30 in ~15 instructions instead of ~25, branch free, relative
  // error <= ~5e-7 (the pairwise value path: N x M evaluations, issue bound):
  //   x <= 0.5 : 2 atanh(s), s = x / (2 + x) <= 0.2, series to s^9 (next term < 1e-8)
  //   x >  0.5 : ln2 * lg2(1 + x): MUFU.LG2 (abs. error 2^-22 below 2, relative above) on a
  //              sum whose rounding is <= 6e-8 of a result >= 0.405
  static GD_HD float log1p_lean(float x) {
    const float s = x * rcp(2.0f + x);
    const float z = s * s;
    float p = fmaf(z, 1.0f / 9.0f, 1.0f / 7.0f);
    p = fmaf(z, p, 1.0f / 5.0f);
    p = fmaf(z, p, 1.0f / 3.0f);
    const float t = s + s;
    const float small = fmaf(t * z, p, t);
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + x));
    const float big = l * 0.693147180559945f;
#else
    const float big = ::log1pf(x);
#endif
    return x <= 0.5f ? small : big;
  }
  // x^(-1/6) for normal positive x: one MUFU.LG2 + one MUFU.EX2 (rel. err ~5e-7)
  static GD_HD float rsixthroot(float x) {
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    l *= -(1.0f / 6.0f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
#else
    return rcbrt(sqrt(x));
#endif
  }
  // The "nice row" screen of make_geom<FAST> (six extents in [1e-4, 1e4], both yaws in
  // [-16, 16], NaN fails) with 3-input NaN-propagating min/max: 5 FMNMX + 5 FSETP instead
  // of 16 FSETP.  Same decision for every input.
  static GD_HD bool nice_row_minmax(float p3, float p4, float p5, float t3, float t4, float t5,
                                    float yp, float yt) {
    const float lo = 1e-4f, hi = 1e4f, ymax = 16.0f;
#if defined(__CUDA_ARCH__)
    float mn, mx, ya;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(mn) : "f"(p3), "f"(p4), "f"(p5));
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(mn) : "f"(mn), "f"(t3), "f"(t4));
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(p3), "f"(p4), "f"(p5));
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(mx), "f"(t3), "f"(t4));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(ya) : "f"(fabsf(yp)), "f"(fabsf(yt)));
    return mn >= lo && t5 >= lo && mx <= hi && t5 <= hi && ya <= ymax;
#else
    return p3 >= lo && p4 >= lo && p5 >= lo && t3 >= lo && t4 >= lo && t5 >= lo && p3 <= hi &&
           p4 <= hi && p5 <= hi && t3 <= hi && t4 <= hi && t5 <= hi && yp >= -ymax &&
           yp <= ymax && yt >= -ymax && yt <= ymax;
#endif
  }
};

// torch.clamp(min, max) semantics (NaN propagates); returns the gradient mask
// of the CLOSED interval (clamp passes grad on [min,max]).
template <typename T>
GD_HD T clamp_extent(T v, T* mask) {
  const T lo = (T)1e-7, hi = (T)1e7;           // ref:13-14
  *mask = (v >= lo && v <= hi) ? (T)1 : (T)0;
  return v < lo ? lo : (v > hi ? hi : v);
}

// sqrt(clamp(x, 0)) and its derivative factor 1/(2 sqrt(x)):
// 0 below zero (clamp blocks), +inf at exactly zero (ref:95,99,139,184,197;
// torch autograd convention probed in SURVEY.md appendix A).
template <typename T>
GD_HD T sqrt_clamp0(T x, T* dfac) {
  const typename Mth<T>::mask neg = x < (T)0;  // NaN falls through and propagates
  const T xc = Mth<T>::sel(neg, (T)0, x);
  const T k = (T)0.5 * Mth<T>::rsqrt(xc);      // +inf at x == 0
  *dfac = Mth<T>::sel(neg, (T)0, k);
  return Mth<T>::sqrt(xc);
}

// post map ref:24-39: f = log1p | expm1 | nlog | identity, then tau >= 1 ->
// 1 - tau/(tau+f) = f/(tau+f).  Returns value, multiplies *dfac by d out / d in.
template <typename T, bool FAST>
GD_HD T post_map(T d, const PairParams<T>& P, T* dfac, typename Mth<T>::mask* rare) {
  T f = d, df = (T)1;
  if (P.fun == kFunLog1p) {
    if constexpr (FAST) {
      *rare |= !(d < (T)1e30);                 // inf / nan distance: robust path
      if constexpr (std::is_same<T, float>::value) {
        f = P.lean ? Mth<float>::log1p_lean(d) : Mth<float>::log1p_pos(d);
      } else {
        f = Mth<T>::log1p_pos(d);
      }
    } else {
      f = Mth<T>::log1p(d);                    // ref:26
    }
    df = Mth<T>::rcp((T)1 + d);
  } else if (P.fun == kFunExpm1) {
    f = Mth<T>::expm1(d);                      // ref:28
    df = f + (T)1;
  } else if (P.fun == kFunNlog) {
    // -log(1 - d + 1e-7) (ref:30) as -log1p(1e-7 - d): no rounding of 1 - d for small d
    f = -Mth<T>::log1p((T)1e-7 - d);
    df = Mth<T>::rcp((T)1 - d + (T)1e-7);
  }
  if (P.tau_on) {                              // ref:36-37
    T inv = Mth<T>::rcp(P.tau + f);
    df = df * P.tau * inv * inv;
    f = f * inv;
  }
  *dfac = *dfac * df;
  return f;
}

// ---------------------------------------------------------------------------
// a1: the two boxes reduced to what every distance needs         ref:8-21
// ---------------------------------------------------------------------------
template <typename T>
struct PairGeom {
  T dx, dy, dz;             // c_p - c_t, centre = xyz + off * UNCLAMPED extents (ref:12)
  T ap, bp, ep;             // pred half extents (clamped, ref:13-14, x0.5 ref:19-20)
  T at, bt, et;             // target half extents
  T ma, mb, me;             // clamp gradient masks of the pred extents
  T sp, cp;                 // sin / cos of pred yaw              (ref:16-17)
  T sd, cd;                 // sin / cos of (r_p - r_t)
  // Quantities that depend on ONE box only.  The element-wise path derives them per
  // pair (geom_derive, same arithmetic as before); the pairwise path computes them
  // once per box (BoxGauss) so the N x M inner loop does not repeat them.
  T A, B, E, C, D, F;       // squared half extents (pred: A B E, target: C D F)
  T abp, abt;               // a_p b_p, a_t b_t                    (ref:91-92)
  T amb, cmd;               // A - B, C - D formed as (a-b)(a+b)
  T iap, ibp, iep;          // 1 / pred half extents
  T iat, ibt, iet;          // 1 / target half extents
  T r6p, r6t;               // (a b e)^(-1/6) per box, valid iff has_r6
  bool has_r6;
};

template <typename T>
GD_HD void geom_derive(PairGeom<T>* g) {
  g->A = g->ap * g->ap;
  g->B = g->bp * g->bp;
  g->E = g->ep * g->ep;
  g->C = g->at * g->at;
  g->D = g->bt * g->bt;
  g->F = g->et * g->et;
  g->abp = g->ap * g->bp;
  g->abt = g->at * g->bt;
  g->amb = (g->ap - g->bp) * (g->ap + g->bp);
  g->cmd = (g->at - g->bt) * (g->at + g->bt);
  g->iap = Mth<T>::rcp(g->ap);
  g->ibp = Mth<T>::rcp(g->bp);
  g->iep = Mth<T>::rcp(g->ep);
  g->iat = Mth<T>::rcp(g->at);
  g->ibt = Mth<T>::rcp(g->bt);
  g->iet = Mth<T>::rcp(g->et);
  g->r6p = g->r6t = (T)0;
  g->has_r6 = false;
}

// FAST: the caller promises to re-run the row on the robust path when *rare is
// set.  A row is "nice" when all six extents lie in [1e-4, 1e4] (clamps inactive,
// gradient masks 1, no overflow in products of ratios) and both yaws are within
// +-16 rad (so r_p - r_t is formed with an absolute error below 1e-6 rad and the
// branch-free range reduction is exact); NaNs fail every test and land on the
// robust path too.
template <typename T, bool NEED_PRED_ROT, bool FAST, int DIET = 0>
GD_HD PairGeom<T> make_geom(const T* p, const T* t, const PairParams<T>& P,
                            typename Mth<T>::mask* rare) {
  PairGeom<T> g;
  if constexpr ((DIET & kDietStd) != 0) {      // center_offset == (0, 0, 0.5)
    g.dx = p[0] - t[0];
    g.dy = p[1] - t[1];
    g.dz = (p[2] - t[2]) + (T)0.5 * (p[5] - t[5]);
  } else {
  g.dx = (p[0] - t[0]) + P.off[0] * (p[3] - t[3]);
  g.dy = (p[1] - t[1]) + P.off[1] * (p[4] - t[4]);
  g.dz = (p[2] - t[2]) + P.off[2] * (p[5] - t[5]);
  }
  if constexpr (FAST) {
    if constexpr ((DIET & kDietGuards) != 0) {
      *rare |= !Mth<T>::nice_row_minmax(p[3], p[4], p[5], t[3], t[4], t[5], p[6], t[6]);
    } else {
    const T lo = (T)1e-4, hi = (T)1e4;
    typename Mth<T>::mask ok = p[3] >= lo && p[4] >= lo && p[5] >= lo && t[3] >= lo && t[4] >= lo && t[5] >= lo;
    ok = ok && p[3] <= hi && p[4] <= hi && p[5] <= hi && t[3] <= hi && t[4] <= hi && t[5] <= hi;
    const T ymax = (T)16;
    ok = ok && (p[6] >= -ymax && p[6] <= ymax && t[6] >= -ymax && t[6] <= ymax);
    *rare |= !ok;
    }
    g.ap = (T)0.5 * p[3];
    g.bp = (T)0.5 * p[4];
    g.ep = (T)0.5 * p[5];
    g.at = (T)0.5 * t[3];
    g.bt = (T)0.5 * t[4];
    g.et = (T)0.5 * t[5];
    g.ma = g.mb = g.me = (T)1;
    Mth<T>::sincos_fast(p[6] - t[6], &g.sd, &g.cd);
    if (NEED_PRED_ROT) {
      Mth<T>::sincos_fast(p[6], &g.sp, &g.cp);
    } else {
      g.sp = (T)0;
      g.cp = (T)1;
    }
    geom_derive(&g);
    return g;
  } else {
  T dummy;
  g.ap = (T)0.5 * clamp_extent(p[3], &g.ma);
  g.bp = (T)0.5 * clamp_extent(p[4], &g.mb);
  g.ep = (T)0.5 * clamp_extent(p[5], &g.me);
  g.at = (T)0.5 * clamp_extent(t[3], &dummy);
  g.bt = (T)0.5 * clamp_extent(t[4], &dummy);
  g.et = (T)0.5 * clamp_extent(t[5], &dummy);
  const T ybig = (T)16;
  if (p[6] >= -ybig && p[6] <= ybig && t[6] >= -ybig && t[6] <= ybig) {
    // ordinary yaws: the difference is formed first so that sin(r_p - r_t) keeps its
    // RELATIVE accuracy when the boxes are nearly parallel
    Mth<T>::sincos(p[6] - t[6], &g.sd, &g.cd);
    if (NEED_PRED_ROT) {
      Mth<T>::sincos(p[6], &g.sp, &g.cp);
    } else {
      g.sp = (T)0;
      g.cp = (T)1;
    }
  } else {
    // huge yaws: r_p - r_t would lose its low bits in float, so take sin/cos of
    // each angle (as the reference does, ref:16-17) and use the difference identities
    T st, ct;
    Mth<T>::sincos(p[6], &g.sp, &g.cp);
    Mth<T>::sincos(t[6], &st, &ct);
    g.sd = g.sp * ct - g.cp * st;
    g.cd = g.cp * ct + g.sp * st;
  }
  geom_derive(&g);
  return g;
  }
}

// Local gradient (before the outer chain factor): w.r.t. centre difference in
// the global frame, the three pred half extents and the pred yaw.
template <typename T>
struct LocalGrad {
  T gdx, gdy, gdz, ga, gb, ge, gr;
};

template <typename T, int DIET = 0>
GD_HD void store_grad(const PairGeom<T>& g, const PairParams<T>& P,
                      const LocalGrad<T>& L, T fac, T* out) {
  // d/dw = 0.5 * [1e-7 <= w <= 1e7] * d/da + off_x * d/dc_x   
  out[0] = fac * L.gdx;
  out[1] = fac * L.gdy;
  out[2] = fac * L.gdz;
  if constexpr ((DIET & kDietStd) != 0) {      // center_offset == (0, 0, 0.5)
    out[3] = fac * ((T)0.5 * g.ma * L.ga);
    out[4] = fac * ((T)0.5 * g.mb * L.gb);
    out[5] = fac * ((T)0.5 * g.me * L.ge + (T)0.5 * L.gdz);
    out[6] = fac * L.gr;
    return;
  }
  out[3] = fac * ((T)0.5 * g.ma * L.ga + P.off[0] * L.gdx);
  out[4] = fac * ((T)0.5 * g.mb * L.gb + P.off[1] * L.gdy);
  out[5] = fac * ((T)0.5 * g.me * L.ge + P.off[2] * L.gdz);
  out[6] = fac * L.gr;
}

// ---------------------------------------------------------------------------
// a2: GWD                                                        ref:42-106
// ---------------------------------------------------------------------------
template <typename T, bool GRAD, bool FAST, int DIET = 0>
GD_HD T gwd_core(const PairGeom<T>& g, const PairParams<T>& P, T gscale, T* grad, typename Mth<T>::mask* rare) {
  const T A = g.A, B = g.B, C = g.C, D = g.D;
  const T s2 = g.sd * g.sd;
  const T K = g.abp * g.abt;                                       // ref:91-92
  const T V = g.ap * g.at + g.bp * g.bt;
  const T amb = g.amb;                                             // A - B
  const T cmd = g.cmd;                                             // C - D
  const T eps = amb * cmd * s2;                                    // V^2 - U
  const T c2 = g.cd * g.cd;
  // U = tr(Sigma_p Sigma_t) + 2 sqrt(det det): all terms positive    ref:88-95
  const T U = (A * C + B * D) * c2 + (A * D + B * C) * s2 + (T)2 * K;
  T kU;                                                            // 1/(2 sqrt U)
  const T rU = sqrt_clamp0(U, &kU);                                // ref:95
  const T eta = eps * Mth<T>::rcp(V + rU);                         // V - sqrt U
  const T da = g.ap - g.at, db = g.bp - g.bt, de = g.ep - g.et;
  const T W = da * da + db * db + (T)2 * eta + de * de;            // ref:81-97
  const T d2 = g.dx * g.dx + g.dy * g.dy + g.dz * g.dz + a2_times<DIET>(P, W);  // ref:79,99
  T k;                                                             // 1/(2 d)
  T d = sqrt_clamp0(d2, &k);                                       // ref:99
  T inv_n = (T)1;
  if (P.flag) {                                                    // ref:101-104
    // 1 / (2 (K e_p e_t)^(1/6)), product split so it cannot overflow
    if (g.has_r6) {                            // pairwise: per-box factors
      inv_n = (T)0.5 * (g.r6p * g.r6t);
    } else if constexpr (FAST) {
      inv_n = (T)0.5 * Mth<T>::rsixthroot((g.abp * g.ep) * (g.abt * g.et));
    } else {
      const T vp = Mth<T>::sqrt(g.ap * g.bp * g.ep);
      const T vt = Mth<T>::sqrt(g.at * g.bt * g.et);
      inv_n = (T)0.5 * Mth<T>::rcbrt(vp * vt);
    }
  }
  const T gval = d * inv_n;
  T fac = (T)1;
  const T out = post_map<T, FAST>(gval, P, &fac, rare);
  if (GRAD) {
    const T irU = (T)2 * kU;                                       // 1/sqrt U
    const T rot = cmd * s2;                                        // (C-D) s2
    // dW/da_p = [2 (a_p-a_t) V - 2 a_p (eta - (C-D) s2)] / sqrt U  
    const T dWa = (T)2 * (da * V - g.ap * (eta - rot)) * irU;
    const T dWb = (T)2 * (db * V - g.bp * (eta + rot)) * irU;
    const T dWr = amb * cmd * ((T)2 * g.sd * g.cd) * irU;
    const T kn = k * inv_n;                                        // 1/(2 d n)
    LocalGrad<T> L;
    L.gdx = (T)2 * g.dx * kn;
    L.gdy = (T)2 * g.dy * kn;
    L.gdz = (T)2 * g.dz * kn;
    L.ga = a2_times<DIET>(P, dWa) * kn;
    L.gb = a2_times<DIET>(P, dWb) * kn;
    L.ge = a2_times<DIET>(P, (T)2) * de * kn;
    L.gr = a2_times<DIET>(P, dWr) * kn;
    if (P.flag) {                                                  // d ln n / da = 1/(6a)
      const T g6 = gval * (T)(1.0 / 6.0);
      L.ga -= g6 * g.iap;
      L.gb -= g6 * g.ibp;
      L.ge -= g6 * g.iep;
    }
    store_grad<T, DIET>(g, P, L, fac * gscale, grad);
  }
  return out;
}

// ---------------------------------------------------------------------------
// a3: KLD, both directions, in the pred frame                   ref:109-141
// ---------------------------------------------------------------------------
// Shared pieces of KL(t||p) ("fwd", what kld3d_loss(pred, target) computes:
// it inverts Sigma_p, ref:114-116) and KL(p||t) ("rev", kld3d_loss(target,
// pred), needed by jd / symmax / symmin, ref:193,207,220).
template <typename T>
struct KldCommon {
  T u, v;            // centre difference rotated into the pred frame   (ref:119-123)
  T iA, iB, iE;      // 1/A, 1/B, 1/E
  T cmd;             // C - D
  T s2, s2x2;        // sin^2 dl, sin(2 dl)
};

template <typename T, bool FAST, int DIET = 0>
GD_HD T kld_fwd(const PairGeom<T>& g, const PairParams<T>& P, bool want_grad,
                LocalGrad<T>* L, T* ul, T* vl, typename Mth<T>::mask* rare) {
  // value: 0.5 (u^2/A + v^2/B + dz^2/E)/alpha^2 + 0.5 tr(Sp^-1 St) + 0.5 F/E
  //        + ln(a_p b_p e_p / a_t b_t e_t) - 1.5                   ref:122-137
  const T u = g.cp * g.dx + g.sp * g.dy;
  const T v = -g.sp * g.dx + g.cp * g.dy;
  const T iap = g.iap, ibp = g.ibp, iep = g.iep;
  const T iA = iap * iap, iB = ibp * ibp, iE = iep * iep;
  const T s2 = g.sd * g.sd;
  const T cmd = g.cmd;
  // delta_i = (t_i - p_i)/p_i ; C/A = (1+delta_a)^2 ...
  const T qa = (g.at - g.ap) * iap, qb = (g.bt - g.bp) * ibp, qe = (g.et - g.ep) * iep;
  const T ra = g.at * iap, rb = g.bt * ibp, re = g.et * iep;       // = 1 + q
  const T maha = times_ia2<DIET>((T)0.5 * (u * u * iA + v * v * iB + g.dz * g.dz * iE), P);
  // sum(delta + delta^2/2) - log((1+da)(1+db)(1+de)) with a single log:
  const T pair = qa * qb + qa * qe + qb * qe + qa * qb * qe;       // Pi(1+d) - 1 - sum d
  const T shape = (T)0.5 * (qa * qa + qb * qb + qe * qe)
      + Mth<T>::template sum_minus_log_ratios<FAST>(qa + qb + qe, pair, ra, rb, re, rare)
      + (T)0.5 * cmd * s2 * (iB - iA);
  if (want_grad) {
    const T s2x2 = (T)2 * g.sd * g.cd;
    const T rot = cmd * s2;
    const T lu = times_ia2<DIET>(u * iA, P), lv = times_ia2<DIET>(v * iB, P);
    *ul = lu;                                  // gradient w.r.t. (u, v): pred frame
    *vl = lv;
    L->gdz = times_ia2<DIET>(g.dz * iE, P);
    L->ga = iap * (-qa * ((T)2 + qa) + rot * iA - times_ia2<DIET>(u * u * iA, P));
    L->gb = ibp * (-qb * ((T)2 + qb) - rot * iB - times_ia2<DIET>(v * v * iB, P));
    L->ge = iep * (-qe * ((T)2 + qe) - times_ia2<DIET>(g.dz * g.dz * iE, P));
    L->gr = (iA - iB) * (times_ia2<DIET>(u * v, P) - (T)0.5 * cmd * s2x2);
  }
  return maha + shape;
}

template <typename T, bool FAST, int DIET = 0>
GD_HD T kld_rev(const PairGeom<T>& g, const PairParams<T>& P, bool want_grad,
                LocalGrad<T>* L, T* ul, T* vl, typename Mth<T>::mask* rare) {
  // KL with Sigma_t inverted (kld3d_loss(target, pred)); gradient still w.r.t. pred.
  const T u = g.cp * g.dx + g.sp * g.dy;
  const T v = -g.sp * g.dx + g.cp * g.dy;
  // rotate into the target frame: (u', v') = R(dl) (u, v)
  const T ut = g.cd * u - g.sd * v;
  const T vt = g.sd * u + g.cd * v;
  const T iat = g.iat, ibt = g.ibt, iet = g.iet;
  const T iC = iat * iat, iD = ibt * ibt, iF = iet * iet;
  const T s2 = g.sd * g.sd;
  const T amb = g.amb;
  const T qa = (g.ap - g.at) * iat, qb = (g.bp - g.bt) * ibt, qe = (g.ep - g.et) * iet;
  const T ra = g.ap * iat, rb = g.bp * ibt, re = g.ep * iet;       // = 1 + q
  const T maha = times_ia2<DIET>((T)0.5 * (ut * ut * iC + vt * vt * iD + g.dz * g.dz * iF), P);
  const T pair = qa * qb + qa * qe + qb * qe + qa * qb * qe;
  const T shape = (T)0.5 * (qa * qa + qb * qb + qe * qe)
      + Mth<T>::template sum_minus_log_ratios<FAST>(qa + qb + qe, pair, ra, rb, re, rare)
      + (T)0.5 * amb * s2 * (iD - iC);
  if (want_grad) {
    const T s2x2 = (T)2 * g.sd * g.cd;
    const T A = g.A, B = g.B;
    // gradient w.r.t. (u', v') rotated back into the pred frame: R(-dl)
    const T lut = times_ia2<DIET>(ut * iC, P), lvt = times_ia2<DIET>(vt * iD, P);
    *ul = g.cd * lut + g.sd * lvt;
    *vl = -g.sd * lut + g.cd * lvt;
    L->gdz = times_ia2<DIET>(g.dz * iF, P);
    L->ga = g.iap * (qa * ((T)2 + qa) + A * s2 * (iD - iC));
    L->gb = g.ibp * (qb * ((T)2 + qb) + B * s2 * (iC - iD));
    L->ge = g.iep * (qe * ((T)2 + qe));
    L->gr = (T)0.5 * amb * s2x2 * (iD - iC);
  }
  return maha + shape;
}

template <typename T>
GD_HD void rotate_centre_grad(const PairGeom<T>& g, T ul, T vl, LocalGrad<T>* L) {
  // pred frame -> global frame: R(r_p)
  L->gdx = g.cp * ul - g.sp * vl;
  L->gdy = g.sp * ul + g.cp * vl;
}

template <typename T, int LOSS, bool GRAD, bool FAST, int DIET = 0>
GD_HD T kld_family_core(const PairGeom<T>& g, const PairParams<T>& P, T gscale, T* grad,
                        typename Mth<T>::mask* rare) {
  LocalGrad<T> L;
  T ul = (T)0, vl = (T)0;
  T fac = (T)1;
  T val;
  if constexpr (LOSS == kKld) {
    val = kld_fwd<T, FAST, DIET>(g, P, GRAD, &L, &ul, &vl, rare);
    if (P.flag) {                                                  // ref:138-139
      T k;
      val = sqrt_clamp0(val, &k);
      fac = k;
    }
  } else {
    LocalGrad<T> Lr;
    T ulr = (T)0, vlr = (T)0;
    T f = kld_fwd<T, FAST, DIET>(g, P, GRAD, &L, &ul, &vl, rare);
    T r = kld_rev<T, FAST, DIET>(g, P, GRAD, &Lr, &ulr, &vlr, rare);
    T wf, wr;                                  // d val / d f, d val / d r
    if constexpr (LOSS == kJd) {                                   // ref:191-197
      val = (T)0.5 * (f + r);
      wf = wr = (T)0.5;
      if (P.flag) {
        T k;
        val = sqrt_clamp0(val, &k);
        wf *= k;
        wr *= k;
      }
    } else {                                                       // ref:204-223
      T kf = (T)1, kr = (T)1;
      if (P.flag) {
        f = sqrt_clamp0(f, &kf);
        r = sqrt_clamp0(r, &kr);
      }
      // torch.max / torch.min backward: exact ties split 1/2 - 1/2 (appendix A)
      const bool take_f = (LOSS == kSymMax) ? (f > r) : (f < r);
      const bool tie = (f == r);
      val = tie ? f : (take_f ? f : r);
      if (f != f || r != r) val = f + r;       // NaN propagates like torch.max/min
      wf = tie ? (T)0.5 * kf : (take_f ? kf : (T)0);
      wr = tie ? (T)0.5 * kr : (take_f ? (T)0 : kr);
    }
    if (GRAD) {
      ul = wf * ul + wr * ulr;
      vl = wf * vl + wr * vlr;
      L.gdz = wf * L.gdz + wr * Lr.gdz;
      L.ga = wf * L.ga + wr * Lr.ga;
      L.gb = wf * L.gb + wr * Lr.gb;
      L.ge = wf * L.ge + wr * Lr.ge;
      L.gr = wf * L.gr + wr * Lr.gr;
    }
  }
  const T out = post_map<T, FAST>(val, P, &fac, rare);
  if (GRAD) {
    rotate_centre_grad(g, ul, vl, &L);
    store_grad<T, DIET>(g, P, L, fac * gscale, grad);
  }
  return out;
}

// ---------------------------------------------------------------------------
// a4: Bhattacharyya                                             ref:144-186
// ---------------------------------------------------------------------------
template <typename T, bool GRAD, bool FAST, int DIET = 0>
GD_HD T bd_core(const PairGeom<T>& g, const PairParams<T>& P, T gscale, T* grad, typename Mth<T>::mask* rare) {
  const T A = g.A, B = g.B, C = g.C, D = g.D;
  const T E = g.E, F = g.F;
  const T s2 = g.sd * g.sd, c2 = g.cd * g.cd, sc = g.sd * g.cd;
  const T u = g.cp * g.dx + g.sp * g.dy;
  const T v = -g.sp * g.dx + g.cp * g.dy;
  const T amb = g.amb;
  const T cmd = g.cmd;
  // Sigma_t in the pred frame
  const T t00 = C * c2 + D * s2, t11 = C * s2 + D * c2, t01 = -cmd * sc;
  // M = (Sigma_p + Sigma_t)/2 in the pred frame                    ref:152
  const T M00 = (T)0.5 * (A + t00), M11 = (T)0.5 * (B + t11), M01 = (T)0.5 * t01;
  const T Ml = (T)0.5 * (E + F);                                   // ref:153
  const T eps = amb * cmd * s2;
  const T detN = (A + C) * (B + D) + eps;                          // det(Sp+St)
  const T det_raw = (T)0.25 * detN;                                // ref:155-157
  const typename Mth<T>::mask clamp_m = !(det_raw >= (T)1e-7);     // ref:158
  bool clamped = false;
  if constexpr (FAST) {                    // clamp active (or nan): robust path redoes the row
    *rare |= clamp_m;
  } else {
    clamped = clamp_m;
  }
  const T det = clamped ? (T)1e-7 : det_raw;
  const T idet = Mth<T>::rcp(det);
  const T iMl = Mth<T>::rcp(Ml);
  const T Q2 = u * u * M11 - (T)2 * u * v * M01 + v * v * M00;     // d^T adj(M) d
  const T maha = times_ia2<DIET>((T)0.125 * (Q2 * idet + g.dz * g.dz * iMl), P);  // ref:170-172,182
  // shape: 0.5 ln det + 0.5 ln Ml - 0.25 ln(ABE) - 0.25 ln(CDF)    ref:174-180
  const T K = g.abp * g.abt;
  const T da = g.ap - g.at, db = g.bp - g.bt, de = g.ep - g.et;
  const T xe = de * de * ((T)0.5 * g.iep * g.iet);                 // Ml/(e_p e_t) - 1
  T shape;
  if (!clamped) {
    // det/(sqrt(AB) sqrt(CD)) = (1+xa)(1+xb) + eps/(4K),  xa = (a_p-a_t)^2/(2 a_p a_t)
    const T xa = da * da * ((T)0.5 * g.iap * g.iat);
    const T xb = db * db * ((T)0.5 * g.ibp * g.ibt);
    const T q = xa + xb + xa * xb + (T)0.25 * eps * ((g.iap * g.ibp) * (g.iat * g.ibt));
    const T qq = q + xe + q * xe;              // (1+q)(1+xe) - 1
    if constexpr (FAST) {
      *rare |= !(qq >= (T)0 && qq < (T)1e30);
      shape = (T)0.5 * Mth<T>::log1p_pos(qq);
    } else {
      shape = (qq < (T)1e37) ? (T)0.5 * Mth<T>::log1p(qq)
                             : (T)0.5 * (Mth<T>::log1p(q) + Mth<T>::log1p(xe));  // 1e7-vs-1e-7 extents
    }
  } else {
    shape = (T)0.5 * (Mth<T>::log((T)1e-7) - Mth<T>::log(K)) + (T)0.5 * Mth<T>::log1p(xe);
  }
  T val = maha + shape;
  T fac = (T)1;
  if (P.flag) {                                                    // ref:183-184
    T k;
    val = sqrt_clamp0(val, &k);
    fac = k;
  }
  const T out = post_map<T, FAST>(val, P, &fac, rare);
  if (GRAD) {
    const T c8 = times_ia2<DIET>((T)0.125, P) * idet;              // 1/(8 alpha^2 det)
    LocalGrad<T> L;
    // centre: M^-1 d / (4 alpha^2) in the pred frame, then rotate
    const T ul = (T)2 * c8 * (M11 * u - M01 * v);
    const T vl = (T)2 * c8 * (M00 * v - M01 * u);
    rotate_centre_grad(g, ul, vl, &L);
    L.gdz = times_ia2<DIET>((T)0.25 * g.dz * iMl, P);
    const T iap = g.iap, ibp = g.ibp, iep = g.iep;
    T sa, sb, gam;                             // shape parts and d/d det_raw factor
    if (!clamped) {
      const T idN = Mth<T>::rcp(detN);
      // a_p M11/det/2... - 1/(2 a_p) = [(A - t00)(B + t11) + t01^2] / (2 a_p detN).  Expanding
      // with t00 t11 - t01^2 = C D, t00 = C - (C-D) s2, t11 = D + (C-D) s2 gives
      //   (A - t00)(B + t11) + t01^2 = (A - C)(B + D) + (A + B)(C - D) s2
      //   (B - t11)(A + t00) + t01^2 = (B - D)(A + C) - (A + B)(C - D) s2
      // -- the same shape as detN.  No term of size C^2 appears (the product form cancels two
      // of them, which costs float32 every digit once the extents differ by ~10^3), and both
      // terms vanish separately as the boxes coincide.
      const T rot2 = (A + B) * (cmd * s2);
      sa = (T)0.5 * (da * (g.ap + g.at) * (B + D) + rot2) * idN * iap;
      sb = (T)0.5 * (db * (g.bp + g.bt) * (A + C) - rot2) * idN * ibp;
      gam = -Q2 * c8 * idet;                   // Mahalanobis part of d/d det
      L.gr = -amb * (u * v * c8 + (gam + (T)0.5 * idet) * M01);
    } else {
      sa = -(T)0.5 * iap;
      sb = -(T)0.5 * ibp;
      gam = (T)0;
      L.gr = -amb * (u * v * c8);
    }
    L.ga = g.ap * (v * v * c8 + gam * M11) + sa;
    L.gb = g.bp * (u * u * c8 + gam * M00) + sb;
    L.ge = times_ia2<DIET>(-(T)0.125 * g.dz * g.dz * iMl * iMl, P) * g.ep
        + (T)0.25 * de * (g.ep + g.et) * iep * iMl;
    store_grad<T, DIET>(g, P, L, fac * gscale, grad);
  }
  return out;
}

// ---------------------------------------------------------------------------
// a7: KFIoU (ignores centres, alpha, sqrt; tau forced 0)         ref:227-248
// ---------------------------------------------------------------------------
template <typename T, bool GRAD, bool FAST>
GD_HD T kfiou_core(const PairGeom<T>& g, const PairParams<T>& Pin, T gscale, T* grad,
                   typename Mth<T>::mask* rare) {
  PairParams<T> P = Pin;
  P.tau_on = 0;                                                    // ref:247
  const T A = g.ap * g.ap, B = g.bp * g.bp, C = g.at * g.at, D = g.bt * g.bt;
  const T E = g.ep * g.ep, F = g.et * g.et;
  const T s2 = g.sd * g.sd, c2 = g.cd * g.cd;
  const T amb = (g.ap - g.bp) * (g.ap + g.bp);
  const T cmd = (g.at - g.bt) * (g.at + g.bt);
  const T detN = (A + C) * (B + D) + amb * cmd * s2;               // ref:234-236
  const T Zraw = detN * (E + F);                                   // ref:237-238
  const bool zc = !(Zraw >= (T)1e-7);                              // ref:243
  const T Z = zc ? (T)1e-7 : Zraw;
  const T volp = g.ap * g.bp * g.ep, volt = g.at * g.bt * g.et;    // ref:240-241
  const T irZ = Mth<T>::rcp(Mth<T>::sqrt(Z));
  const T I = volp * volt * irZ;                                   // ref:243
  const T Uraw = volp + volt - I;
  const bool uc = !(Uraw >= (T)1e-7);                              // ref:245
  const T Uc = uc ? (T)1e-7 : Uraw;
  const T iU = Mth<T>::rcp(Uc);
  const T kf = I * iU;                                             // ref:246
  const T cK = (T)4.656854249492381;
  const T R0 = (T)5.656854249492381;           // 4 sqrt(2) = cK + 1
  if (!zc && !uc) {
    // No clamp active: 1 - cK I/U is a difference of nearly equal numbers when the
    // boxes almost coincide (I/U -> 1/cK).  With R = (vol_p+vol_t) sqrt(Z)/(vol_p vol_t)
    // = U/I + 1 one has  R = R0 sqrt(1+xi),  1+xi = (1+x1)(1+q)(1+xe)  with the small
    // non-negative-ish pieces below, and  1 - cK I/U = (R - R0)/(R - 1)
    //                                   = R0 xi / ((sqrt(1+xi) + 1)(R - 1)).
    const T K = (g.ap * g.bp) * (g.at * g.bt);
    const T da = g.ap - g.at, db = g.bp - g.bt, de = g.ep - g.et;
    const T i4K = (T)0.25 * Mth<T>::rcp(K);
    // vol_p - vol_t.  The telescoped form is exact-ish when the boxes nearly coincide (each
    // term is small), but its terms grow without bound when two extents move in opposite
    // directions at constant volume (w x 1000, h / 1000) and then cancel; the plain
    // difference has an error of ~ulp(vol) whatever the shapes.  Take whichever is bounded
    // better.
    const T tv1 = da * g.bp * g.ep, tv2 = g.at * (db * g.ep), tv3 = g.at * (g.bt * de);
    const T tmag = (tv1 < (T)0 ? -tv1 : tv1) + (tv2 < (T)0 ? -tv2 : tv2) + (tv3 < (T)0 ? -tv3 : tv3);
    const T dv = (tmag > (T)2 * (volp + volt)) ? (volp - volt) : (tv1 + (tv2 + tv3));
    const T i4vv = (T)0.25 * Mth<T>::rcp(volp * volt);
    const T x1 = dv * dv * i4vv;
    const T xa = da * da * Mth<T>::rcp((T)2 * g.ap * g.at);
    const T xb = db * db * Mth<T>::rcp((T)2 * g.bp * g.bt);
    const T xe = de * de * Mth<T>::rcp((T)2 * g.ep * g.et);
    const T q = xa + xb + xa * xb + amb * cmd * s2 * i4K;
    const T xi = x1 + q + xe + x1 * q + x1 * xe + q * xe + x1 * q * xe;
    if (xi < (T)1e30) {                        // else (1e7-vs-1e-7 extents): plain form below
    const T sq = Mth<T>::sqrt((T)1 + xi);
    const T Rm1 = R0 * sq - (T)1;
    const T iRm1 = Mth<T>::rcp(Rm1);
    T fac = (T)1;
    const T out = post_map<T, false>(R0 * xi * Mth<T>::rcp(sq + (T)1) * iRm1, P, &fac, rare);
    if (GRAD) {
      // d/dx = cK R0 sqrt(1+xi) / (2 (R-1)^2) * [dx1/(1+x1) + dq/(1+q) + dxe/(1+xe)]
      const T coef = (T)0.5 * cK * R0 * sq * iRm1 * iRm1;
      const T i1 = Mth<T>::rcp((T)1 + x1), iq = Mth<T>::rcp((T)1 + q), ie = Mth<T>::rcp((T)1 + xe);
      const T iap = Mth<T>::rcp(g.ap), ibp = Mth<T>::rcp(g.bp), iep = Mth<T>::rcp(g.ep);
      const T dx1v = dv * (volp + volt) * i4vv * i1;       // (dx1/dvol_p) vol_p / (1+x1)
      const T rot = cmd * s2 * (A + B);
      LocalGrad<T> L;
      L.gdx = L.gdy = L.gdz = (T)0;
      L.ga = coef * iap * (dx1v + ((B + D) * da * (g.ap + g.at) + rot) * i4K * iq);
      L.gb = coef * ibp * (dx1v + ((A + C) * db * (g.bp + g.bt) - rot) * i4K * iq);
      L.ge = coef * iep * (dx1v + (T)0.5 * de * (g.ep + g.et) * Mth<T>::rcp(g.ep * g.et) * ie);
      L.gr = coef * amb * cmd * ((T)2 * g.sd * g.cd) * i4K * iq;
      store_grad(g, P, L, fac * gscale, grad);
    }
    return out;
    }
  }
  T fac = (T)1;
  const T out = post_map<T, false>((T)1 - cK * kf, P, &fac, rare);  // ref:247
  if (GRAD) {
    // d ln I / dx = d ln volp / dx - 0.5 [Z unclamped] d ln Z / dx
    const T t00 = C * c2 + D * s2, t11 = C * s2 + D * c2;
    const T hz = zc ? (T)0 : (T)0.5 * Mth<T>::rcp(Z);
    const T iap = Mth<T>::rcp(g.ap), ibp = Mth<T>::rcp(g.bp), iep = Mth<T>::rcp(g.ep);
    const T EF = E + F;
    const T dIa = I * (iap - hz * EF * (T)2 * g.ap * (B + t11));
    const T dIb = I * (ibp - hz * EF * (T)2 * g.bp * (A + t00));
    const T dIe = I * (iep - hz * (T)2 * g.ep * detN);
    const T dIr = I * (-hz * EF * amb * cmd * ((T)2 * g.sd * g.cd));
    const T mu = uc ? (T)0 : (T)1;
    // k = I/U: dk = (dI U - I dU)/U^2, dU = mu (dvol - dI)
    const T c = -cK * iU;                      // d(1 - cK k)/dk * 1/U
    LocalGrad<T> L;
    L.gdx = L.gdy = L.gdz = (T)0;
    L.ga = c * (dIa - kf * mu * (volp * iap - dIa));
    L.gb = c * (dIb - kf * mu * (volp * ibp - dIb));
    L.ge = c * (dIe - kf * mu * (volp * iep - dIe));
    L.gr = c * (dIr + kf * mu * dIr);
    store_grad(g, P, L, fac * gscale, grad);
  }
  return out;
}

// ---------------------------------------------------------------------------
// dispatchers
// ---------------------------------------------------------------------------
template <typename T, int LOSS, bool GRAD, bool FAST, int DIET = 0>
GD_HD T core_eval(const PairGeom<T>& g, const PairParams<T>& P, T gscale, T* grad, typename Mth<T>::mask* rare) {
  if constexpr (LOSS == kGwd) return gwd_core<T, GRAD, FAST, DIET>(g, P, gscale, grad, rare);
  else if constexpr (LOSS == kBd) return bd_core<T, GRAD, FAST, DIET>(g, P, gscale, grad, rare);
  else if constexpr (LOSS == kKfiou) return kfiou_core<T, GRAD, FAST>(g, P, gscale, grad, rare);
  else return kld_family_core<T, LOSS, GRAD, FAST, DIET>(g, P, gscale, grad, rare);
}

// element-wise path, ROBUST version: one (pred row, target row) pair for ANY input
// (clamped / degenerate extents, huge yaws, nan): returns the loss value and
// writes grad[0..6] = gscale * d value / d pred (gscale folds weight * scale into
// the single chain-rule factor instead of seven extra multiplies).
template <typename T, int LOSS, bool GRAD>
GD_HD T pair_eval(const T* p, const T* t, const PairParams<T>& P, T gscale, T* grad) {
  constexpr bool kNeedRot = !(LOSS == kGwd || LOSS == kKfiou);
  bool unused = false;
  const PairGeom<T> g = make_geom<T, kNeedRot, false>(p, t, P, &unused);
  return core_eval<T, LOSS, GRAD, false>(g, P, gscale, grad, &unused);
}

// element-wise path, FAST version: straight-line code (no branch, no clamp, no
// slow path) so that several rows of one thread interleave in the scheduler.
// Identical formulas; *rare is set when the row is not "nice" (see make_geom) or
// hits a guard inside the distance -- the caller must then call pair_eval on it.
// kfiou3d has no fast variant (it is not a headline loss): always rare = true.
template <typename T, int LOSS, bool GRAD, int DIET = 0>
GD_HD T pair_eval_fast(const T* p, const T* t, const PairParams<T>& P, T gscale, T* grad,
                       typename Mth<T>::mask* rare) {
  constexpr bool kNeedRot = !(LOSS == kGwd || LOSS == kKfiou);
  if constexpr (LOSS == kKfiou) {
    *rare = true;
    return (T)0;
  } else {
    const PairGeom<T> g = make_geom<T, kNeedRot, true, DIET>(p, t, P, rare);
    return core_eval<T, LOSS, GRAD, true, DIET>(g, P, gscale, grad, rare);
  }
}

// Explicit round-to-nearest operations of the pairwise value path (see the long comment at
// namespace pw further down): mul.rn / add.rn / fma.rn are never re-fused by ptxas.
namespace pw {
GD_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
GD_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
GD_HD float sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
GD_HD float fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return ::fmaf(a, b, c);
#endif
}
}  // namespace pw

// ---------------------------------------------------------------------------
// a12: pairwise N x M path.  Everything that depends on ONE box is computed
// once per box (BoxGauss); the per-pair work is the geometry difference and the
// same value cores as above.  sin/cos of the yaw difference come from the
// angle-difference identities instead of a per-pair sincos.
// ---------------------------------------------------------------------------
// (16-byte aligned: the kernels read it from shared memory with 128-bit loads)
template <typename T>
struct alignas(16) BoxGauss {
  T cx, cy, cz;     // centre incl. center_offset * unclamped extents   ref:12
  T a, b, e;        // clamped half extents                             ref:13-14,19-20
  T s, c;           // sin / cos yaw                                    ref:16-17
  T A, B, E;        // squared half extents
  T ab;             // a b
  T amb;            // (a - b)(a + b)
  T r6;             // (a b e)^(-1/6): the box's factor of the gwd normaliser   ref:101-104
  T ia, ib, ie;     // reciprocal half extents
  T iA, iB, iE;     // their squares
  T iBmA;           // 1/B - 1/A
  T iab;            // 1 / (a b)
  int nice;         // extents in [1e-4, 1e4], aspect <= 256: the FAST cores may be used for this box
};

template <typename T>
GD_HD BoxGauss<T> box_gauss(const T* row, const PairParams<T>& P) {
  BoxGauss<T> b;
  T m;
  if constexpr (std::is_same<T, float>::value) {
    // float: every operation with its rounding spelled out, so that a box converts to the same
    // bits in every kernel that inlines this function (see namespace pw below)
    b.cx = pw::fma(P.off[0], row[3], row[0]);
    b.cy = pw::fma(P.off[1], row[4], row[1]);
    b.cz = pw::fma(P.off[2], row[5], row[2]);
    b.a = pw::mul(0.5f, clamp_extent(row[3], &m));
    b.b = pw::mul(0.5f, clamp_extent(row[4], &m));
    b.e = pw::mul(0.5f, clamp_extent(row[5], &m));
    Mth<T>::sincos(row[6], &b.s, &b.c);
    b.A = pw::mul(b.a, b.a);
    b.B = pw::mul(b.b, b.b);
    b.E = pw::mul(b.e, b.e);
    b.ab = pw::mul(b.a, b.b);
    b.amb = pw::mul(pw::sub(b.a, b.b), pw::add(b.a, b.b));
    b.r6 = Mth<T>::rcbrt(Mth<T>::sqrt(pw::mul(b.ab, b.e)));
    b.ia = Mth<T>::rcp(b.a);
    b.ib = Mth<T>::rcp(b.b);
    b.ie = Mth<T>::rcp(b.e);
    b.iA = pw::mul(b.ia, b.ia);
    b.iB = pw::mul(b.ib, b.ib);
    b.iE = pw::mul(b.ie, b.ie);
    b.iBmA = pw::sub(b.iB, b.iA);
    b.iab = pw::mul(b.ia, b.ib);
  } else {
  b.cx = row[0] + P.off[0] * row[3];
  b.cy = row[1] + P.off[1] * row[4];
  b.cz = row[2] + P.off[2] * row[5];
  b.a = (T)0.5 * clamp_extent(row[3], &m);
  b.b = (T)0.5 * clamp_extent(row[4], &m);
  b.e = (T)0.5 * clamp_extent(row[5], &m);
  Mth<T>::sincos(row[6], &b.s, &b.c);        // once per box: accurate version, any magnitude
  b.A = b.a * b.a;
  b.B = b.b * b.b;
  b.E = b.e * b.e;
  b.ab = b.a * b.b;
  b.amb = (b.a - b.b) * (b.a + b.b);
  b.r6 = Mth<T>::rcbrt(Mth<T>::sqrt(b.ab * b.e));   // finite for every clamped extent
  b.ia = Mth<T>::rcp(b.a);
  b.ib = Mth<T>::rcp(b.b);
  b.ie = Mth<T>::rcp(b.e);
  b.iA = b.ia * b.ia;
  b.iB = b.ib * b.ib;
  b.iE = b.ie * b.ie;
  b.iBmA = b.iB - b.iA;
  b.iab = b.ia * b.ib;
  }
  const T lo = (T)1e-4, hi = (T)1e4;
  // ... and an in-plane aspect ratio <= 256 (the short form of U in gwd_core's value path)
  b.nice = (row[3] >= lo && row[3] <= hi && row[4] >= lo && row[4] <= hi && row[5] >= lo &&
            row[5] <= hi && row[3] <= (T)256 * row[4] && row[4] <= (T)256 * row[3]) ? 1 : 0;
  return b;
}

template <typename T>
GD_HD PairGeom<T> geom_from_gauss(const BoxGauss<T>& p, const BoxGauss<T>& t) {
  PairGeom<T> g;
  g.dx = p.cx - t.cx;
  g.dy = p.cy - t.cy;
  g.dz = p.cz - t.cz;
  g.ap = p.a; g.bp = p.b; g.ep = p.e;
  g.at = t.a; g.bt = t.b; g.et = t.e;
  g.ma = g.mb = g.me = (T)0;
  g.sp = p.s; g.cp = p.c;
  g.sd = p.s * t.c - p.c * t.s;       // sin(r_p - r_t)
  g.cd = p.c * t.c + p.s * t.s;       // cos(r_p - r_t)
  g.A = p.A; g.B = p.B; g.E = p.E;
  g.C = t.A; g.D = t.B; g.F = t.E;
  g.abp = p.ab; g.abt = t.ab;
  g.amb = p.amb; g.cmd = t.amb;
  g.iap = p.ia; g.ibp = p.ib; g.iep = p.ie;
  g.iat = t.ia; g.ibt = t.ib; g.iet = t.ie;
  g.r6p = p.r6; g.r6t = t.r6;
  g.has_r6 = true;
  return g;
}

// robust value (any input)
template <typename T, int LOSS>
GD_HD T pair_value(const BoxGauss<T>& p, const BoxGauss<T>& t, const PairParams<T>& P) {
  const PairGeom<T> g = geom_from_gauss(p, t);
  bool unused = false;
  return core_eval<T, LOSS, false, false>(g, P, (T)1, (T*)0, &unused);
}

// ---------------------------------------------------------------------------
// Pairwise value path with the rounding of every operation FIXED IN THE SOURCE.
//
// The matrix kernel, the fused-reduction kernels and the top-k kernel are different
// instantiations, with the two boxes in different roles (registers / shared memory, loop
// invariant / variant).  Left to the compiler, the same formula is contracted into FMAs
// differently from one context to the next (a product that is loop invariant or shared with the
// cold robust path stays a separate multiply), and the values differ in the last place -- so an
// index derived in one kernel need not be the index of the matrix written by another.  Here
// every multiply, add and FMA is an explicit round-to-nearest intrinsic (mul.rn / add.rn /
// fma.rn are never re-fused by ptxas) and the MUFU approximations are functions of their input
// bits: the value of a pair is the same in every kernel BY CONSTRUCTION, which is what
// "assignment indices bit-exact with the matrix" rests on.
// ---------------------------------------------------------------------------
namespace pw {
// Mth<float>::log1p_lean, operation by operation
GD_HD float log1p_lean(float x) {
  const float s = mul(x, Mth<float>::rcp(add(2.0f, x)));
  const float z = mul(s, s);
  float q = fma(z, 1.0f / 9.0f, 1.0f / 7.0f);
  q = fma(z, q, 1.0f / 5.0f);
  q = fma(z, q, 1.0f / 3.0f);
  const float t = add(s, s);
  const float small = fma(mul(t, z), q, t);
#if defined(__CUDA_ARCH__) && !GD_PRECISE_MATH
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(add(1.0f, x)));
  const float big = mul(l, 0.693147180559945f);
#else
  const float big = ::log1pf(x);
#endif
  return x <= 0.5f ? small : big;
}
// Mth<float>::log1p_pos, operation by operation (log(1+x), 0 <= x < ~1e30, ~1 ulp)
GD_HD float log1p_pos(float x) {
  const float y = add(1.0f, x);
  uint32_t yb;
  memcpy(&yb, &y, 4);
  const int k = (int)((int32_t)(yb - 0x3f3504f3u) >> 23);
  const uint32_t mb = yb - ((uint32_t)k << 23);
  float m;
  memcpy(&m, &mb, 4);
  const float f = (k == 0) ? x : sub(m, 1.0f);
  const float s = mul(f, Mth<float>::rcp(add(2.0f, f)));
  const float z = mul(s, s);
  float q = 1.0f / 13.0f;
  q = fma(q, z, 1.0f / 11.0f);
  q = fma(q, z, 1.0f / 9.0f);
  q = fma(q, z, 1.0f / 7.0f);
  q = fma(q, z, 1.0f / 5.0f);
  q = fma(q, z, 1.0f / 3.0f);
  const float kf = (float)k;
  const float lo = fma(kf, 1.428606765330187e-06f, mul(mul(mul(2.0f, s), z), q));
  return fma(kf, 0.693145751953125f, fma(2.0f, s, lo));
}
// Mth<float>::sum_minus_log_ratios<FAST>, operation by operation:
// S - log(r1 r2 r3) with S = q1 + q2 + q3, r_i = 1 + q_i, pair = r1 r2 r3 - 1 - S
GD_HD float sum_minus_log_ratios(float S, float pair, float r1, float r2, float r3, bool* rare) {
  const float y = mul(mul(r1, r2), r3);
  *rare |= !(y > 1.0e-30f && y < 1.0e30f);
  uint32_t yb;
  memcpy(&yb, &y, 4);
  const int k = (int)((int32_t)(yb - 0x3f3504f3u) >> 23);
  const uint32_t mb = yb - ((uint32_t)k << 23);
  float m;
  memcpy(&m, &mb, 4);
  const float x = add(S, pair);
  const float f = (k == 0) ? x : sub(m, 1.0f);
  const float s = mul(f, Mth<float>::rcp(add(2.0f, f)));
  const float z = mul(s, s);
  float q = 1.0f / 13.0f;
  q = fma(q, z, 1.0f / 11.0f);
  q = fma(q, z, 1.0f / 9.0f);
  q = fma(q, z, 1.0f / 7.0f);
  q = fma(q, z, 1.0f / 5.0f);
  q = fma(q, z, 1.0f / 3.0f);
  const float tail = mul(mul(mul(2.0f, s), z), q);
  const float kf = (float)k;
  float r = fma(-kf, 0.693145751953125f, S);               // ln2 hi (k * hi exact)
  r = fma(-kf, 1.428606765330187e-06f, r);                 // ln2 lo
  const float direct = sub(fma(-2.0f, s, r), tail);
  return (k == 0) ? sub(fma(x, s, -tail), pair) : direct;
}
// sqrt(clamp(x, 0)): negative -> 0, NaN propagates                ref:95,99,139,184,197
GD_HD float sqrt_clamp0(float x) { return Mth<float>::sqrt(x < 0.0f ? 0.0f : x); }

// post map of the FAST value path (fun in {none, log1p}; anything else is left to the robust
// path), ref:24-39
GD_HD float post(float d, const PairParams<float>& P, bool* rare) {
  float f = d;
  if (P.fun == kFunLog1p) {
    *rare |= !(d < 1e30f);                     // inf / nan distance: robust path
    f = log1p_lean(d);
  } else if (P.fun != kFunNone) {
    *rare = true;
  }
  if (P.tau_on) f = mul(f, Mth<float>::rcp(add(P.tau, f)));       // 1 - tau/(tau+f)   ref:36-37
  return f;
}
// GWD (ref:42-106) for two NICE boxes (box_gauss: extents in [1e-4, 1e4], in-plane aspect
// ratio <= 256).  tr(Sp St) + 2K = (AC + BD) + 2K - (A-B)(C-D) s2 = V^2 - eps: under the
// aspect bound U >= 4K >= V^2 / 2^15, so the one subtraction keeps >= 9 bits more than the 1e-5
// of the result needs (error sweep against fp64: equal or better quantiles than the
// all-positive long form up to 1000:1, profiles/r03_pairwise.md), and U > 0: the clamp of
// ref:95 could only act on a NaN, which the root propagates anyway.
GD_HD float gwd_value(const BoxGauss<float>& p, const BoxGauss<float>& t,
                      const PairParams<float>& P, bool* rare) {
  const float dx = sub(p.cx, t.cx), dy = sub(p.cy, t.cy), dz = sub(p.cz, t.cz);
  const float sd = fma(p.s, t.c, -mul(p.c, t.s));                  // sin(r_p - r_t)
  const float s2 = mul(sd, sd);
  const float V = fma(p.a, t.a, mul(p.b, t.b));
  const float eps = mul(mul(p.amb, t.amb), s2);                    // V^2 - U
  const float U = fma(V, V, -eps);
  const float rU = Mth<float>::sqrt(U);                            // ref:95
  const float eta = mul(eps, Mth<float>::rcp(add(V, rU)));         // V - sqrt U
  const float da = sub(p.a, t.a), db = sub(p.b, t.b), de = sub(p.e, t.e);
  float W = fma(da, da, mul(db, db));                              // ref:81-97
  W = fma(2.0f, eta, W);
  W = fma(de, de, W);
  float d2 = fma(dx, dx, mul(dy, dy));                             // ref:79,99
  d2 = fma(dz, dz, d2);
  d2 = fma(P.alpha2, W, d2);
  float d = sqrt_clamp0(d2);                                       // NaN falls through
  if (P.flag) d = mul(d, mul(0.5f, mul(p.r6, t.r6)));              // ref:101-104
  return post(d, P, rare);
}
// KL divergence with Sigma_p inverted -- what kld3d_loss(pred = p, target = t) evaluates
// (ref:109-141) before its optional sqrt; the reverse direction of jd / symmax / symmin
// (kld3d_loss(target, pred), ref:193,207,220) is the same function with the boxes swapped.
// Same formulas as kld_fwd above (ratios q = (t - p)/p, one log through
// sum_minus_log_ratios), every operation explicit.
GD_HD float kld_dir(const BoxGauss<float>& p, const BoxGauss<float>& t,
                    const PairParams<float>& P, bool* rare) {
  const float dx = sub(p.cx, t.cx), dy = sub(p.cy, t.cy), dz = sub(p.cz, t.cz);
  const float u = fma(p.c, dx, mul(p.s, dy));                      // ref:119-123
  const float v = fma(p.c, dy, -mul(p.s, dx));
  const float sd = fma(p.s, t.c, -mul(p.c, t.s));
  const float s2 = mul(sd, sd);
  const float qa = mul(sub(t.a, p.a), p.ia), qb = mul(sub(t.b, p.b), p.ib),
              qe = mul(sub(t.e, p.e), p.ie);
  const float ra = mul(t.a, p.ia), rb = mul(t.b, p.ib), re = mul(t.e, p.ie);   // = 1 + q
  float mh = mul(mul(u, u), p.iA);
  mh = fma(mul(v, v), p.iB, mh);
  mh = fma(mul(dz, dz), p.iE, mh);
  const float maha = mul(mul(0.5f, P.inv_alpha2), mh);             // ref:122-124,137
  const float ab = mul(qa, qb);
  float pr = fma(qa, qe, ab);                                      // Pi(1+q) - 1 - sum q
  pr = fma(qb, qe, pr);
  pr = fma(ab, qe, pr);
  float h = mul(qa, qa);
  h = fma(qb, qb, h);
  h = fma(qe, qe, h);
  const float S = add(add(qa, qb), qe);
  const float lg = sum_minus_log_ratios(S, pr, ra, rb, re, rare);
  const float rot = mul(mul(mul(0.5f, t.amb), s2), p.iBmA);        // 0.5 (C-D) s2 (1/B - 1/A)
  const float shape = add(fma(0.5f, h, lg), rot);
  return add(maha, shape);
}
template <int LOSS>
GD_HD float kld_family_value(const BoxGauss<float>& p, const BoxGauss<float>& t,
                             const PairParams<float>& P, bool* rare) {
  float val;
  if constexpr (LOSS == kKld) {
    val = kld_dir(p, t, P, rare);
    if (P.flag) val = sqrt_clamp0(val);                            // ref:138-139
  } else {
    float f = kld_dir(p, t, P, rare);
    float r = kld_dir(t, p, P, rare);
    if constexpr (LOSS == kJd) {                                   // ref:191-197
      val = mul(0.5f, add(f, r));
      if (P.flag) val = sqrt_clamp0(val);
    } else {                                                       // ref:204-223
      if (P.flag) {
        f = sqrt_clamp0(f);
        r = sqrt_clamp0(r);
      }
      const bool take_f = (LOSS == kSymMax) ? (f > r) : (f < r);
      val = (f == r) ? f : (take_f ? f : r);
      if (f != f || r != r) val = add(f, r);   // NaN propagates like torch.max / torch.min
    }
  }
  return post(val, P, rare);
}
// Bhattacharyya (ref:144-186), same formulas as bd_core above, every operation explicit.
GD_HD float bd_value(const BoxGauss<float>& p, const BoxGauss<float>& t,
                     const PairParams<float>& P, bool* rare) {
  const float dx = sub(p.cx, t.cx), dy = sub(p.cy, t.cy), dz = sub(p.cz, t.cz);
  const float u = fma(p.c, dx, mul(p.s, dy));
  const float v = fma(p.c, dy, -mul(p.s, dx));
  const float sd = fma(p.s, t.c, -mul(p.c, t.s));                  // sin / cos (r_p - r_t)
  const float cd = fma(p.c, t.c, mul(p.s, t.s));
  const float s2 = mul(sd, sd), c2 = mul(cd, cd), sc = mul(sd, cd);
  // Sigma_t in the pred frame; M = (Sigma_p + Sigma_t)/2              ref:152
  const float t00 = fma(t.A, c2, mul(t.B, s2)), t11 = fma(t.A, s2, mul(t.B, c2));
  const float t01 = -mul(t.amb, sc);
  const float M00 = mul(0.5f, add(p.A, t00)), M11 = mul(0.5f, add(p.B, t11)), M01 = mul(0.5f, t01);
  const float Ml = mul(0.5f, add(p.E, t.E));                       // ref:153
  const float eps = mul(mul(p.amb, t.amb), s2);
  const float detN = fma(add(p.A, t.A), add(p.B, t.B), eps);       // det(Sp + St)
  const float det = mul(0.25f, detN);                              // ref:155-157
  *rare |= !(det >= 1e-7f);                                        // clamp active (or nan), ref:158
  const float idet = Mth<float>::rcp(det), iMl = Mth<float>::rcp(Ml);
  float Q2 = mul(mul(u, u), M11);                                  // d^T adj(M) d
  Q2 = fma(mul(-2.0f, mul(u, v)), M01, Q2);
  Q2 = fma(mul(v, v), M00, Q2);
  const float mh = fma(Q2, idet, mul(mul(dz, dz), iMl));
  const float maha = mul(mul(0.125f, P.inv_alpha2), mh);           // ref:170-172,182
  // shape: 0.5 ln det + 0.5 ln Ml - 0.25 ln(ABE) - 0.25 ln(CDF)    ref:174-180
  const float da = sub(p.a, t.a), db = sub(p.b, t.b), de = sub(p.e, t.e);
  const float xe = mul(mul(de, de), mul(0.5f, mul(p.ie, t.ie)));   // Ml/(e_p e_t) - 1
  const float xa = mul(mul(da, da), mul(0.5f, mul(p.ia, t.ia)));
  const float xb = mul(mul(db, db), mul(0.5f, mul(p.ib, t.ib)));
  float q = add(xa, xb);
  q = fma(xa, xb, q);
  q = fma(mul(0.25f, eps), mul(p.iab, t.iab), q);
  const float qq = fma(q, xe, add(q, xe));                         // (1+q)(1+xe) - 1
  *rare |= !(qq >= 0.0f && qq < 1e30f);
  const float shape = mul(0.5f, log1p_pos(qq));
  float val = add(maha, shape);
  if (P.flag) val = sqrt_clamp0(val);                              // ref:183-184
  return post(val, P, rare);
}
// the FAST value of one pair of NICE boxes for every distance that has one
template <int LOSS>
GD_HD float value(const BoxGauss<float>& p, const BoxGauss<float>& t, const PairParams<float>& P,
                  bool* rare) {
  if constexpr (LOSS == kGwd) return gwd_value(p, t, P, rare);
  else if constexpr (LOSS == kBd) return bd_value(p, t, P, rare);
  else return kld_family_value<LOSS>(p, t, P, rare);
}
}  // namespace pw

// losses whose pairwise FAST value is the explicit-rounding form above
template <int LOSS>
#if defined(GD_PW_IMPLICIT)   // host experiments only: the templated cores with implicit contraction
struct PairwiseExact { static constexpr bool value = false; };
#else
struct PairwiseExact { static constexpr bool value = (LOSS != kKfiou); };   // kfiou3d: robust only
#endif

// The robust value behind a call: ONE body per translation unit, the same machine code for
// every kernel that reaches it (cold: degenerate boxes, tripped guards).
template <int LOSS>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
float pair_value_robust(const BoxGauss<float> p, const BoxGauss<float> t,
                        const PairParams<float> P) {       // by value: no address of the
  return pair_value<float, LOSS>(p, t, P);                 // caller's registers escapes
}

// FAST value for two boxes the caller knows to be nice (exact-form losses only): *rare is
// OR-ed with "a guard inside the distance tripped" -- the value is then meaningless and the
// caller must redo the pair with pair_value_auto.
template <typename T, int LOSS>
GD_HD T pair_value_fast(const BoxGauss<T>& p, const BoxGauss<T>& t, const PairParams<T>& P,
                        bool* rare) {
  static_assert(PairwiseExact<LOSS>::value && std::is_same<T, float>::value,
                "only the explicit-rounding value cores may be called outside pair_value_auto");
  return pw::value<LOSS>(p, t, P, rare);
}

// value through the branch-free FAST cores when both boxes are nice and no guard
// trips, else through the robust cores (a cold branch).
template <typename T, int LOSS>
GD_HD T pair_value_auto(const BoxGauss<T>& p, const BoxGauss<T>& t, const PairParams<T>& P) {
  if constexpr (PairwiseExact<LOSS>::value && std::is_same<T, float>::value) {
    // the FAST core runs unconditionally (straight-line; on boxes that are not nice its result
    // is discarded): one cold branch per pair instead of two
    bool rare = !(p.nice && t.nice);
    const T v = pw::value<LOSS>(p, t, P, &rare);
    if (!rare) return v;
    return pair_value_robust<LOSS>(p, t, P);
  } else {
    const PairGeom<T> g = geom_from_gauss(p, t);
    if (LOSS != kKfiou) {
      bool rare = !(p.nice && t.nice);
      const T v = core_eval<T, LOSS, false, true>(g, P, (T)1, (T*)0, &rare);
      if (!rare) return v;
    }
    bool unused = false;
    return core_eval<T, LOSS, false, false>(g, P, (T)1, (T*)0, &unused);
  }
}

}  // namespace gd
