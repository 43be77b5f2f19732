// Two box pairs per thread in one 64-bit register: the packed-FP32 instantiation of the
// per-pair math in gd_math.cuh (T = gd::f2).
//
// Blackwell (sm_100a) has packed single-precision instructions -- FADD2 / FMUL2 / FFMA2
// (PTX add / mul / fma .f32x2) -- that process the two halves of an aligned register pair
// in ONE issue slot.  They do not raise the FMA-per-second peak (tools/micro/
// ffma2_probe.cu: 36.8 T FMA/s either way) but halve the instructions issued, fetched
// and decoded for the ~60 % of the fused kernel's instruction stream that is FP32
// arithmetic -- and instruction energy is what bounds that kernel under the 1000 W cap
// (DESIGN.md section 4).
//
// gd_math.cuh is generic in its value type T, its comparison-result type
// (Mth<T>::mask) and its select (Mth<T>::sel), so the SAME formulas (and the same
// reference citations) serve float, double and f2; this header only supplies the f2
// arithmetic.  Only the branch-free FAST cores are meant to be instantiated with f2
// (gwd3d, kld3d, bd3d); rows a FAST core flags are redone by the scalar robust path,
// exactly as in the scalar kernels.
//
// mul/add are emitted WITHOUT a rounding modifier so ptxas may contract them into FFMA2
// exactly as it contracts the scalar code into FFMA; where the algorithm needs a true
// fused multiply-add (range reductions, polynomial tails) `fma()` is explicit.
// On the host (tests/host_math) the halves are evaluated with plain float arithmetic,
// which makes the f2 instantiation bit-identical to the float instantiation there.
#pragma once
#include "gd_math.cuh"

namespace gd {

struct m2 {            // one comparison result per half
  bool lo, hi;
};
GD_HD m2 operator!(m2 a) { return m2{!a.lo, !a.hi}; }
GD_HD m2 operator&&(m2 a, m2 b) { return m2{a.lo && b.lo, a.hi && b.hi}; }
GD_HD m2 operator||(m2 a, m2 b) { return m2{a.lo || b.lo, a.hi || b.hi}; }
GD_HD m2& operator|=(m2& a, m2 b) {
  a.lo = a.lo || b.lo;
  a.hi = a.hi || b.hi;
  return a;
}

struct f2 {
  unsigned long long v;        // {lo, hi} as the two halves of one aligned register pair
  f2() = default;
  template <typename S>
  GD_HD explicit f2(S s);      // broadcast
};

GD_HD f2 mk2(float lo, float hi) {
  f2 r;
#if defined(__CUDA_ARCH__)
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
  uint32_t a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  r.v = (unsigned long long)a | ((unsigned long long)b << 32);
#endif
  return r;
}
GD_HD float lo2(f2 x) {
#if defined(__CUDA_ARCH__)
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
  (void)b;
  return a;
#else
  const uint32_t a = (uint32_t)(x.v & 0xffffffffull);
  float f;
  memcpy(&f, &a, 4);
  return f;
#endif
}
GD_HD float hi2(f2 x) {
#if defined(__CUDA_ARCH__)
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
  (void)a;
  return b;
#else
  const uint32_t b = (uint32_t)(x.v >> 32);
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
template <typename S>
GD_HD f2::f2(S s) {
  *this = mk2((float)s, (float)s);
}

GD_HD f2 operator+(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  f2 r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return mk2(lo2(a) + lo2(b), hi2(a) + hi2(b));
#endif
}
GD_HD f2 operator-(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  f2 r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return mk2(lo2(a) - lo2(b), hi2(a) - hi2(b));
#endif
}
GD_HD f2 operator*(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
  f2 r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return mk2(lo2(a) * lo2(b), hi2(a) * hi2(b));
#endif
}
GD_HD f2 operator-(f2 a) {   // per-half negation: ptxas folds it into the -R operand modifiers
  return mk2(-lo2(a), -hi2(a));
}
GD_HD f2& operator-=(f2& a, f2 b) {
  a = a - b;
  return a;
}
GD_HD f2& operator+=(f2& a, f2 b) {
  a = a + b;
  return a;
}
GD_HD f2& operator*=(f2& a, f2 b) {
  a = a * b;
  return a;
}
GD_HD f2 fma2(f2 a, f2 b, f2 c) {              // fused, round-to-nearest, both halves
#if defined(__CUDA_ARCH__)
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
#else
  return mk2(fmaf(lo2(a), lo2(b), lo2(c)), fmaf(hi2(a), hi2(b), hi2(c)));
#endif
}

GD_HD m2 operator<(f2 a, f2 b) { return m2{lo2(a) < lo2(b), hi2(a) < hi2(b)}; }
GD_HD m2 operator>(f2 a, f2 b) { return m2{lo2(a) > lo2(b), hi2(a) > hi2(b)}; }
GD_HD m2 operator<=(f2 a, f2 b) { return m2{lo2(a) <= lo2(b), hi2(a) <= hi2(b)}; }
GD_HD m2 operator>=(f2 a, f2 b) { return m2{lo2(a) >= lo2(b), hi2(a) >= hi2(b)}; }
GD_HD m2 operator==(f2 a, f2 b) { return m2{lo2(a) == lo2(b), hi2(a) == hi2(b)}; }
GD_HD m2 operator!=(f2 a, f2 b) { return m2{lo2(a) != lo2(b), hi2(a) != hi2(b)}; }

#define GD_F2_PER_HALF(fn) \
  static GD_HD f2 fn(f2 x) { return mk2(Mth<float>::fn(lo2(x)), Mth<float>::fn(hi2(x))); }

template <>
struct Mth<f2> {
  typedef m2 mask;
  static GD_HD f2 sel(m2 c, f2 a, f2 b) {
    return mk2(c.lo ? lo2(a) : lo2(b), c.hi ? hi2(a) : hi2(b));
  }
  // MUFU-class functions work on one half at a time (the special-function unit is scalar)
  GD_F2_PER_HALF(rcp)
  GD_F2_PER_HALF(sqrt)
  GD_F2_PER_HALF(rsqrt)
  GD_F2_PER_HALF(rsixthroot)
  // robust-path functions: never reached from the FAST cores, present so that the
  // run-time `fun` / clamp branches of the generic code compile
  GD_F2_PER_HALF(log)
  GD_F2_PER_HALF(log1p)
  GD_F2_PER_HALF(expm1)
  GD_F2_PER_HALF(exp)
  GD_F2_PER_HALF(rcbrt)
  static GD_HD void sincos(f2 x, f2* s, f2* c) {
    float s0, c0, s1, c1;
    Mth<float>::sincos(lo2(x), &s0, &c0);
    Mth<float>::sincos(hi2(x), &s1, &c1);
    *s = mk2(s0, s1);
    *c = mk2(c0, c1);
  }
  static GD_HD f2 inf() { return f2(Mth<float>::inf()); }
  static GD_HD m2 nice_row_minmax(f2 p3, f2 p4, f2 p5, f2 t3, f2 t4, f2 t5, f2 yp, f2 yt) {
    return m2{Mth<float>::nice_row_minmax(lo2(p3), lo2(p4), lo2(p5), lo2(t3), lo2(t4), lo2(t5),
                                          lo2(yp), lo2(yt)),
              Mth<float>::nice_row_minmax(hi2(p3), hi2(p4), hi2(p5), hi2(t3), hi2(t4), hi2(t5),
                                          hi2(yp), hi2(yt))};
  }

  // Same algorithm and constants as Mth<float>::sincos_fast; the quadrant index is
  // taken per half, the Cody-Waite reduction and the two polynomials run packed.
  static GD_HD void sincos_fast(f2 x, f2* sn, f2* cs) {
#if defined(__CUDA_ARCH__)
    const int q0 = __float2int_rn(lo2(x) * 0.63661974668502807617f);
    const int q1 = __float2int_rn(hi2(x) * 0.63661974668502807617f);
#else
    const int q0 = (int)::rintf(lo2(x) * 0.63661974668502807617f);
    const int q1 = (int)::rintf(hi2(x) * 0.63661974668502807617f);
#endif
    const f2 k = mk2((float)q0, (float)q1);
    f2 r = fma2(k, f2(-1.5707962512969970703f), x);
    r = fma2(k, f2(-7.5497894158615963534e-08f), r);
    r = fma2(k, f2(-5.3903029534742383927e-15f), r);
    const f2 z = r * r;
    f2 ps = fma2(z, f2(-1.9515295891e-4f), f2(8.3327032626e-3f));
    ps = fma2(z, ps, f2(-1.6666662693e-1f));
    const f2 sr = fma2(r * z, ps, r);
    f2 pc = fma2(z, f2(2.44331570e-5f), f2(-1.38878601e-3f));
    pc = fma2(z, pc, f2(4.16667275e-2f));
    pc = fma2(z, pc, f2(-4.99999970e-1f));
    const f2 cr = fma2(z, pc, f2(1.0f));
    const float sr0 = lo2(sr), sr1 = hi2(sr), cr0 = lo2(cr), cr1 = hi2(cr);
    const float s0 = (q0 & 1) ? cr0 : sr0, c0 = (q0 & 1) ? sr0 : cr0;
    const float s1 = (q1 & 1) ? cr1 : sr1, c1 = (q1 & 1) ? sr1 : cr1;
    *sn = mk2((q0 & 2) ? -s0 : s0, (q1 & 2) ? -s1 : s1);
    *cs = mk2(((q0 + 1) & 2) ? -c0 : c0, ((q1 + 1) & 2) ? -c1 : c1);
  }

  // range reduction y = 2^k m of one half (Mth<float>::log1p_pos / sum_minus_log_ratios)
  static GD_HD void split(float y, int* k, float* m) {
    uint32_t yb;
    memcpy(&yb, &y, 4);
    *k = (int)((int32_t)(yb - 0x3f3504f3u) >> 23);
    const uint32_t mb = yb - ((uint32_t)(*k) << 23);
    memcpy(m, &mb, 4);
  }
  // 2 s^3 P(s^2) / (2 s z) = P(z): the shared series of both log helpers
  static GD_HD f2 log_series(f2 z) {
    f2 p = f2(1.0f / 13.0f);
    p = fma2(p, z, f2(1.0f / 11.0f));
    p = fma2(p, z, f2(1.0f / 9.0f));
    p = fma2(p, z, f2(1.0f / 7.0f));
    p = fma2(p, z, f2(1.0f / 5.0f));
    p = fma2(p, z, f2(1.0f / 3.0f));
    return p;
  }
  // log(1+x), 0 <= x < ~1e30 (Mth<float>::log1p_pos, same operations per half)
  static GD_HD f2 log1p_pos(f2 x) {
    const f2 y = f2(1.0f) + x;
    int k0, k1;
    float m0, m1;
    split(lo2(y), &k0, &m0);
    split(hi2(y), &k1, &m1);
    const f2 f = mk2((k0 == 0) ? lo2(x) : (m0 - 1.0f), (k1 == 0) ? hi2(x) : (m1 - 1.0f));
    const f2 s = f * rcp(f2(2.0f) + f);
    const f2 z = s * s;
    const f2 p = log_series(z);
    const f2 kf = mk2((float)k0, (float)k1);
    const f2 lo = fma2(kf, f2(1.428606765330187e-06f), f2(2.0f) * s * z * p);
    return fma2(kf, f2(0.693145751953125f), fma2(f2(2.0f), s, lo));
  }
  // S - log(r1 r2 r3) (Mth<float>::sum_minus_log_ratios); FAST only: a half whose
  // product leaves [1e-30, 1e30] is flagged and redone by the scalar robust path
  template <bool FAST>
  static GD_HD f2 sum_minus_log_ratios(f2 S, f2 pair, f2 r1, f2 r2, f2 r3, m2* rare) {
    static_assert(FAST, "the packed instantiation only provides the FAST cores");
    const f2 y = r1 * r2 * r3;
    *rare |= !(y > f2(1.0e-30f) && y < f2(1.0e30f));
    int k0, k1;
    float m0, m1;
    split(lo2(y), &k0, &m0);
    split(hi2(y), &k1, &m1);
    const f2 x = S + pair;
    const f2 f = mk2((k0 == 0) ? lo2(x) : (m0 - 1.0f), (k1 == 0) ? hi2(x) : (m1 - 1.0f));
    const f2 s = f * rcp(f2(2.0f) + f);
    const f2 z = s * s;
    const f2 p = log_series(z);
    const f2 tail = f2(2.0f) * s * z * p;
    const f2 kf = mk2((float)k0, (float)k1);
    f2 r = fma2(-kf, f2(0.693145751953125f), S);
    r = fma2(-kf, f2(1.428606765330187e-06f), r);
    const f2 direct = (r - f2(2.0f) * s) - tail;
    const f2 near1 = fma2(x, s, -tail) - pair;
    return mk2((k0 == 0) ? lo2(near1) : lo2(direct), (k1 == 0) ? hi2(near1) : hi2(direct));
  }
};
#undef GD_F2_PER_HALF

// broadcast the (scalar) loss parameters to both halves
GD_HD PairParams<f2> broadcast_params(const PairParams<float>& P) {
  PairParams<f2> Q;
  for (int i = 0; i < 3; ++i) Q.off[i] = f2(P.off[i]);
  Q.alpha2 = f2(P.alpha2);
  Q.inv_alpha2 = f2(P.inv_alpha2);
  Q.tau = f2(P.tau);
  Q.fun = P.fun;
  Q.tau_on = P.tau_on;
  Q.flag = P.flag;
  return Q;
}

// Two (pred, target) rows through the FAST cores at once.  pa/ta, pb/tb: the two rows;
// wsa/wsb their weight*scale factors; ga/gb receive the gradients; rare_a/rare_b are
// OR-ed with "this row must be redone on the robust path".  Returns the two values.
template <int LOSS, bool GRAD, int DIET = 0>
GD_HD void pair_eval_fast2(const float* pa, const float* ta, const float* pb, const float* tb,
                           const PairParams<f2>& P, float wsa, float wsb, float* ga, float* gb,
                           bool* rare_a, bool* rare_b, float* la, float* lb) {
  static_assert(LOSS == kGwd || LOSS == kKld || LOSS == kBd, "packed cores: gwd3d, kld3d, bd3d");
  f2 p[7], t[7], g[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    p[c] = mk2(pa[c], pb[c]);
    t[c] = mk2(ta[c], tb[c]);
  }
  m2 rare = m2{*rare_a, *rare_b};
  const f2 l = pair_eval_fast<f2, LOSS, GRAD, DIET>(p, t, P, mk2(wsa, wsb), g, &rare);
  *la = lo2(l);
  *lb = hi2(l);
  if (GRAD) {
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      ga[c] = lo2(g[c]);
      gb[c] = hi2(g[c]);
    }
  }
  *rare_a = rare.lo;
  *rare_b = rare.hi;
}

}  // namespace gd
