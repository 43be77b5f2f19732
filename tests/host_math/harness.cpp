// TEST-ONLY host build of the per-pair math in csrc/gd_math.cuh.
//
// Purpose: verify the hand-derived formulas and gradients against the fp64
// oracle on a machine with no GPU (float64 instantiation: formula check to
// ~1e-12; float32 instantiation: accuracy preview of the kernel's arithmetic).
// It is compiled by tests/test_host_math.py into a temp dir, is never part of
// the shipped library and is not a CPU fallback: the product path
// (libgdloss_b200.so) contains device code only.
#include "../../mmdet3d_gaussian_b200/csrc/gd_math.cuh"

template <typename T>
static void run(int loss, long n, const T* pred, const T* target, const double* off,
                double alpha, double tau, int fun, int flag, T* out_loss, T* out_grad) {
  gd::PairParams<T> P;
  for (int i = 0; i < 3; ++i) P.off[i] = (T)(float)off[i];  // ref:10 builds a float32 tensor first
  P.alpha2 = (T)(alpha * alpha);
  P.inv_alpha2 = (T)(1.0 / (alpha * alpha));
  P.tau = (T)tau;
  P.tau_on = tau >= 1.0;
  P.fun = fun;
  P.flag = flag;
  for (long i = 0; i < n; ++i) {
    const T* p = pred + 7 * i;
    const T* t = target + 7 * i;
    T* g = out_grad + 7 * i;
    switch (loss) {
      case 0: out_loss[i] = gd::pair_eval<T, 0, true>(p, t, P, (T)1, g); break;
      case 1: out_loss[i] = gd::pair_eval<T, 1, true>(p, t, P, (T)1, g); break;
      case 2: out_loss[i] = gd::pair_eval<T, 2, true>(p, t, P, (T)1, g); break;
      case 3: out_loss[i] = gd::pair_eval<T, 3, true>(p, t, P, (T)1, g); break;
      case 4: out_loss[i] = gd::pair_eval<T, 4, true>(p, t, P, (T)1, g); break;
      case 5: out_loss[i] = gd::pair_eval<T, 5, true>(p, t, P, (T)1, g); break;
      case 6: out_loss[i] = gd::pair_eval<T, 6, true>(p, t, P, (T)1, g); break;
    }
  }
}

// FAST path: returns per-row rare flags; rows flagged rare carry unspecified values
template <typename T>
static void run_fast(int loss, long n, const T* pred, const T* target, const double* off,
                     double alpha, double tau, int fun, int flag, T* out_loss, T* out_grad,
                     int* out_rare) {
  gd::PairParams<T> P;
  for (int i = 0; i < 3; ++i) P.off[i] = (T)(float)off[i];
  P.alpha2 = (T)(alpha * alpha);
  P.inv_alpha2 = (T)(1.0 / (alpha * alpha));
  P.tau = (T)tau;
  P.tau_on = tau >= 1.0;
  P.fun = fun;
  P.flag = flag;
  for (long i = 0; i < n; ++i) {
    const T* p = pred + 7 * i;
    const T* t = target + 7 * i;
    T* g = out_grad + 7 * i;
    bool rare = false;
    switch (loss) {
      case 0: out_loss[i] = gd::pair_eval_fast<T, 0, true>(p, t, P, (T)1, g, &rare); break;
      case 1: out_loss[i] = gd::pair_eval_fast<T, 1, true>(p, t, P, (T)1, g, &rare); break;
      case 2: out_loss[i] = gd::pair_eval_fast<T, 2, true>(p, t, P, (T)1, g, &rare); break;
      case 3: out_loss[i] = gd::pair_eval_fast<T, 3, true>(p, t, P, (T)1, g, &rare); break;
      case 4: out_loss[i] = gd::pair_eval_fast<T, 4, true>(p, t, P, (T)1, g, &rare); break;
      case 5: out_loss[i] = gd::pair_eval_fast<T, 5, true>(p, t, P, (T)1, g, &rare); break;
      case 6: out_loss[i] = gd::pair_eval_fast<T, 6, true>(p, t, P, (T)1, g, &rare); break;
    }
    out_rare[i] = rare ? 1 : 0;
  }
}

extern "C" {
void gd_host_eval_fast_f64(int loss, long n, const double* pred, const double* target,
                           const double* off, double alpha, double tau, int fun, int flag,
                           double* out_loss, double* out_grad, int* out_rare) {
  run_fast<double>(loss, n, pred, target, off, alpha, tau, fun, flag, out_loss, out_grad,
                   out_rare);
}
void gd_host_eval_fast_f32(int loss, long n, const float* pred, const float* target,
                           const double* off, double alpha, double tau, int fun, int flag,
                           float* out_loss, float* out_grad, int* out_rare) {
  run_fast<float>(loss, n, pred, target, off, alpha, tau, fun, flag, out_loss, out_grad,
                  out_rare);
}
void gd_host_eval_f64(int loss, long n, const double* pred, const double* target,
                      const double* off, double alpha, double tau, int fun, int flag,
                      double* out_loss, double* out_grad) {
  run<double>(loss, n, pred, target, off, alpha, tau, fun, flag, out_loss, out_grad);
}
void gd_host_eval_f32(int loss, long n, const float* pred, const float* target,
                      const double* off, double alpha, double tau, int fun, int flag,
                      float* out_loss, float* out_grad) {
  run<float>(loss, n, pred, target, off, alpha, tau, fun, flag, out_loss, out_grad);
}
float gd_host_sum_minus_log_ratios_f32(float S, float pair, float r1, float r2, float r3) {
  bool rare = false;
  return gd::Mth<float>::sum_minus_log_ratios<false>(S, pair, r1, r2, r3, &rare);
}
}

// pairwise value path (per-box precompute + shared cores)
template <typename T>
static void run_pairwise(int loss, long n, long m, const T* b1, const T* b2, const double* off,
                         double alpha, double tau, int fun, int flag, T* out) {
  gd::PairParams<T> P;
  for (int i = 0; i < 3; ++i) P.off[i] = (T)(float)off[i];
  P.alpha2 = (T)(alpha * alpha);
  P.inv_alpha2 = (T)(1.0 / (alpha * alpha));
  P.tau = (T)tau;
  P.tau_on = tau >= 1.0;
  P.fun = fun;
  P.flag = flag;
  for (long i = 0; i < n; ++i) {
    const gd::BoxGauss<T> p = gd::box_gauss(b1 + 7 * i, P);
    for (long j = 0; j < m; ++j) {
      const gd::BoxGauss<T> t = gd::box_gauss(b2 + 7 * j, P);
      T v = 0;
      switch (loss) {
        case 0: v = gd::pair_value_auto<T, 0>(p, t, P); break;
        case 1: v = gd::pair_value_auto<T, 1>(p, t, P); break;
        case 2: v = gd::pair_value_auto<T, 2>(p, t, P); break;
        case 3: v = gd::pair_value_auto<T, 3>(p, t, P); break;
        case 4: v = gd::pair_value_auto<T, 4>(p, t, P); break;
        case 5: v = gd::pair_value_auto<T, 5>(p, t, P); break;
        case 6: v = gd::pair_value_auto<T, 6>(p, t, P); break;
      }
      out[i * m + j] = v;
    }
  }
}
extern "C" {
void gd_host_pairwise_f64(int loss, long n, long m, const double* b1, const double* b2,
                          const double* off, double alpha, double tau, int fun, int flag,
                          double* out) {
  run_pairwise<double>(loss, n, m, b1, b2, off, alpha, tau, fun, flag, out);
}
void gd_host_pairwise_f32(int loss, long n, long m, const float* b1, const float* b2,
                          const double* off, double alpha, double tau, int fun, int flag,
                          float* out) {
  run_pairwise<float>(loss, n, m, b1, b2, off, alpha, tau, fun, flag, out);
}
}

// ---- decode front ends (csrc/gd_decode.cuh) --------------------------------
#include "../../mmdet3d_gaussian_b200/csrc/gd_decode.cuh"

template <typename T>
static gd::PairParams<T> mk_params(const double* off, double alpha, double tau, int fun, int flag) {
  gd::PairParams<T> P;
  for (int i = 0; i < 3; ++i) P.off[i] = (T)(float)off[i];
  P.alpha2 = (T)(alpha * alpha);
  P.inv_alpha2 = (T)(1.0 / (alpha * alpha));
  P.tau = (T)tau;
  P.tau_on = tau >= 1.0;
  P.fun = fun;
  P.flag = flag;
  return P;
}

template <typename T>
static void run_anchor(int loss, long n, const T* an, const T* dp, const T* dt, const double* off,
                       double alpha, double tau, int fun, int flag, T* out_loss, T* out_grad) {
  const gd::PairParams<T> P = mk_params<T>(off, alpha, tau, fun, flag);
  for (long i = 0; i < n; ++i) {
    const T *a = an + 7 * i, *p = dp + 7 * i, *t = dt + 7 * i;
    T* g = out_grad + 7 * i;
    switch (loss) {
      case 0: out_loss[i] = gd::anchor_pair_eval<T, 0, true>(a, p, t, P, (T)1, g); break;
      case 1: out_loss[i] = gd::anchor_pair_eval<T, 1, true>(a, p, t, P, (T)1, g); break;
      case 2: out_loss[i] = gd::anchor_pair_eval<T, 2, true>(a, p, t, P, (T)1, g); break;
      case 3: out_loss[i] = gd::anchor_pair_eval<T, 3, true>(a, p, t, P, (T)1, g); break;
      case 4: out_loss[i] = gd::anchor_pair_eval<T, 4, true>(a, p, t, P, (T)1, g); break;
      case 5: out_loss[i] = gd::anchor_pair_eval<T, 5, true>(a, p, t, P, (T)1, g); break;
      case 6: out_loss[i] = gd::anchor_pair_eval<T, 6, true>(a, p, t, P, (T)1, g); break;
    }
  }
}

template <typename T>
static void run_center(int loss, long n, const T* pr, long pstride, const long long* locs,
                       const T* tg, long tstride, const double* coder, int norm_bbox,
                       const double* off, double alpha, double tau, int fun, int flag,
                       T* out_loss, T* out_grad) {
  const gd::PairParams<T> P = mk_params<T>(off, alpha, tau, fun, flag);
  gd::CenterDecodeParams D;
  D.sx = coder[0];
  D.sy = coder[1];
  D.x0 = coder[2];
  D.y0 = coder[3];
  D.norm_bbox = norm_bbox;
  for (long i = 0; i < n; ++i) {
    const T *p = pr + pstride * i, *t = tg + tstride * i;
    const long long lx = locs[2 * i], ly = locs[2 * i + 1];
    T* g = out_grad + 7 * i;
    switch (loss) {
      case 0: out_loss[i] = gd::center_pair_eval<T, 0, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 1: out_loss[i] = gd::center_pair_eval<T, 1, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 2: out_loss[i] = gd::center_pair_eval<T, 2, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 3: out_loss[i] = gd::center_pair_eval<T, 3, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 4: out_loss[i] = gd::center_pair_eval<T, 4, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 5: out_loss[i] = gd::center_pair_eval<T, 5, true>(p, lx, ly, t, D, P, (T)1, g); break;
      case 6: out_loss[i] = gd::center_pair_eval<T, 6, true>(p, lx, ly, t, D, P, (T)1, g); break;
    }
  }
}

extern "C" {
void gd_host_anchor_f64(int loss, long n, const double* an, const double* dp, const double* dt,
                        const double* off, double alpha, double tau, int fun, int flag,
                        double* out_loss, double* out_grad) {
  run_anchor<double>(loss, n, an, dp, dt, off, alpha, tau, fun, flag, out_loss, out_grad);
}
void gd_host_anchor_f32(int loss, long n, const float* an, const float* dp, const float* dt,
                        const double* off, double alpha, double tau, int fun, int flag,
                        float* out_loss, float* out_grad) {
  run_anchor<float>(loss, n, an, dp, dt, off, alpha, tau, fun, flag, out_loss, out_grad);
}
void gd_host_center_f64(int loss, long n, const double* pr, long pstride, const long long* locs,
                        const double* tg, long tstride, const double* coder, int norm_bbox,
                        const double* off, double alpha, double tau, int fun, int flag,
                        double* out_loss, double* out_grad) {
  run_center<double>(loss, n, pr, pstride, locs, tg, tstride, coder, norm_bbox, off, alpha, tau,
                     fun, flag, out_loss, out_grad);
}
void gd_host_center_f32(int loss, long n, const float* pr, long pstride, const long long* locs,
                        const float* tg, long tstride, const double* coder, int norm_bbox,
                        const double* off, double alpha, double tau, int fun, int flag,
                        float* out_loss, float* out_grad) {
  run_center<float>(loss, n, pr, pstride, locs, tg, tstride, coder, norm_bbox, off, alpha, tau,
                    fun, flag, out_loss, out_grad);
}
}
