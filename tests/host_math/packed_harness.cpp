// TEST-ONLY host build of csrc/gd_packed.cuh (the packed two-rows-per-register
// instantiation of the FAST cores).  On the host the two halves are evaluated with plain
// float arithmetic, so pair_eval_fast2 must reproduce pair_eval_fast<float> BIT FOR BIT,
// value, gradient and "redo on the robust path" flag: this checks the generic plumbing
// (operators, masks, selects, the packed sin/cos and log helpers), not the device
// instruction selection.  Compiled and run by tests/test_host_math.py; exit code 0 = no
// mismatch.  Not part of the shipped library.
#include "../../mmdet3d_gaussian_b200/csrc/gd_packed.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <random>
template <int LOSS>
long run(int fun, int tau_on, int flag, long n, unsigned seed, long* nrare) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> N01(0.f, 1.f);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  gd::PairParams<float> P;
  P.off[0] = 0; P.off[1] = 0; P.off[2] = 0.5f; P.alpha2 = 1; P.inv_alpha2 = 1; P.tau = tau_on ? 2.5f : 0.f;
  P.fun = fun; P.tau_on = tau_on; P.flag = flag;
  const gd::PairParams<gd::f2> Q = gd::broadcast_params(P);
  long bad = 0;
  for (long i = 0; i < n; ++i) {
    float p[2][7], t[2][7];
    for (int h = 0; h < 2; ++h) {
      float sig = (i % 3 == 0) ? 0.005f : ((i % 3 == 1) ? 0.05f : 0.3f);
      t[h][0] = 70 * U(rng); t[h][1] = 80 * U(rng) - 40; t[h][2] = -1 + 0.5f * N01(rng);
      t[h][3] = 1.7f * expf(0.3f * N01(rng)); t[h][4] = 0.7f * expf(0.3f * N01(rng)); t[h][5] = 1.6f * expf(0.2f * N01(rng));
      t[h][6] = 6.2831853f * U(rng) - 3.1415927f;
      for (int c = 0; c < 3; ++c) p[h][c] = t[h][c] + sig * N01(rng);
      for (int c = 3; c < 6; ++c) p[h][c] = t[h][c] * expf(0.66f * sig * N01(rng));
      p[h][6] = t[h][6] + sig * N01(rng);
      if (i % 997 == 0 && h == 1) p[h][4] = 1e-9f;          // a row the FAST path must flag
      if (i % 1013 == 0 && h == 0) t[h][6] = 1000.f;
    }
    float g[2][7], gs[2][7], l[2], ls[2];
    bool r[2] = {false, false}, rs[2] = {false, false};
    gd::pair_eval_fast2<LOSS, true>(p[0], t[0], p[1], t[1], Q, 0.7f, 1.3f, g[0], g[1], &r[0], &r[1], &l[0], &l[1]);
    ls[0] = gd::pair_eval_fast<float, LOSS, true>(p[0], t[0], P, 0.7f, gs[0], &rs[0]);
    ls[1] = gd::pair_eval_fast<float, LOSS, true>(p[1], t[1], P, 1.3f, gs[1], &rs[1]);
    for (int h = 0; h < 2; ++h) {
      if (r[h] != rs[h]) { ++bad; continue; }
      if (r[h]) { ++*nrare; continue; }
      if (memcmp(&l[h], &ls[h], 4)) ++bad;
      if (memcmp(g[h], gs[h], 28)) ++bad;
    }
    // The instruction "diets" (gd::Diet: min/max row screen, alpha == 1 and
    // center_offset == (0,0,.5) folded) change nothing on the host either: x * 1.0f and
    // y + 0.0f * x are exact and -ffp-contract=off forbids the FMA contractions that
    // make the device results differ by <= 1 ulp per operation.
    constexpr int kAll = gd::kDietGuards | gd::kDietStd;
    float gd_[2][7], ld[2], g2[2][7], l2[2];
    bool rd[2] = {false, false}, r2[2] = {false, false};
    ld[0] = gd::pair_eval_fast<float, LOSS, true, kAll>(p[0], t[0], P, 0.7f, gd_[0], &rd[0]);
    ld[1] = gd::pair_eval_fast<float, LOSS, true, kAll>(p[1], t[1], P, 1.3f, gd_[1], &rd[1]);
    gd::pair_eval_fast2<LOSS, true, kAll>(p[0], t[0], p[1], t[1], Q, 0.7f, 1.3f, g2[0], g2[1], &r2[0], &r2[1], &l2[0], &l2[1]);
    for (int h = 0; h < 2; ++h) {
      if (rd[h] != rs[h] || r2[h] != rs[h]) { ++bad; continue; }
      if (rs[h]) continue;
      if (memcmp(&ld[h], &ls[h], 4) || memcmp(&l2[h], &ls[h], 4)) ++bad;
      if (memcmp(gd_[h], gs[h], 28) || memcmp(g2[h], gs[h], 28)) ++bad;
    }
  }
  return bad;
}
int main() {
  long total_bad = 0;
  for (int fun = 0; fun < 2; ++fun) for (int tau = 0; tau < 2; ++tau) for (int flag = 0; flag < 2; ++flag) {
    long nr = 0;
    long b0 = run<gd::kGwd>(fun, tau, flag, 20000, 1, &nr);
    long b1 = run<gd::kKld>(fun, tau, flag, 20000, 2, &nr);
    long b5 = run<gd::kBd>(fun, tau, flag, 20000, 3, &nr);
    printf("fun %d tau %d flag %d: mismatches gwd %ld kld %ld bd %ld (rare rows %ld)\n", fun, tau, flag, b0, b1, b5, nr);
    total_bad += b0 + b1 + b5;
  }
  return total_bad ? 1 : 0;
}
