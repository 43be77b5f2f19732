// TEST-ONLY driver: runs the emulated loss kernels (loss_emul.cpp) under AddressSanitizer / UBSan on
// exact-size, 16-byte aligned heap buffers (tools/asan_emulation.sh).  Not part of pytest (~2 min).
#include <vector>
#include <cstdio>
#include <cmath>
extern "C" int gd_emul_loss(int loss, int kind, int spec, int pack, int grid, int warps,
                            const float* pred, const float* target, const float* weight,
                            int wmode, long long n, float scale, float tau, int mask_zero,
                            float* loss_sum, float* row_loss, float* grad);
static void fill(std::vector<float>& b, unsigned seed, float jitter) {
  unsigned x = seed;
  auto rnd = [&]() { x = x * 1664525u + 1013904223u; return (float)(x >> 8) / 16777216.0f; };
  for (size_t i = 0; i + 6 < b.size(); i += 7) {
    b[i] = 70 * rnd() + jitter * rnd(); b[i+1] = 80 * rnd() - 40; b[i+2] = -1 + rnd();
    b[i+3] = 0.5f + 4 * rnd(); b[i+4] = 0.5f + 2 * rnd(); b[i+5] = 1 + rnd(); b[i+6] = 6.28f * rnd() - 3.14f;
  }
}
int main() {
  for (long long n : {4LL, 131LL, 1003LL, 5003LL}) {
    // 16-byte aligned exact-size buffers (the bulk path needs the alignment; ASAN guards the ends)
    float *pred = nullptr, *target = nullptr, *w = nullptr, *w7 = nullptr, *grad = nullptr, *rows = nullptr;
    posix_memalign((void**)&pred, 16, n * 28); posix_memalign((void**)&target, 16, n * 28);
    posix_memalign((void**)&w, 16, n * 4 + 16); posix_memalign((void**)&w7, 16, n * 28);
    posix_memalign((void**)&grad, 16, n * 28); posix_memalign((void**)&rows, 16, n * 4 + 16);
    std::vector<float> tp(n * 7), tt(n * 7);
    fill(tp, 1, 0.3f); fill(tt, 1, 0.0f);
    for (long long i = 0; i < n * 7; ++i) { pred[i] = tp[i]; target[i] = tt[i]; w7[i] = 0.5f; }
    for (long long i = 0; i < n; ++i) w[i] = (i % 3) ? 1.0f : 0.0f;
    if (n > 200) { pred[5*7+4] = 1e-9f; pred[130*7+6] = 1000.f; target[(n-2)*7+3] = 2e4f; pred[7*7] = NAN; }
    float total = 0;
    for (int loss : {0, 1, 5}) {
      for (int spec : {8, 9, 13}) for (int pack = 0; pack < 2; ++pack) for (int wm = 0; wm < 2; ++wm) {
        int rc = gd_emul_loss(loss, 1, spec, pack, 3, (n % 2) ? 5 : 12, pred, target, wm ? w : nullptr, wm, n,
                              1.0f / n, spec == 13 ? 1.0f : 0.0f, 1, &total, nullptr, grad);
        if (rc) { printf("rc %d\n", rc); return 1; }
      }
      // run-time-parameter kernel: [N,7] weights, row loss, forward only; staged kernel
      if (gd_emul_loss(loss, 1, -1, 0, 2, 3, pred, target, w7, 2, n, 1.0f, 0.0f, 0, &total, rows, grad)) return 1;
      if (gd_emul_loss(loss, 1, -1, 0, 2, 3, pred, target, w, 1, n, 1.0f, 0.0f, 0, &total, rows, nullptr)) return 1;
      if (gd_emul_loss(loss, 0, -1, 0, 3, 8, pred, target, w, 1, n, 1.0f, 0.0f, 0, &total, rows, grad)) return 1;
      if (gd_emul_loss(loss, 0, -1, 0, 1, 8, pred, target, w7, 2, n, 1.0f, 0.0f, 0, &total, nullptr, nullptr)) return 1;
    }
    printf("n %lld ok (last sum %g)\n", n, total);
    free(pred); free(target); free(w); free(w7); free(grad); free(rows);
  }
  return 0;
}
