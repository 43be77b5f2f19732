// TEST-ONLY driver: runs the emulated pairwise kernels (pairwise_emul.cpp) under AddressSanitizer /
// UBSan on exact-size heap buffers (tools/asan_emulation.sh).  Not part of pytest (takes ~1 min).
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cmath>
extern "C" int gd_emul_pairwise(int loss, int packed, int reduce, int force_cpl, unsigned cap,
                                const float* b1, long long n, const float* b2, long long m,
                                float* out, float* row_min, int* row_argmin, float* col_min,
                                int* col_argmin, int similarity);
static void fill(std::vector<float>& b, unsigned seed) {
  unsigned x = seed;
  auto rnd = [&]() { x = x * 1664525u + 1013904223u; return (float)(x >> 8) / 16777216.0f; };
  for (size_t i = 0; i + 6 < b.size(); i += 7) {
    b[i] = 70 * rnd(); b[i+1] = 80 * rnd() - 40; b[i+2] = -1 + rnd();
    b[i+3] = 0.5f + 4 * rnd(); b[i+4] = 0.5f + 2 * rnd(); b[i+5] = 1 + rnd(); b[i+6] = 6.28f * rnd() - 3.14f;
  }
}
int main() {
  const int shapes[][2] = {{1,1},{63,5},{65,33},{131,129},{200,256},{70,300}};
  for (auto& s : shapes) {
    long long n = s[0], m = s[1];
    std::vector<float> b1(n * 7), b2(m * 7);
    fill(b1, 1); fill(b2, 2);
    if (n > 20) { b1[10*7+4] = 1e-9f; b1[11*7+3] = 2e7f; b1[12*7] = NAN; }
    for (int packed = 0; packed < 2; ++packed) for (int reduce = 0; reduce < 2; ++reduce)
      for (int loss : {0, 1, 5}) {
        std::vector<float> out(n * m), rmin(n), cmin(m);
        std::vector<int> ridx(n), cidx(m);
        int rc = gd_emul_pairwise(loss, packed, reduce, 0, 3, b1.data(), n, b2.data(), m, out.data(),
                                  rmin.data(), ridx.data(), cmin.data(), cidx.data(), 0);
        if (rc) { printf("dirty workspace\n"); return 1; }
        if (packed && reduce && m > 64) {
          rc = gd_emul_pairwise(loss, 1, 1, 2, 2, b1.data(), n, b2.data(), m, out.data(), rmin.data(),
                                ridx.data(), cmin.data(), cidx.data(), 1);
          if (rc) { printf("dirty workspace\n"); return 1; }
        }
      }
    printf("shape %lld x %lld ok\n", n, m);
  }
  return 0;
}
