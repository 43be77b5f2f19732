// Pairwise N x M Gaussian-distance matrix for B200 (sm_100a).
//
// New surface (the reference has only IoU matrices,
// core/bbox/assigners/sim_ota_3d_assigner.py:91-93): out[i,j] =
// postprocess(distance(boxes1[i], boxes2[j])), equal to the element-wise loss
// path of gaussian_distance_loss.py on the broadcast-expanded pairs.
//
// Mapping: thread <-> column j.  A CTA owns kRowsPerCta rows of boxes1; it
// converts them once to BoxGauss (centre, half extents, sin/cos yaw) in shared
// memory, each thread converts its own column box once into registers, and the
// inner loop over rows reads the row Gaussian as a shared-memory broadcast and
// writes out[i, j0 + tid] -- consecutive threads write consecutive floats, so
// the 4 B/pair output stream is fully coalesced.  FP32 CUDA-core math, no
// tensor cores (not a contraction).  The fused arg-reduction variant reduces
// (value, column) keys with warp shuffles + shared memory and need not write the
// matrix at all.
#include "gd_common.cuh"

namespace gdk {

constexpr int kRowsPerCta = 64;
constexpr int kWarps = kThreads / 32;

// (value, column) packed so that an unsigned 64-bit min is "smaller value, then
// lower column"; NaN maps below everything (torch.min / argmin propagate NaN).
__device__ __forceinline__ unsigned long long pack_key(float v, unsigned int j) {
  unsigned int b = __float_as_uint(v);
  b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;
  if (v != v) b = 0u;
  return ((unsigned long long)b << 32) | j;
}
__device__ __forceinline__ float unpack_value(unsigned long long k) {
  unsigned int b = (unsigned int)(k >> 32);
  if (b == 0u) return __uint_as_float(0x7fc00000u);
  b ^= (b >> 31) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(b);
}

// One kernel for both uses so the matrix and the fused arg-reduction run the
// SAME per-pair instruction sequence (indices derived from either are then
// bit-identical): WRITE stores the matrix, ARGMIN keeps per-row minima.
template <int LOSS, bool WRITE, bool ARGMIN>
__global__ void __launch_bounds__(kThreads) gd_pairwise_kernel(
    const float* __restrict__ b1, long long n, const float* __restrict__ b2, long long m,
    float* __restrict__ out, long long out_stride, float* __restrict__ row_min,
    int* __restrict__ row_argmin, const gd::PairParams<float> pp) {
  __shared__ gd::BoxGauss<float> s_rows[kRowsPerCta];
  __shared__ unsigned long long s_best[ARGMIN ? kRowsPerCta : 1][kWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row0 = (long long)blockIdx.x * kRowsPerCta;
  const int rows = (int)min((long long)kRowsPerCta, n - row0);
  if (tid < rows) s_rows[tid] = gd::box_gauss(b1 + (row0 + tid) * 7, pp);
  if (ARGMIN) {
    for (int i = tid; i < kRowsPerCta * kWarps; i += kThreads)
      (&s_best[0][0])[i] = ~0ull;
  }
  __syncthreads();
  // whole-CTA column chunks so that every lane takes part in the warp reductions
  for (long long c0 = (long long)blockIdx.y * kThreads; c0 < m;
       c0 += (long long)gridDim.y * kThreads) {
    const long long j = c0 + tid;
    const bool live = j < m;
    gd::BoxGauss<float> t;
    if (live) t = gd::box_gauss(b2 + j * 7, pp);
    else t = s_rows[0];                    // any valid box: result is discarded
    float* o = out + row0 * out_stride + j;
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
      const float v = gd::pair_value_auto<float, LOSS>(s_rows[r], t, pp);
      if (WRITE && live) __stcs(o + (long long)r * out_stride, v);
      if (ARGMIN) {
        unsigned long long k = live ? pack_key(v, (unsigned int)j) : ~0ull;
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, sh);
          k = other < k ? other : k;
        }
        if (lane == 0 && k < s_best[r][warp]) s_best[r][warp] = k;
      }
    }
  }
  if (ARGMIN) {
    __syncthreads();
    if (tid < rows) {
      unsigned long long k = s_best[tid][0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) k = s_best[tid][w] < k ? s_best[tid][w] : k;
      row_min[row0 + tid] = unpack_value(k);
      row_argmin[row0 + tid] = (int)(unsigned int)(k & 0xffffffffu);
    }
  }
}

template <int LOSS>
int launch_pairwise(const float* b1, long long n, const float* b2, long long m, float* out,
                    long long out_stride, float* row_min, int* row_argmin,
                    const gd::PairParams<float>& pp, cudaStream_t st) {
  const long long gx = (n + kRowsPerCta - 1) / kRowsPerCta;
  if (gx > 2147483647LL || m > 0xffffffffLL) return GD_ERR_BAD_ARG;
  if (row_argmin) {                        // fused arg-reduction: one CTA walks all columns
    dim3 grid((unsigned)gx, 1);
    if (out)
      gd_pairwise_kernel<LOSS, true, true><<<grid, kThreads, 0, st>>>(
          b1, n, b2, m, out, out_stride, row_min, row_argmin, pp);
    else
      gd_pairwise_kernel<LOSS, false, true><<<grid, kThreads, 0, st>>>(
          b1, n, b2, m, out, out_stride, row_min, row_argmin, pp);
  } else {
    long long gy = (m + kThreads - 1) / kThreads;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    gd_pairwise_kernel<LOSS, true, false><<<grid, kThreads, 0, st>>>(
        b1, n, b2, m, out, out_stride, row_min, row_argmin, pp);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS>
int launch_argmin(const float* b1, long long n, const float* b2, long long m, float* row_min,
                  int* row_argmin, const gd::PairParams<float>& pp, cudaStream_t st) {
  return launch_pairwise<LOSS>(b1, n, b2, m, nullptr, m, row_min, row_argmin, pp, st);
}

}  // namespace gdk

extern "C" {

int gd_pairwise(const gd_loss_config* cfg, const float* boxes1, int64_t n, const float* boxes2,
                int64_t m, float* out, int64_t out_row_stride, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m < 0 || out_row_stride < m) return GD_ERR_BAD_ARG;
  if (n == 0 || m == 0) return 0;
  if (!boxes1 || !boxes2 || !out) return GD_ERR_BAD_ARG;
  const gd::PairParams<float> pp = make_pair_params(*cfg);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (cfg->loss_type) {
    case GD_LOSS_GWD3D: return launch_pairwise<gd::kGwd>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_KLD3D: return launch_pairwise<gd::kKld>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_JD3D: return launch_pairwise<gd::kJd>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_KLD3D_SYMMAX: return launch_pairwise<gd::kSymMax>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_KLD3D_SYMMIN: return launch_pairwise<gd::kSymMin>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_BD3D: return launch_pairwise<gd::kBd>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
    case GD_LOSS_KFIOU3D: return launch_pairwise<gd::kKfiou>(boxes1, n, boxes2, m, out, out_row_stride, nullptr, nullptr, pp, st);
  }
  return GD_ERR_BAD_ARG;
}

int gd_pairwise_row_argmin(const gd_loss_config* cfg, const float* boxes1, int64_t n,
                           const float* boxes2, int64_t m, float* row_min, int32_t* row_argmin,
                           void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m <= 0) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!boxes1 || !boxes2 || !row_min || !row_argmin) return GD_ERR_BAD_ARG;
  const gd::PairParams<float> pp = make_pair_params(*cfg);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (cfg->loss_type) {
    case GD_LOSS_GWD3D: return launch_argmin<gd::kGwd>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_KLD3D: return launch_argmin<gd::kKld>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_JD3D: return launch_argmin<gd::kJd>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_KLD3D_SYMMAX: return launch_argmin<gd::kSymMax>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_KLD3D_SYMMIN: return launch_argmin<gd::kSymMin>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_BD3D: return launch_argmin<gd::kBd>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
    case GD_LOSS_KFIOU3D: return launch_argmin<gd::kKfiou>(boxes1, n, boxes2, m, row_min, row_argmin, pp, st);
  }
  return GD_ERR_BAD_ARG;
}

}  // extern "C"
