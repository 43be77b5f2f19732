// C ABI of the pairwise N x M Gaussian-distance kernels (gd_pairwise.cuh; one
// instantiation per loss type in gd_pairwise_inst_*.cu).
#include "gd_pairwise.cuh"

namespace gdk {
extern template int launch_pairwise<gd::kGwd>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kKld>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kJd>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kSymMax>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kSymMin>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kBd>(const PairwiseArgs&, cudaStream_t);
extern template int launch_pairwise<gd::kKfiou>(const PairwiseArgs&, cudaStream_t);

// --- MaxIoUAssigner-style labels from the fused minima (similarity = 1 - distance) ---
__global__ void __launch_bounds__(kThreads) gd_assign_lowq_kernel(
    const float* __restrict__ col_min, const int* __restrict__ col_argmin, long long m,
    long long n, float min_pos, long long* __restrict__ assigned) {
  const long long j = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (j >= m) return;
  const float sim = 1.0f - col_min[j];
  const long long anchor = col_argmin[j];
  // mmdet walks the GTs in order and overwrites: the highest GT index wins
  if (sim >= min_pos && anchor >= 0 && anchor < n)
    atomicMax(reinterpret_cast<unsigned long long*>(assigned + anchor), (unsigned long long)(j + 1));
}

__global__ void __launch_bounds__(kThreads) gd_assign_rows_kernel(
    const float* __restrict__ row_min, const int* __restrict__ row_argmin, long long n,
    float pos_thr, float neg_lo, float neg_hi, long long* __restrict__ assigned,
    float* __restrict__ max_overlaps) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float sim = 1.0f - row_min[i];
  long long lab = -1;                                   // ignore
  if (sim >= neg_lo && sim < neg_hi) lab = 0;           // negative
  if (sim >= pos_thr) lab = (long long)row_argmin[i] + 1;
  const long long lowq = assigned[i];                   // set by gd_assign_lowq_kernel, else 0
  if (lowq > 0) lab = lowq;
  assigned[i] = lab;
  if (max_overlaps) max_overlaps[i] = sim;
}

static int dispatch_pairwise(const gd_loss_config* cfg, const PairwiseArgs& a, cudaStream_t st) {
  switch (cfg->loss_type) {
    case GD_LOSS_GWD3D: return launch_pairwise<gd::kGwd>(a, st);
    case GD_LOSS_KLD3D: return launch_pairwise<gd::kKld>(a, st);
    case GD_LOSS_JD3D: return launch_pairwise<gd::kJd>(a, st);
    case GD_LOSS_KLD3D_SYMMAX: return launch_pairwise<gd::kSymMax>(a, st);
    case GD_LOSS_KLD3D_SYMMIN: return launch_pairwise<gd::kSymMin>(a, st);
    case GD_LOSS_BD3D: return launch_pairwise<gd::kBd>(a, st);
    case GD_LOSS_KFIOU3D: return launch_pairwise<gd::kKfiou>(a, st);
  }
  return GD_ERR_BAD_ARG;
}
}  // namespace gdk

extern "C" {

int gd_pairwise(const gd_loss_config* cfg, const float* boxes1, int64_t n, const float* boxes2,
                int64_t m, float* out, int64_t out_row_stride, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m < 0 || out_row_stride < m) return GD_ERR_BAD_ARG;
  if (n == 0 || m == 0) return 0;
  if (!boxes1 || !boxes2 || !out) return GD_ERR_BAD_ARG;
  PairwiseArgs a{};
  a.b1 = boxes1;
  a.n = n;
  a.b2 = boxes2;
  a.m = m;
  a.out = out;
  a.out_stride = out_row_stride;
  a.pp = make_pair_params(*cfg);
  return dispatch_pairwise(cfg, a, reinterpret_cast<cudaStream_t>(stream));
}

int gd_pairwise_row_argmin(const gd_loss_config* cfg, const float* boxes1, int64_t n,
                           const float* boxes2, int64_t m, float* row_min, int32_t* row_argmin,
                           void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m <= 0) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!boxes1 || !boxes2 || !row_min || !row_argmin) return GD_ERR_BAD_ARG;
  PairwiseArgs a{};
  a.b1 = boxes1;
  a.n = n;
  a.b2 = boxes2;
  a.m = m;
  a.out_stride = m;
  a.row_min = row_min;
  a.row_argmin = row_argmin;
  a.pp = make_pair_params(*cfg);
  return dispatch_pairwise(cfg, a, reinterpret_cast<cudaStream_t>(stream));
}

size_t gd_pairwise_workspace_bytes(int64_t m) {
  return 256 + sizeof(unsigned long long) * (size_t)(m > 0 ? m : 0);
}

int gd_pairwise_assign(const gd_loss_config* cfg, const float* boxes1, int64_t n,
                       const float* boxes2, int64_t m, float* row_min, int32_t* row_argmin,
                       float* col_min, int32_t* col_argmin, float* out, int64_t out_row_stride,
                       int32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || m <= 0 ||
      (flags & ~(GD_PAIR_SIMILARITY | GD_PAIR_CPL1 | GD_PAIR_INDEX64)))
    return GD_ERR_BAD_ARG;
  if (out && out_row_stride < m) return GD_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < gd_pairwise_workspace_bytes(m)) return GD_ERR_WORKSPACE;
  if (n == 0) return 0;
  if (!boxes1 || !boxes2 || !row_min || !row_argmin || !col_min || !col_argmin)
    return GD_ERR_BAD_ARG;
  PairwiseArgs a{};
  a.b1 = boxes1;
  a.n = n;
  a.b2 = boxes2;
  a.m = m;
  a.out = out;
  a.out_stride = out ? out_row_stride : m;
  a.similarity = (flags & GD_PAIR_SIMILARITY) ? 1 : 0;
  a.force_cpl1 = (flags & GD_PAIR_CPL1) ? 1 : 0;
  a.idx64 = (flags & GD_PAIR_INDEX64) ? 1 : 0;
  a.row_min = row_min;
  a.row_argmin = row_argmin;
  a.ticket = reinterpret_cast<unsigned int*>(workspace);
  a.col_keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  a.col_min = col_min;
  a.col_argmin = col_argmin;
  a.pp = make_pair_params(*cfg);
  return dispatch_pairwise(cfg, a, reinterpret_cast<cudaStream_t>(stream));
}

int gd_assign_from_minima(const float* row_min, const int32_t* row_argmin, int64_t n,
                          const float* col_min, const int32_t* col_argmin, int64_t m,
                          float pos_thr, float neg_lo, float neg_hi, float min_pos,
                          int32_t match_low_quality, int64_t* assigned_gt_inds,
                          float* max_overlaps, void* stream) {
  using namespace gdk;
  if (n < 0 || m < 0) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!row_min || !row_argmin || !assigned_gt_inds) return GD_ERR_BAD_ARG;
  if (match_low_quality && m > 0 && (!col_min || !col_argmin)) return GD_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(assigned_gt_inds, 0, sizeof(int64_t) * (size_t)n, st);
  if (e != cudaSuccess) return (int)e;
  if (match_low_quality && m > 0) {
    gd_assign_lowq_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        col_min, col_argmin, m, n, min_pos, reinterpret_cast<long long*>(assigned_gt_inds));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  gd_assign_rows_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, st>>>(
      row_min, row_argmin, n, pos_thr, neg_lo, neg_hi,
      reinterpret_cast<long long*>(assigned_gt_inds), max_overlaps);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

}  // extern "C"
