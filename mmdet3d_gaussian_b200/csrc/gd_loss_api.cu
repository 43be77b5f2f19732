// C ABI of the fused GD-loss kernels + the small helper kernels (autograd fold,
// early-return probe).  The heavy kernel templates live in gd_loss_kernels.cuh and
// are instantiated one loss type per translation unit (gd_loss_inst_*.cu) so the
// build parallelises.
#include "gd_loss_kernels.cuh"

namespace gdk {

std::atomic<int64_t> g_launches{0};

const DeviceInfo& device_info() {
  static DeviceInfo cache[64];
  static bool have[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!have[dev]) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cache[dev].sm_count = sms > 0 ? sms : 148;
    have[dev] = true;
  }
  return cache[dev];
}

// ---------------------------------------------------------------------------
// small helpers: autograd fold, early-return probe
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gd_scale_grad_kernel(float* __restrict__ grad,
                                                                 long long nel,
                                                                 const float* __restrict__ go) {
  const float s = __ldg(go);
  if (s == 1.0f) return;                  // the common case costs one launch, no traffic
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(grad) & 15u) == 0) {
    float4* g4 = reinterpret_cast<float4*>(grad);
    const long long nv = nel >> 2;
    for (long long j = i; j < nv; j += stride) {
      float4 v = g4[j];
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      g4[j] = v;
    }
    for (long long j = (nv << 2) + i; j < nel; j += stride) grad[j] *= s;
  } else {
    for (; i < nel; i += stride) grad[i] *= s;
  }
}

__global__ void __launch_bounds__(kThreads) gd_scale_grad_rows_kernel(
    float* __restrict__ grad, long long nel, const float* __restrict__ go, long long go_stride) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nel; i += stride) {
    grad[i] *= __ldg(go + (i / 7) * go_stride);
  }
}

__global__ void __launch_bounds__(kThreads) gd_any_positive_kernel(const float* __restrict__ w,
                                                                   long long count,
                                                                   int* __restrict__ flag) {
  const long long stride = (long long)gridDim.x * kThreads;
  bool any = false;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < count; i += stride) {
    any |= (w[i] > 0.0f);
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag, 1);
}

extern template int launch_loss<gd::kGwd>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kKld>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kJd>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kSymMax>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kSymMin>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kBd>(const LossArgs&, int, int, cudaStream_t);
extern template int launch_loss<gd::kKfiou>(const LossArgs&, int, int, cudaStream_t);

}  // namespace gdk

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int gd_abi_version(void) { return GD_ABI_VERSION; }

size_t gd_loss_workspace_bytes(int64_t n) {
  (void)n;
  return 256 + sizeof(double) * (size_t)gdk::kMaxGrid;   // ticket (padded) + partials
}

int gd_loss_fwd_bwd(const gd_loss_config* cfg, const float* pred, int64_t pred_row_stride,
                    const float* target, int64_t target_row_stride, const float* weight,
                    int32_t weight_mode, int64_t weight_row_stride, int64_t n, float scale,
                    float* loss_sum, float* row_loss, float* grad_pred, void* workspace,
                    size_t workspace_bytes, int32_t variant, int32_t flags, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || n < 0 || (flags & ~GD_FLAG_MASK_ZERO_WEIGHT) || weight_mode < GD_WEIGHT_NONE || weight_mode > GD_WEIGHT_ROW7 ||
      variant < GD_VARIANT_AUTO || variant > GD_VARIANT_BULK_PACKED)
    return GD_ERR_BAD_ARG;
  if (n > 0 && (!pred || !target || (weight_mode != GD_WEIGHT_NONE && !weight)))
    return GD_ERR_BAD_ARG;
  if (loss_sum && (!workspace || workspace_bytes < gd_loss_workspace_bytes(n)))
    return GD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0 && !loss_sum) return 0;       // empty batch, nothing to write
  const int wcols = weight_mode == GD_WEIGHT_ROW7 ? 7 : 1;
  const bool bulk_ok = pred_row_stride == 7 && target_row_stride == 7 && aligned16(pred) &&
                       aligned16(target) &&
                       (weight_mode == GD_WEIGHT_NONE ||
                        (weight_row_stride == wcols && aligned16(weight))) &&
                       (!grad_pred || aligned16(grad_pred)) &&
                       (!row_loss || aligned16(row_loss)) && n >= 4;
  // n == 0 with a loss_sum falls through to a 1-CTA staged launch that writes
  // scale * 0 (nan when scale is nan: torch's mean of an empty tensor).
  const bool want_bulk = variant == GD_VARIANT_BULK || variant == GD_VARIANT_BULK_R2 ||
                         variant == GD_VARIANT_BULK_PACKED;
  if (want_bulk && !bulk_ok) return GD_ERR_LAYOUT;
  const int v = want_bulk ? variant
                          : (variant == GD_VARIANT_AUTO && bulk_ok ? GD_VARIANT_BULK
                                                                   : GD_VARIANT_STAGED);

  LossArgs a;
  a.pred = pred;
  a.target = target;
  a.weight = weight;
  a.pstride = pred_row_stride;
  a.tstride = target_row_stride;
  a.wstride = weight_row_stride;
  a.n = n;
  a.wmode = weight_mode;
  a.mask_zero_w = (flags & GD_FLAG_MASK_ZERO_WEIGHT) ? 1 : 0;
  a.scale = scale;
  a.loss_sum = loss_sum;
  a.row_loss = row_loss;
  a.grad = grad_pred;
  a.ticket = reinterpret_cast<unsigned int*>(workspace);
  a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  a.pp = make_pair_params(*cfg);

  switch (cfg->loss_type) {
    case GD_LOSS_GWD3D: return launch_loss<gd::kGwd>(a, v, kMaxGrid, st);
    case GD_LOSS_KLD3D: return launch_loss<gd::kKld>(a, v, kMaxGrid, st);
    case GD_LOSS_JD3D: return launch_loss<gd::kJd>(a, v, kMaxGrid, st);
    case GD_LOSS_KLD3D_SYMMAX: return launch_loss<gd::kSymMax>(a, v, kMaxGrid, st);
    case GD_LOSS_KLD3D_SYMMIN: return launch_loss<gd::kSymMin>(a, v, kMaxGrid, st);
    case GD_LOSS_BD3D: return launch_loss<gd::kBd>(a, v, kMaxGrid, st);
    case GD_LOSS_KFIOU3D: return launch_loss<gd::kKfiou>(a, v, kMaxGrid, st);
  }
  return GD_ERR_BAD_ARG;
}

int gd_scale_grad(float* grad, int64_t n, const float* grad_output_scalar, void* stream) {
  using namespace gdk;
  if (n < 0 || (n > 0 && (!grad || !grad_output_scalar))) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  const long long nel = n * 7;
  long long grid = (nel / 4 + kThreads - 1) / kThreads;
  const long long cap = (long long)device_info().sm_count * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  gd_scale_grad_kernel<<<(int)grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad, nel, grad_output_scalar);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gd_scale_buffer(float* buf, int64_t count, const float* scalar, void* stream) {
  using namespace gdk;
  if (count < 0 || (count > 0 && (!buf || !scalar))) return GD_ERR_BAD_ARG;
  if (count == 0) return 0;
  long long grid = (count / 4 + kThreads - 1) / kThreads;
  const long long cap = (long long)device_info().sm_count * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  gd_scale_grad_kernel<<<(int)grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      buf, count, scalar);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gd_scale_grad_rows(float* grad, int64_t n, const float* grad_output_rows,
                       int64_t grad_output_stride, void* stream) {
  using namespace gdk;
  if (n < 0 || (n > 0 && (!grad || !grad_output_rows))) return GD_ERR_BAD_ARG;
  if (n == 0) return 0;
  const long long nel = n * 7;
  long long grid = (nel + kThreads - 1) / kThreads;
  const long long cap = (long long)device_info().sm_count * 8;
  if (grid > cap) grid = cap;
  gd_scale_grad_rows_kernel<<<(int)grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      grad, nel, grad_output_rows, grad_output_stride);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gd_any_positive(const float* weight, int64_t count, int32_t* flag, void* stream) {
  using namespace gdk;
  if (count < 0 || !flag || (count > 0 && !weight)) return GD_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  if (count == 0) return 0;
  long long grid = (count + kThreads - 1) / kThreads;
  const long long cap = (long long)device_info().sm_count * 8;
  if (grid > cap) grid = cap;
  gd_any_positive_kernel<<<(int)grid, kThreads, 0, st>>>(weight, count, flag);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int64_t gd_launch_count(void) { return gdk::g_launches.load(std::memory_order_relaxed); }

const char* gd_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case GD_ERR_BAD_ARG: return "gd_loss_b200: bad argument";
    case GD_ERR_WORKSPACE: return "gd_loss_b200: workspace missing or too small";
    case GD_ERR_LAYOUT: return "gd_loss_b200: bulk variant needs contiguous 16-byte aligned tensors";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gd_loss_b200: unknown error";
  }
}

}  // extern "C"
