// C ABI of the fused GD-loss kernels + the small helper kernels (autograd fold,
// early-return probe).  The heavy kernel templates live in gd_loss_kernels.cuh and
// are instantiated one loss type per translation unit (gd_loss_inst_*.cu) so the
// build parallelises.
#include <mutex>

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
  // The answer is almost always "yes" after the first few elements: every CTA re-reads the
  // flag (L2) before each grid-stride step and stops as soon as any CTA has set it, so the
  // probe costs a few microseconds instead of a pass over the weights.
  const long long stride = (long long)gridDim.x * kThreads;
  bool any = false;
  for (long long i0 = (long long)blockIdx.x * kThreads; i0 < count; i0 += stride) {
    // ONE thread looks at the flag; the vote makes the decision CTA-uniform
    const bool stop = threadIdx.x == 0 && *reinterpret_cast<volatile int*>(flag) != 0;
    const long long i = i0 + threadIdx.x;
    any = i < count && w[i] > 0.0f;
    const int vote = __syncthreads_or((any ? 1 : 0) | (stop ? 2 : 0));
    if ((vote & 1) && threadIdx.x == 0) atomicOr(flag, 1);
    if (vote) return;
  }
}

// positives of the anchor head's labels mode; ticket[0] = CTA ticket, ticket[2] = count
__global__ void __launch_bounds__(kThreads) gd_count_labels_kernel(
    const long long* __restrict__ labels, long long total, long long num_classes,
    float* __restrict__ out, unsigned int* ticket) {
  const long long stride = (long long)gridDim.x * kThreads;
  unsigned int c = 0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const long long lab = labels[i];
    c += (lab >= 0 && lab < num_classes) ? 1u : 0u;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(ticket + 2, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      const unsigned int tot = __ldcg(ticket + 2);
      *out = (float)(tot > 0u ? tot : 1u);
      ticket[2] = 0u;
      ticket[0] = 0u;
    }
  }
}

std::atomic<int> g_loss_grid{0};           // gd_set_loss_grid(); policy: warp_kernel_ctas()
std::atomic<int> g_loss_grid_cal[kMaxDevices];
thread_local int t_loss_grid_try = 0;
static std::mutex g_cal_mutex;

// Times `launch` (the launch the caller asked for: idempotent, it writes its outputs) on one CTA
// per SM and on a ladder of smaller grids (the bandwidth jump sat at 138, 132, 130, < 128 and 120
// CTAs on the boxes it was looked for: profiles/r03_grid.md) and keeps the fastest grid for the
// device.  What is to be measured is the power-capped steady state, which a cold board only
// reaches after tens of milliseconds of work (a first version that measured straight away kept
// 148 CTAs on a cold box: everything is fast before the cap bites): ~100 ms of launches first,
// then two interleaved rounds of ~5 ms per grid.  Host-synchronising, ~0.25 s once per device
// and process; any failure keeps one CTA per SM.
template <typename F>
static void calibrate_grid(F&& launch, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_cal_mutex);
  const int dev = current_device();
  if (g_loss_grid_cal[dev].load(std::memory_order_relaxed) > 0) return;
  const int sms = device_info().sm_count;
  constexpr int kCand = 7;
  const int ladder[kCand] = {148, 136, 132, 128, 124, 120, 116};   // of 148 SMs
  int cand[kCand];
  for (int c = 0; c < kCand; ++c) cand[c] = ladder[c] * sms / 148 > 0 ? ladder[c] * sms / 148 : 1;
  float ms[kCand] = {};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool ok = cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess;
  auto timed = [&](int reps, float* out) {       // `reps` launches, elapsed ms
    ok = ok && cudaEventRecord(e0, st) == cudaSuccess;
    for (int i = 0; ok && i < reps; ++i) ok = launch() == 0;
    ok = ok && cudaEventRecord(e1, st) == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess;
    ok = ok && cudaEventElapsedTime(out, e0, e1) == cudaSuccess;
  };
  // into the steady state, and how long one launch takes
  t_loss_grid_try = cand[0];
  float warm = 0.0f, one = 0.0f;
  timed(4, &one);
  one = one > 0.0f ? one / 4.0f : 1.0f;
  for (int batch = 0; ok && warm < 100.0f && batch < 64; ++batch) {
    float t = 0.0f;
    timed(16, &t);
    warm += t;
  }
  int reps = (int)(5.0f / one);                  // ~5 ms per measurement
  reps = reps < 4 ? 4 : (reps > 64 ? 64 : reps);
  for (int round = 0; ok && round < 2; ++round) {
    for (int k = 0; ok && k < kCand; ++k) {
      const int c = round == 0 ? k : kCand - 1 - k;          // second round in reverse order
      t_loss_grid_try = cand[c];
      for (int i = 0; ok && i < 2; ++i) ok = launch() == 0;
      float t = 0.0f;
      timed(reps, &t);
      ms[c] += t;
    }
  }
  t_loss_grid_try = 0;
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  int best = 0;
  if (ok) {
    for (int c = 1; c < kCand; ++c)
      if (ms[c] < ms[best] * 0.995f) best = c;               // a smaller grid has to win by 0.5 %
  } else {
    (void)cudaGetLastError();
  }
  g_loss_grid_cal[dev].store(cand[best], std::memory_order_relaxed);
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
  gd_loss_io io{};
  io.pred = pred;
  io.pred_row_stride = pred_row_stride;
  io.target = target;
  io.target_row_stride = target_row_stride;
  io.weight = weight;
  io.weight_mode = weight_mode;
  io.weight_row_stride = weight_row_stride;
  io.n = n;
  io.scale = scale;
  io.loss_sum = loss_sum;
  io.row_loss = row_loss;
  io.grad_pred = grad_pred;
  io.workspace = workspace;
  io.workspace_bytes = workspace_bytes;
  io.variant = variant;
  io.flags = flags;
  return gd_loss_launch(cfg, &io, stream);
}

int gd_loss_launch(const gd_loss_config* cfg, const gd_loss_io* io, void* stream) {
  using namespace gdk;
  if (!io) return GD_ERR_BAD_ARG;
  const int64_t n = io->n;
  const int32_t weight_mode = io->weight_mode, variant = io->variant, flags = io->flags;
  if (!config_ok(cfg) || n < 0 || (flags & ~GD_FLAG_MASK_ZERO_WEIGHT) ||
      weight_mode < GD_WEIGHT_NONE || weight_mode > GD_WEIGHT_ROW7 || variant < GD_VARIANT_AUTO ||
      variant > GD_VARIANT_BULK_ANY)
    return GD_ERR_BAD_ARG;
  if (n > 0 && (!io->pred || !io->target || (weight_mode != GD_WEIGHT_NONE && !io->weight)))
    return GD_ERR_BAD_ARG;
  if (io->status && !io->loss_sum) return GD_ERR_BAD_ARG;
  if (io->loss_sum && (!io->workspace || io->workspace_bytes < gd_loss_workspace_bytes(n)))
    return GD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0 && !io->loss_sum) return 0;       // empty batch, nothing to write
  const int wcols = weight_mode == GD_WEIGHT_ROW7 ? 7 : 1;
  const bool out_ok = (!io->grad_pred || aligned16(io->grad_pred)) &&
                      (!io->row_loss || aligned16(io->row_loss));
  const bool bulk_ok = io->pred_row_stride == 7 && io->target_row_stride == 7 &&
                       aligned16(io->pred) && aligned16(io->target) &&
                       (weight_mode == GD_WEIGHT_NONE ||
                        (io->weight_row_stride == wcols && aligned16(io->weight))) &&
                       out_ok && n >= 4;

  LossArgs a;
  a.pred = io->pred;
  a.target = io->target;
  a.weight = io->weight;
  a.pstride = io->pred_row_stride;
  a.tstride = io->target_row_stride;
  a.wstride = weight_mode == GD_WEIGHT_NONE ? 0 : io->weight_row_stride;
  a.n = n;
  a.wmode = weight_mode;
  a.mask_zero_w = (flags & GD_FLAG_MASK_ZERO_WEIGHT) ? 1 : 0;
  a.scale = io->scale;
  a.scale_div = io->scale_div;
  a.status = io->status;
  if (io->early_return) {
    if (!io->loss_sum || weight_mode == GD_WEIGHT_NONE || io->er_weight_row_stride < 0 ||
        io->er_weight_col_stride < 0)
      return GD_ERR_BAD_ARG;
    a.early_return = 1;
    a.er_wrow = io->er_weight_row_stride;
    a.er_wcol = io->er_weight_col_stride;
  }
  if (io->any_positive_host) {
    if (!io->loss_sum || weight_mode == GD_WEIGHT_NONE) return GD_ERR_BAD_ARG;
    a.host_flag = reinterpret_cast<int*>(io->any_positive_host);
  }
  a.loss_sum = io->loss_sum;
  a.row_loss = io->row_loss;
  a.grad = io->grad_pred;
  a.ticket = reinterpret_cast<unsigned int*>(io->workspace);
  a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(io->workspace) + 256);
  a.pp = make_pair_params(*cfg);
  if (io->peer_sum && io->peer_sum->world > 1) {
    const gd_peer_sum& ps = *io->peer_sum;
    if (!io->loss_sum || ps.world > GD_MAX_PEERS || ps.rank < 0 || ps.rank >= ps.world)
      return GD_ERR_BAD_ARG;
    for (int r = 0; r < ps.world; ++r)
      if (!ps.peer_buf[r]) return GD_ERR_BAD_ARG;
    a.peer = ps;
  }

  // Row-strided / unaligned inputs: bulk copies of whole wide rows (GD_VARIANT_BULK_ANY) when
  // the strides are sane and a useful number of warps fits the shared memory.
  bool any_ok = false;
  if (!bulk_ok && out_ok && n >= 16) {
    const bool strides_ok =
        a.pstride >= 7 && a.tstride >= 7 && a.pstride <= 64 && a.tstride <= 64 &&
        (weight_mode == GD_WEIGHT_NONE || (a.wstride >= wcols && a.wstride <= 64)) &&
        (reinterpret_cast<uintptr_t>(a.pred) & 3u) == 0 &&
        (reinterpret_cast<uintptr_t>(a.target) & 3u) == 0 &&
        (reinterpret_cast<uintptr_t>(a.weight) & 3u) == 0;
    if (strides_ok) {
      const WarpLayout L = warp_layout(4, weight_mode, a.grad != nullptr, a.row_loss != nullptr,
                                       true, a.pstride, a.tstride, a.wstride);
      any_ok = kSmemBudget / L.per_warp >= 6;
    }
  }
  // n == 0 with a loss_sum falls through to a 1-CTA staged launch that writes
  // scale * 0 (nan when scale is nan: torch's mean of an empty tensor).
  const bool want_bulk = variant == GD_VARIANT_BULK || variant == GD_VARIANT_BULK_R2 ||
                         variant == GD_VARIANT_BULK_PACKED;
  if (want_bulk && !bulk_ok) return GD_ERR_LAYOUT;
  if (variant == GD_VARIANT_BULK_ANY && !(any_ok || (bulk_ok && n >= 16))) return GD_ERR_LAYOUT;
  int v = variant;
  if (variant == GD_VARIANT_AUTO)
    v = bulk_ok ? ((GD_TUNE_DEFAULT & kTunePacked) ? GD_VARIANT_BULK_PACKED : GD_VARIANT_BULK)
                : (any_ok ? GD_VARIANT_BULK_ANY : GD_VARIANT_STAGED);
  if (v == GD_VARIANT_BULK_ANY) plan_any(&a);

  auto dispatch = [&]() -> int {
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
  };
  // grid of the persistent kernel: measured once per device (gd_loss_kernels.cuh, warp_kernel_ctas)
  const bool light = (v == GD_VARIANT_BULK || v == GD_VARIANT_BULK_PACKED || v == GD_VARIANT_BULK_R2) &&
                     weight_mode != GD_WEIGHT_ROW7;
  if (light && n >= (1LL << 22) && g_loss_grid.load(std::memory_order_relaxed) == 0 &&
      !loss_multi_process() && !a.peer.world &&
      g_loss_grid_cal[current_device()].load(std::memory_order_relaxed) == 0) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone)
      calibrate_grid(dispatch, st);
  }
  return dispatch();
}

size_t gd_peer_sum_buffer_bytes(void) { return sizeof(gdk::PeerBuf); }

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

int gd_host_flag_wait(const int32_t* flag, void* stream) {
  if (!flag) return GD_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (unsigned long long spin = 1;; ++spin) {
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) != 0) return 0;
    if ((spin & 0x3fffu) == 0) {           // now and then: did the launch die / the stream drain?
      const cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess) {              // everything queued has run: the flag is final
        return __atomic_load_n(flag, __ATOMIC_ACQUIRE) != 0 ? 0 : GD_ERR_BAD_ARG;
      }
      if (e != cudaErrorNotReady) return (int)e;
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

int gd_count_positive_labels(const int64_t* labels, int64_t total, int64_t num_classes, float* out,
                             void* workspace, size_t workspace_bytes, void* stream) {
  using namespace gdk;
  if (total < 0 || !out || (total > 0 && !labels)) return GD_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < gd_loss_workspace_bytes(total)) return GD_ERR_WORKSPACE;
  long long grid = (total + kThreads * 4 - 1) / (kThreads * 4);
  const long long cap = (long long)device_info().sm_count * 4;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  gd_count_labels_kernel<<<(int)grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(labels), total, num_classes, out,
      reinterpret_cast<unsigned int*>(workspace));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gd_set_loss_grid(int32_t ctas) {
  if (ctas < 0) return GD_ERR_BAD_ARG;
  gdk::g_loss_grid.store(ctas, std::memory_order_relaxed);
  return 0;
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
