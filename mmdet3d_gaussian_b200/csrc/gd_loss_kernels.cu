// Fused forward+backward Gaussian-distance loss kernels for B200 (sm_100a).
//
// Replaces, in ONE pass over HBM (88 algorithmic bytes per box pair: pred 28 +
// target 28 + weight 4 read, grad 28 written), what the reference does with
// ~160-250 eager torch ops and autograd: GDLoss.forward
// (mmdet3d_gaussian/models/losses/gaussian_distance_loss.py:280-310, "ref")
// = preprocess x2 (ref:8-21) -> distance (ref:42-248) -> postprocess (ref:24-39)
// -> weighted reduction (mmdet weight_reduce_loss) -> x loss_weight, plus
// d loss / d pred.
//
// Two kernels, same per-row math (gd_math.cuh):
//   * gd_bulk_kernel   -- persistent; a 3-stage ring of [256-row] tiles filled by
//                         1-D bulk async copies (TMA engine, cp.async.bulk +
//                         mbarrier complete_tx), rows read from shared memory at
//                         stride 7 (odd => bank-conflict free), gradients staged
//                         in shared memory and written back with bulk stores.
//                         Needs contiguous, 16-byte aligned tensors.
//   * gd_staged_kernel -- same tiling with plain (vector when possible) loads
//                         and stores; takes any row stride / alignment (the
//                         CenterGDHead call site passes row-strided views).
// The AoS [N,7] layout is the reference's contract; the transpose to
// one-row-per-thread happens in shared memory, never in HBM.
//
// Loss sum: per-thread fp32 -> warp shuffle -> per-CTA fp64 partial -> the last
// CTA (atomic ticket, one atomic per CTA) adds the partials in fixed order, so
// the result is deterministic for a given grid.
#include "gd_common.cuh"

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

struct LossArgs {
  const float* pred;
  const float* target;
  const float* weight;
  long long pstride, tstride, wstride;   // row strides in elements
  long long n;
  int wmode;
  int mask_zero_w;                        // GD_FLAG_MASK_ZERO_WEIGHT
  float scale;
  float* loss_sum;
  float* row_loss;
  float* grad;
  double* partials;                       // [grid]
  unsigned int* ticket;                   // zero on entry, zero again on exit
  gd::PairParams<float> pp;
};

// ---------------------------------------------------------------------------
// deterministic grid-wide sum
// ---------------------------------------------------------------------------
__device__ __forceinline__ void finish_sum(float acc, const LossArgs& a) {
  __shared__ float s_warp[kThreads / 32];
  __shared__ double s_dwarp[kThreads / 32];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  acc = warp_sum(acc);
  if (lane == 0) s_warp[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += (double)s_warp[w];
    a.partials[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double s = 0.0;
    for (unsigned int i = tid; i < gridDim.x; i += kThreads) s += __ldcg(a.partials + i);
    s = warp_sum(s);
    if (lane == 0) s_dwarp[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) tot += s_dwarp[w];
      *a.loss_sum = (float)(tot * (double)a.scale);
      *a.ticket = 0u;                     // leave the workspace reusable
    }
  }
}

// weight of one row: None -> 1, [N] -> w, [N,7] -> mean(-1)          ref:295-296
__device__ __forceinline__ float row_weight_smem(const float* sw, int wmode, int r) {
  if (wmode == GD_WEIGHT_ROW) return sw[r];
  if (wmode == GD_WEIGHT_ROW7) {
    const float* w = sw + 7 * r;
    return (((((w[0] + w[1]) + w[2]) + w[3]) + w[4]) + w[5] + w[6]) / 7.0f;
  }
  return 1.0f;
}

// One row, registers only.  Returns w_i * loss_i (unscaled) for the sum;
// g[] <- scale * w_i * dloss_i/dpred_i, *rl <- scale * w_i * loss_i.
template <int LOSS, bool GRAD>
__device__ __forceinline__ float eval_row(const float* p, const float* t, float w,
                                          const LossArgs& a, float* g, float* rl) {
  float l = gd::pair_eval<float, LOSS, GRAD>(p, t, a.pp, g);
  if (a.mask_zero_w && w == 0.0f) {        // masked row: exactly zero, nan/inf do not leak
    l = 0.0f;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) g[c] = 0.0f;
    }
  }
  const float ws = w * a.scale;
  if (GRAD) {
#pragma unroll
    for (int c = 0; c < 7; ++c) g[c] *= ws;
  }
  *rl = l * ws;
  return l * w;
}

// ---------------------------------------------------------------------------
// staged kernel: any stride / alignment
// ---------------------------------------------------------------------------
__device__ __forceinline__ void load_rows(float* __restrict__ s, const float* __restrict__ g,
                                          long long stride, int cols, long long row0, int rows,
                                          int tid) {
  const int nel = rows * cols;
  if (stride == cols) {
    const float* base = g + row0 * cols;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
      const int nv = nel >> 2;
      const float4* b4 = reinterpret_cast<const float4*>(base);
      float4* s4 = reinterpret_cast<float4*>(s);
      for (int i = tid; i < nv; i += kThreads) s4[i] = __ldcs(b4 + i);
      for (int i = (nv << 2) + tid; i < nel; i += kThreads) s[i] = __ldcs(base + i);
    } else {
      for (int i = tid; i < nel; i += kThreads) s[i] = __ldcs(base + i);
    }
  } else {
    for (int i = tid; i < nel; i += kThreads) {
      const int r = i / cols, c = i - r * cols;
      s[i] = __ldcs(g + (row0 + r) * stride + c);
    }
  }
}

__device__ __forceinline__ void store_rows(float* __restrict__ g, const float* __restrict__ s,
                                           int cols, long long row0, int rows, int tid) {
  const int nel = rows * cols;
  float* base = g + row0 * cols;
  if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
    const int nv = nel >> 2;
    float4* b4 = reinterpret_cast<float4*>(base);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = tid; i < nv; i += kThreads) __stcs(b4 + i, s4[i]);
    for (int i = (nv << 2) + tid; i < nel; i += kThreads) __stcs(base + i, s[i]);
  } else {
    for (int i = tid; i < nel; i += kThreads) __stcs(base + i, s[i]);
  }
}

template <int LOSS, bool GRAD>
__global__ void __launch_bounds__(kThreads) gd_staged_kernel(const LossArgs a) {
  __shared__ __align__(16) float s_pred[kTile * 7];    // reused for the gradient tile
  __shared__ __align__(16) float s_tgt[kTile * 7];
  __shared__ __align__(16) float s_w[kTile * 7];       // [N,7] weights only
  const int tid = threadIdx.x;
  const long long ntiles = (a.n + kTile - 1) / kTile;
  float acc = 0.0f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kTile;
    const int rows = (int)min((long long)kTile, a.n - row0);
    load_rows(s_pred, a.pred, a.pstride, 7, row0, rows, tid);
    load_rows(s_tgt, a.target, a.tstride, 7, row0, rows, tid);
    if (a.wmode == GD_WEIGHT_ROW7) load_rows(s_w, a.weight, a.wstride, 7, row0, rows, tid);
    float w1 = 1.0f;
    if (a.wmode == GD_WEIGHT_ROW && tid < rows) w1 = __ldcs(a.weight + (row0 + tid) * a.wstride);
    __syncthreads();
    if (tid < rows) {
      float p[7], t[7], g[7], rl;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        p[c] = s_pred[7 * tid + c];
        t[c] = s_tgt[7 * tid + c];
      }
      const float w = (a.wmode == GD_WEIGHT_ROW7) ? row_weight_smem(s_w, a.wmode, tid) : w1;
      acc += eval_row<LOSS, GRAD>(p, t, w, a, g, &rl);
      if (a.row_loss) __stcs(a.row_loss + row0 + tid, rl);
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < 7; ++c) s_pred[7 * tid + c] = g[c];   // own row only: no hazard
      }
    }
    __syncthreads();
    if (GRAD) store_rows(a.grad, s_pred, 7, row0, rows, tid);
    __syncthreads();
  }
  if (a.loss_sum) finish_sum(acc, a);
}

// ---------------------------------------------------------------------------
// bulk pipeline kernel: persistent CTAs, STAGES-deep ring of tiles
// ---------------------------------------------------------------------------
struct BulkLayout {
  int wtile;          // bytes of the weight tile (0, 1024 or 7168)
  int stage;          // bytes per stage
  int out;            // bytes per output buffer
  int total;          // dynamic shared memory bytes
};

__host__ __device__ inline BulkLayout bulk_layout(int wmode, bool grad, bool rows, int stages) {
  BulkLayout L;
  L.wtile = wmode == GD_WEIGHT_ROW7 ? kTileBytes : (wmode == GD_WEIGHT_ROW ? kTile * 4 : 0);
  L.stage = 2 * kTileBytes + L.wtile;
  L.out = (grad ? kTileBytes : 0) + (rows ? kTile * 4 : 0);
  L.total = stages * L.stage + 2 * L.out + stages * 8;
  return L;
}

template <int LOSS, bool GRAD, int STAGES>
__global__ void __launch_bounds__(kThreads) gd_bulk_kernel(const LossArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const bool want_rows = a.row_loss != nullptr;
  const BulkLayout L = bulk_layout(a.wmode, GRAD, want_rows, STAGES);
  unsigned char* out_base = smem + STAGES * L.stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_base + 2 * L.out);

  // rows the bulk path can move: a multiple of 4 rows keeps every copy a multiple of 16 B
  const long long n_main = a.n & ~3LL;
  const long long ntiles = (n_main + kTile - 1) / kTile;
  const int my_n = (ntiles > (long long)blockIdx.x)
                       ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1
                       : 0;
  const int wcols = a.wmode == GD_WEIGHT_ROW7 ? 7 : 1;
  uint64_t policy = 0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    policy = policy_evict_first();
  }
  __syncthreads();

  auto issue = [&](int i) {               // thread 0 only
    const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
    const long long row0 = tile * kTile;
    const uint32_t rows = (uint32_t)min((long long)kTile, n_main - row0);
    const int s = i % STAGES;
    unsigned char* st = smem + s * L.stage;
    const uint32_t box_bytes = rows * kRowBytes;
    const uint32_t w_bytes = a.wmode ? rows * wcols * 4u : 0u;
    mbar_arrive_expect_tx(&bars[s], 2 * box_bytes + w_bytes);
    bulk_load(st, a.pred + row0 * 7, box_bytes, &bars[s], policy);
    bulk_load(st + kTileBytes, a.target + row0 * 7, box_bytes, &bars[s], policy);
    if (a.wmode) bulk_load(st + 2 * kTileBytes, a.weight + row0 * wcols, w_bytes, &bars[s], policy);
  };

  if (tid == 0) {
    const int pre = my_n < STAGES ? my_n : STAGES;
    for (int i = 0; i < pre; ++i) issue(i);
  }

  float acc = 0.0f;
  for (int i = 0; i < my_n; ++i) {
    const int s = i % STAGES;
    const long long tile = (long long)blockIdx.x + (long long)i * gridDim.x;
    const long long row0 = tile * kTile;
    const int rows = (int)min((long long)kTile, n_main - row0);
    const unsigned char* st = smem + s * L.stage;
    const float* sp = reinterpret_cast<const float*>(st);
    const float* stg = reinterpret_cast<const float*>(st + kTileBytes);
    const float* sw = reinterpret_cast<const float*>(st + 2 * kTileBytes);

    mbar_wait(&bars[s], (uint32_t)((i / STAGES) & 1));
    float p[7], t[7], w = 1.0f;
    if (tid < rows) {
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        p[c] = sp[7 * tid + c];
        t[c] = stg[7 * tid + c];
      }
      w = row_weight_smem(sw, a.wmode, tid);
    }
    // the store issued two tiles ago read from the output buffer we are about to reuse
    if (tid == 0 && L.out) bulk_wait_read<1>();
    __syncthreads();                      // stage s fully consumed; out buffer free
    if (tid == 0 && i + STAGES < my_n) issue(i + STAGES);

    unsigned char* ob = out_base + (i & 1) * L.out;
    float* og = reinterpret_cast<float*>(ob);
    float* orow = reinterpret_cast<float*>(ob + (GRAD ? kTileBytes : 0));
    if (tid < rows) {
      float g[7], rl;
      acc += eval_row<LOSS, GRAD>(p, t, w, a, g, &rl);
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < 7; ++c) og[7 * tid + c] = g[c];
      }
      if (want_rows) orow[tid] = rl;
    }
    if (L.out) {
      fence_proxy_async_smem();           // generic-proxy writes -> visible to the bulk engine
      __syncthreads();
      if (tid == 0) {
        if (GRAD) bulk_store(a.grad + row0 * 7, og, (uint32_t)rows * kRowBytes);
        if (want_rows) bulk_store(a.row_loss + row0, orow, (uint32_t)rows * 4u);
        bulk_commit();
      }
    }
  }

  // <= 3 leftover rows (n % 4): block 0, straight from global memory
  if (blockIdx.x == 0 && tid < (int)(a.n - n_main)) {
    const long long r = n_main + tid;
    float p[7], t[7], g[7], rl, w = 1.0f;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      p[c] = a.pred[r * 7 + c];
      t[c] = a.target[r * 7 + c];
    }
    if (a.wmode == GD_WEIGHT_ROW) w = a.weight[r];
    if (a.wmode == GD_WEIGHT_ROW7) w = row_weight_smem(a.weight + r * 7, GD_WEIGHT_ROW7, 0);
    acc += eval_row<LOSS, GRAD>(p, t, w, a, g, &rl);
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) a.grad[r * 7 + c] = g[c];
    }
    if (want_rows) a.row_loss[r] = rl;
  }
  if (tid == 0 && L.out) bulk_wait_all<0>();
  if (a.loss_sum) finish_sum(acc, a);
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

// ---------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------
constexpr int kStages = 3;

template <int LOSS, bool GRAD>
int launch_staged(const LossArgs& a, int grid, cudaStream_t stream) {
  gd_staged_kernel<LOSS, GRAD><<<grid, kThreads, 0, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS, bool GRAD>
int bulk_ctas_per_sm(int smem_bytes) {
  int n = 0;
  cudaFuncSetAttribute(gd_bulk_kernel<LOSS, GRAD, kStages>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gd_bulk_kernel<LOSS, GRAD, kStages>,
                                                kThreads, smem_bytes);
  return n > 0 ? n : 1;
}

template <int LOSS, bool GRAD>
int launch_bulk(const LossArgs& a, int max_grid, cudaStream_t stream) {
  const BulkLayout L = bulk_layout(a.wmode, GRAD, a.row_loss != nullptr, kStages);
  static int occ_cache[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // [wmode][rows]
  int& occ = occ_cache[a.wmode][a.row_loss ? 1 : 0];
  if (occ == 0) occ = bulk_ctas_per_sm<LOSS, GRAD>(L.total);
  const long long ntiles = ((a.n & ~3LL) + kTile - 1) / kTile;
  long long grid = (long long)device_info().sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  gd_bulk_kernel<LOSS, GRAD, kStages><<<(int)grid, kThreads, L.total, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS>
int launch_loss(const LossArgs& a, bool bulk, int max_grid, cudaStream_t stream) {
  const bool grad = a.grad != nullptr;
  if (bulk) {
    return grad ? launch_bulk<LOSS, true>(a, max_grid, stream)
                : launch_bulk<LOSS, false>(a, max_grid, stream);
  }
  long long grid = (a.n + kTile - 1) / kTile;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  return grad ? launch_staged<LOSS, true>(a, (int)grid, stream)
              : launch_staged<LOSS, false>(a, (int)grid, stream);
}

constexpr int kMaxGrid = 65536;           // partials capacity of the workspace

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
      variant < GD_VARIANT_AUTO || variant > GD_VARIANT_BULK)
    return GD_ERR_BAD_ARG;
  if (n > 0 && (!pred || !target || (weight_mode != GD_WEIGHT_NONE && !weight)))
    return GD_ERR_BAD_ARG;
  if (loss_sum && (!workspace || workspace_bytes < gd_loss_workspace_bytes(n)))
    return GD_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {                            // empty batch: the sum of nothing
    if (loss_sum) {
      const cudaError_t e = cudaMemsetAsync(loss_sum, 0, sizeof(float), st);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }
  const int wcols = weight_mode == GD_WEIGHT_ROW7 ? 7 : 1;
  const bool bulk_ok = pred_row_stride == 7 && target_row_stride == 7 && aligned16(pred) &&
                       aligned16(target) &&
                       (weight_mode == GD_WEIGHT_NONE ||
                        (weight_row_stride == wcols && aligned16(weight))) &&
                       (!grad_pred || aligned16(grad_pred)) &&
                       (!row_loss || aligned16(row_loss)) && n >= 4;
  if (variant == GD_VARIANT_BULK && !bulk_ok) return GD_ERR_LAYOUT;
  const bool bulk = variant == GD_VARIANT_BULK || (variant == GD_VARIANT_AUTO && bulk_ok);

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
    case GD_LOSS_GWD3D: return launch_loss<gd::kGwd>(a, bulk, kMaxGrid, st);
    case GD_LOSS_KLD3D: return launch_loss<gd::kKld>(a, bulk, kMaxGrid, st);
    case GD_LOSS_JD3D: return launch_loss<gd::kJd>(a, bulk, kMaxGrid, st);
    case GD_LOSS_KLD3D_SYMMAX: return launch_loss<gd::kSymMax>(a, bulk, kMaxGrid, st);
    case GD_LOSS_KLD3D_SYMMIN: return launch_loss<gd::kSymMin>(a, bulk, kMaxGrid, st);
    case GD_LOSS_BD3D: return launch_loss<gd::kBd>(a, bulk, kMaxGrid, st);
    case GD_LOSS_KFIOU3D: return launch_loss<gd::kKfiou>(a, bulk, kMaxGrid, st);
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
