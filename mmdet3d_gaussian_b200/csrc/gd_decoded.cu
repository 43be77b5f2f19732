// Head front ends: positive-row gather + box decode + GD loss + gradient to the raw
// network outputs in ONE launch.
//
// Replaces, for the reference's two call sites,
//   GDAnchor3DHead.loss_single   models/dense_heads/gd_anchor3d_head.py:102-112 (nonzero +
//                                four row gathers), :128-131 (weights), :133-136 (decode x2),
//                                :137-141 (GDLoss) and the autograd backward of all of it
//                                (index_put into a zero [T,7] gradient);
//   CenterGDHead.loss            models/dense_heads/gd_centerpoint_head.py:421-423 (decode) +
//                                :433-434 (GDLoss) and their backward.
// Per-pair math: gd_decode.cuh (decode prologue / Jacobian epilogue) around gd_math.cuh.
//
// Anchor head, two row-selection modes:
//   * index mode  -- pos_inds[P] given (the reference's `nonzero` result): one thread per
//                    positive; gradient rows written compact [P,7] or scattered into a
//                    caller-zeroed dense [T,7].
//   * mask mode   -- labels[T] given: positives are 0 <= label < num_classes
//                    (gd_anchor3d_head.py:102-104) decided in-kernel, so there is no
//                    `nonzero`, no device->host sync and no separate zero-fill: the kernel
//                    streams the labels (8 B/row) and writes the whole dense gradient
//                    (28 B/row, zero rows for negatives) with coalesced 16-byte stores.
//                    Algorithmic bytes: 36 B per anchor row + 112 B per positive.
#include "gd_decode.cuh"
#include "gd_loss_kernels.cuh"

namespace gdk {

struct AnchorArgs {
  const float* anchors;
  long long anchor_rows;                 // anchors repeat with this period (:110-111)
  const float* dpred;
  const float* dtarget;
  const float* bbox_w;                   // nullable [T,7]
  long long dp_stride, dt_stride, bw_stride;
  float decode_w[7];
  const long long* pos_inds;
  long long num_pos;
  const long long* labels;
  long long num_classes;
  long long total_rows;
  float* grad;
  int grad_mode;
  LossArgs sum;                          // scale, loss_sum, partials, ticket, pp, mask_zero_w
};

// weight of a positive row: mean_c(bbox_weights[i,c] * decode_weight[c])
// (gd_anchor3d_head.py:128-131 then gaussian_distance_loss.py:295-296)
__device__ __forceinline__ float anchor_row_weight(const AnchorArgs& a, long long i) {
  if (!a.bbox_w) return 1.0f;
  const float* w = a.bbox_w + i * a.bw_stride;
  float s = __ldg(w) * a.decode_w[0];
#pragma unroll
  for (int c = 1; c < 7; ++c) s += __ldg(w + c) * a.decode_w[c];
  return s / 7.0f;
}

template <int LOSS, bool GRAD>
__device__ __forceinline__ float anchor_row(const AnchorArgs& a, long long i, float scale, float* g) {
  float an[7], dp[7], dt[7];
  const float* ar = a.anchors + (i % a.anchor_rows) * 7;
  const float* pr = a.dpred + i * a.dp_stride;
  const float* tr = a.dtarget + i * a.dt_stride;
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    an[c] = __ldg(ar + c);
    dp[c] = __ldg(pr + c);
    dt[c] = __ldg(tr + c);
  }
  const float w = anchor_row_weight(a, i);
  const float ws = w * scale;
  float l = gd::anchor_pair_eval<float, LOSS, GRAD>(an, dp, dt, a.sum.pp, ws, g);
  if (a.sum.mask_zero_w && w == 0.0f) {
    l = 0.0f;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) g[c] = 0.0f;
    }
  }
  return l * w;
}

template <int LOSS, bool GRAD>
__global__ void __launch_bounds__(kThreads) gd_anchor_index_kernel(const AnchorArgs a) {
  const float scale = effective_scale(a.sum);
  float acc = 0.0f;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < a.num_pos; k += stride) {
    const long long i = a.pos_inds[k];
    if (i < 0 || i >= a.total_rows) continue;      // never dereference a bad index
    float g[7];
    acc += anchor_row<LOSS, GRAD>(a, i, scale, g);
    if (GRAD) {
      float* o = a.grad_mode == GD_GRAD_SCATTER ? a.grad + i * 7 : a.grad + k * 7;
#pragma unroll
      for (int c = 0; c < 7; ++c) o[c] = g[c];
    }
  }
  if (a.sum.loss_sum) finish_sum(acc, a.sum, scale);
}

template <int LOSS, bool GRAD>
__global__ void __launch_bounds__(kThreads) gd_anchor_mask_kernel(const AnchorArgs a) {
  const float scale = effective_scale(a.sum);
  __shared__ __align__(16) float s_g[kTile * 7];
  const int tid = threadIdx.x;
  const long long ntiles = (a.total_rows + kTile - 1) / kTile;
  float acc = 0.0f;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kTile;
    const int rows = (int)min((long long)kTile, a.total_rows - row0);
    const long long i = row0 + tid;
    float g[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) g[c] = 0.0f;
    if (tid < rows) {
      const long long lab = __ldcs(a.labels + i);
      if (lab >= 0 && lab < a.num_classes) acc += anchor_row<LOSS, GRAD>(a, i, scale, g);
    }
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 7; ++c) s_g[7 * tid + c] = g[c];
      __syncthreads();
      store_rows(a.grad, s_g, 7, row0, rows, tid);
      __syncthreads();
    }
  }
  if (a.sum.loss_sum) finish_sum(acc, a.sum, scale);
}

struct CenterArgs {
  const float* preds;
  const long long* locs;
  const float* target;
  const float* weight;
  long long p_stride, l_stride, t_stride, w_stride;
  int wmode;
  long long n;
  float* grad;
  long long g_stride;
  int g_cols;
  gd::CenterDecodeParams dec;
  LossArgs sum;
};

template <int LOSS, bool GRAD>
__global__ void __launch_bounds__(kThreads) gd_center_kernel(const CenterArgs a) {
  const float scale = effective_scale(a.sum);
  float acc = 0.0f;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < a.n; k += stride) {
    float pr[7], t[7], g[7];
    const float* p = a.preds + k * a.p_stride;
    const float* tr = a.target + k * a.t_stride;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      pr[c] = __ldg(p + c);
      t[c] = __ldg(tr + c);
    }
    const long long lx = a.locs[k * a.l_stride], ly = a.locs[k * a.l_stride + 1];
    float w = 1.0f;
    if (a.wmode == GD_WEIGHT_ROW) w = __ldg(a.weight + k * a.w_stride);
    if (a.wmode == GD_WEIGHT_ROW7) w = row_weight_smem(a.weight + k * a.w_stride, GD_WEIGHT_ROW7, 0);
    const float ws = w * scale;
    float l = gd::center_pair_eval<float, LOSS, GRAD>(pr, lx, ly, t, a.dec, a.sum.pp, ws, g);
    if (a.sum.mask_zero_w && w == 0.0f) {
      l = 0.0f;
      if (GRAD) {
#pragma unroll
        for (int c = 0; c < 7; ++c) g[c] = 0.0f;
      }
    }
    acc += l * w;
    if (GRAD) {
      float* o = a.grad + k * a.g_stride;
#pragma unroll
      for (int c = 0; c < 7; ++c) o[c] = g[c];
      for (int c = 7; c < a.g_cols; ++c) o[c] = 0.0f;   // dir / vel columns take no GD gradient
    }
  }
  if (a.sum.loss_sum) finish_sum(acc, a.sum, scale);
}

template <int LOSS>
int launch_anchor(const AnchorArgs& a, cudaStream_t st) {
  const bool grad = a.grad != nullptr;
  const long long sms = device_info().sm_count;
  if (a.labels) {
    long long grid = (a.total_rows + kTile - 1) / kTile;
    if (grid > sms * 8) grid = sms * 8;
    if (grid < 1) grid = 1;
    if (grad) gd_anchor_mask_kernel<LOSS, true><<<(int)grid, kThreads, 0, st>>>(a);
    else gd_anchor_mask_kernel<LOSS, false><<<(int)grid, kThreads, 0, st>>>(a);
  } else {
    long long grid = (a.num_pos + kThreads - 1) / kThreads;
    if (grid > sms * 8) grid = sms * 8;
    if (grid < 1) grid = 1;
    if (grad) gd_anchor_index_kernel<LOSS, true><<<(int)grid, kThreads, 0, st>>>(a);
    else gd_anchor_index_kernel<LOSS, false><<<(int)grid, kThreads, 0, st>>>(a);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

template <int LOSS>
int launch_center(const CenterArgs& a, cudaStream_t st) {
  long long grid = (a.n + kThreads - 1) / kThreads;
  const long long sms = device_info().sm_count;
  if (grid > sms * 8) grid = sms * 8;
  if (grid < 1) grid = 1;
  if (a.grad) gd_center_kernel<LOSS, true><<<(int)grid, kThreads, 0, st>>>(a);
  else gd_center_kernel<LOSS, false><<<(int)grid, kThreads, 0, st>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

static bool fill_sum(LossArgs* s, const gd_loss_config* cfg, float scale, const float* scale_div,
                     float* loss_sum, void* workspace, size_t workspace_bytes, int32_t flags) {
  if (loss_sum && (!workspace || workspace_bytes < gd_loss_workspace_bytes(0))) return false;
  *s = LossArgs{};
  s->scale = scale;
  s->scale_div = scale_div;
  s->loss_sum = loss_sum;
  s->mask_zero_w = (flags & GD_FLAG_MASK_ZERO_WEIGHT) ? 1 : 0;
  s->ticket = reinterpret_cast<unsigned int*>(workspace);
  s->partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + 256);
  s->pp = make_pair_params(*cfg);
  return true;
}

}  // namespace gdk

#define GD_DISPATCH_LOSS(FN, ...)                                        \
  switch (cfg->loss_type) {                                              \
    case GD_LOSS_GWD3D: return FN<gd::kGwd>(__VA_ARGS__);                \
    case GD_LOSS_KLD3D: return FN<gd::kKld>(__VA_ARGS__);                \
    case GD_LOSS_JD3D: return FN<gd::kJd>(__VA_ARGS__);                  \
    case GD_LOSS_KLD3D_SYMMAX: return FN<gd::kSymMax>(__VA_ARGS__);      \
    case GD_LOSS_KLD3D_SYMMIN: return FN<gd::kSymMin>(__VA_ARGS__);      \
    case GD_LOSS_BD3D: return FN<gd::kBd>(__VA_ARGS__);                  \
    case GD_LOSS_KFIOU3D: return FN<gd::kKfiou>(__VA_ARGS__);            \
  }                                                                      \
  return GD_ERR_BAD_ARG;

extern "C" {

int gd_anchor_decoded_loss_fwd_bwd(const gd_loss_config* cfg, const float* anchors,
                                   int64_t anchor_rows, const float* deltas_pred,
                                   int64_t deltas_pred_row_stride, const float* deltas_target,
                                   int64_t deltas_target_row_stride, const float* bbox_weights,
                                   int64_t bbox_weights_row_stride,
                                   const float* decode_weight_host, const int64_t* pos_inds,
                                   int64_t num_pos, const int64_t* labels, int64_t num_classes,
                                   int64_t total_rows, float scale, const float* scale_div,
                                   float* loss_sum, float* grad_deltas, int32_t grad_mode, void* workspace,
                                   size_t workspace_bytes, int32_t flags, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || total_rows < 0 || num_pos < 0 || anchor_rows <= 0 ||
      (flags & ~GD_FLAG_MASK_ZERO_WEIGHT) || grad_mode < GD_GRAD_NONE ||
      grad_mode > GD_GRAD_DENSE)
    return GD_ERR_BAD_ARG;
  // exactly one of labels / pos_inds selects the rows (pos_inds may be null when num_pos == 0)
  const bool mask_mode = labels != nullptr;
  if (!mask_mode && num_pos > 0 && !pos_inds) return GD_ERR_BAD_ARG;
  if (mask_mode && pos_inds) return GD_ERR_BAD_ARG;
  const bool want_grad = grad_mode != GD_GRAD_NONE;
  if (want_grad && !grad_deltas) return GD_ERR_BAD_ARG;
  if (mask_mode && want_grad && grad_mode != GD_GRAD_DENSE) return GD_ERR_BAD_ARG;
  if (!mask_mode && grad_mode == GD_GRAD_DENSE) return GD_ERR_BAD_ARG;
  const long long work = mask_mode ? total_rows : num_pos;
  if (work > 0 && (!anchors || !deltas_pred || !deltas_target)) return GD_ERR_BAD_ARG;
  if (bbox_weights && !decode_weight_host) return GD_ERR_BAD_ARG;
  AnchorArgs a{};
  if (!fill_sum(&a.sum, cfg, scale, scale_div, loss_sum, workspace, workspace_bytes, flags))
    return GD_ERR_WORKSPACE;
  if (work == 0 && !loss_sum) return 0;
  a.anchors = anchors;
  a.anchor_rows = anchor_rows;
  a.dpred = deltas_pred;
  a.dtarget = deltas_target;
  a.bbox_w = bbox_weights;
  a.dp_stride = deltas_pred_row_stride;
  a.dt_stride = deltas_target_row_stride;
  a.bw_stride = bbox_weights_row_stride;
  for (int c = 0; c < 7; ++c) a.decode_w[c] = bbox_weights ? decode_weight_host[c] : 1.0f;
  a.pos_inds = reinterpret_cast<const long long*>(pos_inds);
  a.num_pos = num_pos;
  a.labels = reinterpret_cast<const long long*>(labels);
  a.num_classes = num_classes;
  a.total_rows = total_rows;
  a.grad = want_grad ? grad_deltas : nullptr;
  a.grad_mode = grad_mode;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GD_DISPATCH_LOSS(launch_anchor, a, st)
}

int gd_center_decoded_loss_fwd_bwd(const gd_loss_config* cfg, const gd_center_coder* coder,
                                   const float* preds, int64_t preds_row_stride,
                                   const int64_t* locs, int64_t locs_row_stride,
                                   const float* target, int64_t target_row_stride,
                                   const float* weight, int32_t weight_mode,
                                   int64_t weight_row_stride, int64_t n, float scale,
                                   const float* scale_div, float* loss_sum, float* grad_preds,
                                   int64_t grad_row_stride,
                                   int32_t grad_cols, void* workspace, size_t workspace_bytes,
                                   int32_t flags, void* stream) {
  using namespace gdk;
  if (!config_ok(cfg) || !coder || n < 0 || (flags & ~GD_FLAG_MASK_ZERO_WEIGHT) ||
      weight_mode < GD_WEIGHT_NONE || weight_mode > GD_WEIGHT_ROW7)
    return GD_ERR_BAD_ARG;
  if (n > 0 && (!preds || !locs || !target || (weight_mode != GD_WEIGHT_NONE && !weight)))
    return GD_ERR_BAD_ARG;
  if (grad_preds && (grad_cols < 7 || grad_row_stride < grad_cols)) return GD_ERR_BAD_ARG;
  CenterArgs a{};
  if (!fill_sum(&a.sum, cfg, scale, scale_div, loss_sum, workspace, workspace_bytes, flags))
    return GD_ERR_WORKSPACE;
  if (n == 0 && !loss_sum) return 0;
  a.preds = preds;
  a.locs = reinterpret_cast<const long long*>(locs);
  a.target = target;
  a.weight = weight;
  a.p_stride = preds_row_stride;
  a.l_stride = locs_row_stride;
  a.t_stride = target_row_stride;
  a.w_stride = weight_row_stride;
  a.wmode = weight_mode;
  a.n = n;
  a.grad = grad_preds;
  a.g_stride = grad_row_stride;
  a.g_cols = grad_cols;
  a.dec.sx = (double)coder->out_size_factor * (double)coder->voxel_size[0];
  a.dec.sy = (double)coder->out_size_factor * (double)coder->voxel_size[1];
  a.dec.x0 = (double)coder->pc_range[0];
  a.dec.y0 = (double)coder->pc_range[1];
  a.dec.norm_bbox = coder->norm_bbox ? 1 : 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GD_DISPATCH_LOSS(launch_center, a, st)
}

}  // extern "C"
