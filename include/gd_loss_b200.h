/* gd_loss_b200.h -- C ABI of the B200 (sm_100a) Gaussian-distance loss library.
 *
 * Drop-in boundary for the ONE hot path of zhanggefan/mmdet3d-gaussian:
 *   mmdet3d_gaussian/models/losses/gaussian_distance_loss.py  (cited as ref:LINE)
 * The reference is pure PyTorch; its "FFI" is the Python call
 *   GDLoss.forward(pred, target, weight, avg_factor, reduction_override)  ref:280-310
 * followed by autograd's backward.  Each entry point below names the piece of
 * that call it replaces.  The Python binding a maintainer adds is a ctypes stub
 * (INTEGRATION.md); the shipped one is mmdet3d_gaussian_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary
 *   - every pointer is DEVICE memory unless the name ends in _host
 *   - the caller allocates all outputs and the workspace; nothing here
 *     allocates device memory, synchronises the device or throws (the *_host
 *     entry point is the documented exception: it owns a staging pipeline)
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*)
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     GD_ERR_* argument error
 *   - box rows are 7 fp32 (x, y, z, w, h, l, r)                       ref:8
 */
#ifndef GD_LOSS_B200_H_
#define GD_LOSS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
/* Only the entry points below are exported: the library is built with
 * -fvisibility=hidden -fno-gnu-unique so two builds of it (e.g. the fast-math and the
 * IEEE-math variant) can live in one process without sharing any state. */
#if defined(__GNUC__)
#define GD_API __attribute__((visibility("default")))
#else
#define GD_API
#endif

extern "C" {
#endif

#define GD_ABI_VERSION 3

/* loss_type: keys of GDLoss.BAG_GD_LOSS                                ref:253-259 */
enum {
  GD_LOSS_GWD3D = 0,        /* 'gwd3d'         ref:42-106  */
  GD_LOSS_KLD3D = 1,        /* 'kld3d'         ref:109-141 */
  GD_LOSS_JD3D = 2,         /* 'jd3d'          ref:189-198 */
  GD_LOSS_KLD3D_SYMMAX = 3, /* 'kld3d_symmax'  ref:201-211 */
  GD_LOSS_KLD3D_SYMMIN = 4, /* 'kld3d_symmin'  ref:214-224 */
  GD_LOSS_BD3D = 5,         /* 'bd3d'          ref:144-186 */
  GD_LOSS_KFIOU3D = 6       /* 'kfiou3d'       ref:227-248 */
};

/* fun: postprocess() non-linearity                                     ref:24-34 */
enum { GD_FUN_NONE = 0, GD_FUN_LOG1P = 1, GD_FUN_EXPM1 = 2, GD_FUN_NLOG = 3 };

/* weight layout                                                        ref:295-296 */
enum {
  GD_WEIGHT_NONE = 0,  /* weight is None                                           */
  GD_WEIGHT_ROW = 1,   /* [N]                                                      */
  GD_WEIGHT_ROW7 = 2   /* [N,7]: the kernel takes mean(-1) per row                 */
};

/* kernel variant */
enum {
  GD_VARIANT_AUTO = 0,   /* bulk pipeline when layout allows, else staged          */
  GD_VARIANT_STAGED = 1, /* LDG/STG staged through shared memory; any stride       */
  GD_VARIANT_BULK = 2,   /* persistent per-warp cp.async.bulk (TMA 1-D) + mbarrier
                            rings, 4 rows per lane                                 */
  GD_VARIANT_BULK_R2 = 3, /* same, 2 rows per lane / twice the warps (tuning aid)  */
  GD_VARIANT_BULK_PACKED = 4, /* GD_VARIANT_BULK with the FP32 math of two rows packed into
                                FFMA2/FMUL2/FADD2 (gwd3d / kld3d / bd3d with gradient; other
                                cases run GD_VARIANT_BULK).  Opt-in: not selected by AUTO.  */
  GD_VARIANT_BULK_ANY = 5    /* the bulk pipeline for ROW-STRIDED and/or 16-byte-unaligned
                                inputs (4-byte aligned): whole wide rows are copied, columns
                                0..6 picked in shared memory.  AUTO takes it whenever
                                GD_VARIANT_BULK cannot run and the rows fit shared memory.   */
};

/* flags of gd_loss_fwd_bwd */
enum {
  /* Rows whose (mean) weight is exactly 0 contribute exactly 0 loss and 0 gradient,
   * even if their distance is inf/nan.  With non-negative weights this makes the
   * all-zero-weight batch equal to the reference's early return (ref:290-292)
   * without the host sync that `torch.any(weight > 0)` costs there. */
  GD_FLAG_MASK_ZERO_WEIGHT = 1
};

enum {
  GD_ERR_BAD_ARG = -1,       /* null pointer / negative size / unknown enum        */
  GD_ERR_WORKSPACE = -2,     /* workspace too small                                */
  GD_ERR_LAYOUT = -3         /* GD_VARIANT_BULK requested on unaligned / strided   */
};

/* The constructor arguments of GDLoss that reach the distance function
 * (ref:261-278): loss_type, center_offset, fun, tau, alpha and the one extra
 * kwarg each distance accepts -- `normalize` for gwd3d (ref:43), `sqrt` for the
 * others (ref:110,145,190,202,215,228) -- passed as `flag`. */
typedef struct gd_loss_config {
  int32_t loss_type;
  int32_t fun;
  int32_t flag;
  float tau;               /* the kernel applies the tau map iff tau >= 1.0   ref:36 */
  float alpha;
  float center_offset[3];  /* ref:9-12 */
} gd_loss_config;

GD_API int gd_abi_version(void);

/* Bytes of device workspace gd_loss_* needs for `n` rows.  The workspace must
 * be zero-filled ONCE when allocated; the kernels leave it zeroed again, so it
 * can be reused by later calls on the same stream without clearing. */
GD_API size_t gd_loss_workspace_bytes(int64_t n);

/* Fused forward + backward: replaces GDLoss.forward (ref:298-310: preprocess x2,
 * distance, postprocess, weighted reduction, x loss_weight) AND the autograd
 * backward to `pred`, in one pass over HBM.
 *
 *   pred, target : [n,7] fp32, row strides in ELEMENTS (7 when contiguous;
 *                  the CenterGDHead call site passes row-strided views,
 *                  gd_centerpoint_head.py:413-423)
 *   weight       : per weight_mode; weight_row_stride in elements (1 or 7 when
 *                  contiguous); ignored for GD_WEIGHT_NONE
 *   scale        : every known-at-forward scalar folded together:
 *                  loss_weight * (1/n | 1/avg_factor | 1)            ref:310 + mmdet weight_reduce_loss
 *   loss_sum     : nullable, 1 fp32 := scale * sum_i w_i * loss_i   (written, not accumulated)
 *   row_loss     : nullable, [n] fp32 := scale * w_i * loss_i       (reduction='none')
 *   grad_pred    : nullable, [n,7] fp32 contiguous := scale * w_i * d loss_i / d pred_i
 *                  (nullable: forward only, e.g. under torch.no_grad)
 * The sum is deterministic: per-CTA partials in fp64, combined in fixed order
 * by the last CTA to finish. */
GD_API int gd_loss_fwd_bwd(const gd_loss_config* cfg,
                    const float* pred, int64_t pred_row_stride,
                    const float* target, int64_t target_row_stride,
                    const float* weight, int32_t weight_mode, int64_t weight_row_stride,
                    int64_t n, float scale,
                    float* loss_sum, float* row_loss, float* grad_pred,
                    void* workspace, size_t workspace_bytes,
                    int32_t variant, int32_t flags, void* stream);

/* Cross-GPU sum of the scalar loss INSIDE the fused launch (the one
 * collective of the row-sharded path), over peer memory instead of a separate NCCL launch.
 * Every rank owns an exchange buffer of gd_peer_sum_buffer_bytes() bytes that all ranks of
 * the box can address (a symmetric / P2P-mapped allocation, e.g. torch symmetric memory),
 * zero-filled once before first use.  The last CTA of the launch stores its scaled partial
 * into slot [rank] of EVERY peer's buffer over NVLink (one 8-byte store + one release flag per
 * peer), waits until all `world` flags of its own buffer carry the call's sequence number, and
 * adds the `world` partials in rank order -- so loss_sum is the global sum, bit-identical on
 * all ranks, deterministic.  All ranks must issue the same sequence of launches with the same
 * buffers, one stream per buffer (as for any collective); the launch completes on a rank only
 * once every rank has contributed.  Gradients stay local (DDP semantics). */
#define GD_MAX_PEERS 16
typedef struct gd_peer_sum {
  int32_t world;                       /* 0 or 1: no exchange */
  int32_t rank;
  void* peer_buf[GD_MAX_PEERS];        /* rank r's exchange buffer as addressable from THIS device */
} gd_peer_sum;
GD_API size_t gd_peer_sum_buffer_bytes(void);

/* The same launch with every argument in one struct, plus what the positional entry point
 * cannot express (ABI >= 2).  Zero-initialise the struct, then fill what applies. */
typedef struct gd_loss_io {
  const float* pred;   int64_t pred_row_stride;     /* as gd_loss_fwd_bwd */
  const float* target; int64_t target_row_stride;
  const float* weight; int32_t weight_mode; int64_t weight_row_stride;
  int64_t n;
  float scale;
  /* nullable, 1 fp32 on the DEVICE: the effective scale is scale / *scale_div.  Lets
   * avg_factor stay a device scalar (gd_centerpoint_head.py:407 computes it with .item(),
   * gd_anchor3d_head.py:102-105 with nonzero + len: both host syncs) -- pass
   * scale = loss_weight, scale_div = avg_factor. */
  const float* scale_div;
  float* loss_sum; float* row_loss; float* grad_pred;
  /* nullable, 1 fp32 := 1 if ANY element of `weight` is > 0 else 0 -- the reference's
   * early-return condition `torch.any(weight > 0)` (ref:290) evaluated inside the same
   * launch (needs loss_sum). */
  float* status;
  /* Early-return branch of GDLoss.forward WITHOUT a host sync (ref:290-292; needs loss_sum and
   * a weight).  Non-zero: when no weight element is > 0, the last CTA of the launch replaces
   * the outputs by what the reference returns on that branch,
   *   loss_sum := sum_{i,c} pred[i,c] * weight(i,c)          (pred * weight).sum(), ref:292
   *   grad_pred[i,c] := weight(i,c)                          its autograd gradient
   * with weight(i,c) = weight[i * er_weight_row_stride + c * er_weight_col_stride]
   * ([n,7] weights: (row stride, 1); a [7] weight against [7,7] rows broadcasts over columns:
   * (0, 1); [1]: (0, 0)) and no loss_weight / avg_factor (the reference applies none there).
   * The branch is rare (an empty batch in practice), so one CTA doing the rewrite is fine. */
  int32_t early_return;
  int64_t er_weight_row_stride, er_weight_col_stride;
  void* workspace; size_t workspace_bytes;
  int32_t variant; int32_t flags;
  /* nullable HOST pointer: sum loss_sum over the ranks of the box inside the launch */
  const gd_peer_sum* peer_sum;
  /* ABI >= 3.  Nullable, one int32 of PINNED host memory (cudaHostAlloc: device-addressable under
   * UVA) that the CALLER sets to 0 before the launch: the host-visible form of
   * `torch.any(weight > 0)` (ref:290) for the weight shapes where the reference's early return
   * RAISES, reported from inside the fused launch instead of by a probe launch in front of it.
   * The kernel stores 1 as soon as the first tile of its first warp holds a positive weight
   * (microseconds after it starts: the common case), otherwise its last CTA stores 1 (some
   * element > 0) or 2 (none).  Wait with gd_host_flag_wait.  Needs loss_sum and a weight. */
  int32_t* any_positive_host;
} gd_loss_io;

GD_API int gd_loss_launch(const gd_loss_config* cfg, const gd_loss_io* io, void* stream);

/* Autograd fold: grad[i,:] *= *grad_output (0-dim upstream gradient; replaces the
 * first step of the reference's autograd backward).  Reads the scalar on the
 * device -- no host sync. */
GD_API int gd_scale_grad(float* grad, int64_t n, const float* grad_output_scalar, void* stream);

/* Same fold for a gradient buffer of any shape: buf[0..count) *= *scalar (head front
 * ends, whose gradient rows are wider than 7). */
GD_API int gd_scale_buffer(float* buf, int64_t count, const float* scalar, void* stream);

/* Autograd fold for reduction='none': grad[i,:] *= grad_output[i]. */
GD_API int gd_scale_grad_rows(float* grad, int64_t n, const float* grad_output_rows,
                       int64_t grad_output_stride, void* stream);

/* Early-return branch of GDLoss.forward (ref:290-292): when a weight is given
 * and no element is > 0 the reference returns (pred*weight).sum().  This writes
 * flag[0] = 1 if any weight element is > 0 else 0 (count elements, any layout
 * flattened by the caller) so the shim can take that branch. */
GD_API int gd_any_positive(const float* weight, int64_t count, int32_t* flag, void* stream);

/* Spins (no CUDA call on the fast path) until *flag != 0, `flag` being the
 * gd_loss_io.any_positive_host word of a launch queued on `stream`.  Returns 0 once the flag is
 * set; if the stream drains or fails with the flag still 0, the CUDA error (or GD_ERR_BAD_ARG). */
GD_API int gd_host_flag_wait(const int32_t* flag, void* stream);

/* out[0] := max(#{i : 0 <= labels[i] < num_classes}, 1) as fp32 -- the avg_factor of the
 * anchor head's labels mode when the caller passes none (reduction='mean' over the positives,
 * gd_anchor3d_head.py:102-105 without nonzero / len), for gd_loss_io.scale_div.
 * Workspace as gd_loss_workspace_bytes. */
GD_API int gd_count_positive_labels(const int64_t* labels, int64_t total, int64_t num_classes,
                                    float* out, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* Pairwise matrix (new surface):
 *   out[i, j] = postprocess(distance(boxes1[i], boxes2[j]))   i<n, j<m
 * equal to the element-wise path on the broadcast-expanded pairs.  boxes
 * contiguous [n,7] / [m,7]; out row stride in elements (>= m). */
GD_API int gd_pairwise(const gd_loss_config* cfg,
                const float* boxes1, int64_t n,
                const float* boxes2, int64_t m,
                float* out, int64_t out_row_stride, void* stream);

/* Pairwise with the consumer fused (assigner use): per row i the argmin/min
 * over j, per column j the argmin over i is left to the caller via the matrix.
 *   row_min [n] fp32, row_argmin [n] int32; the matrix itself is not written. */
GD_API int gd_pairwise_row_argmin(const gd_loss_config* cfg,
                           const float* boxes1, int64_t n,
                           const float* boxes2, int64_t m,
                           float* row_min, int32_t* row_argmin, void* stream);

/* flags of gd_pairwise_assign */
enum {
  GD_PAIR_SIMILARITY = 1,  /* the optional matrix holds 1 - value ("larger is closer", the
                              iou_calculator convention of mmdet assigners); the minima
                              are always minima of the distance value                   */
  GD_PAIR_CPL1 = 2         /* measurement / test aid: the column-lane kernel with one column per
                              lane (the round-1 mapping) also where the row-lane kernel or two
                              columns per lane would run; same per-pair arithmetic (explicitly
                              rounded operations), bit-identical results                      */
  ,
  GD_PAIR_INDEX64 = 4      /* row_argmin / col_argmin point to int64 arrays (torch's index type:
                              saves the caller two conversion launches)                       */
};

/* Bytes of device workspace gd_pairwise_assign needs for m columns.  Zero-filled ONCE
 * when allocated; the kernel leaves it zeroed again. */
GD_API size_t gd_pairwise_workspace_bytes(int64_t m);

/* Pairwise distances with BOTH assigner reductions fused (SURVEY.md section 8 row f2;
 * what a MaxIoUAssigner-style consumer takes from an N x M cost matrix, cf. the
 * reference's IoU-based core/bbox/assigners/sim_ota_3d_assigner.py:91-123):
 *   row_min[i], row_argmin[i] = min / argmin over j of D[i,j]     (per anchor: closest GT)
 *   col_min[j], col_argmin[j] = min / argmin over i of D[i,j]     (per GT: closest anchor)
 * NaN is the minimum (torch.min semantics); ties go to the lowest index.  `out` is
 * optional: when null the matrix is never written.  The reductions and the matrix come
 * from the same per-pair instruction sequence, so they are bit-consistent. */
GD_API int gd_pairwise_assign(const gd_loss_config* cfg,
                       const float* boxes1, int64_t n,
                       const float* boxes2, int64_t m,
                       float* row_min, int32_t* row_argmin,
                       float* col_min, int32_t* col_argmin,
                       float* out, int64_t out_row_stride, int32_t flags,
                       void* workspace, size_t workspace_bytes, void* stream);

/* MaxIoUAssigner-style labels from the fused minima, with similarity = 1 - distance
 * (for tau >= 1 that is tau / (tau + f(d)) in (0, 1], an IoU-like score).  Restates
 * upstream mmdet MaxIoUAssigner.assign_wrt_overlaps with gt_max_assign_all=False
 * (mmdet is not in the reference checkout: parity unpinned, contract in DESIGN.md):
 *   assigned = -1;  neg_lo <= sim < neg_hi -> 0;  sim >= pos_thr -> row_argmin + 1;
 *   match_low_quality: for each GT j in order with (1 - col_min[j]) >= min_pos:
 *                      assigned[col_argmin[j]] = j + 1   (the last GT wins)
 * assigned_gt_inds [n] int64, max_overlaps [n] fp32 (nullable) = 1 - row_min. */
GD_API int gd_assign_from_minima(const float* row_min, const int32_t* row_argmin, int64_t n,
                          const float* col_min, const int32_t* col_argmin, int64_t m,
                          float pos_thr, float neg_lo, float neg_hi, float min_pos,
                          int32_t match_low_quality, int64_t* assigned_gt_inds,
                          float* max_overlaps, void* stream);

/* SimOTA-style consumer without the N x M matrix (second branch;
 * mmdet3d_gaussian/core/bbox/assigners/sim_ota_3d_assigner.py:184-211 dynamic_k_matching with the
 * Gaussian similarity 1 - D in the role of the IoU and a cost monotone in D).
 *
 * gd_pairwise_col_topk: for every column j the k (<= 16) smallest D[i,j] over the rows, ascending
 * (NaN first, ties -> lowest row):  topk_val [k,m] fp32, topk_row [k,m] int32 (-1 / +inf where
 * n < k) -- what ref:187-188 `torch.topk(pairwise_ious, candidate_topk, dim=0)` and the per-GT
 * `torch.topk(cost[:, gt], k=dynamic_k, largest=False)` of ref:192-193 read -- plus the row
 * (min, argmin) that the conflict rule ref:198-203 needs.  `out` (nullable, row stride in
 * elements) additionally receives the matrix.  Row minima, matrix and column lists come from
 * separate launches (row-lane kernel; matrix kernel; threshold sample + filter + per-column
 * selection) that evaluate a pair with the same explicitly rounded operations, so they are
 * reductions of one and the same matrix bit for bit (tests).  Exact for any input: a column
 * whose candidate buffer overflows (masses of equal values) is recomputed by brute force.
 * Workspace: scratch, no zeroing needed, gd_pairwise_topk_workspace_bytes(n, m) bytes. */
GD_API size_t gd_pairwise_topk_workspace_bytes(int64_t n, int64_t m);
GD_API int gd_pairwise_col_topk(const gd_loss_config* cfg,
                                const float* boxes1, int64_t n,
                                const float* boxes2, int64_t m, int32_t k,
                                float* row_min, int32_t* row_argmin,
                                float* topk_val, int32_t* topk_row,
                                float* out, int64_t out_row_stride,
                                void* workspace, size_t workspace_bytes, void* stream);

/* dynamic_k_matching (ref:184-211) from those lists:
 *   dynamic_ks[j] = clamp(int(sum_i (1 - topk_val[i,j])), min=1)                    ref:188-190
 *   rows topk_row[0..dynamic_ks[j]) , j  are matched                                ref:191-194
 *   a row matched by several GTs keeps row_argmin (its lowest cost over ALL GTs)    ref:198-203
 * assigned_gt_inds [n] int64: 0 = background, else GT index + 1 (ref:64-66,112);
 * matched_sim [n] nullable: 1 - D of the kept match, `unmatched_sim` elsewhere (ref:116-118 uses
 * -INF); dynamic_ks [m] nullable; scratch: 3 * n int32, any content. */
GD_API int gd_simota_from_topk(const float* topk_val, const int32_t* topk_row, int64_t m, int32_t k,
                               const float* row_min, const int32_t* row_argmin, int64_t n,
                               int64_t* assigned_gt_inds, float* matched_sim, int32_t* dynamic_ks,
                               float unmatched_sim, int32_t* scratch, void* stream);

/* ---------------------------------------------------------------------------
 * Head front ends: positive-row gather + box decode +
 * GD loss + gradient w.r.t. the RAW network outputs in one launch.
 * ------------------------------------------------------------------------- */

/* where the gradient rows go (anchor front end) */
enum {
  GD_GRAD_NONE = 0,     /* forward only                                                */
  GD_GRAD_COMPACT = 1,  /* index mode: [num_pos,7], row k belongs to pos_inds[k]       */
  GD_GRAD_SCATTER = 2,  /* index mode: rows pos_inds[k] of a caller-ZEROED [T,7]        */
  GD_GRAD_DENSE = 3     /* mask mode: every row of [T,7] is written (0 for negatives)  */
};

/* Replaces GDAnchor3DHead.loss_single's GD branch
 * (mmdet3d_gaussian/models/dense_heads/gd_anchor3d_head.py:102-112 positive-row
 * selection and gathers, :128-131 weights, :133-136 bbox_coder.decode x2 -- upstream
 * mmdet3d DeltaXYZWLHRBBoxCoder.decode --, :137-141 GDLoss) and its backward.
 *
 *   anchors       : [anchor_rows,7]; row i of the batch uses anchors[i % anchor_rows]
 *                   (anchor_list.repeat(mini_batch,1), :110-111)
 *   deltas_pred   : [T,7] bbox_pred after permute/reshape (:97-98), row stride in elements
 *   deltas_target : [T,7] bbox_targets (:99)
 *   bbox_weights  : nullable [T,7] (:100); with decode_weight_host[7] (HOST memory, the
 *                   train_cfg 'decode_weight' broadcast to 7) the row weight is
 *                   mean_c(bbox_weights[i,c] * decode_weight[c]).  Null => weight None.
 *   row selection : EITHER pos_inds[num_pos] (int64, the reference's nonzero result)
 *                   OR labels[T] (int64) + num_classes: positive <=> 0 <= label < num_classes
 *                   (:102-104) decided in-kernel -- no nonzero, no host sync.
 *   scale         : loss_weight / avg_factor (as gd_loss_fwd_bwd)
 *   scale_div     : nullable DEVICE scalar; effective scale = scale / *scale_div (as
 *                   gd_loss_io.scale_div: avg_factor without the host sync of :102-105)
 *   loss_sum      : nullable, 1 fp32 := scale * sum_pos w_i loss_i
 *   grad_deltas   : per grad_mode; = scale * w_i * d loss_i / d deltas_pred[i]
 * No positives => loss 0 and zero gradient (the reference's `pos_bbox_pred.sum()`
 * branch, :160-161). */
GD_API int gd_anchor_decoded_loss_fwd_bwd(const gd_loss_config* cfg,
                                   const float* anchors, int64_t anchor_rows,
                                   const float* deltas_pred, int64_t deltas_pred_row_stride,
                                   const float* deltas_target, int64_t deltas_target_row_stride,
                                   const float* bbox_weights, int64_t bbox_weights_row_stride,
                                   const float* decode_weight_host,
                                   const int64_t* pos_inds, int64_t num_pos,
                                   const int64_t* labels, int64_t num_classes,
                                   int64_t total_rows, float scale, const float* scale_div,
                                   float* loss_sum, float* grad_deltas, int32_t grad_mode,
                                   void* workspace, size_t workspace_bytes,
                                   int32_t flags, void* stream);

/* CenterPointBBoxCoderRev.__init__ constants
 * (mmdet3d_gaussian/core/bbox/coders/centerpoint_bbox_coders.py:9-20) */
typedef struct gd_center_coder {
  double pc_range[2];       /* pc_range[0], pc_range[1]      */
  double voxel_size[2];     /* voxel_size[0], voxel_size[1]  */
  int32_t out_size_factor;
  int32_t norm_bbox;        /* dims = exp(pred)              */
} gd_center_coder;

/* Replaces CenterGDHead.loss's GD branch
 * (mmdet3d_gaussian/models/dense_heads/gd_centerpoint_head.py:421-423
 * CenterPointBBoxYawCoder.decode(locs, preds, correct_yaw=False)[..., :7] --
 * core/bbox/coders/centerpoint_bbox_yaw_coders.py:18-56 --, :433-434 GDLoss) and
 * its backward.
 *   preds  : [n, >=7] gathered head outputs (reg2, height, dim3, yaw, ...), row stride
 *   locs   : [n, 2] int64 (x_ind, y_ind) = pos_ind[..., 1:], row stride in elements
 *   target : [n, >=7] real-world boxes (anno_boxes / encode()[..., :7]), row stride
 *   weight : as gd_loss_fwd_bwd (the reference passes none)
 *   grad_preds : nullable [n, grad_cols >= 7] with row stride; columns >= 7 are zeroed */
GD_API int gd_center_decoded_loss_fwd_bwd(const gd_loss_config* cfg, const gd_center_coder* coder,
                                   const float* preds, int64_t preds_row_stride,
                                   const int64_t* locs, int64_t locs_row_stride,
                                   const float* target, int64_t target_row_stride,
                                   const float* weight, int32_t weight_mode,
                                   int64_t weight_row_stride,
                                   int64_t n, float scale, const float* scale_div,
                                   float* loss_sum, float* grad_preds,
                                   int64_t grad_row_stride, int32_t grad_cols,
                                   void* workspace, size_t workspace_bytes,
                                   int32_t flags, void* stream);

/* How gd_loss_fwd_bwd_host cuts n rows into chunks (pure host arithmetic, callable without
 * a GPU).  Chunks are `chunk_rows` rows (rounded up to a multiple of 256; <= 0 selects the
 * default 2^20), except that the LAST chunk of a multi-chunk input is tapered -- halved
 * repeatedly down to chunk_rows / 8 -- so that the device->host copy of the final piece,
 * which nothing can overlap, is short (n <= chunk_rows stays a single chunk).  Every chunk start is a multiple of 256 rows (16-byte aligned rows).
 * Writes up to `capacity` (start, rows) pairs and returns the number of chunks, or a negative
 * GD_ERR_* (more chunks than capacity / than the pipeline supports). */
GD_API int64_t gd_host_chunk_plan(int64_t n, int64_t chunk_rows, int64_t* starts, int64_t* rows,
                                  int64_t capacity);

/* End-to-end entry with HOST buffers (bench `e2e`): chunks rows, overlaps
 * H2D of chunk k+1 / kernel of chunk k / D2H of chunk k-1 on internal streams,
 * returns when loss_host and grad_host are complete.  Host buffers may be
 * pageable or pinned (pinned is what overlaps).  grad_host nullable. */
GD_API int gd_loss_fwd_bwd_host(const gd_loss_config* cfg,
                         const float* pred_host, const float* target_host,
                         const float* weight_host, int32_t weight_mode,
                         int64_t n, float scale,
                         float* loss_host, float* grad_host,
                         int32_t device, int64_t chunk_rows);

/* Number of kernel launches issued by this library in this process (bench
 * `gpu_launches`). */
GD_API int64_t gd_launch_count(void);

/* CTAs of the persistent fused loss kernel (at most one per SM).  Built-in policy (0): the board
 * runs this kernel at its power cap and on most chips 128 of the 148 SMs deliver 5-6 % more
 * bandwidth than all of them, on some they deliver less (profiles/r03_grid.md) -- so the first
 * launch of a process with >= 2^22 rows on a contiguous [N] / no-weight layout TIMES the launch it
 * was asked for on a ladder of grids (148, 136, 132, 128, 124, 120, 116 CTAs) after ~100 ms of
 * launches that bring the board into its power-capped steady state (~0.25 s once per device; this
 * one call synchronises with the host) and keeps the fastest.  Never calibrated, one CTA per SM: [N,7] weights, row-strided /
 * unaligned inputs, ranks of a multi-process job (WORLD_SIZE > 1), launches captured into a CUDA
 * graph.  A positive value pins the grid for every later launch of the process and switches the
 * calibration off (measurements: tools/ab_grid.py; latency-critical callers). */
GD_API int gd_set_loss_grid(int32_t ctas);

GD_API const char* gd_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif  /* GD_LOSS_B200_H_ */
