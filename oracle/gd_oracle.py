"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Gaussian-distance (GD) loss.

A restatement, in plain eager PyTorch, of the algorithm in the reference file
``mmdet3d_gaussian/models/losses/gaussian_distance_loss.py`` (cited below as
``ref:LINE``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / CPU baseline -- the product path
(``mmdet3d_gaussian_b200``) never touches it and has no CPU fallback.

Parity status: **pinned**.  The reference ships no golden vectors or tests
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, imported unmodified in the build container through
``oracle/ref_loader.py``: ``oracle/make_golden.py`` writes the fixtures in
``tests/golden/`` from the *reference*, and ``tests/test_oracle.py`` checks this
restatement against those fixtures (and, when ``/root/reference`` is present,
against the live reference on fresh random inputs).

The structure deliberately keeps the reference's op granularity (batched 2x2
matrix products, materialised intermediates, autograd backward) so that timing
it on host cores is a fair stand-in ("port") for the reference's CPU torch path;
evaluated in float64 it is the numerical ground truth for the CUDA kernels.

Box row convention (ref:8, ``xyzwhlr``): ``(x, y, z, w, h, l, r)`` -- columns
3,4 are the BEV extents, column 5 the vertical extent, column 6 the yaw.
"""
import copy
from collections import namedtuple

import torch

EXTENT_MIN = 1e-7   # ref:13-14
EXTENT_MAX = 1e7    # ref:13-14
DET_FLOOR = 1e-7    # ref:158, ref:243, ref:245
KFIOU_SCALE = 4.656854249492381  # ref:247

LOSS_TYPES = ('gwd3d', 'kld3d', 'jd3d', 'kld3d_symmax', 'kld3d_symmin',
              'bd3d', 'kfiou3d')   # ref:253-259

Gauss = namedtuple('Gauss', 'mean rot half_bev half_z')


# --------------------------------------------------------------------------
# a1: box -> Gaussian pieces                                        ref:8-21
# --------------------------------------------------------------------------
def box_to_gaussian(boxes, center_offset=(0, 0, 0.5)):
    """ref:8-21.  Centre uses the UNCLAMPED extents (ref:12); the half-axes use
    extents clamped to [1e-7, 1e7] (ref:13-14)."""
    boxes = boxes.reshape(-1, 7)                                   # ref:11
    off = torch.as_tensor(center_offset).to(boxes)                 # ref:9-10
    mean = boxes[:, 0:3] + off.unsqueeze(0) * boxes[:, 3:6]        # ref:12
    bev = boxes[:, 3:5].clamp(min=EXTENT_MIN, max=EXTENT_MAX)      # ref:13
    vert = boxes[:, 5].clamp(min=EXTENT_MIN, max=EXTENT_MAX)       # ref:14
    yaw = boxes[:, 6]                                              # ref:15
    c, s = torch.cos(yaw), torch.sin(yaw)                          # ref:16-17
    rot = torch.stack((c, -s, s, c), dim=-1).reshape(-1, 2, 2)     # ref:18
    half_bev = torch.diag_embed(bev) * 0.5                         # ref:19
    half_z = vert * 0.5                                            # ref:20
    return Gauss(mean, rot, half_bev, half_z)


def _diag(m):
    return m.diagonal(dim1=-2, dim2=-1)


def _cov_bev(g, inverse=False):
    """R diag(a^2, b^2) R^T (ref:86-87,117,149-150) or its inverse (ref:114-116)."""
    d = _diag(g.half_bev)
    if inverse:
        d = d.reciprocal()
    core = torch.diag_embed(d).square()
    return g.rot.bmm(core).bmm(g.rot.transpose(1, 2))


# --------------------------------------------------------------------------
# a9: post map                                                     ref:24-39
# --------------------------------------------------------------------------
def post_map(d, fun='log1p', tau=1.0):
    if fun == 'log1p':
        d = torch.log1p(d)                                         # ref:26
    elif fun == 'expm1':
        d = torch.expm1(d)                                         # ref:28
    elif fun == 'nlog':
        d = -torch.log(1 - d + 1e-7)                               # ref:30
    elif fun != 'none':
        raise ValueError(f'Invalid non-linear function {fun}')     # ref:34
    if tau >= 1.0:                                                 # ref:36
        return 1 - tau / (tau + d)                                 # ref:37
    return d                                                       # ref:39


# --------------------------------------------------------------------------
# a2: Gaussian Wasserstein distance                               ref:42-106
# --------------------------------------------------------------------------
def gwd3d(p, t, fun='log1p', tau=1.0, alpha=1.0, normalize=True):
    centre_sq = (p.mean - t.mean).square().sum(-1)                 # ref:79
    tr_sum = _diag(p.half_bev).square().sum(-1) \
        + _diag(t.half_bev).square().sum(-1)                       # ref:81-84
    cross = _cov_bev(p).bmm(_cov_bev(t))                           # ref:86-88
    cross_tr = _diag(cross).sum(-1)                                # ref:90
    det_sqrt = _diag(p.half_bev).prod(-1) * _diag(t.half_bev).prod(-1)  # ref:91-92
    shape_sq = tr_sum - 2 * (cross_tr + 2 * det_sqrt).clamp(0).sqrt()   # ref:94-95
    shape_sq = shape_sq + (p.half_z - t.half_z).square()           # ref:97
    d = (centre_sq + alpha * alpha * shape_sq).clamp(0).sqrt()     # ref:99
    if normalize:                                                  # ref:101-104
        logsum = det_sqrt.log() + p.half_z.log() + t.half_z.log()
        d = d / (2 * (logsum / 6).exp())
    return post_map(d, fun, tau)                                   # ref:106


# --------------------------------------------------------------------------
# a3: Kullback-Leibler divergence (uses Sigma_p^-1)              ref:109-141
# --------------------------------------------------------------------------
def kld3d(p, t, fun='log1p', tau=1.0, alpha=1.0, sqrt=True):
    inv_p = _cov_bev(p, inverse=True)                              # ref:114-116
    inv_pz = p.half_z.reciprocal()                                 # ref:115
    cov_t = _cov_bev(t)                                            # ref:117
    dxy = (p.mean[:, :2] - t.mean[:, :2]).unsqueeze(-1)            # ref:119
    dz = p.mean[:, 2] - t.mean[:, 2]                               # ref:120
    maha = 0.5 * dxy.transpose(1, 2).bmm(inv_p).bmm(dxy).reshape(-1)   # ref:122-123
    maha = maha + 0.5 * dz.square() * inv_pz.square()              # ref:124
    shape = 0.5 * _diag(inv_p.bmm(cov_t)).sum(-1)                  # ref:126-127
    shape = shape + 0.5 * inv_pz.square() * t.half_z.square()      # ref:128
    logdet_p = _diag(p.half_bev).log().sum(-1) + p.half_z.log()    # ref:130-131
    logdet_t = _diag(t.half_bev).log().sum(-1) + t.half_z.log()    # ref:132-133
    shape = shape + (logdet_p - logdet_t) - 1.5                    # ref:134-136
    d = maha / (alpha * alpha) + shape                             # ref:137
    if sqrt:
        d = d.clamp(0).sqrt()                                      # ref:138-139
    return post_map(d, fun, tau)                                   # ref:141


# --------------------------------------------------------------------------
# a4: Bhattacharyya distance                                     ref:144-186
# --------------------------------------------------------------------------
def bd3d(p, t, fun='log1p', tau=1.0, alpha=1.0, sqrt=True):
    mid = 0.5 * (_cov_bev(p) + _cov_bev(t))                        # ref:149-152
    mid_z = 0.5 * (p.half_z.square() + t.half_z.square())          # ref:153
    det = mid[:, 0, 0] * mid[:, 1, 1] - mid[:, 1, 0] * mid[:, 0, 1]   # ref:155-157
    det = det.clamp(min=DET_FLOOR)                                 # ref:158
    adj = torch.stack((mid[:, 1, 1], -mid[:, 0, 1],
                       -mid[:, 1, 0], mid[:, 0, 0]), -1).reshape(-1, 2, 2)  # ref:160-164
    mid_inv = adj * det.reciprocal()[:, None, None]                # ref:165-166
    dxy = (p.mean[:, :2] - t.mean[:, :2]).unsqueeze(-1)            # ref:168
    dz = p.mean[:, 2] - t.mean[:, 2]                               # ref:169
    maha = 0.125 * dxy.transpose(1, 2).bmm(mid_inv).bmm(dxy).reshape(-1)  # ref:170-171
    maha = maha + 0.125 * dz.square() * mid_z.reciprocal()         # ref:172
    shape = 0.5 * (det.log() + mid_z.log())                        # ref:174
    shape = shape - 0.25 * (_diag(p.half_bev.square()).log().sum(-1)
                            + p.half_z.square().log())             # ref:175-177
    shape = shape - 0.25 * (_diag(t.half_bev.square()).log().sum(-1)
                            + t.half_z.square().log())             # ref:178-180
    d = maha / (alpha * alpha) + shape                             # ref:182
    if sqrt:
        d = d.clamp(0).sqrt()                                      # ref:183-184
    return post_map(d, fun, tau)                                   # ref:186


# --------------------------------------------------------------------------
# a5/a6: Jeffreys and symmetric KLD max/min                      ref:189-224
# --------------------------------------------------------------------------
def jd3d(p, t, fun='log1p', tau=1.0, alpha=1.0, sqrt=True):
    fwd = kld3d(p, t, fun='none', tau=0, alpha=alpha, sqrt=False)  # ref:191-192
    rev = kld3d(t, p, fun='none', tau=0, alpha=alpha, sqrt=False)  # ref:193-194
    d = (fwd + rev) * 0.5                                          # ref:195
    if sqrt:
        d = d.clamp(0).sqrt()                                      # ref:196-197
    return post_map(d, fun, tau)                                   # ref:198


def kld3d_symmax(p, t, fun='log1p', tau=1.0, alpha=1.0, sqrt=True):
    fwd = kld3d(p, t, fun='none', tau=0, alpha=alpha, sqrt=sqrt)   # ref:204-206
    rev = kld3d(t, p, fun='none', tau=0, alpha=alpha, sqrt=sqrt)   # ref:207-209
    return post_map(torch.max(fwd, rev), fun, tau)                 # ref:210-211


def kld3d_symmin(p, t, fun='log1p', tau=1.0, alpha=1.0, sqrt=True):
    fwd = kld3d(p, t, fun='none', tau=0, alpha=alpha, sqrt=sqrt)   # ref:217-219
    rev = kld3d(t, p, fun='none', tau=0, alpha=alpha, sqrt=sqrt)   # ref:220-222
    return post_map(torch.min(fwd, rev), fun, tau)                 # ref:223-224


# --------------------------------------------------------------------------
# a7: Kalman-filter IoU                                          ref:227-248
# --------------------------------------------------------------------------
def kfiou3d(p, t, fun='expm1', tau=0.0, alpha=1.0, sqrt=False):
    """Ignores centres, ``alpha``, ``sqrt`` and forces tau=0 (ref:247)."""
    tot = _cov_bev(p) + _cov_bev(t)                                # ref:232-234
    det_bev = tot[:, 0, 0] * tot[:, 1, 1] - tot[:, 1, 0] * tot[:, 0, 1]  # ref:235-236
    det_all = det_bev * (p.half_z.square() + t.half_z.square())    # ref:237-238
    vol_p = _diag(p.half_bev).prod(-1) * p.half_z                  # ref:240
    vol_t = _diag(t.half_bev).prod(-1) * t.half_z                  # ref:241
    inter = vol_p * vol_t / det_all.clamp(min=DET_FLOOR).sqrt()    # ref:243
    union = (vol_p + vol_t - inter).clamp(min=DET_FLOOR)           # ref:245
    return post_map(1 - KFIOU_SCALE * (inter / union), fun, 0.0)   # ref:246-247


DISTANCES = {'gwd3d': gwd3d, 'kld3d': kld3d, 'jd3d': jd3d,
             'kld3d_symmax': kld3d_symmax, 'kld3d_symmin': kld3d_symmin,
             'bd3d': bd3d, 'kfiou3d': kfiou3d}


# --------------------------------------------------------------------------
# a11: upstream mmdet ``weighted_loss`` / ``weight_reduce_loss`` (not in the
# reference checkout; contract written down in SURVEY.md section 8 row a11)
# --------------------------------------------------------------------------
def weight_reduce(loss, weight=None, reduction='mean', avg_factor=None):
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            return loss.mean()          # divides by N, not by sum(weight)
        if reduction == 'sum':
            return loss.sum()
        return loss
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction == 'none':
        return loss
    raise ValueError('avg_factor can not be used with reduction="sum"')


# --------------------------------------------------------------------------
# a8/a10: the module                                             ref:251-310
# --------------------------------------------------------------------------
class GDLossOracle(torch.nn.Module):
    """Restatement of ``GDLoss`` (ref:251-310): same constructor, same forward."""

    def __init__(self, loss_type, center_offset=(0, 0, 0.5), fun='log1p',
                 tau=1.0, alpha=1.0, reduction='mean', loss_weight=1.0,
                 **kwargs):
        super().__init__()
        assert reduction in ('none', 'sum', 'mean')                # ref:265
        assert loss_type in DISTANCES                              # ref:266
        if loss_type != 'kfiou3d':
            assert fun in ('log1p', 'none')                        # ref:267-268
        else:
            assert fun in ('nlog', 'expm1', 'none')                # ref:269-270
        self.loss_type = loss_type
        self.center_offset = center_offset
        self.fun, self.tau, self.alpha = fun, tau, alpha
        self.reduction, self.loss_weight = reduction, loss_weight
        self.kwargs = kwargs                                       # ref:278

    def forward(self, pred, target, weight=None, avg_factor=None,
                reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')  # ref:287
        reduction = reduction_override or self.reduction           # ref:288-289
        if weight is not None and reduction != 'none' \
                and not torch.any(weight > 0):                     # ref:290-291
            return (pred * weight).sum()                           # ref:292
        extra = copy.deepcopy(self.kwargs)                         # ref:293
        extra.update(kwargs)                                       # ref:294
        if weight is not None and weight.shape == pred.shape:      # ref:295
            weight = weight.mean(-1)                               # ref:296
        p = box_to_gaussian(pred, self.center_offset)              # ref:298
        t = box_to_gaussian(target, self.center_offset)            # ref:299
        rows = DISTANCES[self.loss_type](
            p, t, fun=self.fun, tau=self.tau, alpha=self.alpha, **extra)
        return weight_reduce(rows, weight, reduction, avg_factor) \
            * self.loss_weight                                     # ref:301-310


# --------------------------------------------------------------------------
# a12: pairwise N x M matrix (new surface; oracle = element-wise path on the
# broadcast-expanded pairs), chunked over rows.
# --------------------------------------------------------------------------
def pairwise_distance(boxes1, boxes2, loss_type, center_offset=(0, 0, 0.5),
                      fun='log1p', tau=1.0, alpha=1.0, chunk_rows=4096,
                      **kwargs):
    n, m = boxes1.shape[0], boxes2.shape[0]
    out = boxes1.new_empty((n, m))
    fn = DISTANCES[loss_type]
    for lo in range(0, n, chunk_rows):
        b1 = boxes1[lo:lo + chunk_rows]
        k = b1.shape[0]
        p = box_to_gaussian(b1.repeat_interleave(m, 0), center_offset)
        t = box_to_gaussian(boxes2.repeat(k, 1), center_offset)
        out[lo:lo + k] = fn(p, t, fun=fun, tau=tau, alpha=alpha,
                            **kwargs).reshape(k, m)
    return out


# --------------------------------------------------------------------------
# f1: the decoders the reference runs in front of the loss, and the two head
# call sites restated end to end.
# --------------------------------------------------------------------------
def decode_delta_xyzwlhr(anchors, deltas):
    """Upstream mmdet3d ``DeltaXYZWLHRBBoxCoder.decode`` (called at
    ``models/dense_heads/gd_anchor3d_head.py:133-136``).  mmdet3d is NOT in the
    reference checkout and is un-pinned by it, so this decoder is **parity
    unpinned**; the positional arithmetic below is the same in every mmdet3d
    release that ships the coder (the releases only rename columns 3/4)."""
    a0, a1, a2, a3, a4, a5, a6 = torch.split(anchors[..., :7], 1, dim=-1)
    d0, d1, d2, d3, d4, d5, d6 = torch.split(deltas[..., :7], 1, dim=-1)
    a2 = a2 + a5 / 2
    diagonal = torch.sqrt(a4 ** 2 + a3 ** 2)
    x = d0 * diagonal + a0
    y = d1 * diagonal + a1
    z = d2 * a5 + a2
    e4 = torch.exp(d4) * a4
    e3 = torch.exp(d3) * a3
    e5 = torch.exp(d5) * a5
    r = d6 + a6
    z = z - e5 / 2
    return torch.cat([x, y, z, e3, e4, e5, r], dim=-1)


def decode_centerpoint_yaw(locs, preds, pc_range, out_size_factor, voxel_size,
                           norm_bbox=True):
    """``CenterPointBBoxYawCoder.decode(locs, preds, correct_yaw=False)``
    (``core/bbox/coders/centerpoint_bbox_yaw_coders.py:18-56``; constants from
    ``centerpoint_bbox_coders.py:9-20``).  Pinned against the unmodified
    reference class by ``oracle/make_golden.py`` (``gd_decode_golden.npz``)."""
    x = (preds[..., 0] + locs[..., 0]) * out_size_factor * voxel_size[0] \
        + pc_range[0]                                              # coder:30-31
    y = (preds[..., 1] + locs[..., 1]) * out_size_factor * voxel_size[1] \
        + pc_range[1]                                              # coder:32-33
    z = preds[..., 2]                                              # coder:34
    dim = preds[..., 3:6]                                          # coder:35
    if norm_bbox:
        dim = dim.exp()                                            # coder:36-37
    yaw = preds[..., 6]                                            # coder:38
    others = preds[..., 9:]                                        # coder:51
    return torch.cat((x.unsqueeze(-1), y.unsqueeze(-1), z.unsqueeze(-1), dim,
                      yaw.unsqueeze(-1), others), dim=-1)          # coder:53-55


def anchor_head_gd_loss(module, anchors, bbox_pred, bbox_targets, bbox_weights,
                        pos_inds, decode_weight=None, avg_factor=None):
    """The GD branch of ``GDAnchor3DHead.loss_single``
    (``gd_anchor3d_head.py:107-141``): ``anchors`` is ``[A0,7]`` and repeats over
    the mini-batch (:110-111); ``bbox_pred/bbox_targets/bbox_weights`` are
    ``[T,7]``; returns the scalar the head adds into ``loss_bbox``."""
    total = bbox_pred.shape[0]
    pos_pred = bbox_pred[pos_inds]                                 # head:107
    pos_targets = bbox_targets[pos_inds]                           # head:108
    pos_weights = bbox_weights[pos_inds]                           # head:109
    reps = -(-total // anchors.shape[0])
    pos_anchors = anchors.repeat(reps, 1)[:total][pos_inds]        # head:110-112
    if len(pos_inds) == 0:
        return pos_pred.sum()                                      # head:160-161
    weight = None
    if decode_weight:                                              # head:128-131
        weight = pos_weights * bbox_weights.new_tensor(decode_weight)
    pred_dec = decode_delta_xyzwlhr(pos_anchors, pos_pred)         # head:133-134
    tgt_dec = decode_delta_xyzwlhr(pos_anchors, pos_targets)       # head:135-136
    return module(pred_dec, tgt_dec, weight, avg_factor=avg_factor)  # head:137-141


def center_head_gd_loss(module, pred, pos_ind, target_box, coder, avg_factor=None):
    """The GD branch of ``CenterGDHead.loss``
    (``gd_centerpoint_head.py:413-434``): ``pred`` ``[P,C]`` gathered head
    outputs, ``pos_ind`` ``[P,3]`` (batch, x, y), ``target_box`` ``[P,>=7]``;
    ``coder`` = dict(pc_range, out_size_factor, voxel_size, norm_bbox)."""
    target_gd = target_box[..., :7]                                # head:415
    pred_gd = decode_centerpoint_yaw(pos_ind[..., 1:], pred, **coder)[..., :7]  # head:421-422
    return module(pred_gd, target_gd, avg_factor=avg_factor)       # head:433-434


# --------------------------------------------------------------------------
# f2: MaxIoUAssigner semantics on a similarity matrix (upstream mmdet
# ``MaxIoUAssigner.assign_wrt_overlaps``, ``gt_max_assign_all=False``; mmdet is not in
# the reference checkout: parity unpinned, the loop below is the written contract).
# ``sim`` is [N, M] (anchors x GTs), i.e. the transpose of mmdet's ``overlaps``.
# --------------------------------------------------------------------------
def max_sim_assign(sim, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0,
                   match_low_quality=True):
    n, m = sim.shape
    assigned = sim.new_full((n,), -1, dtype=torch.long)
    if m == 0:
        return assigned.zero_(), sim.new_zeros((n,))
    max_overlaps, argmax_overlaps = sim.max(dim=1)
    gt_max_overlaps, gt_argmax_overlaps = sim.max(dim=0)
    lo, hi = (neg_iou_thr if isinstance(neg_iou_thr, (tuple, list))
              else (0.0, neg_iou_thr))
    assigned[(max_overlaps >= lo) & (max_overlaps < hi)] = 0
    pos = max_overlaps >= pos_iou_thr
    assigned[pos] = argmax_overlaps[pos] + 1
    if match_low_quality:
        for i in range(m):
            if gt_max_overlaps[i] >= min_pos_iou:
                assigned[gt_argmax_overlaps[i]] = i + 1
    return assigned, max_overlaps


# --------------------------------------------------------------------------
# helper used by tests and bench: loss + d loss / d pred in one call
# --------------------------------------------------------------------------
def loss_and_grad(module, pred, target, weight=None, avg_factor=None,
                  reduction_override=None, grad_output=None, **kwargs):
    """Run ``module`` (oracle or reference ``GDLoss``) forward + autograd
    backward; returns ``(loss.detach(), pred.grad)``."""
    pred = pred.detach().clone().requires_grad_(True)
    loss = module(pred, target, weight, avg_factor=avg_factor,
                  reduction_override=reduction_override, **kwargs)
    if loss.dim() == 0:
        loss.backward(None if grad_output is None else grad_output)
    else:
        loss.backward(torch.ones_like(loss) if grad_output is None
                      else grad_output)
    return loss.detach(), pred.grad


# --------------------------------------------------------------------------
# f2 (second branch): SimOTA's dynamic-k matching, restated from
# core/bbox/assigners/sim_ota_3d_assigner.py:184-211 ("sim:LINE").  Pinned against the
# reference's own method by tests/test_oracle.py (golden fixture written by
# oracle/make_simota_golden.py from the unmodified file).
# --------------------------------------------------------------------------
def simota_dynamic_k_matching(cost, pairwise_ious, candidate_topk=10):
    """``cost`` / ``pairwise_ious`` ``[num_priors, num_gt]``.  Returns
    ``(assigned_gt_inds [num_priors] int64 -- 0 background, k+1 = GT k (sim:112) --,
    matched_ious [num_priors] (0 where unmatched), dynamic_ks [num_gt])``.

    Ties: ``torch.topk`` leaves the order of equal entries unspecified; this restatement (and
    the CUDA kernels) take the LOWEST row index first, which is one of the reference's
    admissible outcomes."""
    n, num_gt = cost.shape
    matching = torch.zeros_like(cost)                                    # sim:185
    topk = min(candidate_topk, n)                                        # sim:187
    rows = torch.arange(n, device=cost.device)

    def lowest(values, k, largest):
        # k smallest / largest entries, ties -> lowest row (stable sort on the value)
        order = torch.sort(-values if largest else values, stable=True).indices
        return order[:k]
    dynamic_ks = torch.empty(num_gt, dtype=torch.int64, device=cost.device)
    for j in range(num_gt):
        top = lowest(pairwise_ious[:, j], topk, True)                    # sim:188
        dynamic_ks[j] = max(int(pairwise_ious[top, j].sum().int()), 1)   # sim:190
    for j in range(num_gt):                                              # sim:191-194
        pos = lowest(cost[:, j], int(dynamic_ks[j]), False)
        matching[pos, j] = 1.0
    multi = matching.sum(1) > 1                                          # sim:198
    if multi.sum() > 0:                                                  # sim:199-203
        cost_argmin = torch.min(cost[multi, :], dim=1).indices
        matching[multi, :] *= 0.0
        matching[rows[multi], cost_argmin] = 1.0
    fg = matching.sum(1) > 0.0                                           # sim:205
    assigned = torch.zeros(n, dtype=torch.int64, device=cost.device)
    assigned[fg] = matching[fg, :].argmax(1) + 1                         # sim:208, sim:112
    matched = (matching * pairwise_ious).sum(1)                          # sim:209-210
    return assigned, matched, dynamic_ks
