"""Assigner-side consumers of the pairwise Gaussian distance.

The reference's assigners consume an N x M IoU matrix
(``core/bbox/assigners/sim_ota_3d_assigner.py:91-93``; the shipped configs use
upstream ``MaxIoUAssigner`` with ``BboxOverlapsNearest3D``).  Here the pairwise
Gaussian distance plays that role with the consumer fused into the producing
kernel:

* ``GDSimilarity3D`` -- an ``iou_calculator``-compatible callable
  (``__call__(bboxes1, bboxes2) -> [N, M]``) returning ``1 - D``; with ``tau >= 1``
  that is ``tau / (tau + f(d))`` in ``(0, 1]``, an IoU-like score, so an unmodified
  ``MaxIoUAssigner`` can run on it.
* ``GDMaxSimAssigner`` -- ``MaxIoUAssigner.assign_wrt_overlaps`` semantics
  (``gt_max_assign_all=False``) computed from the fused row / column minima: the
  N x M matrix is never written (the kernel is then FP32-pipe bound instead of
  write bound).
"""
import torch
from torch import nn

from . import _lib, ops
from .losses.gaussian_distance_loss import _FLAG


def _make_cfg(loss_type, fun, tau, alpha, center_offset, kwargs):
    assert loss_type in _lib.LOSS_TYPES
    if loss_type != 'kfiou3d':
        assert fun in ['log1p', 'none']
    else:
        assert fun in ['nlog', 'expm1', 'none']
    name, default = _FLAG[loss_type]
    unknown = set(kwargs) - {name}
    if unknown:
        raise TypeError(f'unexpected keyword argument {sorted(unknown)[0]!r}')
    return _lib.make_config(loss_type, fun, kwargs.get(name, default), tau, alpha, center_offset)


class GDSimilarity3D(nn.Module):
    """``iou_calculator`` drop-in: ``sim[i, j] = 1 - post(distance(b1[i], b2[j]))``."""

    def __init__(self, loss_type='gwd3d', center_offset=(0, 0, 0.5), fun='log1p', tau=1.0,
                 alpha=1.0, **kwargs):
        super().__init__()
        self.cfg = _make_cfg(loss_type, fun, tau, alpha, center_offset, kwargs)

    @torch.no_grad()
    def forward(self, bboxes1, bboxes2, mode='iou', is_aligned=False):
        if is_aligned:
            raise NotImplementedError('aligned mode: use GDLoss(reduction="none")')
        return ops.pairwise_assign(bboxes1[..., :7], bboxes2[..., :7], self.cfg,
                                   want_matrix=True, similarity=True)[4]


class GDMaxSimAssigner(nn.Module):
    """``MaxIoUAssigner`` on the Gaussian similarity without materialising the matrix.

    ``assign(bboxes [N,>=7], gt_bboxes [M,>=7]) -> dict(assigned_gt_inds [N] int64
    (-1 ignore, 0 negative, k+1 = GT k), max_overlaps [N], gt_max_overlaps [M],
    gt_argmax_overlaps [M])``.  Thresholds as in mmdet: ``pos_iou_thr``,
    ``neg_iou_thr`` (float or (lo, hi)), ``min_pos_iou``, ``match_low_quality``;
    ``gt_max_assign_all`` must be False (ties need the matrix)."""

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, match_low_quality=True,
                 gt_max_assign_all=False, loss_type='gwd3d', center_offset=(0, 0, 0.5),
                 fun='log1p', tau=1.0, alpha=1.0, **kwargs):
        super().__init__()
        if gt_max_assign_all:
            raise NotImplementedError('gt_max_assign_all=True needs the full matrix; use '
                                      'GDSimilarity3D with a stock MaxIoUAssigner')
        self.pos_iou_thr = float(pos_iou_thr)
        if isinstance(neg_iou_thr, (tuple, list)):
            self.neg_lo, self.neg_hi = float(neg_iou_thr[0]), float(neg_iou_thr[1])
        else:
            self.neg_lo, self.neg_hi = 0.0, float(neg_iou_thr)
        self.min_pos_iou = float(min_pos_iou)
        self.match_low_quality = bool(match_low_quality)
        self.cfg = _make_cfg(loss_type, fun, tau, alpha, center_offset, kwargs)

    @torch.no_grad()
    def assign(self, bboxes, gt_bboxes):
        row_min, row_arg, col_min, col_arg, _ = ops.pairwise_assign(
            bboxes[..., :7], gt_bboxes[..., :7], self.cfg, index64=False)   # int32: fed straight back
        assigned, max_ov = ops.assign_from_minima(
            row_min, row_arg, col_min, col_arg, self.pos_iou_thr, self.neg_lo, self.neg_hi,
            self.min_pos_iou, self.match_low_quality)
        if gt_bboxes.shape[0] == 0:
            assigned.zero_()                   # mmdet: no GT -> everything is background
            max_ov.zero_()
        return dict(assigned_gt_inds=assigned, max_overlaps=max_ov,
                    gt_max_overlaps=1.0 - col_min, gt_argmax_overlaps=col_arg.long())

    forward = assign


class GDSimOTAAssigner(nn.Module):
    """SimOTA's dynamic-k matching on the Gaussian similarity without the N x M matrix
    (the reference's ``SimOTABEVAssigner.dynamic_k_matching``,
    ``core/bbox/assigners/sim_ota_3d_assigner.py:184-211``, fed with ``pairwise_ious = 1 - D`` and
    a cost monotone in ``D``; the reference builds both from a materialised BEV-IoU matrix,
    :91-107).

    Per GT the ``candidate_topk`` most similar boxes give ``dynamic_k = clamp(int(sum of their
    similarities), min=1)`` (:187-190); the ``dynamic_k`` cheapest boxes of that GT are matched
    (:191-194); a box matched by several GTs keeps the GT of its own lowest cost (:198-203).
    Everything this reads is the column top-k and the row minima of the distance matrix, which
    the fused kernel reduces on the fly (``gd_pairwise_col_topk`` + ``gd_simota_from_topk``).
    ``tau >= 1`` keeps ``D`` in ``[0, 1)`` so that ``1 - D`` is an IoU-like score.  Ties between
    equal distances go to the lowest box index (``torch.topk`` leaves them unspecified).

    ``assign(bboxes [N,>=7], gt_bboxes [M,>=7]) -> dict(assigned_gt_inds [N] int64 (0 background,
    k+1 = GT k), max_overlaps [N] (similarity of the kept match, -INF elsewhere as :116-118),
    dynamic_ks [M], topk_overlaps [k,M], topk_inds [k,M])``.  The class / centre-prior terms of
    the reference's cost (:94-107) are the caller's: this assigner is the GD-only core."""

    INF = 100000000                                   # ref:12

    def __init__(self, candidate_topk=10, loss_type='gwd3d', center_offset=(0, 0, 0.5), fun='log1p',
                 tau=1.0, alpha=1.0, **kwargs):
        super().__init__()
        if not 1 <= int(candidate_topk) <= 16:
            raise ValueError('candidate_topk must be in [1, 16]')
        if float(tau) < 1.0:
            raise ValueError('GDSimOTAAssigner needs tau >= 1 (similarity 1 - D in (0, 1])')
        self.candidate_topk = int(candidate_topk)
        self.cfg = _make_cfg(loss_type, fun, tau, alpha, center_offset, kwargs)

    @torch.no_grad()
    def assign(self, bboxes, gt_bboxes, want_matrix=False):
        n, m = bboxes.shape[0], gt_bboxes.shape[0]
        dev = bboxes.device
        if n == 0 or m == 0:                          # ref:67-79
            return dict(assigned_gt_inds=torch.zeros(n, dtype=torch.int64, device=dev),
                        max_overlaps=torch.zeros(n, device=dev),
                        dynamic_ks=torch.zeros(m, dtype=torch.int64, device=dev),
                        topk_overlaps=torch.zeros(0, m, device=dev),
                        topk_inds=torch.zeros(0, m, dtype=torch.int64, device=dev))
        k = min(self.candidate_topk, n)               # ref:187
        row_min, row_arg, tv, tr, mat = ops.pairwise_col_topk(
            bboxes[..., :7], gt_bboxes[..., :7], self.cfg, k, want_matrix=want_matrix)
        assigned, sim, dks = ops.simota_from_topk(tv, tr, row_min, row_arg,
                                                  unmatched_sim=-float(self.INF))
        out = dict(assigned_gt_inds=assigned, max_overlaps=sim, dynamic_ks=dks,
                   topk_overlaps=1.0 - tv, topk_inds=tr)
        if want_matrix:
            out['distance_matrix'] = mat              # debugging / tests: the values the lists hold
        return out

    forward = assign
