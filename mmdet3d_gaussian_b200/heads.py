"""Fused head front ends: the GD branches of the
reference's two loss call sites with the row gather, the box decode, the loss and
the gradient w.r.t. the RAW head outputs in one kernel launch.

* ``GDAnchorHeadLoss``  -- ``GDAnchor3DHead.loss_single``
  (``mmdet3d_gaussian/models/dense_heads/gd_anchor3d_head.py:102-141``)
* ``GDCenterHeadLoss``  -- ``CenterGDHead.loss``
  (``mmdet3d_gaussian/models/dense_heads/gd_centerpoint_head.py:413-434``)

Both wrap a ``GDLoss`` (config dict or instance) and reuse its
``loss_type / fun / tau / alpha / center_offset / loss_weight / normalize|sqrt``;
the arithmetic runs in ``csrc/gd_decoded.cu``.  INTEGRATION.md shows the
three-line patch of each head.
"""
from torch import nn

from . import _lib, ops
from .losses.gaussian_distance_loss import GDLoss


def _as_loss(cfg_or_module):
    if isinstance(cfg_or_module, GDLoss):
        return cfg_or_module
    cfg = dict(cfg_or_module)
    if cfg.pop('type', 'GDLoss') != 'GDLoss':
        raise ValueError('the fused head losses wrap a GDLoss config')
    return GDLoss(**cfg)


def _scale(loss, avg_factor, count):
    """mmdet ``weight_reduce_loss`` folded into one scalar for a
    HOST ``avg_factor`` (the rule the shim applies; kept in Python for the CPU tests)."""
    if loss.reduction == 'none':
        raise NotImplementedError('the fused head losses return the reduced scalar only')
    if avg_factor is not None:
        if loss.reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return loss.loss_weight / float(avg_factor)
    if loss.reduction == 'sum':
        return loss.loss_weight
    if count is None:
        raise ValueError('labels mode needs avg_factor (the number of positives is not known '
                         'on the host) or reduction="sum"')
    return loss.loss_weight / count if count > 0 else float('nan')


def _scale_args(loss, avg_factor, count):
    """(loss_weight, scale_mode, avg_factor) for the shim: mode 0 = final scale, 1 = divide by
    ``avg_factor`` (number or device tensor), 2 = divide by the device-side count."""
    if loss.reduction == 'none':
        raise NotImplementedError('the fused head losses return the reduced scalar only')
    if avg_factor is not None:
        if loss.reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return loss.loss_weight, 1, avg_factor
    if loss.reduction == 'sum':
        return loss.loss_weight, 0, None
    if count is None:
        return loss.loss_weight, 2, None
    return (loss.loss_weight / count if count > 0 else float('nan')), 0, None


class GDAnchorHeadLoss(nn.Module):
    """``loss_bbox`` contribution of the decoded-box GD loss in
    ``GDAnchor3DHead.loss_single`` (gd_anchor3d_head.py:102-141).

    ``forward(anchors [A0,7], bbox_pred [T,7], bbox_targets [T,7], bbox_weights [T,7],
    pos_inds=None | labels=None, num_classes=None, avg_factor=None)``:

    * ``pos_inds`` (int64 ``[P]``): the reference's ``nonzero`` result (:102-105);
    * ``labels`` (int64 ``[T]``) + ``num_classes``: positives are decided inside the
      kernel -- no ``nonzero``, no device->host sync, CUDA-graph capturable (f4).

    ``avg_factor``: a number, a one-element CUDA tensor (read by the kernel, no sync) or None.
    None with ``reduction='mean'`` averages over the positives like the reference's
    ``loss.mean()``: their number is ``len(pos_inds)``, or in ``labels`` mode counted ON THE
    DEVICE (``gd_count_positive_labels``; at least 1, so no positives give 0, :160-161).

    ``decode_weight`` is ``train_cfg['decode_weight']`` (:128-131): falsy -> no
    weights; a scalar or 7 values -> row weight
    ``mean(bbox_weights[i] * decode_weight)`` (gaussian_distance_loss.py:295-296).
    Rows with weight exactly 0 are masked (as ``GDLoss(host_sync=False)``).
    With no positives the loss is 0 with a zero gradient (:160-161).
    """

    def __init__(self, loss_decoded_bbox, decode_weight=None):
        super().__init__()
        self.loss_decoded_bbox = _as_loss(loss_decoded_bbox)
        self.decode_weight = decode_weight

    def forward(self, anchors, bbox_pred, bbox_targets, bbox_weights=None, pos_inds=None,
                labels=None, num_classes=None, avg_factor=None, **kwargs):
        loss = self.loss_decoded_bbox
        extra = dict(loss.kwargs)
        extra.update(kwargs)
        cfg = loss._shim_config(extra)
        count = int(pos_inds.numel()) if pos_inds is not None else None
        lw, mode, avg = _scale_args(loss, avg_factor, count)
        dw = self.decode_weight if self.decode_weight else None
        return ops.anchor_decoded_loss(
            anchors, bbox_pred.reshape(-1, bbox_pred.shape[-1]),
            bbox_targets.reshape(-1, bbox_targets.shape[-1]),
            None if bbox_weights is None else bbox_weights.reshape(-1, bbox_weights.shape[-1]),
            dw, cfg, lw, mode, avg, pos_inds=pos_inds, labels=labels, num_classes=num_classes)


class GDCenterHeadLoss(nn.Module):
    """``loss_gd`` of ``CenterGDHead.loss`` (gd_centerpoint_head.py:413-434):
    ``CenterPointBBoxYawCoder.decode(pos_ind[..., 1:], pred, correct_yaw=False)[..., :7]``
    (centerpoint_bbox_yaw_coders.py:18-56) fused with the loss.

    ``bbox_coder``: dict with ``pc_range``, ``out_size_factor``, ``voxel_size`` and
    optionally ``norm_bbox`` (centerpoint_bbox_coders.py:9-20; extra keys such as
    ``type`` / ``code_size`` are ignored).
    ``forward(pred [P,C], pos_ind [P,3] int64, target_box [P,>=7], weight=None,
    avg_factor=None)``; the gradient w.r.t. ``pred`` has zeros in columns >= 7.
    ``avg_factor`` may be a one-element CUDA tensor -- e.g.
    ``heatmap.eq(1).float().sum().clamp(min=1)`` WITHOUT the reference's ``.item()`` (:407):
    the kernel divides by it, the call never syncs and captures into a CUDA graph.
    """

    def __init__(self, loss_gd, bbox_coder):
        super().__init__()
        self.loss_gd = _as_loss(loss_gd)
        self._coder_args = (tuple(bbox_coder['pc_range']), bbox_coder['out_size_factor'],
                            tuple(bbox_coder['voxel_size']), bbox_coder.get('norm_bbox', True))
        self._coder = None

    @property
    def coder(self):
        if self._coder is None:
            self._coder = _lib.make_shim_center_coder(*self._coder_args)
        return self._coder

    def forward(self, pred, pos_ind, target_box, weight=None, avg_factor=None, **kwargs):
        loss = self.loss_gd
        extra = dict(loss.kwargs)
        extra.update(kwargs)
        cfg = loss._shim_config(extra)
        lw, mode, avg = _scale_args(loss, avg_factor, int(pred.shape[0]))
        return ops.center_decoded_loss(pred, pos_ind, target_box, weight, self.coder, cfg, lw,
                                       mode, avg)
