"""Drop-in ``GDLoss`` backed by the sm_100a kernels.

Host-side mirror of the reference module
``mmdet3d_gaussian/models/losses/gaussian_distance_loss.py`` (``ref:LINE``):
same registry name, constructor (ref:261-278), ``forward`` signature
(ref:280-286), assertions and error behaviour; the arithmetic of ref:8-248 and of
mmdet's ``weighted_loss`` wrapper runs in one fused CUDA kernel
(``csrc/gd_loss_kernels.cu``) reached through the C ABI.
"""
from copy import deepcopy

import torch
from torch import nn

from .. import _lib, ops
from ..registry import register_everywhere

# loss_type -> (name of the one extra kwarg the distance accepts, its default)
_FLAG = {'gwd3d': ('normalize', True),        # ref:43
         'kld3d': ('sqrt', True),             # ref:110
         'jd3d': ('sqrt', True),              # ref:190
         'kld3d_symmax': ('sqrt', True),      # ref:202-203
         'kld3d_symmin': ('sqrt', True),      # ref:215-216
         'bd3d': ('sqrt', True),              # ref:145
         'kfiou3d': ('sqrt', False)}          # ref:228 (accepted, ignored)


def _scale_and_mode(reduction, avg_factor, n, loss_weight):
    """mmdet ``weight_reduce_loss`` folded into one scalar, as the
    C++ shim applies it (``csrc/torch_shim.cpp``; this Python statement of the rule is what
    the CPU tests check).  Returns (scale, rows_out)."""
    if avg_factor is None:
        if reduction == 'mean':
            # loss.mean(): divides by N (not by sum of weights); mean of empty is nan
            return (loss_weight / n if n > 0 else float('nan')), False
        if reduction == 'sum':
            return loss_weight, False
        return loss_weight, True
    if reduction == 'mean':
        return loss_weight / float(avg_factor), False
    if reduction == 'none':
        return loss_weight, True
    raise ValueError('avg_factor can not be used with reduction="sum"')


@register_everywhere
class GDLoss(nn.Module):
    """Gaussian-distance box regression loss (GWD / KLD / JD / sym-KLD / BCD / KFIoU).

    Args mirror the reference (ref:261-263).  ``**kwargs`` are forwarded to the
    distance: ``normalize`` for ``gwd3d``, ``sqrt`` for the others.  Two extra,
    backwards-compatible keys are consumed here and never reach the distance:

    * ``variant`` ('auto' | 'staged' | 'bulk' | 'bulk_packed' | 'bulk_any') pins the kernel
      variant (tests / measurements; 'auto' picks the fastest one the layout allows);
    * ``host_sync`` (default True).  True = the reference's semantics exactly, early return
      (ref:290-292) included, WITHOUT stalling the GPU: for ``[N,7]`` weights (the KITTI head's
      call, ``gd_anchor3d_head.py:128-141``) and wherever else ``pred * weight`` has the shape
      of ``pred``, ``any(weight > 0)`` is evaluated inside the fused launch and a second, tiny
      launch swaps in ``(pred * weight).sum()`` / ``grad = weight`` when it is false -- no
      host involvement at all, CUDA-graph capturable.  For other ``[N]`` weights the
      reference RAISES a broadcasting error when the branch is taken, so the host must learn
      the answer: the fused launch itself reports ``any(weight > 0)`` into a word of pinned host
      memory (its first warp, microseconds after the kernel starts, when the first tile holds a
      positive weight; its last CTA otherwise) and the host waits for that word -- the GPU never
      idles.  ``weight=None`` and ``reduction='none'`` never probe (as the reference).
      False never waits and never raises: rows whose weight is exactly 0 are masked inside
      the kernel (0 loss, 0 gradient, even where the distance is inf/nan); for non-negative
      weights the value and gradient equal the reference's on every finite row.
      ``'overlap'`` is accepted as an alias of True (round-1 name of this behaviour).

    ``avg_factor`` may be a Python number or a one-element CUDA tensor; the tensor is read by
    the kernel (no ``.item()``), which is what lets the heads drop their host syncs
    (``gd_centerpoint_head.py:407``).
    """

    BAG_GD_LOSS = tuple(_lib.LOSS_TYPES)          # ref:253-259

    def __init__(self, loss_type, center_offset=(0, 0, 0.5), fun='log1p',
                 tau=1.0, alpha=1.0, reduction='mean', loss_weight=1.0, **kwargs):
        super().__init__()
        assert reduction in ['none', 'sum', 'mean']               # ref:265
        assert loss_type in self.BAG_GD_LOSS                      # ref:266
        if loss_type not in ['kfiou3d']:
            assert fun in ['log1p', 'none']                       # ref:267-268
        else:
            assert fun in ['nlog', 'expm1', 'none']               # ref:269-270
        self.loss_type = loss_type
        self.center_offset = center_offset
        self.fun = fun
        self.tau = tau
        self.alpha = alpha
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.variant = kwargs.pop('variant', 'auto')
        hs = kwargs.pop('host_sync', True)
        self.host_sync = True if hs == 'overlap' else bool(hs)
        self.kwargs = kwargs                                      # ref:278
        self._cfg_cache = {}
        self._peer_sum = None        # set by sharded.ShardedGDLoss(fused=True)

    def _key(self, extra):
        name, default = _FLAG[self.loss_type]
        if extra:
            unknown = set(extra) - {name}
            if unknown:
                # the reference forwards **kwargs into the distance function, which
                # raises TypeError on names it does not accept (ref:301-310)
                raise TypeError(f'{self.loss_type}_loss() got an unexpected keyword '
                                f'argument {sorted(unknown)[0]!r}')
            default = bool(extra.get(name, default))
        return (self.loss_type, self.fun, default, float(self.tau), float(self.alpha),
                tuple(self.center_offset))

    def _config(self, extra):
        """ctypes ``gd_loss_config`` (pairwise / direct C-ABI callers)."""
        key = ('c',) + self._key(extra)
        cfg = self._cfg_cache.get(key)
        if cfg is None:
            cfg = self._cfg_cache[key] = _lib.make_config(*key[1:])
        return cfg

    def _shim_config(self, extra):
        """``gd_loss_config`` held by the C++ shim."""
        key = self._key(extra)
        cfg = self._cfg_cache.get(key)
        if cfg is None:
            cfg = self._cfg_cache[key] = _lib.make_shim_config(*key)
        return cfg

    def forward(self, pred, target, weight=None, avg_factor=None,
                reduction_override=None, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')   # ref:287
        reduction = (
            reduction_override if reduction_override else self.reduction)  # ref:288-289
        if self.kwargs or kwargs:
            _kwargs = deepcopy(self.kwargs)                       # ref:293
            _kwargs.update(kwargs)                                # ref:294
        else:
            _kwargs = None
        # ref:290-292 (early return), ref:295-310 and mmdet's weighted_loss: in the shim
        return ops.gd_loss(pred, target, weight, self._shim_config(_kwargs), self.loss_weight,
                           reduction, avg_factor, self.variant,
                           mask_zero_weight=not self.host_sync, peer_sum=self._peer_sum)

    def extra_repr(self):
        return (f'loss_type={self.loss_type!r}, fun={self.fun!r}, tau={self.tau}, '
                f'alpha={self.alpha}, reduction={self.reduction!r}, '
                f'loss_weight={self.loss_weight}')


class GDPairwiseDistance(nn.Module):
    """Batched pairwise matrix ``D[i,j] = post(distance(boxes1[i], boxes2[j]))``
    (new surface): same ``loss_type`` / ``fun`` / ``tau`` /
    ``alpha`` / ``center_offset`` / ``normalize|sqrt`` meaning as ``GDLoss``; no
    weights, no reduction, no gradient."""

    def __init__(self, loss_type, center_offset=(0, 0, 0.5), fun='log1p', tau=1.0,
                 alpha=1.0, **kwargs):
        super().__init__()
        assert loss_type in _lib.LOSS_TYPES
        if loss_type != 'kfiou3d':
            assert fun in ['log1p', 'none']
        else:
            assert fun in ['nlog', 'expm1', 'none']
        name, default = _FLAG[loss_type]
        unknown = set(kwargs) - {name}
        if unknown:
            raise TypeError(f'unexpected keyword argument {sorted(unknown)[0]!r}')
        self.loss_type = loss_type
        self.cfg = _lib.make_config(loss_type, fun, kwargs.get(name, default), tau,
                                    alpha, center_offset)

    @torch.no_grad()
    def forward(self, boxes1, boxes2):
        return ops.pairwise_distance(boxes1, boxes2, self.cfg)

    @torch.no_grad()
    def assign(self, boxes1, boxes2, want_matrix=False, cpl1=False):
        """Row and column ``(min, argmin)`` in one launch, matrix optional:
        ``(row_min, row_argmin, col_min, col_argmin, matrix | None)``."""
        return ops.pairwise_assign(boxes1, boxes2, self.cfg, want_matrix=want_matrix, cpl1=cpl1)

    @torch.no_grad()
    def row_argmin(self, boxes1, boxes2):
        """``(min_j D[i,j], argmin_j D[i,j])`` without writing the matrix."""
        return ops.pairwise_row_argmin(boxes1, boxes2, self.cfg)
