"""Evaluation-side consumer of the pairwise Gaussian distance.

The reference's evaluator matches detections to ground truth through an
*affinity calculator* taken from the ``EVAL_AFFINITYCALS`` registry
(``mmdet3d_gaussian/core/evaluation/builder.py:5,16-17``).  The interface is
(``core/evaluation/affinity.py:5-32``):

* class attribute ``LARGER_CLOSER`` (bool: is a larger value a better match);
* ``__call__(det_bboxes [N,>=7] float numpy, gt_bboxes [M,>=7] float numpy,
  gt_iscrowd=None) -> [N, M] float32 numpy``; crowd annotations are refused with
  the same assertion the reference's calculators raise.

The shipped calculators loop over all pairs on one host thread
(``ops/eval/affinity.cpp:8-105``).  ``LidarGaussianDistance`` offers the GD family
through the same interface with the N x M matrix produced by the pairwise CUDA
kernel (``csrc/gd_pairwise.cu``): host arrays are copied to the current CUDA device,
one launch fills the matrix, and it is copied back.  There is no host fallback.
"""
import numpy as np
import torch

from . import _lib, ops
from .losses.gaussian_distance_loss import _FLAG
from .registry import Registry

EVAL_AFFINITYCALS = Registry('eval_affinity_calculator')   # builder.py:5


def build_eval_affinity_calculator(cfg, **default_args):
    """Same role as ``core/evaluation/builder.py:16-17``."""
    return EVAL_AFFINITYCALS.build(cfg, default_args)


def _register(cls):
    EVAL_AFFINITYCALS.register_module(force=True)(cls)
    try:        # the reference package, when importable, owns the real registry
        from mmdet3d_gaussian.core.evaluation.builder import EVAL_AFFINITYCALS as REF
        REF.register_module(force=True)(cls)
    except Exception:
        pass
    return cls


@_register
class LidarGaussianDistance:
    """``affinity[i, j] = post(distance(det[i], gt[j]))`` -- smaller is closer.

    Same ``loss_type`` / ``fun`` / ``tau`` / ``alpha`` / ``center_offset`` /
    ``normalize|sqrt`` meaning as ``GDLoss``; with ``tau >= 1`` the value lies in
    ``[0, 1)``, so matcher thresholds can be written like ``1 - IoU`` thresholds.
    """
    LARGER_CLOSER = False

    def __init__(self, loss_type='gwd3d', center_offset=(0, 0, 0.5), fun='log1p', tau=1.0,
                 alpha=1.0, device=None, **kwargs):
        assert loss_type in _lib.LOSS_TYPES
        if loss_type != 'kfiou3d':
            assert fun in ['log1p', 'none']
        else:
            assert fun in ['nlog', 'expm1', 'none']
        name, default = _FLAG[loss_type]
        unknown = set(kwargs) - {name}
        if unknown:
            raise TypeError(f'unexpected keyword argument {sorted(unknown)[0]!r}')
        self.cfg = _lib.make_config(loss_type, fun, kwargs.get(name, default), tau, alpha,
                                    center_offset)
        self.device = device

    def _dev(self):
        if not torch.cuda.is_available():
            raise RuntimeError('LidarGaussianDistance needs a CUDA device (no host fallback)')
        return torch.device('cuda', torch.cuda.current_device()) if self.device is None \
            else torch.device(self.device)

    def __call__(self, det_bboxes, gt_bboxes, gt_iscrowd=None):
        assert gt_iscrowd is None, 'Does not support crowd annotation yet'   # affinity.py:10
        dev = self._dev()
        det = torch.as_tensor(np.ascontiguousarray(det_bboxes, dtype=np.float32))
        gt = torch.as_tensor(np.ascontiguousarray(gt_bboxes, dtype=np.float32))
        det = det.reshape(-1, det.shape[-1] if det.ndim > 1 else 7)
        gt = gt.reshape(-1, gt.shape[-1] if gt.ndim > 1 else 7)
        if det.shape[0] == 0 or gt.shape[0] == 0:
            return np.zeros((det.shape[0], gt.shape[0]), dtype=np.float32)
        mat = ops.pairwise_distance(det[:, :7].to(dev), gt[:, :7].to(dev), self.cfg)
        return mat.cpu().numpy()


@_register
class LidarGaussianSimilarity(LidarGaussianDistance):
    """``1 - LidarGaussianDistance``: an IoU-like score (``tau/(tau+f(d))`` for
    ``tau >= 1``) for matchers configured with ``LARGER_CLOSER = True`` thresholds."""
    LARGER_CLOSER = True

    def __call__(self, det_bboxes, gt_bboxes, gt_iscrowd=None):
        return np.float32(1.0) - super().__call__(det_bboxes, gt_bboxes, gt_iscrowd)
