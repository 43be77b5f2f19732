"""Affinity-calculator consumer of the pairwise kernel:
interface of the reference's ``EVAL_AFFINITYCALS`` entries
(core/evaluation/affinity.py:5-32, builder.py:5,16-17)."""
import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import synth
from mmdet3d_gaussian_b200.evaluation import (EVAL_AFFINITYCALS, LidarGaussianDistance,
                                              LidarGaussianSimilarity,
                                              build_eval_affinity_calculator)
from oracle import gd_oracle

HAVE_GPU = torch.cuda.is_available()


def test_registry_and_interface_host_side():
    calc = build_eval_affinity_calculator(dict(type='LidarGaussianDistance', loss_type='kld3d',
                                               fun='log1p', tau=1.0))
    assert isinstance(calc, LidarGaussianDistance) and calc.LARGER_CLOSER is False
    assert LidarGaussianSimilarity.LARGER_CLOSER is True
    assert 'LidarGaussianSimilarity' in EVAL_AFFINITYCALS
    with pytest.raises(AssertionError):                 # ctor asserts like GDLoss (ref:265-270)
        LidarGaussianDistance(loss_type='kfiou3d', fun='log1p')
    with pytest.raises(TypeError):
        LidarGaussianDistance(loss_type='gwd3d', sqrt=True)
    with pytest.raises(AssertionError, match='crowd'):  # affinity.py:10
        calc(np.zeros((1, 7), np.float32), np.zeros((1, 7), np.float32), gt_iscrowd=[0])
    if not HAVE_GPU:                                    # no host fallback: must fail loudly
        with pytest.raises(RuntimeError, match='CUDA'):
            calc(np.ones((2, 7), np.float32), np.ones((3, 7), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d', 'jd3d'])
def test_affinity_vs_oracle(loss_type):
    det, _, _ = synth.make_pairs(513, 'kitti', seed=31)
    gt = synth.make_targets(29, 'kitti', seed=32)
    det9 = np.concatenate([det.numpy(), np.random.RandomState(0).rand(513, 2)], 1)  # score cols
    calc = LidarGaussianDistance(loss_type, fun='log1p', tau=1.0)
    aff = calc(det9.astype(np.float64), gt.numpy())
    assert aff.dtype == np.float32 and aff.shape == (513, 29)
    ref = gd_oracle.pairwise_distance(det.double(), gt.double(), loss_type, fun='log1p',
                                      tau=1.0).numpy()
    err = np.abs(aff - ref) / np.maximum(np.abs(ref), 1e-3)
    assert err.max() <= 1e-5, err.max()
    sim = LidarGaussianSimilarity(loss_type, fun='log1p', tau=1.0)(det9, gt.numpy())
    assert np.array_equal(sim, np.float32(1.0) - aff)
    assert (sim > 0).all() and (sim <= 1).all()
    # matching decisions (per-detection best GT) agree with the oracle wherever the
    # two best candidates are separated by more than the tolerance
    top2 = np.sort(ref, axis=1)[:, :2]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5 * np.maximum(np.abs(top2[:, 1]), 1e-3)
    assert np.array_equal(aff.argmin(1)[clear], ref.argmin(1)[clear])


@pytest.mark.gpu
def test_affinity_empty_inputs():
    calc = LidarGaussianDistance('gwd3d')
    assert calc(np.zeros((0, 7), np.float32), np.ones((4, 7), np.float32)).shape == (0, 4)
    assert calc(np.ones((5, 9), np.float32), np.zeros((0, 7), np.float32)).shape == (5, 0)
