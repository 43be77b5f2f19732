"""B200-native (sm_100a) Gaussian-distance box-regression loss.

Drop-in for the hot path of zhanggefan/mmdet3d-gaussian
(``mmdet3d_gaussian/models/losses/gaussian_distance_loss.py``): ``GDLoss`` keeps
the reference's registry name, constructor and ``forward``; the math runs in
hand-written CUDA kernels behind a C ABI (``include/gd_loss_b200.h``).  See
DESIGN.md.
"""
from .assigners import GDMaxSimAssigner, GDSimOTAAssigner, GDSimilarity3D
from .evaluation import LidarGaussianDistance, LidarGaussianSimilarity
from .heads import GDAnchorHeadLoss, GDCenterHeadLoss
from .losses import GDLoss, GDPairwiseDistance
from .registry import LOSSES, build_loss

__all__ = ['GDLoss', 'GDPairwiseDistance', 'GDAnchorHeadLoss', 'GDCenterHeadLoss',
           'GDSimilarity3D', 'GDMaxSimAssigner', 'GDSimOTAAssigner', 'LidarGaussianDistance',
           'LidarGaussianSimilarity', 'LOSSES', 'build_loss']
__version__ = '0.1.0'
