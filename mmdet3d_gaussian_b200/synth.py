"""Deterministic synthetic box-pair generator.

Shared by ``bench.py``, the tests and ``oracle/make_golden.py`` so that every
arm sees the same seeded inputs.  All boxes are rows ``(x, y, z, w, h, l, r)``
(reference naming, ``gaussian_distance_loss.py:8``): columns 3,4 are the BEV
extents, column 5 the vertical extent, column 6 the yaw.

Class size priors (BEV, BEV, vertical) come from the reference configs:

* KITTI  -- ``configs/_base_/models/hv_pointpillars_secfpn_kitti.py:47``
* Waymo  -- ``configs/_base_/models/hv_pointpillars_secfpn_waymo.py:51-55``
* nuScenes -- the reference's CenterPoint configs carry only class names
  (``configs/nuscenes/centerpoint_02pillar_second_secfpn_gwd5_8x4_cyclic_20e_nus.py:11-14``),
  no sizes; the table below uses the customary nuScenes per-class mean sizes
  (car, truck, construction_vehicle, bus, trailer, barrier, motorcycle,
  bicycle, pedestrian, traffic_cone).
"""
import math

import torch

PRIORS = {
    'kitti': [[0.8, 0.6, 1.73], [1.76, 0.6, 1.73], [3.9, 1.6, 1.56]],
    'waymo': [[4.73, 2.08, 1.77], [1.81, 0.84, 1.77], [0.91, 0.84, 1.74]],
    'nuscenes': [[4.63, 1.97, 1.74], [6.93, 2.51, 2.84], [6.37, 2.85, 3.19],
                 [10.5, 2.94, 3.47], [12.29, 2.90, 3.87], [0.50, 2.53, 0.98],
                 [2.11, 0.77, 1.47], [1.70, 0.60, 1.28], [0.73, 0.67, 1.77],
                 [0.41, 0.41, 1.07]],
}


def make_targets(n, dataset='kitti', seed=0, device='cpu', dtype=torch.float32):
    """``n`` target boxes: x~U(0,70), y~U(-40,40), z~N(-1,0.5^2), extents =
    class prior * exp(N(0,0.1^2)), yaw~U(-pi,pi)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pri = torch.tensor(PRIORS[dataset], device=device, dtype=dtype)
    x = torch.rand(n, generator=g, device=device, dtype=dtype) * 70.0
    y = torch.rand(n, generator=g, device=device, dtype=dtype) * 80.0 - 40.0
    z = torch.randn(n, generator=g, device=device, dtype=dtype) * 0.5 - 1.0
    cls = torch.randint(0, pri.shape[0], (n,), generator=g, device=device)
    ext = pri[cls] * torch.exp(
        torch.randn(n, 3, generator=g, device=device, dtype=dtype) * 0.1)
    yaw = (torch.rand(n, generator=g, device=device, dtype=dtype) * 2.0 - 1.0) \
        * math.pi
    return torch.cat([x[:, None], y[:, None], z[:, None], ext, yaw[:, None]], 1)


def perturb(target, sigma_xyz=0.3, sigma_ext=0.2, sigma_yaw=0.3, seed=1):
    """Predictions = targets with xyz += N(0,s^2), extents *= exp(N(0,s^2)),
    yaw += N(0,s^2) (strictly positive extents, well-conditioned regime)."""
    g = torch.Generator(device=target.device)
    g.manual_seed(seed)
    n = target.shape[0]
    noise = torch.randn(n, 7, generator=g, device=target.device,
                        dtype=target.dtype)
    pred = target.clone()
    pred[:, 0:3] += noise[:, 0:3] * sigma_xyz
    pred[:, 3:6] *= torch.exp(noise[:, 3:6] * sigma_ext)
    pred[:, 6] += noise[:, 6] * sigma_yaw
    return pred


def make_pairs(n, dataset='kitti', seed=0, device='cpu', dtype=torch.float32,
               sigma=None, weights='ones'):
    """Returns ``(pred, target, weight)``.

    ``sigma`` scales the three perturbation widths together relative to the
    default (0.3, 0.2, 0.3): ``sigma=0.05`` means (0.05, 0.0333, 0.05).
    ``weights``: ``'ones'`` -> ``[n]`` ones, ``'bernoulli'`` ->
    ``Bernoulli(0.5)*U(0,1)`` ``[n]`` (config C3), ``'rows7'`` -> the same
    expanded to ``[n,7]``, ``None`` -> no weights.
    """
    target = make_targets(n, dataset, seed, device, dtype)
    if sigma is None:
        pred = perturb(target, seed=seed + 1)
    else:
        pred = perturb(target, sigma, sigma * (2.0 / 3.0), sigma, seed=seed + 1)
    if weights is None:
        w = None
    elif weights == 'ones':
        w = torch.ones(n, device=device, dtype=dtype)
    else:
        g = torch.Generator(device=device)
        g.manual_seed(seed + 2)
        keep = (torch.rand(n, generator=g, device=device) < 0.5).to(dtype)
        w = keep * torch.rand(n, generator=g, device=device, dtype=dtype)
        if weights == 'rows7':
            w = w[:, None].expand(n, 7).contiguous()
    return pred, target, w


def make_anchor_grid(n, dataset='waymo', seed=0, device='cpu',
                     dtype=torch.float32):
    """Pairwise config C4: ``n`` anchors = class priors x rotations {0, 1.57}
    on a jittered BEV grid."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pri = torch.tensor(PRIORS[dataset], device=device, dtype=dtype)
    idx = torch.arange(n, device=device)
    k = pri.shape[0]
    cls = idx % k
    rot = ((idx // k) % 2).to(dtype) * 1.57
    cell = idx // (2 * k)
    side = max(int(math.ceil(math.sqrt(max(int(cell.max().item()) + 1, 1)))), 1)
    x = (cell % side).to(dtype) * (150.0 / side) - 75.0
    y = (cell // side).to(dtype) * (150.0 / side) - 75.0
    z = torch.full((n,), -0.0345, device=device, dtype=dtype)
    return torch.cat([x[:, None], y[:, None], z[:, None], pri[cls],
                      rot[:, None]], 1)
