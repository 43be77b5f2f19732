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


# ---------------------------------------------------------------------------
# head-level inputs
# ---------------------------------------------------------------------------
KITTI_ANCHOR_Z = (-0.6, -0.6, -1.78)     # hv_pointpillars_secfpn_kitti.py:42-46

CENTER_CODER_NUS = dict(pc_range=(-51.2, -51.2), out_size_factor=4,
                        voxel_size=(0.2, 0.2), norm_bbox=True)
# configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:1,49-53,58


def make_anchor_head_batch(total_rows, anchor_rows=None, pos_frac=0.002, seed=0,
                           device='cpu', dtype=torch.float32, num_classes=3):
    """Inputs of ``GDAnchor3DHead.loss_single``'s GD branch
    (``gd_anchor3d_head.py:97-141``) for a KITTI-like anchor set.

    Returns ``dict(anchors [A0,7], bbox_pred [T,7], bbox_targets [T,7],
    bbox_weights [T,7], labels [T] int64, pos_inds [P] int64)``: anchors are the
    three KITTI class priors x rotations {0, 1.57} on a BEV grid
    (``hv_pointpillars_secfpn_kitti.py:40-49``) and repeat every ``anchor_rows``
    rows; positives (``0 <= label < num_classes``) are a ``pos_frac`` Bernoulli
    draw; target deltas ~ what ``DeltaXYZWLHRBBoxCoder.encode`` produces for a
    GT close to the anchor; predicted deltas = targets + noise."""
    a0 = int(anchor_rows or total_rows)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pri = torch.tensor(PRIORS['kitti'], device=device, dtype=dtype)
    zc = torch.tensor(KITTI_ANCHOR_Z, device=device, dtype=dtype)
    idx = torch.arange(a0, device=device)
    cls = (idx // 2) % 3
    rot = (idx % 2).to(dtype) * 1.57
    cell = idx // 6
    side = 216
    x = (cell % side).to(dtype) * (69.12 / side)
    y = (cell // side).to(dtype) * (69.12 / side) - 39.68
    anchors = torch.cat([x[:, None], y[:, None], zc[cls][:, None], pri[cls],
                         rot[:, None]], 1)
    t = total_rows
    labels = torch.full((t,), num_classes, device=device, dtype=torch.int64)
    u = torch.rand(t, generator=g, device=device)
    pos = u < pos_frac
    labels[pos] = torch.randint(0, num_classes, (t,), generator=g, device=device)[pos]
    labels[(u > 0.999)] = -1                                   # ignored rows
    scale = torch.tensor([0.3, 0.3, 0.3, 0.15, 0.15, 0.1, 0.4], device=device, dtype=dtype)
    targets = torch.randn(t, 7, generator=g, device=device, dtype=dtype) * scale
    targets = targets * pos[:, None].to(dtype)                 # zeros off the positives
    noise = torch.tensor([0.15, 0.15, 0.15, 0.2, 0.2, 0.2, 0.3], device=device, dtype=dtype)
    pred = targets + torch.randn(t, 7, generator=g, device=device, dtype=dtype) * noise
    weights = pos[:, None].to(dtype).expand(t, 7).contiguous()
    return dict(anchors=anchors, bbox_pred=pred, bbox_targets=targets,
                bbox_weights=weights, labels=labels,
                pos_inds=pos.nonzero(as_tuple=False).reshape(-1))


def make_center_head_batch(n, channels=11, batch=8, seed=0, device='cpu',
                           dtype=torch.float32, coder=None):
    """Inputs of ``CenterGDHead.loss``'s GD branch
    (``gd_centerpoint_head.py:413-434``) on the nuScenes 128x128 map: returns
    ``dict(pred [n,C], pos_ind [n,3] int64 (batch,x,y), target_box [n,C],
    coder)``.  ``target_box`` = gravity-centre GT boxes + (sin, cos, vel);
    ``pred`` = what a trained head would emit for them + noise (sub-cell offset,
    height, log-dims, yaw, dir, vel)."""
    coder = dict(coder or CENTER_CODER_NUS)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    boxes = make_targets(n, 'nuscenes', seed=seed + 7, device=device, dtype=dtype)
    span = -2.0 * coder['pc_range'][0]
    boxes[:, 0] = torch.rand(n, generator=g, device=device, dtype=dtype) * span * 0.98 \
        + coder['pc_range'][0] + 0.01 * span
    boxes[:, 1] = torch.rand(n, generator=g, device=device, dtype=dtype) * span * 0.98 \
        + coder['pc_range'][1] + 0.01 * span
    cellx = coder['out_size_factor'] * coder['voxel_size'][0]
    celly = coder['out_size_factor'] * coder['voxel_size'][1]
    fx = (boxes[:, 0] - coder['pc_range'][0]) / cellx
    fy = (boxes[:, 1] - coder['pc_range'][1]) / celly
    xi, yi = fx.long(), fy.long()
    b = torch.randint(0, batch, (n,), generator=g, device=device)
    pos_ind = torch.stack([b, xi, yi], 1)
    noise = torch.randn(n, channels, generator=g, device=device, dtype=dtype)
    pred = torch.zeros(n, channels, device=device, dtype=dtype)
    pred[:, 0] = fx - xi.to(dtype) + 0.3 * noise[:, 0]
    pred[:, 1] = fy - yi.to(dtype) + 0.3 * noise[:, 1]
    pred[:, 2] = boxes[:, 2] + 0.3 * noise[:, 2]
    pred[:, 3:6] = boxes[:, 3:6].log() + 0.2 * noise[:, 3:6]
    pred[:, 6] = boxes[:, 6] + 0.3 * noise[:, 6]
    pred[:, 7:] = noise[:, 7:]
    others = torch.randn(n, channels - 9, generator=g, device=device, dtype=dtype)
    target_box = torch.cat([boxes, boxes[:, 6:7].sin(), boxes[:, 6:7].cos(), others], 1)
    return dict(pred=pred, pos_ind=pos_ind, target_box=target_box, coder=coder)
