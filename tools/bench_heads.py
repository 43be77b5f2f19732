#!/usr/bin/env python
"""Head-level timing of the fused front ends.

For the GD branch of GDAnchor3DHead.loss_single (KITTI 3-class: 248x216x6 = 321,408
anchors per sample, batch 6 per GPU) and of CenterGDHead.loss (nuScenes), times
  fused     : one launch (gather + decode + loss + gradient to the raw outputs)
  unfused   : the reference's op sequence on the GPU -- torch nonzero / gathers /
              decode in eager torch -- around OUR element-wise GDLoss kernel
  cpu_port  : the oracle port of the same call site on the host cores
forward + backward per call, CUDA events, and prints one JSON document.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import GDAnchorHeadLoss, GDCenterHeadLoss, GDLoss, synth  # noqa: E402
from oracle import gd_oracle  # noqa: E402

KW = dict(loss_type='gwd3d', fun='log1p', tau=1.0, loss_weight=5.0)


def gpu_time(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6   # us


def anchor_case(total, anchor_rows, pos_frac, out):
    b = synth.make_anchor_head_batch(total, anchor_rows, pos_frac=pos_frac, seed=0, device='cuda')
    npos = int(b['pos_inds'].numel())
    avg = float(max(npos, 1))
    head = GDAnchorHeadLoss(dict(type='GDLoss', **KW), 1)
    loss_mod = GDLoss(**KW)
    bp = b['bbox_pred'].clone().requires_grad_(True)

    def fused_labels():
        bp.grad = None
        head(b['anchors'], bp, b['bbox_targets'], b['bbox_weights'], labels=b['labels'],
             num_classes=3, avg_factor=avg).backward()

    def fused_index():
        bp.grad = None
        pos = ((b['labels'] >= 0) & (b['labels'] < 3)).nonzero(as_tuple=False).reshape(-1)
        head(b['anchors'], bp, b['bbox_targets'], b['bbox_weights'], pos_inds=pos,
             avg_factor=avg).backward()

    reps = -(-total // anchor_rows)

    def unfused():
        bp.grad = None
        pos = ((b['labels'] >= 0) & (b['labels'] < 3)).nonzero(as_tuple=False).reshape(-1)
        anchors = b['anchors'].repeat(reps, 1)[:total][pos]
        w = b['bbox_weights'][pos] * b['bbox_weights'].new_tensor(1)
        pd = gd_oracle.decode_delta_xyzwlhr(anchors, bp[pos])
        td = gd_oracle.decode_delta_xyzwlhr(anchors, b['bbox_targets'][pos])
        loss_mod(pd, td, w, avg_factor=avg).backward()

    res = dict(head='anchor', total_rows=total, positives=npos)
    for name, fn in (('fused_labels', fused_labels), ('fused_index', fused_index),
                     ('unfused', unfused)):
        dev_us, wall_us = gpu_time(fn)
        res[name + '_us'] = round(dev_us, 2)
        res[name + '_wall_us'] = round(wall_us, 2)
    res['fused_labels_GBps'] = round((36.0 * total + 112.0 * npos) / res['fused_labels_us'] / 1e3, 1)
    # CPU port of the same call site
    c = {k: v.cpu() for k, v in b.items()}
    mod = gd_oracle.GDLossOracle(**KW)
    best = float('inf')
    for _ in range(3):
        p = c['bbox_pred'].clone().requires_grad_(True)
        t0 = time.perf_counter()
        pos = ((c['labels'] >= 0) & (c['labels'] < 3)).nonzero(as_tuple=False).reshape(-1)
        gd_oracle.anchor_head_gd_loss(mod, c['anchors'], p, c['bbox_targets'], c['bbox_weights'],
                                      pos, decode_weight=1, avg_factor=avg).backward()
        best = min(best, time.perf_counter() - t0)
    res['cpu_port_us'] = round(best * 1e6, 1)
    out.append(res)


def center_case(n, out):
    c = synth.make_center_head_batch(n, seed=0, device='cuda')
    kw = dict(KW, tau=0.0)
    head = GDCenterHeadLoss(dict(type='GDLoss', **kw), c['coder'])
    loss_mod = GDLoss(**kw)
    p = c['pred'].clone().requires_grad_(True)

    def fused():
        p.grad = None
        head(p, c['pos_ind'], c['target_box'], avg_factor=float(n)).backward()

    def unfused():
        p.grad = None
        dec = gd_oracle.decode_centerpoint_yaw(c['pos_ind'][..., 1:], p, **c['coder'])[..., :7]
        loss_mod(dec, c['target_box'][..., :7], avg_factor=float(n)).backward()
    res = dict(head='center', rows=n)
    for name, fn in (('fused', fused), ('unfused', unfused)):
        dev_us, wall_us = gpu_time(fn)
        res[name + '_us'] = round(dev_us, 2)
        res[name + '_wall_us'] = round(wall_us, 2)
    cc = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in c.items()}
    mod = gd_oracle.GDLossOracle(**kw)
    best = float('inf')
    for _ in range(3):
        q = cc['pred'].clone().requires_grad_(True)
        t0 = time.perf_counter()
        gd_oracle.center_head_gd_loss(mod, q, cc['pos_ind'], cc['target_box'], cc['coder'],
                                      avg_factor=float(n)).backward()
        best = min(best, time.perf_counter() - t0)
    res['cpu_port_us'] = round(best * 1e6, 1)
    out.append(res)


def main():
    torch.cuda.set_device(0)
    torch.set_num_threads(os.cpu_count() or 1)
    out = []
    per_sample = 248 * 216 * 6
    anchor_case(per_sample * 6, per_sample, 0.0006, out)       # KITTI, 6 samples / GPU
    anchor_case(per_sample * 12, per_sample, 0.0006, out)      # gwd5tau1 config: 12 / GPU
    anchor_case(1 << 24, 1 << 20, 0.01, out)                   # stress: 16.8M rows
    center_case(500, out)                                      # nuScenes: ~10^2-10^3 objects
    center_case(1 << 20, out)
    print(json.dumps({'cores': os.cpu_count(), 'cases': out}))


if __name__ == '__main__':
    main()
