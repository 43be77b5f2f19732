"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/gd_decode_golden.npz``
(the decoders in front of the loss).

* ``center`` cases: the UNMODIFIED reference ``CenterPointBBoxYawCoder.decode``
  (``core/bbox/coders/centerpoint_bbox_yaw_coders.py:18-56``, loaded by
  ``oracle/ref_loader.py``) followed by the UNMODIFIED reference ``GDLoss``, i.e.
  ``gd_centerpoint_head.py:413-434`` line by line, in float64; gradient w.r.t. the
  gathered head outputs by autograd.
* ``anchor`` cases: ``gd_anchor3d_head.py:107-141`` with the reference ``GDLoss``;
  ``bbox_coder.decode`` there is upstream mmdet3d ``DeltaXYZWLHRBBoxCoder``
  (absent from the reference checkout, version unpinned), so it is the oracle's
  restatement -- these cases pin the gather / weight / decode-Jacobian / reduction
  plumbing, not the upstream coder itself (**parity unpinned** for that piece).

Build container only:   python oracle/make_decode_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gd_oracle, ref_loader  # noqa: E402
from mmdet3d_gaussian_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'gd_decode_golden.npz')


def main():
    ref = ref_loader.load_reference()
    coder_cls = ref_loader.load_reference_center_coder()
    arrays, manifest = {}, []
    cid = 0
    # ---- CenterGDHead.loss, nuScenes coder constants -------------------------
    for lt, fun, tau, extra, channels, norm in (
            ('gwd3d', 'log1p', 0.0, {}, 11, True),      # configs/nuscenes/*gwd5*
            ('gwd3d', 'none', 1.0, {'normalize': False}, 11, True),
            ('kld3d', 'log1p', 0.0, {}, 9, True),
            ('bd3d', 'log1p', 0.0, {}, 11, True),
            ('jd3d', 'none', 0.0, {'sqrt': False}, 11, True),
            ('kld3d', 'log1p', 1.0, {}, 11, False)):
        c = synth.make_center_head_batch(96, channels=channels, seed=50 + cid,
                                         coder=dict(synth.CENTER_CODER_NUS, norm_bbox=norm))
        if not norm:
            c['pred'][:, 3:6] = c['pred'][:, 3:6].exp()
        kw = dict(loss_type=lt, fun=fun, tau=tau, alpha=1.0, reduction='mean', loss_weight=5.0,
                  **extra)
        coder = coder_cls(pc_range=list(c['coder']['pc_range']),
                          out_size_factor=c['coder']['out_size_factor'],
                          voxel_size=list(c['coder']['voxel_size']), code_size=9,
                          norm_bbox=norm)
        pred = c['pred'].double().requires_grad_(True)
        target_box = coder.encode(c['target_box'].double()[:, [0, 1, 2, 3, 4, 5, 6] +
                                                            list(range(9, channels))])
        assert torch.equal(target_box[:, :7], c['target_box'].double()[:, :7])
        target_gd = target_box[..., :7]                               # head:415
        pred_gd = coder.decode(c['pos_ind'][..., 1:], pred, correct_yaw=False)[..., :7]
        avg = 41.0
        loss = ref.GDLoss(**kw)(pred_gd, target_gd, avg_factor=avg)   # head:433-434
        loss.backward()
        # the oracle's restatement must agree with the reference classes
        p2 = c['pred'].double().requires_grad_(True)
        l2 = gd_oracle.center_head_gd_loss(gd_oracle.GDLossOracle(**kw), p2, c['pos_ind'],
                                           c['target_box'].double(), c['coder'], avg_factor=avg)
        l2.backward()
        assert abs(l2.item() - loss.item()) <= 1e-12 * abs(loss.item())
        assert torch.allclose(p2.grad, pred.grad, rtol=1e-10, atol=1e-14)
        key = f'{cid:03d}'
        arrays[f'{key}/pred'] = c['pred'].numpy()
        arrays[f'{key}/pos_ind'] = c['pos_ind'].numpy()
        arrays[f'{key}/target_box'] = c['target_box'].numpy()
        arrays[f'{key}/loss_f64'] = loss.detach().numpy()
        arrays[f'{key}/grad_f64'] = pred.grad.numpy()
        manifest.append(dict(id=key, head='center', kwargs=kw, avg_factor=avg,
                             coder=dict(c['coder'], pc_range=list(c['coder']['pc_range']),
                                        voxel_size=list(c['coder']['voxel_size']))))
        cid += 1
    # ---- GDAnchor3DHead.loss_single, KITTI anchors ---------------------------
    for lt, fun, tau, dw in (('gwd3d', 'log1p', 1.0, 1),     # configs/kitti/*gwd5tau1*
                             ('kld3d', 'log1p', 1.0, 1),
                             ('bd3d', 'log1p', 1.0, [1.0, 1.0, 0.5, 2.0, 2.0, 0.25, 1.5]),
                             ('kfiou3d', 'none', 0.0, 1)):
        b = synth.make_anchor_head_batch(3000, 1200, pos_frac=0.04, seed=70 + cid)
        kw = dict(loss_type=lt, fun=fun, tau=tau, alpha=1.0, reduction='mean', loss_weight=5.0)
        bp = b['bbox_pred'].double().requires_grad_(True)
        avg = float(len(b['pos_inds'])) + 3.0
        loss = gd_oracle.anchor_head_gd_loss(
            ref.GDLoss(**kw), b['anchors'].double(), bp, b['bbox_targets'].double(),
            b['bbox_weights'].double(), b['pos_inds'], decode_weight=dw, avg_factor=avg)
        loss.backward()
        key = f'{cid:03d}'
        for name in ('anchors', 'bbox_pred', 'bbox_targets', 'bbox_weights', 'pos_inds'):
            arrays[f'{key}/{name}'] = b[name].numpy()
        arrays[f'{key}/loss_f64'] = loss.detach().numpy()
        arrays[f'{key}/grad_f64'] = bp.grad.numpy()
        manifest.append(dict(id=key, head='anchor', kwargs=kw, avg_factor=avg,
                             decode_weight=dw))
        cid += 1
    arrays['manifest'] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print(f'wrote {OUT}: {len(manifest)} cases, {os.path.getsize(OUT) / 1e3:.0f} kB')


if __name__ == '__main__':
    main()
