"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/gd_golden.npz``.

Runs the UNMODIFIED reference loss (``oracle/ref_loader.py`` imports
``/root/reference/mmdet3d_gaussian/models/losses/gaussian_distance_loss.py`` by
path under a stub ``mmdet``) on seeded synthetic boxes and stores its float64
outputs (loss + d loss/d pred) as golden vectors, plus its float32 outputs so the
tests can show the reference's own fp32-vs-fp64 error beside ours.

Must be run in the build container (the reference does not travel to the GPU
box):   python oracle/make_golden.py
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from mmdet3d_gaussian_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'gd_golden.npz')
N = 128


def input_sets():
    """name -> (pred, target) float32 CPU tensors."""
    sets = {}
    for name, ds, sigma, seed in (('kitti_s0.3', 'kitti', None, 0),
                                  ('nus_s0.3', 'nuscenes', None, 10),
                                  ('waymo_s0.05', 'waymo', 0.05, 20),
                                  ('kitti_s0.005', 'kitti', 0.005, 30)):
        p, t, _ = synth.make_pairs(N, ds, seed=seed, sigma=sigma)
        sets[name] = (p, t)
    # edge rows: extents at / below / above the clamp, negative extents,
    # large yaw, yaw + pi, identical boxes, axis-aligned equal-extent boxes.
    p, t, _ = synth.make_pairs(32, 'kitti', seed=40)
    p = p.clone()
    t = t.clone()
    p[0, 3] = 1e-7
    p[1, 3] = 5e-8
    p[2, 4] = -0.3
    p[3, 5] = 0.0
    p[4, 3:6] = torch.tensor([1e-7, 1e-7, 1e-7])
    t[5, 3:6] = torch.tensor([1e-7, 2e-8, -1.0])
    p[6, 3] = 2e7
    p[7, 5] = 1e7
    p[8, 6] += 100.0
    p[9, 6] = t[9, 6] + float(np.pi)
    p[10] = t[10]
    p[11, :3] = t[11, :3]
    p[12, 3:6] = t[12, 3:6]
    p[13, 6] = t[13, 6]
    p[14, 3:5] = torch.tensor([1e-3, 2e-3])
    t[14, 3:5] = torch.tensor([2e-3, 1e-3])
    p[15] = torch.tensor([1., 2., 3., 2., 2., 1., 0.])
    t[15] = torch.tensor([4., 6., 3., 2., 2., 1., 0.])
    p[16, 3:6] = -p[16, 3:6]
    sets['edge'] = (p, t)
    return sets


def run(mod, cls_kwargs, pred, target, weight, avg_factor, override, dtype):
    loss_mod = mod.GDLoss(**cls_kwargs)
    p = pred.to(dtype).clone().requires_grad_(True)
    t = target.to(dtype)
    w = None if weight is None else weight.to(dtype)
    try:
        out = loss_mod(p, t, w, avg_factor=avg_factor,
                       reduction_override=override)
    except (ValueError, RuntimeError, AssertionError) as e:
        return None, None, type(e).__name__
    if out.dim() == 0:
        out.backward()
    else:
        out.backward(torch.ones_like(out))
    return out.detach().numpy(), p.grad.numpy(), None


def main():
    mod = ref_loader.load_reference()
    sets = input_sets()
    arrays, manifest = {}, []
    for name, (p, t) in sets.items():
        arrays[f'in/{name}/pred'] = p.numpy()
        arrays[f'in/{name}/target'] = t.numpy()

    def add(case_id, inputs, cls_kwargs, weight_mode=None, avg_factor=None,
            override=None):
        p, t = sets[inputs]
        n = p.shape[0]
        weight = None
        if weight_mode is not None:
            g = torch.Generator().manual_seed(1234)
            weight = (torch.rand(n, generator=g) < 0.6).float() \
                * torch.rand(n, generator=g)
            if weight_mode == 'rows7':
                weight = weight[:, None] * torch.rand(n, 7, generator=g)
            elif weight_mode == 'zeros7':
                weight = torch.zeros(n, 7)
            elif weight_mode == 'zeros':
                weight = torch.zeros(n)
            arrays[f'case/{case_id}/weight'] = weight.numpy()
        entry = dict(id=case_id, inputs=inputs, kwargs=cls_kwargs,
                     weight_mode=weight_mode, avg_factor=avg_factor,
                     override=override)
        for tag, dtype in (('f64', torch.float64), ('f32', torch.float32)):
            loss, grad, err = run(mod, cls_kwargs, p, t, weight, avg_factor,
                                  override, dtype)
            if err is not None:
                entry['raises'] = err
                break
            arrays[f'case/{case_id}/loss_{tag}'] = loss
            if tag == 'f64':
                arrays[f'case/{case_id}/grad_{tag}'] = grad
            else:   # fp32 grads only needed for the error yardstick: keep norms
                arrays[f'case/{case_id}/grad_{tag}'] = grad.astype(np.float32)
        manifest.append(entry)

    cid = 0
    # ---- core grid: every distance x fun x tau x alpha x flag, rows -------
    for lt in ('gwd3d', 'kld3d', 'bd3d', 'jd3d', 'kld3d_symmax',
               'kld3d_symmin', 'kfiou3d'):
        funs = ('nlog', 'expm1', 'none') if lt == 'kfiou3d' else ('log1p', 'none')
        flag = 'normalize' if lt == 'gwd3d' else 'sqrt'
        taus = (0.0, 1.0, 2.5)
        for fun, tau, alpha, fl in itertools.product(
                funs, taus, (1.0, 0.5), (True, False)):
            if lt == 'kfiou3d' and (alpha != 1.0 or tau == 2.5):
                continue        # kfiou ignores alpha; tau is forced to 0
            if tau == 2.5 and (alpha != 1.0 or not fl):
                continue
            for inputs in ('kitti_s0.3', 'waymo_s0.05'):
                if inputs == 'waymo_s0.05' and (alpha != 1.0 or tau == 2.5):
                    continue
                kw = dict(loss_type=lt, fun=fun, tau=tau, alpha=alpha,
                          reduction='none', loss_weight=1.0)
                kw[flag] = fl
                add(f'{cid:04d}', inputs, kw)
                cid += 1
    # ---- near-identical regime + edge rows + other dataset ---------------
    for lt in ('gwd3d', 'kld3d', 'bd3d', 'jd3d', 'kld3d_symmax',
               'kld3d_symmin', 'kfiou3d'):
        fun = 'none' if lt == 'kfiou3d' else 'log1p'
        for inputs in ('kitti_s0.005', 'edge', 'nus_s0.3'):
            for tau in (0.0, 1.0):
                add(f'{cid:04d}', inputs,
                    dict(loss_type=lt, fun=fun, tau=tau, reduction='none'))
                cid += 1
    # ---- weights x reduction x avg_factor x loss_weight -------------------
    for lt in ('gwd3d', 'kld3d', 'bd3d'):
        for wm, red, af in itertools.product(
                (None, 'rows', 'rows7'), ('mean', 'sum', 'none'), (None, 37.5)):
            add(f'{cid:04d}', 'kitti_s0.3',
                dict(loss_type=lt, fun='log1p', tau=0.0, reduction=red,
                     loss_weight=5.0), weight_mode=wm, avg_factor=af)
            cid += 1
    # reduction_override, centre offsets, zero weights (early return ref:290-292)
    add(f'{cid:04d}', 'kitti_s0.3', dict(loss_type='gwd3d', reduction='none'),
        weight_mode='rows', override='mean'); cid += 1
    add(f'{cid:04d}', 'kitti_s0.3', dict(loss_type='kld3d', reduction='mean'),
        weight_mode='rows', override='sum'); cid += 1
    for off in ((0, 0, 0), (0.5, 0.5, 0.5), (0.1, -0.2, 0.3)):
        for lt in ('gwd3d', 'kld3d', 'bd3d'):
            add(f'{cid:04d}', 'kitti_s0.3',
                dict(loss_type=lt, center_offset=off, reduction='sum',
                     tau=1.0)); cid += 1
    add(f'{cid:04d}', 'kitti_s0.3', dict(loss_type='gwd3d', loss_weight=5.0),
        weight_mode='zeros7', avg_factor=11); cid += 1
    add(f'{cid:04d}', 'kitti_s0.3', dict(loss_type='bd3d', reduction='none'),
        weight_mode='zeros7'); cid += 1
    add(f'{cid:04d}', 'kitti_s0.3', dict(loss_type='kld3d'),
        weight_mode='zeros'); cid += 1      # reference raises (broadcast)

    arrays['manifest'] = np.frombuffer(
        json.dumps(manifest).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print(f'wrote {OUT}: {len(manifest)} cases, '
          f'{os.path.getsize(OUT) / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
