"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/gd_mismatch_golden.npz``.

Golden vectors of the UNMODIFIED reference loss (imported by path under a stub ``mmdet``,
``oracle/ref_loader.py``) in float64 on STRONGLY MISMATCHED box pairs: one, two opposite or
all three extents of the target scaled by 10 / 100 / 1000, centres shifted by 10 / 100 /
1000 m, elongated boxes at a 30 degree yaw difference.  This is the regime in which a
float32 formulation can lose every digit without any of the sigma = 0.3 / 0.05 / 0.005
parity distributions noticing (a cancellation in the bd3d shape gradient did: DESIGN.md
section 4); the fixtures pin the oracle there and give the kernels a target.

Must be run in the build container (the reference does not travel to the GPU box):
    python oracle/make_mismatch_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from mmdet3d_gaussian_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'gd_mismatch_golden.npz')
ROWS_PER_CASE = 12


def inputs():
    """[(tag, pred, target)] float32; every row is a different (what, ratio) combination."""
    pred, target, _ = synth.make_pairs(ROWS_PER_CASE * 16, 'kitti', seed=77)
    tags, i = [], 0
    t = target.clone()
    p = pred.clone()
    for what in ('w', 'h', 'l', 'wh', 'all', 'shift'):
        for ratio in (10.0, 100.0, 1000.0):
            for _ in range(ROWS_PER_CASE // 2):
                if what == 'w':
                    t[i, 3] *= ratio
                elif what == 'h':
                    t[i, 4] *= ratio
                elif what == 'l':
                    t[i, 5] *= ratio
                elif what == 'wh':
                    t[i, 3] *= ratio
                    t[i, 4] /= ratio
                elif what == 'all':
                    p[i, 3:6] *= ratio          # the PREDICTION is the big box here
                else:
                    t[i, 0] += ratio
                tags.append(f'{what}x{ratio:g}')
                i += 1
    for aspect in (10.0, 30.0):                  # elongated, 30 degrees apart
        for _ in range(ROWS_PER_CASE // 2):
            p[i, 3] *= aspect
            t[i, 3] *= aspect
            p[i, 6] = t[i, 6] + 0.5236
            tags.append(f'elong{aspect:g}')
            i += 1
    return tags, p[:i].contiguous(), t[:i].contiguous()


def main():
    mod = ref_loader.load_reference()
    tags, pred, target = inputs()
    arrays = {'pred': pred.numpy(), 'target': target.numpy()}
    manifest = {'tags': tags, 'cases': []}
    cid = 0
    for lt in ('gwd3d', 'kld3d', 'bd3d', 'jd3d', 'kld3d_symmax', 'kld3d_symmin', 'kfiou3d'):
        for fun, tau in ((('none', 0.0), ('expm1', 0.0)) if lt == 'kfiou3d'
                         else (('log1p', 0.0), ('none', 0.0), ('log1p', 1.0))):
            kw = dict(loss_type=lt, fun=fun, tau=tau, reduction='none', loss_weight=1.0)
            p = pred.double().clone().requires_grad_(True)
            out = mod.GDLoss(**kw)(p, target.double())
            out.backward(torch.ones_like(out))
            arrays[f'case/{cid:03d}/loss'] = out.detach().numpy()
            arrays[f'case/{cid:03d}/grad'] = p.grad.numpy()
            manifest['cases'].append(dict(id=f'{cid:03d}', kwargs=kw))
            cid += 1
    arrays['manifest'] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print(f'wrote {OUT}: {cid} cases x {pred.shape[0]} rows, {os.path.getsize(OUT) / 1e3:.0f} kB')


if __name__ == '__main__':
    main()
