#!/usr/bin/env python
"""Small workload for `compute-sanitizer` (memcheck / racecheck / synccheck / initcheck):
every kernel family of the library once or twice at sizes of ~10^4 rows with ragged
tails, through the public module -> C ABI.  `tools/gpu_sanitize.sh` runs it under each
tool and keeps the summaries.  Exits non-zero if a result is not finite."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import GDLoss, GDPairwiseDistance, synth  # noqa: E402


def main():
    which = set(sys.argv[1:]) or {'loss', 'strided', 'pairwise'}
    torch.cuda.set_device(0)
    ok = True
    if 'loss' in which:
        for n in (10_003, 131):
            pred, target, w = synth.make_pairs(n, 'kitti', seed=2, device='cuda',
                                               weights='bernoulli')
            w7 = w[:, None].expand(n, 7).contiguous()
            for lt in ('gwd3d', 'kld3d', 'bd3d', 'jd3d'):
                for wt in (w, w7, None):
                    for red in ('mean', 'none'):
                        p = pred.clone().requires_grad_(True)
                        out = GDLoss(lt, fun='log1p', tau=0.0, reduction=red)(p, target, wt)
                        out.sum().backward()
                        ok = ok and bool(torch.isfinite(out).all()) and \
                            bool(torch.isfinite(p.grad).all())
    if 'strided' in which:
        n = 9_001
        pred, target, w = synth.make_pairs(n, 'nuscenes', seed=3, device='cuda')
        wide_p = torch.zeros(n, 9, device='cuda')
        wide_p[:, :7] = pred
        wide_t = torch.zeros(n + 1, 11, device='cuda')
        wide_t[1:, :7] = target
        for lt in ('gwd3d', 'kld3d', 'bd3d'):
            p = wide_p.clone().requires_grad_(True)
            out = GDLoss(lt, fun='log1p', tau=0.0)(p[:, :7], wide_t[1:, :7], None,
                                                   avg_factor=17.0)
            out.backward()
            ok = ok and bool(torch.isfinite(out)) and bool(torch.isfinite(p.grad).all())
            p = pred.clone().requires_grad_(True)                 # 28-byte offset views
            out = GDLoss(lt, fun='none', tau=0.0)(p[1:], target[1:], w[1:])
            out.backward()
            ok = ok and bool(torch.isfinite(out))
    if 'pairwise' in which:
        b1 = synth.make_anchor_grid(6_001, 'waymo', device='cuda')
        b2 = synth.make_targets(97, 'waymo', seed=5, device='cuda')
        for lt in ('gwd3d', 'kld3d', 'bd3d'):
            mod = GDPairwiseDistance(lt, fun='log1p', tau=1.0)
            mat = mod(b1, b2)
            rmin, ridx, cmin, cidx, _ = mod.assign(b1, b2)
            ok = ok and bool(torch.isfinite(mat).all()) and int(ridx.max()) < 97
        from mmdet3d_gaussian_b200 import GDSimOTAAssigner
        res = GDSimOTAAssigner(candidate_topk=10, loss_type='gwd3d', fun='log1p', tau=1.0).assign(
            b1, b2, want_matrix=True)
        ok = ok and int(res['assigned_gt_inds'].max()) <= 97 and int(res['topk_inds'].min()) >= 0
    torch.cuda.synchronize()
    print('sanitize_target finished, ok =', ok)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
