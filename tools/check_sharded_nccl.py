#!/usr/bin/env python
"""Multi-GPU parity of the sharded path (BASELINE.json configs[2], "C3": nuScenes-scale
weighted GD loss, 786,432 rows sharded by contiguous row blocks, one NCCL all-reduce of
the scalar).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29541 tools/check_sharded_nccl.py

Every rank evaluates its shard with the CUDA kernel; rank 0 gathers the gradients and
checks loss and gradients against the fp64 CPU oracle of the whole batch (1e-5 relative),
then times the sharded call (CUDA events, max over ranks).  Prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import GDLoss, GDMaxSimAssigner, sharded, synth  # noqa: E402
from oracle import gd_oracle  # noqa: E402

N = 786_432


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    pred, target, w = synth.make_pairs(N, 'nuscenes', seed=3, weights='bernoulli')
    lo, hi = sharded.shard_bounds(N, rank, world)
    avg = float(max(int((w > 0).sum()), 1))
    report = {'world_size': world, 'rows': N, 'cases': []}
    ok = True
    fused_flag = '--fused' in sys.argv
    report['fused_requested'] = fused_flag
    for lt, red, af, fused in (('gwd3d', 'mean', avg, False), ('gwd3d', 'mean', avg, fused_flag),
                               ('kld3d', 'mean', None, False), ('bd3d', 'sum', None, fused_flag),
                               ('kld3d', 'mean', torch.tensor(avg, device=dev), fused_flag)):
        kw = dict(loss_type=lt, fun='log1p', tau=0.0, loss_weight=5.0, reduction=red)
        mod = sharded.ShardedGDLoss(GDLoss(host_sync=False, **kw), fused=fused)
        p = pred[lo:hi].to(dev).requires_grad_(True)
        t, ww = target[lo:hi].to(dev), w[lo:hi].to(dev)
        loss = mod(p, t, ww, avg_factor=af)
        loss.backward()
        grads = [None] * world                 # shards may be ragged: gather as objects
        dist.all_gather_object(grads, p.grad.cpu())
        vals = [None] * world
        dist.all_gather_object(vals, float(loss.item()))
        if rank == 0 and len(set(vals)) != 1:
            ok = False
            report.setdefault('rank_disagreement', []).append(vals)
        # timing of the sharded call (forward + backward + the collective)
        for _ in range(5):
            p.grad = None
            mod(p, t, ww, avg_factor=af).backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            p.grad = None
            mod(p, t, ww, avg_factor=af).backward()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            full = torch.cat([g.cpu() for g in grads]).double()
            af_host = float(af) if torch.is_tensor(af) else af
            ref_l, ref_g = gd_oracle.loss_and_grad(gd_oracle.GDLossOracle(**kw), pred.double(),
                                                   target.double(), w.double(), avg_factor=af_host)
            rel = abs(loss.item() - ref_l.item()) / abs(ref_l.item())
            gn = ref_g.norm(dim=1).clamp_min(1e-2 * ref_g.norm(dim=1).max().item() * 1e-3)
            fin = torch.isfinite(ref_g).all(dim=1)
            gerr = ((full - ref_g).norm(dim=1) / gn)[fin].max().item()
            good = rel <= 1e-5 and gerr <= 1e-5
            ok = ok and good
            report['cases'].append({'loss_type': lt, 'reduction': red,
                                    'avg_factor': af_host, 'avg_factor_on_device': torch.is_tensor(af),
                                    'fused_in_kernel_sum': bool(mod.fused),
                                    'loss': loss.item(), 'oracle': ref_l.item(),
                                    'loss_rel_err': rel, 'grad_max_row_rel_err': gerr,
                                    'ms_per_call_max_over_ranks': round(float(ms.item()), 4),
                                    'ok': good})
    # pairwise path (C4-like): anchors sharded, GTs replicated, per-GT minima merged with
    # one all-reduce(MIN) of 64-bit (value, anchor) keys; labels must equal the
    # single-GPU assigner's bit for bit
    na, m = 200_000, 256
    anchors = synth.make_anchor_grid(na, 'waymo')
    gts = synth.make_targets(m, 'waymo', seed=5)
    gts[:, 0] = gts[:, 0] * 2.0 - 70.0
    base = GDMaxSimAssigner(0.6, 0.45, 0.45, True, loss_type='gwd3d', fun='log1p', tau=1.0)
    alo, ahi = sharded.shard_bounds(na, rank, world)
    smod = sharded.ShardedGDMaxSimAssigner(base)
    a_loc, g_dev = anchors[alo:ahi].to(dev), gts.to(dev)
    res = smod.assign(a_loc, g_dev, alo)
    labels = [None] * world
    dist.all_gather_object(labels, res['assigned_gt_inds'].cpu())
    for _ in range(5):
        smod.assign(a_loc, g_dev, alo)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        smod.assign(a_loc, g_dev, alo)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = base.assign(anchors.to(dev), g_dev)
        same = bool(torch.equal(torch.cat(labels), full['assigned_gt_inds'].cpu())) and \
            bool(torch.equal(res['gt_argmax_overlaps'].cpu(), full['gt_argmax_overlaps'].cpu())) and \
            bool(torch.equal(res['gt_max_overlaps'].cpu().view(torch.int32),
                             full['gt_max_overlaps'].cpu().view(torch.int32)))
        ok = ok and same
        report['pairwise_assign'] = {'anchors': na, 'gts': m, 'equal_to_single_gpu': same,
                                     'positives': int((full['assigned_gt_inds'] > 0).sum()),
                                     'ms_per_call_max_over_ranks': round(float(ms.item()), 4)}
    if rank == 0:
        report['ok'] = ok
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
