"""CPU tests of the multi-GPU host logic under gloo, world_size 2.

The CUDA kernel cannot run here, so the LOCAL loss is the oracle module injected into
``ShardedGDLoss`` (tests may use the oracle); what is under test is the partition
(``shard_bounds``), the single all-reduce, the global-mean handling and the local
gradient semantics -- the same code path the NCCL run takes on GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmdet3d_gaussian_b200 import sharded, synth
from oracle import gd_oracle


def test_shard_bounds_partition():
    for n in (0, 1, 3, 4, 5, 255, 256, 1000, 786432, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = sharded.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n and lo % 4 == 0 or lo == n
                cover.append((lo, hi))
            assert cover[0][0] == 0 and cover[-1][1] == n
            for (a, b), (c, d) in zip(cover, cover[1:]):
                assert b == c
            sizes = [b - a for a, b in cover]
            assert max(sizes) - min(s for s in sizes if s or True) <= -(-n // world) + 4
    with pytest.raises(ValueError):
        sharded.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pred, target, w = synth.make_pairs(n, 'nuscenes', seed=3, weights='bernoulli')
        pred, target, w = pred.double(), target.double(), w.double()
        lo, hi = sharded.shard_bounds(n, rank, world)
        out = {}
        for red, af in (('mean', None), ('mean', 123.0), ('sum', None)):
            local = gd_oracle.GDLossOracle('gwd3d', fun='log1p', tau=0.0, loss_weight=5.0,
                                           reduction=red)
            mod = sharded.ShardedGDLoss(local)
            p = pred[lo:hi].clone().requires_grad_(True)
            loss = mod(p, target[lo:hi], w[lo:hi], avg_factor=af)
            loss.backward()
            out[(red, af)] = (loss.item(), p.grad.clone(), lo, hi)
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_sharded_loss_world_size_2():
    n, world = 1003, 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, results), nprocs=world, join=True)
    pred, target, w = synth.make_pairs(n, 'nuscenes', seed=3, weights='bernoulli')
    pred, target, w = pred.double(), target.double(), w.double()
    for red, af in (('mean', None), ('mean', 123.0), ('sum', None)):
        full = gd_oracle.GDLossOracle('gwd3d', fun='log1p', tau=0.0, loss_weight=5.0,
                                      reduction=red)
        ref_l, ref_g = gd_oracle.loss_and_grad(full, pred, target, w, avg_factor=af)
        grads = torch.zeros_like(ref_g)
        for rank in range(world):
            loss, g, lo, hi = results[rank][(red, af)]
            assert abs(loss - ref_l.item()) <= 1e-12 * abs(ref_l.item())   # same on every rank
            grads[lo:hi] = g
        assert torch.allclose(grads, ref_g, rtol=1e-12, atol=1e-15)


def test_sharded_single_process_passthrough():
    """No process group: the wrapper is the plain module (rows='none' is always local)."""
    pred, target, w = synth.make_pairs(64, 'kitti', seed=1)
    local = gd_oracle.GDLossOracle('kld3d', reduction='mean')
    mod = sharded.ShardedGDLoss(local)
    a = mod(pred.double(), target.double(), w.double())
    b = local(pred.double(), target.double(), w.double())
    assert abs(a.item() - b.item()) < 1e-14
    rows = mod(pred.double(), target.double(), reduction_override='none')
    assert rows.shape == (64,)
