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


# ---------------------------------------------------------------------------
# pairwise path: anchors sharded, per-GT minima merged with one all-reduce(MIN) of keys
# ---------------------------------------------------------------------------
def test_min_key_packing_round_trip_and_order():
    v = torch.tensor([0.0, -0.0, 1.5, -2.0, float('inf'), float('-inf'), float('nan'), 3e-40,
                      1e38, 1.5, 1.5], dtype=torch.float32)
    i = torch.tensor([5, 3, 9, 2 ** 31 - 2, 7, 8, 9, -1, 4, 2, 11])
    k = sharded.pack_min_keys(v, i)
    vv, ii = sharded.unpack_min_keys(k)
    assert torch.equal(ii, i)
    assert torch.equal(torch.isnan(vv), torch.isnan(v))
    ok = ~torch.isnan(v)
    assert torch.equal(vv[ok].view(torch.int32), v[ok].view(torch.int32))     # bit exact
    order = torch.argsort(k).tolist()
    assert order[0] == 6                                   # NaN first (torch.min propagates it)
    assert order[1] == 5 and order[-1] == 4                # -inf ... +inf
    assert order.index(9) < order.index(2) < order.index(10)   # equal values: lowest index first


def _first_argmin(mat, dim):
    key = torch.where(torch.isnan(mat), torch.full_like(mat, -float('inf')), mat)
    mn = key.min(dim=dim, keepdim=True).values
    idx = (key == mn).to(torch.uint8).argmax(dim=dim)
    return torch.gather(mat, dim, idx.unsqueeze(dim)).squeeze(dim), idx


def _assign_from_minima_ref(thr):
    pos, neg_lo, neg_hi, min_pos, lowq = thr

    def fn(row_min, row_arg, col_min, col_arg_local):
        sim = 1.0 - row_min
        lab = torch.full_like(row_arg, -1)
        lab[(sim >= neg_lo) & (sim < neg_hi)] = 0
        hit = sim >= pos
        lab[hit] = row_arg[hit] + 1
        if lowq:
            for j in range(col_min.shape[0]):              # ascending: the highest GT wins
                a = int(col_arg_local[j])
                if float(1.0 - col_min[j]) >= min_pos and a >= 0:
                    lab[a] = j + 1
        return lab, sim
    return fn


class _Thresholds:
    def __init__(self, thr):
        self.pos_iou_thr, self.neg_lo, self.neg_hi, self.min_pos_iou, self.match_low_quality = thr
        self.cfg = None


def _pairwise_local(b1, b2):
    if b1.shape[0] == 0:
        m = b2.shape[0]
        return (torch.empty(0), torch.empty(0, dtype=torch.int64),
                torch.full((m,), float('inf')), torch.full((m,), -1, dtype=torch.int64))
    mat = gd_oracle.pairwise_distance(b1.double(), b2.double(), 'gwd3d', fun='log1p',
                                      tau=1.0).float()
    rv, ri = _first_argmin(mat, 1)
    cv, ci = _first_argmin(mat, 0)
    return rv, ri, cv, ci


PAIR_THR = ((0.6, 0.0, 0.45, 0.45, True), (0.5, 0.1, 0.3, 0.2, True), (0.6, 0.0, 0.45, 0.0, False))


def _pairwise_inputs(n, m):
    b1 = synth.make_anchor_grid(n, 'waymo')
    b2 = synth.make_targets(m, 'waymo', seed=4)
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    b2[m - 1] = b2[1]                 # duplicate GT: exact ties between columns
    b1[n - 1] = b1[3]                 # duplicate anchors in DIFFERENT shards: ties across ranks
    b1[5] = b1[3]
    return b1, b2


def _pair_worker(rank, world, port, n, m, results):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        b1, b2 = _pairwise_inputs(n, m)
        lo, hi = sharded.shard_bounds(n, rank, world)
        out = {'lo': lo, 'hi': hi}
        for thr in PAIR_THR:
            mod = sharded.ShardedGDMaxSimAssigner(_Thresholds(thr), pairwise_fn=_pairwise_local,
                                                  assign_fn=_assign_from_minima_ref(thr))
            res = mod.assign(b1[lo:hi], b2, lo)
            out[thr] = {k: v.clone() for k, v in res.items()}
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n,world', [(1003, 2), (6, 2)])
def test_sharded_pairwise_assign_world_size_2(n, world):
    """Labels, per-GT maxima and arg-maxima of the row-sharded assigner equal the
    unsharded MaxIoUAssigner restatement; (6, 2): the second shard holds 2 rows only."""
    m = 9
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_pair_worker, args=(world, port, n, m, results), nprocs=world, join=True)
    b1, b2 = _pairwise_inputs(n, m)
    mat = gd_oracle.pairwise_distance(b1.double(), b2.double(), 'gwd3d', fun='log1p',
                                      tau=1.0).float()
    sim = 1.0 - mat
    cv, ci = _first_argmin(mat, 0)
    for thr in PAIR_THR:
        pos, neg_lo, neg_hi, min_pos, lowq = thr
        want, want_max = gd_oracle.max_sim_assign(sim, pos, (neg_lo, neg_hi), min_pos, lowq)
        got = torch.cat([results[r][thr]['assigned_gt_inds'] for r in range(world)])
        got_max = torch.cat([results[r][thr]['max_overlaps'] for r in range(world)])
        assert torch.equal(got, want), thr
        assert torch.equal(got_max, want_max)
        for r in range(world):                              # merged minima: same on every rank
            assert torch.equal(results[r][thr]['gt_argmax_overlaps'], ci)
            assert torch.equal(results[r][thr]['gt_max_overlaps'], 1.0 - cv)
