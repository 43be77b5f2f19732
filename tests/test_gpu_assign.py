"""GPU tests of the pairwise kernel with the assigner reductions fused
(``pytest -m gpu``).

Bit-exactness rule (BASELINE.json north_star: "assignment indices derived from the
pairwise matrix bit-exact"): the fused row / column arg-minima must equal the
arg-minima of the matrix the same kernel materialises -- value bits identical, ties to
the lowest index, NaN first -- for every shape, and the labels of
``GDMaxSimAssigner`` must equal the oracle's MaxIoUAssigner restatement run on that
matrix.  Against the fp64 oracle matrix, values agree to 1e-5 and indices agree on
every row / column whose two best candidates are separated by more than 1e-5
relative (the tie-margin audit)."""
import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import (GDMaxSimAssigner, GDPairwiseDistance, GDSimilarity3D, ops,
                                   synth)
from mmdet3d_gaussian_b200 import _lib
from oracle import gd_oracle

pytestmark = pytest.mark.gpu
ALL_TYPES = ('gwd3d', 'kld3d', 'jd3d', 'kld3d_symmax', 'kld3d_symmin', 'bd3d', 'kfiou3d')


@pytest.fixture(scope='module', autouse=True)
def _need_cuda():
    assert torch.cuda.is_available()
    _lib.load()
    yield


def first_argmin(mat, dim):
    """(min, lowest index attaining it); NaN counts as the minimum (torch.min)."""
    key = torch.where(torch.isnan(mat), torch.full_like(mat, -float('inf')), mat)
    mn = key.min(dim=dim, keepdim=True).values
    hit = (key == mn)
    idx = hit.to(torch.uint8).argmax(dim=dim)
    val = torch.gather(mat, dim, idx.unsqueeze(dim)).squeeze(dim)
    return val, idx


def same_bits(a, b):
    return torch.equal(a.view(torch.int32), b.view(torch.int32))


@pytest.mark.parametrize('loss_type', ALL_TYPES)
@pytest.mark.parametrize('n,m', [(1, 1), (63, 5), (64, 16), (65, 32), (700, 33), (513, 64),
                                 (300, 100), (1000, 129), (2050, 256), (900, 300), (400, 700)])
def test_fused_minima_equal_matrix_minima(loss_type, n, m):
    fun = 'none' if loss_type == 'kfiou3d' else 'log1p'
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=n + m, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    if m > 4:                                   # duplicate GTs and anchors: exact ties
        b2[m - 1] = b2[1]
    if n > 70:
        b1[n - 1] = b1[3]
        b1[69] = b1[3]
    mod = GDPairwiseDistance(loss_type, fun=fun, tau=1.0)
    mat = mod(b1, b2)
    for _ in range(2):                          # twice: the workspace must come back clean
        rmin, ridx, cmin, cidx, none = mod.assign(b1, b2)
        assert none is None
        rv, ri = first_argmin(mat, 1)
        cv, ci = first_argmin(mat, 0)
        assert same_bits(rmin, rv) and torch.equal(ridx, ri), (loss_type, n, m, 'rows')
        assert same_bits(cmin, cv) and torch.equal(cidx, ci), (loss_type, n, m, 'cols')
    # matrix written by the fused kernel == the plain matrix kernel; similarity = 1 - D
    rmin2, ridx2, cmin2, cidx2, mat2 = mod.assign(b1, b2, want_matrix=True)
    assert same_bits(mat2, mat) and torch.equal(ridx2, ridx) and torch.equal(cidx2, cidx)
    sim = ops.pairwise_assign(b1, b2, mod.cfg, want_matrix=True, similarity=True)[4]
    assert same_bits(sim, 1.0 - mat)
    # legacy entry point
    v0, i0 = mod.row_argmin(b1, b2)
    assert same_bits(v0, rmin) and torch.equal(i0, ridx)
    # int32 index outputs (the C ABI's default) == the int64 ones the kernel writes for torch
    r32 = ops.pairwise_assign(b1, b2, mod.cfg, index64=False)
    assert r32[1].dtype == torch.int32 and ridx.dtype == torch.int64
    assert torch.equal(r32[1].long(), ridx) and torch.equal(r32[3].long(), cidx) and same_bits(r32[0], rmin)


def test_nan_and_degenerate_boxes_in_reductions():
    b1 = synth.make_anchor_grid(200, 'waymo', device='cuda')
    b2 = synth.make_targets(40, 'waymo', seed=3, device='cuda')
    b1[7, 0] = float('nan')
    b2[11, 3] = float('nan')
    b1[20, 3:6] = 1e-7
    b2[5, 4] = -1.0
    mod = GDPairwiseDistance('gwd3d', fun='log1p', tau=1.0)
    mat = mod(b1, b2)
    rmin, ridx, cmin, cidx, _ = mod.assign(b1, b2)
    rv, ri = first_argmin(mat, 1)
    cv, ci = first_argmin(mat, 0)
    assert torch.equal(torch.isnan(rmin), torch.isnan(rv)) and torch.isnan(rmin).all()
    assert torch.equal(ridx, ri) and torch.equal(cidx, ci)
    assert torch.isnan(cmin).all() and int(ridx[0]) == 11 and int(cidx[0]) == 7


@pytest.mark.parametrize('loss_type', ('gwd3d', 'kld3d', 'bd3d'))
def test_minima_vs_fp64_oracle_with_tie_audit(loss_type):
    b1 = synth.make_anchor_grid(3000, 'waymo')
    b2 = synth.make_targets(48, 'waymo', seed=9)
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    ref = gd_oracle.pairwise_distance(b1.double(), b2.double(), loss_type, fun='log1p', tau=1.0)
    mod = GDPairwiseDistance(loss_type, fun='log1p', tau=1.0)
    rmin, ridx, cmin, cidx, _ = mod.assign(b1.cuda(), b2.cuda())
    for ours_v, ours_i, dim in ((rmin, ridx, 1), (cmin, cidx, 0)):
        two = torch.topk(ref, 2, dim=dim, largest=False).values
        best, second = (two[:, 0], two[:, 1]) if dim == 1 else (two[0], two[1])
        clear = (second - best) > 1e-5 * second.abs()
        assert clear.float().mean() > 0.9
        assert torch.equal(ours_i.cpu()[clear], ref.argmin(dim=dim)[clear])
        err = (ours_v.cpu().double() - best).abs() / best.abs().clamp_min(1e-3)
        assert err.max() < 1e-5


@pytest.mark.parametrize('m', [1, 7, 40, 256])
def test_max_sim_assigner_equals_oracle_on_our_matrix(m):
    n = 20_000
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=m, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    kw = dict(loss_type='gwd3d', fun='log1p', tau=1.0)
    sim = GDSimilarity3D(**kw)(b1, b2)
    assert sim.shape == (n, m) and float(sim.max()) <= 1.0 and float(sim.min()) >= 0.0
    for pos, neg, min_pos, lowq in ((0.6, 0.45, 0.45, True), (0.5, (0.1, 0.3), 0.2, True),
                                    (0.6, 0.45, 0.0, False)):
        res = GDMaxSimAssigner(pos, neg, min_pos, lowq, **kw).assign(b1, b2)
        want, want_max = gd_oracle.max_sim_assign(sim.cpu(), pos, neg, min_pos, lowq)
        assert torch.equal(res['assigned_gt_inds'].cpu(), want), (m, pos, neg)
        assert same_bits(res['max_overlaps'].cpu(), want_max)
        assert same_bits(res['gt_max_overlaps'].cpu(), sim.cpu().max(0).values)
        assert (res['assigned_gt_inds'] > 0).any() or not lowq
    # no ground truth: everything is background
    res = GDMaxSimAssigner(0.6, 0.45, **kw).assign(b1, b2[:0])
    assert int(res['assigned_gt_inds'].abs().sum()) == 0


def test_c4_shape_consistency():
    """BASELINE configs[3]: 200k anchors x 256 GTs -- fused minima == matrix minima."""
    b1 = synth.make_anchor_grid(200_000, 'waymo', device='cuda')
    b2 = synth.make_targets(256, 'waymo', seed=5, device='cuda')
    for lt in ('gwd3d', 'kld3d', 'bd3d'):
        mod = GDPairwiseDistance(lt, fun='log1p', tau=1.0)
        mat = mod(b1, b2)
        rmin, ridx, cmin, cidx, _ = mod.assign(b1, b2)
        rv, ri = first_argmin(mat, 1)
        cv, ci = first_argmin(mat, 0)
        assert same_bits(rmin, rv) and torch.equal(ridx, ri)
        assert same_bits(cmin, cv) and torch.equal(cidx, ci)


def test_sharded_assigner_emulated_shards():
    """Row-sharded assignment (pairwise): per-shard fused minima, the
    64-bit (value, anchor) keys of ``sharded.pack_min_keys`` merged with an element-wise MIN
    (what the all-reduce does across ranks) and ``gd_assign_from_minima`` fed with shard-local
    arg-minima (-1 = the winner lives in another shard) reproduce the single-GPU labels."""
    from mmdet3d_gaussian_b200 import sharded
    n, m, world = 5003, 40, 3
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=2, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    b1[n - 1] = b1[3]                              # the same anchor in two shards: a cross-shard tie
    kw = dict(loss_type='gwd3d', fun='log1p', tau=1.0)
    base = GDMaxSimAssigner(0.6, 0.45, 0.45, True, **kw)
    full = base.assign(b1, b2)
    keys, parts = None, []
    for r in range(world):
        lo, hi = sharded.shard_bounds(n, r, world)
        rmin, rarg, cmin, carg, _ = ops.pairwise_assign(b1[lo:hi], b2, base.cfg)
        k = sharded.pack_min_keys(cmin, torch.where(carg >= 0, carg + lo, carg))
        keys = k if keys is None else torch.minimum(keys, k)
        parts.append((rmin, rarg, lo, hi))
    gmin, garg = sharded.unpack_min_keys(keys)
    assert torch.equal(garg, full['gt_argmax_overlaps'])
    assert same_bits(1.0 - gmin, full['gt_max_overlaps'])
    labels = []
    for rmin, rarg, lo, hi in parts:
        local = torch.where((garg >= lo) & (garg < hi), garg - lo, torch.full_like(garg, -1))
        assigned, _ = ops.assign_from_minima(rmin, rarg, gmin, local, 0.6, 0.0, 0.45, 0.45, True)
        labels.append(assigned)
    assert torch.equal(torch.cat(labels), full['assigned_gt_inds'])
    # the module wrapper without a process group is the single-GPU assigner
    res = sharded.ShardedGDMaxSimAssigner(base).assign(b1, b2, 0)
    assert torch.equal(res['assigned_gt_inds'], full['assigned_gt_inds'])
    assert torch.equal(res['gt_argmax_overlaps'], full['gt_argmax_overlaps'])


# ---------------------------------------------------------------------------
# SimOTA-style consumer: column top-k + dynamic-k matching without the matrix
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d', 'jd3d'])
@pytest.mark.parametrize('n,m,topk', [(5000, 37, 10), (700, 300, 10), (64, 5, 10), (9, 3, 10),
                                      (20001, 256, 13), (300, 1, 4)])
def test_simota_assigner_equals_matching_on_the_matrix(loss_type, n, m, topk):
    """The matrix-free path (row minima from the row-lane kernel; column top-k lists from the
    threshold sample + filter pass + per-column selection) against the restated reference
    matching (gd_oracle.simota_dynamic_k_matching = sim_ota_3d_assigner.py:184-211) run on the
    fp32 matrix: identical assignment, similarities and dynamic k -- exact ties included (both
    take the lowest index first); the lists equal a stable sort of the matrix columns bit for
    bit; the result does not depend on whether the matrix is written; and the matrix IS the plain
    pairwise kernel's (every launch evaluates a pair with the same explicitly rounded
    operations)."""
    from mmdet3d_gaussian_b200 import GDSimOTAAssigner
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=n + m, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    b2[:, :3] = b1[torch.randint(0, n, (m,), device='cuda'), :3] + 0.3 * torch.randn(m, 3, device='cuda')
    if n > 70:
        b1[n - 1] = b1[3]                         # exact ties between rows
        b1[68] = b1[3]
    asg = GDSimOTAAssigner(candidate_topk=topk, loss_type=loss_type, fun='log1p', tau=1.0)
    res = asg.assign(b1, b2)
    resm = asg.assign(b1, b2, want_matrix=True)
    for key in ('assigned_gt_inds', 'max_overlaps', 'dynamic_ks', 'topk_overlaps', 'topk_inds'):
        assert torch.equal(res[key], resm[key]), key
    mat = resm['distance_matrix']
    plain = GDPairwiseDistance(loss_type, fun='log1p', tau=1.0)(b1, b2)
    assert same_bits(plain, mat)
    k = min(topk, n)
    order = torch.sort(mat, dim=0, stable=True).indices[:k]           # ties -> lowest row
    assert torch.equal(res['topk_inds'], order)
    want_val = torch.gather(mat, 0, order)
    assert torch.equal((1.0 - want_val).view(torch.int32), res['topk_overlaps'].view(torch.int32))
    sims = (1.0 - mat).cpu()
    assigned, matched, dks = gd_oracle.simota_dynamic_k_matching(mat.cpu(), sims, topk)
    assert torch.equal(res['dynamic_ks'].cpu(), dks)
    assert torch.equal(res['assigned_gt_inds'].cpu(), assigned)
    fg = assigned > 0
    assert torch.equal(res['max_overlaps'].cpu()[fg], matched[fg])
    assert bool((res['max_overlaps'].cpu()[~fg] == -float(asg.INF)).all())
    assert int(fg.sum()) >= 1
    # a second call on the same stream (scratch reuse)
    res2 = asg.assign(b1, b2)
    assert torch.equal(res2['assigned_gt_inds'], res['assigned_gt_inds'])


@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d'])
def test_column_topk_at_c4_size_and_with_overflowing_candidate_buffers(loss_type):
    """gd_pairwise_col_topk at BASELINE's C4 shape (200k x 256): lists == a stable sort of the
    matrix columns, bit for bit.  Then an input built to defeat the threshold sample: every row
    the strided sample sees is far away, all the others are close, so nearly every pair passes
    the filter and every column's candidate buffer overflows -- the brute-force path must give
    the same exact answer."""
    n, m, k = 200_000, 256, 10
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=5, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    cfg = GDPairwiseDistance(loss_type, fun='log1p', tau=1.0).cfg

    def check(b1, b2, k):
        rmin, ridx, val, row, _ = ops.pairwise_col_topk(b1, b2, cfg, k)
        mat = ops.pairwise_distance(b1, b2, cfg)
        order = torch.sort(mat, dim=0, stable=True).indices[:k]
        assert torch.equal(row, order)
        assert same_bits(val, torch.gather(mat, 0, order))
        rv, ri = first_argmin(mat, 1)
        assert same_bits(rmin, rv) and torch.equal(ridx, ri)
    check(b1, b2, k)
    n2 = 40_960                                  # sample step 20: rows 0, 20, 40, ... are sampled
    c1 = synth.make_anchor_grid(n2, 'waymo', device='cuda')
    c1[::20, :2] += 500.0                        # the sampled rows sit far from every GT
    check(c1, b2[:40], 16)


def test_simota_assigner_vs_fp64_oracle_and_edge_cases():
    """Against the fp64 oracle matrix: equal assignment wherever the decisions have a margin
    (top-k membership and the conflict argmin separated by more than 1e-5 relative)."""
    from mmdet3d_gaussian_b200 import GDSimOTAAssigner
    n, m, topk = 3000, 24, 10
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=77, device='cuda')
    b2[:, :3] = b1[torch.randint(0, n, (m,), device='cuda'), :3] + 0.2 * torch.randn(m, 3, device='cuda')
    asg = GDSimOTAAssigner(candidate_topk=topk, loss_type='gwd3d', fun='log1p', tau=1.0)
    res = asg.assign(b1, b2)
    ref = gd_oracle.pairwise_distance(b1.cpu().double(), b2.cpu().double(), 'gwd3d', fun='log1p', tau=1.0)
    assigned, matched, dks = gd_oracle.simota_dynamic_k_matching(ref, 1.0 - ref, topk)
    # decisions with a margin: the (k)th and (k+1)th smallest of every column, k = dynamic k
    srt = torch.sort(ref, dim=0).values
    clear = torch.ones(m, dtype=torch.bool)
    for j in range(m):
        kk = int(dks[j])
        clear[j] = (srt[kk, j] - srt[kk - 1, j]) > 1e-5 * srt[kk, j].abs().clamp_min(1e-3)
        s = (1.0 - srt[:topk, j]).sum()
        clear[j] &= abs(s - torch.round(s)) > 1e-4        # int() truncation away from an integer
    if bool(clear.all()):
        assert torch.equal(res['dynamic_ks'].cpu(), dks)
        top2 = ref.topk(2, dim=1, largest=False).values
        sure = (top2[:, 1] - top2[:, 0]) > 1e-5 * top2[:, 1].abs().clamp_min(1e-3)
        assert torch.equal(res['assigned_gt_inds'].cpu()[sure], assigned[sure])
    # no GT / no boxes
    e = asg.assign(b1, b2[:0])
    assert int(e['assigned_gt_inds'].abs().sum()) == 0 and e['assigned_gt_inds'].shape == (n,)
    e = asg.assign(b1[:0], b2)
    assert e['assigned_gt_inds'].shape == (0,)
    with pytest.raises(ValueError):
        GDSimOTAAssigner(candidate_topk=17)
    with pytest.raises(ValueError):
        GDSimOTAAssigner(tau=0.0)
