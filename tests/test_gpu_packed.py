"""GPU parity of the packed-FP32 variant of the fused kernel (``variant='bulk_packed'`` =
GD_VARIANT_BULK_PACKED, csrc/gd_packed.cuh -- what ``variant='auto'`` runs for gwd3d / kld3d /
bd3d since round 2, build_ext.TUNE_DEFAULT) against the fp64 oracle and the scalar kernel
(``variant='bulk'``), of the two column mappings of the pairwise kernel, and of the ``host_sync``
aliases."""
import math

import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import GDLoss, synth
from oracle import gd_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _run(kw, variant, pred, target, w, **fw):
    p = pred.clone().cuda().requires_grad_(True)
    out = GDLoss(variant=variant, host_sync=False, **kw)(p, target.cuda(), w.cuda(), **fw)
    out.backward()
    return out.item(), p.grad.cpu().double().numpy()


@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d'])
@pytest.mark.parametrize('fun,tau', [('log1p', 0.0), ('none', 0.0), ('log1p', 1.0), ('none', 2.0)])
@pytest.mark.parametrize('n', [100_003, 4, 131, 1 << 20])
def test_packed_vs_oracle_and_scalar(loss_type, fun, tau, n):
    pred, target, w = synth.make_pairs(n, 'kitti', seed=11, weights='bernoulli')
    # rows the FAST path must hand to the robust path, in both halves of a pair
    if n > 64:
        pred[5, 4] = 1e-9
        pred[8, 6] = 1000.0
        target[9, 3] = 2e4
    kw = dict(loss_type=loss_type, fun=fun, tau=tau, loss_weight=5.0)
    lp, gp = _run(kw, 'bulk_packed', pred, target, w, avg_factor=float(n))
    ls, gs = _run(kw, 'bulk', pred, target, w, avg_factor=float(n))
    ref_l, ref_g = gd_oracle.loss_and_grad(gd_oracle.GDLossOracle(**kw), pred.double(),
                                           target.double(), w.double(), avg_factor=float(n))
    ref_g = ref_g.numpy()
    assert abs(lp - ref_l.item()) <= RTOL * abs(ref_l.item())
    assert abs(lp - ls) <= 2e-6 * abs(ls)
    fin = np.isfinite(ref_g).all(axis=1)
    gn = np.maximum(np.linalg.norm(ref_g[fin], axis=1), 1e-2 * 5.0 / n)
    assert (np.linalg.norm(gp[fin] - ref_g[fin], axis=1) / gn).max() <= RTOL
    assert (np.linalg.norm(gp[fin] - gs[fin], axis=1) / gn).max() <= 2e-6


def test_packed_falls_back_for_other_configurations():
    """Losses / modes without a packed instantiation run the scalar bulk kernel."""
    pred, target, w = synth.make_pairs(5000, 'kitti', seed=3)
    for kw in (dict(loss_type='jd3d'), dict(loss_type='kfiou3d', fun='none'),
               dict(loss_type='gwd3d', normalize=False), dict(loss_type='kld3d', reduction='none')):
        a = GDLoss(variant='bulk_packed', **kw)(pred.cuda(), target.cuda(), w.cuda())
        b = GDLoss(variant='bulk', **kw)(pred.cuda(), target.cuda(), w.cuda())
        assert torch.equal(a, b)


@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d'])
@pytest.mark.parametrize('fun,tau', [('log1p', 1.0), ('none', 0.0)])
@pytest.mark.parametrize('n,m', [(1, 1), (63, 5), (65, 33), (700, 64), (1001, 129), (2050, 256), (400, 700),
                                 (60001, 40)])
def test_packed_pairwise(loss_type, fun, tau, n, m):
    """Pairwise kernel: the matrix vs the fp64 oracle (1e-5); the fused minima equal the minima
    of the matrix written by the SAME launch bit for bit (ties -> lowest index, NaN first), odd
    row counts and ragged column counts included; degenerate boxes take the robust path; the
    matrix-only launch (two columns per lane for gwd3d / kld3d), the one-column mapping and the
    row-lane reductions all return the same bits."""
    from mmdet3d_gaussian_b200 import GDPairwiseDistance
    b1 = synth.make_anchor_grid(n, 'waymo', device='cuda')
    b2 = synth.make_targets(m, 'waymo', seed=n + m, device='cuda')
    b2[:, 0] = b2[:, 0] * 2.0 - 70.0
    if m > 4:
        b2[m - 1] = b2[1]                       # exact ties between columns
    if n > 70:
        b1[n - 1] = b1[3]                       # exact ties between rows (odd and even)
        b1[68] = b1[3]
        b1[10, 4] = 1e-9                        # rows the FAST cores must hand over
        b1[11, 6] = 1000.0
    mod = GDPairwiseDistance(loss_type, fun=fun, tau=tau)
    plain = mod(b1, b2)
    rmin, ridx, cmin, cidx, mat = mod.assign(b1, b2, want_matrix=True)
    r1, i1, c1, j1, mat1 = mod.assign(b1, b2, want_matrix=True)         # repeatable
    ref = gd_oracle.pairwise_distance(b1.cpu().double(), b2.cpu().double(), loss_type, fun=fun,
                                      tau=tau)
    err = (mat.cpu().double() - ref).abs() / ref.abs().clamp_min(1e-3)
    assert err.max() < RTOL

    def same(a, b):
        return torch.equal(a.view(torch.int32), b.view(torch.int32))
    # matrix and minima of ONE launch come from one instruction sequence
    assert same(mat, mat1)
    assert same(rmin, r1) and torch.equal(ridx, i1) and same(cmin, c1) and torch.equal(cidx, j1)
    # the matrix-only launch runs the same mapping: same bits
    assert same(plain, mat)
    # the other column mapping is another instantiation -- since the pairwise values are
    # computed with explicitly rounded operations (gd_math.cuh, namespace pw) it returns the
    # same bits, and so do its minima (round 2: last-place differences, 2e-6)
    r1c, i1c, c1c, j1c, m1 = mod.assign(b1, b2, want_matrix=True, cpl1=True)
    assert same(m1, mat)
    assert same(r1c, rmin) and torch.equal(i1c, ridx) and same(c1c, cmin) and torch.equal(j1c, cidx)

    def first_argmin(x, dim):
        key = torch.where(torch.isnan(x), torch.full_like(x, -float('inf')), x)
        mn = key.min(dim=dim, keepdim=True).values
        idx = (key == mn).to(torch.uint8).argmax(dim=dim)
        return torch.gather(x, dim, idx.unsqueeze(dim)).squeeze(dim), idx
    for _ in range(2):                          # the workspace must come back clean
        rv, ri = first_argmin(mat, 1)
        cv, ci = first_argmin(mat, 0)
        assert torch.equal(rmin.view(torch.int32), rv.view(torch.int32)) and torch.equal(ridx, ri)
        assert torch.equal(cmin.view(torch.int32), cv.view(torch.int32)) and torch.equal(cidx, ci)
        rmin, ridx, cmin, cidx, none = mod.assign(b1, b2)
        assert none is None


@pytest.mark.parametrize('aspect', [5.0, 30.0, 100.0])
def test_pairwise_short_forms_on_elongated_crossing_boxes(aspect):
    """The pairwise value path takes sin/cos of the yaw difference from per-box sines / cosines
    and a short log1p (csrc/gd_math.cuh, P.lean): elongated boxes at right angles, distances from
    1e-3 to 1e2, aspect ratios up to 100:1 -- still within 1e-5 of the fp64 oracle."""
    from mmdet3d_gaussian_b200 import GDPairwiseDistance
    g = torch.Generator().manual_seed(int(aspect))
    n, m = 600, 96
    b1 = synth.make_targets(n, 'waymo', seed=1)
    b2 = synth.make_targets(m, 'waymo', seed=2)
    b1[:, 3] = b1[:, 4] * aspect
    b2[:, 3] = b2[:, 4] * aspect
    b2[:, :3] = b1[:m, :3] + torch.randn(m, 3, generator=g) * torch.logspace(-3, 1.5, m)[:, None]
    b1[:m, 6] = b2[:, 6] + math.pi / 2                         # exactly crossing pairs on the diagonal
    b1[m:2 * m, 6] = b2[:, 6] + math.pi / 2 + 1e-3
    for lt in ('gwd3d', 'kld3d', 'bd3d'):
        for fun, tau in (('log1p', 1.0), ('log1p', 0.0), ('none', 0.0)):
            mat = GDPairwiseDistance(lt, fun=fun, tau=tau)(b1.cuda(), b2.cuda()).cpu().double()
            ref = gd_oracle.pairwise_distance(b1.double(), b2.double(), lt, fun=fun, tau=tau)
            err = ((mat - ref).abs() / ref.abs().clamp_min(1e-3)).max().item()
            assert err < RTOL, (lt, fun, tau, aspect, err)


def test_overlapped_host_sync_mode():
    """``host_sync='overlap'`` (opt-in): the decisions and results of the default module
    (reference gaussian_distance_loss.py:290-292), with the fused launch queued before the
    host waits for the probe."""
    pred, target, w = synth.make_pairs(20_000, 'kitti', seed=4, weights='bernoulli')
    for kw in (dict(loss_type='kld3d', fun='log1p', tau=0.0), dict(loss_type='gwd3d')):
        outs = []
        for hs in (True, 'overlap'):
            p = pred.clone().cuda().requires_grad_(True)
            loss = GDLoss(host_sync=hs, **kw)(p, target.cuda(), w.cuda(), avg_factor=123.0)
            loss.backward()
            outs.append((loss.item(), p.grad.clone()))
        assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    # early return: no positive weight -> (pred * weight).sum(), gradient = weight
    w7 = -torch.rand(64, 7)
    p = pred[:64].clone().cuda().requires_grad_(True)
    out = GDLoss('kld3d', host_sync='overlap')(p, target[:64].cuda(), w7.cuda())
    out.backward()
    assert torch.allclose(out.cpu(), (pred[:64] * w7).sum()) and torch.equal(p.grad.cpu(), w7)
    # reduction='none' and weight=None never probe
    rows = GDLoss('bd3d', host_sync='overlap', reduction='none')(pred[:50].cuda(), target[:50].cuda(),
                                                                  w[:50].cuda())
    assert rows.shape == (50,)
