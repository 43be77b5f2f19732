"""GPU parity tests of the fused head front ends (``pytest -m gpu``), SURVEY.md
section 8 row f1: positive-row gather + box decode + GD loss + gradient w.r.t. the raw
head outputs in one launch, against the fp64 oracle's restatement of the two
reference call sites (``gd_anchor3d_head.py:102-141``,
``gd_centerpoint_head.py:413-434``) under autograd, and against the golden
vectors written from the unmodified reference classes
(``tests/golden/gd_decode_golden.npz``).  Tolerances as in test_gpu_parity.py.
"""
import json
import os

import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import GDAnchorHeadLoss, GDCenterHeadLoss, GDLoss, ops, synth
from mmdet3d_gaussian_b200 import _lib
from oracle import gd_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = [dict(loss_type='gwd3d', fun='log1p', tau=1.0, loss_weight=5.0),      # KITTI gwd5tau1
           dict(loss_type='kld3d', fun='log1p', tau=1.0, loss_weight=5.0),
           dict(loss_type='bd3d', fun='none', tau=0.0, loss_weight=5.0),
           dict(loss_type='jd3d', fun='log1p', tau=1.0, loss_weight=5.0),
           dict(loss_type='kld3d_symmin', fun='log1p', tau=1.0, loss_weight=5.0, sqrt=False),
           dict(loss_type='kfiou3d', fun='none', tau=0.0, loss_weight=5.0)]


@pytest.fixture(scope='module', autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), 'these tests need a CUDA device'
    _lib.load()
    yield


def cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


def check(loss, grad, ref_loss, ref_grad, what):
    assert abs(loss - ref_loss) <= RTOL * abs(ref_loss) + 1e-12, (what, loss, ref_loss)
    fin = np.isfinite(ref_grad).all(1)
    err = np.linalg.norm((grad - ref_grad)[fin], axis=1)
    tol = RTOL * np.maximum(np.linalg.norm(ref_grad[fin], axis=1),
                            1e-3 * np.abs(ref_grad[fin]).max())
    assert (err <= tol).all(), (what, float((err / tol).max()))


def oracle_anchor(kw, b, decode_weight, avg):
    bp = b['bbox_pred'].double().requires_grad_(True)
    loss = gd_oracle.anchor_head_gd_loss(
        gd_oracle.GDLossOracle(**kw), b['anchors'].double(), bp, b['bbox_targets'].double(),
        b['bbox_weights'].double(), b['pos_inds'], decode_weight=decode_weight, avg_factor=avg)
    loss.backward()
    return loss.item(), bp.grad.numpy()


@pytest.mark.parametrize('kw', CONFIGS, ids=lambda k: k['loss_type'])
@pytest.mark.parametrize('mode', ['index', 'labels'])
def test_anchor_head_vs_oracle(kw, mode):
    b = synth.make_anchor_head_batch(60_000, 21_384, pos_frac=0.01, seed=3)
    avg = float(len(b['pos_inds']))
    ref_loss, ref_grad = oracle_anchor(kw, b, 1, avg)
    g = cuda(b)
    bp = g['bbox_pred'].clone().requires_grad_(True)
    head = GDAnchorHeadLoss(dict(type='GDLoss', **kw), decode_weight=1)
    sel = dict(pos_inds=g['pos_inds']) if mode == 'index' else \
        dict(labels=g['labels'], num_classes=3)
    loss = head(g['anchors'], bp, g['bbox_targets'], g['bbox_weights'], avg_factor=avg, **sel)
    loss.backward()
    check(loss.item(), bp.grad.cpu().double().numpy(), ref_loss, ref_grad, (kw, mode))
    # rows that are not positive get an exactly zero gradient
    neg = torch.ones(60_000, dtype=torch.bool)
    neg[b['pos_inds']] = False
    assert float(bp.grad.cpu()[neg].abs().sum()) == 0.0


def test_anchor_head_weights_and_edge_cases():
    kw = CONFIGS[0]
    b = synth.make_anchor_head_batch(5_003, 1_200, pos_frac=0.05, seed=8)
    g = cuda(b)
    # 7-vector decode_weight, weights that vary per row, avg_factor unrelated to P
    wts = b['bbox_weights'] * torch.rand(5_003, 7)
    b2 = dict(b, bbox_weights=wts)
    dw = [1.0, 1.0, 0.5, 2.0, 2.0, 0.25, 1.5]
    ref_loss, ref_grad = oracle_anchor(kw, b2, dw, 321.0)
    for mode in ('index', 'labels'):
        bp = g['bbox_pred'].clone().requires_grad_(True)
        sel = dict(pos_inds=g['pos_inds']) if mode == 'index' else \
            dict(labels=g['labels'], num_classes=3)
        loss = GDAnchorHeadLoss(dict(type='GDLoss', **kw), decode_weight=dw)(
            g['anchors'], bp, g['bbox_targets'], wts.cuda(), avg_factor=321.0, **sel)
        (loss * 3.0).backward()                               # upstream grad_output != 1
        check(loss.item(), bp.grad.cpu().double().numpy() / 3.0, ref_loss, ref_grad, mode)
    # decode_weight=None -> weight None (gd_anchor3d_head.py:128-131), mean over P
    ref_loss, ref_grad = oracle_anchor(kw, b, None, None)
    bp = g['bbox_pred'].clone().requires_grad_(True)
    loss = GDAnchorHeadLoss(dict(type='GDLoss', **kw))(
        g['anchors'], bp, g['bbox_targets'], g['bbox_weights'], pos_inds=g['pos_inds'])
    loss.backward()
    check(loss.item(), bp.grad.cpu().double().numpy(), ref_loss, ref_grad, 'no weight')
    # row-strided bbox_pred (a view into a wider tensor) and no_grad forward
    wide = torch.zeros(5_003, 9, device='cuda')
    wide[:, :7] = g['bbox_pred']
    with torch.no_grad():
        l2 = GDAnchorHeadLoss(dict(type='GDLoss', **kw))(
            g['anchors'], wide[:, :7], g['bbox_targets'], None, pos_inds=g['pos_inds'])
    assert abs(l2.item() - ref_loss) <= RTOL * abs(ref_loss)
    # no positives: loss 0, zero gradient (gd_anchor3d_head.py:160-161)
    bp = g['bbox_pred'].clone().requires_grad_(True)
    empty = torch.zeros(0, dtype=torch.long, device='cuda')
    loss = GDAnchorHeadLoss(dict(type='GDLoss', **kw), 1)(
        g['anchors'], bp, g['bbox_targets'], g['bbox_weights'], pos_inds=empty, avg_factor=1.0)
    loss.backward()
    assert loss.item() == 0.0 and float(bp.grad.abs().sum()) == 0.0
    bp = g['bbox_pred'].clone().requires_grad_(True)
    loss = GDAnchorHeadLoss(dict(type='GDLoss', **kw), 1)(
        g['anchors'], bp, g['bbox_targets'], g['bbox_weights'],
        labels=torch.full((5_003,), 3, device='cuda'), num_classes=3, avg_factor=1.0)
    loss.backward()
    assert loss.item() == 0.0 and float(bp.grad.abs().sum()) == 0.0


def test_anchor_head_equals_decode_then_gdloss_on_gpu():
    """Fused front end == torch decode + the element-wise GDLoss kernel (both ours)."""
    kw = CONFIGS[0]
    g = cuda(synth.make_anchor_head_batch(40_000, 40_000, pos_frac=0.02, seed=11))
    pos = g['pos_inds']
    bp = g['bbox_pred'].clone().requires_grad_(True)
    a = g['anchors'][pos]
    loss_a = GDLoss(**kw)(gd_oracle.decode_delta_xyzwlhr(a, bp[pos]),
                          gd_oracle.decode_delta_xyzwlhr(a, g['bbox_targets'][pos]),
                          avg_factor=77.0)
    loss_a.backward()
    bp2 = g['bbox_pred'].clone().requires_grad_(True)
    loss_b = GDAnchorHeadLoss(dict(type='GDLoss', **kw))(
        g['anchors'], bp2, g['bbox_targets'], None, pos_inds=pos, avg_factor=77.0)
    loss_b.backward()
    assert abs(loss_a.item() - loss_b.item()) <= 2e-5 * abs(loss_a.item())
    assert torch.allclose(bp.grad, bp2.grad, rtol=2e-3, atol=2e-5 * float(bp.grad.abs().max()))


def test_anchor_head_mask_mode_is_graph_capturable():
    """labels mode has no host sync: the whole loss tail captures into a CUDA graph (f4)."""
    kw = CONFIGS[0]
    g = cuda(synth.make_anchor_head_batch(100_000, 50_000, pos_frac=0.005, seed=4))
    head = GDAnchorHeadLoss(dict(type='GDLoss', **kw), 1)

    def run(bp):
        loss = head(g['anchors'], bp, g['bbox_targets'], g['bbox_weights'], labels=g['labels'],
                    num_classes=3, avg_factor=500.0)
        grad, = torch.autograd.grad(loss, bp)
        return loss, grad
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        # the leaf lives on the capture stream (a leaf first used on the default stream
        # would make autograd sync with it and invalidate the capture)
        bp_static = g['bbox_pred'].clone().requires_grad_(True)
        run(bp_static)                                        # warm-up: workspace, module load
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            cap_loss, cap_grad = run(bp_static)
    torch.cuda.synchronize()
    new_pred = g['bbox_pred'] * 1.25
    with torch.no_grad():
        bp_static.copy_(new_pred)
    graph.replay()
    torch.cuda.synchronize()
    eager_loss, eager_grad = run(new_pred.clone().requires_grad_(True))
    assert cap_loss.item() == eager_loss.item()
    assert torch.equal(cap_grad, eager_grad)


def test_anchor_head_labels_mode_counts_positives_on_the_device():
    """labels mode, reduction='mean', no avg_factor: the mean runs over the positives like the
    reference's loss.mean() over the gathered rows (gd_anchor3d_head.py:102-141), their number
    counted on the device -- equal to the index mode, where the host knows len(pos_inds); no
    positives give 0; a device-tensor avg_factor equals the same number passed from the host.
    All of it captures into one CUDA graph."""
    kw = CONFIGS[0]
    b = synth.make_anchor_head_batch(50_000, 20_000, pos_frac=0.01, seed=13)
    g = cuda(b)
    head = GDAnchorHeadLoss(dict(type='GDLoss', **kw), 1)
    ref_loss, ref_grad = oracle_anchor(kw, b, 1, None)           # mean over the P positives
    bp = g['bbox_pred'].clone().requires_grad_(True)
    loss = head(g['anchors'], bp, g['bbox_targets'], g['bbox_weights'], labels=g['labels'],
                num_classes=3)
    loss.backward()
    check(loss.item(), bp.grad.cpu().double().numpy(), ref_loss, ref_grad, 'device count')
    bp2 = g['bbox_pred'].clone().requires_grad_(True)
    loss2 = head(g['anchors'], bp2, g['bbox_targets'], g['bbox_weights'], pos_inds=g['pos_inds'])
    loss2.backward()
    assert abs(loss.item() - loss2.item()) <= 2e-6 * abs(loss2.item())
    assert torch.allclose(bp.grad, bp2.grad, rtol=2e-6, atol=0)
    # device avg_factor == host avg_factor
    af = torch.tensor([123.0], device='cuda')
    l_dev = head(g['anchors'], g['bbox_pred'], g['bbox_targets'], g['bbox_weights'],
                 labels=g['labels'], num_classes=3, avg_factor=af)
    l_host = head(g['anchors'], g['bbox_pred'], g['bbox_targets'], g['bbox_weights'],
                  labels=g['labels'], num_classes=3, avg_factor=123.0)
    assert abs(l_dev.item() - l_host.item()) <= 2e-7 * abs(l_host.item())
    # no positives
    bp3 = g['bbox_pred'].clone().requires_grad_(True)
    l0 = head(g['anchors'], bp3, g['bbox_targets'], g['bbox_weights'],
              labels=torch.full((50_000,), 3, device='cuda'), num_classes=3)
    l0.backward()
    assert l0.item() == 0.0 and float(bp3.grad.abs().sum()) == 0.0
    # graph capture with the device-side count; replay after the labels changed
    labels = g['labels'].clone()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        bps = g['bbox_pred'].clone().requires_grad_(True)
        torch.autograd.grad(head(g['anchors'], bps, g['bbox_targets'], g['bbox_weights'],
                                 labels=labels, num_classes=3), bps)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            cl = head(g['anchors'], bps, g['bbox_targets'], g['bbox_weights'], labels=labels,
                      num_classes=3)
            cg, = torch.autograd.grad(cl, bps)
    graph.replay()
    torch.cuda.synchronize()
    assert cl.item() == loss.item() and torch.equal(cg, bp.grad)
    labels[b['pos_inds'][::2].cuda()] = 3                         # drop every other positive
    graph.replay()
    torch.cuda.synchronize()
    bp4 = g['bbox_pred'].clone().requires_grad_(True)
    l4 = head(g['anchors'], bp4, g['bbox_targets'], g['bbox_weights'], labels=labels, num_classes=3)
    l4.backward()
    assert cl.item() == l4.item() and torch.equal(cg, bp4.grad)
    assert abs(l4.item() - loss.item()) > 1e-4 * abs(loss.item())


def test_center_head_is_graph_capturable_with_device_avg_factor():
    """CenterGDHead.loss computes avg_factor as heatmap.eq(1).float().sum().item()
    (gd_centerpoint_head.py:407): with the count kept on the device the whole GD branch -- the
    count, the decode, the loss and its gradient -- captures into one CUDA graph (f4)."""
    kw = dict(CONFIGS[0], tau=0.0)
    c = synth.make_center_head_batch(20_000, seed=16)
    g = cuda(c)
    head = GDCenterHeadLoss(dict(type='GDLoss', **kw), c['coder'])
    heat = (torch.rand(2, 128, 128, device='cuda') < 0.01).float()

    def run(p):
        num_pos = heat.eq(1).float().sum().clamp(min=1)            # :407 without .item()
        loss = head(p, g['pos_ind'], g['target_box'], avg_factor=num_pos)
        grad, = torch.autograd.grad(loss, p)
        return loss, grad
    ref_loss, ref_grad = oracle_center(kw, c, float(heat.sum().clamp(min=1)))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ps = g['pred'].clone().requires_grad_(True)
        run(ps)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            cl, cg = run(ps)
    graph.replay()
    torch.cuda.synchronize()
    check(cl.item(), cg.cpu().double().numpy()[:, :7], ref_loss, ref_grad[:, :7], 'center graph')
    heat.zero_()                                                   # no positives -> avg_factor 1
    graph.replay()
    torch.cuda.synchronize()
    ref1, _ = oracle_center(kw, c, 1.0)
    assert abs(cl.item() - ref1) <= RTOL * abs(ref1)
    # and GDLoss itself on the head's strided views with the device scalar
    dec = gd_oracle.decode_centerpoint_yaw(c['pos_ind'][..., 1:], c['pred'], **c['coder']).cuda()
    p7 = dec.detach().requires_grad_(True)
    out = GDLoss(**kw)(p7[..., :7], g['target_box'][..., :7], avg_factor=torch.ones((), device='cuda'))
    out.backward()
    assert abs(out.item() - ref1) <= 2e-5 * abs(ref1)


def oracle_center(kw, c, avg, weight=None):
    p = c['pred'].double().requires_grad_(True)
    mod = gd_oracle.GDLossOracle(**kw)
    if weight is None:
        loss = gd_oracle.center_head_gd_loss(mod, p, c['pos_ind'], c['target_box'].double(),
                                             c['coder'], avg_factor=avg)
    else:
        dec = gd_oracle.decode_centerpoint_yaw(c['pos_ind'][..., 1:], p, **c['coder'])[..., :7]
        loss = mod(dec, c['target_box'].double()[..., :7], weight.double(), avg_factor=avg)
    loss.backward()
    return loss.item(), p.grad.numpy()


@pytest.mark.parametrize('kw', CONFIGS, ids=lambda k: k['loss_type'])
def test_center_head_vs_oracle(kw):
    kw = dict(kw, tau=0.0)                                     # nuScenes configs use tau=0
    c = synth.make_center_head_batch(30_000, seed=6)
    ref_loss, ref_grad = oracle_center(kw, c, 30_000.0)
    g = cuda(c)
    p = g['pred'].clone().requires_grad_(True)
    loss = GDCenterHeadLoss(dict(type='GDLoss', **kw), c['coder'])(
        p, g['pos_ind'], g['target_box'], avg_factor=30_000.0)
    loss.backward()
    grad = p.grad.cpu().double().numpy()
    assert np.abs(grad[:, 7:]).max() == 0.0 and np.abs(ref_grad[:, 7:]).max() == 0.0
    check(loss.item(), grad[:, :7], ref_loss, ref_grad[:, :7], kw)


def test_center_head_variants():
    kw = dict(CONFIGS[0], tau=0.0)
    # 9 channels (no velocity), norm_bbox=False, weights, strided gather result, N=0
    c = synth.make_center_head_batch(4_001, channels=9, seed=9,
                                     coder=dict(synth.CENTER_CODER_NUS, norm_bbox=False))
    c['pred'][:, 3:6] = c['pred'][:, 3:6].exp()
    w = torch.rand(4_001) * (torch.rand(4_001) < 0.7)
    ref_loss, ref_grad = oracle_center(kw, c, 123.0, weight=w)
    g = cuda(c)
    wide = torch.zeros(4_001, 13, device='cuda')
    wide[:, 2:11] = g['pred']
    view = wide[:, 2:11].detach().requires_grad_(True)
    loss = GDCenterHeadLoss(dict(type='GDLoss', **kw), c['coder'])(
        view, g['pos_ind'], g['target_box'], weight=w.cuda(), avg_factor=123.0)
    (loss * 0.5).backward()
    check(loss.item(), view.grad.cpu().double().numpy()[:, :7] * 2.0, ref_loss, ref_grad[:, :7],
          'center variants')
    p0 = torch.zeros(0, 9, device='cuda', requires_grad=True)
    l0 = GDCenterHeadLoss(dict(type='GDLoss', **kw), c['coder'])(
        p0, torch.zeros(0, 3, dtype=torch.long, device='cuda'), torch.zeros(0, 9, device='cuda'),
        avg_factor=1.0)
    l0.backward()
    assert l0.item() == 0.0 and p0.grad.shape == (0, 9)


def test_decode_golden_vectors():
    """Fixtures written from the UNMODIFIED reference classes (CenterPointBBoxYawCoder +
    GDLoss; the anchor decode is the restated upstream coder + reference GDLoss)."""
    path = os.path.join(HERE, 'golden', 'gd_decode_golden.npz')
    z = np.load(path)
    manifest = json.loads(bytes(z['manifest']).decode())
    assert len(manifest) >= 8
    for case in manifest:
        cid, kw = case['id'], case['kwargs']
        t = lambda name: torch.from_numpy(z[f'{cid}/{name}'])   # noqa: E731
        if case['head'] == 'center':
            p = t('pred').cuda().requires_grad_(True)
            loss = GDCenterHeadLoss(dict(type='GDLoss', **kw), case['coder'])(
                p, t('pos_ind').cuda(), t('target_box').cuda(), avg_factor=case['avg_factor'])
            loss.backward()
            grad = p.grad.cpu().double().numpy()[:, :7]
            ref_grad = z[f'{cid}/grad_f64'][:, :7]
        else:
            bp = t('bbox_pred').cuda().requires_grad_(True)
            loss = GDAnchorHeadLoss(dict(type='GDLoss', **kw), case['decode_weight'])(
                t('anchors').cuda(), bp, t('bbox_targets').cuda(), t('bbox_weights').cuda(),
                pos_inds=t('pos_inds').cuda(), avg_factor=case['avg_factor'])
            loss.backward()
            grad = bp.grad.cpu().double().numpy()
            ref_grad = z[f'{cid}/grad_f64']
        check(loss.item(), grad, float(z[f'{cid}/loss_f64']), ref_grad, case)
