"""CPU tests: the oracle restatement vs the reference's golden vectors (and vs
the live reference when /root/reference is mounted), plus analytic known-answer
checks that need no oracle."""
import math

import numpy as np
import os

import pytest
import torch

from oracle import gd_oracle, ref_loader
from mmdet3d_gaussian_b200 import synth


def _run_oracle(entry, z, dtype=torch.float64):
    cid = entry['id']
    pred = torch.from_numpy(z[f"in/{entry['inputs']}/pred"]).to(dtype)
    target = torch.from_numpy(z[f"in/{entry['inputs']}/target"]).to(dtype)
    weight = None
    if entry['weight_mode'] is not None:
        weight = torch.from_numpy(z[f'case/{cid}/weight']).to(dtype)
    mod = gd_oracle.GDLossOracle(**entry['kwargs'])
    return gd_oracle.loss_and_grad(mod, pred, target, weight,
                                   avg_factor=entry['avg_factor'],
                                   reduction_override=entry['override'])


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | same_inf | (np.abs(a - b) <= atol + rtol * np.abs(b))
    return bool(ok.all())


def test_oracle_matches_reference_golden_fp64(golden):
    z, manifest = golden
    assert len(manifest) >= 250
    n_raise = 0
    for entry in manifest:
        if 'raises' in entry:
            n_raise += 1
            with pytest.raises((ValueError, RuntimeError)):
                _run_oracle(entry, z)
            continue
        loss, grad = _run_oracle(entry, z)
        cid = entry['id']
        assert _close(loss.numpy(), z[f'case/{cid}/loss_f64'], 1e-11, 1e-13), entry
        assert _close(grad.numpy(), z[f'case/{cid}/grad_f64'], 1e-9, 1e-12), entry
    assert n_raise >= 4      # sum+avg_factor (x3 losses x3 weights) and [N] zero weights


def test_oracle_matches_reference_golden_fp32(golden):
    """Same op order as the reference => fp32 results agree to a few ulps on the
    well-separated set (the fp32 'port' is a faithful CPU baseline)."""
    z, manifest = golden
    for entry in manifest:
        if 'raises' in entry or entry['inputs'] != 'kitti_s0.3':
            continue
        loss, _ = _run_oracle(entry, z, torch.float32)
        ref = z[f"case/{entry['id']}/loss_f32"]
        assert _close(loss.numpy(), ref, 2e-4, 1e-6), entry


@pytest.mark.skipif(not ref_loader.reference_available(),
                    reason='reference checkout only exists in the build container')
@pytest.mark.parametrize('loss_type', gd_oracle.LOSS_TYPES)
def test_oracle_matches_live_reference(loss_type):
    ref = ref_loader.load_reference()
    pred, target, weight = synth.make_pairs(2000, 'nuscenes', seed=77,
                                            weights='bernoulli')
    pred, target, weight = pred.double(), target.double(), weight.double()
    fun = 'expm1' if loss_type == 'kfiou3d' else 'log1p'
    for tau in (0.0, 1.0):
        kw = dict(loss_type=loss_type, fun=fun, tau=tau, alpha=0.7,
                  loss_weight=5.0)
        a = gd_oracle.loss_and_grad(ref.GDLoss(**kw), pred, target, weight,
                                    avg_factor=123.0)
        b = gd_oracle.loss_and_grad(gd_oracle.GDLossOracle(**kw), pred, target,
                                    weight, avg_factor=123.0)
        assert _close(b[0].numpy(), a[0].numpy(), 1e-12, 0)
        assert _close(b[1].numpy(), a[1].numpy(), 1e-9, 1e-13)


# ---------------------------------------------------------------------------
# analytic known-answer tests (independent of the reference)
# ---------------------------------------------------------------------------
def _rows(loss_type, pred, target, **kw):
    kw.setdefault('fun', 'none')
    kw.setdefault('tau', 0.0)
    mod = gd_oracle.GDLossOracle(loss_type, reduction='none', **kw)
    return mod(pred.double(), target.double())


def test_kat_gwd_axis_aligned_equal_extent():
    # equal extents, equal yaw: GWD = centre distance / (2 (a b e a b e)^(1/6))
    p = torch.tensor([[1., 2., 3., 2., 4., 1., 0.3]])
    t = torch.tensor([[4., 6., 3., 2., 4., 1., 0.3]])
    got = _rows('gwd3d', p, t)
    a, b, e = 1.0, 2.0, 0.5
    want = 5.0 / (2 * (a * b * e) ** (2 / 6))
    assert abs(got.item() - want) < 1e-12
    assert abs(_rows('gwd3d', p, t, normalize=False).item() - 5.0) < 1e-12


@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d', 'jd3d',
                                       'kld3d_symmax', 'kld3d_symmin'])
def test_kat_identity_is_zero(loss_type):
    _, t, _ = synth.make_pairs(64, 'kitti', seed=3)
    d = _rows(loss_type, t, t)
    assert d.abs().max().item() < 1e-6


def test_kat_symmetry_relations():
    p, t, _ = synth.make_pairs(256, 'waymo', seed=5)
    for sq in (True, False):
        kpt = _rows('kld3d', p, t, sqrt=False)
        ktp = _rows('kld3d', t, p, sqrt=False)
        jd = _rows('jd3d', p, t, sqrt=False)
        assert torch.allclose(jd, 0.5 * (kpt + ktp), rtol=1e-12, atol=1e-14)
        smax = _rows('kld3d_symmax', p, t, sqrt=sq)
        smin = _rows('kld3d_symmin', p, t, sqrt=sq)
        jd_s = _rows('jd3d', p, t, sqrt=sq)
        assert (smax >= jd_s - 1e-12).all() and (jd_s >= smin - 1e-12).all()
    # Bhattacharyya and GWD are symmetric
    assert torch.allclose(_rows('bd3d', p, t), _rows('bd3d', t, p), rtol=1e-10)
    assert torch.allclose(_rows('gwd3d', p, t, center_offset=(0, 0, 0)),
                          _rows('gwd3d', t, p, center_offset=(0, 0, 0)), rtol=1e-10)


@pytest.mark.parametrize('loss_type', ['gwd3d', 'kld3d', 'bd3d', 'kfiou3d'])
def test_kat_rigid_motion_and_yaw_period(loss_type):
    p, t, _ = synth.make_pairs(128, 'kitti', seed=9)
    p, t = p.double(), t.double()
    base = _rows(loss_type, p, t, center_offset=(0, 0, 0))
    th = 0.83
    c, s = math.cos(th), math.sin(th)

    def move(b):
        b = b.clone()
        x, y = b[:, 0].clone(), b[:, 1].clone()
        b[:, 0] = c * x - s * y + 3.0
        b[:, 1] = s * x + c * y - 7.0
        b[:, 2] += 1.5
        b[:, 6] += th
        return b
    moved = _rows(loss_type, move(p), move(t), center_offset=(0, 0, 0))
    assert torch.allclose(moved, base, rtol=1e-9, atol=1e-12)
    p2 = p.clone()
    p2[:, 6] += math.pi
    assert torch.allclose(_rows(loss_type, p2, t, center_offset=(0, 0, 0)), base,
                          rtol=1e-9, atol=1e-12)


def test_pairwise_oracle_matches_rows():
    b1, b2, _ = synth.make_pairs(40, 'waymo', seed=2)
    mat = gd_oracle.pairwise_distance(b1.double(), b2[:9].double(), 'kld3d',
                                      fun='log1p', tau=1.0, chunk_rows=16)
    assert mat.shape == (40, 9)
    for i, j in ((0, 0), (17, 8), (39, 3)):
        d = gd_oracle.GDLossOracle('kld3d', fun='log1p', tau=1.0,
                                   reduction='none')(b1[i:i + 1].double(),
                                                     b2[j:j + 1].double())
        assert abs(mat[i, j].item() - d.item()) < 1e-13


def test_weight_reduce_contract():
    """Row a11: the upstream mmdet contract, unit-tested on its own."""
    loss = torch.tensor([1., 2., 3., 4.], dtype=torch.float64)
    w = torch.tensor([1., 0., 0.5, 2.], dtype=torch.float64)
    assert gd_oracle.weight_reduce(loss, w, 'mean').item() == pytest.approx(10.5 / 4)
    assert gd_oracle.weight_reduce(loss, w, 'sum').item() == pytest.approx(10.5)
    assert gd_oracle.weight_reduce(loss, w, 'mean', 3.0).item() == pytest.approx(3.5)
    assert gd_oracle.weight_reduce(loss, w, 'none', 3.0).shape == (4,)
    with pytest.raises(ValueError):
        gd_oracle.weight_reduce(loss, w, 'sum', 3.0)


# ---------------------------------------------------------------------------
# f1: decoders in front of the loss
# ---------------------------------------------------------------------------
def _decode_golden():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                        'gd_decode_golden.npz')
    z = np.load(path)
    return z, json.loads(bytes(z['manifest']).decode())


def test_head_oracles_match_decode_golden():
    """``center_head_gd_loss`` / ``anchor_head_gd_loss`` vs the fixtures written from the
    unmodified reference ``CenterPointBBoxYawCoder`` + ``GDLoss``."""
    z, manifest = _decode_golden()
    assert {c['head'] for c in manifest} == {'center', 'anchor'}
    for case in manifest:
        cid = case['id']
        t = lambda name: torch.from_numpy(z[f'{cid}/{name}'])   # noqa: E731
        mod = gd_oracle.GDLossOracle(**case['kwargs'])
        if case['head'] == 'center':
            p = t('pred').double().requires_grad_(True)
            loss = gd_oracle.center_head_gd_loss(mod, p, t('pos_ind'), t('target_box').double(),
                                                 case['coder'], avg_factor=case['avg_factor'])
        else:
            p = t('bbox_pred').double().requires_grad_(True)
            loss = gd_oracle.anchor_head_gd_loss(
                mod, t('anchors').double(), p, t('bbox_targets').double(),
                t('bbox_weights').double(), t('pos_inds'),
                decode_weight=case['decode_weight'], avg_factor=case['avg_factor'])
        loss.backward()
        assert _close(loss.item(), z[f'{cid}/loss_f64'], 1e-11, 1e-13), case
        assert _close(p.grad.numpy(), z[f'{cid}/grad_f64'], 1e-9, 1e-12), case


@pytest.mark.skipif(not os.path.isdir(ref_loader.CODER_DIR), reason='needs /root/reference')
def test_center_decode_matches_live_reference_coder():
    coder_cls = ref_loader.load_reference_center_coder()
    c = synth.make_center_head_batch(500, seed=77)
    coder = coder_cls(pc_range=[-51.2, -51.2], out_size_factor=4, voxel_size=[0.2, 0.2],
                      code_size=9, norm_bbox=True)
    for dtype in (torch.float64, torch.float32):
        ref = coder.decode(c['pos_ind'][..., 1:], c['pred'].to(dtype), correct_yaw=False)
        ours = gd_oracle.decode_centerpoint_yaw(c['pos_ind'][..., 1:], c['pred'].to(dtype),
                                                **c['coder'])
        assert torch.equal(ref, ours)


def test_anchor_decode_known_answers():
    """Zero deltas decode to the anchor itself; the z shift follows the bottom-centre
    convention (z_bottom + h/2 is the quantity the delta moves)."""
    b = synth.make_anchor_head_batch(64, 64, pos_frac=1.1, seed=0)
    a = b['anchors'].double()
    assert torch.allclose(gd_oracle.decode_delta_xyzwlhr(a, torch.zeros_like(a)), a,
                          rtol=0, atol=1e-12)
    d = torch.zeros_like(a)
    d[:, 5] = math.log(2.0)
    out = gd_oracle.decode_delta_xyzwlhr(a, d)
    assert torch.allclose(out[:, 5], 2 * a[:, 5])
    assert torch.allclose(out[:, 2] + out[:, 5] / 2, a[:, 2] + a[:, 5] / 2)


def _mismatch_golden():
    import json as _json
    import os as _os
    here = _os.path.dirname(_os.path.abspath(__file__))
    z = np.load(_os.path.join(here, 'golden', 'gd_mismatch_golden.npz'))
    return z, _json.loads(bytes(z['manifest']).decode())


def test_oracle_matches_reference_on_mismatched_boxes():
    """Oracle (fp64) == the unmodified reference's fp64 output on strongly mismatched pairs
    (extent ratios 10 ... 1000, shifts to 1000 m, elongated boxes):
    tests/golden/gd_mismatch_golden.npz, written by oracle/make_mismatch_golden.py."""
    z, man = _mismatch_golden()
    pred = torch.from_numpy(z['pred']).double()
    target = torch.from_numpy(z['target']).double()
    assert len(man['cases']) == 20 and pred.shape[0] == len(man['tags']) == 120
    for c in man['cases']:
        loss, grad = gd_oracle.loss_and_grad(gd_oracle.GDLossOracle(**c['kwargs']), pred, target)
        rl, rg = z[f"case/{c['id']}/loss"], z[f"case/{c['id']}/grad"]
        assert np.allclose(loss.numpy(), rl, rtol=1e-11, atol=1e-300), c
        gn = np.abs(rg).max(1, keepdims=True)
        assert (np.abs(grad.numpy() - rg) <= 1e-10 * gn + 1e-300).all(), c


# ---------------------------------------------------------------------------
# SimOTA dynamic-k matching: restatement vs the reference method
# ---------------------------------------------------------------------------
def test_simota_restatement_matches_reference_goldens():
    """``gd_oracle.simota_dynamic_k_matching`` against fixtures written by the UNMODIFIED
    ``SimOTABEVAssigner.dynamic_k_matching`` (oracle/make_simota_golden.py; tie-free inputs)."""
    import json
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                             'gd_simota_golden.npz'))
    manifest = json.loads(bytes(z['manifest']).decode())
    assert len(manifest) >= 7
    for c in manifest:
        cid = c['id']
        cost, ious = torch.from_numpy(z[f'{cid}/cost']), torch.from_numpy(z[f'{cid}/ious'])
        assigned, matched, dks = gd_oracle.simota_dynamic_k_matching(cost, ious, c['candidate_topk'])
        assert torch.equal(assigned, torch.from_numpy(z[f'{cid}/assigned'])), c
        assert torch.allclose(matched, torch.from_numpy(z[f'{cid}/matched_ious']), rtol=0, atol=1e-15)
        assert int(dks.min()) >= 1 and (assigned > 0).sum() >= 1


@pytest.mark.skipif(not os.path.isfile(ref_loader.SIMOTA_FILE), reason='needs /root/reference')
def test_simota_restatement_matches_live_reference():
    cls = ref_loader.load_reference_simota()
    g = torch.Generator().manual_seed(5)
    for n, m, topk in ((400, 9, 10), (33, 4, 10), (7, 2, 10), (800, 25, 6)):
        d = torch.rand(n, m, generator=g, dtype=torch.float64)
        ious = 1.0 / (1.0 + torch.where(torch.rand(n, m, generator=g) < 0.1, d * 0.03, 0.2 + d))
        cost = -torch.log(ious)                       # the reference's own cost shape (sim:94)
        valid = torch.ones(n, dtype=torch.bool)
        mi, mg = cls(candidate_topk=topk).dynamic_k_matching(cost.clone(), ious.clone(), m, valid)
        want = torch.zeros(n, dtype=torch.int64)
        want[valid] = mg + 1
        assigned, matched, _ = gd_oracle.simota_dynamic_k_matching(cost, ious, topk)
        assert torch.equal(assigned, want)
        assert torch.allclose(matched[valid], mi, rtol=0, atol=1e-15)
