"""CPU tests of the fused decode front ends (csrc/gd_decode.cuh), host-compiled.

TEST-ONLY build (tests/host_math/harness.cpp, g++), like test_host_math.py: the
float64 instantiation pins the decode prologue + Jacobian epilogue against the
oracle's restatement of the two reference call sites
(``gd_anchor3d_head.py:107-141``, ``gd_centerpoint_head.py:413-434``) under
autograd; the float32 instantiation previews the device arithmetic and must stay
inside the 1e-5 budget and beat the reference formulation's own fp32 error.
"""
import ctypes
import itertools

import numpy as np
import pytest
import torch

from mmdet3d_gaussian_b200 import synth
from oracle import gd_oracle
from test_host_math import FUN, LT, hostlib  # noqa: F401  (fixture)

VP = ctypes.c_void_p


def _ptr(a):
    return a.ctypes.data_as(VP)


def host_anchor(lib, lt, anchors, dp, dt, off, alpha, tau, fun, flag, dtype):
    n = dp.shape[0]
    npdt = np.float64 if dtype == 'f64' else np.float32
    a, p, t = (np.ascontiguousarray(x.numpy().astype(npdt)) for x in (anchors, dp, dt))
    ol, og = np.zeros(n, npdt), np.zeros((n, 7), npdt)
    getattr(lib, 'gd_host_anchor_' + dtype)(
        ctypes.c_int(LT.index(lt)), ctypes.c_long(n), _ptr(a), _ptr(p), _ptr(t),
        (ctypes.c_double * 3)(*off), ctypes.c_double(alpha), ctypes.c_double(tau),
        ctypes.c_int(FUN[fun]), ctypes.c_int(int(flag)), _ptr(ol), _ptr(og))
    return ol.astype(np.float64), og.astype(np.float64)


def oracle_anchor(lt, anchors, dp, dt, off, alpha, tau, fun, flag, dtype=torch.float64):
    kw = {'normalize' if lt == 'gwd3d' else 'sqrt': flag}
    m = gd_oracle.GDLossOracle(lt, center_offset=off, fun=fun, tau=tau, alpha=alpha,
                               reduction='none', **kw)
    p = dp.to(dtype).clone().requires_grad_(True)
    a = anchors.to(dtype)
    out = m(gd_oracle.decode_delta_xyzwlhr(a, p), gd_oracle.decode_delta_xyzwlhr(a, dt.to(dtype)))
    out.backward(torch.ones_like(out))
    return out.detach().double().numpy(), p.grad.double().numpy()


def host_center(lib, lt, pred, locs, target, coder, off, alpha, tau, fun, flag, dtype):
    n = pred.shape[0]
    npdt = np.float64 if dtype == 'f64' else np.float32
    p = np.ascontiguousarray(pred.numpy().astype(npdt))
    t = np.ascontiguousarray(target.numpy().astype(npdt))
    lo = np.ascontiguousarray(locs.numpy().astype(np.int64))
    cd = (ctypes.c_double * 4)(coder['out_size_factor'] * coder['voxel_size'][0],
                               coder['out_size_factor'] * coder['voxel_size'][1],
                               coder['pc_range'][0], coder['pc_range'][1])
    ol, og = np.zeros(n, npdt), np.zeros((n, 7), npdt)
    getattr(lib, 'gd_host_center_' + dtype)(
        ctypes.c_int(LT.index(lt)), ctypes.c_long(n), _ptr(p), ctypes.c_long(p.shape[1]),
        _ptr(lo), _ptr(t), ctypes.c_long(t.shape[1]), cd, ctypes.c_int(int(coder['norm_bbox'])),
        (ctypes.c_double * 3)(*off), ctypes.c_double(alpha), ctypes.c_double(tau),
        ctypes.c_int(FUN[fun]), ctypes.c_int(int(flag)), _ptr(ol), _ptr(og))
    return ol.astype(np.float64), og.astype(np.float64)


def oracle_center(lt, pred, pos_ind, target, coder, off, alpha, tau, fun, flag,
                  dtype=torch.float64):
    kw = {'normalize' if lt == 'gwd3d' else 'sqrt': flag}
    m = gd_oracle.GDLossOracle(lt, center_offset=off, fun=fun, tau=tau, alpha=alpha,
                               reduction='none', **kw)
    p = pred.to(dtype).clone().requires_grad_(True)
    out = gd_oracle.center_head_gd_loss(m, p, pos_ind, target.to(dtype), coder)
    out.backward(torch.ones_like(out))
    return out.detach().double().numpy(), p.grad.double().numpy()


def _positives(seed, n=1500):
    b = synth.make_anchor_head_batch(n, n, pos_frac=1.1, seed=seed)
    return b['anchors'], b['bbox_pred'], b['bbox_targets']


@pytest.mark.parametrize('lt', LT)
def test_anchor_decode_closed_form_fp64(hostlib, lt):
    anchors, dp, dt = _positives(3)
    funs = ['nlog', 'expm1', 'none'] if lt == 'kfiou3d' else ['log1p', 'none']
    for fun, tau, alpha, flag, off in itertools.product(
            funs, (0.0, 1.0), (1.0, 0.5), (True, False), ((0, 0, 0.5), (0.1, -0.2, 0.3))):
        a = (lt, anchors, dp, dt, off, alpha, tau, fun, flag)
        ol, og = host_anchor(hostlib, *a, 'f64')
        rl, rg = oracle_anchor(*a)
        el = np.max(np.abs(ol - rl) / np.maximum(np.abs(rl), 1e-3))
        eg = np.max(np.abs(og - rg) / np.maximum(np.abs(rg).max(1, keepdims=True), 1e-3))
        assert el < 1e-9 and eg < 1e-8, (a[0], a[4:], el, eg)


@pytest.mark.parametrize('lt', ['gwd3d', 'kld3d', 'bd3d', 'jd3d'])
def test_anchor_decode_fp32_budget(hostlib, lt):
    anchors, dp, dt = _positives(4, 20000)
    a = (lt, anchors, dp, dt, (0, 0, 0.5), 1.0, 1.0, 'log1p', True)
    ol, og = host_anchor(hostlib, *a, 'f32')
    rl, rg = oracle_anchor(*a)
    fl, fg = oracle_anchor(*a, dtype=torch.float32)

    def errs(l, g):
        el = np.abs(l - rl) / np.maximum(np.abs(rl), 1e-30)
        eg = np.linalg.norm(g - rg, axis=1) / np.maximum(np.linalg.norm(rg, axis=1), 1e-30)
        return el.max(), eg.max()
    ours, ref32 = errs(ol, og), errs(fl, fg)
    assert ours[0] <= 5e-6 and ours[1] <= 5e-6, (ours, ref32)
    assert ours[0] <= ref32[0] and ours[1] <= ref32[1], (ours, ref32)


@pytest.mark.parametrize('lt', LT)
def test_center_decode_closed_form_fp64(hostlib, lt):
    c = synth.make_center_head_batch(1200, seed=2)
    funs = ['nlog', 'expm1', 'none'] if lt == 'kfiou3d' else ['log1p', 'none']
    for norm_bbox in (True, False):
        coder = dict(c['coder'], norm_bbox=norm_bbox)
        pred = c['pred'].clone()
        if not norm_bbox:
            pred[:, 3:6] = pred[:, 3:6].exp()
        for fun, tau, alpha, flag in itertools.product(funs, (0.0, 1.0), (1.0, 0.5),
                                                       (True, False)):
            off = (0, 0, 0.5)
            ol, og = host_center(hostlib, lt, pred, c['pos_ind'][:, 1:], c['target_box'], coder,
                                 off, alpha, tau, fun, flag, 'f64')
            rl, rg = oracle_center(lt, pred, c['pos_ind'], c['target_box'], coder, off, alpha,
                                   tau, fun, flag)
            assert np.abs(rg[:, 7:]).max() == 0.0
            rg = rg[:, :7]
            el = np.max(np.abs(ol - rl) / np.maximum(np.abs(rl), 1e-3))
            eg = np.max(np.abs(og - rg) / np.maximum(np.abs(rg).max(1, keepdims=True), 1e-3))
            assert el < 1e-9 and eg < 1e-8, (lt, norm_bbox, fun, tau, alpha, flag, el, eg)


@pytest.mark.parametrize('lt', ['gwd3d', 'kld3d', 'bd3d'])
def test_center_decode_fp32_budget(hostlib, lt):
    c = synth.make_center_head_batch(20000, seed=5)
    a = (c['coder'], (0, 0, 0.5), 1.0, 0.0, 'log1p', True)
    ol, og = host_center(hostlib, lt, c['pred'], c['pos_ind'][:, 1:], c['target_box'], *a, 'f32')
    rl, rg = oracle_center(lt, c['pred'], c['pos_ind'], c['target_box'], *a)
    fl, fg = oracle_center(lt, c['pred'], c['pos_ind'], c['target_box'], *a,
                           dtype=torch.float32)
    rg, fg = rg[:, :7], fg[:, :7]

    def errs(l, g):
        el = np.abs(l - rl) / np.maximum(np.abs(rl), 1e-30)
        eg = np.linalg.norm(g - rg, axis=1) / np.maximum(np.linalg.norm(rg, axis=1), 1e-30)
        return el.max(), eg.max()
    ours, ref32 = errs(ol, og), errs(fl, fg)
    assert ours[0] <= 5e-6 and ours[1] <= 5e-6, (ours, ref32)
    assert ours[0] <= ref32[0] and ours[1] <= ref32[1], (ours, ref32)
