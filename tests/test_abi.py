"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads,
exports every symbol ``include/gd_loss_b200.h`` declares, validates arguments
without touching a GPU, and the Python mirror keeps the reference's constructor /
registry / error behaviour (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

from mmdet3d_gaussian_b200 import GDLoss, GDPairwiseDistance, LOSSES, build_loss
from mmdet3d_gaussian_b200 import _lib, build_ext
from mmdet3d_gaussian_b200.losses.gaussian_distance_loss import _scale_and_mode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    return _lib.load()


def test_library_is_in_tree_and_current(lib):
    assert _lib.loaded_path() == build_ext.lib_path()
    assert os.path.dirname(_lib.loaded_path()).endswith('mmdet3d_gaussian_b200')
    assert build_ext.is_current()
    assert lib.gd_abi_version() == 3


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'gd_loss_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(gd_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_config_struct_layout():
    assert ctypes.sizeof(_lib.GDLossConfig) == 32
    cfg = _lib.make_config('bd3d', 'log1p', True, 1.0, 0.5, (0, 0, 0.5))
    assert (cfg.loss_type, cfg.fun, cfg.flag) == (5, 1, 1)
    assert cfg.center_offset[2] == 0.5


def test_sass_is_sm100a_with_bulk_copies():
    """The shipped binary carries sm_100a code using the TMA bulk-copy engine
    (UBLKCP) and mbarriers (SYNCS) -- evidence the hot path is the B200 one."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    elf = subprocess.run([cuobjdump, '-lelf', build_ext.lib_path()], capture_output=True,
                         text=True).stdout
    assert 'sm_100a' in elf
    sass = subprocess.run([cuobjdump, '-sass', '-fun',
                           '_ZN3gdk14gd_warp_kernelILi1ELb1ELi4ELi9ELi1ELb0ELb0EEEvNS_8LossArgsE',
                           build_ext.lib_path()], capture_output=True, text=True).stdout
    assert 'UBLKCP' in sass and 'SYNCS' in sass
    assert 'LDS.128' in sass and 'STS.128' in sass and 'FFMA2' not in sass
    # the opt-in packed variant of the same kernel uses the packed FP32 pipe instructions
    packed = subprocess.run([cuobjdump, '-sass', '-fun',
                             '_ZN3gdk14gd_warp_kernelILi1ELb1ELi4ELi9ELi1ELb1ELb0EEEvNS_8LossArgsE',
                             build_ext.lib_path()], capture_output=True, text=True).stdout
    assert 'FFMA2' in packed and 'FMUL2' in packed and 'UBLKCP' in packed


def test_argument_validation_without_gpu(lib):
    cfg = _lib.make_config('gwd3d', 'log1p', True, 0.0, 1.0, (0, 0, 0.5))
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    call = lib.gd_loss_fwd_bwd
    # negative n, null pointers, unknown enums, missing workspace: rejected up front
    assert call(ctypes.byref(cfg), one, 7, one, 7, null, 0, 0, -1, 1.0, null, null, null,
                null, 0, 0, 0, null) == -1
    assert call(ctypes.byref(cfg), null, 7, one, 7, null, 0, 0, 8, 1.0, null, null, null,
                null, 0, 0, 0, null) == -1
    assert call(ctypes.byref(cfg), one, 7, one, 7, null, 1, 1, 8, 1.0, null, null, null,
                null, 0, 0, 0, null) == -1
    assert call(ctypes.byref(cfg), one, 7, one, 7, null, 0, 0, 8, 1.0, one, null, null,
                null, 0, 0, 0, null) == -2
    assert call(ctypes.byref(cfg), one, 7, one, 7, null, 0, 0, 8, 1.0, null, null, null,
                null, 0, 7, 0, null) == -1
    bad = _lib.make_config('gwd3d', 'log1p', True, 0.0, 1.0, (0, 0, 0.5))
    bad.loss_type = 9
    assert call(ctypes.byref(bad), one, 7, one, 7, null, 0, 0, 8, 1.0, null, null, null,
                null, 0, 0, 0, null) == -1
    # bulk variant on a strided / unaligned layout is an error, not a silent fallback
    assert call(ctypes.byref(cfg), one, 9, one, 7, null, 0, 0, 8, 1.0, null, null, null,
                null, 0, 2, 0, null) == -3
    assert call(ctypes.byref(cfg), ctypes.c_void_p(20), 7, one, 7, null, 0, 0, 8, 1.0, null,
                null, null, null, 0, 2, 0, null) == -3
    assert lib.gd_pairwise(ctypes.byref(cfg), one, 4, one, 4, one, 3, null) == -1
    assert lib.gd_pairwise_row_argmin(ctypes.byref(cfg), one, 4, one, 0, one, one, null) == -1
    assert b'bad argument' in lib.gd_error_string(-1)
    assert lib.gd_loss_workspace_bytes(1 << 24) >= 8 * 65536


def test_module_mirrors_reference_constructor():
    m = GDLoss('gwd3d')
    assert (m.center_offset, m.fun, m.tau, m.alpha, m.reduction, m.loss_weight) == \
        ((0, 0, 0.5), 'log1p', 1.0, 1.0, 'mean', 1.0)        # ref:261-263 defaults
    assert len(m.state_dict()) == 0 and len(list(m.parameters())) == 0
    assert set(GDLoss.BAG_GD_LOSS) == {'gwd3d', 'kld3d', 'jd3d', 'kld3d_symmax',
                                       'kld3d_symmin', 'bd3d', 'kfiou3d'}
    for bad in (dict(loss_type='iou3d'), dict(loss_type='gwd3d', reduction='max'),
                dict(loss_type='gwd3d', fun='expm1'), dict(loss_type='kfiou3d', fun='log1p')):
        with pytest.raises(AssertionError):
            GDLoss(**bad)
    GDLoss('kfiou3d', fun='nlog')
    assert GDLoss('kld3d', sqrt=False).kwargs == {'sqrt': False}
    # shipped config dicts build unchanged through the registry (configs/kitti/*gwd5tau1*)
    cfg = dict(type='GDLoss', loss_type='gwd3d', fun='log1p', tau=1.0, loss_weight=5.0)
    assert isinstance(build_loss(cfg), GDLoss) and 'GDLoss' in LOSSES
    assert isinstance(GDPairwiseDistance('bd3d', sqrt=False), torch.nn.Module)


def test_reduction_contract_folding():
    """mmdet weight_reduce_loss folded to (scale, rows_out) -- SURVEY.md section 8 a11."""
    assert _scale_and_mode('mean', None, 10, 5.0) == (0.5, False)
    assert _scale_and_mode('sum', None, 10, 5.0) == (5.0, False)
    assert _scale_and_mode('none', None, 10, 5.0) == (5.0, True)
    assert _scale_and_mode('mean', 4.0, 10, 5.0) == (1.25, False)
    assert _scale_and_mode('none', 4.0, 10, 5.0) == (5.0, True)
    with pytest.raises(ValueError):
        _scale_and_mode('sum', 4.0, 10, 5.0)
    s, _ = _scale_and_mode('mean', None, 0, 1.0)
    assert s != s


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        GDLoss('gwd3d')(torch.zeros(4, 7), torch.zeros(4, 7))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        GDPairwiseDistance('gwd3d')(torch.zeros(4, 7), torch.zeros(4, 7))


def test_head_front_end_argument_validation_without_gpu(lib):
    cfg = _lib.make_config('gwd3d', 'log1p', True, 1.0, 1.0, (0, 0, 0.5))
    null, one = ctypes.c_void_p(0), ctypes.c_void_p(16)
    dw = (ctypes.c_float * 7)(*[1.0] * 7)
    anchor = lib.gd_anchor_decoded_loss_fwd_bwd

    def call_anchor(pos=one, npos=4, labels=null, grad=one, mode=_lib.GRAD_SCATTER, bw=null,
                    dwp=None, arows=4, loss=null, ws=null, wsb=0, total=16):
        return anchor(ctypes.byref(cfg), one, arows, one, 7, one, 7, bw, 7, dwp, pos, npos,
                      labels, 3, total, 1.0, null, loss, grad, mode, ws, wsb, 0, null)
    # well-formed empty calls are accepted in every mode (nothing is launched)
    assert call_anchor(pos=null, npos=0, labels=one, mode=_lib.GRAD_DENSE, total=0) == 0
    assert call_anchor(npos=0, mode=_lib.GRAD_SCATTER) == 0
    assert call_anchor(npos=0, mode=_lib.GRAD_COMPACT) == 0
    assert call_anchor(npos=0, grad=null, mode=_lib.GRAD_NONE) == 0
    assert call_anchor(arows=0) == -1                          # no anchors
    assert call_anchor(pos=null) == -1                         # index mode without indices
    assert call_anchor(labels=one) == -1                       # both selectors
    assert call_anchor(mode=_lib.GRAD_DENSE) == -1             # dense needs labels
    assert call_anchor(pos=null, npos=0, labels=one, mode=_lib.GRAD_COMPACT) == -1
    assert call_anchor(grad=null) == -1                        # gradient requested, no buffer
    assert call_anchor(bw=one, dwp=None) == -1                 # weights without decode_weight
    assert call_anchor(bw=one, dwp=dw, loss=one) == -2         # loss_sum without workspace
    assert call_anchor(mode=9) == -1

    coder = _lib.make_center_coder((-51.2, -51.2), 4, (0.2, 0.2))
    assert ctypes.sizeof(_lib.GDCenterCoder) == 40
    center = lib.gd_center_decoded_loss_fwd_bwd

    def call_center(c=ctypes.byref(coder), preds=one, n=4, wmode=0, w=null, grad=one, gstride=11,
                    gcols=11, loss=null):
        return center(ctypes.byref(cfg), c, preds, 11, one, 3, one, 11, w, wmode, 1, n, 1.0,
                      null, loss, grad, gstride, gcols, null, 0, 0, null)
    assert call_center(c=None) == -1
    assert call_center(preds=null) == -1
    assert call_center(n=-1) == -1
    assert call_center(wmode=1) == -1                          # weight mode without weights
    assert call_center(gcols=6) == -1                          # gradient rows narrower than 7
    assert call_center(gstride=9) == -1                        # stride < columns
    assert call_center(loss=one) == -2
    assert lib.gd_scale_buffer(null, 8, one, null) == -1
    assert lib.gd_scale_buffer(null, 0, null, null) == 0


def test_head_modules_validate_on_the_host():
    from mmdet3d_gaussian_b200 import GDAnchorHeadLoss, GDCenterHeadLoss
    from mmdet3d_gaussian_b200.heads import _scale
    cfg = dict(type='GDLoss', loss_type='gwd3d', fun='log1p', tau=1.0, loss_weight=5.0)
    head = GDAnchorHeadLoss(cfg, decode_weight=1)
    assert isinstance(head.loss_decoded_bbox, GDLoss) and len(head.state_dict()) == 0
    assert _scale(head.loss_decoded_bbox, 10.0, None) == 0.5
    assert _scale(head.loss_decoded_bbox, None, 4) == 1.25
    with pytest.raises(ValueError):
        _scale(head.loss_decoded_bbox, None, None)             # labels mode needs avg_factor
    with pytest.raises(NotImplementedError):
        _scale(GDLoss('gwd3d', reduction='none'), None, 4)
    with pytest.raises(ValueError):
        GDAnchorHeadLoss(dict(type='SmoothL1Loss'))
    z = torch.zeros(4, 7)
    with pytest.raises(ValueError, match='exactly one'):
        head(z, z, z, z, avg_factor=1.0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        head(z, z, z, z, pos_inds=torch.zeros(1, dtype=torch.long), avg_factor=1.0)
    center = GDCenterHeadLoss(dict(cfg, tau=0.0), dict(
        type='CenterPointBBoxYawCoder', pc_range=(-51.2, -51.2), out_size_factor=4,
        voxel_size=(0.2, 0.2), code_size=9))
    assert center.coder.norm_bbox == 1 and center.coder.out_size_factor == 4
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        center(torch.zeros(4, 11), torch.zeros(4, 3, dtype=torch.long), torch.zeros(4, 11))


def test_pairwise_assign_argument_validation_without_gpu(lib):
    cfg = _lib.make_config('gwd3d', 'log1p', True, 1.0, 1.0, (0, 0, 0.5))
    null, one = ctypes.c_void_p(0), ctypes.c_void_p(16)
    call = lib.gd_pairwise_assign
    need = lib.gd_pairwise_workspace_bytes(256)
    assert need == 256 + 8 * 256 and lib.gd_pairwise_workspace_bytes(-3) == 256
    ok = (one, 8, one, 256, one, one, one, one)
    assert call(ctypes.byref(cfg), *ok, null, 256, 0, null, 0, null) == -2       # no workspace
    assert call(ctypes.byref(cfg), *ok, null, 256, 0, one, need - 1, null) == -2
    assert call(ctypes.byref(cfg), *ok, one, 255, 0, one, need, null) == -1      # stride < m
    assert call(ctypes.byref(cfg), *ok, null, 256, 8, one, need, null) == -1     # unknown flag
    assert call(ctypes.byref(cfg), one, 8, one, 0, one, one, one, one, null, 0, 0, one, need,
                null) == -1                                                       # m == 0
    assert call(ctypes.byref(cfg), one, 8, one, 256, one, one, null, one, null, 256, 0, one,
                need, null) == -1                                                 # null output
    assert call(ctypes.byref(cfg), one, 0, one, 256, one, one, one, one, null, 256, 0, one,
                need, null) == 0                                                  # n == 0
    lab = lib.gd_assign_from_minima
    assert lab(one, one, -1, one, one, 4, 0.6, 0.0, 0.45, 0.45, 1, one, null, null) == -1
    assert lab(null, one, 8, one, one, 4, 0.6, 0.0, 0.45, 0.45, 1, one, null, null) == -1
    assert lab(one, one, 8, null, one, 4, 0.6, 0.0, 0.45, 0.45, 1, one, null, null) == -1
    assert lab(one, one, 0, null, null, 0, 0.6, 0.0, 0.45, 0.45, 1, null, null, null) == 0


def test_assigner_modules_validate_on_the_host():
    from mmdet3d_gaussian_b200 import GDMaxSimAssigner, GDSimilarity3D
    with pytest.raises(NotImplementedError):
        GDMaxSimAssigner(0.6, 0.45, gt_max_assign_all=True)
    a = GDMaxSimAssigner(0.6, (0.1, 0.45), min_pos_iou=0.3, loss_type='bd3d', sqrt=False)
    assert (a.neg_lo, a.neg_hi, a.cfg.flag) == (0.1, 0.45, 0)
    with pytest.raises(TypeError):
        GDSimilarity3D('gwd3d', sqrt=True)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        GDSimilarity3D()(torch.zeros(4, 7), torch.zeros(2, 7))


def test_host_chunk_plan(lib):
    """gd_host_chunk_plan (how gd_loss_fwd_bwd_host cuts rows; pure host arithmetic): the
    chunks tile [0, n) exactly, every start is a multiple of 256 rows, no chunk exceeds
    chunk_rows, a multi-chunk input ends in a tapered tail of at most chunk_rows / 8 rows and
    a single-chunk input stays one launch."""
    import ctypes
    for n, chunk in ((0, 0), (1, 0), (1000, 0), (1 << 20, 0), ((1 << 20) + 1, 0), (1 << 24, 0),
                     ((1 << 24) + 12345, 1 << 20), (5_000_000, 1 << 18), (777, 256), (70_000, 1000)):
        cnt = lib.gd_host_chunk_plan(n, chunk, None, None, 0)
        assert cnt >= 0
        starts = (ctypes.c_int64 * max(cnt, 1))()
        rows = (ctypes.c_int64 * max(cnt, 1))()
        assert lib.gd_host_chunk_plan(n, chunk, ctypes.cast(starts, ctypes.c_void_p),
                                      ctypes.cast(rows, ctypes.c_void_p), cnt) == cnt
        eff = ((chunk if chunk > 0 else 1 << 20) + 255) // 256 * 256
        pos = 0
        for i in range(cnt):
            assert starts[i] == pos and starts[i] % 256 == 0 and 0 < rows[i] <= eff
            pos += rows[i]
        assert pos == n
        if n <= eff:
            assert cnt == (1 if n else 0)
        else:
            assert rows[cnt - 1] <= (eff // 8 + 255) // 256 * 256
            assert cnt <= n // eff + 1 + 4                      # a handful of extra launches
    # capacity too small / bad arguments
    buf = (ctypes.c_int64 * 2)()
    assert lib.gd_host_chunk_plan(1 << 24, 1 << 20, ctypes.cast(buf, ctypes.c_void_p),
                                  ctypes.cast(buf, ctypes.c_void_p), 2) < 0
    assert lib.gd_host_chunk_plan(-1, 0, None, None, 0) < 0
