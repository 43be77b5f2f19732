"""Bindings of ``libgdloss_b200.so`` -- the C ABI in ``include/gd_loss_b200.h``.

Two binders over the same exported C symbols: ``shim()`` is the torch C++ extension
(``csrc/torch_shim.cpp`` -> ``_C.so``) that the loss module and the head front ends call,
``load()`` is a ctypes handle used by the pairwise / assignment surface, the tools and the
tests that exercise the C ABI directly.

There is no fallback of any kind: if a binary is missing (and cannot be
built because nvcc / g++ is absent) or a call fails, this raises.  Tensors are passed
as raw device pointers; work is enqueued on torch's current CUDA stream.
"""
import ctypes
import os

from . import build_ext

LOSS_TYPES = {'gwd3d': 0, 'kld3d': 1, 'jd3d': 2, 'kld3d_symmax': 3,
              'kld3d_symmin': 4, 'bd3d': 5, 'kfiou3d': 6}
FUNS = {'none': 0, 'log1p': 1, 'expm1': 2, 'nlog': 3}
WEIGHT_NONE, WEIGHT_ROW, WEIGHT_ROW7 = 0, 1, 2
VARIANTS = {'auto': 0, 'staged': 1, 'bulk': 2, 'bulk_r2': 3, 'bulk_packed': 4, 'bulk_any': 5}
FLAG_MASK_ZERO_WEIGHT = 1
ABI_VERSION = 3
PAIR_SIMILARITY = 1
PAIR_CPL1 = 2              # one column per lane (measurement / test aid)
PAIR_INDEX64 = 4           # index outputs are int64 arrays
GRAD_NONE, GRAD_COMPACT, GRAD_SCATTER, GRAD_DENSE = 0, 1, 2, 3


class GDLossConfig(ctypes.Structure):
    """``struct gd_loss_config``."""
    _fields_ = [('loss_type', ctypes.c_int32), ('fun', ctypes.c_int32),
                ('flag', ctypes.c_int32), ('tau', ctypes.c_float),
                ('alpha', ctypes.c_float), ('center_offset', ctypes.c_float * 3)]


MAX_PEERS = 16


class GDPeerSum(ctypes.Structure):
    """``struct gd_peer_sum``."""
    _fields_ = [('world', ctypes.c_int32), ('rank', ctypes.c_int32),
                ('peer_buf', ctypes.c_void_p * MAX_PEERS)]


class GDLossIO(ctypes.Structure):
    """``struct gd_loss_io``."""
    _fields_ = [('pred', ctypes.c_void_p), ('pred_row_stride', ctypes.c_int64),
                ('target', ctypes.c_void_p), ('target_row_stride', ctypes.c_int64),
                ('weight', ctypes.c_void_p), ('weight_mode', ctypes.c_int32),
                ('weight_row_stride', ctypes.c_int64), ('n', ctypes.c_int64),
                ('scale', ctypes.c_float), ('scale_div', ctypes.c_void_p),
                ('loss_sum', ctypes.c_void_p), ('row_loss', ctypes.c_void_p),
                ('grad_pred', ctypes.c_void_p), ('status', ctypes.c_void_p),
                ('early_return', ctypes.c_int32), ('er_weight_row_stride', ctypes.c_int64),
                ('er_weight_col_stride', ctypes.c_int64),
                ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_size_t),
                ('variant', ctypes.c_int32), ('flags', ctypes.c_int32),
                ('peer_sum', ctypes.POINTER(GDPeerSum)),
                ('any_positive_host', ctypes.c_void_p)]


class GDCenterCoder(ctypes.Structure):
    """``struct gd_center_coder``."""
    _fields_ = [('pc_range', ctypes.c_double * 2), ('voxel_size', ctypes.c_double * 2),
                ('out_size_factor', ctypes.c_int32), ('norm_bbox', ctypes.c_int32)]


_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float
_cfgp = ctypes.POINTER(GDLossConfig)

# name -> (restype, argtypes); must list every symbol include/gd_loss_b200.h declares
SIGNATURES = {
    'gd_abi_version': (ctypes.c_int, []),
    'gd_loss_workspace_bytes': (ctypes.c_size_t, [_i64]),
    'gd_loss_fwd_bwd': (ctypes.c_int, [_cfgp, _vp, _i64, _vp, _i64, _vp, _i32, _i64,
                                       _i64, _f32, _vp, _vp, _vp, _vp,
                                       ctypes.c_size_t, _i32, _i32, _vp]),
    'gd_loss_launch': (ctypes.c_int, [_cfgp, ctypes.POINTER(GDLossIO), _vp]),
    'gd_peer_sum_buffer_bytes': (ctypes.c_size_t, []),
    'gd_host_flag_wait': (ctypes.c_int, [_vp, _vp]),
    'gd_count_positive_labels': (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, ctypes.c_size_t, _vp]),
    'gd_scale_grad': (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    'gd_scale_buffer': (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    'gd_scale_grad_rows': (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp]),
    'gd_any_positive': (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    'gd_pairwise': (ctypes.c_int, [_cfgp, _vp, _i64, _vp, _i64, _vp, _i64, _vp]),
    'gd_pairwise_row_argmin': (ctypes.c_int, [_cfgp, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    'gd_pairwise_workspace_bytes': (ctypes.c_size_t, [_i64]),
    'gd_pairwise_assign': (ctypes.c_int, [_cfgp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                                          _i64, _i32, _vp, ctypes.c_size_t, _vp]),
    'gd_pairwise_topk_workspace_bytes': (ctypes.c_size_t, [_i64, _i64]),
    'gd_pairwise_col_topk': (ctypes.c_int, [_cfgp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp,
                                            _vp, _i64, _vp, ctypes.c_size_t, _vp]),
    'gd_simota_from_topk': (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp,
                                           _f32, _vp, _vp]),
    'gd_assign_from_minima': (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _f32, _f32, _f32, _f32,
                                             _i32, _vp, _vp, _vp]),
    'gd_anchor_decoded_loss_fwd_bwd': (ctypes.c_int, [
        _cfgp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, ctypes.POINTER(ctypes.c_float),
        _vp, _i64, _vp, _i64, _i64, _f32, _vp, _vp, _vp, _i32, _vp, ctypes.c_size_t, _i32, _vp]),
    'gd_center_decoded_loss_fwd_bwd': (ctypes.c_int, [
        _cfgp, ctypes.POINTER(GDCenterCoder), _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32,
        _i64, _i64, _f32, _vp, _vp, _vp, _i64, _i32, _vp, ctypes.c_size_t, _i32, _vp]),
    'gd_loss_fwd_bwd_host': (ctypes.c_int, [_cfgp, _vp, _vp, _vp, _i32, _i64, _f32,
                                            _vp, _vp, _i32, _i64]),
    'gd_host_chunk_plan': (_i64, [_i64, _i64, _vp, _vp, _i64]),
    'gd_launch_count': (_i64, []),
    'gd_set_loss_grid': (ctypes.c_int, [_i32]),
    'gd_error_string': (ctypes.c_char_p, [ctypes.c_int]),
}

_LIB = None
_LIB_PATH = None


def load(path=None):
    """Load (building first if needed) the shared library; cached."""
    global _LIB, _LIB_PATH
    if _LIB is not None and path is None:
        return _LIB
    if path is None:
        path = os.environ.get('GD_LOSS_B200_LIB')
    if path is None:
        path = build_ext.lib_path()
        if not build_ext.is_current():
            # build() raises if nvcc is unavailable: no silent fallback
            path = build_ext.build()
    if not os.path.exists(path):
        raise RuntimeError(f'gd_loss_b200: native library {path} is missing; run '
                           f'`python -m mmdet3d_gaussian_b200.build_ext`')
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gd_abi_version() != ABI_VERSION:
        raise RuntimeError('gd_loss_b200: ABI version mismatch')
    _LIB, _LIB_PATH = lib, path
    return lib


def loaded_path():
    return _LIB_PATH


_SHIM = None


def shim():
    """The torch C++ extension, bound to the same library file as ``load()``; cached."""
    global _SHIM
    if _SHIM is None:
        import importlib.util
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        load()
        path = build_ext.shim_path()
        if not build_ext.shim_is_current():
            path = build_ext.build_shim()       # raises if g++ is unavailable
        spec = importlib.util.spec_from_file_location(__package__ + '._C', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.bind(_LIB_PATH)
        _SHIM = mod
    return _SHIM


def make_shim_config(loss_type, fun, flag, tau, alpha, center_offset):
    """``gd_loss_config`` held by the C++ shim (same fields as ``make_config``)."""
    return shim().LossConfig(LOSS_TYPES[loss_type], FUNS[fun], bool(flag), float(tau), float(alpha),
                             [float(x) for x in center_offset])


def make_shim_center_coder(pc_range, out_size_factor, voxel_size, norm_bbox=True):
    """``gd_center_coder`` held by the C++ shim."""
    return shim().CenterCoder([float(x) for x in pc_range[:2]], int(out_size_factor),
                              [float(x) for x in voxel_size[:2]], bool(norm_bbox))


def check(code, what):
    if code != 0:
        msg = load().gd_error_string(code).decode()
        raise RuntimeError(f'gd_loss_b200.{what} failed ({code}): {msg}')


def make_center_coder(pc_range, out_size_factor, voxel_size, norm_bbox=True):
    c = GDCenterCoder()
    for i in range(2):
        c.pc_range[i] = float(pc_range[i])
        c.voxel_size[i] = float(voxel_size[i])
    c.out_size_factor = int(out_size_factor)
    c.norm_bbox = 1 if norm_bbox else 0
    return c


def make_config(loss_type, fun, flag, tau, alpha, center_offset):
    cfg = GDLossConfig()
    cfg.loss_type = LOSS_TYPES[loss_type]
    cfg.fun = FUNS[fun]
    cfg.flag = 1 if flag else 0
    cfg.tau = float(tau)
    cfg.alpha = float(alpha)
    for i in range(3):
        cfg.center_offset[i] = float(center_offset[i])
    return cfg
