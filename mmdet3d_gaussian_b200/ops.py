"""Torch-facing operators over the C ABI (``include/gd_loss_b200.h``).

PyTorch is plumbing here: it owns the device memory, the stream and the autograd
graph edge.  All arithmetic happens in the CUDA library; there is no eager or CPU
fallback -- non-CUDA tensors raise.  The loss and the head front ends go through the torch
C++ extension (``csrc/torch_shim.cpp`` -> ``_C.so``: a few microseconds of host time per
call); the pairwise / assignment surface, which is not latency critical, binds the same C
ABI with ctypes.
"""
import ctypes

import torch

from . import _lib

_WORKSPACES = {}
_MAX_WORKSPACES = 256


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor, got {type(t)}')
    if not t.is_cuda:
        raise RuntimeError(
            f'gd_loss_b200: {name} is on {t.device}; this implementation is CUDA-only '
            f'(sm_100a) and has no CPU fallback')


def _raw_stream(index=None):
    """cudaStream_t of torch's current stream on device ``index`` as an int (the raw
    getter: ``torch.cuda.current_stream()`` costs ~10 us of Python per call, which is more
    than the kernel launch itself at training sizes)."""
    if index is None:
        index = torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(index)


def _stream_ptr():
    return ctypes.c_void_p(_raw_stream())


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current."""
    __slots__ = ('ctx',)

    def __init__(self, dev):
        self.ctx = None if dev.index == torch._C._cuda_getDevice() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _workspace(device):
    """Zero-initialised scratch (ticket + per-CTA partials), one per (device, stream).
    The kernels leave it zeroed, so it is cleared exactly once."""
    key = (device.index, _raw_stream(device.index))
    ws = _WORKSPACES.get(key)
    if ws is None:
        if len(_WORKSPACES) >= _MAX_WORKSPACES:      # bounded: start over
            _WORKSPACES.clear()
        nbytes = _lib.load().gd_loss_workspace_bytes(0)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


REDUCTIONS = {'none': 0, 'mean': 1, 'sum': 2}
SYNC_MASK_ZERO, SYNC_EXACT = 0, 1


def gd_loss(pred, target, weight, cfg, loss_weight=1.0, reduction='mean', avg_factor=None,
            variant='auto', mask_zero_weight=False, peer_sum=None):
    """``GDLoss.forward`` after its Python-only steps: the torch C++ shim
    (``csrc/torch_shim.cpp``) flattens ``[..., 7]``, dispatches the weight shape
    (``None`` / ``[N]`` / ``[N,7]``: mean over the last dim, reference
    ``gaussian_distance_loss.py:295-296``), folds mmdet's ``weight_reduce_loss`` scalars,
    allocates the outputs, launches the fused kernel on the current stream and hooks the
    gradient into autograd.  ``cfg``: ``_lib.make_shim_config(...)``.
    ``mask_zero_weight=False`` keeps the reference's early return (ref:290-292), decided on
    the device where its result is shape-valid; ``True`` never probes and masks rows whose
    weight is exactly 0.  ``peer_sum`` (``sharded.PeerSumContext``): the reduced loss is summed
    over the GPUs of the box inside the launch."""
    return _lib.shim().gd_loss(pred, target, weight, cfg, float(loss_weight),
                               REDUCTIONS[reduction], avg_factor, _lib.VARIANTS[variant],
                               SYNC_MASK_ZERO if mask_zero_weight else SYNC_EXACT, peer_sum)


def any_positive(weight):
    """``bool(torch.any(weight > 0))`` -- the early-return probe of
    ``GDLoss.forward`` (reference ``gaussian_distance_loss.py:290``) as a blocking call
    (one device->host sync; the module itself never uses this form)."""
    _require_cuda(weight, 'weight')
    w = weight.detach()
    if w.dtype != torch.float32:
        w = w.float()
    if not w.is_contiguous():
        w = w.contiguous()
    flag = torch.empty((1,), dtype=torch.int32, device=w.device)
    with _on_device(w.device):
        code = _lib.load().gd_any_positive(_ptr(w), w.numel(), _ptr(flag), _stream_ptr())
    _lib.check(code, 'gd_any_positive')
    return bool(flag.item())


def _boxes(t, name):
    _require_cuda(t, name)
    if t.dim() != 2 or t.shape[1] != 7:
        raise ValueError(f'{name} must be [K,7], got {tuple(t.shape)}')
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def pairwise_distance(boxes1, boxes2, cfg, out=None):
    """``[N,M]`` matrix ``out[i,j] = post(distance(boxes1[i], boxes2[j]))``."""
    b1, b2 = _boxes(boxes1, 'boxes1'), _boxes(boxes2, 'boxes2')
    n, m = b1.shape[0], b2.shape[0]
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=b1.device)
    with _on_device(b1.device):
        code = _lib.load().gd_pairwise(ctypes.byref(cfg), _ptr(b1), n, _ptr(b2), m,
                                       _ptr(out), out.stride(0) if n > 1 else max(m, 1),
                                       _stream_ptr())
    _lib.check(code, 'gd_pairwise')
    return out


def pairwise_row_argmin(boxes1, boxes2, cfg):
    """Per row of the (never materialised) matrix: ``(min_j, argmin_j)``."""
    b1, b2 = _boxes(boxes1, 'boxes1'), _boxes(boxes2, 'boxes2')
    n, m = b1.shape[0], b2.shape[0]
    if m == 0:
        raise ValueError('pairwise_row_argmin needs at least one column box')
    vmin = torch.empty((n,), dtype=torch.float32, device=b1.device)
    idx = torch.empty((n,), dtype=torch.int32, device=b1.device)
    with _on_device(b1.device):
        code = _lib.load().gd_pairwise_row_argmin(ctypes.byref(cfg), _ptr(b1), n, _ptr(b2), m,
                                                  _ptr(vmin), _ptr(idx), _stream_ptr())
    _lib.check(code, 'gd_pairwise_row_argmin')
    return vmin, idx.long()


def launch_count():
    return int(_lib.load().gd_launch_count())


# ---------------------------------------------------------------------------
# head front ends: gather + decode + loss + gradient to the raw outputs (f1)
# ---------------------------------------------------------------------------
def anchor_decoded_loss(anchors, bbox_pred, bbox_targets, bbox_weights, decode_weight, cfg,
                        loss_weight, scale_mode=0, avg_factor=None, pos_inds=None, labels=None,
                        num_classes=None, mask_zero_weight=True):
    """Fused GD branch of ``GDAnchor3DHead.loss_single``
    (reference ``gd_anchor3d_head.py:102-141``); see ``gd_anchor_decoded_loss_fwd_bwd``
    in ``include/gd_loss_b200.h``.  Returns the 0-dim loss; differentiable w.r.t.
    ``bbox_pred`` (dense ``[T,7]`` gradient, zero off the positives).
    ``scale_mode``: 0 -> the loss is scaled by ``loss_weight`` as given; 1 -> by
    ``loss_weight / avg_factor`` (number or 0-dim CUDA tensor, no sync); 2 (labels mode) -> by
    ``loss_weight / max(#positives, 1)`` counted on the device."""
    if (pos_inds is None) == (labels is None):
        raise ValueError('pass exactly one of pos_inds / labels')
    if labels is not None and num_classes is None:
        raise ValueError('labels mode needs num_classes')
    dw = None
    if decode_weight is not None and bbox_weights is not None:
        dw = ([float(decode_weight)] * 7 if not hasattr(decode_weight, '__len__')
              else [float(x) for x in decode_weight])
        if len(dw) != 7:
            raise ValueError('decode_weight must be a scalar or 7 values')
    return _lib.shim().anchor_decoded_loss(
        anchors, bbox_pred, bbox_targets, bbox_weights if dw is not None else None, dw, pos_inds,
        labels, int(num_classes or 0), cfg, float(loss_weight), int(scale_mode), avg_factor,
        bool(mask_zero_weight))


def center_decoded_loss(pred, pos_ind, target_box, weight, coder, cfg, loss_weight, scale_mode=0,
                        avg_factor=None, mask_zero_weight=True):
    """Fused GD branch of ``CenterGDHead.loss`` (reference
    ``gd_centerpoint_head.py:413-434``); see ``gd_center_decoded_loss_fwd_bwd``.
    ``pred`` ``[P,C>=7]`` gathered head outputs, ``pos_ind`` ``[P,3]`` int64
    (batch, x, y), ``target_box`` ``[P,>=7]``.  Differentiable w.r.t. ``pred``.
    ``scale_mode``: 0 / 1 as ``anchor_decoded_loss``; 2 -> ``loss_weight / P``."""
    return _lib.shim().center_decoded_loss(pred, pos_ind, target_box, weight, coder, cfg,
                                           float(loss_weight), int(scale_mode), avg_factor,
                                           bool(mask_zero_weight))


# ---------------------------------------------------------------------------
# pairwise distances with the assigner reductions fused (f2)
# ---------------------------------------------------------------------------
_PAIR_WORKSPACES = {}
_MAX_WORKSPACES = 256


def _pair_workspace(device, m):
    key = (device.index, _raw_stream(device.index))
    need = _lib.load().gd_pairwise_workspace_bytes(m)
    ws = _PAIR_WORKSPACES.get(key)
    if ws is None and len(_PAIR_WORKSPACES) >= _MAX_WORKSPACES:
        _PAIR_WORKSPACES.clear()
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 256 + 8 * 1024), dtype=torch.uint8, device=device)
        _PAIR_WORKSPACES[key] = ws
    return ws


def pairwise_assign(boxes1, boxes2, cfg, want_matrix=False, similarity=False, cpl1=False,
                    index64=True):
    """Row and column minima / arg-minima of the pairwise distance matrix in one launch
    (``gd_pairwise_assign``); the matrix itself is written only when ``want_matrix``.
    Returns ``(row_min [N], row_argmin [N] int64, col_min [M], col_argmin [M] int64,
    matrix | None)``.  Indices are -1 / values +inf on an empty axis.  ``cpl1=True`` pins the
    one-column-per-lane mapping (``GD_PAIR_CPL1``: same arithmetic, bit-identical results).
    The kernel writes the indices as int64 itself (``GD_PAIR_INDEX64``); ``index64=False`` keeps
    the C ABI's int32 (what ``assign_from_minima`` consumes: no conversion launch either way)."""
    b1, b2 = _boxes(boxes1, 'boxes1'), _boxes(boxes2, 'boxes2')
    n, m = b1.shape[0], b2.shape[0]
    dev = b1.device
    row_min = torch.empty((n,), dtype=torch.float32, device=dev)
    idt = torch.int64 if index64 else torch.int32
    row_idx = torch.empty((n,), dtype=idt, device=dev)             # GD_PAIR_INDEX64: no conversion
    col_min = torch.empty((m,), dtype=torch.float32, device=dev)
    col_idx = torch.empty((m,), dtype=idt, device=dev)
    mat = torch.empty((n, m), dtype=torch.float32, device=dev) if want_matrix else None
    if n == 0 or m == 0:
        row_min.fill_(float('inf'))
        col_min.fill_(float('inf'))
        row_idx.fill_(-1)
        col_idx.fill_(-1)
        return row_min, row_idx, col_min, col_idx, mat
    ws = _pair_workspace(dev, m)
    with _on_device(dev):
        code = _lib.load().gd_pairwise_assign(
            ctypes.byref(cfg), _ptr(b1), n, _ptr(b2), m, _ptr(row_min), _ptr(row_idx),
            _ptr(col_min), _ptr(col_idx), _ptr(mat), m,
            (_lib.PAIR_SIMILARITY if similarity else 0) | (_lib.PAIR_CPL1 if cpl1 else 0) |
            (_lib.PAIR_INDEX64 if index64 else 0),
            _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(code, 'gd_pairwise_assign')
    return row_min, row_idx, col_min, col_idx, mat


def assign_from_minima(row_min, row_argmin, col_min, col_argmin, pos_thr, neg_lo, neg_hi,
                       min_pos, match_low_quality=True):
    """``(assigned_gt_inds [N] int64, max_overlaps [N])`` -- ``gd_assign_from_minima``."""
    n, m = row_min.shape[0], col_min.shape[0]
    dev = row_min.device
    assigned = torch.empty((n,), dtype=torch.int64, device=dev)
    max_ov = torch.empty((n,), dtype=torch.float32, device=dev)
    ri = row_argmin.to(torch.int32)
    ci = col_argmin.to(torch.int32)
    with _on_device(dev):
        code = _lib.load().gd_assign_from_minima(
            _ptr(row_min), _ptr(ri), n, _ptr(col_min), _ptr(ci), m, float(pos_thr),
            float(neg_lo), float(neg_hi), float(min_pos), 1 if match_low_quality else 0,
            _ptr(assigned), _ptr(max_ov), _stream_ptr())
    _lib.check(code, 'gd_assign_from_minima')
    return assigned, max_ov


# ---------------------------------------------------------------------------
# SimOTA-style consumer: column top-k + dynamic-k matching, no matrix (f2)
# ---------------------------------------------------------------------------
_TOPK_WORKSPACES = {}


def pairwise_col_topk(boxes1, boxes2, cfg, k, want_matrix=False):
    """``(row_min [N], row_argmin [N] int64, topk_val [k,M], topk_row [k,M] int64, matrix | None)``:
    per column of the distance matrix its ``k`` smallest entries, ascending (NaN first, ties ->
    lowest row; -1 / +inf where N < k), and per row its minimum -- ``gd_pairwise_col_topk``.  The
    matrix is only written when ``want_matrix`` (same instruction sequence: bit-consistent)."""
    b1, b2 = _boxes(boxes1, 'boxes1'), _boxes(boxes2, 'boxes2')
    n, m = b1.shape[0], b2.shape[0]
    if m == 0:
        raise ValueError('pairwise_col_topk needs at least one column box')
    if not 1 <= int(k) <= 16:
        raise ValueError('k must be in [1, 16]')
    dev = b1.device
    lib = _lib.load()
    row_min = torch.empty((n,), dtype=torch.float32, device=dev)
    row_idx = torch.empty((n,), dtype=torch.int32, device=dev)
    val = torch.empty((int(k), m), dtype=torch.float32, device=dev)
    row = torch.empty((int(k), m), dtype=torch.int32, device=dev)
    mat = torch.empty((n, m), dtype=torch.float32, device=dev) if want_matrix else None
    need = lib.gd_pairwise_topk_workspace_bytes(n, m)
    key = (dev.index, _raw_stream(dev.index))
    ws = _TOPK_WORKSPACES.get(key)
    if ws is None or ws.numel() < need:
        if len(_TOPK_WORKSPACES) >= _MAX_WORKSPACES:
            _TOPK_WORKSPACES.clear()
        ws = _TOPK_WORKSPACES[key] = torch.empty(need, dtype=torch.uint8, device=dev)
    with _on_device(dev):
        code = lib.gd_pairwise_col_topk(ctypes.byref(cfg), _ptr(b1), n, _ptr(b2), m, int(k),
                                        _ptr(row_min), _ptr(row_idx), _ptr(val), _ptr(row),
                                        _ptr(mat), m, _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(code, 'gd_pairwise_col_topk')
    return row_min, row_idx.long(), val, row.long(), mat


def simota_from_topk(topk_val, topk_row, row_min, row_argmin, unmatched_sim=-1e8):
    """SimOTA ``dynamic_k_matching`` (reference ``sim_ota_3d_assigner.py:184-211``) from the
    column top-k lists: ``(assigned_gt_inds [N] int64 (0 = background, else GT + 1),
    matched_sim [N], dynamic_ks [M] int64)`` -- ``gd_simota_from_topk``."""
    k, m = topk_val.shape
    n = row_min.shape[0]
    dev = row_min.device
    assigned = torch.empty((n,), dtype=torch.int64, device=dev)
    sim = torch.empty((n,), dtype=torch.float32, device=dev)
    dks = torch.empty((m,), dtype=torch.int32, device=dev)
    scratch = torch.empty((3 * max(n, 1),), dtype=torch.int32, device=dev)
    tr = topk_row.to(torch.int32).contiguous()
    ri = row_argmin.to(torch.int32)
    with _on_device(dev):
        code = _lib.load().gd_simota_from_topk(
            _ptr(topk_val.contiguous()), _ptr(tr), m, k, _ptr(row_min), _ptr(ri), n, _ptr(assigned),
            _ptr(sim), _ptr(dks), float(unmatched_sim), _ptr(scratch), _stream_ptr())
    _lib.check(code, 'gd_simota_from_topk')
    return assigned, sim, dks.long()
