"""Torch-facing operators over the C ABI (``include/gd_loss_b200.h``).

PyTorch is plumbing here: it owns the device memory, the stream and the autograd
graph edge.  All arithmetic happens in the CUDA library; there is no eager or CPU
fallback -- non-CUDA tensors raise.
"""
import ctypes

import torch

from . import _lib

_WORKSPACES = {}


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
        nbytes = _lib.load().gd_loss_workspace_bytes(0)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _rows7(t):
    """[..., 7] -> [N, 7] fp32 with unit inner stride (row stride free)."""
    t = t.reshape(-1, 7)
    if t.dtype != torch.float32:
        t = t.float()
    if t.shape[0] > 0 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _row_stride(t):
    """Row stride in elements of a [N,7] tensor (7 when there is at most one row)."""
    return t.stride(0) if t.shape[0] > 1 else 7


def _launch(cfg, pred, target, weight, wmode, scale, want_sum, want_rows, want_grad,
            variant, flags=0):
    lib = _lib.load()
    n = pred.shape[0]
    dev = pred.device
    loss = torch.empty((), dtype=torch.float32, device=dev) if want_sum else None
    rows = torch.empty((n,), dtype=torch.float32, device=dev) if want_rows else None
    grad = torch.empty((n, 7), dtype=torch.float32, device=dev) if want_grad else None
    ws = _workspace(dev) if want_sum else None
    wstride = 0
    if wmode == _lib.WEIGHT_ROW:
        wstride = weight.stride(0) if n > 1 else 1
    elif wmode == _lib.WEIGHT_ROW7:
        wstride = _row_stride(weight)
    with _on_device(dev):
        code = lib.gd_loss_fwd_bwd(
            ctypes.byref(cfg), _ptr(pred), _row_stride(pred), _ptr(target),
            _row_stride(target), _ptr(weight), wmode, wstride, n, float(scale),
            _ptr(loss), _ptr(rows), _ptr(grad), _ptr(ws),
            ws.numel() if ws is not None else 0, _lib.VARIANTS[variant], flags,
            _stream_ptr())
    _lib.check(code, 'gd_loss_fwd_bwd')
    return loss, rows, grad


class _GDLossFunction(torch.autograd.Function):
    """Fused forward+backward.  The gradient w.r.t. ``pred`` is produced by the
    forward launch (one HBM pass, 88 B/pair) with every known-at-forward scalar
    folded in; ``backward`` only folds the incoming ``grad_output`` (a kernel that
    exits immediately when it is exactly 1)."""

    @staticmethod
    def forward(ctx, pred, target, weight, cfg, wmode, scale, rows_out, variant, flags):
        need_grad = bool(ctx.needs_input_grad[0])
        loss, rows, grad = _launch(cfg, pred, target, weight, wmode, scale,
                                   not rows_out, rows_out, need_grad, variant, flags)
        ctx.gd = (cfg, wmode, scale, rows_out, variant, flags)
        ctx.grad_buf = grad
        ctx.save_for_backward(pred, target, weight)
        ctx.set_materialize_grads(False)
        return rows if rows_out else loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None or not ctx.needs_input_grad[0]:
            return (None,) * 9
        cfg, wmode, scale, rows_out, variant, flags = ctx.gd
        pred, target, weight = ctx.saved_tensors
        grad = ctx.grad_buf
        ctx.grad_buf = None
        if grad is None:
            # second backward through the same node (retain_graph=True) or grad
            # mode was off at forward time: regenerate with the fused kernel
            _, _, grad = _launch(cfg, pred, target, weight, wmode, scale, False, False,
                                 True, variant, flags)
        lib = _lib.load()
        n = grad.shape[0]
        go = grad_out.detach()
        if go.dtype != torch.float32:
            go = go.float()
        with _on_device(grad.device):
            if rows_out:
                go = go.reshape(-1)
                code = lib.gd_scale_grad_rows(_ptr(grad), n, _ptr(go),
                                              go.stride(0) if n > 1 else 1, _stream_ptr())
                _lib.check(code, 'gd_scale_grad_rows')
            else:
                code = lib.gd_scale_grad(_ptr(grad), n, _ptr(go), _stream_ptr())
                _lib.check(code, 'gd_scale_grad')
        return (grad,) + (None,) * 8


def gd_loss(pred, target, weight, cfg, scale, rows_out=False, variant='auto',
            mask_zero_weight=False):
    """``scale * sum_i w_i loss_i`` (0-dim) or ``scale * w_i * loss_i`` ([N]).

    ``pred``/``target``: ``[..., 7]``; ``weight``: ``None``, ``[N]`` or ``[N,7]``
    (mean over the last dim, reference ``gaussian_distance_loss.py:295-296``)."""
    _require_cuda(pred, 'pred')
    _require_cuda(target, 'target')
    if target.requires_grad:
        raise NotImplementedError(
            'gd_loss_b200: gradients w.r.t. `target` are not produced (the reference '
            'call sites build targets without grad); detach the target')
    out_shape_rows = pred.shape[:-1]
    in_dtype = pred.dtype
    p2 = _rows7(pred)
    t2 = _rows7(target.detach())
    if p2.shape != t2.shape:
        raise ValueError(f'pred {tuple(pred.shape)} and target {tuple(target.shape)} differ')
    n = p2.shape[0]
    wmode, w2 = _lib.WEIGHT_NONE, None
    if weight is not None:
        _require_cuda(weight, 'weight')
        w = weight.detach()
        if w.dtype != torch.float32:
            w = w.float()
        if w.shape == pred.shape:
            wmode, w2 = _lib.WEIGHT_ROW7, w.reshape(-1, 7)
            if n > 0 and w2.stride(1) != 1:
                w2 = w2.contiguous()
        elif w.numel() == n and w.shape == pred.shape[:-1]:
            wmode, w2 = _lib.WEIGHT_ROW, w.reshape(-1)
        else:
            raise ValueError(f'weight shape {tuple(weight.shape)} must be '
                             f'{tuple(pred.shape)} or {tuple(pred.shape[:-1])}')
    flags = _lib.FLAG_MASK_ZERO_WEIGHT if mask_zero_weight else 0
    out = _GDLossFunction.apply(p2, t2, w2, cfg, wmode, scale, rows_out, variant, flags)
    if rows_out and len(out_shape_rows) != 1:
        out = out.reshape(out_shape_rows)
    if in_dtype != torch.float32 and in_dtype.is_floating_point:
        out = out.to(in_dtype)
    return out


def any_positive(weight):
    """``bool(torch.any(weight > 0))`` -- the early-return probe of
    ``GDLoss.forward`` (reference ``gaussian_distance_loss.py:290``); like the
    reference it costs one device->host sync."""
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


class _PositiveProbe:
    """``any(weight > 0)`` in flight: the probe kernel and a copy of its flag into pinned
    host memory are queued, an event marks the copy.  ``result()`` blocks the HOST until that
    event only -- work queued on the stream after the probe (the speculative fused launch)
    keeps the GPU busy meanwhile."""
    __slots__ = ('host', 'event')

    def __init__(self, host, event):
        self.host, self.event = host, event

    def result(self):
        self.event.synchronize()
        return bool(int(self.host[0]))


_PROBE_SLOTS = {}


def any_positive_begin(weight):
    """Start the early-return probe of ``GDLoss.forward`` (reference
    ``gaussian_distance_loss.py:290``) without waiting for it; see ``_PositiveProbe``.
    One pinned slot per (device, stream): the caller must consume the result before it
    starts the next probe on that stream (``GDLoss.forward`` does)."""
    _require_cuda(weight, 'weight')
    w = weight.detach()
    if w.dtype != torch.float32:
        w = w.float()
    if not w.is_contiguous():
        w = w.contiguous()
    dev = w.device
    key = (dev.index, _raw_stream(dev.index))
    slot = _PROBE_SLOTS.get(key)
    if slot is None:
        slot = (torch.empty((1,), dtype=torch.int32, device=dev),
                torch.zeros((1,), dtype=torch.int32).pin_memory(), torch.cuda.Event())
        _PROBE_SLOTS[key] = slot
    flag, host, event = slot
    with _on_device(dev):
        code = _lib.load().gd_any_positive(_ptr(w), w.numel(), _ptr(flag), _stream_ptr())
        _lib.check(code, 'gd_any_positive')
        host.copy_(flag, non_blocking=True)
        event.record()
    return _PositiveProbe(host, event)


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
def _rows_f32(t, name, min_cols=7):
    _require_cuda(t, name)
    if t.dim() != 2 or t.shape[1] < min_cols:
        raise ValueError(f'{name} must be [K,>={min_cols}], got {tuple(t.shape)}')
    if t.dtype != torch.float32:
        t = t.float()
    if t.shape[0] > 0 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _stride0(t):
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


def _fold_grad_output(grad, grad_out):
    """grad *= grad_out (0-dim) on the device; the kernel exits at once when it is 1."""
    go = grad_out.detach()
    if go.dtype != torch.float32:
        go = go.float()
    with _on_device(grad.device):
        code = _lib.load().gd_scale_buffer(_ptr(grad), grad.numel(), _ptr(go), _stream_ptr())
    _lib.check(code, 'gd_scale_buffer')


class _AnchorDecodedLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bbox_pred, anchors, bbox_targets, bbox_weights, decode_weight, pos_inds,
                labels, num_classes, cfg, scale, flags):
        lib = _lib.load()
        dev = bbox_pred.device
        total = bbox_pred.shape[0]
        need_grad = bool(ctx.needs_input_grad[0])
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad, mode = None, _lib.GRAD_NONE
        if need_grad:
            if labels is not None:
                grad, mode = torch.empty((total, 7), dtype=torch.float32, device=dev), _lib.GRAD_DENSE
            else:
                grad, mode = torch.zeros((total, 7), dtype=torch.float32, device=dev), _lib.GRAD_SCATTER
        ws = _workspace(dev)
        dw = None
        if bbox_weights is not None:
            dw = (ctypes.c_float * 7)(*[float(x) for x in decode_weight])
        with _on_device(dev):
            code = lib.gd_anchor_decoded_loss_fwd_bwd(
                ctypes.byref(cfg), _ptr(anchors), anchors.shape[0], _ptr(bbox_pred),
                _stride0(bbox_pred), _ptr(bbox_targets), _stride0(bbox_targets),
                _ptr(bbox_weights), _stride0(bbox_weights) if bbox_weights is not None else 7,
                dw, _ptr(pos_inds), pos_inds.numel() if pos_inds is not None else 0,
                _ptr(labels), int(num_classes or 0), total, float(scale), _ptr(loss), _ptr(grad),
                mode, _ptr(ws), ws.numel(), flags, _stream_ptr())
        _lib.check(code, 'gd_anchor_decoded_loss_fwd_bwd')
        ctx.grad_buf = grad
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None or not ctx.needs_input_grad[0]:
            return (None,) * 11
        grad = ctx.grad_buf
        ctx.grad_buf = None
        if grad is None:
            raise RuntimeError('gd_loss_b200: the fused head loss supports one backward pass '
                               'per forward (retain_graph re-use is not supported)')
        _fold_grad_output(grad, grad_out)
        return (grad,) + (None,) * 10


def anchor_decoded_loss(anchors, bbox_pred, bbox_targets, bbox_weights, decode_weight, cfg,
                        scale, pos_inds=None, labels=None, num_classes=None,
                        mask_zero_weight=True):
    """Fused GD branch of ``GDAnchor3DHead.loss_single``
    (reference ``gd_anchor3d_head.py:102-141``); see ``gd_anchor_decoded_loss_fwd_bwd``
    in ``include/gd_loss_b200.h``.  Returns the 0-dim loss; differentiable w.r.t.
    ``bbox_pred`` (dense ``[T,7]`` gradient, zero off the positives)."""
    if (pos_inds is None) == (labels is None):
        raise ValueError('pass exactly one of pos_inds / labels')
    anchors = _rows_f32(anchors.detach(), 'anchors').contiguous()[:, :7].contiguous()
    bp = _rows_f32(bbox_pred, 'bbox_pred')
    bt = _rows_f32(bbox_targets.detach(), 'bbox_targets')
    if bp.shape[0] != bt.shape[0]:
        raise ValueError('bbox_pred and bbox_targets row counts differ')
    bw = None
    if decode_weight is not None and bbox_weights is not None:
        bw = _rows_f32(bbox_weights.detach(), 'bbox_weights')
        if bw.shape[0] != bp.shape[0]:
            raise ValueError('bbox_weights row count differs from bbox_pred')
        if not hasattr(decode_weight, '__len__'):
            decode_weight = [decode_weight] * 7
        if len(decode_weight) != 7:
            raise ValueError('decode_weight must be a scalar or 7 values')
    if pos_inds is not None:
        _require_cuda(pos_inds, 'pos_inds')
        pos_inds = pos_inds.detach().reshape(-1).to(torch.int64).contiguous()
    else:
        _require_cuda(labels, 'labels')
        if num_classes is None:
            raise ValueError('labels mode needs num_classes')
        labels = labels.detach().reshape(-1).to(torch.int64).contiguous()
        if labels.numel() != bp.shape[0]:
            raise ValueError('labels must have one entry per bbox_pred row')
    flags = _lib.FLAG_MASK_ZERO_WEIGHT if mask_zero_weight else 0
    return _AnchorDecodedLossFunction.apply(bp, anchors, bt, bw, decode_weight, pos_inds,
                                            labels, num_classes, cfg, scale, flags)


class _CenterDecodedLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, locs, target, weight, wmode, coder, cfg, scale, flags):
        lib = _lib.load()
        dev = pred.device
        n, cols = pred.shape
        need_grad = bool(ctx.needs_input_grad[0])
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grad = torch.empty((n, cols), dtype=torch.float32, device=dev) if need_grad else None
        ws = _workspace(dev)
        wstride = 0
        if wmode == _lib.WEIGHT_ROW:
            wstride = weight.stride(0) if n > 1 else 1
        elif wmode == _lib.WEIGHT_ROW7:
            wstride = _stride0(weight)
        with _on_device(dev):
            code = lib.gd_center_decoded_loss_fwd_bwd(
                ctypes.byref(cfg), ctypes.byref(coder), _ptr(pred), _stride0(pred), _ptr(locs),
                locs.stride(0) if n > 1 else 2, _ptr(target), _stride0(target), _ptr(weight),
                wmode, wstride, n, float(scale), _ptr(loss), _ptr(grad), cols, cols, _ptr(ws),
                ws.numel(), flags, _stream_ptr())
        _lib.check(code, 'gd_center_decoded_loss_fwd_bwd')
        ctx.grad_buf = grad
        ctx.set_materialize_grads(False)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None or not ctx.needs_input_grad[0]:
            return (None,) * 9
        grad = ctx.grad_buf
        ctx.grad_buf = None
        if grad is None:
            raise RuntimeError('gd_loss_b200: the fused head loss supports one backward pass '
                               'per forward (retain_graph re-use is not supported)')
        _fold_grad_output(grad, grad_out)
        return (grad,) + (None,) * 8


def center_decoded_loss(pred, pos_ind, target_box, weight, coder, cfg, scale,
                        mask_zero_weight=True):
    """Fused GD branch of ``CenterGDHead.loss`` (reference
    ``gd_centerpoint_head.py:413-434``); see ``gd_center_decoded_loss_fwd_bwd``.
    ``pred`` ``[P,C>=7]`` gathered head outputs, ``pos_ind`` ``[P,3]`` int64
    (batch, x, y), ``target_box`` ``[P,>=7]``.  Differentiable w.r.t. ``pred``."""
    p = _rows_f32(pred, 'pred')
    t = _rows_f32(target_box.detach(), 'target_box')
    _require_cuda(pos_ind, 'pos_ind')
    if pos_ind.dim() != 2 or pos_ind.shape[1] != 3 or pos_ind.shape[0] != p.shape[0]:
        raise ValueError(f'pos_ind must be [{p.shape[0]},3] (batch, x, y)')
    if t.shape[0] != p.shape[0]:
        raise ValueError('pred and target_box row counts differ')
    locs = pos_ind.detach().to(torch.int64)[:, 1:]          # (x_ind, y_ind), a strided view
    if locs.stride(1) != 1:
        locs = locs.contiguous()
    wmode, w2 = _lib.WEIGHT_NONE, None
    if weight is not None:
        _require_cuda(weight, 'weight')
        w = weight.detach().float()
        if w.dim() == 2 and w.shape == (p.shape[0], 7):
            wmode, w2 = _lib.WEIGHT_ROW7, (w if w.stride(1) == 1 else w.contiguous())
        elif w.numel() == p.shape[0]:
            wmode, w2 = _lib.WEIGHT_ROW, w.reshape(-1)
        else:
            raise ValueError('weight must be [P] or [P,7]')
    flags = _lib.FLAG_MASK_ZERO_WEIGHT if mask_zero_weight else 0
    return _CenterDecodedLossFunction.apply(p, locs, t, w2, wmode, coder, cfg, scale, flags)


# ---------------------------------------------------------------------------
# pairwise distances with the assigner reductions fused (f2)
# ---------------------------------------------------------------------------
_PAIR_WORKSPACES = {}


def _pair_workspace(device, m):
    key = (device.index, _raw_stream(device.index))
    need = _lib.load().gd_pairwise_workspace_bytes(m)
    ws = _PAIR_WORKSPACES.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 256 + 8 * 1024), dtype=torch.uint8, device=device)
        _PAIR_WORKSPACES[key] = ws
    return ws


def pairwise_assign(boxes1, boxes2, cfg, want_matrix=False, similarity=False, packed=False):
    """Row and column minima / arg-minima of the pairwise distance matrix in one launch
    (``gd_pairwise_assign``); the matrix itself is written only when ``want_matrix``.
    Returns ``(row_min [N], row_argmin [N] int64, col_min [M], col_argmin [M] int64,
    matrix | None)``.  Indices are -1 / values +inf on an empty axis.  ``packed=True``
    selects the opt-in packed-FP32 kernel (``GD_PAIR_PACKED``)."""
    b1, b2 = _boxes(boxes1, 'boxes1'), _boxes(boxes2, 'boxes2')
    n, m = b1.shape[0], b2.shape[0]
    dev = b1.device
    row_min = torch.empty((n,), dtype=torch.float32, device=dev)
    row_idx = torch.empty((n,), dtype=torch.int32, device=dev)
    col_min = torch.empty((m,), dtype=torch.float32, device=dev)
    col_idx = torch.empty((m,), dtype=torch.int32, device=dev)
    mat = torch.empty((n, m), dtype=torch.float32, device=dev) if want_matrix else None
    if n == 0 or m == 0:
        row_min.fill_(float('inf'))
        col_min.fill_(float('inf'))
        row_idx.fill_(-1)
        col_idx.fill_(-1)
        return row_min, row_idx.long(), col_min, col_idx.long(), mat
    ws = _pair_workspace(dev, m)
    with _on_device(dev):
        code = _lib.load().gd_pairwise_assign(
            ctypes.byref(cfg), _ptr(b1), n, _ptr(b2), m, _ptr(row_min), _ptr(row_idx),
            _ptr(col_min), _ptr(col_idx), _ptr(mat), m,
            (_lib.PAIR_SIMILARITY if similarity else 0) | (_lib.PAIR_PACKED if packed else 0),
            _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(code, 'gd_pairwise_assign')
    return row_min, row_idx.long(), col_min, col_idx.long(), mat


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
