#!/usr/bin/env python
"""Kernel-only GB/s of the fused loss at 2^24 pairs for the layouts the headline does not
cover: [N,7] weights (112 B/pair), reduction='none' (92 B/pair),
row-strided `[..., :7]` views of 9- / 11-wide rows (CenterGDHead, gd_centerpoint_head.py:413-423;
bytes = what the kernel TOUCHES: whole wide rows in, 28 B gradient out) and 28-byte-offset
slices, each for 'auto' and for the staged kernel.  CUDA events around bare C-ABI launches.
One JSON document on stdout."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, ops, synth  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    torch.cuda.set_device(0)
    lib = _lib.load()
    pred, target, w = synth.make_pairs(n + 8, 'nuscenes', seed=0, device='cuda', weights='bernoulli')
    w7 = w[:, None].expand(n + 8, 7).contiguous()
    wide_p = torch.zeros(n + 8, 9, device='cuda')
    wide_p[:, :7] = pred
    wide_t = torch.zeros(n + 8, 11, device='cuda')
    wide_t[:, :7] = target
    grad = torch.empty(n + 8, 7, device='cuda')
    rows = torch.empty(n + 8, device='cuda')
    loss = torch.empty((), device='cuda')
    ws = ops._workspace(pred.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {'n': n, 'cases': []}

    def run(name, lt, p, ps, t, ts, wt, wmode, wstride, rows_out, bytes_per_pair, variants):
        cfg = _lib.make_config(lt, 'log1p', True, 0.0, 1.0, (0, 0, 0.5))
        rec = {'case': name, 'loss': lt, 'bytes_per_pair': bytes_per_pair}
        for v in variants:
            io = _lib.GDLossIO()
            io.pred, io.pred_row_stride = p.data_ptr(), ps
            io.target, io.target_row_stride = t.data_ptr(), ts
            io.weight = wt.data_ptr() if wt is not None else None
            io.weight_mode, io.weight_row_stride = wmode, wstride
            io.n, io.scale = n, 5.0 / n
            io.loss_sum = None if rows_out else loss.data_ptr()
            io.row_loss = rows.data_ptr() if rows_out else None
            io.grad_pred = grad.data_ptr()
            io.workspace, io.workspace_bytes = ws.data_ptr(), ws.numel()
            io.variant = _lib.VARIANTS[v]

            def launch():
                code = lib.gd_loss_launch(ctypes.byref(cfg), ctypes.byref(io), stream)
                if code != 0:
                    raise RuntimeError(f'{name}/{v}: {code}')
            try:
                ms = timed(launch)
            except RuntimeError as e:
                rec[v] = str(e)
                continue
            rec[v] = {'ms': round(ms, 4), 'GBps': round(bytes_per_pair * n / ms / 1e6, 1)}
        sys.stderr.write(json.dumps(rec) + '\n')
        out['cases'].append(rec)

    for lt in ('gwd3d', 'kld3d', 'bd3d'):
        run('contiguous [N] weights', lt, pred, 7, target, 7, w, 1, 1, False, 88, ('auto', 'bulk', 'staged'))
        run('contiguous [N,7] weights', lt, pred, 7, target, 7, w7, 2, 7, False, 112, ('auto', 'staged'))
        run("reduction='none' [N] weights", lt, pred, 7, target, 7, w, 1, 1, True, 92, ('auto', 'staged'))
        run('strided 9/11-wide rows, no weight', lt, wide_p, 9, wide_t, 11, None, 0, 0, False,
            36 + 44 + 28, ('auto', 'staged'))
        run('28-byte offset slices [N] weights', lt, pred[1:], 7, target[1:], 7, w[1:], 1, 1, False, 88,
            ('auto', 'staged'))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
