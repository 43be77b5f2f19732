#!/usr/bin/env python
"""Interleaved A/B of the fused kernel's grid size for the layouts the headline does not cover
([N,7] weights; row-strided views; 28-byte-offset slices; reduction='none'), 2^24 pairs, kld3d and
bd3d, `gd_set_loss_grid` on the production library: ROUNDS rounds, every grid in every round."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, ops, synth  # noqa: E402

GRIDS = [148, 140, 136, 132, 128, 124, 120]
ROUNDS = 4


def main():
    n = 1 << 24
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

    def make(lt, p, ps, t, ts, wt, wmode, wstride, rows_out):
        cfg = _lib.make_config(lt, 'log1p', True, 0.0, 1.0, (0, 0, 0.5))
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
        io.variant = _lib.VARIANTS['auto']

        def launch():
            code = lib.gd_loss_launch(ctypes.byref(cfg), ctypes.byref(io), stream)
            if code != 0:
                raise RuntimeError(code)
        launch.keep = (cfg, io)
        return launch
    cases = {}
    for lt in ('kld3d', 'bd3d'):
        cases[f'[N,7] weights {lt}'] = (make(lt, pred, 7, target, 7, w7, 2, 7, False), 112)
        cases[f'strided 9/11 {lt}'] = (make(lt, wide_p, 9, wide_t, 11, None, 0, 0, False), 108)
        cases[f'28-byte offset {lt}'] = (make(lt, pred[1:], 7, target[1:], 7, w[1:], 1, 1, False), 88)
        cases[f"'none' {lt}"] = (make(lt, pred, 7, target, 7, w, 1, 1, True), 92)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms = {g: {c: [] for c in cases} for g in GRIDS}
    for r in range(ROUNDS):
        for g in (GRIDS if r % 2 == 0 else GRIDS[::-1]):
            lib.gd_set_loss_grid(g)
            for c, (fn, _) in cases.items():
                ms[g][c].append(timed(fn))
    out = {}
    for c, (_, bpp) in cases.items():
        out[c] = {g: round(bpp * n / (sum(ms[g][c]) / ROUNDS) / 1e6, 1) for g in GRIDS}
        sys.stderr.write(f'{c:26s} ' + ' '.join(f'{g}:{out[c][g]:.0f}' for g in GRIDS) + '\n')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
