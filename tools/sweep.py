#!/usr/bin/env python
"""Config C5 sweep: kernel-only fused fwd+bwd time for
N in 2^10..2^26 x {gwd3d,kld3d,bd3d} (+ siblings at 2^24), both kernel variants,
CUDA events around bare C-ABI launches.  Also times the pairwise C4 shape.
Prints one JSON document."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, ops, synth  # noqa: E402


def time_launch(fn, reps):
    for _ in range(3):
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
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', choices=['all', 'elementwise', 'pairwise'], default='all')
    ap.add_argument('--max-log2n', type=int, default=26)
    ap.add_argument('--packed', '--cpl1', dest='packed', action='store_true',
                    help='also time the one-column-per-lane mapping of the pairwise kernel (GD_PAIR_CPL1)')
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    nmax = 1 << (args.max_log2n if args.only != 'pairwise' else 10)
    pred, target, weight = synth.make_pairs(nmax, 'kitti', seed=0, device=dev)
    grad = torch.empty(nmax, 7, device=dev)
    rows = torch.empty(nmax, device=dev)
    loss = torch.empty((), device=dev)
    ws = ops._workspace(dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {'elementwise': [], 'pairwise': []}
    types = ['gwd3d', 'kld3d', 'bd3d']
    for e in ([] if args.only == 'pairwise' else list(range(10, args.max_log2n + 1, 2))):
        n = 1 << e
        for lt in types + (['jd3d', 'kld3d_symmax', 'kld3d_symmin', 'kfiou3d'] if e == 24 else []):
            fun = 'none' if lt == 'kfiou3d' else 'log1p'
            cfg = _lib.make_config(lt, fun, True, 0.0, 1.0, (0, 0, 0.5))
            for variant in ('bulk', 'bulk_packed', 'bulk_r2', 'staged'):
                for mode, g, r, bpp in (('fwd+bwd', grad, None, 88), ('fwd', None, None, 60),
                                        ('fwd+bwd+rows', grad, rows, 92)):
                    if mode != 'fwd+bwd' and (e != 24 or lt not in types):
                        continue

                    def launch():
                        code = lib.gd_loss_fwd_bwd(
                            ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7,
                            weight.data_ptr(), 1, 1, n, 5.0 / n, loss.data_ptr(),
                            r.data_ptr() if r is not None else None,
                            g.data_ptr() if g is not None else None,
                            ws.data_ptr(), ws.numel(), _lib.VARIANTS[variant], 0, stream)
                        _lib.check(code, 'gd_loss_fwd_bwd')
                    ms = time_launch(launch, 50 if e <= 20 else 20)
                    out['elementwise'].append({
                        'n': n, 'log2n': e, 'loss': lt, 'variant': variant, 'mode': mode,
                        'ms': round(ms, 5), 'Gpairs_per_s': round(n / ms / 1e6, 3),
                        'GBps': round(bpp * n / ms / 1e6, 1)})
    # pairwise C4: 200k anchors x 256 GT
    anchors = synth.make_anchor_grid(200_000, 'waymo', device=dev)
    gts = synth.make_targets(256, 'waymo', seed=5, device=dev)
    mat = torch.empty(200_000, 256, device=dev)
    vmin = torch.empty(200_000, device=dev)
    idx = torch.empty(200_000, dtype=torch.int32, device=dev)
    for lt in ([] if args.only == 'elementwise' else types):
        cfg = _lib.make_config(lt, 'log1p', True, 1.0, 1.0, (0, 0, 0.5))
        ms = time_launch(lambda: _lib.check(lib.gd_pairwise(
            ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
            mat.data_ptr(), 256, stream), 'gd_pairwise'), 20)
        ms2 = time_launch(lambda: _lib.check(lib.gd_pairwise_row_argmin(
            ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
            vmin.data_ptr(), idx.data_ptr(), stream), 'gd_pairwise_row_argmin'), 20)
        cmin = torch.empty(256, device=dev)
        cidx = torch.empty(256, dtype=torch.int32, device=dev)
        pws = ops._pair_workspace(dev, 256)
        ms3 = time_launch(lambda: _lib.check(lib.gd_pairwise_assign(
            ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
            vmin.data_ptr(), idx.data_ptr(), cmin.data_ptr(), cidx.data_ptr(), None, 256, 0,
            pws.data_ptr(), pws.numel(), stream), 'gd_pairwise_assign'), 20)
        pairs = 200_000 * 256
        packed = {}
        if args.packed and lt in ('gwd3d', 'kld3d', 'bd3d'):
            # one column per lane (GD_PAIR_CPL1 = 2): reductions only, and with the matrix
            ms4 = time_launch(lambda: _lib.check(lib.gd_pairwise_assign(
                ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
                vmin.data_ptr(), idx.data_ptr(), cmin.data_ptr(), cidx.data_ptr(), None, 256, 2,
                pws.data_ptr(), pws.numel(), stream), 'gd_pairwise_assign'), 20)
            ms5 = time_launch(lambda: _lib.check(lib.gd_pairwise_assign(
                ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
                vmin.data_ptr(), idx.data_ptr(), cmin.data_ptr(), cidx.data_ptr(), mat.data_ptr(),
                256, 2, pws.data_ptr(), pws.numel(), stream), 'gd_pairwise_assign'), 20)
            ms6 = time_launch(lambda: _lib.check(lib.gd_pairwise_assign(
                ctypes.byref(cfg), anchors.data_ptr(), 200_000, gts.data_ptr(), 256,
                vmin.data_ptr(), idx.data_ptr(), cmin.data_ptr(), cidx.data_ptr(), mat.data_ptr(),
                256, 0, pws.data_ptr(), pws.numel(), stream), 'gd_pairwise_assign'), 20)
            packed = {'cpl1_assign_ms': round(ms4, 4),
                      'cpl1_assign_Gpairs_per_s': round(pairs / ms4 / 1e6, 2),
                      'cpl1_assign_matrix_ms': round(ms5, 4),
                      'assign_matrix_ms': round(ms6, 4)}
        out['pairwise'].append({'loss': lt, 'n': 200_000, 'm': 256, 'matrix_ms': round(ms, 4), **packed,
                                'matrix_Gpairs_per_s': round(pairs / ms / 1e6, 2),
                                'matrix_write_GBps': round(4 * pairs / ms / 1e6, 1),
                                'argmin_ms': round(ms2, 4),
                                'argmin_Gpairs_per_s': round(pairs / ms2 / 1e6, 2),
                                'assign_ms': round(ms3, 4),
                                'assign_Gpairs_per_s': round(pairs / ms3 / 1e6, 2)})
    # small-M shapes (a handful of GTs per sample, KITTI-like): fused assign only
    for m in ([] if args.only == 'elementwise' else (8, 32, 64)):
        cfg = _lib.make_config('gwd3d', 'log1p', True, 1.0, 1.0, (0, 0, 0.5))
        g2 = synth.make_targets(m, 'kitti', seed=6, device=dev)
        nn = 321_408
        a2 = synth.make_anchor_grid(nn, 'kitti', device=dev)
        v2 = torch.empty(nn, device=dev)
        i2 = torch.empty(nn, dtype=torch.int32, device=dev)
        cm = torch.empty(m, device=dev)
        ci = torch.empty(m, dtype=torch.int32, device=dev)
        pws = ops._pair_workspace(dev, m)
        ms = time_launch(lambda: _lib.check(lib.gd_pairwise_assign(
            ctypes.byref(cfg), a2.data_ptr(), nn, g2.data_ptr(), m, v2.data_ptr(), i2.data_ptr(),
            cm.data_ptr(), ci.data_ptr(), None, m, 0, pws.data_ptr(), pws.numel(), stream),
            'gd_pairwise_assign'), 20)
        out['pairwise'].append({'loss': 'gwd3d', 'n': nn, 'm': m, 'assign_ms': round(ms, 4),
                                'assign_Gpairs_per_s': round(nn * m / ms / 1e6, 2)})
    print(json.dumps(out))


if __name__ == '__main__':
    main()
