#!/usr/bin/env python
"""Knob sweep of gd_warp_kernel on one GPU (kernel-only, CUDA events around bare
C-ABI launches of ``libgdloss_b200_tune.so``, the -DGD_TUNE=1 build whose knobs are
read per launch from GD_TUNE_FLAGS / GD_TUNE_WARPS; see gd_loss_kernels.cuh).

For every knob setting: the 5 bench configurations at 2^24 pairs (and fwd-only),
the gradient checked against the production library's (bit-identical expected,
except for the measurement-only no-math knob).  Prints one JSON document.
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, build_ext, synth  # noqa: E402

FLAGS = {'late_wait': 1, 'load_normal': 2, 'store_hint': 4, 'no_math': 8, 'generic_store': 16,
         'half_math': 32, 'lane0': 64}
SETTINGS = [
    ('base', 0, None), ('lane0', 64, None), ('late_wait', 1, None), ('load_normal', 2, None), ('store_hint', 4, None),
    ('late_wait+store_hint', 5, None), ('late_wait+load_normal', 3, None),
    ('generic_store', 16, None), ('generic_store+load_normal', 18, None),
    ('no_math', 8, None), ('half_math', 32, None), ('no_math+generic_store', 24, None),
    ('warps10', 0, 10), ('warps8', 0, 8), ('warps6', 0, 6),
    ('late_wait+warps10', 1, 10), ('late_wait+warps8', 1, 8),
]
COMBOS = (('kld3d', 'none'), ('kld3d', 'log1p'), ('bd3d', 'none'), ('bd3d', 'log1p'),
          ('gwd3d', 'log1p'))


def bind(path):
    lib = ctypes.CDLL(path)
    restype, argtypes = _lib.SIGNATURES['gd_loss_fwd_bwd']
    lib.gd_loss_fwd_bwd.restype, lib.gd_loss_fwd_bwd.argtypes = restype, argtypes
    lib.gd_loss_workspace_bytes.restype = ctypes.c_size_t
    lib.gd_loss_workspace_bytes.argtypes = [ctypes.c_int64]
    return lib


def timed(fn, reps=30, warm=5):
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
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--log2n', type=int, default=24)
    ap.add_argument('--reps', type=int, default=30)
    args = ap.parse_args()
    n = 1 << args.log2n
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    prod = bind(build_ext.lib_path())
    tune = bind(build_ext.lib_path(tune=True))
    pred, target, weight = synth.make_pairs(n, 'kitti', seed=0, device=dev)
    grad = torch.empty(n, 7, device=dev)
    grad_ref = torch.empty(n, 7, device=dev)
    loss = torch.empty((), device=dev)
    loss_ref = torch.empty((), device=dev)
    ws = torch.zeros(prod.gd_loss_workspace_bytes(n), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launcher(lib, cfg, g, lo):
        def launch():
            code = lib.gd_loss_fwd_bwd(
                ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7, weight.data_ptr(),
                1, 1, n, 5.0 / n, lo.data_ptr(), None, g.data_ptr() if g is not None else None,
                ws.data_ptr(), ws.numel(), _lib.VARIANTS['bulk'], 0, stream)
            if code != 0:
                raise RuntimeError(f'gd_loss_fwd_bwd -> {code}')
        return launch

    out = {'n': n, 'reps': args.reps, 'results': []}
    cfgs = {c: _lib.make_config(c[0], c[1], True, 0.0, 1.0, (0, 0, 0.5)) for c in COMBOS}
    prod_ms = {}
    for c in COMBOS:
        prod_ms[c] = timed(launcher(prod, cfgs[c], grad_ref, loss_ref), args.reps)
    out['production'] = {f'{c[0]}/{c[1]}': {'ms': round(prod_ms[c], 5),
                                            'GBps': round(88 * n / prod_ms[c] / 1e6, 1)}
                         for c in COMBOS}
    for name, flags, warps in SETTINGS:
        os.environ['GD_TUNE_FLAGS'] = str(flags)
        if warps is None:
            os.environ.pop('GD_TUNE_WARPS', None)
        else:
            os.environ['GD_TUNE_WARPS'] = str(warps)
        row = {'setting': name, 'flags': flags, 'warps': warps, 'configs': {}}
        tot = 0.0
        for c in COMBOS:
            launcher(prod, cfgs[c], grad_ref, loss_ref)()
            fn = launcher(tune, cfgs[c], grad, loss)
            fn()
            torch.cuda.synchronize()
            same = bool(torch.equal(grad, grad_ref)) and bool(torch.equal(loss, loss_ref))
            ms = timed(fn, args.reps)
            ms_fwd = timed(launcher(tune, cfgs[c], None, loss), args.reps)
            tot += ms
            row['configs'][f'{c[0]}/{c[1]}'] = {
                'ms': round(ms, 5), 'GBps': round(88 * n / ms / 1e6, 1),
                'fwd_ms': round(ms_fwd, 5), 'fwd_GBps': round(60 * n / ms_fwd / 1e6, 1),
                'bit_identical_to_production': same}
        row['mean_GBps'] = round(88 * n * len(COMBOS) / tot / 1e6, 1)
        out['results'].append(row)
        sys.stderr.write(f"{name:28s} mean {row['mean_GBps']:8.1f} GB/s  " + ' '.join(
            f"{v['GBps']:.0f}" for v in row['configs'].values()) + '\n')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
