#!/usr/bin/env python
"""Interleaved A/B of the fused kernel's grid size on one GPU: does running on fewer SMs (the board
is power capped: a smaller grid may hold a higher SM clock) move more bytes?  Uses the PRODUCTION
library through `gd_set_loss_grid`, variant 'auto', the five bench configurations at 2^24 pairs;
ROUNDS rounds, every grid in every round, 25 launches each.
Prints one JSON document (mean GB/s per grid and per configuration)."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, build_ext, synth  # noqa: E402

GRIDS = [148, 140, 136, 134, 132, 130, 128, 126, 124, 120]
COMBOS = (('kld3d', 'none'), ('kld3d', 'log1p'), ('bd3d', 'none'), ('bd3d', 'log1p'), ('gwd3d', 'log1p'))
ROUNDS = 6


def main():
    n = 1 << 24
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    lib = ctypes.CDLL(build_ext.lib_path())
    lib.gd_set_loss_grid.argtypes = [ctypes.c_int32]
    restype, argtypes = _lib.SIGNATURES['gd_loss_fwd_bwd']
    lib.gd_loss_fwd_bwd.restype, lib.gd_loss_fwd_bwd.argtypes = restype, argtypes
    lib.gd_loss_workspace_bytes.restype = ctypes.c_size_t
    lib.gd_loss_workspace_bytes.argtypes = [ctypes.c_int64]
    pred, target, weight = synth.make_pairs(n, 'kitti', seed=0, device=dev)
    grad = torch.empty(n, 7, device=dev)
    loss = torch.empty((), device=dev)
    ws = torch.zeros(lib.gd_loss_workspace_bytes(n), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfgs = {c: _lib.make_config(c[0], c[1], True, 0.0, 1.0, (0, 0, 0.5)) for c in COMBOS}

    def launch(cfg):
        code = lib.gd_loss_fwd_bwd(ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7,
                                   weight.data_ptr(), 1, 1, n, 5.0 / n, loss.data_ptr(), None,
                                   grad.data_ptr(), ws.data_ptr(), ws.numel(), _lib.VARIANTS['auto'], 0,
                                   stream)
        if code != 0:
            raise RuntimeError(code)

    def timed(cfg, reps=25):
        for _ in range(3):
            launch(cfg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            launch(cfg)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = {g: {c: [] for c in COMBOS} for g in GRIDS}
    for c in COMBOS:                               # warm the clocks / power state
        timed(cfgs[c], 50)
    for r in range(ROUNDS):
        order = GRIDS if r % 2 == 0 else GRIDS[::-1]
        for g in order:
            lib.gd_set_loss_grid(g)
            for c in COMBOS:
                ms[g][c].append(timed(cfgs[c]))
    out = {'n': n, 'rounds': ROUNDS, 'grids': {}}
    for g in GRIDS:
        per = {f'{c[0]}/{c[1]}': round(88 * n / (sum(v) / len(v)) / 1e6, 1) for c, v in ms[g].items()}
        tot = sum(sum(v) / len(v) for v in ms[g].values())
        out['grids'][g] = {'mean_GBps': round(88 * n * len(COMBOS) / tot / 1e6, 1), 'per_config': per,
                           'spread_ms_kld_none': [round(x, 4) for x in ms[g][COMBOS[0]]]}
        sys.stderr.write(f"grid {g}: mean {out['grids'][g]['mean_GBps']} GB/s {per}\n")
    print(json.dumps(out))


if __name__ == '__main__':
    main()
