#!/usr/bin/env python
"""Does a smaller grid of the fused kernel hold its advantage when the launches go on for seconds
(the board heats up, the power controller settles)?  For each grid: 4 s of back-to-back launches
of the four bench configurations (round robin), GB/s per 0.5-s window, SM clock / power sampled."""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, build_ext, synth  # noqa: E402

COMBOS = (('kld3d', 'none'), ('kld3d', 'log1p'), ('bd3d', 'none'), ('bd3d', 'log1p'))


def smi():
    out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,temperature.gpu',
                          '--format=csv,noheader,nounits', '-i', '0'], capture_output=True, text=True).stdout
    return [float(x) for x in out.strip().split(',')]


def main():
    grids = [int(x) for x in sys.argv[1:]] or [148, 132, 128, 124, 148, 128]
    n = 1 << 24
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    lib = ctypes.CDLL(build_ext.lib_path())
    restype, argtypes = _lib.SIGNATURES['gd_loss_fwd_bwd']
    lib.gd_loss_fwd_bwd.restype, lib.gd_loss_fwd_bwd.argtypes = restype, argtypes
    lib.gd_loss_workspace_bytes.restype = ctypes.c_size_t
    lib.gd_loss_workspace_bytes.argtypes = [ctypes.c_int64]
    lib.gd_set_loss_grid.argtypes = [ctypes.c_int32]
    pred, target, weight = synth.make_pairs(n, 'kitti', seed=0, device=dev)
    grad = torch.empty(n, 7, device=dev)
    loss = torch.empty((), device=dev)
    ws = torch.zeros(lib.gd_loss_workspace_bytes(n), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfgs = [_lib.make_config(c[0], c[1], True, 0.0, 1.0, (0, 0, 0.5)) for c in COMBOS]

    def launch(cfg):
        lib.gd_loss_fwd_bwd(ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7,
                            weight.data_ptr(), 1, 1, n, 5.0 / n, loss.data_ptr(), None, grad.data_ptr(),
                            ws.data_ptr(), ws.numel(), _lib.VARIANTS['auto'], 0, stream)
    out = []
    for g in grids:
        lib.gd_set_loss_grid(g)
        windows, samples = [], []
        t_end = time.time() + 4.0
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(120):                     # ~0.12 s of kernels per window (4 configs x 30 x 4)
                for cfg in cfgs:
                    launch(cfg)
            e1.record()
            samples.append(smi())                    # while the GPU is still busy
            torch.cuda.synchronize()
            windows.append(round(88 * n * 480 / e0.elapsed_time(e1) / 1e6, 1))
        row = {'grid': g, 'GBps_windows': windows, 'sm_mhz': [s[0] for s in samples],
               'power_w': [s[1] for s in samples], 'temp_c': [s[2] for s in samples]}
        out.append(row)
        sys.stderr.write(f"grid {g}: {windows}\n   MHz {row['sm_mhz']}\n   W {row['power_w']}  T {row['temp_c']}\n")
    print(json.dumps(out))


if __name__ == '__main__':
    main()
