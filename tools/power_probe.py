#!/usr/bin/env python
"""Is the fused kernel power-capped?  Runs a few knob settings of the tune build
(tools/tune_sweep.py) for ~2 s each at 2^24 pairs while sampling nvidia-smi (SM / memory
clocks, board power, throttle reasons) and prints one JSON document: achieved GB/s next
to the median clocks and power of each setting."""
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, build_ext, synth  # noqa: E402
from tune_sweep import bind  # noqa: E402

Q = 'clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu'


class Sampler:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(['nvidia-smi', '-i', '0', f'--query-gpu={Q}',
                                      '--format=csv,noheader,nounits', '-lms', '50'],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def window(self, t0, t1):
        sel = [r for ts, r in self.rows if t0 + 0.3 <= ts <= t1]
        if not sel:
            return {}
        num = lambda i: statistics.median(float(r[i]) for r in sel)   # noqa: E731
        return {'sm_mhz': num(0), 'mem_mhz': num(1), 'power_w': num(2), 'temp_c': num(4),
                'power_cap_active': sum(r[3].lower().startswith('active') for r in sel) / len(sel),
                'samples': len(sel)}


def main():
    n = 1 << 24
    torch.cuda.set_device(0)
    tune = bind(build_ext.lib_path(tune=True))
    pred, target, weight = synth.make_pairs(n, 'kitti', seed=0, device='cuda')
    grad = torch.empty(n, 7, device='cuda')
    loss = torch.empty((), device='cuda')
    ws = torch.zeros(tune.gd_loss_workspace_bytes(n), dtype=torch.uint8, device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    src = torch.empty(1 << 28, device='cuda')          # 1 GiB copy for the baseline line
    dst = torch.empty_like(src)
    sampler = Sampler()
    out = []

    def run(name, fn, bytes_per_call):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        calls = 0
        while time.time() - t0 < 2.0:
            for _ in range(200):
                fn()
            calls += 200
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = e0.elapsed_time(e1) / calls
        row = {'setting': name, 'ms': round(ms, 5), 'GBps': round(bytes_per_call / ms / 1e6, 1)}
        row.update(sampler.window(t0, t1))
        out.append(row)
        sys.stderr.write(json.dumps(row) + '\n')
        time.sleep(1.0)

    run('torch copy_ 1 GiB', lambda: dst.copy_(src), 2 * src.numel() * 4)
    for lt, fun in (('kld3d', 'none'), ('bd3d', 'log1p')):
        cfg = _lib.make_config(lt, fun, True, 0.0, 1.0, (0, 0, 0.5))

        def launch():
            code = tune.gd_loss_fwd_bwd(ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7,
                                        weight.data_ptr(), 1, 1, n, 5.0 / n, loss.data_ptr(), None,
                                        grad.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _lib.VARIANTS['bulk'], 0, stream)
            assert code == 0, code
        for name, flags, warps, grid in (('base', 0, None, None), ('half_math', 32, None, None),
                                         ('no_math', 8, None, None), ('warps8', 0, 8, None),
                                         # fewer active SMs under the power cap (GD_TUNE_GRID)
                                         ('grid140', 0, None, 140), ('grid132', 0, None, 132),
                                         ('grid120', 0, None, 120), ('grid104', 0, None, 104),
                                         ('no_math grid132', 8, None, 132)):
            os.environ['GD_TUNE_FLAGS'] = str(flags)
            for key, val in (('GD_TUNE_WARPS', warps), ('GD_TUNE_GRID', grid)):
                if val is None:
                    os.environ.pop(key, None)
                else:
                    os.environ[key] = str(val)
            run(f'{lt}/{fun} {name}', launch, 88 * n)
        os.environ.pop('GD_TUNE_GRID', None)
    sampler.proc.terminate()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
