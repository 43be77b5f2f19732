#!/usr/bin/env python
"""Per-call latency at the sizes training actually uses (C1 = 100k
pairs; real heads pass P ~ 10^2-10^3 positives): module forward+backward wall time,
the bare C-ABI launch, CUDA-graph replay of the module call, and the oracle port on
the host cores.  `--profile` prints a cProfile of the module path.  One JSON line."""
import argparse
import cProfile
import ctypes
import json
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import GDLoss, _lib, ops, synth  # noqa: E402
from oracle import gd_oracle  # noqa: E402

KW = dict(loss_type='gwd3d', fun='log1p', tau=0.0, loss_weight=5.0)     # BASELINE configs[0]


def wall_us(fn, reps, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--profile', action='store_true')
    args = ap.parse_args()
    torch.cuda.set_device(0)
    lib = _lib.load()
    out = []
    for n in (1000, 100_000):
        pred, target, w = synth.make_pairs(n, 'kitti', seed=0, device='cuda')
        pred.requires_grad_(True)
        w7 = w[:, None].expand(n, 7).contiguous()
        res = dict(pairs=n)
        for name, mod, wt in (('module_default_w7', GDLoss(**KW), w7),
                              ('module_default_w1', GDLoss(**KW), w),
                              ('module_default_noweight', GDLoss(**KW), None),
                              ('module_nosync_w7', GDLoss(host_sync=False, **KW), w7)):
            def call():
                pred.grad = None
                mod(pred, target, wt, avg_factor=float(n)).backward()
            res[name + '_us'] = round(wall_us(call, 300), 2)

            def fwd_only():
                with torch.no_grad():
                    mod(pred, target, wt, avg_factor=float(n))
            res[name + '_fwd_us'] = round(wall_us(fwd_only, 300), 2)
        # what torch's autograd engine charges for ANY CUDA backward() (thread hand-off to the
        # device worker): the smallest possible graph, one elementwise op + sum
        def trivial():
            pred.grad = None
            (pred * 1.0).sum().backward()
        res['torch_trivial_fwd_bwd_us'] = round(wall_us(trivial, 300), 2)
        # bare C ABI
        cfg = _lib.make_config('gwd3d', 'log1p', True, 0.0, 1.0, (0, 0, 0.5))
        grad = torch.empty(n, 7, device='cuda')
        loss = torch.empty((), device='cuda')
        ws = ops._workspace(pred.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        pd = pred.detach()

        def abi():
            lib.gd_loss_fwd_bwd(ctypes.byref(cfg), pd.data_ptr(), 7, target.data_ptr(), 7,
                                w.data_ptr(), 1, 1, n, 5.0 / n, loss.data_ptr(), None,
                                grad.data_ptr(), ws.data_ptr(), ws.numel(), 0, 0, stream)
        res['c_abi_us'] = round(wall_us(abi, 2000), 2)
        # the module's forward alone from C++ (no Python module call): shim.gd_loss
        sh = _lib.shim()
        scfg = _lib.make_shim_config('gwd3d', 'log1p', True, 0.0, 1.0, (0, 0, 0.5))

        def shim_call():
            pred.grad = None
            sh.gd_loss(pred, target, w7, scfg, 5.0, 1, float(n), 0, 1).backward()
        res['shim_fwd_bwd_w7_us'] = round(wall_us(shim_call, 1000), 2)
        # CUDA graph of the module call (host_sync=False): replay cost
        mod = GDLoss(**KW)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            p_static = pred.detach().clone().requires_grad_(True)
            g0, = torch.autograd.grad(mod(p_static, target, w7, avg_factor=float(n)), p_static)
            s.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                gl = mod(p_static, target, w7, avg_factor=float(n))
                gg, = torch.autograd.grad(gl, p_static)
        torch.cuda.synchronize()
        res['graph_replay_us'] = round(wall_us(graph.replay, 2000), 2)
        # oracle port on the host (reference algorithm, eager torch)
        torch.set_num_threads(os.cpu_count() or 1)
        cp, ct, cw = pred.detach().cpu(), target.cpu(), w7.cpu()
        omod = gd_oracle.GDLossOracle(**KW)
        best = float('inf')
        for _ in range(5):
            q = cp.clone().requires_grad_(True)
            t0 = time.perf_counter()
            omod(q, ct, cw, avg_factor=float(n)).backward()
            best = min(best, time.perf_counter() - t0)
        res['cpu_port_us'] = round(best * 1e6, 1)
        res['cpu_cores'] = os.cpu_count()
        out.append(res)
        if args.profile and n == 1000:
            mod = GDLoss(**KW)
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(500):
                pred.grad = None
                mod(pred, target, w7, avg_factor=float(n)).backward()
            torch.cuda.synchronize()
            pr.disable()
            pstats.Stats(pr, stream=sys.stderr).sort_stats('cumulative').print_stats(35)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
