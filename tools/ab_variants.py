#!/usr/bin/env python
"""A/B of production-style builds of the fused kernel in ONE process on ONE GPU
(boxes differ by a few percent under the power cap, so variants must be compared inside
one run): `python tools/ab_variants.py 0 64 128 192` times libgdloss_b200_v<bits>.so
(build_ext.build_variant) on the five bench configurations at 2^24 pairs, interleaved
over several rounds, and prints one JSON document.

Bits of <bits> = GD_TUNE_DEFAULT (csrc/gd_loss_kernels.cuh): 1 late store wait, 2 loads
evict_normal, 4 store hint, 16 generic stores, 64 lane-0 issue, 128 strided layout,
256 min/max row screen, 512 alpha == 1 / center_offset == (0,0,.5) folded, 1024 packed-FP32
math.  Variants with bit 512 or 1024 are not bit-identical to variant 0 on the device (FMA
contraction, FFMA2): `grad_max_row_rel_diff_to_first` reports how far they are."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdet3d_gaussian_b200 import _lib, build_ext, synth  # noqa: E402
from tune_sweep import COMBOS, bind, timed  # noqa: E402


def sustained(launch, sampler, seconds=2.0):
    """~2 s of back-to-back launches of ONE variant: ms per launch + clocks / power."""
    import time
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    calls = 0
    while time.time() - t0 < seconds:
        for _ in range(200):
            launch()
        calls += 200
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    row = {'ms': round(e0.elapsed_time(e1) / calls, 5)}
    row.update(sampler.window(t0, t1))
    time.sleep(0.5)
    return row


def main():
    power = '--power' in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith('--')] or ['0']
    n = 1 << 24
    torch.cuda.set_device(0)
    libs = {}
    for nm in names:
        path = build_ext.lib_path() if nm == 'prod' else os.path.join(
            build_ext.PKG_DIR, f'libgdloss_b200_v{nm}.so')
        libs[nm] = bind(path)
    pred, target, weight = synth.make_pairs(n, 'kitti', seed=0, device='cuda')
    grad = torch.empty(n, 7, device='cuda')
    ref = torch.empty(n, 7, device='cuda')
    loss = torch.empty((), device='cuda')
    ws = torch.zeros(next(iter(libs.values())).gd_loss_workspace_bytes(n), dtype=torch.uint8,
                     device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cfgs = {c: _lib.make_config(c[0], c[1], True, 0.0, 1.0, (0, 0, 0.5)) for c in COMBOS}

    def launcher(lib, cfg, g):
        def launch():
            code = lib.gd_loss_fwd_bwd(
                ctypes.byref(cfg), pred.data_ptr(), 7, target.data_ptr(), 7, weight.data_ptr(),
                1, 1, n, 5.0 / n, loss.data_ptr(), None, g.data_ptr(), ws.data_ptr(), ws.numel(),
                _lib.VARIANTS['bulk'], 0, stream)
            if code != 0:
                raise RuntimeError(f'gd_loss_fwd_bwd -> {code}')
        return launch

    rounds = 4
    ms = {nm: {c: [] for c in COMBOS} for nm in names}
    same = {nm: True for nm in names}
    maxrel = {nm: 0.0 for nm in names}     # worst per-row |dgrad| / |grad| vs the first library
    for c in COMBOS:
        launcher(libs[names[0]], cfgs[c], ref)()
        for nm in names[1:]:
            launcher(libs[nm], cfgs[c], grad)()
            torch.cuda.synchronize()
            same[nm] = same[nm] and bool(torch.equal(grad, ref))
            if not same[nm]:
                num = (grad - ref).norm(dim=1)
                den = ref.norm(dim=1).clamp_min(1e-30)
                maxrel[nm] = max(maxrel[nm], float((num / den).max()))
    for _ in range(rounds):
        for c in COMBOS:
            for nm in names:
                ms[nm][c].append(timed(launcher(libs[nm], cfgs[c], grad), 25, 5))
    out = {'n': n, 'rounds': rounds, 'variants': {}}
    for nm in names:
        per = {f'{c[0]}/{c[1]}': round(88 * n / (sum(v) / len(v)) / 1e6, 1) for c, v in ms[nm].items()}
        bench = [per[f'{lt}/{fun}'] for lt, fun in COMBOS[:4]]
        out['variants'][nm] = {'GBps': per, 'bench_mean_GBps': round(4 / sum(1 / x for x in bench), 1),
                               'grad_bit_identical_to_first': same[nm],
                               'grad_max_row_rel_diff_to_first': maxrel[nm]}
        sys.stderr.write(f"{nm:>6s} bench-mean {out['variants'][nm]['bench_mean_GBps']:8.1f}  " +
                         ' '.join(f'{v:.0f}' for v in per.values()) + '\n')
    if power:
        # each variant alone, long enough for the power cap to settle: GB/s next to the
        # SM clock and board power it was achieved at
        from power_probe import Sampler
        sampler = Sampler()
        for nm in names:
            sus = {}
            for c in (COMBOS[0], COMBOS[3]):
                row = sustained(launcher(libs[nm], cfgs[c], grad), sampler)
                row['GBps'] = round(88 * n / row['ms'] / 1e6, 1)
                sus[f'{c[0]}/{c[1]}'] = row
                sys.stderr.write(f'{nm:>6s} sustained {c[0]}/{c[1]} {json.dumps(row)}\n')
            out['variants'][nm]['sustained'] = sus
        sampler.proc.terminate()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
