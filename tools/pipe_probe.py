#!/usr/bin/env python
"""GB/s and board power of the candidate data-movement pipelines for the fused kernel
(tools/micro/pipe_probe.cu: flat LDG/STG, bulk copies only, the shipped bulk -> LDS -> STS
-> bulk-store ring, register-only LDG.128/STG.128, bulk in + STG out), each carrying the
kernel's 88 B/pair traffic mix with one FFMA per element, next to a 1 GiB `copy_`.
nvidia-smi is sampled every 50 ms while each mode runs (~2 s); prints one JSON document.

    python tools/pipe_probe.py [seconds] [log2_rows]
"""
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

SRC = os.path.join(HERE, 'micro', 'pipe_probe.cu')
EXE = os.path.join(HERE, 'micro', 'pipe_probe')


def build():
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= os.path.getmtime(SRC):
        return
    subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
                    '-o', EXE, SRC], check=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    seconds = args[0] if len(args) > 0 else '2'
    lg = args[1] if len(args) > 1 else '24'
    build()
    if '--build-only' in sys.argv:
        return
    from power_probe import Sampler          # needs torch + a GPU
    import torch
    sampler = Sampler()
    out = []
    torch.cuda.set_device(0)
    src = torch.empty(1 << 28, device='cuda')
    dst = torch.empty_like(src)
    for _ in range(5):
        dst.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    calls = 0
    while time.time() - t0 < float(seconds):
        for _ in range(200):
            dst.copy_(src)
        calls += 200
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / calls
    row = {'name': 'torch copy_ 1 GiB', 'ms': round(ms, 5),
           'GBps': round(2 * src.numel() * 4 / ms / 1e6, 1)}
    row.update(sampler.window(t0, t1))
    out.append(row)
    sys.stderr.write(json.dumps(row) + '\n')
    del src, dst
    torch.cuda.empty_cache()
    time.sleep(1.0)
    for mode in range(5):
        res = subprocess.run([EXE, str(mode), seconds, lg], capture_output=True, text=True,
                             timeout=120)
        if res.returncode not in (0, 3):
            out.append({'mode': mode, 'error': res.stderr[-300:]})
            continue
        row = json.loads(res.stdout.strip().splitlines()[-1])
        row.update(sampler.window(row['t0'], row['t1']))
        out.append(row)
        sys.stderr.write(json.dumps(row) + '\n')
        time.sleep(1.0)
    sampler.proc.terminate()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
