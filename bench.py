#!/usr/bin/env python
"""Benchmark of the GD-loss hot path (BASELINE.json metric: box pairs/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1] ("C2"): KLD and BCD
loss, tau=0, f(x)=x and f=log(1+x), fused forward+backward over 2^24 synthetic
KITTI-prior box pairs with [N] weights, loss_weight=5, reduction='mean'.  One
STEP = those four loss evaluations over the batch (4 x 2^24 pairs, 4 launches).
Inputs (1.0 GB) are far larger than the 126 MB L2, so every launch streams from
HBM (no flush needed; stated in `config.l2`).  N>1: one process per GPU
(torchrun), every rank owns 2^24 rows (weak scaling), one NCCL all-reduce of the
scalar loss after each evaluation, timing = max over ranks of CUDA-event time.

`value`   : device-resident throughput (inputs in HBM before the timed region).
`e2e`     : same four evaluations through the C ABI with HOST (pinned) buffers:
            H2D of pred/target/weight, kernels, D2H of grad and loss inside the
            timed region (gd_loss_fwd_bwd_host).
`roofline`: HBM; achieved = 88 B/pair x 2^24 / mean kernel time, peak from
            MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle port (oracle/gd_oracle.py, the
            reference's eager-torch algorithm) in fp32 on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS = 1 << 24
BYTES_PER_PAIR = 88          # pred 28 + target 28 + weight 4 read, grad 28 written
COMBOS = (('kld3d', 'none'), ('kld3d', 'log1p'), ('bd3d', 'none'), ('bd3d', 'log1p'))
LOSS_WEIGHT = 5.0
METRIC = 'box pairs/s, GD loss fwd+bwd (KLD+BCD, tau=0, f=x|log1p)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pairs', type=int, default=N_PAIRS)
    ap.add_argument('--variant', default='auto', choices=['auto', 'bulk', 'bulk_packed', 'bulk_r2', 'staged'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-chunk-log2', type=int, default=20,
                    help='rows per chunk of the host-buffer pipeline (e2e), as a power of two')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--eager-gpu', action='store_true',
                    help='also time the reference algorithm (oracle port, eager torch + autograd) '
                         'on THIS GPU: the incumbent a user of the reference runs today')
    ap.add_argument('--detail', action='store_true', help='extra per-config lines on stderr')
    ap.add_argument('--watchdog', type=float, default=900.0,
                    help='seconds after which all thread stacks are dumped and the process exits')
    ap.add_argument('--verbose', action='store_true', help='phase log with timestamps on stderr')
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get('dram_bytes_per_launch')
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.05):
                continue
            parts = [p.strip() for p in line.split(',')]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[2:]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm, eager torch, fp32, all host threads)
# ---------------------------------------------------------------------------
def cpu_pairs_per_s(sample_rows, chunk_rows, repeats, threads=None):
    import torch
    from oracle import gd_oracle
    from mmdet3d_gaussian_b200 import synth
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    pred, target, w = synth.make_pairs(sample_rows, 'kitti', seed=0)
    mods = [gd_oracle.GDLossOracle(lt, fun=fun, tau=0.0, loss_weight=LOSS_WEIGHT)
            for lt, fun in COMBOS]
    best = float('inf')
    for _ in range(repeats):
        t0 = time.perf_counter()
        for mod in mods:
            for lo in range(0, sample_rows, chunk_rows):
                p = pred[lo:lo + chunk_rows].clone().requires_grad_(True)
                loss = mod(p, target[lo:lo + chunk_rows], w[lo:lo + chunk_rows],
                           avg_factor=float(sample_rows))
                loss.backward()
        best = min(best, time.perf_counter() - t0)
    return len(COMBOS) * sample_rows / best, cores, best


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is allowed all
    # host cores (torch is imported only below, so the override takes effect, including
    # in the autograd engine's thread)
    os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
    os.environ.pop('MKL_NUM_THREADS', None)
    sample, chunk = 1 << 21, 1 << 17
    steps = max(args.steps, 1)
    for _ in range(min(args.warmup, 1)):
        cpu_pairs_per_s(1 << 16, 1 << 16, 1)
    t_budget = time.perf_counter()
    vals = []
    for _ in range(min(steps, 5)):
        v, cores, _ = cpu_pairs_per_s(sample, chunk, 1)
        vals.append(v)
        if time.perf_counter() - t_budget > 120:
            break
    value = max(vals)
    ms = len(COMBOS) * sample / value * 1e3
    desc = (f'oracle port (reference algorithm in eager torch, fp32, autograd backward), '
            f'4 configs x 2^21 pairs per step in 2^17-row chunks, best of {len(vals)} steps')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': len(vals), 'warmup': min(args.warmup, 1),
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'C2: kld3d+bd3d x fun{none,log1p}, tau=0, bounded sample of '
                               '2^21 pairs per config (full workload 2^24)',
                   'pairs_per_step': len(COMBOS) * sample},
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mmdet3d_gaussian_b200 import GDLoss, _lib, ops, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    t_start = time.time()

    def log(msg):
        if args.verbose:
            sys.stderr.write(f'[bench rank {rank} +{time.time() - t_start:6.1f}s] {msg}\n')
            sys.stderr.flush()
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    n = args.pairs
    log('process group and library ready')

    pred, target, weight = synth.make_pairs(n, 'kitti', seed=rank, device=dev)
    pred.requires_grad_(True)
    # host_sync=False: the early-return probe of GDLoss.forward (reference
    # gaussian_distance_loss.py:290, a device->host sync per call) is folded into the
    # kernel, so launches queue back to back; `value_default_module` below times the
    # faithful default (host_sync=True) for comparison.
    mods = [GDLoss(lt, fun=fun, tau=0.0, loss_weight=LOSS_WEIGHT, variant=args.variant,
                   host_sync=False) for lt, fun in COMBOS]
    mods_sync = [GDLoss(lt, fun=fun, tau=0.0, loss_weight=LOSS_WEIGHT, variant=args.variant)
                 for lt, fun in COMBOS]
    avg = float(n * world)
    losses = [None] * len(mods)
    pending = []

    def one_eval(i, modules=None, collective=True):
        pred.grad = None
        loss = (modules or mods)[i](pred, target, weight, avg_factor=avg)
        if world > 1 and collective:
            tot = loss.detach().clone()
            # the one collective of the path (4 bytes).  Asynchronous: its result is only read
            # at the end of the step, so the next evaluation's kernel need not wait for it
            pending.append(dist.all_reduce(tot, async_op=True))
            losses[i] = tot
        elif world == 1:
            losses[i] = loss.detach()
        loss.backward()                        # grad_output == 1: scale kernel exits at once

    def drain():
        while pending:
            pending.pop(0).wait()              # stream-side wait: no host sync

    def step():
        for i in range(len(mods)):
            one_eval(i)
        drain()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log('inputs generated')
    for _ in range(max(args.warmup, 3)):
        step()
    log('warm-up enqueued')
    sync_all()
    log('warm-up done')

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = lib.gd_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    t_wall1 = time.time()
    launches = lib.gd_launch_count() - launches0
    log('timed region done')
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = len(COMBOS) * n * world / (ms_step * 1e-3)

    # ---- same step through the default module (host sync per evaluation, as the reference)
    sync_all()
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_sync = max(2, min(args.steps, 20))
    evs0.record()
    for _ in range(k_sync):
        for i in range(len(mods_sync)):
            one_eval(i, mods_sync)
        drain()
    evs1.record()
    sync_all()
    value_sync = len(COMBOS) * n * world / (evs0.elapsed_time(evs1) / k_sync * 1e-3)
    log('default-module pass done')

    # ---- same semantics as the default, the fused launch queued before the host waits for the
    # probe (host_sync='overlap', opt-in).  Informational; never allowed to break the run.
    value_overlap = None
    try:
        if world > 1:
            raise RuntimeError('single-GPU runs only')
        mods_ov = [GDLoss(lt, fun=fun, tau=0.0, loss_weight=LOSS_WEIGHT, variant=args.variant,
                          host_sync='overlap') for lt, fun in COMBOS]
        for i in range(len(mods_ov)):
            one_eval(i, mods_ov)
        drain()
        sync_all()
        evo0, evo1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evo0.record()
        for _ in range(k_sync):
            for i in range(len(mods_ov)):
                one_eval(i, mods_ov)
            drain()
        evo1.record()
        sync_all()
        value_overlap = len(COMBOS) * n * world / (evo0.elapsed_time(evo1) / k_sync * 1e-3)
    except Exception as exc:                       # noqa: BLE001
        if world == 1:
            sys.stderr.write(f'bench: overlap-mode pass skipped: {exc!r}\n')
    log('overlap-module pass done')

    # ---- kernel-only durations per config (CUDA events around bare C-ABI launches)
    per_cfg = {}
    fused_ms = []
    grad_buf = torch.empty(n, 7, device=dev)
    loss_buf = torch.empty((), device=dev)
    ws = ops._workspace(dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    pd, td, wd = pred.detach(), target, weight
    reps = 20
    for (lt, fun) in COMBOS + (('gwd3d', 'log1p'),):
        cfg = _lib.make_config(lt, fun, True, 0.0, 1.0, (0, 0, 0.5))

        def launch():
            code = lib.gd_loss_fwd_bwd(ctypes.byref(cfg), pd.data_ptr(), 7, td.data_ptr(), 7,
                                       wd.data_ptr(), 1, 1, n, LOSS_WEIGHT / avg,
                                       loss_buf.data_ptr(), None, grad_buf.data_ptr(),
                                       ws.data_ptr(), ws.numel(),
                                       _lib.VARIANTS[args.variant], 0, stream)
            _lib.check(code, 'gd_loss_fwd_bwd')
        for _ in range(3):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        per_cfg[f'{lt}/{fun}'] = {'ms': round(ms, 4),
                                  'GBps': round(BYTES_PER_PAIR * n / ms / 1e6, 1),
                                  'Gpairs_per_s': round(n / ms / 1e6, 2)}
        if (lt, fun) in COMBOS:
            fused_ms.append(ms)
    kernel_ms = sum(fused_ms) / len(fused_ms)
    log('kernel-only timings done')

    # keep the device busy ~1.5 s more so the clock sampler sees it under load (rank 0
    # only, so NO collective in here: the other ranks are already past this point)
    t_end = time.time() + 1.5
    while rank == 0 and time.time() < t_end:
        for _ in range(20):
            for i in range(len(mods)):
                one_eval(i, collective=False)
        torch.cuda.synchronize()
    t_wall_load = time.time()
    clocks = sampler.stop(t_wall0, t_wall_load) if rank == 0 else None
    log('clock sampling done')

    # ---- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        hp, ht, hw = (x.detach().cpu().pin_memory() for x in (pred, target, weight))
        hgrad = torch.empty(n, 7).pin_memory()
        hloss = torch.zeros(1).pin_memory()
        cfgs = [_lib.make_config(lt, fun, True, 0.0, 1.0, (0, 0, 0.5)) for lt, fun in COMBOS]

        def e2e_step():
            for cfg in cfgs:
                code = lib.gd_loss_fwd_bwd_host(
                    ctypes.byref(cfg), hp.data_ptr(), ht.data_ptr(), hw.data_ptr(), 1, n,
                    LOSS_WEIGHT / avg, hloss.data_ptr(), hgrad.data_ptr(), local_rank,
                    1 << args.e2e_chunk_log2)
                _lib.check(code, 'gd_loss_fwd_bwd_host')
        log('host buffers pinned')
        e2e_step()
        log('first e2e step done')
        k = max(2, min(args.steps, 5))
        sync_all()
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': len(COMBOS) * n * world * k / dt, 'unit': 'pairs/s',
               'h2d_bytes_per_step': len(COMBOS) * n * 60,
               'd2h_bytes_per_step': len(COMBOS) * (n * 28 + 4),
               'steps': k, 'ms_per_step': dt / k * 1e3,
               'api': f'gd_loss_fwd_bwd_host (C ABI, pinned host buffers, 2^{args.e2e_chunk_log2}-row '
                      f'chunks, 3 streams)'}

    log('e2e done')
    cpu = None
    # CPU baseline: rank 0 at N=1 only (torchrun pins OMP_NUM_THREADS=1 in its workers,
    # which would serialise the autograd thread of the CPU port)
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_pairs_per_s(1 << 17, 1 << 17, 1)                     # warm the thread pool
        v, cores, secs = cpu_pairs_per_s(1 << 23, 1 << 17, 2)
        cpu = {'value': v, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
               'sample': 'oracle port (reference algorithm, eager torch fp32 + autograd), 4 '
                         f'configs x 2^23 pairs (half the batch) in 2^17-row chunks, best of 2 '
                         f'passes ({secs:.2f} s per pass)'}
        # the single-thread row of SURVEY.md section 8d (bounded: 4 x 2^20 pairs, one pass)
        v1, _, secs1 = cpu_pairs_per_s(1 << 20, 1 << 17, 1, threads=1)
        cpu['value_1thread'] = v1
        cpu['sample_1thread'] = f'4 configs x 2^20 pairs, one pass ({secs1:.2f} s)'
        torch.set_num_threads(os.cpu_count() or 1)

    eager = None
    if rank == 0 and world == 1 and args.eager_gpu:
        # baseline leg only: the reference's own op sequence (oracle port) on CUDA tensors
        from oracle import gd_oracle
        rows, chunk = min(n, 1 << 22), 1 << 20
        emods = [gd_oracle.GDLossOracle(lt, fun=fun, tau=0.0, loss_weight=LOSS_WEIGHT)
                 for lt, fun in COMBOS]

        def eager_pass():
            for mod in emods:
                for lo in range(0, rows, chunk):
                    p = pred.detach()[lo:lo + chunk].clone().requires_grad_(True)
                    mod(p, target[lo:lo + chunk], weight[lo:lo + chunk],
                        avg_factor=float(rows)).backward()
        eager_pass()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        eager_pass()
        g1.record()
        torch.cuda.synchronize()
        eager = {'value': len(COMBOS) * rows / (g0.elapsed_time(g1) * 1e-3), 'unit': 'pairs/s',
                 'kind': 'port on cuda (eager torch + autograd, fp32)',
                 'sample': f'4 configs x 2^{rows.bit_length() - 1} pairs in 2^20-row chunks'}
        log('eager-gpu baseline done')

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
        out = {
            'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'C2: kld3d+bd3d x fun{none,log1p}, tau=0, 2^24 KITTI-prior '
                                   'box pairs per GPU, weights [N], loss_weight=5, mean/avg_factor',
                       'pairs_per_step_per_gpu': len(COMBOS) * n, 'launches_per_step': 2 * len(COMBOS),
                       'l2': 'inputs 1.0 GB per launch >> 126 MB L2, no flush needed',
                       'variant': args.variant, 'module': 'GDLoss(host_sync=False)',
                       'parallelism': f'rows sharded x{world}, 1 NCCL all-reduce of the scalar per evaluation '
                                      f'(async, waited at the end of the step)'
                       if world > 1 else 'single GPU'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': recorded_traffic(),
                         'peak_source': peak_src, 'kernel': 'gd_warp_kernel (fused fwd+bwd, bulk-copy warp pipelines)',
                         'bytes_per_pair': BYTES_PER_PAIR, 'pairs_per_launch': n,
                         'kernel_ms': kernel_ms, 'frac_of_8TBps_nominal': achieved / 8000.0,
                         # the same figure from the timed region itself (module calls: the
                         # fused launch + the grad_output fold that exits at once + Python)
                         'achieved_timed_region': BYTES_PER_PAIR * len(COMBOS) * n / (ms_step * 1e-3) / 1e9,
                         'per_config': per_cfg},
            'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
            'value_default_module': value_sync, 'value_overlap_module': value_overlap,
            'gpu_eager_baseline': eager,
            'lib': os.path.relpath(_lib.loaded_path(), ROOT),
            'losses': [float(x) for x in losses],
        }
        print(json.dumps(out))
    log('report printed')
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    log('exit')


if __name__ == '__main__':
    a = parse()
    import faulthandler
    # a hung collective or device call must not hang the caller: dump every thread's stack
    # to stderr and exit non-zero
    faulthandler.dump_traceback_later(a.watchdog, exit=True)
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
