#!/usr/bin/env python
"""Summarise gpurun_out/<tag>/{prof_bulk.ncu-rep,launches.csv} into profiles/.

    python tools/summarize_ncu.py r01c

Writes profiles/<tag>_ncu_full.md (per-launch table from the `ncu --set full` capture),
profiles/<tag>_launches.md (kernel shares from the gpu__time_duration launch list) and
updates profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by
bench.py for `roofline.traffic`).  Runs on the CPU box: `ncu -i` needs no GPU."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of ncu peak'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma pipe %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu pipe %'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'eligible warps/cycle'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
     'stall long_scoreboard /issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait /issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
     'stall not_selected /issue'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
     'stall no_instruction /issue'),
]


def to_bytes(val, unit):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    return float(val) * mult


def main():
    tag = sys.argv[1]
    src = os.path.join(ROOT, 'gpurun_out', tag)
    out_dir = os.path.join(ROOT, 'profiles')
    os.makedirs(out_dir, exist_ok=True)
    rep = os.path.join(src, 'prof_bulk.ncu-rep')
    raw_csv = os.path.join(src, 'prof_bulk_raw.csv')     # exported on the GPU box (ncu -i)
    if os.path.exists(rep) or os.path.exists(raw_csv):
        if os.path.exists(rep):
            raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                                 text=True).stdout
        else:
            with open(raw_csv) as f:
                raw = f.read()
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        name_i = hdr.index('Kernel Name')
        lines = [f'# ncu --set full capture ({tag}), one row per captured launch', '',
                 'Command: `ncu --set full --clock-control none --import-source on -k '
                 'regex:gd_warp_kernel -s 12 -c 4 python bench.py --steps 2 --warmup 3 --no-e2e '
                 '--no-cpu` (bench workload C2: 2^24 box pairs per launch; template args '
                 '`<loss, grad, rows/lane, spec, weight mode>`, loss 1 = kld3d, 5 = bd3d, '
                 'spec 8 = fun none, 9 = fun log1p).', '',
                 'Algorithmic bytes per launch: 88 B x 2^24 = 1.476 GB (read 1.007 GB, write '
                 '0.470 GB).', '']
        cols = ['kernel'] + [label for _, label in METRICS]
        lines.append('| ' + ' | '.join(cols) + ' |')
        lines.append('|' + '---|' * len(cols))
        traffic = []
        for r in data:
            cells = [r[name_i].replace('void gdk::', '').replace('(gdk::LossArgs)', '')]
            rd = wr = 0.0
            for m, _ in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    v, u = r[i], units[i]
                    if m == 'dram__bytes_read.sum':
                        rd = to_bytes(v, u)
                    if m == 'dram__bytes_write.sum':
                        wr = to_bytes(v, u)
                    try:
                        v = f'{float(v):.4g}'
                    except ValueError:
                        pass
                    cells.append(f'{v} {u}'.strip())
                else:
                    cells.append('n/a')
            traffic.append(rd + wr)
            lines.append('| ' + ' | '.join(cells) + ' |')
        lines += ['', f'DRAM traffic per launch (read+write): '
                  f'{", ".join(f"{t / 1e9:.3f} GB" for t in traffic)} vs 1.476 GB algorithmic '
                  f'(writes still resident in L2 at kernel end are not counted by ncu).']
        with open(os.path.join(out_dir, f'{tag}_ncu_full.md'), 'w') as f:
            f.write('\n'.join(lines) + '\n')
        if traffic:
            with open(os.path.join(out_dir, 'traffic.json'), 'w') as f:
                json.dump({'source': f'profiles/{tag}_ncu_full.md',
                           'kernel': 'gd_warp_kernel',
                           'dram_bytes_per_launch': sum(traffic) / len(traffic)}, f)
    lst = os.path.join(src, 'launches.csv')
    if os.path.exists(lst):
        agg = collections.OrderedDict()
        with open(lst) as f:
            body = [ln for ln in f if ln.startswith('"')]
        for r in csv.DictReader(io.StringIO(''.join(body))):
            if r.get('Metric Name') != 'gpu__time_duration.sum':
                continue
            name = r['Kernel Name'].split('(')[0].replace('void ', '')[:90]
            t = float(r['Metric Value']) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(
                r['Metric Unit'], 1.0)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
        total = sum(v[1] for v in agg.values())
        lines = [f'# ncu launch list ({tag}): gpu__time_duration.sum per kernel', '',
                 'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 120 '
                 'python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu` (first 120 launches '
                 'of the process: synthetic-data generation by torch, then warm-up steps of '
                 'the bench; cold-cache serialised times -- compare SHARES).', '',
                 '| kernel | launches | total us | share |', '|---|---|---|---|']
        for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append(f'| `{name}` | {cnt} | {t:.1f} | {100 * t / total:.1f}% |')
        ours = sum(t for n, (c, t) in agg.items() if 'gdk::' in n)
        ours_main = sum(t for n, (c, t) in agg.items() if 'gd_warp_kernel' in n)
        lines += ['', f'Library kernels (`gdk::*`): {100 * ours / total:.1f}% of captured GPU '
                  f'time; within the library the fused loss kernel is '
                  f'{100 * ours_main / max(ours, 1e-9):.1f}% (the rest: gd_scale_grad_kernel '
                  f'early-outs and gd_any_positive_kernel).']
        with open(os.path.join(out_dir, f'{tag}_launches.md'), 'w') as f:
            f.write('\n'.join(lines) + '\n')
    for name in ('bench_auto.json', 'sweep.json', 'bench_reference.json'):
        p = os.path.join(src, name)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            with open(p) as f, open(os.path.join(out_dir, f'{tag}_{name}'), 'w') as g:
                g.write(f.read())


if __name__ == '__main__':
    main()
