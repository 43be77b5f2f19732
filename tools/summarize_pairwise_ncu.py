"""Condenses `ncu --page raw --csv` exports of the pairwise kernels into profiles/pairwise_ncu.json
(read by bench.py's `pairwise` sub-record as a RECORDED capture).

    python tools/summarize_pairwise_ncu.py gpurun_out/<tag>/prof_pairwise_matrix_raw.csv \
        gpurun_out/<tag>/prof_pairwise_rowlane_raw.csv > profiles/pairwise_ncu.json
"""
import csv
import json
import sys

PAIRS = 200_000 * 256
KEYS = {
    'gpu__time_duration.sum': 'time_us',
    'smsp__inst_executed.sum': 'warp_instructions',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'fma_pipe_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'alu_pipe_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_slots_busy_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__grid_size': 'grid',
    'smsp__warps_eligible.avg.per_cycle_active': 'eligible_warps_per_cycle',
    'sm__cycles_active.min': 'sm_cycles_active_min',
    'sm__cycles_active.avg': 'sm_cycles_active_avg',
    'sm__cycles_active.max': 'sm_cycles_active_max',
    'dram__bytes_write.sum': 'dram_write',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'stall_not_selected_per_issue',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait_per_issue',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_pipe_throttle_per_issue',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier_per_issue',
}


def main():
    out = []
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        r = rows[2]                                    # first captured launch
        d = {'kernel': r[hdr.index('Kernel Name')]}
        for k, name in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    v = round(float(v), 3)
                except ValueError:
                    pass
                d[name] = v
                if name == 'dram_write':
                    d['dram_write_unit'] = units[hdr.index(k)]
        if 'warp_instructions' in d:
            d['thread_instructions_per_pair'] = round(d['warp_instructions'] * 32 / PAIRS, 1)
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
