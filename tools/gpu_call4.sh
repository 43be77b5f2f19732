#!/bin/bash
# pairwise kernel A/B (two columns per lane + lean post map vs one column per lane), the
# pairwise / assign / packed test files, bench with the fixed C5 sweep
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_assign.py tests/test_eval_affinity.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider -k "pairwise or assign or affinity or packed or early or smoke" > $OUT/pytest_pairwise.log 2>&1
stamp "pytest (pairwise surface) exit $?"; tail -8 $OUT/pytest_pairwise.log
timeout -s KILL 300 python tools/sweep.py --only pairwise --cpl1 > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
stamp "sweep pairwise exit $?"; cat $OUT/sweep_pairwise.json
GD_B200_PAIR_CPL=1 timeout -s KILL 300 python tools/sweep.py --only pairwise > $OUT/sweep_pairwise_cpl1.json 2> $OUT/sweep_pairwise_cpl1.err
stamp "sweep pairwise (CPL1 forced) exit $?"; cat $OUT/sweep_pairwise_cpl1.json
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err
stamp "bench exit $?"; python - <<PY
import json
d = json.load(open('$OUT/bench.json'))
print('value', d['value'], 'frac', d['roofline']['frac'])
for r in d['c5']['rows']:
    print(r)
print(d['pairwise'])
PY
ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_kernel -c 2 -o $OUT/pairwise_ncu python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
stamp "ncu pairwise exit $?"
ncu -i $OUT/pairwise_ncu.ncu-rep --page raw --csv > $OUT/pairwise_ncu_raw.csv 2>/dev/null
ls -la $OUT
