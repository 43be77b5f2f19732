#!/bin/bash
# full GPU suite incl. the SimOTA consumer, bench (C4 with simota), sanitizer over the new kernels
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -25 $OUT/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --c5-max-log2 24 > $OUT/bench.json 2> $OUT/bench.err
stamp "bench exit $?"; tail -3 $OUT/bench.err; python - <<PY
import json
d = json.load(open('$OUT/bench.json'))
print('value', d['value'], 'frac', d['roofline']['frac'])
print(d['pairwise'])
PY
