#!/bin/bash
# Validation of the in-kernel early return / single-kernel probe / new bench: full GPU suite
# (no -x), latency, the new bench.py (both arms), layouts.
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -2 $OUT/smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -25 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"; cat $OUT/latency.json; tail -3 $OUT/latency.err
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --verbose > $OUT/bench.json 2> $OUT/bench.err
stamp "bench exit $?"; head -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout -s KILL 400 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
stamp "bench reference exit $?"; cat $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
