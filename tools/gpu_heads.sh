#!/bin/bash
# gpurun call for the head front ends: GPU tests + head bench.  bash tools/gpu_heads.sh [tag]
TAG=${1:-heads}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke exit $?"
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -30 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/bench_heads.py > $OUT/bench_heads.json 2> $OUT/bench_heads.err
echo "bench_heads exit $?"; tail -3 $OUT/bench_heads.err; cat $OUT/bench_heads.json
timeout -s KILL 300 python tools/latency.py --profile > $OUT/latency.json 2> $OUT/latency.err
echo "latency exit $?"; cat $OUT/latency.json; grep -A45 "cumulative" $OUT/latency.err | head -60
