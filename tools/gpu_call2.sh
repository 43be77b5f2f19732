#!/bin/bash
# Round-2 validation call: full GPU suite on the new default build (packed+diet kernel, torch
# C++ shim, device-side early return, strided bulk pipeline), latency, layouts, bench, sanitizer.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -5 $OUT/smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -15 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"; cat $OUT/latency.json; tail -3 $OUT/latency.err
timeout -s KILL 300 python tools/bench_layouts.py > $OUT/layouts.json 2> $OUT/layouts.err
stamp "layouts exit $?"; cat $OUT/layouts.err | tail -20
timeout -s KILL 400 python bench.py --steps 50 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
stamp "bench exit $?"; cat $OUT/bench.json | head -c 3000; tail -3 $OUT/bench.err
timeout -s KILL 300 python tools/bench_heads.py > $OUT/bench_heads.json 2> $OUT/bench_heads.err
stamp "bench_heads exit $?"; cat $OUT/bench_heads.json | head -c 1500
bash tools/gpu_sanitize.sh ${TAG}_san loss strided
stamp "sanitizer done"
