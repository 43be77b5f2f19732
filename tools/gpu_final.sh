#!/bin/bash
# Round evidence run on one GPU: smoke, GPU parity tests, bench (both arms), ncu full capture +
# launch list of the bench command, size sweep, head bench, latency.  bash tools/gpu_final.sh [tag]
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -5 $OUT/smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --eager-gpu > $OUT/bench_auto.json 2> $OUT/bench_auto.err
stamp "bench exit $?"; head -c 900 $OUT/bench_auto.json; echo
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
stamp "bench reference exit $?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gd_warp_kernel \
  -s 12 -c 4 -o $OUT/prof_bulk -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
stamp "ncu full exit $?"
if [ -f $OUT/prof_bulk.ncu-rep ]; then
  ncu -i $OUT/prof_bulk.ncu-rep --page raw --csv > $OUT/prof_bulk_raw.csv 2>/dev/null
  ncu -i $OUT/prof_bulk.ncu-rep --page source --csv --print-source sass > $OUT/prof_bulk_sass.csv 2>/dev/null
  SZ=$(stat -c %s $OUT/prof_bulk.ncu-rep); if [ $SZ -gt 30000000 ]; then rm $OUT/prof_bulk.ncu-rep; fi
fi
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launches.log 2>&1
stamp "ncu launches exit $?"
timeout -s KILL 300 python tools/sweep.py > $OUT/sweep.json 2> $OUT/sweep.err
stamp "sweep exit $?"
timeout -s KILL 200 python tools/bench_heads.py > $OUT/bench_heads.json 2> $OUT/bench_heads.err
stamp "bench_heads exit $?"
timeout -s KILL 200 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"
du -sh $OUT
