#!/bin/bash
# One gpurun call, most valuable first: GPU parity tests, bench (both arms), knob sweep of
# the dominant kernel, ncu (launch list + full captures of the three kernel families),
# size sweep, head bench, latency.   bash tools/gpu_call_a.sh [tag]
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1

timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -5 $OUT/smoke.log

timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --durations=15 > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -40 $OUT/pytest_gpu.log

timeout -s KILL 400 python bench.py --steps 100 --warmup 5 > $OUT/bench_auto.json 2> $OUT/bench_auto.err
stamp "bench exit $?"; head -c 1500 $OUT/bench_auto.json; echo
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
stamp "bench reference exit $?"

timeout -s KILL 300 python tools/tune_sweep.py > $OUT/tune_sweep.json 2> $OUT/tune_sweep.err
stamp "tune sweep exit $?"; cat $OUT/tune_sweep.err | tail -25

# ncu: full capture of the dominant kernel (kld/none and bd/log1p launches) + exports
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gd_warp_kernel \
  -s 12 -c 4 -o $OUT/prof_bulk -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
stamp "ncu full exit $?"
if [ -f $OUT/prof_bulk.ncu-rep ]; then
  ncu -i $OUT/prof_bulk.ncu-rep --page raw --csv > $OUT/prof_bulk_raw.csv 2>/dev/null
  ncu -i $OUT/prof_bulk.ncu-rep --page details --csv > $OUT/prof_bulk_details.csv 2>/dev/null
  ncu -i $OUT/prof_bulk.ncu-rep --page source --csv --print-source sass > $OUT/prof_bulk_sass.csv 2>/dev/null
  ls -la $OUT/prof_bulk.ncu-rep
  SZ=$(stat -c %s $OUT/prof_bulk.ncu-rep); if [ $SZ -gt 30000000 ]; then rm $OUT/prof_bulk.ncu-rep; echo "rep too big, kept csv only"; fi
fi
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launches.log 2>&1
stamp "ncu launches exit $?"

timeout -s KILL 400 python tools/sweep.py > $OUT/sweep.json 2> $OUT/sweep.err
stamp "sweep exit $?"
timeout -s KILL 300 python tools/bench_heads.py > $OUT/bench_heads.json 2> $OUT/bench_heads.err
stamp "bench_heads exit $?"; cat $OUT/bench_heads.json | head -c 1500; echo
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"; cat $OUT/latency.json | head -c 1500; echo

# ncu of the pairwise and decoded kernels (raw csv only)
timeout -s KILL 300 ncu --set full --clock-control none -k regex:gd_pairwise -s 3 -c 6 -o $OUT/prof_pairwise -f \
  python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
stamp "ncu pairwise exit $?"
[ -f $OUT/prof_pairwise.ncu-rep ] && ncu -i $OUT/prof_pairwise.ncu-rep --page raw --csv > $OUT/prof_pairwise_raw.csv 2>/dev/null && rm $OUT/prof_pairwise.ncu-rep
timeout -s KILL 300 ncu --set full --clock-control none -k regex:decoded -s 2 -c 6 -o $OUT/prof_heads -f \
  python tools/bench_heads.py > $OUT/ncu_heads.log 2>&1
stamp "ncu heads exit $?"
[ -f $OUT/prof_heads.ncu-rep ] && ncu -i $OUT/prof_heads.ncu-rep --page raw --csv > $OUT/prof_heads_raw.csv 2>/dev/null && rm $OUT/prof_heads.ncu-rep
du -sh $OUT; ls -la $OUT
