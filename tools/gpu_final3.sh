#!/bin/bash
# Evidence run of the session on one GPU: smoke, full GPU suite, bench (both arms, all sub-records),
# ncu full captures (fused loss kernel; pairwise matrix + row-lane kernels) + launch list of the
# bench command, layouts, latency, heads, pairwise sweeps (A/B of the reductions), sanitizer.
#   bash tools/gpu_final3.sh [tag]
TAG=${1:-r03z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -6 $OUT/smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/sweep.py --only pairwise --cpl1 > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
stamp "sweep pairwise exit $?"; cat $OUT/sweep_pairwise.json; echo
GD_B200_PAIR_ROWLANE=0 timeout -s KILL 300 python tools/sweep.py --only pairwise > $OUT/sweep_pairwise_collane.json 2> $OUT/sweep_pairwise_collane.err
stamp "sweep pairwise (column-lane reductions) exit $?"
timeout -s KILL 700 python bench.py --eager-gpu > $OUT/bench_auto.json 2> $OUT/bench_auto.err
stamp "bench exit $?"; head -c 1200 $OUT/bench_auto.json; echo; tail -2 $OUT/bench_auto.err
timeout -s KILL 400 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
stamp "bench reference exit $?"; head -c 600 $OUT/bench_reference.json; echo
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:gd_warp_kernel \
  -c 10 -o $OUT/prof_bulk -f python tools/ncu_target.py > $OUT/ncu_full.log 2>&1
stamp "ncu full exit $?"; tail -2 $OUT/ncu_full.log
if [ -f $OUT/prof_bulk.ncu-rep ]; then
  ncu -i $OUT/prof_bulk.ncu-rep --page raw --csv > $OUT/prof_bulk_raw.csv 2>/dev/null
  SZ=$(stat -c %s $OUT/prof_bulk.ncu-rep); if [ $SZ -gt 30000000 ]; then rm $OUT/prof_bulk.ncu-rep; fi
fi
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > $OUT/ncu_launches.log 2>&1
stamp "ncu launches exit $?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_kernel \
  -c 2 -o $OUT/prof_pairwise_matrix -f python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
stamp "ncu pairwise matrix exit $?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_rowlane \
  -s 28 -c 2 -o $OUT/prof_pairwise_rowlane -f python tools/sweep.py --only pairwise > $OUT/ncu_rowlane.log 2>&1
stamp "ncu pairwise rowlane exit $?"
for f in prof_pairwise_matrix prof_pairwise_rowlane; do
  if [ -f $OUT/$f.ncu-rep ]; then
    ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
    SZ=$(stat -c %s $OUT/$f.ncu-rep); if [ $SZ -gt 12000000 ]; then rm $OUT/$f.ncu-rep; fi
  fi
done
timeout -s KILL 300 python tools/bench_layouts.py > $OUT/layouts.json 2> $OUT/layouts.err
stamp "layouts exit $?"
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"; cat $OUT/latency.json
timeout -s KILL 300 python tools/bench_heads.py > $OUT/bench_heads.json 2> $OUT/bench_heads.err
stamp "bench_heads exit $?"
bash tools/gpu_sanitize.sh ${TAG}_san loss strided pairwise
stamp "sanitizer done"
du -sh $OUT
