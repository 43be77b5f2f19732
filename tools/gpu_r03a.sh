#!/bin/bash
# Round-2 (session 3) first call: GPU suite, pairwise A/B (row-lane fused reductions, fast
# matrix loop), bench with the in-launch early-return flag, ncu of the pairwise kernels.
TAG=${1:-r03a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -4 $OUT/smoke.log
timeout -s KILL 300 python tools/sweep.py --only pairwise --cpl1 > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
stamp "sweep pairwise exit $?"; tail -2 $OUT/sweep_pairwise.err; cat $OUT/sweep_pairwise.json; echo
GD_B200_PAIR_ROWLANE=0 timeout -s KILL 300 python tools/sweep.py --only pairwise > $OUT/sweep_pairwise_collane.json 2> $OUT/sweep_pairwise_collane.err
stamp "sweep pairwise (column-lane reductions) exit $?"; cat $OUT/sweep_pairwise_collane.json; echo
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -15 $OUT/pytest_gpu.log
timeout -s KILL 600 python bench.py --no-extras > $OUT/bench_auto.json 2> $OUT/bench_auto.err
stamp "bench exit $?"; head -c 1500 $OUT/bench_auto.json; echo; tail -2 $OUT/bench_auto.err
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
stamp "latency exit $?"; cat $OUT/latency.json
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_kernel \
  -c 2 -o $OUT/prof_pairwise_matrix -f python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
stamp "ncu matrix exit $?"; tail -2 $OUT/ncu_pairwise.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_rowlane \
  -s 28 -c 2 -o $OUT/prof_pairwise_rowlane -f python tools/sweep.py --only pairwise > $OUT/ncu_rowlane.log 2>&1
stamp "ncu rowlane exit $?"; tail -2 $OUT/ncu_rowlane.log
for f in prof_pairwise_matrix prof_pairwise_rowlane; do
  if [ -f $OUT/$f.ncu-rep ]; then
    ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
    SZ=$(stat -c %s $OUT/$f.ncu-rep); if [ $SZ -gt 30000000 ]; then rm $OUT/$f.ncu-rep; fi
  fi
done
du -sh $OUT
