#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench (variants), ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1

echo "== smoke" | tee $OUT/smoke.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/smoke.log

echo "== pytest -m gpu"
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log

echo "== bench"
timeout -s KILL 600 python bench.py --steps 100 --warmup 5 > $OUT/bench_auto.json 2> $OUT/bench_auto.err
echo "bench auto exit $?"
timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --variant bulk_r2 --no-e2e --no-cpu > $OUT/bench_r2.json 2> $OUT/bench_r2.err
echo "bench r2 exit $?"
GD_LOSS_B200_LIB=$PWD/mmdet3d_gaussian_b200/libgdloss_b200_precise.so timeout -s KILL 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu > $OUT/bench_precise.json 2> $OUT/bench_precise.err
echo "bench precise exit $?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "bench reference exit $?"
if [ -f tools/sweep.py ]; then
  timeout -s KILL 600 python tools/sweep.py > $OUT/sweep.json 2> $OUT/sweep.err
  echo "sweep exit $?"
fi

echo "== ncu"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gd_warp_kernel \
  -s 12 -c 4 -o $OUT/prof_bulk -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
