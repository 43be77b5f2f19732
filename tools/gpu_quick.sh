#!/bin/bash
# Short gpurun call: smoke, GPU parity tests, bench, ncu of the pairwise kernel.
# Usage (repo root on the GPU box):  bash tools/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout -s KILL 400 python bench.py --steps 50 --warmup 5 > $OUT/bench_auto.json 2> $OUT/bench_auto.err
echo "bench exit $?"
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "bench reference exit $?"
timeout -s KILL 300 python tools/sweep.py --only pairwise > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
echo "sweep pairwise exit $?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_kernel \
  -s 3 -c 2 -o $OUT/prof_pairwise -f python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
echo "ncu pairwise exit $?"
ls -la $OUT
