#!/bin/bash
# gpurun call for the pairwise / assign work: GPU tests, pairwise sweep, ncu of the kernels.
TAG=${1:-pair}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -30 $OUT/pytest_gpu.log
timeout -s KILL 300 python tools/sweep.py --only pairwise > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
echo "sweep exit $?"; tail -3 $OUT/sweep_pairwise.err; cat $OUT/sweep_pairwise.json
timeout -s KILL 300 python tools/latency.py > $OUT/latency.json 2> $OUT/latency.err
echo "latency exit $?"; cat $OUT/latency.json
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gd_pairwise_kernel \
  -s 3 -c 26 -o $OUT/prof_pairwise -f python tools/sweep.py --only pairwise > $OUT/ncu_pairwise.log 2>&1
echo "ncu exit $?"
