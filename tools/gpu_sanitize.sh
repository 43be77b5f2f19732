#!/bin/bash
# compute-sanitizer over the library's kernels: memcheck, racecheck,
# synccheck on tools/sanitize_target.py (initcheck needs >15 min on this workload: left out).  Only OUR library is instrumented
# (--kernel-name kns=3gdk: the mangled namespace of every kernel in csrc/).
#   bash tools/gpu_sanitize.sh [tag] [workloads ...]
TAG=${1:-r02s}
shift
WHAT=${@:-loss strided pairwise}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  [ $tool = initcheck ] && extra=""
  timeout -s KILL 900 $CS --tool $tool $extra --kernel-name kns=3gdk --print-limit 30 \
    --log-file $OUT/sanitizer_$tool.log python tools/sanitize_target.py $WHAT \
    > $OUT/sanitizer_$tool.out 2>&1
  echo "$tool exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_$tool.log | tail -1)"
  tail -2 $OUT/sanitizer_$tool.out
done
