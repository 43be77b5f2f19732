#!/bin/bash
# One GPU call that decides the next default of the fused kernel (run after
# tools/prepare_variants.sh built the libraries on the CPU side):
#   1. smoke + the GPU parity suite on the production library,
#   2. the SAME parity suite against every variant library (GD_LOSS_B200_LIB) plus the
#      opt-in packed tests,
#   3. interleaved A/B timing of all variants + sustained GB/s, SM clock and power of each,
#   4. the data-movement pipeline probe and the knob/power probe of the shipped kernel.
# bash tools/gpu_variants.sh [tag] [bits ...]
TAG=${1:-r02a}
shift
BITS=${@:-256 768 1024 1792 1920}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
stamp "smoke exit $?"; tail -4 $OUT/smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest (production) exit $?"; tail -2 $OUT/pytest_gpu.log
HAVE=""
for b in $BITS; do
  LIB=$PWD/mmdet3d_gaussian_b200/libgdloss_b200_v$b.so
  if [ ! -f $LIB ]; then stamp "variant $b: library missing, skipped"; continue; fi
  HAVE="$HAVE $b"
  GD_LOSS_B200_LIB=$LIB GD_B200_TEST_EXPERIMENTAL=1 timeout -s KILL 600 python -m pytest \
    tests/test_gpu_parity.py tests/test_gpu_packed.py -m gpu -q --timeout=300 -p no:cacheprovider \
    > $OUT/pytest_v$b.log 2>&1
  stamp "pytest (variant $b) exit $?"; tail -2 $OUT/pytest_v$b.log
done
GD_B200_TEST_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_gpu_packed.py -m gpu -q \
  --timeout=300 -p no:cacheprovider > $OUT/pytest_packed.log 2>&1
stamp "pytest (packed, production library) exit $?"; tail -2 $OUT/pytest_packed.log
timeout -s KILL 600 python tools/ab_variants.py prod $HAVE --power > $OUT/ab_variants.json 2> $OUT/ab_variants.err
stamp "ab_variants exit $?"; tail -12 $OUT/ab_variants.err
timeout -s KILL 200 python tools/pipe_probe.py 2 24 > $OUT/pipe_probe.json 2> $OUT/pipe_probe.err
stamp "pipe_probe exit $?"; cat $OUT/pipe_probe.err | tail -8
timeout -s KILL 200 python tools/power_probe.py > $OUT/power_probe.json 2> $OUT/power_probe.err
stamp "power_probe exit $?"
# opt-in packed pairwise kernel: its own gated tests, the whole pairwise surface routed
# through it (environment switch), then timing next to the scalar kernel
GD_B200_TEST_EXPERIMENTAL=1 timeout -s KILL 400 python -m pytest tests/test_gpu_packed.py -m gpu -q \
  -k pairwise --timeout=300 -p no:cacheprovider > $OUT/pytest_pairwise_packed.log 2>&1
stamp "pytest (packed pairwise) exit $?"; tail -2 $OUT/pytest_pairwise_packed.log
GD_B200_PAIRWISE_PACKED=1 timeout -s KILL 400 python -m pytest tests/test_gpu_assign.py \
  tests/test_eval_affinity.py -m gpu -q --timeout=300 -p no:cacheprovider > $OUT/pytest_pairwise_forced.log 2>&1
stamp "pytest (pairwise surface, packed forced) exit $?"; tail -2 $OUT/pytest_pairwise_forced.log
timeout -s KILL 300 python tools/sweep.py --only pairwise --packed > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
stamp "sweep pairwise exit $?"
du -sh $OUT
# host-buffer pipeline: chunk size of the e2e path (tail of the last D2H vs per-chunk overhead)
for L in 17 18 19 20 21; do
  timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-chunk-log2 $L \
    > $OUT/bench_e2e_chunk$L.json 2> $OUT/bench_e2e_chunk$L.err
  python - <<PY
import json
try:
    d = json.load(open('$OUT/bench_e2e_chunk$L.json'))
    print('e2e chunk 2^$L:', round(d['e2e']['value'] / 1e9, 4), 'G pairs/s')
except Exception as e:
    print('e2e chunk 2^$L: failed', e)
PY
done
stamp "e2e chunk sweep done"
