#!/bin/bash
# 2-GPU check of the grid policy: under torchrun (WORLD_SIZE=2) the ranks keep one CTA per SM; alone, 128.
TAG=${1:-r03u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-extras > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "bench N=2 exit $?"
timeout -s KILL 400 python bench.py --gpus 1 --no-cpu --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err
echo "bench N=1 exit $?"
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -2 $OUT/pytest_gpu.log
python - <<PY
import json
for f in ('bench_n1', 'bench_n2'):
    txt = open('$OUT/' + f + '.json').read()
    d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(f, 'value', round(d['value'] / 1e9, 2), 'ms/step', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4))
PY
