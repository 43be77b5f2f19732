#!/bin/bash
# 2-GPU call: in-kernel cross-GPU sum over peer memory (parity vs the fp64 oracle of the whole
# batch, identical bits on both ranks), bench at N=2 with it and with NCCL; then (one GPU)
# the pairwise tests + sweep of the trimmed kernel.
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29541 tools/check_sharded_nccl.py --fused > $OUT/sharded_fused.json 2> $OUT/sharded_fused.err
stamp "sharded check (fused) exit $?"; cat $OUT/sharded_fused.json; tail -5 $OUT/sharded_fused.err
timeout -s KILL 400 $TR --master-port 29542 tools/check_sharded_nccl.py > $OUT/sharded_nccl.json 2> $OUT/sharded_nccl.err
stamp "sharded check (nccl) exit $?"; cat $OUT/sharded_nccl.json | head -c 1500
timeout -s KILL 600 $TR --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --verbose > $OUT/bench_n2.json 2> $OUT/bench_n2.err
stamp "bench N=2 (in-kernel sum) exit $?"; tail -4 $OUT/bench_n2.err
timeout -s KILL 600 $TR --master-port 29544 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --nccl --c5-max-log2 24 > $OUT/bench_n2_nccl.json 2> $OUT/bench_n2_nccl.err
stamp "bench N=2 (NCCL) exit $?"; tail -2 $OUT/bench_n2_nccl.err
python - <<PY
import json
for f in ('bench_n2', 'bench_n2_nccl'):
    try:
        d = json.load(open('$OUT/' + f + '.json'))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], d['config']['parallelism'])
    print('  e2e', d['e2e'] and d['e2e']['value'], d['e2e'] and d['e2e'].get('numa'))
    print('  c3', d.get('c3_strong'))
    for r in d.get('c5', {}).get('rows', [])[:10]:
        print('  ', r)
PY
timeout -s KILL 600 python -m pytest tests/test_gpu_packed.py tests/test_gpu_assign.py tests/test_eval_affinity.py tests/test_gpu_parity.py -m gpu -q --timeout=600 -p no:cacheprovider -k "pairwise or assign or affinity or packed" > $OUT/pytest_pairwise.log 2>&1
stamp "pytest (pairwise surface) exit $?"; tail -4 $OUT/pytest_pairwise.log
timeout -s KILL 300 python tools/sweep.py --only pairwise --cpl1 > $OUT/sweep_pairwise.json 2> $OUT/sweep_pairwise.err
stamp "sweep pairwise exit $?"; cat $OUT/sweep_pairwise.json
