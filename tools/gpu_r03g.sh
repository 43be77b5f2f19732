#!/bin/bash
# 2-GPU call of the session: sharded parity (in-kernel cross-GPU sum and NCCL), bench at N=2; then on
# one GPU the full GPU suite and the bench with all sub-records (int64 index outputs of the fused
# assignment).   gpurun --gpus 2 -- bash tools/gpu_r03g.sh [tag]
TAG=${1:-r03g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29541 tools/check_sharded_nccl.py --fused > $OUT/sharded_fused.json 2> $OUT/sharded_fused.err
stamp "sharded check (fused) exit $?"; head -c 1200 $OUT/sharded_fused.json; echo; tail -3 $OUT/sharded_fused.err
timeout -s KILL 400 $TR --master-port 29542 tools/check_sharded_nccl.py > $OUT/sharded_nccl.json 2> $OUT/sharded_nccl.err
stamp "sharded check (nccl) exit $?"; head -c 600 $OUT/sharded_nccl.json; echo
timeout -s KILL 600 $TR --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > $OUT/bench_n2.json 2> $OUT/bench_n2.err
stamp "bench N=2 (in-kernel sum) exit $?"; tail -2 $OUT/bench_n2.err; head -c 900 $OUT/bench_n2.json; echo
timeout -s KILL 600 python bench.py --gpus 1 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
stamp "bench N=1 exit $?"; tail -2 $OUT/bench_n1.err
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
python - <<PY
import json
for f in ('bench_n1', 'bench_n2'):
    try:
        txt = open('$OUT/' + f + '.json').read()
        d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'] / 1e9, 2), 'G pairs/s, ms/step', round(d['ms_per_step'], 4), '|', d['config'].get('parallelism', '')[:80])
    if d.get('e2e'):
        print('   e2e', round(d['e2e']['value'] / 1e9, 3))
    if d.get('c3_strong'):
        print('   c3', d['c3_strong']['us_per_call_max_over_ranks'], 'us', d['c3_strong']['cross_gpu_sum'])
    if d.get('pairwise'):
        print('   pairwise', d['pairwise']['rows'], d['pairwise']['simota_gwd3d'])
PY
