#!/bin/bash
# 8-GPU re-check of the weak-scaling headline after the session's ABI change: bench at N = 8 and 4
# with the in-kernel cross-GPU sum, sharded parity at 8 ranks.   gpurun --gpus 8 -- bash tools/gpu_r03s.sh
TAG=${1:-r03s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
run() {  # n port extra...
  local n=$1 port=$2; shift 2
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $n --steps 30 --warmup 5 --no-cpu "$@"
}
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
  --master-port 29551 tools/check_sharded_nccl.py --fused > $OUT/sharded_fused_n8.json 2> $OUT/sharded_fused_n8.err
stamp "sharded check N=8 (fused) exit $?"; tail -c 600 $OUT/sharded_fused_n8.json; echo
run 8 29552 --c5-max-log2 26 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
stamp "bench N=8 (in-kernel sum) exit $?"; tail -2 $OUT/bench_n8.err
run 4 29553 --c5-max-log2 24 > $OUT/bench_n4.json 2> $OUT/bench_n4.err
stamp "bench N=4 (in-kernel sum) exit $?"
run 1 29555 --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err
stamp "bench N=1 exit $?"
python - <<PY
import json
for f in ('bench_n1', 'bench_n4', 'bench_n8'):
    try:
        txt = open('$OUT/' + f + '.json').read()
        d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value'] / 1e9, 2), 'G pairs/s, ms/step', round(d['ms_per_step'], 4), '|', d['config'].get('parallelism', '')[:70])
    if d.get('e2e'):
        print('   e2e', round(d['e2e']['value'] / 1e9, 3), 'h2d GB/s per GPU', round(d['e2e']['h2d_GBps_per_gpu'], 1))
    if d.get('c3_strong'):
        print('   c3', d['c3_strong']['us_per_call_max_over_ranks'], 'us', d['c3_strong']['cross_gpu_sum'])
PY
