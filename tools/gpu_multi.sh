#!/bin/bash
# Multi-GPU bench on one box: bash tools/gpu_multi.sh <tag> "<N list>"   (gpurun --gpus 8)
TAG=${1:-r01m}
NS=${2:-"1 2 4 8"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
PORT=29517
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout -s KILL 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu > $OUT/scale_n1.json 2> $OUT/scale_n1.err
  else
    NCCL_DEBUG=WARN timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 100 --warmup 5 --no-cpu \
      > $OUT/scale_n$N.json 2> $OUT/scale_n$N.err
  fi
  echo "N=$N exit $?"; tail -c 600 $OUT/scale_n$N.json | head -c 300; echo
  PORT=$((PORT+1))
done
# GPU parity tests must also pass when another device is current / visible
timeout -s KILL 600 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider -k "golden or tail or strided" > $OUT/pytest_multi.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/pytest_multi.log
