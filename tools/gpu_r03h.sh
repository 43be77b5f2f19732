#!/bin/bash
# Quick single-GPU check after a kernel change: assign / SimOTA / pairwise tests, then the bench
# (all sub-records) and a launch list of the SimOTA assigner.
TAG=${1:-r03h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*"; }
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
stamp "pytest exit $?"; tail -6 $OUT/pytest_gpu.log
timeout -s KILL 600 python bench.py --no-cpu > $OUT/bench_auto.json 2> $OUT/bench_auto.err
stamp "bench exit $?"; tail -2 $OUT/bench_auto.err
python - <<PY
import json
txt = open('$OUT/bench_auto.json').read()
d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('value', round(d['value'] / 1e9, 2), d['roofline']['frac'])
print(d['pairwise']['rows']); print(d['pairwise']['simota_gwd3d'])
PY
cat > /tmp/simota_probe.py <<PY
import torch, sys
sys.path.insert(0, '.')
from mmdet3d_gaussian_b200 import GDSimOTAAssigner, synth
a = synth.make_anchor_grid(200_000, 'waymo', device='cuda'); g = synth.make_targets(256, 'waymo', seed=5, device='cuda')
g[:, 0] = g[:, 0] * 2.0 - 70.0
asg = GDSimOTAAssigner(candidate_topk=10, loss_type='gwd3d', fun='log1p', tau=1.0)
for _ in range(3): asg.assign(a, g)
torch.cuda.synchronize()
PY
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/simota_launches.csv python /tmp/simota_probe.py > $OUT/simota_ncu.log 2>&1
stamp "ncu simota exit $?"
python - <<PY
import csv
rows = [r for r in csv.reader(open('$OUT/simota_launches.csv')) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
for r in rows[-14:]:
    print(r[ki][:70], r[vi])
PY
