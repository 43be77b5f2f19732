#!/bin/bash
# CPU side (no GPU needed): build the A/B libraries the next GPU call will measure.
#   bash tools/prepare_variants.sh [bits ...]        default: 256 768 1024 1792 1920
# Each becomes mmdet3d_gaussian_b200/libgdloss_b200_v<bits>.so (git-ignored; it travels to
# the GPU box with the snapshot).  Bits: see tools/ab_variants.py.
set -e
cd "$(dirname "$0")/.."
BITS=${@:-256 768 1024 1792 1920}
python -m mmdet3d_gaussian_b200.build_ext            # production library
python -m mmdet3d_gaussian_b200.build_ext --precise
python -m mmdet3d_gaussian_b200.build_ext --tune
for b in $BITS; do
  python -m mmdet3d_gaussian_b200.build_ext --variant $b
done
python tools/pipe_probe.py --build-only
ls -la mmdet3d_gaussian_b200/*.so tools/micro/pipe_probe
