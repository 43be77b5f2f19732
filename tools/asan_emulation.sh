#!/bin/bash
# CPU only: the kernels' source under the execution-model emulation (tests/host_math/cuda_emul.h),
# built with AddressSanitizer + UBSan -- catches out-of-bounds reads / writes of global buffers
# and static shared memory, misaligned vector accesses and signed overflow in the index math.
#   bash tools/asan_emulation.sh [GD_TUNE_DEFAULT bits ...]       default: 0 1920
set -e
cd "$(dirname "$0")/.."
OUT=$(mktemp -d)
FLAGS="-O1 -g -std=c++20 -ffp-contract=off -pthread -w -fsanitize=address,undefined -fno-omit-frame-pointer -I /usr/local/cuda/include -I include -x c++"
g++ $FLAGS tests/host_math/pairwise_emul.cpp tests/host_math/asan_pairwise_main.cpp -o $OUT/asan_pairwise
$OUT/asan_pairwise
for bits in ${@:-0 1920}; do
  g++ $FLAGS -DGD_TUNE_DEFAULT=$bits tests/host_math/loss_emul.cpp tests/host_math/asan_loss_main.cpp -o $OUT/asan_loss_$bits
  echo "== loss kernels, GD_TUNE_DEFAULT=$bits"
  $OUT/asan_loss_$bits
done
echo "sanitizers: clean"
