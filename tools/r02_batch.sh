#!/bin/bash
# One gpurun call of the round-2 kernel work: GPU parity suite for every library given, then A/B timing of all of them,
# then (optional, NCU=tag) an ncu --set full capture of the first library.   usage: tools/r02_batch.sh <tag> <lib.so> [<lib.so> ...]
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
for lib in "$@"; do
  echo "== pytest $lib"
  OCEAN_LIB=$PWD/$lib timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_${tag}_$(basename $lib .so).log
done
echo "== A/B"
tools/ab_bench.sh godotoceanwaves_b200/libocean_prev.so "$@" 2>&1 | tee gpurun_out/ab_$tag.txt
if [ -n "$NCU" ]; then
  tools/ncu_capture.sh ${tag} $1
fi
