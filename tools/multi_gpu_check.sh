#!/bin/bash
# Multi-GPU evidence in ONE gpurun --gpus N call: the sharded == single-GPU parity test, then bench.py at every power of two up
# to N for the weak (cfg2) and the strong (cfg4: 1024x1024 x 8 split over the GPUs) workloads.   usage: tools/multi_gpu_check.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== $N GPUs"; nvidia-smi topo -m 2>/dev/null | head -14 | tee gpurun_out/topo_$tag.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k multi_gpu --tb=short 2>&1 | tail -60 | tee gpurun_out/pytest_multigpu_$tag.log
run() {   # $1 = gpus, $2 = workload
  local out=gpurun_out/bench_${tag}_$2_n$1.json
  if [ "$1" = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --workload $2 --no-cpu-baseline > $out 2> gpurun_out/bench_${tag}_$2_n$1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 --workload $2 > $out 2> gpurun_out/bench_${tag}_$2_n$1.err; fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], "value %.4g %s, %.4f ms/step, frac %.3f, e2e %.4g (%.2f ms/step), scaling %s, sharding_check %s, bound %s" % (
        d["value"], d["unit"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["scaling"],
        d["sharding_check"]["ok"], d["host_binding"].get("bound")))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
g=1
while [ $g -le $N ]; do run $g cfg2; run $g cfg4-strong; g=$((g * 2)); done
