#!/bin/bash
# Everything the round's record needs from ONE single-GPU box (run as `gpurun -- bash tools/final_capture.sh <tag>`): parity suite, the
# other BASELINE configs, smoke, both bench arms, ncu launch list, ncu --set full capture of the dominant kernel, DRAM bytes, sanitizers.
# Every step has its own time-out: a hung step must not eat the box time of the others.  (The round-2b capture ran exactly these steps.)
tag=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_$tag.log
timeout 300 python tools/run_configs.py > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; cut -c1-200 gpurun_out/configs_$tag.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_$tag.log
timeout 400 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-700 gpurun_out/bench_${tag}_n1.json
timeout 300 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; cut -c1-300 gpurun_out/bench_${tag}_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_update_persistent --launch-skip 10 --launch-count 1 -f \
  -o gpurun_out/prof_${tag}_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
timeout 200 bash tools/ncu_dram.sh godotoceanwaves_b200/libocean.so | tee gpurun_out/dram_$tag.txt
timeout 420 bash tools/sanitize.sh 2>&1 | tee gpurun_out/sanitize_$tag.log
# afterwards, here: python tools/summarize_ncu.py gpurun_out/prof_${tag}_final.ncu-rep profiles/${tag}_final_persistent ; copy the logs into profiles/
