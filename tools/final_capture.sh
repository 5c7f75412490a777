#!/bin/bash
# Everything the round's record needs from ONE single-GPU box (run under gpurun): tests, smoke, bench arms, ncu launch list,
# ncu --set full captures of the dominant kernel and of the spectrum kernel, DRAM bytes, the other BASELINE configs, sanitizers.
#   usage: tools/final_capture.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-400 gpurun_out/bench_${tag}_n1.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; cut -c1-300 gpurun_out/bench_${tag}_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_persistent --launch-skip 10 --launch-count 1 -f \
  -o gpurun_out/prof_${tag}_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spectrum_compute --launch-skip 2 --launch-count 1 -f \
  -o gpurun_out/prof_${tag}_spectrum python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_spectrum_$tag.log 2>&1
timeout 600 python tools/run_configs.py > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; cut -c1-260 gpurun_out/configs_$tag.jsonl
tools/sanitize.sh 2>&1 | tee gpurun_out/sanitize_$tag.log
tools/ncu_dram.sh godotoceanwaves_b200/libocean.so | tee gpurun_out/dram_$tag.txt
