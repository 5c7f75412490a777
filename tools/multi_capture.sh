timeout 300 python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_r01b_n2.json 2> gpurun_out/bench_r01b_n2.err; cut -c1-300 gpurun_out/bench_r01b_n2.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_r01b_ref.json 2> gpurun_out/bench_r01b_ref.err; cut -c1-200 gpurun_out/bench_r01b_ref.json; grep -o '"cpu_baseline.*' gpurun_out/bench_r01b_ref.json | cut -c1-200
