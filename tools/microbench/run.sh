#!/bin/bash
# builds and runs the pipe-rate microbenchmark on the GPU box (under gpurun); output -> gpurun_out/pipe_rates.txt
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/pipe_rates pipe_rates.cu && /tmp/pipe_rates | tee ../../gpurun_out/pipe_rates.txt
