#!/bin/bash
# Round 2, first GPU call: the whole GPU suite (incl. the new bench-geometry pins and the formerly gated edge cases), then the
# bring-up + A/B of the two experimental kernels, then the pure-read HBM microbenchmark.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
(cd tools/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu && timeout 120 ./read_bw) > gpurun_out/read_bw.txt 2>&1
tail -12 gpurun_out/read_bw.txt
bash tools/gpu_layer_ab.sh
