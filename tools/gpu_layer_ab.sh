#!/bin/bash
# Round-2 bring-up of the experimental persistent per-layer GEMM kernel (csrc/gemm_layer.cu) on one B200:
#   1. the gated parity tests (every case in a child process under a timeout: a broken grid barrier hangs, it does not crash);
#   2. if green, a bench A/B of the default chain against SSRB_LAYER_KERNEL=1 with the in-kernel timeline of both.
# Usage:  gpurun --timeout 1500 -- 'bash tools/gpu_layer_ab.sh'
set -u
mkdir -p gpurun_out
SSRB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_layer_kernel.py -m gpu -x -q > gpurun_out/pytest_layer.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_layer.log
tail -15 gpurun_out/pytest_layer.log
if [ $rc -ne 0 ]; then echo "layer kernel parity failed: no A/B"; exit $rc; fi
bash tools/gpu_ab.sh base:SSRB_LAYER_KERNEL=0,TL=1 layer:SSRB_LAYER_KERNEL=1,TL=1
