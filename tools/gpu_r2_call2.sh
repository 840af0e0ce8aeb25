#!/bin/bash
# Round 2, second GPU call: the tests fixed after call 1, the experimental kernels' full parity lists, then their same-box A/B.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attn_ops.py tests/test_gpu_pin_bench_geometry.py tests/test_gpu_lm.py tests/test_gpu_codec.py -m gpu -q > gpurun_out/pytest_fix.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fix.log
tail -8 gpurun_out/pytest_fix.log
SSRB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_layer_kernel.py -m gpu -q > gpurun_out/pytest_layer.log 2>&1
rcl=$?; echo "pytest rc=$rcl" >> gpurun_out/pytest_layer.log; tail -12 gpurun_out/pytest_layer.log
SSRB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_flat2.py -m gpu -q > gpurun_out/pytest_flat2.log 2>&1
rcf=$?; echo "pytest rc=$rcf" >> gpurun_out/pytest_flat2.log; tail -12 gpurun_out/pytest_flat2.log
bash tools/gpu_ab.sh base:SSRB_LAYER_KERNEL=0,TL=1 layer:SSRB_LAYER_KERNEL=1,TL=1
bash tools/gpu_ab.sh pair:SSRB_FLAT_2CTA=1
