#!/bin/bash
# A/B of the decode-attention L2 run-ahead depth (SSRB_ATTN_L2_AHEAD) on the bench batch: decode-iteration time per setting.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attn_ops.py -m gpu -q > gpurun_out/pytest_attn.log 2>&1; tail -2 gpurun_out/pytest_attn.log
for a in ${AHEADS:-0 2 4 8 16}; do
  SSRB_ATTN_L2_AHEAD=$a timeout 300 python tools/small_batch_probe.py --batch 32 --skip 150 --iters 200 > gpurun_out/attn_ahead_$a.json 2>gpurun_out/attn_ahead_$a.err
  echo "ahead=$a $(cat gpurun_out/attn_ahead_$a.json)"
done
