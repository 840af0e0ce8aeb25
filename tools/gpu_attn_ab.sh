#!/bin/bash
# A/B of the decode attention's on-demand share (SSRB_ATTN_DYN per mille, SSRB_ATTN_UNIT tiles) on the bench batch.
# CFGS="dyn:unit ..." ; TRACE=dyn:unit additionally prints the per-CTA phase trace of that setting.
set -u
mkdir -p gpurun_out
for cfg in ${CFGS:-0:2 125:2}; do
  d=${cfg%%:*}; u=${cfg##*:}
  if [ "${SKIP_TESTS:-0}" != "1" ]; then
    SSRB_ATTN_DYN=$d SSRB_ATTN_UNIT=$u timeout 600 python -m pytest tests/test_gpu_attn_ops.py -m gpu -q -x > gpurun_out/pytest_attn_${d}_$u.log 2>&1; echo "dyn=$d unit=$u $(tail -1 gpurun_out/pytest_attn_${d}_$u.log)"
  fi
  SSRB_ATTN_DYN=$d SSRB_ATTN_UNIT=$u timeout 300 python tools/small_batch_probe.py --batch 32 --skip 150 --iters 200 > gpurun_out/attn_dyn_${d}_$u.json 2>gpurun_out/attn_dyn_${d}_$u.err
  echo "dyn=$d unit=$u $(cat gpurun_out/attn_dyn_${d}_$u.json | cut -c1-110) $(tail -1 gpurun_out/attn_dyn_${d}_$u.err | cut -c1-200)"
done
if [ -n "${TRACE:-}" ]; then
  d=${TRACE%%:*}; u=${TRACE##*:}
  SSRB_ATTN_DYN=$d SSRB_ATTN_UNIT=$u timeout 300 python tools/timeline.py --out /tmp/tl_dyn.npy > gpurun_out/tl_dyn.txt 2>&1; python tools/attn_dyn_trace.py /tmp/tl_dyn.npy
fi
