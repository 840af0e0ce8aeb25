#!/bin/bash
# Bring-up / A/B of the persistent layer kernel and the CTA-pair prefill GEMM on one B200 (results: profiles/r02a_summary.md — the
# layer kernel lost and stays off, the CTA-pair GEMM won and is the default):
#   A. persistent per-layer decode GEMM kernel (csrc/gemm_layer.cu, SSRB_LAYER_KERNEL=1)
#   B. CTA-pair prefill GEMM (csrc/gemm_flat2.cu, SSRB_FLAT_2CTA=1)
# For each: the gated parity tests first (every case in a child process under a timeout; the kernels' spins trap after 2 s, so a
# protocol bug fails a launch instead of hanging the box), and only if they are green a same-box bench A/B with in-kernel timelines.
# Usage:  gpurun --timeout 2400 -- 'bash tools/gpu_layer_ab.sh'        (SECTIONS="A" or "B" to run one)
set -u
mkdir -p gpurun_out
SECTIONS=${SECTIONS:-"A B"}
if [[ " $SECTIONS " == *" A "* ]]; then
  SSRB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_zz_layer_kernel.py -m gpu -x -q > gpurun_out/pytest_layer.log 2>&1
  rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_layer.log
  tail -15 gpurun_out/pytest_layer.log
  if [ $rc -eq 0 ]; then bash tools/gpu_ab.sh base:SSRB_LAYER_KERNEL=0,TL=1 layer:SSRB_LAYER_KERNEL=1,TL=1
  else echo "layer kernel parity failed: no A/B"; fi
fi
if [[ " $SECTIONS " == *" B "* ]]; then
  SSRB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_flat2.py -m gpu -x -q > gpurun_out/pytest_flat2.log 2>&1
  rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_flat2.log
  tail -8 gpurun_out/pytest_flat2.log
  if [ $rc -eq 0 ]; then bash tools/gpu_ab.sh base2:SSRB_FLAT_2CTA=0 pair:SSRB_FLAT_2CTA=1      # compare "lm_prefill_ms" of the two lines
  else echo "CTA-pair GEMM parity failed: no A/B"; fi
fi
