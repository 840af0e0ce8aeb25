#!/bin/bash
# A/B of the decode-GEMM ring depth at 16 activation rows (batch 1 / 8): build the variants first, e.g.
#   python ssr-speech_b200/build.py --variant ssr-speech_b200/libssr_q16_4.so DEC_STAGES_Q16=4
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -2
for lib in "" ssr-speech_b200/libssr_q16_5.so ssr-speech_b200/libssr_q16_4.so; do
  for b in 1 8; do
    SSRB_LIB=$lib timeout 300 python tools/small_batch_probe.py --batch $b 2>/dev/null | cut -c1-130 | sed "s|^|lib=${lib:-default(6)} |"
  done
done
