#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the tensor-core codec kernels (persistent conv / resblock, 3 x TF32 encoder, cp.async LSTM)
# and one bf16 LM roll-out (decode attention incl. the split-stream merge).
set -u
mkdir -p gpurun_out
TESTS="tests/test_gpu_codec.py tests/test_gpu_lm.py::test_bf16_greedy_runs_and_is_deterministic"
for tool in memcheck racecheck; do
  SSRB_NO_GRAPH=1 timeout ${SAN_TIMEOUT:-700} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $TESTS -m gpu -x -q -k "tc_decode_waveform or tc_wmdecode_waveform or tc_encode or reloading or bf16_greedy" > gpurun_out/sanitize_codec_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_codec_$tool.log
  grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_codec_$tool.log
  tail -3 gpurun_out/sanitize_codec_$tool.log
done
