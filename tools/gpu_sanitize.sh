#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family (race / memory / sync checks).  One B200, ~10 min:
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh'
# memcheck: out-of-bounds and misaligned accesses (also shared / distributed shared memory); racecheck: shared-memory hazards
# inside a CTA; synccheck: divergent barriers.  The tests chosen keep every kernel of the decode chain, the prefill, the
# sampler and the codec in play at sizes a 10-50x sanitizer slowdown can afford.  SSRB_NO_GRAPH=1: the sanitizer attributes
# errors to launches, and graph replays hide the launch site.
set -u
mkdir -p gpurun_out
TESTS="tests/test_gpu_lm.py::test_fp32_tokens_match_reference_golden tests/test_gpu_lm.py::test_bf16_greedy_runs_and_is_deterministic tests/test_gpu_lm.py::test_continuous_batching_bf16_runs tests/test_gpu_gemm.py tests/test_gpu_codec.py"
for tool in memcheck racecheck synccheck; do
  SSRB_NO_GRAPH=1 timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest $TESTS -m gpu -x -q -k "not fullsize and not match_generic" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_$tool.log
  tail -3 gpurun_out/sanitize_$tool.log
done
if [ "${SSRB_EXPERIMENTAL:-0}" = "1" ]; then   # the experimental layer kernel: one small case under memcheck (watchdog raised: 50x slowdown)
  SSRB_NO_GRAPH=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 \
      python -m pytest "tests/test_gpu_zz_layer_kernel.py::test_layer_kernel_matches_per_gemm_chain" -m gpu -x -q -k "M16-D512" > gpurun_out/sanitize_layer.log 2>&1
  echo "layer memcheck rc=$?" | tee -a gpurun_out/sanitize_layer.log
fi
