#!/bin/bash
# A/B of environment switches on the bench workload (no CPU baseline):  tools/gpu_ab.sh name:VAR=val,VAR2=val ...
# A variant whose list holds TL=1 also dumps its in-kernel timeline.  SSRB_LIB=<path> selects another build of the library.
# Optional first: PYTEST="tests/test_gpu_modes.py ..." to run tests before; TIMELINE=1 dumps the in-kernel timeline of each variant.
set -u
mkdir -p gpurun_out
if [ -n "${PYTEST:-}" ]; then
  timeout 900 python -m pytest $PYTEST -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ab.log
  tail -4 gpurun_out/pytest_ab.log
fi
run() { name=$1; shift
  if [ "${TIMELINE:-0}" = "1" ] || [[ " $* " == *" TL=1 "* ]]; then env "$@" timeout 300 python tools/timeline.py --out gpurun_out/tl_$name.npy > gpurun_out/tl_$name.txt 2>&1; tail -1 gpurun_out/tl_$name.txt; fi
  env "$@" timeout 400 python bench.py --no-cpu-baseline --steps 2 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    b=d["roofline"]["breakdown"]
    print("$name", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "iter_ms", round(d["roofline"]["iteration_ms_avg"],4),
          "attn", round(b["attention_us"]), "gemm", round(b["gemm_us"]), "ln", round(b["layernorm_us"]), "phases", {k:round(v,1) for k,v in d["e2e"]["phase_ms_per_step"].items()})
except Exception as e:
    print("$name failed", e)
PY
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name ${envs//,/ }
done
