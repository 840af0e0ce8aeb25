#!/bin/bash
# pipeline + codec tests and one bench line (host-path changes)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_codec.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["phase_ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
print({k:(v["e2e_codec_tokens_per_sec"], v["ms_per_step"]) for k,v in d["other_batches"].items()}, {k:(v["e2e_codec_tokens_per_sec"]) for k,v in d["other_configs"].items()})
PY
