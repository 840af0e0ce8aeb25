#!/bin/bash
# Codec verification + timing on one B200: the codec parity tests, the 32 x 10 s timing (tools/profile_codec.py) and the launch list.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_z_codec_shapes.py tests/test_gpu_pin_bench_geometry.py -m gpu -q -k "not rollout and not teacher" > gpurun_out/pytest_codec.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_codec.log
tail -15 gpurun_out/pytest_codec.log
python tools/profile_codec.py --batch 32 --chunk 32 > gpurun_out/codec_time.txt 2>&1; cat gpurun_out/codec_time.txt
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_codec.csv python tools/profile_codec.py --batch 32 --chunk 32 > gpurun_out/ncu_codec.log 2>&1; tail -2 gpurun_out/ncu_codec.log
fi
