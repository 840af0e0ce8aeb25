#!/bin/bash
# Round verification on one B200: GPU parity tests, smoke, both bench arms, launch list, one full capture of the
# dominant kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py --iters 2 --no-graph > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include 'profiled_iterations/' -k regex:attn_decode_tma -s 8 -c 1 -o gpurun_out/attn_decode_full -f \
    python tools/profile_step.py --iters 2 --skip-iters 250 --no-graph > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include 'profiled_iterations/' -k regex:gemm_ -s 4 -c 5 -o gpurun_out/gemm_full -f \
    python tools/profile_step.py --iters 2 --skip-iters 250 --no-graph > gpurun_out/ncu_gemm.log 2>&1
fi
timeout 200 python tools/timeline.py --out gpurun_out/tl_final.npy > gpurun_out/tl_final.txt 2>&1; tail -1 gpurun_out/tl_final.txt
ls -la gpurun_out | tail -5
