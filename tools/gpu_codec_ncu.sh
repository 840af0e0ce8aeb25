#!/bin/bash
# launch list of the codec on 32 x 10 s + condensed `ncu --set full` pages of the tensor-core codec kernels (exported on the box: the
# .ncu-rep itself is too large to bring back)
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_codec.csv python tools/profile_codec.py --batch 32 --chunk 32 > gpurun_out/ncu_codec.log 2>&1; tail -1 gpurun_out/ncu_codec.log
timeout 900 ncu --set full --clock-control none -k regex:"resblock_tc|conv_tc32|conv_tc_kernel" -c 44 -o /tmp/codec_full -f python tools/profile_codec.py --batch 32 --chunk 32 > gpurun_out/ncu_codec_full.log 2>&1; tail -1 gpurun_out/ncu_codec_full.log
ncu -i /tmp/codec_full.ncu-rep --page raw --csv > gpurun_out/codec_full_raw.csv; ls -la gpurun_out/codec_full_raw.csv
