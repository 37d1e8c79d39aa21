#!/bin/bash
# GPU box: launch list of one sampler step + one full ncu capture of the dominant kernel.
# Usage: bash tools/ncu_capture.sh <tag>   (outputs under gpurun_out/)
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 450 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 400 -c 3 \
    -o gpurun_out/prof_gemm_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/ | tail -12
tail -3 gpurun_out/ncu_list_$TAG.log gpurun_out/ncu_full_$TAG.log
