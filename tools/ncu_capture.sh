#!/bin/bash
# GPU box: launch list of one sampler step, full ncu captures of the two dominant kernels, and the DRAM
# traffic of every gemm_kernel launch of one decoder forward.  Usage: bash tools/ncu_capture.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick --no-strong"
QP="python tools/quick_perf.py --batch 64 --reps 0 --no-graph"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 450 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$QP > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attn_weights" -s 1 -c 4 \
    -o gpurun_out/prof_top_$TAG -f $QP > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_kernel -s 0 -c ${GEMMS:-322} --csv --log-file gpurun_out/gemm_traffic_$TAG.csv $QP > gpurun_out/ncu_traffic_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
tail -n 2 gpurun_out/ncu_list_$TAG.log gpurun_out/ncu_full_$TAG.log gpurun_out/ncu_traffic_$TAG.log
