#!/bin/bash
# GPU box: full ncu captures (source-level) of the memory-bound kernels of one decoder forward (first launch of
# each: stack-entry prep, depthwise conv K=31, BiasNorm+bypass, down/up-sampling) and of the attention kernel.
# Usage: bash tools/ncu_focus.sh <tag>
TAG=${1:-r1}
QP="python tools/quick_perf.py --batch 64 --reps 0 --no-graph --steps 1"
mkdir -p gpurun_out
$QP > gpurun_out/plain_focus_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"dwconv|biasnorm|sample_|stream_prep|attn_weights" -c 8 \
    -o gpurun_out/prof_mem_$TAG -f $QP > gpurun_out/ncu_mem_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.* | tail; tail -n 2 gpurun_out/ncu_mem_$TAG.log
