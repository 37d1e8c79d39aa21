#!/bin/bash
# Runs every parity check in its own process with a timeout; summary in gpurun_out/checks.jsonl
mkdir -p gpurun_out
: > gpurun_out/checks.jsonl
for c in ${@:-$(python tools/gpu_check.py --list)}; do
  timeout 180 python tools/gpu_check.py "$c" >> gpurun_out/checks.jsonl 2>> gpurun_out/checks.err || echo "{\"check\": \"$c\", \"rc\": $?}" >> gpurun_out/checks.jsonl
done
cat gpurun_out/checks.jsonl
