#!/bin/bash
# first-contact script for a GPU box: every stage in its own process, everything logged
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for mode in "fp32" "selftest" "bf16" "time fp32 256" "time bf16 256" "time bf16 1024"; do
  tag=$(echo $mode | tr ' ' '_')
  timeout 300 python tools/gpu_check.py $mode > gpurun_out/check_$tag.log 2>&1
  echo "== $mode exit $?" | tee -a gpurun_out/summary.txt
  tail -n 60 gpurun_out/check_$tag.log
done
