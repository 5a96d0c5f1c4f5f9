#!/bin/bash
mkdir -p gpurun_out
for mode in "bf16" "fp16" "time bf16 256" "time fp16 256" "time bf16 1024"; do
  tag=$(echo $mode | tr ' ' '_')
  timeout 300 python tools/gpu_check.py $mode > gpurun_out/fused_$tag.log 2>&1
  echo "== fused $mode exit $?" | tee -a gpurun_out/fused_summary.txt
  tail -n 25 gpurun_out/fused_$tag.log
done
