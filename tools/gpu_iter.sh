#!/bin/bash
# developer loop on a GPU box: correctness first, then per-stage profile and timelines.  usage: tools/gpu_iter.sh TAG [full]
TAG=$1
OUT=gpurun_out/r02/$TAG
mkdir -p $OUT
timeout 300 python tools/gpu_check.py fp16 > $OUT/check_fp16.txt 2>&1
timeout 300 python tools/gpu_check.py bf16 > $OUT/check_bf16.txt 2>&1
grep -E "forward vs|rk4_50|cfg3|Error|error" $OUT/check_fp16.txt $OUT/check_bf16.txt | head -30
for B in 256 1024; do timeout 120 python tools/gpu_profile.py bf16 $B > $OUT/prof_B$B.txt 2>&1; done
for B in 148 1024; do timeout 120 python tools/gpu_timeline.py bf16 $B > $OUT/tl_B$B.txt 2>&1; done
for B in 256 1024; do timeout 120 python tools/gpu_check.py time bf16 $B > $OUT/time_$B.txt 2>&1; grep "samples/s" $OUT/time_$B.txt | tail -1; done
if [ "$2" == "full" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.txt 2>&1; tail -5 $OUT/pytest.txt; fi
