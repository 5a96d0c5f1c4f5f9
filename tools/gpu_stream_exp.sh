#!/bin/bash
OUT=gpurun_out/r02/$1
mkdir -p $OUT
for NL in 2 4 1; do
  FLO_PROD_LANES=$NL timeout 120 python tools/gpu_profile.py bf16 256 > $OUT/prof_B256_nl$NL.txt 2>&1
  FLO_PROD_LANES=$NL timeout 120 python tools/gpu_profile.py bf16 1024 > $OUT/prof_B1024_nl$NL.txt 2>&1
  FLO_PROD_LANES=$NL timeout 120 python tools/gpu_timeline.py bf16 256 > $OUT/tl_B256_nl$NL.txt 2>&1
  echo "== lanes $NL"; head -3 $OUT/prof_B256_nl$NL.txt; head -1 $OUT/prof_B1024_nl$NL.txt
  grep -A1 "stage 12 S_up1" $OUT/tl_B256_nl$NL.txt | tail -1 | cut -c1-700
done
