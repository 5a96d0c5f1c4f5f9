#!/bin/bash
# Round deliverables on one B200: bench line, launch list of the same command, one full ncu capture of the top kernels.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
