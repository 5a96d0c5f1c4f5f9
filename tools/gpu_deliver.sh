#!/bin/bash
# Round deliverables on one B200: tests, smoke, bench line, reference arm, stage timeline, launch list of the bench
# command and one full ncu capture of a forward (each ncu pass only after the same command exited 0 without it).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 200 python tools/gpu_timeline.py bf16 256 > gpurun_out/timeline.txt 2>&1
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch-list rc=$?"
timeout 100 python tools/one_forward.py bf16 256 > /dev/null 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_attn" --launch-skip 19 --launch-count 19 \
    -o gpurun_out/r02_full_forward -f python tools/one_forward.py bf16 256 > gpurun_out/ncu_full.log 2>&1
echo "full-capture rc=$?"
timeout 200 python tools/gpu_profile.py bf16 4096 layerwise > gpurun_out/layerwise_b4096.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --batch 1024 --no-cpu > gpurun_out/bench_b1024.json 2> gpurun_out/bench_b1024.err; echo "b1024 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --dtype fp16 --no-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "fp16 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --cfg 3.0 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --config c5 --no-cpu > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; echo "c5 rc=$?"
timeout 200 python tools/gpu_timeline.py bf16 1024 > gpurun_out/timeline_B1024.txt 2>&1
