#!/bin/bash
# first-contact GPU script: smoke, FP64 peaks, parity tests, a short bench and an ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== fp64 peak"; timeout 120 tools/fp64_peak | tee gpurun_out/fp64_peak.json
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
