#!/bin/bash
# evidence for the round: GPU tests, bench lines (with the CPU baseline), reference arm, launch list, full ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_1024.json | cut -c1-200
echo "== bench 1024 cold"; SCPP_WARM=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024_cold.json | cut -c1-200
echo "== bench 4096"; timeout 900 python bench.py --steps 2 --warmup 3 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096.json | cut -c1-200
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-300
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log | cut -c1-120
echo "== ncu full, k_solve launch 30"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 30 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log | cut -c1-200
echo "== ncu full, k_discretize launch 0"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_discretize -s 0 -c 1 -f -o gpurun_out/prof_k1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1; tail -1 gpurun_out/ncu_k1.log | cut -c1-200
tail -3 gpurun_out/bench.err
