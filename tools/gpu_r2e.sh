#!/bin/bash
# round 2: full -m gpu suite on the final engine, bench (default command), ncu captures for the FP64 / DRAM accounting, launch list
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -12 | tee gpurun_out/r02e_pytest_gpu.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02e_bench_1024.json | cut -c1-300; tail -2 gpurun_out/bench.err
echo "== ncu k_solve (solver 0), launch 60 of a 1024-instance solve"
SCPP_SOLVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k k_solve -s 60 -c 1 -o gpurun_out/prof_k2s0 -f python tools/prof_cta.py 1024 15 > gpurun_out/ncu_k2s0.log 2>&1; tail -1 gpurun_out/ncu_k2s0.log
echo "== ncu k_discretize, first launch (1024 instances)"
SCPP_SOLVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k k_discretize -s 0 -c 1 -o gpurun_out/prof_k1 -f python tools/prof_cta.py 1024 2 > gpurun_out/ncu_k1.log 2>&1; tail -1 gpurun_out/ncu_k1.log
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launch_bench.log 2>&1; tail -1 gpurun_out/launch_bench.log | cut -c1-100
