#!/bin/bash
# round 2, session 3: full -m gpu suite at HEAD (incl. MPC K6, dual-number Jacobians, per-instance parameters), smoke, default bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02h_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r02h_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02h_bench_1024.json | cut -c1-400; tail -2 gpurun_out/bench.err
