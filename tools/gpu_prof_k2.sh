#!/bin/bash
# one full ncu capture of a steady-state k_solve launch (batch 1024, round 30)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 30 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log | cut -c1-200
