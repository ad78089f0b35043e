#!/bin/bash
mkdir -p gpurun_out
python tools/prof_cta.py 296 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_cta -s 2 -c 1 -o gpurun_out/prof_cta -f python tools/prof_cta.py 296 3 > gpurun_out/ncu_cta.log 2>&1
tail -3 gpurun_out/ncu_cta.log
