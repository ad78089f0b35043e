#!/bin/bash
# source-counter profile of one warm k_solve launch (second outer iteration), batch 1024
mkdir -p gpurun_out
timeout 1200 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -f -o gpurun_out/prof_v3 python bench.py --steps 1 --warmup 0 --batch 1024 --no-cpu-baseline > gpurun_out/ncu_v3.log 2>&1
tail -2 gpurun_out/ncu_v3.log | cut -c1-300
