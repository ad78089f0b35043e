#!/bin/bash
# perf round: bench at the metric's batch (1024) and in the throughput regime, then a source-counter profile
mkdir -p gpurun_out
echo "== bench 1024"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024.json | cut -c1-400
echo "== bench 4096"; timeout 900 python bench.py --steps 2 --warmup 3 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096.json | cut -c1-400
if [ -n "$WARM" ]; then echo "== bench 1024 cold"; SCPP_WARM=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_1024_cold.json | cut -c1-400; fi
echo "== ncu source counters"
timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -o gpurun_out/prof_cur python bench.py --steps 1 --warmup 0 --batch 1480 --no-cpu-baseline > gpurun_out/ncu_cur.log 2>&1
tail -2 gpurun_out/ncu_cur.log | cut -c1-300
