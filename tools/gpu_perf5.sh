#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== bench 1024"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024.json | cut -c1-200
echo "== bench 4096"; timeout 900 python bench.py --steps 2 --warmup 3 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096.json | cut -c1-200
echo "== iteration statistics"; timeout 600 python tools/iter_stats.py 1024 2>&1 | tail -14 | tee gpurun_out/iter_stats.txt
