#!/bin/bash
# round 2, session 3: full -m gpu suite with the plugin-surface model (model id 2), default bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02k_pytest_gpu_full.txt; tail -8 gpurun_out/r02k_pytest_gpu_full.txt | tee gpurun_out/r02k_pytest_gpu.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02k_bench_1024.json | cut -c1-300; tail -2 gpurun_out/bench.err
