#!/bin/bash
# quick GPU check: parity tests + device-resident bench at 1024 and 4096 instances
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== bench 1024"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024.json | cut -c1-120; python -c "
import json; d=json.load(open('gpurun_out/bench_1024.json')); print(d['kernel_ms'], d['gpu_launches'])"
echo "== bench 4096"; timeout 900 python bench.py --steps 2 --warmup 3 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096.json | cut -c1-120
tail -3 gpurun_out/bench.err
