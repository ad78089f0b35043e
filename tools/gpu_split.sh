#!/bin/bash
# split pipeline: parity tests, then bench at 1024 / 4096 in both modes
mkdir -p gpurun_out
echo "== pytest split"; timeout 900 compute-sanitizer --version >/dev/null 2>&1; timeout 1200 python -m pytest tests -x -q -m gpu -k "split or tensor_core" 2>&1 | tail -15 | tee gpurun_out/pytest_split.txt
for sl in -1 1; do
echo "== bench 1024 slice $sl"; SCPP_SLICE=$sl timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024_s$sl.json | cut -c1-120; python -c "
import json; d=json.load(open('gpurun_out/bench_1024_s$sl.json')); print(d['kernel_ms'], d['gpu_launches'])"
done
echo "== bench 4096 split"; SCPP_SLICE=-1 timeout 900 python bench.py --steps 2 --warmup 3 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096_s-1.json | cut -c1-120
tail -3 gpurun_out/bench.err
