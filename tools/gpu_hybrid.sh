#!/bin/bash
# hybrid rounds: split pipeline below T active instances
mkdir -p gpurun_out
for sl in -256 -512 -768 -1000; do
echo "== bench 1024 slice $sl"; SCPP_SLICE=$sl timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_1024_s$sl.json | cut -c1-110; python -c "
import json; d=json.load(open('gpurun_out/bench_1024_s$sl.json')); print(d['kernel_ms'], d['gpu_launches'])"
done
tail -3 gpurun_out/bench.err
