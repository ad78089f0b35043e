#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in 8 7; do
echo "== bench 4096 persistent wpb $w"; SCPP_LARGE_WPB=$w timeout 900 python bench.py --steps 2 --warmup 2 --batch 4096 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_4096_w$w.json | cut -c1-100; python -c "
import json; d=json.load(open('gpurun_out/bench_4096_w$w.json')); print(d['kernel_ms'])"
done
echo "== bench 16384"; timeout 900 python bench.py --steps 1 --warmup 1 --batch 16384 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_16384.json | cut -c1-100
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_1024.json | cut -c1-100
tail -2 gpurun_out/bench.err
