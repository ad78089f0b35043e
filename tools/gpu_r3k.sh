#!/bin/bash
# round 2, third session, final state (re-run after the last K1 change): full -m gpu suite, smoke, default bench (with the CPU arm), reference arm
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r03k_pytest_gpu_full.txt; tail -8 gpurun_out/r03k_pytest_gpu_full.txt | tee gpurun_out/r03k_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r03k_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r03k_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_cold']), d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d['roofline_fp64']['frac'], d['cpu_baseline']['value'], d['failed_fraction'])"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench.err | tee gpurun_out/r03k_bench_ref.json | cut -c1-300
