#!/bin/bash
# round 2, third session: last check of the committed state -- full -m gpu suite, smoke, default bench line
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r03m_pytest_gpu_full.txt; tail -4 gpurun_out/r03m_pytest_gpu_full.txt | tee gpurun_out/r03m_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r03m_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r03m_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_cold']), d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['failed_fraction'])"
