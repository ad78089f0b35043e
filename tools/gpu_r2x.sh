#!/bin/bash
# round 2: final state -- full -m gpu suite, smoke, default bench; the stalled-step knob at 3 for the record
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02z_pytest_gpu_full.txt; tail -8 gpurun_out/r02z_pytest_gpu_full.txt | tee gpurun_out/r02z_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02z_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02z_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_cold']), d['ms_per_step'], d['kernel_ms'], d['interior_point_iterations_per_instance_iteration'], d['roofline']['launches_per_step'], d['failed_fraction'])"
echo "== bench 4096"; timeout 900 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02z_bench_4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['interior_point_iterations_per_instance_iteration'], d['roofline']['launches_per_step'], d['failed_fraction'])"
