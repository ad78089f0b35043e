#!/bin/bash
# round 2: interior warm start 50x closer + step fraction 0.9999 after a stalled outer iteration -- full -m gpu suite, default bench, batch 4096, Starship
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02w_pytest_gpu_full.txt; tail -8 gpurun_out/r02w_pytest_gpu_full.txt | tee gpurun_out/r02w_pytest_gpu.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02w_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_cold']), d['ms_per_step'], d['kernel_ms'], d['interior_point_iterations_per_instance_iteration'], d['subproblem_exit_status'], d['failed_fraction'])"
echo "== bench 4096"; timeout 900 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02w_bench_4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['interior_point_iterations_per_instance_iteration'], d['failed_fraction'])"
echo "== Starship K=100 N=4096"; timeout 900 python bench.py --config RocketQuatStarship --K 100 --batch 4096 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02w_starship_K100_N4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['interior_point_iterations_per_instance_iteration'], d.get('failed_fraction'))"
echo "== sweep K=50 N=65536 (failures at scale?)"; timeout 900 python bench.py --batch 65536 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02w_sweep_K50_N65536.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['interior_point_iterations_per_instance_iteration'], d['subproblem_exit_status'], d['failed_fraction'])"
