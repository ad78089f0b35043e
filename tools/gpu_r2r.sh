#!/bin/bash
# round 2, session 3: full -m gpu suite, smoke, default bench (with cold / other-solver / CPU arm), reference arm, launch list of the bench command
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02r_pytest_gpu_full.txt; tail -8 gpurun_out/r02r_pytest_gpu_full.txt | tee gpurun_out/r02r_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02r_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02r_bench_1024.json | cut -c1-200; tail -2 gpurun_out/bench.err
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/r02r_bench_ref.json | cut -c1-300
echo "== bench 4096"; timeout 900 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02r_bench_4096.json | cut -c1-120
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02r_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launch_bench.log 2>&1; tail -1 gpurun_out/launch_bench.log | cut -c1-100
echo "== ncu k_solve"; SCPP_SOLVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 60 -c 1 -o gpurun_out/prof_k2_r02r -f python tools/prof_cta.py 1024 15 > gpurun_out/ncu_k2_r02r.log 2>&1; tail -1 gpurun_out/ncu_k2_r02r.log | cut -c1-200
