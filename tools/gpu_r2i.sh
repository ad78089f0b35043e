#!/bin/bash
# round 2, session 3: full -m gpu suite after the oracle thread-safety fix, fresh ncu capture of K2 (warp solver) with source, 8-warp CTA variant at 4096
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02i_pytest_gpu_full.txt; tail -8 gpurun_out/r02i_pytest_gpu_full.txt | tee gpurun_out/r02i_pytest_gpu.txt
echo "== bench 4096 wpb 7"; timeout 600 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02i_bench_4096_w7.json | cut -c1-120
echo "== bench 4096 wpb 8"; SCPP_B200_LIB=$PWD/scpp_b200/libscpp_b200_w8.so timeout 600 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02i_bench_4096_w8.json | cut -c1-120
echo "== ncu k_solve (solver 0), launch 60 of a 1024-instance solve"
SCPP_SOLVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 60 -c 1 -o gpurun_out/prof_k2_r02i -f python tools/prof_cta.py 1024 15 > gpurun_out/ncu_k2_r02i.log 2>&1; tail -1 gpurun_out/ncu_k2_r02i.log | cut -c1-200
