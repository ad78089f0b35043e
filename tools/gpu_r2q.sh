#!/bin/bash
# round 2, session 3: retry of a failed factorisation through a kernel exit (no loop / back-edge around the factorisation) vs the round-1 driver; SCvx K = 50
# test (the path that needs the retry); fresh ncu capture of k_solve
mkdir -p gpurun_out
echo "== pytest (SCvx K=50 full batch, bench-shape sample, slicing)"; timeout 1500 python -m pytest tests -q -m gpu -k "scvx or bench_shape or slicing or split_pipeline" 2>&1 | tail -3 | tee gpurun_out/r02q_pytest.txt
for v in main r01solve main r01solve; do
  lib=$PWD/scpp_b200/libscpp_b200_$v.so; [ $v = main ] && lib=$PWD/scpp_b200/libscpp_b200.so
  echo "== bench 1024 $v"; SCPP_B200_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02q_bench_1024_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
echo "== ncu k_solve (solver 0), launch 60 of a 1024-instance solve"
SCPP_SOLVER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 60 -c 1 -o gpurun_out/prof_k2_r02q -f python tools/prof_cta.py 1024 15 > gpurun_out/ncu_k2_r02q.log 2>&1; tail -1 gpurun_out/ncu_k2_r02q.log | cut -c1-200
