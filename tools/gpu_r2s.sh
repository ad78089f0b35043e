#!/bin/bash
# A/B: global loads of k_solve bypassing L1 (-Xptxas -dlcm=cg: L1 left to the local-memory frame) against main
mkdir -p gpurun_out
for v in main cg main cg; do
  lib=$PWD/scpp_b200/libscpp_b200_$v.so; [ $v = main ] && lib=$PWD/scpp_b200/libscpp_b200.so
  echo "== bench 1024 $v"; SCPP_B200_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02s_bench_1024_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
