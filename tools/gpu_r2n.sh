#!/bin/bash
# round 2, session 3: full -m gpu suite with the roll-control plugin model; A/B of the on-demand cap (capd) in the stage passes
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r02n_pytest_gpu_full.txt; tail -8 gpurun_out/r02n_pytest_gpu_full.txt | tee gpurun_out/r02n_pytest_gpu.txt
for v in main nodcap capsm main nodcap capsm; do
  lib=$PWD/scpp_b200/libscpp_b200_$v.so; [ $v = main ] && lib=$PWD/scpp_b200/libscpp_b200.so
  echo "== bench 1024 $v"; SCPP_B200_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02n_bench_1024_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
