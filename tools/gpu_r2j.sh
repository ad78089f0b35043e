#!/bin/bash
# round 2, session 3: code-size variants of k_solve (out-of-line Cholesky / stage passes) against the main build, batch 1024
mkdir -p gpurun_out
echo "== pytest (two adjusted tests)"; timeout 900 python -m pytest tests -q -m gpu -k "converging_rocketquat or mpc_vs" 2>&1 | tail -3
for v in main nichol nipass niboth; do
  lib=$PWD/scpp_b200/libscpp_b200_$v.so; [ $v = main ] && lib=$PWD/scpp_b200/libscpp_b200.so
  echo "== bench 1024 $v"; SCPP_B200_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02j_bench_1024_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
