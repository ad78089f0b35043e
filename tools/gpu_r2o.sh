#!/bin/bash
# round 2, session 3: fixed-final-time SC on the device; A/B of the round-1 driver of the interior-point loop (no retry loop / cap state) against main
mkdir -p gpurun_out
echo "== pytest (fixed final time, error paths)"; timeout 1200 python -m pytest tests -q -m gpu -k "fixed_final or warm_start_and_errors" 2>&1 | tail -3 | tee gpurun_out/r02p_pytest_fixed.txt
for v in main r01solve main r01solve; do
  lib=$PWD/scpp_b200/libscpp_b200_$v.so; [ $v = main ] && lib=$PWD/scpp_b200/libscpp_b200.so
  echo "== bench 1024 $v"; SCPP_B200_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02p_bench_1024_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
