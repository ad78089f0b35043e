#!/bin/bash
mkdir -p gpurun_out
for c in 1 2; do
echo "== bench solver 1 cta_per_sm $c"; SCPP_CTA_PER_SM=$c SCPP_SOLVER=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02d_bench_c${c}.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['gpu_launches'])"
done
SCPP_CTA_PER_SM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_cta -s 2 -c 1 -o gpurun_out/prof_cta1 -f python tools/prof_cta.py 148 3 > gpurun_out/ncu_cta.log 2>&1
tail -2 gpurun_out/ncu_cta.log
