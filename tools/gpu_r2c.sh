#!/bin/bash
# CTA solver iteration: parity tests + bench (warm / cold) + a short ncu capture
mkdir -p gpurun_out
echo "== pytest cta"; timeout 1200 python -m pytest tests -x -q -m gpu -k "cta_solver" 2>&1 | tail -5 | tee gpurun_out/r02c_pytest_cta.txt
for w in 0.995 0; do
echo "== bench solver 1 warm $w"; SCPP_SOLVER=1 SCPP_WARM=$w timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02c_bench_s1_w${w}.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['gpu_launches'], d['converged_fraction'], d['failed_fraction'])"
tail -2 gpurun_out/bench.err
done
if [ -n "$PROF" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_cta -s 2 -c 1 -o gpurun_out/prof_cta -f python tools/prof_cta.py 296 3 > gpurun_out/ncu_cta.log 2>&1
tail -2 gpurun_out/ncu_cta.log
fi
