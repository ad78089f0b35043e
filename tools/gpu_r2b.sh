#!/bin/bash
# round 2: the CTA-per-instance solver: parity tests, then bench with both solvers (warm and cold)
mkdir -p gpurun_out
echo "== pytest cta"; timeout 1200 python -m pytest tests -x -q -m gpu -k "cta_solver" 2>&1 | tail -15 | tee gpurun_out/r02b_pytest_cta.txt
for sv in 1 0; do for w in 0.995 0; do
echo "== bench solver $sv warm $w"; SCPP_SOLVER=$sv SCPP_WARM=$w timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02b_bench_s${sv}_w${w}.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['gpu_launches'], d['converged_fraction'], d['failed_fraction'])"
tail -2 gpurun_out/bench.err
done; done
