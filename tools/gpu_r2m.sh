#!/bin/bash
# A/B: the final round-1 tree (43.1 k then) on today's box against HEAD (single inlined copy of the factorisation)
mkdir -p gpurun_out
echo "== r01 tree"; (cd _r01 && timeout 600 python bench.py --steps 3 --warmup 2 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])") | tee gpurun_out/r02m_ab_r01.txt
echo "== HEAD"; timeout 600 python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02m_bench_1024_head.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
echo "== r01 tree again"; (cd _r01 && timeout 600 python bench.py --steps 3 --warmup 2 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])")
