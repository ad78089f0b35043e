#!/bin/bash
# round 2, final kernel: N sweep at K = 50, Starship K = 100 x 4096, SCvx K = 50 x 1024, CTA solver, MPC Monte-Carlo
mkdir -p gpurun_out
for n in 16384 65536; do
  echo "== sweep K=50 N=$n"; timeout 900 python bench.py --batch $n --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02t_sweep_K50_N$n.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
done
echo "== Starship K=100 N=4096"; timeout 900 python bench.py --config RocketQuatStarship --K 100 --batch 4096 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02t_starship_K100_N4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d.get('failed_fraction'))"
echo "== SCvx K=50 N=1024"; timeout 900 python bench.py --algorithm SCvx --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench.err | tee gpurun_out/r02t_scvx_K50_N1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d.get('failed_fraction'), d.get('converged_fraction'))"
echo "== CTA solver N=1024"; timeout 900 python bench.py --solver 1 --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02t_bench_1024_solver1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'])"
