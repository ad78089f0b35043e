#!/bin/bash
# round 2: batch-size sweep at K = 50 (solver 0, chunked engine), Starship K = 100 x 4096, SCvx K = 50
mkdir -p gpurun_out
for n in 1024 4096 16384 65536; do
st=2; [ $n -ge 16384 ] && st=1
echo "== sweep N=$n"; timeout 1500 python bench.py --batch $n --steps $st --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/bench.err | tee gpurun_out/r02f_sweep_K50_N$n.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'])"
tail -1 gpurun_out/bench.err
done
echo "== Starship K=100 batch 4096"; timeout 1500 python bench.py --config RocketQuatStarship --K 100 --batch 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>gpurun_out/bench.err | tee gpurun_out/r02f_starship_K100_N4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['converged_fraction'], d['failed_fraction'], d['subproblem_exit_status'])"
tail -1 gpurun_out/bench.err
echo "== SCvx K=50 batch 1024"; timeout 900 python bench.py --algorithm SCvx --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench.err | tee gpurun_out/r02f_scvx_K50_N1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['converged_fraction'], d['failed_fraction'])"
tail -1 gpurun_out/bench.err
