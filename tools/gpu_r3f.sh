#!/bin/bash
# round 2, third session: state with the shared-linearisation K1 as the bench path -- full -m gpu suite, smoke, default bench, 4096, Starship K = 100, SCvx K = 50
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E    *+\|^E    *where" > gpurun_out/r03f_pytest_gpu_full.txt; tail -8 gpurun_out/r03f_pytest_gpu_full.txt | tee gpurun_out/r03f_pytest_gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r03f_smoke.txt
echo "== bench default"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r03f_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['value_cold']), d['ms_per_step'], d['kernel_ms'], d['interior_point_iterations_per_instance_iteration'], d['roofline']['launches_per_step'], d['failed_fraction'], d['roofline_fp64']['frac'])"
echo "== bench 4096"; timeout 900 python bench.py --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r03f_bench_4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['failed_fraction'])"
echo "== starship K=100 4096"; timeout 900 python bench.py --config RocketQuatStarship --K 100 --batch 4096 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r03f_starship_K100_N4096.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['failed_fraction'])"
echo "== SCvx K=50 1024"; timeout 900 python bench.py --algorithm SCvx --steps 1 --warmup 1 --no-extras 2>gpurun_out/bench.err | tee gpurun_out/r03f_scvx_K50_N1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['failed_fraction'], d.get('converged_fraction'))"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03f_launches.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > /dev/null 2>&1; python - <<'P'
import csv,collections
t=collections.Counter(); n=collections.Counter()
for r in csv.reader(open('gpurun_out/r03f_launches.csv')):
    if len(r)>14 and r[0].isdigit():
        k=r[4].split('<')[0].replace('void ',''); t[k]+=float(r[-1]); n[k]+=1
tot=sum(t.values())
for k,v in t.most_common(): print(k, n[k], round(v/1e6,2),'ms', round(100*v/tot,1),'%')
P
