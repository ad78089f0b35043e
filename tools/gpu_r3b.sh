#!/bin/bash
# round 2, third session: K1 with the shared linearisation (cfg.jacobian = 2) -- parity tests, A/B bench against the column kernel, launch times, counters
mkdir -p gpurun_out
echo "== pytest K1"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "discretize or shared_linearisation or dual_number" 2>&1 | grep -v "^E    *+\|^E    *where" | tail -30 > gpurun_out/r03b_pytest_k1.txt; tail -3 gpurun_out/r03b_pytest_k1.txt
for j in 2 1; do
echo "== bench jacobian=$j"; SCPP_JACOBIAN=$j timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r03b_bench_1024_j$j.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['roofline']['launches_per_step'], d['failed_fraction'])"
done
echo "== ncu launch list"; SCPP_JACOBIAN=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_discretize -c 20 --csv --log-file gpurun_out/r03b_k1_launches.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r03b_ncu.log 2>&1; grep k_discretize gpurun_out/r03b_k1_launches.csv | cut -d, -f5,9,15 | head -4
echo "== ncu full, first K1 launch"; SCPP_JACOBIAN=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_discretize -c 1 -f -o gpurun_out/r03b_k1 python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r03b_ncu2.log 2>&1
python tools/ncu_counts.py gpurun_out/r03b_k1.ncu-rep 1024 | tee gpurun_out/r03b_ncu_counts_k1.json
ncu -i gpurun_out/r03b_k1.ncu-rep --page details 2>/dev/null | grep -i "Duration\|Executed Ipc Active\|Issue Slots Busy\|Registers Per\|Theoretical Occ\|Achieved Occ\|Dynamic Shared\|L2 Hit\|Stall\|FP64\|Bank conf" | head -30 | tee gpurun_out/r03b_k1_details.txt
