#!/bin/bash
# round 2, third session: K1 with two 8-warp CTAs per SM (the default CTA shape from here on) -- K1 tests, bench, counters of the first K1 launch
mkdir -p gpurun_out
echo "== pytest K1"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shared_linearisation or zero_order or discretize or dual_number" 2>&1 | grep -v "^E    *+\|^E    *where" | tail -30 > gpurun_out/r03j_pytest_k1.txt; tail -3 gpurun_out/r03j_pytest_k1.txt
echo "== bench"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r03j_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['roofline']['launches_per_step'], d['failed_fraction'])"
echo "== ncu full, first K1 launch"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_discretize -c 1 -f -o gpurun_out/r03j_k1 python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r03j_ncu2.log 2>&1
python tools/ncu_counts.py gpurun_out/r03j_k1.ncu-rep 1024 > gpurun_out/r03j_ncu_counts_k1.json; grep "duration\|warp_instr\|flop_per_unit" gpurun_out/r03j_ncu_counts_k1.json
ncu -i gpurun_out/r03j_k1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for k,x in zip(h,v):
    if ('pcsamp_warps_issue_stalled' in k and 'not_issued' not in k) or k in ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','dram__bytes_read.sum','dram__bytes_write.sum'): print(k,x)
" | sort -t' ' -k2 -n -r | head -16 | tee gpurun_out/r03j_k1_stalls.txt
