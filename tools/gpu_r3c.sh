#!/bin/bash
# round 2, third session: quick A/B of a K1 build -- shared-linearisation parity test, bench at 1024, one full ncu capture of the first K1 launch
TAG=${1:-r03c}
mkdir -p gpurun_out
echo "== pytest K1"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shared_linearisation" 2>&1 | grep -v "^E    *+\|^E    *where" | tail -30 > gpurun_out/${TAG}_pytest_k1.txt; tail -3 gpurun_out/${TAG}_pytest_k1.txt
echo "== bench"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/${TAG}_bench_1024.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['roofline']['launches_per_step'], d['failed_fraction'])"
echo "== ncu full, first K1 launch"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_discretize -c 1 -f -o gpurun_out/${TAG}_k1 python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
python tools/ncu_counts.py gpurun_out/${TAG}_k1.ncu-rep 1024 > gpurun_out/${TAG}_ncu_counts_k1.json; grep "duration\|warp_instr\|flop_per_unit" gpurun_out/${TAG}_ncu_counts_k1.json
