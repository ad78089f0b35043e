#!/bin/bash
# round 2, third session: zero-order-hold inputs on the GPU + the tests around it
mkdir -p gpurun_out
echo "== pytest zoh"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "zero_order or shared_linearisation or rocket2d_config0 or warm_start_closed or k50_batch or reports" 2>&1 | grep -v "^E    *+\|^E    *where" | tail -40 > gpurun_out/r03g_pytest_zoh.txt; tail -5 gpurun_out/r03g_pytest_zoh.txt
