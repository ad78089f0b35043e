#!/bin/bash
# round 2, first GPU pass: the whole -m gpu suite (with the new parity tests, -s to keep their printed summaries) + the round-1 bench line
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -x -q -m gpu -s 2>&1 | tail -40 | tee gpurun_out/r02a_pytest_gpu.txt
echo "== bench 1024"; timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/r02a_bench_1024.json | cut -c1-200
echo "== bench scvx"; timeout 600 python bench.py --steps 2 --warmup 2 --algorithm SCvx 2>>gpurun_out/bench.err | tee gpurun_out/r02a_bench_scvx_1024.json | cut -c1-200
tail -3 gpurun_out/bench.err
