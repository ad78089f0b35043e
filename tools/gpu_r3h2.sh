#!/bin/bash
# round 2, third session, final state on 2 GPUs: the sharded == single-GPU test and the 2-GPU bench line
mkdir -p gpurun_out
echo "== pytest 2 gpu"; timeout 1200 python -m pytest tests -q -m gpu -k "two_gpu" 2>&1 | tail -5 | tee gpurun_out/r03h_pytest_2gpu.txt
echo "== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/bench2.err | tail -1 | tee gpurun_out/r03h_bench_sc_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['n_gpus'], d['kernel_ms'])"
