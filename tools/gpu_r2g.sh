#!/bin/bash
# round 2 (2 GPUs): new parity tests (dual-number Jacobians, per-instance parameters, 2-GPU bit-identity), bench at 1 and 2 GPUs with both Jacobian paths
mkdir -p gpurun_out
echo "== pytest new"; timeout 1500 python -m pytest tests -x -q -m gpu -s -k "dual_number or per_instance or two_gpu or scvx_k50" 2>&1 | tail -12 | tee gpurun_out/r02g_pytest_new.txt
echo "== bench 1 GPU (jacobian = 1 default)"; timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02g_bench_1gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'])"
echo "== bench 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench2.err | tee gpurun_out/r02g_bench_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms'], d['n_gpus'])"
tail -2 gpurun_out/bench2.err
