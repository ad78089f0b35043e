#!/bin/bash
# round 2, session 3 (2 GPUs): sharded == single-GPU bit-identity for SC (converging workload) and SCvx (BASELINE configs[2]), bench lines at 2 GPUs for
# SC and SCvx, MPC Monte-Carlo batch (configs[3] in spirit)
mkdir -p gpurun_out
echo "== 2-GPU pytest"; timeout 1200 python -m pytest tests -q -m gpu -k two_gpu 2>&1 | tail -3 | tee gpurun_out/r02l_pytest_2gpu.txt
echo "== SCvx sharded vs single GPU"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multi_gpu_check.py 256 SCvx 2>&1 | grep -v "^W\|^\*\*\*\|OMP" | tail -6 | tee gpurun_out/r02l_scvx_2gpu_check.txt
echo "== bench SC 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench2.err | grep '^{' | tee gpurun_out/r02l_bench_sc_2gpu.json | cut -c1-200
echo "== bench SCvx 2 GPUs x 1024"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --algorithm SCvx --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench3.err | grep '^{' | tee gpurun_out/r02l_bench_scvx_2gpu.json | cut -c1-200
echo "== bench SCvx 1 GPU x 1024"; timeout 900 python bench.py --algorithm SCvx --steps 2 --warmup 1 --no-extras 2>gpurun_out/bench4.err | grep '^{' | tee gpurun_out/r02l_bench_scvx_1gpu.json | cut -c1-200
echo "== MPC Monte-Carlo 16384 x horizon 20"; timeout 600 python tools/mpc_mc.py 16384 21 10 2>&1 | tail -1 | tee gpurun_out/r02l_mpc_mc_16384.json | cut -c1-400
echo "== MPC Monte-Carlo 16384 x K 7 (MPC.info)"; timeout 600 python tools/mpc_mc.py 16384 7 10 2>&1 | tail -1 | tee gpurun_out/r02l_mpc_mc_16384_K7.json | cut -c1-400
