#!/bin/bash
# round 2, final kernel (4 GPUs): weak scaling 1 / 2 / 4 x 1024, rounds per step at each N
mkdir -p gpurun_out
for n in 1 2 4; do
  echo "== bench SC $n GPUs"
  if [ $n = 1 ]; then timeout 600 python bench.py --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | grep '^{' > gpurun_out/r02u_bench_sc_${n}gpu.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --no-extras --no-cpu-baseline 2>gpurun_out/bench$n.err | grep '^{' > gpurun_out/r02u_bench_sc_${n}gpu.json; fi
  python -c "
import json,sys; d=json.loads(open('gpurun_out/r02u_bench_sc_${n}gpu.json').read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['roofline']['launches_per_step'])"
done
