#!/bin/bash
# round 2: hybrid tail (cfg.solver = 2): parity test, bench against solver 0 at batch 1024 and 4096
mkdir -p gpurun_out
echo "== pytest hybrid"; timeout 1500 python -m pytest tests -q -m gpu -k "hybrid_tail or cta_solver" 2>&1 | grep -v "^E    *+\|^E    *where" | tail -12 | tee gpurun_out/r02y_pytest_hybrid.txt
for sv in 0 2 0 2; do
  echo "== bench 1024 solver $sv"; timeout 600 python bench.py --solver $sv --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02y_bench_1024_solver$sv.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['roofline']['launches_per_step'], d['failed_fraction'])"
done
echo "== bench 4096 solver 2"; timeout 600 python bench.py --solver 2 --batch 4096 --steps 2 --warmup 1 --no-extras --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r02y_bench_4096_solver2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['failed_fraction'])"
