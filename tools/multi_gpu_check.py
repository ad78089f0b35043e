#!/usr/bin/env python
"""Multi-GPU parity check (launched by torchrun, one rank per GPU): the converging RocketQuat workload (non-reference weights, see
tests/test_gpu_parity.py::test_converging_rocketquat_workload) sharded UNEVENLY over the ranks -- rank r gets a share proportional to r + 1, so
the shards differ in size and in how early their instances converge -- must give, instance by instance, the bit-identical result of a
single-GPU solve of the whole batch; every rank issues one flag all-gather per outer iteration and learns the global state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import scpp_b200 as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group(backend="gloo")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
algorithm = sys.argv[2] if len(sys.argv) > 2 else "SC"      # "SCvx": BASELINE.json configs[2] (free-final-time SCvx, K = 50, sharded), shipped SCvx.info
if algorithm == "SCvx":
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, algorithm="SCvx")
else:
    over = dict(weight_trust_region_trajectory=2.0, weight_virtual_control=1e4, nu_tol=1e-3, delta_tol=1e-2)
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=20, **over)
cfg.ipm.warm = 0.995
xi_all = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), N)
w = np.arange(1, world + 1, dtype=float); cuts = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * N).astype(int)])
lo, hi = int(cuts[rank]), int(cuts[rank + 1])
eng = S.SCAlgorithm(model, params, cfg, hi - lo, device=local)
from scpp_b200.sharding import broadcast_unique_id
eng.comm_init(world, rank, broadcast_unique_id(dist, S.comm_unique_id, rank))
eng.set_boundary_states(xi_all[lo:hi], x_final)
eng.solve()
sol = eng.get_solution(); ga = eng.global_active(); r = eng.last_rounds()
eng.close()
parts = [None] * world
dist.all_gather_object(parts, (lo, hi, sol["X"], sol["U"], sol["t"], sol["iterations"], sol["flags"], ga, r["rounds"]))
ok = True
if rank == 0:
    ref = S.SCAlgorithm(model, params, cfg, N, device=local)
    ref.set_boundary_states(xi_all, x_final)
    ref.solve()
    rs = ref.get_solution(); rr = ref.last_rounds()
    ref.close()
    for (a, b, X, U, t, it, fl, g, rd) in parts:
        same = np.array_equal(X, rs["X"][a:b]) and np.array_equal(U, rs["U"][a:b]) and np.array_equal(t, rs["t"][a:b]) and np.array_equal(it, rs["iterations"][a:b]) and np.array_equal(fl, rs["flags"][a:b])
        print(f"shard [{a},{b}): bit-identical to the single-GPU solve: {same}; rounds {rd} (single GPU {rr['rounds']}); converged {int((fl == 1).sum())}/{b - a}; global_active reported {g}")
        ok = ok and same and g == int((rs['flags'] == 0).sum() * 0)      # every instance has stopped: nothing is active anywhere
    print("iteration counts over the batch:", np.bincount(rs["iterations"]).tolist())
flag = [ok]
dist.broadcast_object_list(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag[0] else 1)
