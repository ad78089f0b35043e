#!/usr/bin/env python
"""Monte-Carlo closed loop on the GPU (scpp/src/SC_sim.cpp for a batch): perturbed initial states, solve -> K4 step -> warm solve ...
Prints per-step timing and how many instances are still flying.  usage: closed_loop_mc.py [batch] [steps] [K]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scpp_b200 as S
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=K)
cfg.ipm.warm = float(os.environ.get("SCPP_WARM", "0.995"))
xi = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), batch)
eng = S.SCAlgorithm(model, params, cfg, batch)
eng.set_boundary_states(xi, x_final)
rows = []
t_all = time.perf_counter()
for s in range(steps):
    t0 = time.perf_counter()
    eng.solve(warm_start=s > 0)
    tm = eng.last_timing()
    r = eng.sim_step(0.05)
    dt = time.perf_counter() - t0
    sol_flags = eng.get_solution()["flags"]
    rows.append(dict(step=s, wall_ms=1e3 * dt, solve_ms=tm["ms_total"], instance_iterations=tm["instance_iterations"], flying=int((r["reached"] == 0).sum()),
                     failed=int((sol_flags == 2).sum()), altitude_mean=float(r["x"][:, 3].mean())))
    print(json.dumps(rows[-1]), flush=True)
wall = time.perf_counter() - t_all
print(json.dumps(dict(batch=batch, K=K, steps=steps, wall_s=wall, closed_loop_steps_per_s=batch * steps / wall,
                      instance_iterations_per_s=sum(r["instance_iterations"] for r in rows) / wall)))
eng.close()
