#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scpp_b200 as S, orc_py as O
p, rpy = O.falcon9()
plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(60, 66)] + [p]
xi = np.array([list(q.x_init) for q in plist])
for rep in range(2):
  for solver in (1, 0):
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=6, keep_history=1)
    cfg.solver = solver
    eng = S.SCAlgorithm(model, params, cfg, len(plist))
    eng.set_boundary_states(xi, x_final); eng.solve()
    sol = eng.get_solution(); info = eng.get_info()
    print("solver", solver, "iters", sol["iterations"], "flags", sol["flags"])
    print("  ipm its", info[:, :, 5].astype(int).tolist())
    print("  status", info[:, :, 6].astype(int).tolist())
    eng.close()
