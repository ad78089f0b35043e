#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scpp_b200 as S, orc_py as O
p, rpy = O.falcon9()
plist = [p, O.rq_perturb(p, rpy, 0x5C99, 3)]
xi = np.array([list(q.x_init) for q in plist])
ro = [O.sc_solve(O.ROCKETQUAT, q, O.sc_config(K=50, max_iterations=4)) for q in plist]
out = {}
for jac in (0, 1, 0):
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=4, keep_history=1, jacobian=jac)
    print("cfg.jacobian", cfg.jacobian, "solver", cfg.solver, "nsub", cfg.nsub)
    eng = S.SCAlgorithm(model, params, cfg, 2)
    eng.set_boundary_states(xi, x_final); eng.solve()
    Xh, Uh, th = eng.get_all_solutions(); t = eng.last_timing(); eng.close()
    for i in range(2):
        print(" jac", jac, "inst", i, ["%.1e" % np.abs(Xh[i, it] - ro[i]["X_all"][it]).max() for it in range(5)], t["ms_discretize"])
