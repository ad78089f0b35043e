#!/usr/bin/env python
"""split pipeline vs monolithic kernel on the same batch: where do they differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scpp_b200 as S
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=8, keep_history=1)
cfg.ipm.warm = 0.995
xi = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), 200)
out = []
for sl in (1, -1):
    cfg.ipm_slice = sl
    eng = S.SCAlgorithm(model, params, cfg, 200)
    eng.set_boundary_states(xi, x_final); eng.solve()
    out.append((eng.get_all_solutions(), eng.get_info(), eng.get_solution())); eng.close()
(Xa, Ua, ta), (Xb, Ub, tb) = out[0][0], out[1][0]
print("iters equal", np.array_equal(out[0][2]["iterations"], out[1][2]["iterations"]), "flags equal", np.array_equal(out[0][2]["flags"], out[1][2]["flags"]))
dX = np.abs(Xa - Xb).max(axis=(2, 3)); dU = np.abs(Ua - Ub).max(axis=(2, 3))
print("max dX per iterate", dX.max(axis=0)); print("max dU per iterate", dU.max(axis=0))
ia, ib = out[0][1], out[1][1]
print("w_tr equal", np.array_equal(ia[:, :, 4], ib[:, :, 4]), "max ipm-iteration difference", np.abs(ia[:, :, 5] - ib[:, :, 5]).max())
bad = np.argwhere(np.abs(ia[:, :, 5] - ib[:, :, 5]) > 2)
for n, it in bad[:10]: print("instance", n, "outer", it, "its", ia[n, it, 5], ib[n, it, 5], "status", ia[n, it, 6], ib[n, it, 6], "pres", ia[n, it, 7], ib[n, it, 7], "dres", ia[n, it, 8], ib[n, it, 8])
print("status counts mono", np.unique(ia[:, :, 6], return_counts=True), "split", np.unique(ib[:, :, 6], return_counts=True))
