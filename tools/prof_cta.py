#!/usr/bin/env python
"""small driver for ncu captures of the CTA-per-instance solver: N instances, a few outer iterations"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scpp_b200 as S
N = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mit = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=mit)
cfg.solver = int(os.environ.get("SCPP_SOLVER", "1")); cfg.ipm.warm = float(os.environ.get("SCPP_WARM", "0.995"))
xi = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), N)
eng = S.SCAlgorithm(model, params, cfg, N)
eng.set_boundary_states(xi, x_final)
eng.solve()
print(eng.last_timing(), eng.get_info()[:, :, 5].sum(axis=0))
eng.close()
