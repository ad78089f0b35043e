"""compute-sanitizer payload: a few K1 launches on the shared-linearisation path (ragged batch, both integrator schedules, zero-order hold through
a short solve); run as  compute-sanitizer --tool racecheck|memcheck python tools/k1_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scpp_b200 as S
rng = np.random.default_rng(0)
K = 12
X = rng.normal(size=(5, K, 14)) * 0.1; X[:, :, 0] = 1.0; X[:, :, 7] = 1.0
U = rng.normal(size=(5, K, 4)) * 0.01; U[:, :, 2] = 0.02
par = np.array([0.3, 0.0, 0.0, -0.01, 0.4, 0.4, 0.01, 0.0, 0.0, -0.02])
for nsub in (-5, 3):
    a = S.discretize(S.ROCKETQUAT, X, U, 10.0, par, nsub=nsub, jacobian=2); b = S.discretize(S.ROCKETQUAT, X, U, 10.0, par, nsub=nsub, jacobian=0)
    print(nsub, max(np.abs(a[k] - b[k]).max() for k in a))
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=10, max_iterations=2, jacobian=2, interpolate_input=0)
eng = S.SCAlgorithm(model, params, cfg, 3)
eng.set_boundary_states(np.tile(x_init, (3, 1)), x_final)
eng.solve()
print("solve", eng.get_solution()["iterations"])
eng.close()
