#!/usr/bin/env python
"""SCvx on the GPU vs the oracle: per-instance decision sequences, prefix parity, final costs"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scpp_b200 as S, orc_py as O
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", algorithm="SCvx", keep_history=1)
cfg.ipm.warm = float(os.environ.get("SCPP_WARM", "0.995"))
p, rpy = O.falcon9()
plist = [p] + [O.rq_perturb(p, rpy, 0x5C99, i) for i in (0, 2, 3, 5, 7)]
xi = np.array([list(q.x_init) for q in plist])
eng = S.SCAlgorithm(model, params, cfg, len(plist))
eng.set_boundary_states(xi, x_final); eng.solve()
sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions(); print(eng.last_timing()); eng.close()
ocfg = O.scvx_config(K=30, model=O.ROCKETQUAT)
for i, q in enumerate(plist):
    ro = O.scvx_solve(O.ROCKETQUAT, q, ocfg)
    n = int(sol["iterations"][i])
    print(f"instance {i}: oracle its {ro['iterations']} conv {ro['converged']} | gpu its {n} flag {sol['flags'][i]}")
    print("   oracle solves", [a.solves for a in ro["info"]], " J_last %.4e" % (ro["info"][-1].nonlinear_cost if ro["info"] else -1))
    print("   gpu    solves", [int(v) for v in info[i, :n, 4]], " J_last %.4e" % info[i, n - 1, 1])
    for it in range(min(abs(ro["iterations"]), n, 6)):
        a = ro["info"][it]
        print(f"     it {it+1} dX {np.abs(Xh[i, it+1] - ro['X_all'][it+1]).max():.2e} dU {np.abs(Uh[i, it+1] - ro['U_all'][it+1]).max():.2e} rho {a.rho:.4f}/{info[i,it,2]:.4f} TR {a.trust_region_used:.4g}/{info[i,it,3]:.4g} J {a.nonlinear_cost:.5e}/{info[i,it,1]:.5e}")
