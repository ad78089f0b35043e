"""Generates tests/golden/*.npz from the CPU ORACLE (oracle/*.c).

These are NOT reference outputs: EmbersArc/SCpp ships no golden vectors and cannot be built in this environment
(PARITY UNPINNED).  The fixtures freeze the oracle's own certified results so that (a) the oracle
cannot drift silently and (b) the GPU tests have committed vectors to compare against on the box.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc_py as O  # noqa: E402


def run(model, p, K, max_it):
    cfg = O.sc_config(K=K, model=model, max_iterations=max_it)
    r = O.sc_solve(model, p, cfg)
    info = np.array([[i.norm1_nu, i.sum_delta, i.delta_sigma, i.sigma, i.weight_tr_used, i.ipm.iterations, i.ipm.status, i.ipm.pres, i.ipm.dres, i.ipm.relgap]
                     for i in r["info"]])
    return dict(X_all=r["X_all"], U_all=r["U_all"], t_all=r["t_all"], X=r["X"], U=r["U"], t=np.array(r["t"]),
                iterations=np.array(r["iterations"]), converged=np.array(int(r["converged"])), info=info)


def run_scvx(model, p, K):
    cfg = O.scvx_config(K=K, model=model)
    r = O.scvx_solve(model, p, cfg)
    info = np.array([[i.norm1_nu, i.nonlinear_cost, i.rho, i.trust_region_used, i.solves, i.ipm.iterations, i.ipm.status, i.ipm.pres, i.ipm.dres, i.ipm.relgap]
                     for i in r["info"]])
    return dict(X_all=r["X_all"], U_all=r["U_all"], X=r["X"], U=r["U"], t=np.array(r["t"]), iterations=np.array(r["iterations"]),
                converged=np.array(int(r["converged"])), info=info)


if __name__ == "__main__":
    O.build()
    np.savez_compressed(os.path.join(HERE, "rocket2d_K30.npz"), **run(O.ROCKET2D, O.rocket2d(), 30, 15))
    p, rpy = O.falcon9()
    np.savez_compressed(os.path.join(HERE, "rocketquat_K20_nominal.npz"), **run(O.ROCKETQUAT, p, 20, 5))
    np.savez_compressed(os.path.join(HERE, "rocketquat_K50_inst7.npz"), **run(O.ROCKETQUAT, O.rq_perturb(p, rpy, 0x5C99, 7), 50, 4))
    np.savez_compressed(os.path.join(HERE, "rocketquat_scvx_K30_nominal.npz"), **run_scvx(O.ROCKETQUAT, p, 30))
    print("written", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
