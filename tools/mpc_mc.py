#!/usr/bin/env python
"""BASELINE.json configs[3] in spirit (the reference ships no RocketEuler model: Rocket2D is its only model with an operating point and an
MPC.info): receding-horizon MPC for a Monte-Carlo batch of initial states, horizon N = 20 (K = 21), closed loop of scpp/src/MPC_sim.cpp:64-70 on
the device (kernel K6 + K4).  Prints one JSON line: MPC solves per second (device time of the solves), exit-status census, closed-loop drift."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scpp_b200 as S

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 21
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
model, params, x_init, x_final, _ = S.load_model("Rocket2D")
params.constrain_initial_final = 0
cfg = S.load_mpc_info(os.path.join(S.CONFIG_DIR, "Rocket2D", "MPC.info"), S.ROCKET2D)
cfg.K = K; cfg.time_horizon = 0.125 * (K - 1)
rng = np.random.default_rng(0x5C99)
x0 = np.array([-20., 100., 2., -10., 0.05, 0.0]) * (1 + 0.1 * rng.standard_normal((N, 6)))
xf = np.array([0., 0, 0, -1, 0, 0.])
mpc = S.MPCAlgorithm(model, params, cfg, N)
mpc.set_states(x0, xf)
mpc.solve()                                   # warm-up
ms, ok, iters = [], [], []
for _ in range(steps):
    mpc.solve()
    ms.append(mpc.last_ms())
    sol = mpc.get_solution()
    ok.append(float(np.isin(sol["status"], (0, 3)).mean())); iters.append(float(sol["iterations"].mean()))
    x = mpc.sim_step(0.05)
mpc.close()
print(json.dumps({"metric": "mpc_solves_per_sec", "value": N * steps / (sum(ms) * 1e-3), "unit": "solves/s", "batch": N, "K": K, "steps": steps,
                  "ms_per_solve_batch": float(np.mean(ms)), "solved_fraction_per_step": ok, "mean_ipm_iterations": iters,
                  "mean_distance_to_target_after": float(np.linalg.norm(x[:, :2] - xf[:2], axis=1).mean()),
                  "mean_distance_to_target_before": float(np.linalg.norm(x0[:, :2] - xf[:2], axis=1).mean())}))
