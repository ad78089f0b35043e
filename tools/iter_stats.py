#!/usr/bin/env python
"""per-outer-iteration statistics of the interior-point iteration counts over a batch (GPU): mean / max / histogram"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scpp_b200 as S
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50)
cfg.ipm.warm = float(os.environ.get("SCPP_WARM", "0.995")); cfg.ipm.stalled_step = int(os.environ.get("SCPP_STALLED_STEP", "0"))
rpy = np.deg2rad([-20.0, 20.0, 0.0])
xi = S.perturbed_initial_states(x_init, rpy, batch)
eng = S.SCAlgorithm(model, params, cfg, batch)
eng.set_boundary_states(xi, x_final)
eng.solve()
info = eng.get_info()
it = info[:, :, 5]; st = info[:, :, 6]
tot_mean = 0; tot_max = 0
for o in range(info.shape[1]):
    v = it[:, o]
    tot_mean += v.mean(); tot_max += v.max()
    print(f"outer {o:2d}: ipm its mean {v.mean():5.1f} p50 {np.percentile(v,50):4.0f} p90 {np.percentile(v,90):4.0f} p99 {np.percentile(v,99):4.0f} max {v.max():4.0f}  status!=0: {(st[:,o]!=0).sum()}")
print(f"sum of means {tot_mean:.1f}  sum of maxima {tot_max:.1f}  (kernel time follows the maxima: ratio {tot_max/tot_mean:.2f})")
print(eng.last_timing())
tot = it.sum(axis=1)
print("per-instance total ipm iterations: mean %.1f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
order = np.argsort(-tot)[:6]
for i in order:
    print(f"instance {i}: total {tot[i]:.0f} per outer {it[i].astype(int).tolist()} n1 {info[i,-1,0]:.2e} sd {info[i,-1,1]:.2e} wtr {info[i,-1,4]:.0f}")
i = np.argsort(tot)[len(tot)//2]
print(f"median instance {i}: total {tot[i]:.0f} per outer {it[i].astype(int).tolist()} n1 {info[i,-1,0]:.2e} sd {info[i,-1,1]:.2e} wtr {info[i,-1,4]:.0f}")
rounds = int(tot.max()) + 2
act = [(tot > r).sum() for r in range(0, rounds, 10)]
print("active instances every 10 rounds:", act)
print("work conserving bound: sum/1024 = %.1f rounds of a full batch; actual rounds ~ %d" % (tot.sum() / len(tot), rounds))
