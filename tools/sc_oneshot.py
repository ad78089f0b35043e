#!/usr/bin/env python
"""SC_oneshot / SC_sim on the B200 engine with the reference's OUTPUT LAYOUT, so that its evaluation scripts keep working:

  scpp/src/SC_oneshot.cpp:29-63   output/<Model>/SC/<time>/<iteration>/{X.txt, U.txt, t.txt}     (every iterate, redimensionalised)
  scpp/src/SC_sim.cpp:68-107      output/<Model>/SC_sim/<time>/0/{X.txt, U.txt, t.txt}           (closed-loop states / applied inputs)

Rows are nodes, columns are comma-separated in Eigen's default stream precision (6 significant digits), which is what
evaluation/RocketQuat/plot_RocketQuat.py:31-32 and evaluation/Rocket2D/plot_rocket2d.py read with np.loadtxt(delimiter=",").

usage: sc_oneshot.py [--model RocketQuat|Rocket2D|RocketQuatStarship] [--K n] [--sim steps] [--time-step 0.05] [--out DIR] [--config DIR]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scpp_b200 as S


def eigen_csv(path, M):
    """Eigen::IOFormat(StreamPrecision, DontAlignCols, ", ", "\\n") of row vectors, one per line"""
    with open(path, "w") as f:
        for row in np.atleast_2d(M):
            f.write(", ".join("%.6g" % v for v in row) + "\n")


def redimensionalize(model_name, params, x_init, X, U):
    """Model::redimensionalizeTrajectory (rocketQuat.cpp:188-201, rocket2d.cpp:109-119) with the scales of nondimensionalize()"""
    X, U = X.copy(), U.copy()
    if model_name.startswith("RocketQuat"):
        ms, rs = x_init[0], np.linalg.norm(x_init[1:4])
        X[..., 0] *= ms; X[..., 1:7] *= rs; U[..., 0:3] *= ms * rs; U[..., 3] *= ms * rs * rs
    else:
        ms, rs = params.m, np.linalg.norm(x_init[0:2])       # rocket2d.cpp:198-214: m_scale = m, r_scale = |r_init|
        X[..., 0:4] *= rs; U[..., 1] *= ms * rs
    return X, U


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="RocketQuat")
    ap.add_argument("--K", type=int, default=None)
    ap.add_argument("--sim", type=int, default=0, help="closed-loop steps (SC_sim); 0 = SC_oneshot")
    ap.add_argument("--time-step", type=float, default=0.05)
    ap.add_argument("--out", default=os.path.join("..", "output"))
    args = ap.parse_args()
    kw = {"keep_history": 1}
    if args.K:
        kw["K"] = args.K
    model, params, x_init, x_final, cfg = S.load_model(args.model, **kw)
    eng = S.SCAlgorithm(model, params, cfg, 1)
    eng.set_boundary_states(x_init, x_final)
    stamp = time.strftime("%Y-%m-%d_%H-%M-%S")
    name = "RocketQuat" if args.model.startswith("RocketQuat") else args.model
    if args.sim <= 0:
        eng.solve()
        n = int(eng.get_solution()["iterations"][0])
        Xh, Uh, th = eng.get_all_solutions()
        root = os.path.join(args.out, name, "SC", stamp)
        for it in range(n + 1):
            X, U = (Xh[0, it], Uh[0, it])
            if cfg.nondimensionalize:
                X, U = redimensionalize(args.model, params, np.asarray(x_init, float), X, U)
            d = os.path.join(root, str(it)); os.makedirs(d, exist_ok=True)
            eigen_csv(os.path.join(d, "X.txt"), X); eigen_csv(os.path.join(d, "U.txt"), U)
            with open(os.path.join(d, "t.txt"), "w") as f:
                f.write("%.6g" % th[0, it])
        print(root)
    else:
        Xs, Us, step = [], [], 0
        while step < args.sim:                                  # SC_sim.cpp:41-65
            eng.solve(warm_start=step > 0)
            r = eng.sim_step(args.time_step)
            Xs.append(r["x"][0]); Us.append(r["u0"][0])
            if r["reached"][0]:
                break
            step += 1
        d = os.path.join(args.out, name, "SC_sim", stamp, "0"); os.makedirs(d, exist_ok=True)
        eigen_csv(os.path.join(d, "X.txt"), np.array(Xs)); eigen_csv(os.path.join(d, "U.txt"), np.array(Us))
        with open(os.path.join(d, "t.txt"), "w") as f:
            f.write("%.6g" % (step * args.time_step))
        print(d)
    eng.close()


if __name__ == "__main__":
    main()
