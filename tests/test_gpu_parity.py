"""GPU parity tests (run with -m gpu on the B200 box): the CUDA engine, called through the C-ABI, against the CPU oracle.

PARITY UNPINNED: the reference has no golden vectors; the oracle (oracle/*.c) is a literal restatement certified by KKT
residuals (tests/test_oracle.py).  Tolerances are north_star's: 1e-5 on the state, 1e-4 on the control, in the units the
algorithm iterates on (nondimensional; thrusts are ~1e-2 there, SURVEY §7 "parity definition")."""
import ctypes as C
import os
import numpy as np
import pytest

import orc_py as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL_X, TOL_U = 1e-5, 1e-4
RPY_F9 = np.deg2rad([-20.0, 20.0, 0.0])


@pytest.fixture(scope="module")
def S():
    import scpp_b200
    assert scpp_b200.device_count() > 0, "no CUDA device"
    return scpp_b200


def _oracle_nondim_par(p):
    pn = O.RQParams.from_buffer_copy(p); O.lib().orc_rq_nondimensionalize(C.byref(pn))
    par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    return pn, par


def test_tensor_core_block_products(S):
    """mma.sync.m8n8k4.f64 block products (predicated 18-wide operands, lower-tile mode, 14-row contraction) == scalar loops"""
    assert S.selftest_blockops() < 1e-12


def test_discretize_matches_rkf78_oracle(S):
    """hot path 1 alone: RK4 x nsub forward-sensitivity kernel vs the literal RKF78 x 5 Phi^-1-form oracle"""
    p, rpy = O.falcon9()
    r = O.sc_solve(O.ROCKETQUAT, p, O.sc_config(K=50, max_iterations=3))
    pn, par = _oracle_nondim_par(p)
    for it in (0, 3):
        X, U, t = r["X_all"][it], r["U_all"][it], r["t_all"][it]
        ref = O.discretize(O.ROCKETQUAT, X, U, t, par)
        got = S.discretize(S.ROCKETQUAT, X, U, t, par, nsub=20)
        for key in ("A", "B", "C", "s", "z"):
            scale = np.abs(ref[key]).max()
            assert np.abs(got[key][0] - ref[key]).max() <= 2e-10 * max(1.0, scale), key
        # the identity the SOCP relies on: linear model == nonlinear propagation at the linearisation point
        for k in (0, 17, 48):
            lin = got["A"][0, k] @ X[k] + got["B"][0, k] @ U[k] + got["C"][0, k] @ U[k + 1] + got["s"][0, k] * t + got["z"][0, k]
            assert np.allclose(lin, O.simulate(O.ROCKETQUAT, t / 49, U[k], U[k + 1], par, X[k]), atol=1e-9)


def test_discretize_rocket2d_and_ragged_sizes(S):
    p2 = O.rocket2d()
    pn = O.R2DParams.from_buffer_copy(p2); O.lib().orc_r2d_nondimensionalize(C.byref(pn))
    par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    for K in (3, 7, 30):      # K=3: smallest the engine accepts; odd sizes exercise partial warps / blocks
        tol = 2e-10 if K >= 30 else 1e-5      # coarse grids: 6 s intervals, RK4 x 20 and RKF78 x 5 both carry truncation error
        X = np.zeros((K, 6)); U = np.zeros((K, 2)); t = C.c_double()
        O.lib().orc_r2d_initial_trajectory(C.byref(pn), K, X.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), C.byref(t))
        ref = O.discretize(O.ROCKET2D, X, U, t.value, par)
        got = S.discretize(S.ROCKET2D, X, U, t.value, par, nsub=20)
        for key in ("A", "B", "C", "s", "z"):
            assert np.abs(got[key][0] - ref[key]).max() <= tol * max(1.0, np.abs(ref[key]).max()), (K, key)


def _compare_run(S, name, model_o, params_list, K, max_it, tol_x=TOL_X, tol_u=TOL_U, xi=None, cfg_over=None, warm=0.0, ipm_slice=None):
    model, params, x_init, x_final, cfg = S.load_model(name, K=K, max_iterations=max_it, keep_history=1, **(cfg_over or {}))
    cfg.ipm.warm = warm
    if ipm_slice is not None:
        cfg.ipm_slice = ipm_slice
    N = len(params_list)
    if xi is None:
        xi = np.array([list(p.x_init) for p in params_list])
    eng = S.SCAlgorithm(model, params, cfg, N)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info()
    Xh, Uh, th = eng.get_all_solutions()
    ocfg = O.sc_config(K=K, model=model_o, max_iterations=max_it)
    for k, v in (cfg_over or {}).items():
        if hasattr(ocfg, k):
            setattr(ocfg, k, v)
    report = []
    for i, p in enumerate(params_list):
        ro = O.sc_solve(model_o, p, ocfg)
        n = abs(ro["iterations"])
        assert ro["iterations"] > 0, "oracle failed"
        assert sol["iterations"][i] == n, f"instance {i}: iteration count {sol['iterations'][i]} vs oracle {n}"
        assert (sol["flags"][i] == 1) == ro["converged"]
        for it in range(n + 1):
            ku = K if ocfg.interpolate_input else K - 1      # zero-order hold: the reference has no input column K - 1 (a pinned placeholder here)
            dX = np.abs(Xh[i, it] - ro["X_all"][it]).max(); dU = np.abs(Uh[i, it, :ku] - ro["U_all"][it][:ku]).max()
            assert dX < tol_x and dU < tol_u, f"instance {i} iterate {it}: dX {dX:.2e} dU {dU:.2e}"
            report.append((dX, dU))
        # same discrete decisions: weight doubling (SCAlgorithm.cpp:112-115) and convergence test (:131)
        for it in range(n):
            assert info[i, it, 4] == ro["info"][it].weight_tr_used
            assert abs(info[i, it, 0] - ro["info"][it].norm1_nu) < 1e-6 and abs(info[i, it, 1] - ro["info"][it].sum_delta) < 1e-4 * max(1., ro["info"][it].sum_delta)      # epigraph values: accurate to the duality gap, reltol 1e-8 x a cost of ~1e3-1e4
        # final trajectory is redimensionalised (SCAlgorithm.cpp:182-187)
        assert np.allclose(sol["X"][i], ro["X"], rtol=1e-6, atol=1e-4 * np.abs(ro["X"]).max())
    eng.close()
    return report


def test_sc_rocket2d_config0(S):
    """BASELINE.json configs[0]: Rocket2D SC K=30 single instance (the reference's CPU-runnable case)"""
    rep = _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15)
    assert len(rep) >= 3


def test_sc_rocketquat_k50_batch(S):
    """BASELINE.json configs[1] at oracle-checkable batch size: RocketQuat K=50, perturbed initial states"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(6)] + [p]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=6)


def test_sc_rocketquat_k50_interior_warm_start(S):
    """the opt-in interior warm start of the sub-problems (ipm.warm, what bench.py uses) reaches the same optimum: same parity bar"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(40, 46)] + [p]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=15, warm=0.995)
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15, warm=0.995)


def test_round_slicing_is_bit_identical(S):
    """the engine's rounds (one interior-point iteration per K2 launch, batch re-formed in between) and the lock-step mode
    (whole sub-problem per launch) run the same arithmetic: identical bits, instance by instance"""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=6, keep_history=1)
    cfg.ipm.warm = 0.995
    xi = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), 96)
    out = []
    for sl in (0, 1, 3):
        cfg.ipm_slice = sl
        eng = S.SCAlgorithm(model, params, cfg, 96)
        eng.set_boundary_states(xi, x_final)
        eng.solve()
        out.append((eng.get_all_solutions(), eng.get_info(), eng.get_solution()))
        eng.close()
    for o in out[1:]:
        for a, b in zip(o[0], out[0][0]):
            assert np.array_equal(a, b)
        assert np.array_equal(o[1], out[0][1])
        assert np.array_equal(o[2]["iterations"], out[0][2]["iterations"]) and np.array_equal(o[2]["flags"], out[0][2]["flags"])


def test_split_pipeline_vs_oracle(S):
    """the split pipeline (cfg.ipm_slice = -1: one kernel per step of the interior-point iteration, several warps per instance in the
    stage-parallel passes) against the oracle: K=50 (two 32-stage parts), K=30 (one), K=100 (four), cold and interior warm starts"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(20, 26)] + [p]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=6, ipm_slice=-1)
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist[:4], K=50, max_it=15, warm=0.995, ipm_slice=-1)
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15, warm=0.995, ipm_slice=-1)
    ps, rpys = O.starship()
    _compare_run(S, "RocketQuatStarship", O.ROCKETQUAT, [ps, O.rq_perturb(ps, rpys, 0x5C99, 1)], K=100, max_it=4, ipm_slice=-1)


def test_split_pipeline_matches_monolithic_kernel(S):
    """same algorithm, different decomposition over kernels and warps: identical iteration counts and decisions, iterates equal to
    rounding (the partial sums are added in a different order)"""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=8, keep_history=1)
    cfg.ipm.warm = 0.995
    xi = S.perturbed_initial_states(x_init, np.deg2rad([-20.0, 20.0, 0.0]), 200)
    out = []
    for sl in (1, -1):
        cfg.ipm_slice = sl
        eng = S.SCAlgorithm(model, params, cfg, 200)
        eng.set_boundary_states(xi, x_final)
        eng.solve()
        out.append((eng.get_all_solutions(), eng.get_info(), eng.get_solution()))
        eng.close()
    (Xa, Ua, ta), (Xb, Ub, tb) = out[0][0], out[1][0]
    assert np.array_equal(out[0][2]["iterations"], out[1][2]["iterations"]) and np.array_equal(out[0][2]["flags"], out[1][2]["flags"])
    assert np.abs(Xa - Xb).max() < 1e-5 and np.abs(Ua - Ub).max() < 1e-4
    assert np.array_equal(out[0][1][:, :, 4], out[1][1][:, :, 4])                 # trust-region weights used
    # interior-point iteration counts: the dual residual of the last iterates sits at the rounding floor of the condensed system
    # (~1e-8), so the iteration at which the test fires moves by a few when sums are re-ordered
    d_it = np.abs(out[0][1][:, :, 5] - out[1][1][:, :, 5])
    assert d_it.mean() < 0.5 and d_it.max() <= 10


def test_sc_rocketquat_starship_k100(S):
    """BASELINE.json configs[4]: Starship parameters, K=100"""
    p, rpy = O.starship()
    _compare_run(S, "RocketQuatStarship", O.ROCKETQUAT, [p, O.rq_perturb(p, rpy, 0x5C99, 1)], K=100, max_it=4)


def test_committed_fixtures(S):
    """the committed (oracle-generated) fixtures in tests/golden/, compared on the box without running the oracle"""
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    p, rpy = O.falcon9()
    cases = [("rocket2d_K30", "Rocket2D", 30, 15, np.array(O.rocket2d().x_init)),
             ("rocketquat_K20_nominal", "RocketQuat", 20, 5, np.array(p.x_init)),
             ("rocketquat_K50_inst7", "RocketQuat", 50, 4, np.array(O.rq_perturb(p, rpy, 0x5C99, 7).x_init))]
    for name, cfgname, K, max_it, xi in cases:
        g = np.load(os.path.join(gdir, name + ".npz"))
        model, params, x_init, x_final, cfg = S.load_model(cfgname, K=K, max_iterations=max_it, keep_history=1)
        eng = S.SCAlgorithm(model, params, cfg, 1)
        eng.set_boundary_states(xi, x_final)
        eng.solve()
        sol = eng.get_solution(); Xh, Uh, th = eng.get_all_solutions()
        n = int(g["iterations"])
        assert sol["iterations"][0] == n and int(sol["flags"][0] == 1) == int(g["converged"])
        assert np.abs(Xh[0, :n + 1] - g["X_all"]).max() < TOL_X and np.abs(Uh[0, :n + 1] - g["U_all"]).max() < TOL_U
        assert np.allclose(sol["X"][0], g["X"], rtol=1e-6, atol=1e-4 * np.abs(g["X"]).max())
        eng.close()


def test_full_batch_properties(S):
    """BASELINE.json configs[1] at full size (1024): size-independent properties instead of oracle comparison"""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50)
    N = 1024
    xi = S.perturbed_initial_states(x_init, RPY_F9, N)
    eng = S.SCAlgorithm(model, params, cfg, N)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    a = eng.get_solution(); info = eng.get_info(); t1 = eng.last_timing()
    eng.solve()
    b = eng.get_solution()
    # (1) determinism / idempotence of a cold solve
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["U"], b["U"]) and np.array_equal(a["iterations"], b["iterations"])
    # (2) no instance failed, every instance iterated 1..max_iterations times, accounting adds up
    assert (a["flags"] != 2).all()
    assert a["iterations"].min() >= 1 and a["iterations"].max() <= cfg.max_iterations
    assert t1["instance_iterations"] == int(a["iterations"].sum())
    # (3) boundary conditions hold for every instance (pinned variables): x_0 = x_init, final rows, last input
    assert np.allclose(a["X"][:, 0, :], xi, rtol=1e-12, atol=1e-9)
    fin = [1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13]
    assert np.abs(a["X"][:, -1, fin] - x_final[fin]).max() < 1e-6
    assert np.abs(a["U"][:, -1, [0, 1, 3]]).max() < 1e-6 and np.abs(a["U"][:, :, 3]).max() < 1e-9
    # (4) application constraints hold (dimensional): thrust bounds, gimbal, glide slope, dry mass
    Tn = np.linalg.norm(a["U"][:, :, :3], axis=2)
    assert Tn.max() <= params.T_max * (1 + 1e-6)
    assert (a["U"][:, :, 2] >= params.T_min * (1 - 1e-6)).all()                      # n = (0,0,1) on a cold start
    assert (np.linalg.norm(a["U"][:, :, :2], axis=2) <= np.tan(params.gimbal_max) * a["U"][:, :, 2] * (1 + 1e-6) + 1e-3).all()
    assert (np.linalg.norm(a["X"][:, :, 1:3], axis=2) <= np.tan(params.gamma_gs) * a["X"][:, :, 3] * (1 + 1e-6) + 1e-3).all()
    assert (a["X"][:, :, 0] >= x_final[0] * (1 - 1e-9)).all()
    # (5) every sub-problem carries a certificate: status optimal / reduced accuracy, small residuals
    for i in range(N):
        n = a["iterations"][i]
        assert set(info[i, :n, 6].astype(int)) <= {0, 3}
        assert info[i, :n, 7].max() < 1e-6 and info[i, :n, 8].max() < 1e-6
    # (6) batch independence: instance 777 solved alone gives the identical trajectory
    eng1 = S.SCAlgorithm(model, params, cfg, 1)
    eng1.set_boundary_states(xi[777:778], x_final)
    eng1.solve()
    c = eng1.get_solution()
    assert np.array_equal(c["X"][0], a["X"][777]) and c["iterations"][0] == a["iterations"][777]
    eng1.close(); eng.close()


def test_warm_start_and_errors(S):
    model, params, x_init, x_final, cfg = S.load_model("Rocket2D", K=30)
    eng = S.SCAlgorithm(model, params, cfg, 3)
    with pytest.raises(S.ScppError):
        eng.solve()                                  # boundary states not set
    eng.set_boundary_states(x_init, x_final)
    with pytest.raises(S.ScppError):
        eng.solve(warm_start=True)                   # nothing to warm-start from
    eng.solve()
    cold = eng.get_solution()
    assert (cold["flags"] == 1).all()
    eng.solve(warm_start=True)                       # SCAlgorithm::solve(true): starts from the converged trajectory
    warm = eng.get_solution()
    assert (warm["flags"] == 1).all() and (warm["iterations"] <= cold["iterations"]).all()
    assert np.allclose(warm["X"], cold["X"], atol=1e-3 * np.abs(cold["X"]).max())
    eng.close()
    bad = S.default_config(model); bad.interpolate_input = 0; bad.algorithm = 1      # zero-order-hold inputs are built for SC only: SCvx with them is reported, not ignored
    bad.scvx_trust_region = 1.0; bad.scvx_alpha = 2.0; bad.scvx_beta = 3.2
    with pytest.raises(S.ScppError, match="zero-order hold"):
        S.SCAlgorithm(model, params, bad, 1)


def test_k4_simulate_vs_oracle(S):
    """K4 alone through the C-ABI: scpp::simulate for a batch of states"""
    p, _ = O.falcon9()
    par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(p), par.ctypes.data_as(C.c_void_p))
    rng = np.random.default_rng(7)
    n = 64
    x = np.tile(np.array(p.x_init), (n, 1)); x[:, 1:7] *= 1 + 0.1 * rng.standard_normal((n, 6))
    u0 = np.tile([1e4, -2e4, 3e5, 0.], (n, 1)) * (1 + 0.1 * rng.standard_normal((n, 4))); u1 = u0 * 1.05
    got = S.simulate(S.ROCKETQUAT, x, u0, u1, par, 0.05)
    for i in range(n):
        ref = O.simulate(O.ROCKETQUAT, 0.05, u0[i], u1[i], par, x[i])
        assert np.abs(got[i] - ref).max() <= 1e-12 * np.abs(ref).max()


def test_closed_loop_vs_oracle(S):
    """scpp/src/SC_sim.cpp on the device: solve, K4 step (x_init advanced on the device), warm-started solve ... against the oracle's
    literal loop, for a small Monte-Carlo batch; instances that reach the end are frozen"""
    for name, model_o, K, steps, plist in (("Rocket2D", O.ROCKET2D, 30, 4, [O.rocket2d()]),
                                          ("RocketQuat", O.ROCKETQUAT, 20, 2, [O.falcon9()[0]] + [O.rq_perturb(*O.falcon9(), 0x5C99, i) for i in range(3)])):
        model, params, x_init, x_final, cfg = S.load_model(name, K=K, max_iterations=15)
        xi = np.array([list(p.x_init) for p in plist])
        eng = S.SCAlgorithm(model, params, cfg, len(plist))
        eng.set_boundary_states(xi, x_final)
        Xs, Us, its = [], [], []
        for s in range(steps):
            eng.solve(warm_start=s > 0)
            its.append(eng.get_solution()["iterations"].copy())
            r = eng.sim_step(0.05)
            Xs.append(r["x"]); Us.append(r["u0"])
        eng.close()
        Xs, Us, its = np.array(Xs), np.array(Us), np.array(its)
        ocfg = O.sc_config(K=K, model=model_o, max_iterations=15)
        for i, p in enumerate(plist):
            ro = O.sc_sim(model_o, p, ocfg, 0.05, steps)
            assert ro["steps"] == steps and np.array_equal(its[:, i], ro["iters"][:steps])
            sx, su = np.abs(ro["X_sim"]).max(), np.abs(ro["U_sim"]).max()
            assert np.abs(Xs[:, i] - ro["X_sim"]).max() < 1e-6 * sx and np.abs(Us[:, i] - ro["U_sim"]).max() < 1e-4 * su


def test_sc_oneshot_output_layout(S, tmp_path):
    """tools/sc_oneshot.py writes what scpp/src/SC_oneshot.cpp:29-63 writes: output/<Model>/SC/<time>/<k>/{X,U,t}.txt, redimensionalised,
    comma-separated, 6 significant digits -- readable by the reference's plotting scripts (np.loadtxt(..., delimiter=','))"""
    import subprocess, sys, glob
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "sc_oneshot.py"), "--model", "RocketQuat", "--K", "20", "--out", str(tmp_path)], text=True)
    root = out.strip().splitlines()[-1]
    its = sorted(int(os.path.basename(d)) for d in glob.glob(os.path.join(root, "*")))
    p, _ = O.falcon9()
    ro = O.sc_solve(O.ROCKETQUAT, p, O.sc_config(K=20, max_iterations=15))
    assert its == list(range(ro["iterations"] + 1))
    X = np.loadtxt(os.path.join(root, str(its[-1]), "X.txt"), delimiter=","); U = np.loadtxt(os.path.join(root, str(its[-1]), "U.txt"), delimiter=",")
    t = float(open(os.path.join(root, str(its[-1]), "t.txt")).read())
    assert X.shape == (20, 14) and U.shape == (20, 4)
    assert np.allclose(X, ro["X"], rtol=2e-5, atol=2e-5 * np.abs(ro["X"]).max()) and np.allclose(U, ro["U"], rtol=2e-4, atol=2e-5 * np.abs(ro["U"]).max())
    assert abs(t - ro["t"]) < 1e-4 * ro["t"]
    X0 = np.loadtxt(os.path.join(root, "0", "X.txt"), delimiter=",")           # iterate 0 = the initial guess, redimensionalised
    xi_, xf_ = np.array(p.x_init), np.array(p.x_final)
    assert np.allclose(X0[0], xi_, rtol=1e-5)
    # last node of the initial guess: alpha2 = (K-1)/K, the reference's denominator is K (rocketQuat.cpp:45-46)
    assert np.allclose(X0[-1, 1:7], (xi_[1:7] + 19 * xf_[1:7]) / 20, rtol=1e-5, atol=1e-3)


def test_k5_lqr_gains_vs_oracle(S):
    """LQRTracker gains through the C-ABI (one warp per (instance, node)) vs the oracle, Rocket2D K=30 with the reference's LQR.info
    weights; getInput's interpolation is the host helper lqr_input"""
    model, params, x_init, x_final, cfg = S.load_model("Rocket2D", K=30, max_iterations=15)
    eng = S.SCAlgorithm(model, params, cfg, 3)
    eng.set_boundary_states(x_init, x_final)
    eng.solve()
    sol = eng.get_solution()
    q, r = np.ones(6), np.array([2.0, 2.0])
    G, ok = eng.lqr_gains(q, r)
    eng.close()
    p2 = O.rocket2d()
    par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(p2), par.ctypes.data_as(C.c_void_p))
    pp = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(C.c_void_p)
    Gr = np.zeros((30, 2, 6)); okr = np.zeros(30, np.int32)
    O.lib().orc_lqr_tracker_gains(1, 30, pp(sol["X"][0]), pp(sol["U"][0]), pp(par), pp(q), pp(r), Gr.ctypes.data_as(C.c_void_p), okr.ctypes.data_as(C.c_void_p))
    assert ok.all() and okr.all()
    for i in range(3):
        rel = np.abs(G[i] - Gr).max(axis=(1, 2)) / np.abs(Gr).max(axis=(1, 2))
        assert rel.max() < 1e-4
    u = S.lqr_input(0.37 * sol["t"][0], sol["X"][0][3] * 1.01, sol["X"][0], sol["U"][0], sol["t"][0], G[0])
    assert u.shape == (2,) and np.isfinite(u).all()


def test_ragged_batch_sizes_and_a_failing_instance(S):
    """batch sizes that do not fill the launch shape (1, 149 = one more than the SMs, a K with few stages), and an instance whose
    sub-problem is infeasible (initial mass below the dry mass): the reference would std::terminate (SCAlgorithm.cpp:94-98); here it is
    flagged (flags == 2) and every other instance is bit-identical to its solo solve"""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=10, max_iterations=4)
    N = 149
    xi = S.perturbed_initial_states(x_init, RPY_F9, N)
    xi[5, 0] = 0.9 * x_final[0]                          # below m_dry: m_k >= m_dry cannot hold at k = 0
    eng = S.SCAlgorithm(model, params, cfg, N)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    a = eng.get_solution()
    eng.close()
    assert a["flags"][5] == 2 and (np.delete(a["flags"], 5) != 2).all()
    for i in (0, 6, 148):
        e1 = S.SCAlgorithm(model, params, cfg, 1)
        e1.set_boundary_states(xi[i:i + 1], x_final)
        e1.solve()
        c = e1.get_solution()
        e1.close()
        assert np.array_equal(c["X"][0], a["X"][i]) and np.array_equal(c["U"][0], a["U"][i]) and c["iterations"][0] == a["iterations"][i]
    # smallest horizon the engine accepts
    m3, p3, xi3, xf3, c3 = S.load_model("Rocket2D", K=3, max_iterations=3)
    e3 = S.SCAlgorithm(m3, p3, c3, 2)
    e3.set_boundary_states(xi3, xf3)
    e3.solve()
    s3 = e3.get_solution()
    e3.close()
    assert np.isfinite(s3["X"]).all() and np.array_equal(s3["X"][0], s3["X"][1])
    with pytest.raises(S.ScppError):
        S.load_model("Rocket2D", K=2) and S.SCAlgorithm(*S.load_model("Rocket2D", K=2)[:2], S.load_model("Rocket2D", K=2)[4], 1)


def test_scvx_vs_oracle(S):
    """SCvx on the device (algorithm = 1): nominal Falcon-9 instance and perturbed ones, K = 30, the reference's SCvx.info.
    The reference's loop is not reproducible decision by decision: after a rejected step it overwrites last_nonlinear_cost, and while
    the trust region is inactive the re-solve returns the same candidate, so rho = (rounding noise) / predicted and its SIGN decides
    accept or reject (SCvxAlgorithm.cpp:116-139).  The number of solves per iteration and the radii therefore differ between any two
    implementations, while the accepted iterates agree as long as the radius is inactive.  Asserted: every instance converges like the
    oracle's, the leading iterates agree to 5e-5 / 5e-6 (at least the first three), the final nonlinear cost agrees to 5 %."""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", algorithm="SCvx", keep_history=1)
    assert cfg.algorithm == 1 and cfg.K == 30 and cfg.max_iterations == 30
    cfg.ipm.warm = 0.995
    p, rpy = O.falcon9()
    plist = [p] + [O.rq_perturb(p, rpy, 0x5C99, i) for i in (0, 2, 3)]
    xi = np.array([list(q.x_init) for q in plist])
    eng = S.SCAlgorithm(model, params, cfg, len(plist))
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions()
    eng.close()
    ocfg = O.scvx_config(K=30, model=O.ROCKETQUAT)
    for i, q in enumerate(plist):
        ro = O.scvx_solve(O.ROCKETQUAT, q, ocfg)
        n = int(sol["iterations"][i])
        assert ro["converged"] and sol["flags"][i] == 1 and abs(n - ro["iterations"]) <= 8
        m = 0
        for it in range(1, min(ro["iterations"], n) + 1):
            if np.abs(Xh[i, it] - ro["X_all"][it]).max() < 5e-5 and np.abs(Uh[i, it] - ro["U_all"][it]).max() < 5e-6:
                m = it
            else:
                break
        assert m >= 3, (i, m)
        for it in range(m):
            assert abs(ro["info"][it].norm1_nu - info[i, it, 0]) < 1e-4 * ro["info"][it].norm1_nu
            assert abs(ro["info"][it].nonlinear_cost - info[i, it, 1]) < 5e-3 * ro["info"][it].nonlinear_cost
        Jo, Jg = ro["info"][-1].nonlinear_cost, info[i, n - 1, 1]
        assert abs(Jo - Jg) < 0.05 * Jo


def test_cpp_host_mirror_runs_sc_and_scvx(S, tmp_path):
    """the C++ host mirror (include/scpp_b200.hpp) end to end on the device: SC and SCvx for the nominal RocketQuat instance"""
    import subprocess
    exe = str(tmp_path / "cpp_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "scpp_b200"), "-lscpp_b200", "-Wl,-rpath," + os.path.join(ROOT, "scpp_b200")])
    out = subprocess.run([exe, os.path.join(ROOT, "configs"), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "ok (GPU)" in out.stdout and "flag=1" in out.stdout, out.stdout + out.stderr


# ---- round 2: parity at the shapes that are measured ------------------------------------------------------------------------
def _oracle_many(fn, items, threads=None):
    """the oracle on several instances at once (ctypes releases the GIL; one instance per host thread)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 4)) as ex:
        return list(ex.map(fn, items))


def _status_fractions(info, iters):
    """fractions of the solved sub-problems by exit status: 0 = tolerances met, 3 = reduced accuracy (best iterate inside the band)"""
    st = np.concatenate([info[i, :n, 6] for i, n in enumerate(iters)]).astype(int)
    sc = np.concatenate([np.maximum(info[i, :n, 7], info[i, :n, 8]) for i, n in enumerate(iters)])
    return {s: float((st == s).mean()) for s in (0, 1, 2, 3)}, float(sc.max())


def test_bench_shape_sample_from_the_full_batch(S):
    """BASELINE.json configs[1] exactly as bench.py runs it (1024 perturbed RocketQuat instances, K = 50, interior warm start, one
    interior-point iteration per launch, Richardson RK4): 64 randomly indexed instances OF THAT BATCH against the oracle, all 15 iterates,
    plus the exit-status census of the 15 360 sub-problems"""
    N = 1024
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, keep_history=1)
    cfg.ipm.warm = 0.995
    xi = S.perturbed_initial_states(x_init, RPY_F9, N)
    eng = S.SCAlgorithm(model, params, cfg, N)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions()
    eng.close()
    idx = np.sort(np.random.default_rng(2024).choice(N, 64, replace=False))
    p, rpy = O.falcon9()
    ocfg = O.sc_config(K=50, max_iterations=cfg.max_iterations)
    ros = _oracle_many(lambda i: O.sc_solve(O.ROCKETQUAT, O.rq_perturb(p, rpy, 0x5C99, int(i)), ocfg), idx)
    worst = (0.0, 0.0)
    for i, ro in zip(idx, ros):
        n = abs(ro["iterations"])
        assert ro["iterations"] > 0 and sol["iterations"][i] == n and (sol["flags"][i] == 1) == ro["converged"]
        assert np.allclose(Xh[i, 0], ro["X_all"][0], atol=1e-12)          # the batch generator and the oracle's perturbation are the same instance
        for it in range(n + 1):
            dX = np.abs(Xh[i, it] - ro["X_all"][it]).max(); dU = np.abs(Uh[i, it] - ro["U_all"][it]).max()
            assert dX < TOL_X and dU < TOL_U, f"instance {i} iterate {it}: dX {dX:.2e} dU {dU:.2e}"
            worst = (max(worst[0], dX), max(worst[1], dU))
        for it in range(n):
            assert info[i, it, 4] == ro["info"][it].weight_tr_used
    frac, resid = _status_fractions(info, sol["iterations"])
    print(f"\n[bench shape] worst |dX| {worst[0]:.2e} |dU| {worst[1]:.2e}; sub-problem exits: {frac}; worst residual {resid:.2e}")
    # every sub-problem ends with its tolerances met (0) or, at the accuracy floor of the condensed system, inside the 10x band (3)
    assert frac[1] == 0.0 and frac[2] == 0.0 and frac[3] <= 0.35 and resid < 1e-6


def test_converging_rocketquat_workload(S):
    """NON-REFERENCE weights (w_tr = 2, w_vc = 1e4, nu_tol = 1e-3, delta_tol = 1e-2 instead of SC.info's 50 / 1e3 / 1e-5 / 1e-3), found by
    experiment: with the shipped weights no RocketQuat instance converges.  Here some instances converge after 6-7
    iterations and others run to the limit, so convergence (SCAlgorithm.cpp:131), weight doubling (:112-115), early exit and the
    compaction of an unevenly finishing batch run on nx = 14.  Same decisions as the oracle; iterates to 5e-5 / 5e-4 for every instance that
    converges and over iterates 0-3 of the others (the weak trust region w_tr = 2 leaves the sub-problem optimum weakly determined: two correct
    solvers, or one solver from two starting points, differ by 2-3e-5 there), 1e-4 / 1e-3 over their iterates 4-6 (a stalled loop re-solves a weakly determined
    problem: the 1e-10 difference between the two discretisation schemes is amplified to ~2e-5 by iterate 4 and ~1e-4 later)."""
    over = dict(weight_trust_region_trajectory=2.0, weight_virtual_control=1e4, nu_tol=1e-3, delta_tol=1e-2)
    ids = [8, 20, 22, 0, 4, 5, 12, 15, 17, 19]
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=20, keep_history=1, **over)
    p, rpy = O.falcon9()
    plist = [p] + [O.rq_perturb(p, rpy, 0x5C99, i) for i in ids]
    xi = np.array([list(q.x_init) for q in plist])
    eng = S.SCAlgorithm(model, params, cfg, len(plist))
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions()
    eng.close()
    ocfg = O.sc_config(K=50, max_iterations=20)
    for k, v in over.items():
        setattr(ocfg, k, v)
    ros = _oracle_many(lambda q: O.sc_solve(O.ROCKETQUAT, q, ocfg), plist)
    n_conv = 0
    for i, ro in enumerate(ros):
        n = abs(ro["iterations"])
        assert ro["iterations"] > 0 and sol["iterations"][i] == n and (sol["flags"][i] == 1) == ro["converged"], (i, sol["iterations"][i], ro["iterations"])
        n_conv += int(ro["converged"])
        upto = n if ro["converged"] else 6
        for it in range(upto + 1):
            dX = np.abs(Xh[i, it] - ro["X_all"][it]).max(); dU = np.abs(Uh[i, it] - ro["U_all"][it]).max()
            loose = 10. if (not ro["converged"] and it > 3) else 5.      # w_tr = 2 instead of 50: the sub-problems are weakly determined
            assert dX < loose * TOL_X and dU < loose * TOL_U, f"instance {i} iterate {it}: dX {dX:.2e} dU {dU:.2e}"
        for it in range(n):
            assert info[i, it, 4] == ro["info"][it].weight_tr_used
    assert n_conv >= 4 and n_conv < len(plist)            # a mixed batch: early exits next to instances that use every iteration
    assert len(set(int(v) for v in sol["iterations"])) >= 3


def test_starship_k100_batch(S):
    """BASELINE.json configs[4]: Starship parameters, K = 100, 32 perturbed instances, all 15 iterations against the oracle"""
    ps, rpys = O.starship()
    plist = [ps] + [O.rq_perturb(ps, rpys, 0x5C99, i) for i in range(31)]
    _compare_run_threaded(S, "RocketQuatStarship", plist, K=100, max_it=15, warm=0.995)


def _compare_run_threaded(S, name, plist, K, max_it, warm=0.0):
    model, params, x_init, x_final, cfg = S.load_model(name, K=K, max_iterations=max_it, keep_history=1)
    cfg.ipm.warm = warm
    xi = np.array([list(q.x_init) for q in plist])
    eng = S.SCAlgorithm(model, params, cfg, len(plist))
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions()
    eng.close()
    ocfg = O.sc_config(K=K, max_iterations=max_it)
    ros = _oracle_many(lambda q: O.sc_solve(O.ROCKETQUAT, q, ocfg), plist)
    for i, ro in enumerate(ros):
        n = abs(ro["iterations"])
        assert ro["iterations"] > 0 and sol["iterations"][i] == n and (sol["flags"][i] == 1) == ro["converged"]
        for it in range(n + 1):
            dX = np.abs(Xh[i, it] - ro["X_all"][it]).max(); dU = np.abs(Uh[i, it] - ro["U_all"][it]).max()
            assert dX < TOL_X and dU < TOL_U, f"instance {i} iterate {it}: dX {dX:.2e} dU {dU:.2e}"
        for it in range(n):
            assert info[i, it, 4] == ro["info"][it].weight_tr_used


def test_scvx_k50_full_batch_no_failures_and_oracle_sample(S):
    """the metric's own shape for the SCvx variant: K = 50, 1024 perturbed instances.  (1) no instance ends flagged: sub-problems whose
    optimal cost is zero used to lose positive definiteness (2.7 % of the batch in round 1); the on-demand regularisation (ipm.cuh: dcap)
    carries them through.  (2) a sample of the batch against the oracle: same convergence, the leading accepted iterates to north_star's
    1e-5 / 1e-4 (the loop is not reproducible decision by decision once rounding decides a ratio test, see test_scvx_vs_oracle)."""
    N = 1024
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", algorithm="SCvx", K=50, keep_history=1)
    cfg.ipm.warm = 0.995
    xi = S.perturbed_initial_states(x_init, RPY_F9, N)
    eng = S.SCAlgorithm(model, params, cfg, N)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); info = eng.get_info(); Xh, Uh, th = eng.get_all_solutions()
    eng.close()
    failed = float((sol["flags"] == 2).mean())
    print(f"\n[SCvx K=50 x {N}] failed fraction {failed}, converged fraction {(sol['flags'] == 1).mean():.3f}, mean outer iterations {sol['iterations'].mean():.1f}")
    assert failed == 0.0
    assert (sol["flags"] == 1).mean() > 0.95
    idx = [24, 29, 93, 7, 100, 513, 777, 1000]              # 24, 29, 93: zero-cost second sub-problem (failed in round 1)
    p, rpy = O.falcon9()
    ocfg = O.scvx_config(K=50, model=O.ROCKETQUAT)
    ros = _oracle_many(lambda i: O.scvx_solve(O.ROCKETQUAT, O.rq_perturb(p, rpy, 0x5C99, int(i)), ocfg), idx)
    lead, xs = [], []
    for i, ro in zip(idx, ros):
        n = int(sol["iterations"][i])
        if ro["iterations"] <= 0:
            continue                                        # the oracle's own solver gave up on this instance (zero-cost sub-problem)
        assert ro["converged"] and sol["flags"][i] == 1
        # The sub-problem minimises |nu|_1 alone: U and the costs are unique, X is only weakly determined (states trade against virtual
        # control at equal cost).  Leading iterates are counted on U, norm1_nu and the nonlinear cost; X is recorded beside them.
        m, zero_cost = 0, False
        for it in range(1, min(ro["iterations"], n) + 1):
            a = ro["info"][it - 1]
            if a.norm1_nu < 1e-8:      # optimal cost zero: the optimal set is a face, U is not unique either; the two loops part here and both converge
                zero_cost = True
                break
            if (np.abs(Uh[i, it] - ro["U_all"][it]).max() < TOL_U and abs(a.norm1_nu - info[i, it - 1, 0]) < 1e-4 * max(a.norm1_nu, 1e-6)
                    and abs(a.nonlinear_cost - info[i, it - 1, 1]) < 2e-2 * a.nonlinear_cost):      # J is evaluated on the weakly determined X
                m = it
                xs.append(np.abs(Xh[i, it] - ro["X_all"][it]).max() < TOL_X)
            else:
                break
        if not zero_cost:
            lead.append(m)
        Jo, Jg = ro["info"][-1].nonlinear_cost, info[i, n - 1, 1]
        assert abs(Jo - Jg) < 0.05 * max(Jo, 1e-3), (i, Jo, Jg)
    print(f"[SCvx K=50] leading iterates equal to the oracle's (U to 1e-4, costs): {lead}; of those with X to 1e-5: {np.mean(xs):.2f}")
    assert len(lead) >= 4 and min(lead) >= 3 and np.mean(xs) >= 0.6


# ---- round 2: the CTA-per-instance solver (cfg.solver = 1) -------------------------------------------------------------------------------
def _solve_batch(S, name, xi, x_final_override=None, N=None, **over):
    solver = over.pop("solver", 1); warm = over.pop("warm", 0.0)
    model, params, x_init, x_final, cfg = S.load_model(name, keep_history=1, **over)
    cfg.solver = solver; cfg.ipm.warm = warm
    eng = S.SCAlgorithm(model, params, cfg, len(xi))
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    out = (eng.get_solution(), eng.get_info(), eng.get_all_solutions(), eng.last_timing())
    eng.close()
    return out


def test_cta_solver_vs_oracle(S):
    """cfg.solver = 1 (one CTA per instance, factor in shared memory, whole sub-problem per launch) against the oracle: RocketQuat K = 50
    cold and with the interior warm start, Rocket2D K = 30 (converges), Starship K = 100 (one CTA per SM)"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(60, 66)] + [p]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=6, cfg_over=dict(solver=1))
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist[:4], K=50, max_it=15, warm=0.995, cfg_over=dict(solver=1))
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15, cfg_over=dict(solver=1))
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15, warm=0.995, cfg_over=dict(solver=1))
    ps, rpys = O.starship()
    _compare_run(S, "RocketQuatStarship", O.ROCKETQUAT, [ps, O.rq_perturb(ps, rpys, 0x5C99, 1)], K=100, max_it=4, cfg_over=dict(solver=1))


def test_cta_solver_matches_warp_solver_on_a_batch(S):
    """same method, different mapping (sums in a different order): identical decisions and iteration counts on 300 instances (more than
    the 296 resident CTAs: the queue hands out a second instance to some), iterates equal far inside the parity bar; and the batch result
    does not depend on which CTA solved which instance (two runs are bit-identical)"""
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=8)
    xi = S.perturbed_initial_states(x_init, RPY_F9, 300)
    a = _solve_batch(S, "RocketQuat", xi, K=50, max_iterations=8, solver=0, warm=0.995)
    b = _solve_batch(S, "RocketQuat", xi, K=50, max_iterations=8, solver=1, warm=0.995)
    c = _solve_batch(S, "RocketQuat", xi, K=50, max_iterations=8, solver=1, warm=0.995)
    assert np.array_equal(a[0]["iterations"], b[0]["iterations"]) and np.array_equal(a[0]["flags"], b[0]["flags"])
    (Xa, Ua, ta), (Xb, Ub, tb) = a[2], b[2]
    assert np.abs(Xa - Xb).max() < TOL_X and np.abs(Ua - Ub).max() < TOL_U
    assert np.array_equal(a[1][:, :, 4], b[1][:, :, 4])
    assert np.array_equal(b[2][0], c[2][0]) and np.array_equal(b[2][1], c[2][1]) and np.array_equal(b[1], c[1])


def test_two_gpu_sharded_solve_is_bit_identical(S):
    """multi-GPU (needs >= 2 devices, skipped on a one-GPU box; run with gpurun --gpus 2): an unevenly sharded batch of the converging workload
    == the single-GPU solve, bit for bit, with the asynchronous per-outer-iteration flag exchange (tools/multi_gpu_check.py)"""
    import subprocess, sys
    if S.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "multi_gpu_check.py"), "300"], capture_output=True, text=True, timeout=900)
    print(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


def test_dual_number_jacobians_equal_the_hand_derived_ones(S):
    """K1 with cfg.jacobian = 1 (forward-mode dual numbers over the model's generic-scalar flow map: nothing model-specific beyond
    systemFlowMap, the role CppAD plays in the reference, systemDynamics.hpp:110-235) against jacobian = 0 (hand-derived sparse Jacobian):
    equal to rounding on both models; a whole SC solve on the optional fast path still matches the oracle"""
    p, rpy = O.falcon9()
    r = O.sc_solve(O.ROCKETQUAT, p, O.sc_config(K=50, max_iterations=2))
    pn, par = _oracle_nondim_par(p)
    X, U, t = r["X_all"][2], r["U_all"][2], r["t_all"][2]
    a = S.discretize(S.ROCKETQUAT, X, U, t, par, nsub=-5, jacobian=1); b = S.discretize(S.ROCKETQUAT, X, U, t, par, nsub=-5, jacobian=0)
    for key in ("A", "B", "C", "s", "z"):
        assert np.abs(a[key] - b[key]).max() <= 1e-13 * max(1.0, np.abs(b[key]).max()), key
    p2 = O.rocket2d()
    pn2 = O.R2DParams.from_buffer_copy(p2); O.lib().orc_r2d_nondimensionalize(C.byref(pn2))
    par2 = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn2), par2.ctypes.data_as(C.c_void_p))
    X2 = np.zeros((30, 6)); U2 = np.zeros((30, 2)); t2 = C.c_double()
    O.lib().orc_r2d_initial_trajectory(C.byref(pn2), 30, X2.ctypes.data_as(C.c_void_p), U2.ctypes.data_as(C.c_void_p), C.byref(t2))
    a = S.discretize(S.ROCKET2D, X2, U2, t2.value, par2, nsub=20, jacobian=1); b = S.discretize(S.ROCKET2D, X2, U2, t2.value, par2, nsub=20, jacobian=0)
    for key in ("A", "B", "C", "s", "z"):
        assert np.abs(a[key] - b[key]).max() <= 1e-13 * max(1.0, np.abs(b[key]).max()), key
    assert S.default_config(S.ROCKETQUAT).jacobian == 1
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, [p, O.rq_perturb(p, rpy, 0x5C99, 3)], K=50, max_it=4, cfg_over=dict(jacobian=0))


def test_shared_linearisation_k1_equals_the_column_kernels(S):
    """K1 with cfg.jacobian = 2 (discretize_shared.cuh: per CTA a chain warp integrates x(tau) two steps ahead, a lineariser warp leaves the stage records
    in shared memory one step ahead, and one column warp per interval integrates the 24 columns out of them) against the column kernels (0: same hand-derived Jacobian; 1: dual numbers) and the
    RKF78 oracle; batch sizes that leave the last CTA partly empty; a whole SC solve on this path against the oracle"""
    p, rpy = O.falcon9()
    r = O.sc_solve(O.ROCKETQUAT, p, O.sc_config(K=50, max_iterations=2))
    pn, par = _oracle_nondim_par(p)
    X, U, t = r["X_all"][2], r["U_all"][2], r["t_all"][2]
    ref = O.discretize(O.ROCKETQUAT, X, U, t, par)
    rng = np.random.default_rng(11)
    for n in (1, 3, 37):      # 49, 147, 1813 intervals: none a multiple of the 6 (or 13) intervals of a CTA
        Xb = X[None] * (1.0 + 1e-3 * rng.standard_normal((n,) + X.shape)); Ub = U[None] * (1.0 + 1e-3 * rng.standard_normal((n,) + U.shape))
        Xb[0], Ub[0] = X, U
        tb = t * (1.0 + 0.01 * rng.standard_normal(n)); tb[0] = t
        for nsub in (-5, 20):
            a = S.discretize(S.ROCKETQUAT, Xb, Ub, tb, par, nsub=nsub, jacobian=2)
            b = S.discretize(S.ROCKETQUAT, Xb, Ub, tb, par, nsub=nsub, jacobian=0)
            c = S.discretize(S.ROCKETQUAT, Xb, Ub, tb, par, nsub=nsub, jacobian=1)
            for key in ("A", "B", "C", "s", "z"):
                sc = max(1.0, np.abs(b[key]).max())
                assert np.abs(a[key] - b[key]).max() <= 1e-13 * sc, (n, nsub, key)
                assert np.abs(a[key] - c[key]).max() <= 1e-13 * sc, (n, nsub, key)
                assert np.abs(a[key][0] - ref[key]).max() <= 2e-10 * max(1.0, np.abs(ref[key]).max()), (n, nsub, key)
    p2 = O.rocket2d()
    pn2 = O.R2DParams.from_buffer_copy(p2); O.lib().orc_r2d_nondimensionalize(C.byref(pn2))
    par2 = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn2), par2.ctypes.data_as(C.c_void_p))
    for K in (3, 30):
        X2 = np.zeros((K, 6)); U2 = np.zeros((K, 2)); t2 = C.c_double()
        O.lib().orc_r2d_initial_trajectory(C.byref(pn2), K, X2.ctypes.data_as(C.c_void_p), U2.ctypes.data_as(C.c_void_p), C.byref(t2))
        a = S.discretize(S.ROCKET2D, X2, U2, t2.value, par2, nsub=-5, jacobian=2); b = S.discretize(S.ROCKET2D, X2, U2, t2.value, par2, nsub=-5, jacobian=0)
        for key in ("A", "B", "C", "s", "z"):
            assert np.abs(a[key] - b[key]).max() <= 1e-13 * max(1.0, np.abs(b[key]).max()), (K, key)
    # plugin models (Lin generated by dual numbers, dense): the small one fits the stash, the 14-state one does not and takes path 1
    a = S.discretize(S.ROCKET2D_PLUGIN, X2, U2, t2.value, par2, nsub=-5, jacobian=2); b = S.discretize(S.ROCKET2D_PLUGIN, X2, U2, t2.value, par2, nsub=-5, jacobian=1)
    for key in ("A", "B", "C", "s", "z"):
        assert np.abs(a[key] - b[key]).max() <= 1e-13 * max(1.0, np.abs(b[key]).max()), key
    a = S.discretize(S.ROCKETQUAT_ROLL, X, U, t, par, nsub=-5, jacobian=2); b = S.discretize(S.ROCKETQUAT_ROLL, X, U, t, par, nsub=-5, jacobian=1)
    assert all(np.array_equal(a[key], b[key]) for key in a)
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, [p, O.rq_perturb(p, rpy, 0x5C99, 3), O.rq_perturb(p, rpy, 0x5C99, 4)], K=50, max_it=4, cfg_over=dict(jacobian=2))
    _compare_run(S, "Rocket2D", O.ROCKET2D, [p2], K=30, max_it=15, cfg_over=dict(jacobian=2))


def test_zero_order_hold_inputs(S):
    """interpolate_input = false (discretizationImplementation.hpp:41-50,96-101; SCProblem.cpp:49-56,114-121; final-input constraints on
    column K - 2, rocketQuat.cpp:109-111): every iterate against the oracle's literal K - 1 input columns, on both K1 paths and both models"""
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=15, cfg_over=dict(interpolate_input=0))
    p, rpy = O.falcon9()
    plist = [p, O.rq_perturb(p, rpy, 0x5C99, 3), O.rq_perturb(p, rpy, 0x5C99, 5)]
    for jac in (1, 2):
        _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=5, cfg_over=dict(interpolate_input=0, jacobian=jac))
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist[:2], K=50, max_it=5, cfg_over=dict(interpolate_input=0, jacobian=2), warm=0.995)
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=20, max_iterations=3, keep_history=1, interpolate_input=0)
    eng = S.SCAlgorithm(model, params, cfg, 2)
    eng.set_boundary_states(np.array([list(q.x_init) for q in plist[:2]]), x_final)
    eng.solve()
    Xh, Uh, th = eng.get_all_solutions(); sol = eng.get_solution()
    assert np.array_equal(Uh[:, 1:, 19], Uh[:, :-1, 19]) and (Uh[:, :, 19, 2] > 0).all()      # the placeholder column never moves
    assert np.abs(sol["U"][:, 18, [0, 1, 3]]).max() < 1e-9                                  # u_x = u_y = torque = 0 on the last real column
    eng.close()


def test_per_instance_model_parameters(S):
    """scpp_b200_set_instance_params: every instance of the batch has its own vehicle (inertia, specific impulse, thrust limits, glide
    slope); each one against the oracle run with that instance's parameters"""
    import copy
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=30, max_iterations=5, keep_history=1)
    p, rpy = O.falcon9()
    plist, mlist = [], []
    for i, (isp, jscale, tmax, gs) in enumerate([(275., 1.0, 420000., 30.), (300., 1.2, 400000., 35.), (260., 0.8, 450000., 25.), (290., 1.1, 430000., 40.)]):
        q = O.rq_perturb(p, rpy, 0x5C99, 200 + i)
        q.alpha_m = 1. / (isp * 9.81); q.T_max = tmax; q.gamma_gs = np.deg2rad(gs)
        for a in range(3):
            q.J_B[a] = p.J_B[a] * jscale
        plist.append(q)
        m = S.ModelParams.from_buffer_copy(params)
        m.alpha_m = q.alpha_m; m.T_max = q.T_max; m.gamma_gs = q.gamma_gs
        for a in range(3):
            m.J_B[a] = q.J_B[a]
        mlist.append(m)
    xi = np.array([list(q.x_init) for q in plist])
    eng = S.SCAlgorithm(model, params, cfg, len(plist))
    eng.set_instance_params(mlist)
    eng.set_boundary_states(xi, x_final)
    eng.solve()
    sol = eng.get_solution(); Xh, Uh, th = eng.get_all_solutions()
    ocfg = O.sc_config(K=30, max_iterations=5)
    for i, q in enumerate(plist):
        ro = O.sc_solve(O.ROCKETQUAT, q, ocfg)
        n = abs(ro["iterations"])
        assert ro["iterations"] > 0 and sol["iterations"][i] == n
        for it in range(n + 1):
            assert np.abs(Xh[i, it] - ro["X_all"][it]).max() < TOL_X and np.abs(Uh[i, it] - ro["U_all"][it]).max() < TOL_U, (i, it)
    # back to the shared parameters: instance 0 (whose vehicle is the nominal one) is unchanged, the others are not
    eng.set_instance_params(None)
    eng.solve()
    sol2 = eng.get_solution()
    eng.close()
    assert np.array_equal(sol2["X"][0], sol["X"][0]) and not np.array_equal(sol2["X"][1], sol["X"][1])


def test_mpc_vs_oracle(S):
    """MPCAlgorithm on the device (kernel K6, one thread per instance) for a Monte-Carlo batch of Rocket2D states against the
    oracle's conic solver on the full problem of buildMPCProblem + addApplicationConstraints (tests/mpc_ref.py); exactLinearDiscretization
    against scipy's matrix exponential; K = 7 (the shipped MPC.info) and K = 21 (a 20-step horizon, BASELINE configs[3])"""
    import mpc_ref as R
    model, params, x_init, x_final, _ = S.load_model("Rocket2D")
    params.constrain_initial_final = 0
    p = O.rocket2d()
    rng = np.random.default_rng(11)
    for K, horizon, gimbal_bar, x_bar in ((7, 1.5, 1e-4, 1e-5), (21, 2.5, 2e-2, 1e-4)):     # bars: see tests/test_host.py::test_mpc_dense_conic_source_vs_oracle
        cfg = S.load_mpc_info(os.path.join(S.CONFIG_DIR, "Rocket2D", "MPC.info"), S.ROCKET2D)
        cfg.K = K; cfg.time_horizon = horizon
        N = 64
        x0 = np.array([-20., 100., 2., -10., 0.05, 0.0]) * (1 + 0.1 * rng.standard_normal((N, 6)))
        xf = np.array([0., 0, 0, -1, 0, 0.])
        mpc = S.MPCAlgorithm(model, params, cfg, N)
        A, B, z = mpc.discretization()
        Ar, Br, zr = R.discretize(p, horizon / (K - 1))
        assert np.abs(A - Ar).max() < 1e-8 and np.abs(B - Br).max() < 1e-8 * max(1., np.abs(Br).max()) and np.abs(z - zr).max() < 1e-8
        mpc.set_states(x0, xf)
        mpc.solve()
        sol = mpc.get_solution()
        assert (np.isin(sol["status"], (0, 3))).mean() > 0.9 and sol["iterations"].max() < 60
        w_term = np.array(list(cfg.state_weights_terminal)[:6]); w_in = np.array(list(cfg.input_weights)[:2])
        compared = 0
        for i in range(0, N, 8):
            P = R.full_socp(p, K, A, B, z, x0[i], xf, w_term, w_in)
            r = R.solve_with_oracle(O, P)
            if r["status"] != 0:
                continue
            compared += 1
            assert sol["status"][i] in (0, 3)
            Xo = np.array([[r["x"][P["iX"](k, j)] for j in range(6)] for k in range(K)]); Uo = np.array([[r["x"][P["iU"](k, j)] for j in range(2)] for k in range(K - 1)])
            cost = lambda X_, U_: np.linalg.norm(w_term * (X_[-1] - xf)) + np.linalg.norm(U_ * w_in)
            assert abs(cost(sol["X"][i], sol["U"][i]) - cost(Xo, Uo)) < 1e-6 * cost(Xo, Uo), (K, i)
            assert abs(sol["U"][i][0, 0] - Uo[0, 0]) < 1e-4                       # the input MPC applies (MPC_sim.cpp:64-70): north_star's bar
            assert np.abs(sol["U"][i][:, 0] - Uo[:, 0]).max() < gimbal_bar and np.abs(sol["U"][i][:, 1] - Uo[:, 1]).max() < 1e-5 * np.abs(Uo[:, 1]).max(), (K, i)
            assert np.abs(sol["X"][i] - Xo).max() < x_bar * max(1., np.abs(Xo).max()), (K, i)
        assert compared >= 5
        # closed loop (MPC_sim.cpp:64-70): the simulated state is the next initial state; the oracle's RKF78 gives the same step
        x1 = mpc.sim_step(0.05)
        par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(p), par.ctypes.data_as(C.c_void_p))
        ref = O.simulate(O.ROCKET2D, 0.05, sol["U"][0][0], sol["U"][0][0], par, x0[0])
        assert np.abs(x1[0] - ref).max() < 1e-9 * max(1., np.abs(ref).max())
        mpc.solve()
        assert (np.isin(mpc.get_solution()["status"], (0, 3))).mean() > 0.9
        mpc.close()
    with pytest.raises(S.ScppError):
        S.MPCAlgorithm(S.ROCKETQUAT, S.load_model("RocketQuat")[1], cfg, 1)          # no operating point: the reference throws too


def test_plugin_surface_model_vs_oracle(S):
    """a model added with NO hand-written Jacobian and NO hand-written row table (SCPP_B200_MODEL_ROCKET2D_PLUGIN: the planar
    rocket written against systemFlowMap / getInitializedTrajectory / addApplicationConstraints only, scpp_b200/plugins/rocket2d_plugin.hpp)
    on the device through the C-ABI against the oracle's Rocket2D: a batch of different initial states, every iterate, same decisions; K1
    with its dual-number Jacobians against the RKF78 oracle; SCvx and the closed loop run on it too"""
    d2r = np.pi / 180
    plist = []
    for i, (rx, ry, vy, eta) in enumerate([(-200, 800, -100, -20), (150, 700, -80, 10), (-50, 900, -120, 25), (300, 600, -60, -15), (0, 500, -90, 5)]):
        p = O.rocket2d(); p.x_init[:] = [rx, ry, 0, vy, eta * d2r, 0]
        plist.append(p)
    rep = _compare_run(S, "Rocket2DPlugin", O.ROCKET2D, plist, K=30, max_it=15)
    assert len(rep) >= 15
    _compare_run(S, "Rocket2DPlugin", O.ROCKET2D, plist[:2], K=30, max_it=15, warm=0.995, cfg_over=dict(solver=1))      # CTA-per-instance solver
    # K1 alone on the plugin model (dual numbers only: it has no hand-derived Jacobian) against the oracle's RKF78 discretisation
    p2 = O.rocket2d()
    pn2 = O.R2DParams.from_buffer_copy(p2); O.lib().orc_r2d_nondimensionalize(C.byref(pn2))
    par2 = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn2), par2.ctypes.data_as(C.c_void_p))
    ro = O.sc_solve(O.ROCKET2D, p2, O.sc_config(K=30, model=O.ROCKET2D, max_iterations=2))
    X, U, t = ro["X_all"][2], ro["U_all"][2], ro["t_all"][2]
    a = S.discretize(S.ROCKET2D_PLUGIN, X, U, t, par2, nsub=-5, jacobian=1); b = S.discretize(S.ROCKET2D_PLUGIN, X, U, t, par2, nsub=-5, jacobian=0)
    ref = O.discretize(O.ROCKET2D, X, U, t, par2)
    for key in ("A", "B", "C", "s", "z"):
        assert np.abs(a[key][0] - ref[key]).max() <= 2e-9 * max(1.0, np.abs(ref[key]).max()), key
        assert np.abs(a[key] - b[key]).max() <= 1e-13 * max(1.0, np.abs(b[key]).max()), key       # AutoJacobian (dense, duals) == the dual-number products


def test_rocketquat_roll_control_vs_oracle(S):
    """RocketQuat with enable_roll_control = true (model id 3, written against the plugin surface only: scpp_b200/plugins/rocketquat_plugin.hpp)
    on the device against the oracle's RocketQuat with enable_roll_control = 1: K = 50, a perturbed batch with initial roll rates, every
    iterate to 1e-5 / 1e-4, same decisions; both K2 mappings"""
    p, rpy = O.falcon9()
    plist = []
    for i in range(5):
        q = O.rq_perturb(p, rpy, 0x5C99, 40 + i)
        q.enable_roll_control = 1
        q.x_init[13] = 0.01 * (i - 2)
        plist.append(q)
    over = dict()
    model, params, x_init, x_final, cfg = S.load_model("RocketQuatRoll", K=50, max_iterations=6, keep_history=1)
    assert params.enable_roll_control == 1
    _compare_run(S, "RocketQuatRoll", O.ROCKETQUAT, plist, K=50, max_it=6)
    _compare_run(S, "RocketQuatRoll", O.ROCKETQUAT, plist[:3], K=50, max_it=6, warm=0.995, cfg_over=dict(solver=1))
    with pytest.raises(S.ScppError):
        S.SCAlgorithm(S.ROCKETQUAT, params, cfg, 1)             # the hand-written model refuses roll control and names model 3


def test_fixed_final_time_sc_vs_oracle(S):
    """SC with free_final_time = false (a setting the reference dispatches on, SCProblem.cpp:27-35 / discretization.cpp:42-55; not used by the
    shipped SC.info): RocketQuat K = 50 perturbed batch and Rocket2D K = 30 on the device against the oracle with free_final_time = 0, both K2
    mappings; the final time of every instance stays at final_time"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(70, 74)] + [p]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=6, cfg_over=dict(free_final_time=0))
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist[:2], K=50, max_it=6, warm=0.995, cfg_over=dict(free_final_time=0, solver=1))
    _compare_run(S, "Rocket2D", O.ROCKET2D, [O.rocket2d()], K=30, max_it=8, cfg_over=dict(free_final_time=0))
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=20, max_iterations=3, free_final_time=0)
    eng = S.SCAlgorithm(model, params, cfg, 2)
    eng.set_boundary_states(np.tile(x_init, (2, 1)), x_final)
    eng.solve()
    assert np.all(eng.get_solution()["t"] == params.final_time)
    eng.close()


def test_hybrid_tail_solver_vs_oracle_and_warp_solver(S):
    """cfg.solver = 2: the warp solver, with the sub-problems that start in the TAIL of a solve (fewer unfinished instances than resident CTAs)
    on the CTA-per-instance solver.  A batch of 200 is a tail from the first round on (every sub-problem on the CTA solver), 1024 switches near
    the end: same decisions and iteration counts as the warp solver alone, iterates far inside the parity bar; a sample against the oracle"""
    p, rpy = O.falcon9()
    plist = [O.rq_perturb(p, rpy, 0x5C99, i) for i in range(90, 94)]
    _compare_run(S, "RocketQuat", O.ROCKETQUAT, plist, K=50, max_it=6, warm=0.995, cfg_over=dict(solver=2))
    model, params, x_init, x_final, cfg = S.load_model("RocketQuat", K=50, max_iterations=15)
    for N in (200, 1024):
        xi = S.perturbed_initial_states(x_init, RPY_F9, N)
        a = _solve_batch(S, "RocketQuat", xi, K=50, max_iterations=15, solver=0, warm=0.995)
        b = _solve_batch(S, "RocketQuat", xi, K=50, max_iterations=15, solver=2, warm=0.995)
        assert np.array_equal(a[0]["iterations"], b[0]["iterations"]) and np.array_equal(a[0]["flags"], b[0]["flags"])
        (Xa, Ua, ta), (Xb, Ub, tb) = a[2], b[2]
        assert np.abs(Xa - Xb).max() < TOL_X and np.abs(Ua - Ub).max() < TOL_U
        assert np.array_equal(a[1][:, :, 4], b[1][:, :, 4])                      # same weight-doubling decisions
