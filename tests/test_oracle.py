"""CPU tests of the oracle (oracle/*.c): the reference ships no golden vectors (PARITY UNPINNED), so the oracle is
pinned by identities and certificates instead: exact-derivative checks, the linearisation identity of the
discretisation, KKT certificates verified independently in numpy, HiGHS on LP sub-cases and analytic SOCPs."""
import ctypes as C
import numpy as np
import pytest
import orc_py as O


def _rand_state(model, rng):
    nx, nu, npar = O.DIMS[model]
    if model == O.ROCKETQUAT:
        x = 0.3 * rng.normal(size=nx); x[0] = 1.0 + 0.2 * rng.random(); x[7] += 1.0; x[7:11] /= np.linalg.norm(x[7:11])
        u = rng.normal(size=nu) * 0.02; u[2] += 0.03
        par = np.array([0.3, 0, 0, -0.0115, 0.3, 0.3, 0.004, 0.001, -0.002, -0.0177])
    else:
        x = 0.3 * rng.normal(size=nx); u = np.array([0.1 * rng.normal(), 0.01 + 0.01 * rng.random()])
        par = np.array([1.0, 0.3, 0.0, -0.012, 0.0, -0.018])
    return x, u, par


@pytest.mark.parametrize("model", [O.ROCKETQUAT, O.ROCKET2D])
def test_jacobian_matches_central_differences(model):
    rng = np.random.default_rng(1)
    nx, nu, _ = O.DIMS[model]
    for _ in range(5):
        x, u, par = _rand_state(model, rng)
        A, B = O.jac(model, x, u, par)
        h = 1e-6
        for j in range(nx):
            e = np.zeros(nx); e[j] = h
            fd = (O.f(model, x + e, u, par) - O.f(model, x - e, u, par)) / (2 * h)
            assert np.allclose(A[:, j], fd, atol=2e-8), (j, A[:, j], fd)
        for j in range(nu):
            e = np.zeros(nu); e[j] = h
            fd = (O.f(model, x, u + e, par) - O.f(model, x, u - e, par)) / (2 * h)
            assert np.allclose(B[:, j], fd, atol=2e-8), (j, B[:, j], fd)


def test_rocketquat_literal_quirks():
    # w x w == 0 (no gyroscopic term); R(q) polynomial form without normalisation
    x = np.zeros(14); x[0] = 2.0; x[7:11] = [0.9, 0.1, -0.2, 0.3]; x[11:14] = [0.5, -0.4, 0.3]
    u = np.array([0.1, -0.2, 0.7, 0.05])
    par = np.array([0.3, 0.1, 0.2, -1.0, 2.0, 3.0, 4.0, 0.1, 0.2, -1.5])
    f = O.f(O.ROCKETQUAT, x, u, par)
    w, qx, qy, qz = x[7:11]
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - w * qz), 2 * (qx * qz + w * qy)],
                  [2 * (qx * qy + w * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - w * qx)],
                  [2 * (qx * qz - w * qy), 2 * (qy * qz + w * qx), 1 - 2 * (qx * qx + qy * qy)]])
    assert np.allclose(f[4:7], R @ u[:3] / x[0] + par[1:4], atol=1e-15)
    assert np.allclose(f[11:14], (np.cross(par[7:10], u[:3]) + [0, 0, u[3]]) / par[4:7], atol=1e-15)
    assert np.isclose(f[0], -par[0] * np.linalg.norm(u[:3]))


def test_rkf78_tableau_consistency():
    c = np.zeros(13); a = np.zeros((13, 13)); b = np.zeros(13)
    O.lib().orc_rkf78_tableau(c.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert np.allclose(a.sum(1), c, atol=1e-15)
    assert np.isclose(b.sum(), 1.0, atol=1e-15)
    # order conditions up to 3 for the 8th-order weights
    assert np.isclose(b @ c, 0.5) and np.isclose(b @ c ** 2, 1 / 3) and np.isclose(b @ (a @ c), 1 / 6)


@pytest.mark.parametrize("model,foh", [(O.ROCKETQUAT, True), (O.ROCKET2D, True)])
def test_discretisation_linearisation_identity(model, foh):
    """x_{k+1}^{nonlinear} == A x + B u + C u+ + s sigma + z at the linearisation point (exact identity)."""
    rng = np.random.default_rng(2)
    nx, nu, _ = O.DIMS[model]
    K = 6
    X = np.zeros((K, nx)); U = np.zeros((K, nu))
    for k in range(K):
        X[k], U[k], par = _rand_state(model, rng)
    sigma = 2.5
    dd = O.discretize(model, X, U, sigma, par, foh=True, free_time=True)
    for k in range(K - 1):
        xn = O.simulate(model, sigma / (K - 1), U[k], U[k + 1], par, X[k])
        lin = dd["A"][k] @ X[k] + dd["B"][k] @ U[k] + dd["C"][k] @ U[k + 1] + dd["s"][k] * sigma + dd["z"][k]
        assert np.allclose(lin, xn, atol=1e-11)
    # fixed final time instantiation: z absorbs f, no s
    ddf = O.discretize(model, X, U, sigma, par, foh=True, free_time=False)
    for k in range(K - 1):
        xn = O.simulate(model, sigma / (K - 1), U[k], U[k + 1], par, X[k])
        lin = ddf["A"][k] @ X[k] + ddf["B"][k] @ U[k] + ddf["C"][k] @ U[k + 1] + ddf["z"][k]
        assert np.allclose(lin, xn, atol=1e-11)
        assert np.allclose(ddf["A"][k], dd["A"][k], atol=1e-12)


def test_discretisation_sensitivities_by_finite_difference():
    """A_k = d x_{k+1}/d x_k, B_k, C_k, s_k by differencing the nonlinear simulation."""
    model = O.ROCKETQUAT
    rng = np.random.default_rng(3)
    x, u0, par = _rand_state(model, rng)
    _, u1, _ = _rand_state(model, rng)
    K, sigma = 5, 2.0
    X = np.tile(x, (K, 1)); U = np.tile(u0, (K, 1)); U[1] = u1
    dd = O.discretize(model, X, U, sigma, par)
    dt = sigma / (K - 1)
    h = 1e-6
    for j in range(14):
        e = np.zeros(14); e[j] = h
        fd = (O.simulate(model, dt, u0, u1, par, x + e) - O.simulate(model, dt, u0, u1, par, x - e)) / (2 * h)
        assert np.allclose(dd["A"][0][:, j], fd, atol=5e-9)
    for j in range(4):
        e = np.zeros(4); e[j] = h
        fd = (O.simulate(model, dt, u0 + e, u1, par, x) - O.simulate(model, dt, u0 - e, u1, par, x)) / (2 * h)
        assert np.allclose(dd["B"][0][:, j], fd, atol=5e-9)
        fd = (O.simulate(model, dt, u0, u1 + e, par, x) - O.simulate(model, dt, u0, u1 - e, par, x)) / (2 * h)
        assert np.allclose(dd["C"][0][:, j], fd, atol=5e-9)
    hs = 1e-5
    fd = (O.simulate(model, (sigma + hs) / (K - 1), u0, u1, par, x) - O.simulate(model, (sigma - hs) / (K - 1), u0, u1, par, x)) / (2 * hs)
    assert np.allclose(dd["s"][0], fd, atol=5e-9)


def test_conic_solver_against_highs_lp():
    from scipy.optimize import linprog
    rng = np.random.default_rng(4)
    n, p, m = 12, 4, 20
    A = rng.normal(size=(p, n)); x0 = rng.random(n); b = A @ x0
    G = np.vstack([-np.eye(n), rng.normal(size=(m - n, n))]); h = G @ x0 + rng.random(m) + 0.1
    c = rng.normal(size=n)
    ref = linprog(c, A_ub=G, b_ub=h, A_eq=A, b_eq=b, bounds=(None, None), method="highs")
    assert ref.status == 0
    Ai, Aj = np.nonzero(A); Gi, Gj = np.nonzero(G)
    r = O.conic_solve(c, b, h, m, [], (Ai, Aj, A[Ai, Aj]), (Gi, Gj, G[Gi, Gj]))
    assert r["status"] == 0
    assert abs(c @ r["x"] - ref.fun) < 1e-7 * max(1, abs(ref.fun))
    assert np.allclose(r["x"], ref.x, atol=1e-6)


def test_conic_solver_analytic_socp():
    # min c'x s.t. ||x|| <= 1  ->  x = -c/||c||
    c = np.array([1.0, -2.0, 0.5])
    G = np.vstack([np.zeros((1, 3)), -np.eye(3)]); h = np.array([1.0, 0, 0, 0])
    Gi, Gj = np.nonzero(G)
    r = O.conic_solve(c, np.zeros(0), h, 0, [4], (np.zeros(0, int), np.zeros(0, int), np.zeros(0)), (Gi, Gj, G[Gi, Gj]))
    assert r["status"] == 0
    assert np.allclose(r["x"], -c / np.linalg.norm(c), atol=1e-7)


def _setup_rq(K, inst=None):
    p, rpy = O.falcon9()
    if inst is not None:
        p = O.rq_perturb(p, rpy, 0x5C99, inst)
    cfg = O.sc_config(K=K)
    pn = O.RQParams.from_buffer_copy(p); O.lib().orc_rq_nondimensionalize(C.byref(pn))
    par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    X = np.zeros((K, 14)); U = np.zeros((K, 4)); t = C.c_double()
    O.lib().orc_rq_initial_trajectory(C.byref(pn), K, X.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), C.byref(t))
    return p, pn, cfg, par, X, U, t.value


def test_socp_dimensions_match_survey():
    # n=47K-25, p=16K+14, l=30K-26, cones 6K+1 of total dim 36K+3
    for K in (10, 50):
        p, pn, cfg, par, X, U, t = _setup_rq(K)
        dd = O.discretize(O.ROCKETQUAT, X, U, t, par)
        ex = O.export(O.ROCKETQUAT, pn, cfg, 50.0, X, U, t, dd)
        assert len(ex["c"]) == 47 * K - 25 and len(ex["b"]) == 16 * K + 14 and ex["l"] == 30 * K - 26
        assert len(ex["q"]) == 6 * K + 1 and ex["q"].sum() == 36 * K + 3
    p2 = O.rocket2d(); cfg = O.sc_config(K=30, model=O.ROCKET2D)
    pn2 = O.R2DParams.from_buffer_copy(p2); O.lib().orc_r2d_nondimensionalize(C.byref(pn2))
    par2 = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn2), par2.ctypes.data_as(C.c_void_p))
    X = np.zeros((30, 6)); U = np.zeros((30, 2)); t = C.c_double()
    O.lib().orc_r2d_initial_trajectory(C.byref(pn2), 30, X.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p), C.byref(t))
    dd = O.discretize(O.ROCKET2D, X, U, t.value, par2)
    ex = O.export(O.ROCKET2D, pn2, cfg, 1.0, X, U, t.value, dd)
    assert (len(ex["c"]), len(ex["b"]), ex["l"], len(ex["q"]), int(ex["q"].sum())) == (621, 187, 590, 61, 333)


def _kkt_certificate(ex, r):
    """independent numpy check of the optimality conditions of min c'x, Ax=b, Gx+s=h, s in K, z in K*"""
    x, y, s, z = r["x"], r["y"], r["s"], r["z"]
    A, G, c, b, h, l, q = ex["A"], ex["G"], ex["c"], ex["b"], ex["h"], ex["l"], ex["q"]
    pres = max(np.linalg.norm(A @ x - b) / max(1, np.linalg.norm(b)), np.linalg.norm(G @ x + s - h) / max(1, np.linalg.norm(h)))
    dres = np.linalg.norm(A.T @ y + G.T @ z + c) / max(1, np.linalg.norm(c))
    viol = max(0.0, -s[:l].min(), -z[:l].min())
    o = l
    for d in q:
        viol = max(viol, np.linalg.norm(s[o + 1:o + d]) - s[o], np.linalg.norm(z[o + 1:o + d]) - z[o]); o += d
    return pres, dres, s @ z, viol


@pytest.mark.parametrize("K,inst", [(12, None), (20, 3)])
def test_sc_subproblem_kkt_certificate(K, inst):
    p, pn, cfg, par, X, U, t = _setup_rq(K, inst)
    dd = O.discretize(O.ROCKETQUAT, X, U, t, par)
    ex = O.export(O.ROCKETQUAT, pn, cfg, 50.0, X, U, t, dd)
    Ai, Aj, Av, Gi, Gj, Gv = ex["raw"]
    r = O.conic_solve(ex["c"], ex["b"], ex["h"], ex["l"], ex["q"], (Ai, Aj, Av), (Gi, Gj, Gv))
    assert r["status"] == 0
    pres, dres, gap, viol = _kkt_certificate(ex, r)
    pcost = ex["c"] @ r["x"]
    assert pres < 1e-8 and dres < 1e-8 and viol <= 1e-12 and gap < 1e-7 * max(1, abs(pcost))
    # same solve through the SC entry point gives the same X,U,sigma
    sub = O.subproblem(O.ROCKETQUAT, pn, cfg, 50.0, X, U, t, dd)
    assert sub["status"] == 0
    assert np.allclose(sub["X"], r["x"][ex["iX"]], atol=1e-9) and np.isclose(sub["sigma"], r["x"][ex["isigma"]], atol=1e-9)
    # initial state pinned, final state rows pinned, virtual control = dynamics defect
    assert np.allclose(sub["X"][0], np.array(pn.x_init), atol=1e-9)
    fin = [1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13]
    assert np.allclose(sub["X"][-1, fin], np.array(pn.x_final)[fin], atol=1e-9)
    for k in range(K - 1):
        lin = dd["A"][k] @ sub["X"][k] + dd["B"][k] @ sub["U"][k] + dd["C"][k] @ sub["U"][k + 1] + dd["s"][k] * sub["sigma"] + dd["z"][k]
        assert np.allclose(sub["X"][k + 1] - lin, sub["nu"][k], atol=1e-8)
    assert abs(np.abs(sub["nu"]).sum() - sub["norm1_nu"]) < 1e-6
    assert np.allclose(np.linalg.norm(np.hstack([X - sub["X"], U - sub["U"]]), axis=1), sub["delta"], atol=1e-6)


def test_sc_loop_rocket2d_config0():
    """BASELINE.json configs[0]: Rocket2D SC K=30, single instance (the reference's own CPU-runnable case)."""
    p = O.rocket2d(); cfg = O.sc_config(K=30, model=O.ROCKET2D)
    r = O.sc_solve(O.ROCKET2D, p, cfg)
    assert r["converged"] and 2 <= r["iterations"] <= 15
    last = r["info"][-1]
    assert last.norm1_nu < cfg.nu_tol and last.sum_delta < cfg.delta_tol
    for i in r["info"]:
        assert i.ipm.status == 0 and i.ipm.pres < 1e-8 and i.ipm.dres < 1e-8
    # redimensionalised final trajectory lands: position/velocity match x_final
    assert np.allclose(r["X"][-1], np.array(p.x_final), atol=1e-6)
    assert np.allclose(r["X"][0], np.array(p.x_init), atol=1e-6)
    # the weight doubles only after norm1_nu < nu_tol (SCAlgorithm.cpp:112-115)
    ws = [i.weight_tr_used for i in r["info"]]
    for a, b, i in zip(ws, ws[1:], r["info"]):
        assert b == (2 * a if i.norm1_nu < cfg.nu_tol else a)


def test_sc_loop_rocketquat_runs_and_is_deterministic():
    p, rpy = O.falcon9()
    cfg = O.sc_config(K=20, max_iterations=4)
    r1 = O.sc_solve(O.ROCKETQUAT, p, cfg); r2 = O.sc_solve(O.ROCKETQUAT, p, cfg)
    assert r1["iterations"] == r2["iterations"] == 4
    assert np.array_equal(r1["X_all"], r2["X_all"])
    # initial guess quirks (rocketQuat.cpp:43-44,64): alpha2=k/K, U=(0,0,(Tmax-Tmin)/2,0) in nondimensional units
    pn = O.RQParams.from_buffer_copy(p); O.lib().orc_rq_nondimensionalize(C.byref(pn))
    assert np.allclose(r1["U_all"][0], [0, 0, (pn.T_max - pn.T_min) / 2, 0])
    assert np.isclose(r1["X_all"][0][-1, 0], (1 / 20) * pn.x_init[0] + (19 / 20) * pn.x_final[0])
    assert np.isclose(r1["X"][0, 0], 24000.0) and np.allclose(r1["X"][0, 1:4], [200, 200, 800])


def test_perturbation_recipe_is_reproducible():
    p, rpy = O.falcon9()
    a = O.rq_perturb(p, rpy, 0x5C99, 17); b = O.rq_perturb(p, rpy, 0x5C99, 17); c = O.rq_perturb(p, rpy, 0x5C99, 18)
    assert list(a.x_init) == list(b.x_init) and list(a.x_init) != list(c.x_init)
    assert a.x_init[0] == p.x_init[0] and a.x_init[3] == p.x_init[3]          # mass and r_z unchanged
    assert abs(a.x_init[1]) <= abs(p.x_init[1]) and abs(a.x_init[6]) <= 1.2 * abs(p.x_init[6])
    assert np.isclose(np.linalg.norm(np.array(a.x_init)[7:11]), 1.0)


@pytest.mark.parametrize("name,model,K,max_it,inst", [("rocket2d_K30", O.ROCKET2D, 30, 15, None), ("rocketquat_K20_nominal", O.ROCKETQUAT, 20, 5, None),
                                                      ("rocketquat_K50_inst7", O.ROCKETQUAT, 50, 4, 7)])
def test_oracle_reproduces_committed_fixtures(name, model, K, max_it, inst):
    """tests/golden/*.npz are ORACLE-generated (the reference has no golden vectors): they pin the oracle against drift"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    if model == O.ROCKET2D:
        p = O.rocket2d()
    else:
        p, rpy = O.falcon9()
        if inst is not None:
            p = O.rq_perturb(p, rpy, 0x5C99, inst)
    r = O.sc_solve(model, p, O.sc_config(K=K, model=model, max_iterations=max_it))
    assert r["iterations"] == int(g["iterations"]) and int(r["converged"]) == int(g["converged"])
    assert np.allclose(r["X_all"], g["X_all"], atol=1e-9) and np.allclose(r["U_all"], g["U_all"], atol=1e-9)
    assert np.allclose(r["t_all"], g["t_all"], atol=1e-9)


# ---- SCvx variant of the oracle (restated, not yet built on the GPU) --------------------------
def test_scvx_restatement_rocket2d_runs_and_is_certified():
    """literal SCvxAlgorithm loop (buildSCvxProblem + ratio test against the simulated nonlinear cost) on the reference's own
    SCvx.info for Rocket2D: every sub-problem is solved to the ECOS tolerances, the bookkeeping of iterate() holds"""
    p = O.rocket2d()
    cfg = O.scvx_config(K=30, model=O.ROCKET2D, max_iterations=6)
    r = O.scvx_solve(O.ROCKET2D, p, cfg)
    assert r["iterations"] >= 3
    last = None
    for i, inf in enumerate(r["info"]):
        assert inf.ipm.status == 0 and inf.ipm.pres < 1e-8 and inf.ipm.dres < 1e-8
        assert inf.norm1_nu >= -1e-9 and inf.nonlinear_cost > 0 and inf.solves >= 1
        if i == 0:
            assert inf.solves == 1 and inf.rho == 0.0          # the first iterate is accepted unconditionally (:110-114)
        elif not (r["converged"] and i == len(r["info"]) - 1):
            assert inf.rho >= cfg.rho_0                        # the last solve of an iteration is an accepted step
            assert abs(inf.rho - inf.actual_change / inf.predicted_change) < 1e-12
        last = inf
    # the nonlinear cost reported for an accepted iterate is the simulated defect of that iterate (getNonlinearCost, :262-278)
    par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(p), par.ctypes.data_as(C.c_void_p))
    J = O.scvx_nonlinear_cost(O.ROCKET2D, r["X_all"][-1], r["U_all"][-1], p.final_time, par)
    assert abs(J - last.nonlinear_cost) < 1e-9 * max(1.0, J)


def test_scvx_subproblem_optimum_is_not_unique_in_the_states():
    """Finding that decides how SCvx parity can be stated: the SCvx sub-problem minimises only w_vc*|nu|_1, so its optimal VALUE and
    inputs are determined but the intermediate states are not (flat directions of the 1-norm).  Two solves that differ only in the
    solver's regularisation agree on norm1_nu to 1e-9 relative and on U to 1e-6, and differ in X by O(1): iterate-level parity of X
    against another interior-point code (the reference's ECOS) is not defined for SCvx; parity has to be stated on U, norm1_nu, the
    nonlinear cost and the accept/reject decisions."""
    p = O.rocket2d()
    out = []
    for reg in (1e-8, 2e-7):
        C.c_double.in_dll(O.lib(), "orc_scvx_static_reg").value = reg
        out.append(O.scvx_solve(O.ROCKET2D, p, O.scvx_config(K=30, model=O.ROCKET2D, max_iterations=1)))
    C.c_double.in_dll(O.lib(), "orc_scvx_static_reg").value = 2e-7
    a, b = out
    assert a["info"][0].ipm.status == 0 and b["info"][0].ipm.status == 0
    assert abs(a["info"][0].norm1_nu - b["info"][0].norm1_nu) < 1e-8 * a["info"][0].norm1_nu
    assert np.abs(a["U_all"][1] - b["U_all"][1]).max() < 1e-5
    assert np.abs(a["X_all"][1] - b["X_all"][1]).max() > 1e-2


def test_lqr_restatement_against_scipy_care():
    """solveSchurIterative + FullPivLU restatement (LQR.cpp:7-109) reproduces the stabilising CARE solution on well-conditioned systems"""
    import scipy.linalg as sl
    rng = np.random.default_rng(0)
    p = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(C.c_void_p)
    for nx, nu in ((6, 2), (14, 4)):
        A = rng.standard_normal((nx, nx)); B = rng.standard_normal((nx, nu)); q = 0.5 + rng.random(nx); r = 1.0 + rng.random(nu)
        K = np.zeros((nu, nx))
        assert O.lib().orc_lqr_gain(nx, nu, p(q), p(r), p(A), p(B), K.ctypes.data_as(C.c_void_p)) == 1
        P = sl.solve_continuous_are(A, B, np.diag(q), np.diag(r))
        Kref = np.linalg.solve(np.diag(r), B.T @ P)
        assert np.abs(K - Kref).max() < 1e-10 * np.abs(Kref).max()
        assert np.linalg.eigvals(A - B @ K).real.max() < 0          # stabilising


def test_scvx_oracle_reproduces_committed_fixture():
    """tests/golden/rocketquat_scvx_K30_nominal.npz (oracle-generated): the SCvx restatement on the reference's RocketQuat SCvx.info --
    16 outer iterations, converged; pins the oracle (equilibration, regularisation, ratio-test loop) against drift"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rocketquat_scvx_K30_nominal.npz"))
    p, _ = O.falcon9()
    r = O.scvx_solve(O.ROCKETQUAT, p, O.scvx_config(K=30, model=O.ROCKETQUAT))
    assert r["iterations"] == int(g["iterations"]) == 16 and int(r["converged"]) == int(g["converged"]) == 1
    assert np.allclose(r["X_all"], g["X_all"], atol=1e-9) and np.allclose(r["U_all"], g["U_all"], atol=1e-9)
    assert np.array_equal(np.array([i.solves for i in r["info"]]), g["info"][:, 4].astype(int))
    J = g["info"][:, 1]
    assert J[-1] < 0.15 * J[0]                                 # the nonlinear defect falls by an order of magnitude
