"""CPU tests of the host logic and of the kernel SOURCE compiled for the host (LANES == 1, tests/_hostsim):
the C-ABI library loads and exports every declared symbol, the INFO loader, the dual-number check of the device
Jacobians, kernel-source parity against the oracle, and the multi-rank plumbing over gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import hostsim as H
import orc_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S():
    from scpp_b200 import build
    build.build()          # nvcc cross-compiles sm_100a without a GPU
    import scpp_b200
    return scpp_b200


def test_cabi_exports_every_declared_symbol(S):
    hdr = open(os.path.join(ROOT, "include", "scpp_b200.h")).read()
    names = sorted(set(re.findall(r"\b(scpp_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 18
    lib = S.lib()
    for n in names:
        assert hasattr(lib, n), f"libscpp_b200.so does not export {n}"
    assert lib.scpp_b200_version() >= 100
    nx, nu, npar = S.model_dims(S.ROCKETQUAT)
    assert (nx, nu, npar) == (14, 4, 10) and S.model_dims(S.ROCKET2D) == (6, 2, 6)


def test_no_cpu_execution_path(S):
    """without a CUDA device every compute entry point must fail loudly (never fall back)"""
    if S.device_count() > 0:
        pytest.skip("a GPU is present")
    model, params, xi, xf, cfg = S.load_model("RocketQuat")
    with pytest.raises(S.ScppError, match="no CUDA device"):
        S.SCAlgorithm(model, params, cfg, 2)
    with pytest.raises(S.ScppError, match="no CUDA device"):
        S.discretize(model, np.zeros((3, 14)), np.zeros((3, 4)), 1.0, np.zeros(10))
    # the package never imports the oracle or the host simulation
    src = open(os.path.join(ROOT, "scpp_b200", "__init__.py")).read() + open(os.path.join(ROOT, "scpp_b200", "csrc", "engine.cu")).read()
    assert "orc_" not in src and "hostsim" not in src and "liborc" not in src


def test_info_loader_matches_reference_parameter_sets(S):
    model, p, xi, xf, cfg = S.load_model("RocketQuat")
    po, rpy = O.falcon9()
    assert np.array_equal(xi, np.array(po.x_init)) and np.array_equal(xf, np.array(po.x_final))
    for f in ("alpha_m", "T_min", "T_max", "t_max", "gimbal_max", "theta_max", "gamma_gs", "w_B_max", "final_time"):
        assert getattr(p, f) == getattr(po, f), f
    assert (cfg.K, cfg.max_iterations, cfg.weight_trust_region_trajectory, cfg.weight_virtual_control) == (15, 15, 50.0, 1000.0)
    assert cfg.nu_tol == 1e-5 and cfg.delta_tol == 1e-3 and cfg.free_final_time == 1 and cfg.nondimensionalize == 1
    _, ps, xis, _, _ = S.load_model("RocketQuatStarship")
    pso, _ = O.starship()
    assert np.array_equal(xis, np.array(pso.x_init)) and ps.T_max == pso.T_max and ps.gamma_gs == pso.gamma_gs
    _, p2, xi2, xf2, cfg2 = S.load_model("Rocket2D")
    p2o = O.rocket2d()
    assert np.array_equal(xi2, np.array(p2o.x_init)) and np.array_equal(xf2, np.array(p2o.x_final)) and p2.m == p2o.m and cfg2.K == 25


def test_info_loader_error_semantics(S, tmp_path):
    """parameterServer.hpp:66-77,95-103: missing scalar, missing / redundant vector entries raise"""
    good = open(os.path.join(ROOT, "configs", "RocketQuat", "model.info")).read()
    def load(txt):
        f = tmp_path / "m.info"; f.write_text(txt)
        return S.load_model_info(str(f), S.ROCKETQUAT)
    load(good)
    with pytest.raises(S.ScppError, match="Failed to load scalar"):
        load(good.replace("I_sp    275.", ""))
    with pytest.raises(S.ScppError, match="Missing entries"):
        load(good.replace("g_I   { (0) 0.0  (1) 0.0  (2) -9.81 }", "g_I   { (0) 0.0  (1) 0.0 }"))
    with pytest.raises(S.ScppError, match="Redundant entries"):
        load(good.replace("g_I   { (0) 0.0  (1) 0.0  (2) -9.81 }", "g_I   { (0) 0.0  (1) 0.0  (2) -9.81 (3) 1. }"))
    with pytest.raises(S.ScppError):
        S.load_model_info(str(tmp_path / "nope.info"), S.ROCKETQUAT)
    # scaling and comments
    p, xi, xf = load(good.replace("r_init    { (0) 200.  (1) 200.  (2) 800. }", "r_init { scaling 100. ; cm\n (0) 2. \n (1) 2. \n (2) 8. }"))
    assert np.allclose(xi[1:4], [200, 200, 800])


def test_perturbation_recipe_matches_oracle(S):
    model, p, xi, xf, cfg = S.load_model("RocketQuat")
    po, rpy = O.falcon9()
    b = S.perturbed_initial_states(xi, rpy, 5, first=3)
    for i in range(5):
        assert np.array_equal(b[i], np.array(O.rq_perturb(po, rpy, 0x5C99, 3 + i).x_init))


@pytest.mark.parametrize("model", [0, 1])
def test_device_jacobian_equals_dual_number_jacobian(model):
    """hand-derived sparse Jacobian used by K1 == forward-mode AD of the generic-scalar flow map == oracle Jacobian"""
    rng = np.random.default_rng(7)
    nx, nu = H.DIMS[model]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for _ in range(10):
        if model == 0:
            x = 0.5 * rng.normal(size=nx); x[0] = 1 + rng.random(); x[7] += 1
            u = 0.02 * rng.normal(size=nu); u[2] += 0.03
            par = np.array([0.3, 0.01, -0.02, -0.0115, 0.3, 0.25, 0.004, 0.001, -0.002, -0.0177])
        else:
            x = rng.normal(size=nx); u = np.array([0.2 * rng.normal(), 0.02 * rng.random() + 0.01]); par = np.array([1.0, 0.3, 0.001, -0.012, 0.002, -0.018])
        f = np.zeros(nx); Aad = np.zeros((nx, nx)); Bad = np.zeros((nx, nu)); Al = np.zeros((nx, nx)); Bl = np.zeros((nx, nu))
        H.lib().hs_jacobians(model, p(x), p(u), p(par), p(f), p(Aad), p(Bad), p(Al), p(Bl))
        assert not np.isnan(f).any()
        assert np.allclose(Al, Aad, atol=1e-13) and np.allclose(Bl, Bad, atol=1e-13)
        Ao, Bo = O.jac(model, x, u, par)
        assert np.allclose(Ao, Aad, atol=1e-13) and np.allclose(Bo, Bad, atol=1e-13) and np.allclose(O.f(model, x, u, par), f, atol=1e-14)


def test_kernel_source_discretisation_vs_oracle():
    p, rpy = O.falcon9()
    r = O.sc_solve(O.ROCKETQUAT, p, O.sc_config(K=30, max_iterations=2))
    pn = O.RQParams.from_buffer_copy(p); O.lib().orc_rq_nondimensionalize(C.byref(pn))
    par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    X, U, t = r["X_all"][2], r["U_all"][2], r["t_all"][2]
    ref = O.discretize(O.ROCKETQUAT, X, U, t, par)
    errs = []
    for nsub in (5, 10, 20):
        got = H.discretize(0, X, U, t, par, nsub)
        errs.append(max(np.abs(got[k] - ref[k]).max() / max(1.0, np.abs(ref[k]).max()) for k in ("A", "B", "C", "s", "z")))
    assert errs[2] < 2e-10                       # RK4 x 20
    rich = H.discretize(0, X, U, t, par, -5)     # the shipped integrator: RK4 x 5 and x 10, Richardson-extrapolated (60 instead of 80 evaluations)
    assert max(np.abs(rich[k] - ref[k]).max() / max(1.0, np.abs(ref[k]).max()) for k in ("A", "B", "C", "s", "z")) < 2e-10
    assert 8 < errs[0] / errs[1] < 24 and 8 < errs[1] / errs[2] < 24   # 4th-order convergence towards the RKF78 result
    # the shared-linearisation mapping (cfg.jacobian = 2, discretize_shared.cuh): producer one RK4 step ahead, two stash buffers, the
    # Richardson pair as one 15-step schedule -- same arithmetic as the column kernel with the hand-derived Jacobian
    for nsub in (-5, 20, 3):
        shared = H.discretize(0, X, U, t, par, nsub, jacobian=2); column = H.discretize(0, X, U, t, par, nsub, jacobian=0)
        for k in ("A", "B", "C", "s", "z"):
            assert np.abs(shared[k] - column[k]).max() <= 1e-13 * max(1.0, np.abs(column[k]).max()), (nsub, k)
    rng = np.random.default_rng(3)               # Rocket2D (12 columns, dense B record)
    X2 = rng.normal(size=(9, 6)) * 0.3; U2 = np.column_stack([rng.normal(size=9) * 0.1, 1.0 + rng.random(9)]); par2 = np.array([1.0, 0.02, 0.0, -0.1, 0.0, -0.05])
    shared = H.discretize(1, X2, U2, 3.0, par2, -5, jacobian=2); column = H.discretize(1, X2, U2, 3.0, par2, -5, jacobian=0)
    for k in ("A", "B", "C", "s", "z"):
        assert np.abs(shared[k] - column[k]).max() <= 1e-13 * max(1.0, np.abs(column[k]).max()), k


@pytest.mark.parametrize("warm,ipm_slice", [(0.0, 1), (0.995, 1), (0.0, 0), (0.995, 3), (0.0, -1), (0.995, -1)])
@pytest.mark.parametrize("name,model,K,max_it", [("Rocket2D", 1, 30, 15), ("RocketQuat", 0, 20, 5)])
def test_kernel_source_sc_loop_vs_oracle(name, model, K, max_it, warm, ipm_slice):
    """the K2/K3 source (structured IPM, one 'warp' of 1 lane) reproduces the literal ECOS-form oracle iterate by iterate"""
    if model == 0:
        p, _ = O.falcon9()
    else:
        p = O.rocket2d()
    ocfg = O.sc_config(K=K, model=model, max_iterations=max_it)
    ro = O.sc_solve(model, p, ocfg)
    P, xi, xf = H.params_from_oracle(model, p)
    rh = H.sc_solve(model, P, H.sc_config(ocfg, tol=1e-8, warm=warm, ipm_slice=ipm_slice), xi, xf)
    n = ro["iterations"]
    assert n > 0 and rh["iters"][0] == n and bool(rh["converged"][0] == 1) == ro["converged"]
    for it in range(n + 1):
        assert np.abs(rh["X_all"][0, it] - ro["X_all"][it]).max() < 1e-5
        assert np.abs(rh["U_all"][0, it] - ro["U_all"][it]).max() < 1e-4
    for it in range(n):
        assert rh["info"][0, it, 4] == ro["info"][it].weight_tr_used
        assert int(rh["info"][0, it, 6]) in (0, 3)
    assert np.allclose(rh["X"][0], ro["X"], rtol=1e-6, atol=1e-4 * np.abs(ro["X"]).max())


@pytest.mark.parametrize("name,model,K,max_it", [("Rocket2D", 1, 30, 15), ("RocketQuat", 0, 20, 5)])
def test_kernel_source_zero_order_hold_vs_oracle(name, model, K, max_it):
    """interpolate_input = false (discretizationImplementation.hpp:41-50,96-101; SCProblem.cpp:49-56,114-121; the model's final-input
    constraints on column K - 2): the engine's K-column layout with a pinned placeholder column against the oracle's literal K - 1 columns,
    iterate by iterate; the discretisation alone against the oracle's zero-order-hold RKF78"""
    p = O.falcon9()[0] if model == 0 else O.rocket2d()
    ocfg = O.sc_config(K=K, model=model, max_iterations=max_it)
    foh = O.sc_solve(model, p, ocfg)
    ocfg.interpolate_input = 0
    ro = O.sc_solve(model, p, ocfg)
    if model == 0:
        pn = O.RQParams.from_buffer_copy(p); O.lib().orc_rq_nondimensionalize(C.byref(pn))
        par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    else:
        pn = O.R2DParams.from_buffer_copy(p); O.lib().orc_r2d_nondimensionalize(C.byref(pn))
        par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(pn), par.ctypes.data_as(C.c_void_p))
    Xm, Um, tm = ro["X_all"][2], ro["U_all"][2], ro["t_all"][2]
    ref = O.discretize(model, Xm, Um, tm, par, foh=False)
    for jac in (0, 1, 2):          # all three K1 paths hold the input: B is the whole input matrix, C is exactly zero
        got = H.discretize(model, Xm, Um, tm, par, -5, jacobian=jac, zoh=True)
        for key in ("A", "B", "s", "z"):
            assert np.abs(got[key] - ref[key]).max() <= 2e-10 * max(1.0, np.abs(ref[key]).max()), (jac, key)
        assert not got["C"].any()
    k = K // 2                   # linear model == nonlinear propagation with the input held, at the linearisation point
    lin = ref["A"][k] @ Xm[k] + ref["B"][k] @ Um[k] + ref["s"][k] * tm + ref["z"][k]
    assert np.allclose(lin, O.simulate(model, tm / (K - 1), Um[k], Um[k], par, Xm[k]), atol=1e-9)
    P, xi, xf = H.params_from_oracle(model, p)
    rh = H.sc_solve(model, P, H.sc_config(ocfg, tol=1e-8, warm=0.0, ipm_slice=1), xi, xf)
    n = ro["iterations"]
    assert n > 0 and rh["iters"][0] == n and bool(rh["converged"][0] == 1) == ro["converged"]
    assert np.abs(ro["X_all"][n] - foh["X_all"][min(n, foh["iterations"])]).max() > 1e-4          # the two holds are different problems
    for it in range(1, n + 1):
        assert np.abs(rh["X_all"][0, it] - ro["X_all"][it]).max() < 1e-5, it
        assert np.abs(rh["U_all"][0, it, :K - 1] - ro["U_all"][it][:K - 1]).max() < 1e-4, it     # column K - 1: placeholder here, absent in the reference
    for it in range(n):
        assert rh["info"][0, it, 4] == ro["info"][it].weight_tr_used
        assert int(rh["info"][0, it, 6]) in (0, 3)
        assert abs(rh["info"][0, it, 0] - ro["info"][it].norm1_nu) < 1e-6 and abs(rh["info"][0, it, 1] - ro["info"][it].sum_delta) < 1e-5 * max(1.0, ro["info"][it].sum_delta)


def test_sliced_solver_is_bit_identical_to_unsliced():
    """parking the interior-point state between launches (cfg.ipm_slice) must not change a single bit of the iterates"""
    p, _ = O.falcon9()
    ocfg = O.sc_config(K=20, model=0, max_iterations=4)
    P, xi, xf = H.params_from_oracle(0, p)
    ref = H.sc_solve(0, P, H.sc_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=0), xi, xf)
    for sl in (1, 2, 5):
        r = H.sc_solve(0, P, H.sc_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=sl), xi, xf)
        assert np.array_equal(r["X_all"], ref["X_all"]) and np.array_equal(r["U_all"], ref["U_all"]) and np.array_equal(r["info"], ref["info"])


def test_shard_range_partitions_the_batch():
    from scpp_b200.sharding import shard_range
    for n, w in ((8192, 8), (1000, 3), (5, 8), (1, 1)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch.distributed as dist
import scpp_b200 as S, hostsim as H, orc_py as O
from scpp_b200.sharding import shard_range, broadcast_unique_id, global_active, reduce_timing
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
uid = broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
assert uid == bytes(range(128))
N, K, max_it = 5, 12, 4
lo, hi = shard_range(N, world, rank)
model, p, xi, xf, cfg = S.load_model("RocketQuat", K=K, max_iterations=max_it)
xis = S.perturbed_initial_states(xi, np.deg2rad([-20., 20., 0.]), hi - lo, first=lo)
# per-rank engine stand-in on CPU: the kernel source compiled for the host (test-only)
po, _ = O.falcon9()
P, _, _ = H.params_from_oracle(0, po)
r = H.sc_solve(0, P, H.sc_config(O.sc_config(K=K, max_iterations=max_it), tol=1e-8), xis, xf)
pad = max(b - a for a, b in (shard_range(N, world, q) for q in range(world)))
n_act, allf = global_active(dist, (r["converged"] != 0).astype(np.uint8), pad)
t, c = reduce_timing(dist, [1.0 + rank], [int(r["iters"].sum())])
full = S.perturbed_initial_states(xi, np.deg2rad([-20., 20., 0.]), N)
assert np.array_equal(full[lo:hi], xis)
out = dict(rank=rank, n_act=n_act, tmax=float(t[0]), iters=float(c[0]), local_iters=int(r["iters"].sum()), shape=list(allf.shape))
import json
with open(os.path.join(sys.argv[2], f"rank{rank}.json"), "w") as fh:      # one file per rank: stdout lines of concurrent ranks can interleave
    json.dump(out, fh)
dist.destroy_process_group()
'''


def test_two_rank_plumbing_over_gloo(tmp_path):
    """world_size 2 on CPU: shard assignment, unique-id broadcast, flag all-gather and timing reduction agree on both ranks"""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    import socket
    with socket.socket() as sk:          # a free port, so parallel test sessions do not collide
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT, str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    import json
    outs = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    a, b = sorted(outs, key=lambda o: o["rank"])
    assert a["n_act"] == b["n_act"] and a["tmax"] == b["tmax"] == 2.0
    assert a["iters"] == b["iters"] == a["local_iters"] + b["local_iters"]
    assert a["shape"] == [2, 3]


def test_split_pipeline_two_parts_vs_monolithic_source():
    """K = 50: the stage-parallel passes of the split pipeline run as two 32-stage parts with partial sums; same iterates as the
    one-warp driver up to rounding, same iteration counts"""
    p, _ = O.falcon9()
    ocfg = O.sc_config(K=50, model=0, max_iterations=3)
    P, xi, xf = H.params_from_oracle(0, p)
    ref = H.sc_solve(0, P, H.sc_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=1), xi, xf)
    r = H.sc_solve(0, P, H.sc_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=-1), xi, xf)
    assert np.array_equal(r["iters"], ref["iters"]) and np.array_equal(r["info"][:, :, 5], ref["info"][:, :, 5])
    assert np.abs(r["X_all"] - ref["X_all"]).max() < 1e-7 and np.abs(r["U_all"] - ref["U_all"]).max() < 1e-7


def test_k4_simulate_source_is_the_oracle_integrator():
    """K4 body (RKF78, 20 steps, first-order-hold input) vs the oracle's restatement of scpp::simulate: same arithmetic"""
    p, _ = O.falcon9()
    par = np.zeros(10); O.lib().orc_rq_model_par(C.byref(p), par.ctypes.data_as(C.c_void_p))
    x = np.array(p.x_init); u0 = np.array([1e4, -2e4, 3e5, 0.]); u1 = np.array([-1e4, 1e4, 3.5e5, 0.])
    a = H.simulate(0, 0.05, u0, u1, par, x); b = O.simulate(0, 0.05, u0, u1, par, x)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max() and np.abs(a - x).max() > 1e-3
    p2 = O.rocket2d()
    par2 = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(p2), par2.ctypes.data_as(C.c_void_p))
    x2 = np.array(p2.x_init)
    a2 = H.simulate(1, 0.05, [0.01, 3e5], [0.02, 3.2e5], par2, x2); b2 = O.simulate(1, 0.05, [0.01, 3e5], [0.02, 3.2e5], par2, x2)
    assert np.abs(a2 - b2).max() <= 1e-12 * np.abs(b2).max()


@pytest.mark.parametrize("name,model,K,steps", [("Rocket2D", 1, 30, 4), ("RocketQuat", 0, 20, 2)])
def test_closed_loop_source_vs_oracle(name, model, K, steps):
    """SC_sim closed loop (scpp/src/SC_sim.cpp:41-65): cold solve, then warm-started solves from the simulated state; the kernel source
    (warm-start re-scaling, interpolated input, K4, carried trust-region weight) against the oracle's literal loop"""
    p = O.falcon9()[0] if model == 0 else O.rocket2d()
    ocfg = O.sc_config(K=K, model=model, max_iterations=15)
    ro = O.sc_sim(model, p, ocfg, 0.05, steps)
    assert ro["steps"] == steps
    P, xi, xf = H.params_from_oracle(model, p)
    rh = H.sc_sim(model, P, H.sc_config(ocfg, tol=1e-8, warm=0.0, ipm_slice=1, history=False), xi, xf, 0.05, steps)
    assert np.array_equal(rh["iters"][:, 0], ro["iters"][:steps])
    sx, su = np.abs(ro["X_sim"]).max(), np.abs(ro["U_sim"]).max()
    assert np.abs(rh["X_sim"][:, 0] - ro["X_sim"]).max() < 1e-7 * sx
    assert np.abs(rh["U_sim"][:, 0] - ro["U_sim"]).max() < 1e-5 * su


def test_k5_lqr_gain_source_vs_oracle():
    """K5 body (Gauss-Jordan inverse inside the sign iteration, complete-pivoting solve) vs the oracle's restatement of LQR.cpp along a
    solved Rocket2D trajectory with the reference's LQR.info weights.  The Hamiltonians of this model are ill-conditioned (double
    integrators): two correct implementations of the same iteration agree to ~1e-6 relative, which is the bar here."""
    p2 = O.rocket2d()
    rs = O.sc_solve(1, p2, O.sc_config(K=30, model=1, max_iterations=6))
    par = np.zeros(6); O.lib().orc_r2d_model_par(C.byref(p2), par.ctypes.data_as(C.c_void_p))
    X, U = rs["X"], rs["U"]
    q, r = np.ones(6), np.array([2.0, 2.0])
    pp = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(C.c_void_p)
    G = np.zeros((30, 2, 6)); ok = np.zeros(30, np.int32); G2 = np.zeros((30, 2, 6)); ok2 = np.zeros(30, np.int32)
    O.lib().orc_lqr_tracker_gains(1, 30, pp(X), pp(U), pp(par), pp(q), pp(r), G.ctypes.data_as(C.c_void_p), ok.ctypes.data_as(C.c_void_p))
    H.lib().hs_lqr_gains(1, 30, pp(X), pp(U), pp(par), pp(q), pp(r), G2.ctypes.data_as(C.c_void_p), ok2.ctypes.data_as(C.c_void_p))
    assert ok.all() and ok2.all()
    rel = np.abs(G - G2).max(axis=(1, 2)) / np.abs(G).max(axis=(1, 2))
    assert rel.max() < 1e-4
    # the gains stabilise the linearised dynamics at every node
    for k in (0, 10, 29):
        A, B = O.jac(1, X[k], U[k], par)
        assert np.linalg.eigvals(A - B @ G2[k]).real.max() < 0


def _scvx_common_prefix(ro, info_h, n_h):
    """number of leading outer iterations in which both runs took the same decisions (solves per iteration, trust region used)"""
    m = 0
    for it in range(min(abs(ro["iterations"]), n_h)):
        a = ro["info"][it]
        if a.solves != int(info_h[it, 4]) or abs(a.trust_region_used - info_h[it, 3]) > 1e-12 * a.trust_region_used:
            break
        m += 1
    return m


@pytest.mark.parametrize("warm", [0.0, 0.995])
def test_scvx_source_vs_oracle_nominal(warm):
    """SCvx variant (SCvxProblem.cpp:6-71, SCvxAlgorithm.cpp:61-164) in the kernel source -- trust-region cone on the input rows with a
    pinned radius, pinned sigma, K4 cost, ratio test -- against the oracle's literal loop on the reference's RocketQuat SCvx.info:
    same decisions, same radii, converged in the same iteration; iterates to 5e-5 / 5e-6 (the sub-problem minimises only |nu|_1, its
    states are weakly determined, so the SC bar of 1e-5 is not reachable between two interior-point codes)"""
    p = O.falcon9()[0]
    ocfg = O.scvx_config(K=30, model=0)
    ro = O.scvx_solve(0, p, ocfg)
    P, xi, xf = H.params_from_oracle(0, p)
    rh = H.sc_solve(0, P, H.scvx_config(ocfg, tol=1e-8, warm=warm), xi, xf)
    n = ro["iterations"]
    assert n > 5 and ro["converged"] and rh["iters"][0] == n and rh["converged"][0] == 1
    for it in range(n + 1):
        assert np.abs(rh["X_all"][0, it] - ro["X_all"][it]).max() < 5e-5 and np.abs(rh["U_all"][0, it] - ro["U_all"][it]).max() < 5e-6
    for it in range(n):
        a, h = ro["info"][it], rh["info"][0, it]
        assert abs(a.norm1_nu - h[0]) < 1e-5 * a.norm1_nu and abs(a.nonlinear_cost - h[1]) < 2e-4 * a.nonlinear_cost
        assert abs(a.rho - h[2]) < 2e-3 * max(1.0, abs(a.rho))
    assert abs(ro["info"][n - 1].trust_region_used - rh["info"][0, n - 1, 3]) < 1e-12


def test_scvx_source_vs_oracle_perturbed_prefix():
    """The reference's SCvx loop is not reproducible decision by decision: after a rejected step last_nonlinear_cost is overwritten, and
    while the trust region is inactive the re-solve returns the same candidate, so rho = (rounding noise) / predicted and its sign
    decides accept or reject (SCvxAlgorithm.cpp:116-139).  Solves per iteration and radii differ between implementations; the accepted
    iterates agree while the radius is inactive.  Parity is asserted on the leading iterates and on the outcome."""
    p, rpy = O.falcon9()
    pp = O.rq_perturb(p, rpy, 0x5C99, 0)
    ocfg = O.scvx_config(K=30, model=0)
    ro = O.scvx_solve(0, pp, ocfg)
    P, xi, xf = H.params_from_oracle(0, pp)
    rh = H.sc_solve(0, P, H.scvx_config(ocfg, tol=1e-8, warm=0.995), xi, xf)
    n = int(rh["iters"][0])
    assert ro["converged"] and rh["converged"][0] == 1 and abs(n - ro["iterations"]) <= 8
    m = 0
    for it in range(1, min(ro["iterations"], n) + 1):
        if np.abs(rh["X_all"][0, it] - ro["X_all"][it]).max() < 5e-5 and np.abs(rh["U_all"][0, it] - ro["U_all"][it]).max() < 5e-6:
            m = it
        else:
            break
    assert m >= 3
    Jo, Jh = ro["info"][-1].nonlinear_cost, rh["info"][0, n - 1, 1]
    assert abs(Jo - Jh) < 0.05 * Jo


def test_scvx_info_loader(S, tmp_path):
    """scpp_b200_load_scvx_info == SCvxAlgorithm::loadParameters (SCvxAlgorithm.cpp:23-44) on the shipped SCvx.info values; a missing
    key raises like the reference's ParameterServer"""
    for name, K, thr, nd, mit in (("RocketQuat", 30, 1e-3, 1, 30), ("Rocket2D", 30, 1e-2, 0, 20)):
        model, params, xi, xf, cfg = S.load_model(name, algorithm="SCvx")
        assert cfg.algorithm == 1 and cfg.K == K and cfg.max_iterations == mit and cfg.nondimensionalize == nd and cfg.interpolate_input == 1
        assert (cfg.scvx_rho_0, cfg.scvx_rho_1, cfg.scvx_rho_2, cfg.scvx_alpha, cfg.scvx_beta) == (0.0, 0.25, 0.9, 2.0, 3.2)
        assert cfg.scvx_change_threshold == thr and cfg.weight_virtual_control == 1e3 and cfg.scvx_trust_region == 5.0
    bad = tmp_path / "SCvx.info"
    bad.write_text("K 30\nnondimensionalize true\nmax_iterations 5\n")
    with pytest.raises(S.ScppError):
        S.load_scvx_info(str(bad), S.ROCKETQUAT)
    cfg = S.default_config(S.ROCKETQUAT)
    assert cfg.algorithm == 0


def test_scvx_in_the_split_pipeline_source():
    """SCvx mode of every step of the split pipeline (K = 30: one 32-stage part, so the sums are the monolithic kernel's): identical iterates"""
    p = O.falcon9()[0]
    ocfg = O.scvx_config(K=30, model=0, max_iterations=8)
    P, xi, xf = H.params_from_oracle(0, p)
    a = H.sc_solve(0, P, H.scvx_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=1), xi, xf)
    b = H.sc_solve(0, P, H.scvx_config(ocfg, tol=1e-8, warm=0.995, ipm_slice=-1), xi, xf)
    assert a["iters"][0] == b["iters"][0] == 8 and np.array_equal(a["X_all"], b["X_all"]) and np.array_equal(a["U_all"], b["U_all"])


def _build_cpp_mirror(tmp_path):
    import subprocess
    exe = str(tmp_path / "cpp_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "scpp_b200"), "-lscpp_b200", "-Wl,-rpath," + os.path.join(ROOT, "scpp_b200")])
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_a_gpu(S, tmp_path):
    """include/scpp_b200.hpp (SCAlgorithm / SCvxAlgorithm with the reference's method names) builds against the C-ABI library; parameter
    loading and the error behaviour work on the host; with no CUDA device initialize() throws 'no CUDA device' (no fallback)"""
    import subprocess
    if S.device_count() > 0:
        pytest.skip("a GPU is present (covered by the gpu-marked test)")
    out = subprocess.run([_build_cpp_mirror(tmp_path), os.path.join(ROOT, "configs")], capture_output=True, text=True)
    assert out.returncode == 0 and "ok (no GPU)" in out.stdout, out.stdout + out.stderr


def test_mpc_dense_conic_source_vs_oracle():
    """K6 body on the host (mpc.cuh: dense Mehrotra / Nesterov-Todd solver, one problem per thread on the device) on the CONDENSED MPC problem
    against the oracle's generic conic solver on the FULL problem of buildMPCProblem (MPCProblem.cpp:6-87: X, U, equality-constrained
    dynamics), both built in tests/mpc_ref.py from the same exact discretisation; several states, K = 7 (MPC.info, 1.5 s) and K = 21 (2.5 s).
    Bars: optimal cost 1e-7 relative, thrust 1e-5 relative; gimbal angle 1e-4 rad at the shipped K = 7 (north_star's bar on the control) and
    5e-3 rad at K = 21, where the condensed normal equations reach their accuracy floor (dual residual ~1e-7, exit status 3) and the gimbal
    angle is only weakly determined by the cost (the two solutions differ by < 5e-8 relative in cost)."""
    import ctypes as C
    import mpc_ref as R
    p = O.rocket2d()
    w_term = np.array([5., 5, 5, 1, 1, 1]); w_in = np.array([0.1, 0.1])
    x_final = np.array([0., 0, 0, -1, 0, 0.])
    pp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(3)
    for K, horizon, gimbal_bar in ((7, 1.5, 1e-4), (21, 2.5, 5e-3)):
        A, B, z = R.discretize(p, horizon / (K - 1))
        compared = 0
        for trial in range(6):
            x_init = np.array([-20., 100., 2., -10., 0.05, 0.0]) * (1 + 0.15 * rng.standard_normal(6))
            P = R.full_socp(p, K, A, B, z, x_init, x_final, w_term, w_in)
            r = R.solve_with_oracle(O, P)
            if r["status"] != 0:
                continue                      # an instance the oracle cannot solve to tolerance (glide slope / minimum thrust make it infeasible)
            compared += 1
            Uo = np.array([[r["x"][P["iU"](k, i)] for i in range(2)] for k in range(K - 1)])
            nv, nl, cdim, G, c, h, _ = R.condensed(p, K, A, B, z, x_init, x_final, w_term, w_in)
            y = np.zeros(nv); it = C.c_int()
            st = H.lib().hs_dense_conic(nv, nl, len(cdim), pp(np.array(cdim, np.int32)), pp(G), pp(c), pp(h), C.c_double(1e-8), pp(y), C.byref(it))
            assert st in (0, 3) and it.value < 40
            assert abs(y[-2] + y[-1] - r["info"].pcost) < 1e-7 * abs(r["info"].pcost)
            U = y[:-2].reshape(K - 1, 2)
            assert np.abs(U[:, 0] - Uo[:, 0]).max() < gimbal_bar and np.abs(U[:, 1] - Uo[:, 1]).max() < 1e-5 * np.abs(Uo[:, 1]).max()
        assert compared >= 4


def test_mpc_info_loader(S):
    """scpp_b200_load_mpc_info == MPCAlgorithm::loadParameters (MPCAlgorithm.cpp:17-32) on the shipped Rocket2D MPC.info"""
    c = S.load_mpc_info(os.path.join(S.CONFIG_DIR, "Rocket2D", "MPC.info"), S.ROCKET2D)
    assert (c.K, c.time_horizon, c.nondimensionalize, c.constant_dynamics, c.intermediate_cost_active) == (7, 1.5, 0, 1, 0)
    assert list(c.state_weights_terminal)[:6] == [5., 5., 5., 1., 1., 1.] and np.allclose(list(c.input_weights)[:2], [0.1, 0.1])
    with pytest.raises(S.ScppError):
        S.load_mpc_info(os.path.join(S.CONFIG_DIR, "Rocket2D", "SC.info"), S.ROCKET2D)


def test_cvx_shim_lowers_the_reference_constraints(S, tmp_path):
    """include/scpp_cvx.hpp (recording shim of the reference's constraint DSL) + include/scpp_plugin.hpp (lowering to stage-wise tables):
    RocketQuat's and Rocket2d's addApplicationConstraints written call by call as in rocketQuat.cpp:70-144 / rocket2d.cpp:46-84 lower to
    exactly the tables the engine uses (scpp_b200_model_rows through the C-ABI), the pinned variables included; the problem-builder subset
    (SCProblem.cpp:16-134) records and evaluates (tests/cvx_shim_test.cpp)"""
    import subprocess
    exe = str(tmp_path / "cvx_shim_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cvx_shim_test.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "scpp_b200"), "-lscpp_b200", "-Wl,-rpath," + os.path.join(ROOT, "scpp_b200")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def test_generated_plugin_table_is_current(tmp_path):
    """scpp_b200/csrc/gen/*.inc are what tools/gen_plugin.cpp produces from scpp_b200/plugins/*.hpp today"""
    import subprocess
    exe = str(tmp_path / "gen_plugin")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "gen_plugin.cpp"), "-o", exe])
    subprocess.check_call([exe, str(tmp_path)])
    for f in ("rocket2d_plugin.inc", "rocketquat_roll_plugin.inc"):
        assert open(tmp_path / f).read() == open(os.path.join(ROOT, "scpp_b200", "csrc", "gen", f)).read(), f


def test_plugin_surface_model_source_vs_oracle(S):
    """a model written ONLY against the plugin surface (scpp_b200/plugins/rocket2d_plugin.hpp: generic-scalar flow map + constraints in the
    cvx:: DSL; Jacobians by dual numbers, row table / pins / constant slots generated at build time) through the kernel source on the host:
    same iteration count and convergence as the oracle's Rocket2D, iterates to 1e-5 / 1e-4; equal to the hand-written Rocket2d traits to
    solver accuracy; its dual-number Jacobian equals the hand-derived one; the engine reports the generated table through the C-ABI"""
    p = O.rocket2d()
    ocfg = O.sc_config(K=30, model=1, max_iterations=15)
    P, xi, xf = H.params_from_oracle(1, p)
    ro = O.sc_solve(O.ROCKET2D, p, ocfg)
    res = {}
    for model in (1, 2):
        r = H.sc_solve(model, P, H.sc_config(ocfg, nsub=-5, tol=1e-8), xi, xf)
        n = int(r["iters"][0])
        assert n == abs(ro["iterations"]) and bool(r["converged"][0]) == ro["converged"]
        for it in range(n + 1):
            assert np.abs(r["X_all"][0, it] - ro["X_all"][it]).max() < 1e-5 and np.abs(r["U_all"][0, it] - ro["U_all"][it]).max() < 1e-4, (model, it)
        res[model] = r
    assert np.abs(res[1]["X_all"] - res[2]["X_all"]).max() < 1e-6 and np.abs(res[1]["U_all"] - res[2]["U_all"]).max() < 1e-6
    # Jacobians: AutoJacobian (duals over the plugin's flow map) == Rocket2d's hand-derived Lin
    rng = np.random.default_rng(5)
    x = rng.standard_normal(6); u = np.array([0.1, 0.7]); par = np.array([1.0, 0.3, 0.0, -0.01, 0.0, -0.02])
    f = np.zeros(6); A_ad = np.zeros((6, 6)); B_ad = np.zeros((6, 2)); A1 = np.zeros((6, 6)); B1 = np.zeros((6, 2)); A2 = np.zeros((6, 6)); B2 = np.zeros((6, 2))
    pp = lambda a: a.ctypes.data_as(C.c_void_p)
    H.lib().hs_jacobians(1, pp(x), pp(u), pp(par), pp(f), pp(A_ad), pp(B_ad), pp(A1), pp(B1))
    H.lib().hs_jacobians(2, pp(x), pp(u), pp(par), pp(f), pp(A_ad), pp(B_ad), pp(A2), pp(B2))
    assert np.abs(A1 - A2).max() < 1e-14 and np.abs(B1 - B2).max() < 1e-14 and np.abs(A2 - A_ad).max() < 1e-14
    # the table in use, through the C-ABI: 8 LP rows (4 boxes) and the glide-slope cone, numerically equal to the hand-written model's
    model, params, x_init, x_final, _ = S.load_model("Rocket2D")
    lp_a, cones_a = S.model_rows(S.ROCKET2D, params, x_init, x_final); lp_b, cones_b = S.model_rows(S.ROCKET2D_PLUGIN, params, x_init, x_final)
    key = lambda r: (sorted(r[0].items()), r[1])
    assert sorted(map(key, lp_a)) == sorted(map(key, lp_b)) and [list(map(key, c)) for c in cones_a] == [list(map(key, c)) for c in cones_b]


def test_rocketquat_with_roll_control_through_the_plugin_surface(S):
    """enable_roll_control = true (rocketQuat.cpp:135-138: |roll torque| <= t_max, w_z and the torque free) is built as a model written against
    the plugin surface only (scpp_b200/plugins/rocketquat_plugin.hpp, model id 3: RocketQuat's flow map, dual-number Jacobians, generated
    table with 4 LP rows): kernel source on the host against the oracle with enable_roll_control = 1, the torque really used; the
    hand-written model refuses the setting and points at model 3"""
    p, rpy = O.falcon9()
    for inst in (None, 3):
        q = p if inst is None else O.rq_perturb(p, rpy, 0x5C99, inst)
        q.enable_roll_control = 1
        q.x_init[13] = 0.02                       # an initial roll rate: the roll torque has something to do
        ocfg = O.sc_config(K=20, max_iterations=5)
        ro = O.sc_solve(O.ROCKETQUAT, q, ocfg)
        P, xi, xf = H.params_from_oracle(3, q)
        r = H.sc_solve(3, P, H.sc_config(ocfg, nsub=-5, tol=1e-8), xi, xf)
        n = int(r["iters"][0])
        assert n == abs(ro["iterations"]) == 5
        for it in range(n + 1):
            assert np.abs(r["X_all"][0, it] - ro["X_all"][it]).max() < 1e-5 and np.abs(r["U_all"][0, it] - ro["U_all"][it]).max() < 1e-4, (inst, it)
        t_max = q.t_max / (q.x_init[0] * np.linalg.norm(q.x_init[1:4]) ** 2)
        tq = r["U_all"][0, n, :, 3]
        assert np.abs(tq).max() > 0.5 * t_max and np.abs(tq).max() <= t_max * (1 + 1e-6)          # active and inside its box
    model, params, x_init, x_final, cfg = S.load_model("RocketQuatRoll")
    assert model == S.ROCKETQUAT_ROLL and params.enable_roll_control == 1
    lp, cones = S.model_rows(model, params, x_init, x_final)
    assert len(lp) == 4 and [len(c) for c in cones] == [3, 3, 4, 4, 3]
    assert sorted(r[1] for r in lp if list(r[0]) == [17]) == [params.t_max, params.t_max] and sorted(list(r[0].values())[0] for r in lp if list(r[0]) == [17]) == [-1., 1.]
    with pytest.raises(S.ScppError):
        S.load_model_info(os.path.join(S.CONFIG_DIR, "RocketQuatRoll", "model.info"), S.ROCKETQUAT)      # the hand-written table has no roll rows


def test_fixed_final_time_sc_source_vs_oracle():
    """SC with free_final_time = false (SCProblem.cpp:27-35,49-56,82-100: no sigma / delta_sigma variables; fixed-time discretisation
    discretizationImplementation.hpp:111-116): the interior-point solver pins sigma (the global rows and the border of the Newton system are
    skipped, z_fixed = z + s sigma falls out of the free-time tile) -- kernel source on the host, monolithic and split pipeline, against the
    oracle with free_final_time = 0; the final time stays at model.info's final_time"""
    for model, p, K in ((1, O.rocket2d(), 30), (0, O.falcon9()[0], 20)):
        ocfg = O.sc_config(K=K, model=model, max_iterations=6)
        ocfg.free_final_time = 0
        ro = O.sc_solve(model, p, ocfg)
        P, xi, xf = H.params_from_oracle(model, p)
        for sl in (1, -1):
            r = H.sc_solve(model, P, H.sc_config(ocfg, nsub=-5, tol=1e-8, ipm_slice=sl), xi, xf)
            n = int(r["iters"][0])
            assert n == abs(ro["iterations"]) and bool(r["converged"][0]) == ro["converged"]
            assert r["t"][0] == p.final_time and np.all(ro["t_all"] == p.final_time)
            for it in range(n + 1):
                assert np.abs(r["X_all"][0, it] - ro["X_all"][it]).max() < 1e-5 and np.abs(r["U_all"][0, it] - ro["U_all"][it]).max() < 1e-4, (model, sl, it)
