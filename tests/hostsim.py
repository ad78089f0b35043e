"""TEST-ONLY: builds and binds the host-simulation library (kernel bodies compiled by g++ with LANES == 1).
Never imported by the scpp_b200 package."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class ModelParamsHost(C.Structure):
    _fields_ = [("g_I", C.c_double * 3), ("J_B", C.c_double * 3), ("r_T_B", C.c_double * 3), ("alpha_m", C.c_double),
                ("m", C.c_double), ("T_min", C.c_double), ("T_max", C.c_double), ("t_max", C.c_double),
                ("gimbal_max", C.c_double), ("theta_max", C.c_double), ("gamma_gs", C.c_double), ("w_B_max", C.c_double),
                ("final_time", C.c_double), ("exact_minimum_thrust", C.c_int), ("enable_roll_control", C.c_int),
                ("constrain_initial_final", C.c_int), ("pad_", C.c_int)]


class IpmSettings(C.Structure):
    _fields_ = [("feastol", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double), ("maxit", C.c_int), ("stalled_step", C.c_int), ("warm", C.c_double)]


class ScConfig(C.Structure):
    _fields_ = [("K", C.c_int), ("free_final_time", C.c_int), ("interpolate_input", C.c_int), ("nondimensionalize", C.c_int),
                ("weight_time", C.c_double), ("weight_trust_region_time", C.c_double),
                ("weight_trust_region_trajectory", C.c_double), ("weight_virtual_control", C.c_double),
                ("nu_tol", C.c_double), ("delta_tol", C.c_double), ("max_iterations", C.c_int), ("nsub", C.c_int),
                ("keep_history", C.c_int), ("ipm_slice", C.c_int), ("ipm", IpmSettings),
                ("algorithm", C.c_int), ("solver", C.c_int), ("scvx_rho_0", C.c_double), ("scvx_rho_1", C.c_double), ("scvx_rho_2", C.c_double),
                ("scvx_alpha", C.c_double), ("scvx_beta", C.c_double), ("scvx_change_threshold", C.c_double), ("scvx_trust_region", C.c_double),
                ("jacobian", C.c_int), ("pad3_", C.c_int)]


def build():
    out = os.path.join(ROOT, "tests", "_hostsim")
    os.makedirs(out, exist_ok=True)
    src = os.path.join(ROOT, "scpp_b200", "csrc")
    lib = os.path.join(out, "libhostsim.so")
    deps = [os.path.join(src, f) for f in os.listdir(src) if f.endswith((".cuh", ".cpp", ".hpp"))]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-w", "-I" + src,
                           os.path.join(src, "hostsim.cpp"), "-o", lib])
    return lib


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        assert _LIB.hs_sizes(0) == C.sizeof(ModelParamsHost) and _LIB.hs_sizes(1) == C.sizeof(ScConfig)
    return _LIB


def params_from_oracle(model, p):
    """orc_py.RQParams / R2DParams -> (ModelParamsHost, x_init, x_final)"""
    P = ModelParamsHost()
    if model in (0, 3):
        P.g_I[:] = list(p.g_I); P.J_B[:] = list(p.J_B); P.r_T_B[:] = list(p.r_T_B)
        P.alpha_m = p.alpha_m; P.T_min = p.T_min; P.T_max = p.T_max; P.t_max = p.t_max
        P.gimbal_max = p.gimbal_max; P.theta_max = p.theta_max; P.gamma_gs = p.gamma_gs; P.w_B_max = p.w_B_max
        P.final_time = p.final_time; P.exact_minimum_thrust = p.exact_minimum_thrust; P.enable_roll_control = p.enable_roll_control
    else:
        P.g_I[:] = [p.g_I[0], p.g_I[1], 0]; P.J_B[:] = [p.J_B, 0, 0]; P.r_T_B[:] = [p.r_T_B[0], p.r_T_B[1], 0]
        P.m = p.m; P.T_min = p.T_min; P.T_max = p.T_max
        P.gimbal_max = p.gimbal_max; P.theta_max = p.theta_max; P.gamma_gs = p.gamma_gs; P.w_B_max = p.w_B_max
        P.final_time = p.final_time; P.constrain_initial_final = p.constrain_initial_final
    return P, np.array(p.x_init, float), np.array(p.x_final, float)


def sc_config(ocfg, nsub=20, tol=1e-9, maxit=100, history=True, warm=0.0, ipm_slice=1):
    c = ScConfig()
    for f in ("K", "free_final_time", "interpolate_input", "nondimensionalize", "weight_time", "weight_trust_region_time",
              "weight_trust_region_trajectory", "weight_virtual_control", "nu_tol", "delta_tol", "max_iterations"):
        setattr(c, f, getattr(ocfg, f))
    c.nsub = nsub; c.keep_history = int(history); c.ipm_slice = ipm_slice
    c.ipm.feastol = tol; c.ipm.abstol = tol; c.ipm.reltol = tol; c.ipm.maxit = maxit; c.ipm.warm = warm
    return c


def scvx_config(ov, final_time_free=False, nsub=20, tol=1e-8, maxit=100, history=True, warm=0.0, ipm_slice=1):
    """orc_py.SCvxConfig -> ScConfig with algorithm = 1 (SCvx)"""
    c = ScConfig()
    c.K = ov.K; c.free_final_time = 1; c.interpolate_input = ov.interpolate_input; c.nondimensionalize = ov.nondimensionalize
    c.weight_time = 0.; c.weight_trust_region_time = 0.; c.weight_trust_region_trajectory = 0.; c.weight_virtual_control = ov.weight_virtual_control
    c.nu_tol = 0.; c.delta_tol = 0.; c.max_iterations = ov.max_iterations
    c.nsub = nsub; c.keep_history = int(history); c.ipm_slice = ipm_slice
    c.ipm.feastol = tol; c.ipm.abstol = tol; c.ipm.reltol = tol; c.ipm.maxit = maxit; c.ipm.warm = warm
    c.algorithm = 1
    c.scvx_rho_0 = ov.rho_0; c.scvx_rho_1 = ov.rho_1; c.scvx_rho_2 = ov.rho_2; c.scvx_alpha = ov.alpha; c.scvx_beta = ov.beta
    c.scvx_change_threshold = ov.change_threshold; c.scvx_trust_region = ov.trust_region
    return c


DIMS = {0: (14, 4), 1: (6, 2), 2: (6, 2), 3: (14, 4)}      # 2: Rocket2dPlugin, 3: RocketQuatRollPlugin


def sc_solve(model, P, cfg, x_init, x_final):
    nx, nu = DIMS[model]
    x_init = np.ascontiguousarray(np.atleast_2d(x_init), float); x_final = np.ascontiguousarray(np.atleast_2d(x_final), float)
    N, K, M = x_init.shape[0], cfg.K, cfg.max_iterations
    if x_final.shape[0] == 1 and N > 1:
        x_final = np.ascontiguousarray(np.tile(x_final, (N, 1)))
    X = np.zeros((N, K, nx)); U = np.zeros((N, K, nu)); sg = np.zeros(N)
    iters = np.zeros(N, np.int32); status = np.zeros(N, np.int32); conv = np.zeros(N, np.int32)
    hist = np.zeros((N, M + 1, K * (nx + nu) + 1)); info = np.zeros((N, M, 10))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().hs_sc_solve(model, C.byref(P), C.byref(cfg), N, p(x_init), p(x_final), p(X), p(U), p(sg), p(iters), p(status), p(conv), p(hist), p(info))
    H = hist[:, :, :-1].reshape(N, M + 1, K, nx + nu)
    return dict(X=X, U=U, t=sg, iters=iters, status=status, converged=conv, X_all=H[..., :nx], U_all=H[..., nx:], t_all=hist[:, :, -1], info=info)


def sc_sim(model, P, cfg, x_init, x_final, time_step, steps):
    nx, nu = DIMS[model]
    x_init = np.ascontiguousarray(np.atleast_2d(x_init), float); N = x_init.shape[0]
    x_final = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(x_final), (N, nx)), float)
    Xs = np.zeros((steps, N, nx)); Us = np.zeros((steps, N, nu)); iters = np.zeros((steps, N), np.int32); reached = np.zeros(N, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().hs_sc_sim(model, C.byref(P), C.byref(cfg), N, p(x_init), p(x_final), C.c_double(time_step), steps, p(Xs), p(Us), p(iters), p(reached))
    return dict(X_sim=Xs, U_sim=Us, iters=iters, reached=reached)


def simulate(model, dt, u0, u1, par, x):
    x = np.array(x, float)
    p = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(C.c_void_p)
    lib().hs_simulate(model, C.c_double(dt), x.ctypes.data_as(C.c_void_p), p(u0), p(u1), p(par))
    return x


def discretize(model, X, U, sigma, par, nsub, jacobian=0, zoh=False):
    nx, nu = DIMS[model]
    K = X.shape[0]
    out = np.zeros((K - 1, nx, nx + 2 * nu + 2))
    p = lambda a: np.ascontiguousarray(a, float).ctypes.data_as(C.c_void_p)
    (lib().hs_discretize_zoh if zoh else lib().hs_discretize2)(model, K, p(X), p(U), C.c_double(sigma), p(par), nsub, jacobian, out.ctypes.data_as(C.c_void_p))
    return dict(A=out[:, :, :nx], B=out[:, :, nx:nx + nu], C=out[:, :, nx + nu:nx + 2 * nu], s=out[:, :, nx + 2 * nu], z=out[:, :, nx + 2 * nu + 1])
