"""ctypes binding of the CPU oracle (oracle/liborc.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

ROCKETQUAT, ROCKET2D = 0, 1
DIMS = {ROCKETQUAT: (14, 4, 10), ROCKET2D: (6, 2, 6)}


class RQParams(C.Structure):
    _fields_ = [("g_I", C.c_double * 3), ("J_B", C.c_double * 3), ("r_T_B", C.c_double * 3),
                ("alpha_m", C.c_double), ("T_min", C.c_double), ("T_max", C.c_double), ("t_max", C.c_double),
                ("gimbal_max", C.c_double), ("theta_max", C.c_double), ("gamma_gs", C.c_double), ("w_B_max", C.c_double),
                ("x_init", C.c_double * 14), ("x_final", C.c_double * 14), ("final_time", C.c_double),
                ("exact_minimum_thrust", C.c_int), ("enable_roll_control", C.c_int),
                ("m_scale", C.c_double), ("r_scale", C.c_double)]


class R2DParams(C.Structure):
    _fields_ = [("g_I", C.c_double * 2), ("J_B", C.c_double), ("r_T_B", C.c_double * 2), ("m", C.c_double),
                ("T_min", C.c_double), ("T_max", C.c_double),
                ("gimbal_max", C.c_double), ("theta_max", C.c_double), ("gamma_gs", C.c_double), ("w_B_max", C.c_double),
                ("x_init", C.c_double * 6), ("x_final", C.c_double * 6), ("final_time", C.c_double),
                ("constrain_initial_final", C.c_int), ("m_scale", C.c_double), ("r_scale", C.c_double)]


class SCConfig(C.Structure):
    _fields_ = [("K", C.c_int), ("free_final_time", C.c_int), ("interpolate_input", C.c_int), ("nondimensionalize", C.c_int),
                ("weight_time", C.c_double), ("weight_trust_region_time", C.c_double),
                ("weight_trust_region_trajectory", C.c_double), ("weight_virtual_control", C.c_double),
                ("nu_tol", C.c_double), ("delta_tol", C.c_double), ("max_iterations", C.c_int)]


class IpmInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("iterations", C.c_int), ("pres", C.c_double), ("dres", C.c_double),
                ("gap", C.c_double), ("relgap", C.c_double), ("pcost", C.c_double), ("dcost", C.c_double),
                ("cone_viol", C.c_double), ("kkt_resid", C.c_double)]


class IterInfo(C.Structure):
    _fields_ = [("norm1_nu", C.c_double), ("sum_delta", C.c_double), ("delta_sigma", C.c_double), ("sigma", C.c_double),
                ("weight_tr_used", C.c_double), ("ipm", IpmInfo), ("t_discretize_ms", C.c_double), ("t_solve_ms", C.c_double)]


class SocpDims(C.Structure):
    _fields_ = [("n", C.c_int), ("p", C.c_int), ("m", C.c_int), ("l", C.c_int), ("ncones", C.c_int),
                ("nnzA", C.c_int), ("nnzG", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liborc.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_uniform_pm1.restype = C.c_double
        _LIB.orc_uniform_pm1.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint]
        _LIB.orc_rq_perturb.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_void_p]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def f(model, x, u, par):
    nx = DIMS[model][0]
    out = np.zeros(nx)
    lib().orc_f(model, _p(np.ascontiguousarray(x, float)), _p(np.ascontiguousarray(u, float)), _p(np.ascontiguousarray(par, float)), _p(out))
    return out


def jac(model, x, u, par):
    nx, nu, _ = DIMS[model]
    A = np.zeros((nx, nx), order="F")
    B = np.zeros((nx, nu), order="F")
    lib().orc_jac(model, _p(np.ascontiguousarray(x, float)), _p(np.ascontiguousarray(u, float)), _p(np.ascontiguousarray(par, float)), _p(A), _p(B))
    return A, B


def discretize(model, X, U, sigma, par, foh=True, free_time=True):
    nx, nu, _ = DIMS[model]
    K = X.shape[0]
    X = np.ascontiguousarray(X, float); U = np.ascontiguousarray(U, float); par = np.ascontiguousarray(par, float)
    A = np.zeros((K - 1, nx, nx)); B = np.zeros((K - 1, nu, nx)); Cc = np.zeros((K - 1, nu, nx))
    s = np.zeros((K - 1, nx)); z = np.zeros((K - 1, nx))
    lib().orc_discretize(model, K, _p(X), _p(U), C.c_double(sigma), _p(par), int(foh), int(free_time), _p(A), _p(B), _p(Cc), _p(s), _p(z))
    # column-major blocks -> numpy [k][row][col]
    return dict(A=A.transpose(0, 2, 1).copy(), B=B.transpose(0, 2, 1).copy(), C=Cc.transpose(0, 2, 1).copy(), s=s, z=z)


def dd_colmajor(dd):
    """numpy [k][row][col] -> flat column-major arrays the C API expects"""
    return (np.ascontiguousarray(dd["A"].transpose(0, 2, 1)), np.ascontiguousarray(dd["B"].transpose(0, 2, 1)),
            np.ascontiguousarray(dd["C"].transpose(0, 2, 1)), np.ascontiguousarray(dd["s"]), np.ascontiguousarray(dd["z"]))


def simulate(model, dt, u0, u1, par, x):
    x = np.array(x, float)
    lib().orc_simulate(model, C.c_double(dt), _p(np.ascontiguousarray(u0, float)), _p(np.ascontiguousarray(u1, float)), _p(np.ascontiguousarray(par, float)), _p(x))
    return x


def subproblem(model, params, cfg, weight_tr, Xbar, Ubar, sigmabar, dd, thrust_dir=None):
    nx, nu, _ = DIMS[model]
    K = cfg.K
    A, B, Cc, s, z = dd_colmajor(dd)
    X = np.zeros((K, nx)); U = np.zeros((K, nu)); nu_v = np.zeros((K - 1, nx)); delta = np.zeros(K)
    sigma = C.c_double(); n1 = C.c_double(); ds = C.c_double(); info = IpmInfo()
    td = np.ascontiguousarray(thrust_dir, float) if thrust_dir is not None else None
    st = lib().orc_sc_subproblem(model, C.byref(params), C.byref(cfg), C.c_double(weight_tr),
                                 _p(np.ascontiguousarray(Xbar, float)), _p(np.ascontiguousarray(Ubar, float)), C.c_double(sigmabar),
                                 _p(A), _p(B), _p(Cc), _p(s), _p(z), _p(td),
                                 _p(X), _p(U), C.byref(sigma), _p(nu_v), _p(delta), C.byref(n1), C.byref(ds), C.byref(info))
    return dict(status=st, X=X, U=U, sigma=sigma.value, nu=nu_v, delta=delta, norm1_nu=n1.value, delta_sigma=ds.value, info=info)


def export(model, params, cfg, weight_tr, Xbar, Ubar, sigmabar, dd, thrust_dir=None):
    import scipy.sparse as sp
    nx, nu, _ = DIMS[model]
    K = cfg.K
    A, B, Cc, s, z = dd_colmajor(dd)
    td = np.ascontiguousarray(thrust_dir, float) if thrust_dir is not None else None
    dims = SocpDims()
    args = [model, C.byref(params), C.byref(cfg), C.c_double(weight_tr), _p(np.ascontiguousarray(Xbar, float)),
            _p(np.ascontiguousarray(Ubar, float)), C.c_double(sigmabar), _p(A), _p(B), _p(Cc), _p(s), _p(z), _p(td)]
    lib().orc_sc_export(*args, C.byref(dims), *([None] * 13))
    c = np.zeros(dims.n); b = np.zeros(dims.p); h = np.zeros(dims.m); q = np.zeros(dims.ncones, np.int32)
    Ai = np.zeros(dims.nnzA, np.int32); Aj = np.zeros(dims.nnzA, np.int32); Av = np.zeros(dims.nnzA)
    Gi = np.zeros(dims.nnzG, np.int32); Gj = np.zeros(dims.nnzG, np.int32); Gv = np.zeros(dims.nnzG)
    iX = np.zeros((K, nx), np.int32); iU = np.zeros((K, nu), np.int32); isg = C.c_int()
    lib().orc_sc_export(*args, C.byref(dims), _p(c), _p(b), _p(h), _p(q), _p(Ai), _p(Aj), _p(Av), _p(Gi), _p(Gj), _p(Gv), _p(iX), _p(iU), C.byref(isg))
    Am = sp.coo_matrix((Av, (Ai, Aj)), shape=(dims.p, dims.n)).tocsr()
    Gm = sp.coo_matrix((Gv, (Gi, Gj)), shape=(dims.m, dims.n)).tocsr()
    return dict(c=c, b=b, h=h, q=q, l=dims.l, A=Am, G=Gm, iX=iX, iU=iU, isigma=isg.value,
                raw=(Ai, Aj, Av, Gi, Gj, Gv))


def conic_solve(c, b, h, l, q, Acoo, Gcoo):
    n, p, m = len(c), len(b), len(h)
    Ai, Aj, Av = Acoo; Gi, Gj, Gv = Gcoo
    x = np.zeros(n); y = np.zeros(p); s = np.zeros(m); z = np.zeros(m); info = IpmInfo()
    q = np.ascontiguousarray(q, np.int32)
    mk = lambda a, t: np.ascontiguousarray(a, t)
    st = lib().orc_conic_solve(n, p, m, l, len(q), _p(q), _p(mk(c, float)), _p(mk(b, float)), _p(mk(h, float)),
                               len(Av), _p(mk(Ai, np.int32)), _p(mk(Aj, np.int32)), _p(mk(Av, float)),
                               len(Gv), _p(mk(Gi, np.int32)), _p(mk(Gj, np.int32)), _p(mk(Gv, float)),
                               _p(x), _p(y), _p(s), _p(z), C.byref(info))
    return dict(status=st, x=x, y=y, s=s, z=z, info=info)


def sc_solve(model, params, cfg):
    nx, nu, _ = DIMS[model]
    K, M = cfg.K, cfg.max_iterations + 1
    Xa = np.zeros((M, K, nx)); Ua = np.zeros((M, K, nu)); ta = np.zeros(M)
    info = (IterInfo * cfg.max_iterations)()
    Xo = np.zeros((K, nx)); Uo = np.zeros((K, nu)); to = C.c_double(); conv = C.c_int()
    it = lib().orc_sc_solve(model, C.byref(params), C.byref(cfg), _p(Xa), _p(Ua), _p(ta), info, _p(Xo), _p(Uo), C.byref(to), C.byref(conv))
    n = abs(it)
    return dict(iterations=it, converged=bool(conv.value), X_all=Xa[:n + 1], U_all=Ua[:n + 1], t_all=ta[:n + 1],
                info=[info[i] for i in range(n)], X=Xo, U=Uo, t=to.value)


def rq_perturb(nominal, rpy_init, seed, instance):
    out = RQParams()
    r = np.ascontiguousarray(rpy_init, float)
    lib().orc_rq_perturb(C.byref(nominal), _p(r), seed, instance, C.byref(out))
    return out


def euler_to_quat_xyz(rpy):
    q = np.zeros(4)
    lib().orc_euler_to_quat_xyz(_p(np.ascontiguousarray(rpy, float)), _p(q))
    return q


# ---- the reference's parameter sets (scpp_models/config/RocketQuat/model.info:111-213 Falcon 9 block,
#      :1-105 Starship block; Rocket2D/model.info) expressed directly, used when no .info loader is involved
def falcon9():
    d2r = np.pi / 180
    p = RQParams()
    p.g_I[:] = [0, 0, -9.81]; p.J_B[:] = [5e6, 5e6, 7e4]; p.r_T_B[:] = [0, 0, -15.]
    p.alpha_m = 1. / (275. * 9.81)
    p.T_min, p.T_max, p.t_max = 200000., 420000., 17500.
    p.gimbal_max, p.theta_max, p.gamma_gs, p.w_B_max = 15 * d2r, 90 * d2r, 30 * d2r, 60 * d2r
    rpy = np.array([-20., 20., 0.]) * d2r
    q0 = euler_to_quat_xyz(rpy); q1 = euler_to_quat_xyz(np.zeros(3))
    p.x_init[:] = [24000., 200, 200, 800, -40, -40, -80, *q0, 0, 0, 0]
    p.x_final[:] = [22000., 0, 0, 0, 0, 0, 0, *q1, 0, 0, 0]
    p.final_time = 12.
    p.exact_minimum_thrust, p.enable_roll_control = 1, 0
    return p, rpy


def starship():
    d2r = np.pi / 180
    p = RQParams()
    p.g_I[:] = [0, 0, -9.81]; p.J_B[:] = [27741458., 27741458., 1316250.]; p.r_T_B[:] = [0, 0, -25.]
    p.alpha_m = 1. / (330. * 9.81)
    p.T_min, p.T_max, p.t_max = 2000000., 3000000., 17500.
    p.gimbal_max, p.theta_max, p.gamma_gs, p.w_B_max = 15 * d2r, 90 * d2r, 45 * d2r, 60 * d2r
    rpy = np.array([70., 0., 0.]) * d2r
    q0 = euler_to_quat_xyz(rpy); q1 = euler_to_quat_xyz(np.zeros(3))
    p.x_init[:] = [140000., 0, 250, 500, 0, -15, -60, *q0, 0, 0, 0]
    p.x_final[:] = [120000., 0, 0, 0, 0, 0, 0, *q1, 0, 0, 0]
    p.final_time = 11.
    p.exact_minimum_thrust, p.enable_roll_control = 1, 0
    return p, rpy


def rocket2d():
    d2r = np.pi / 180
    p = R2DParams()
    p.g_I[:] = [0, -9.81]; p.J_B = 5e6; p.r_T_B[:] = [0, -15.]; p.m = 24000.
    p.T_min, p.T_max = 10000., 420000.
    p.gimbal_max, p.theta_max, p.gamma_gs, p.w_B_max = 15 * d2r, 60 * d2r, 45 * d2r, 20 * d2r
    p.x_init[:] = [-200, 800, 0, -100, -20 * d2r, 0]
    p.x_final[:] = [0, 0, 0, -1, 0, 0]
    p.final_time = 12.
    p.constrain_initial_final = 1
    return p


def sc_config(K=50, model=ROCKETQUAT, max_iterations=15):
    c = SCConfig()
    c.K = K; c.free_final_time = 1; c.interpolate_input = 1; c.nondimensionalize = 1
    c.weight_time = 1.; c.weight_trust_region_time = 1.
    c.weight_trust_region_trajectory = 50. if model == ROCKETQUAT else 1.
    c.weight_virtual_control = 1000.; c.nu_tol = 1e-5; c.delta_tol = 1e-3; c.max_iterations = max_iterations
    return c


# ---- SCvx variant (oracle/orc_sc.c: orc_scvx_*) -----------------------------------------------------------------------
class SCvxConfig(C.Structure):
    _fields_ = [("K", C.c_int), ("interpolate_input", C.c_int), ("nondimensionalize", C.c_int),
                ("rho_0", C.c_double), ("rho_1", C.c_double), ("rho_2", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
                ("change_threshold", C.c_double), ("weight_virtual_control", C.c_double), ("trust_region", C.c_double),
                ("max_iterations", C.c_int)]


class SCvxInfo(C.Structure):
    _fields_ = [("norm1_nu", C.c_double), ("nonlinear_cost", C.c_double), ("actual_change", C.c_double), ("predicted_change", C.c_double),
                ("rho", C.c_double), ("trust_region_used", C.c_double), ("trust_region_next", C.c_double),
                ("solves", C.c_int), ("pad_", C.c_int), ("ipm", IpmInfo)]


def scvx_config(K=30, model=ROCKETQUAT, max_iterations=None):
    """scpp_models/config/{RocketQuat,Rocket2D}/SCvx.info with K overridable"""
    c = SCvxConfig()
    c.K = K; c.interpolate_input = 1; c.nondimensionalize = 1 if model == ROCKETQUAT else 0
    c.rho_0 = 0.0; c.rho_1 = 0.25; c.rho_2 = 0.9; c.alpha = 2.0; c.beta = 3.2
    c.change_threshold = 1e-3 if model == ROCKETQUAT else 1e-2
    c.weight_virtual_control = 1e3; c.trust_region = 5.
    c.max_iterations = max_iterations if max_iterations is not None else (30 if model == ROCKETQUAT else 20)
    return c


def scvx_solve(model, params, cfg):
    nx, nu, _ = DIMS[model]
    K, M = cfg.K, cfg.max_iterations + 1
    Xa = np.zeros((M, K, nx)); Ua = np.zeros((M, K, nu))
    info = (SCvxInfo * cfg.max_iterations)()
    Xo = np.zeros((K, nx)); Uo = np.zeros((K, nu)); to = C.c_double(); conv = C.c_int()
    it = lib().orc_scvx_solve(model, C.byref(params), C.byref(cfg), _p(Xa), _p(Ua), info, _p(Xo), _p(Uo), C.byref(to), C.byref(conv))
    n = abs(it)
    return dict(iterations=it, converged=bool(conv.value), X_all=Xa[:n + 1], U_all=Ua[:n + 1], info=[info[i] for i in range(n)], X=Xo, U=Uo, t=to.value)


def scvx_subproblem(model, params, K, weight_vc, trust_region, Ubar, dd, thrust_dir=None):
    nx, nu, _ = DIMS[model]
    A, B, Cc, s, z = dd_colmajor(dd)
    X = np.zeros((K, nx)); U = np.zeros((K, nu)); nu_v = np.zeros((K - 1, nx)); n1 = C.c_double(); info = IpmInfo()
    td = np.ascontiguousarray(thrust_dir, float) if thrust_dir is not None else None
    st = lib().orc_scvx_subproblem(model, C.byref(params), K, C.c_double(weight_vc), C.c_double(trust_region), _p(np.ascontiguousarray(Ubar, float)),
                                   _p(A), _p(B), _p(Cc), _p(z), _p(td), _p(X), _p(U), _p(nu_v), C.byref(n1), C.byref(info))
    return dict(status=st, X=X, U=U, nu=nu_v, norm1_nu=n1.value, info=info)


def scvx_nonlinear_cost(model, X, U, t, par):
    lib().orc_scvx_nonlinear_cost.restype = C.c_double
    return lib().orc_scvx_nonlinear_cost(model, X.shape[0], _p(np.ascontiguousarray(X, float)), _p(np.ascontiguousarray(U, float)), C.c_double(t),
                                         _p(np.ascontiguousarray(par, float)))


# ---- closed loop (scpp/src/SC_sim.cpp) ----------------------------------------------------------------------------------
def sc_sim(model, params, cfg, time_step=0.05, max_steps=100):
    nx, nu, _ = DIMS[model]
    Xs = np.zeros((max_steps, nx)); Us = np.zeros((max_steps, nu)); iters = np.zeros(max_steps, np.int32); reached = C.c_int()
    n = lib().orc_sc_sim(model, C.byref(params), C.byref(cfg), C.c_double(time_step), max_steps, _p(Xs), _p(Us), _p(iters), C.byref(reached))
    m = n if n >= 0 else -n - 1
    return dict(steps=n, X_sim=Xs[:m], U_sim=Us[:m], iters=iters[:max(m, 1)], reached_end=bool(reached.value))
