"""scpp_b200 — B200-native batched successive convexification behind SCpp's SCAlgorithm / model-plugin surface.

This Python module is the host-side mirror of the reference's C++ API used by the tests and bench.py; it only
marshals numpy arrays into the C-ABI of ``libscpp_b200.so`` (include/scpp_b200.h).  All compute runs in the CUDA
library: there is no CPU fallback, and every compute call raises if the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCPP_B200_LIB", os.path.join(_HERE, "libscpp_b200.so"))   # the override is for A/B experiments with kernel variants
CONFIG_DIR = os.path.join(os.path.dirname(_HERE), "configs")

ROCKETQUAT, ROCKET2D, ROCKET2D_PLUGIN, ROCKETQUAT_ROLL = 0, 1, 2, 3      # 2, 3: models written only against the plugin surface (scpp_b200/plugins/)
MODEL_NAMES = {ROCKETQUAT: "RocketQuat", ROCKET2D: "Rocket2D", ROCKET2D_PLUGIN: "Rocket2D", ROCKETQUAT_ROLL: "RocketQuatRoll"}
INFO_STRIDE = 10
INFO_FIELDS = ("norm1_nu", "sum_delta", "delta_sigma", "sigma", "weight_tr_used", "ipm_iterations", "ipm_status", "pres", "dres", "relgap")


class ModelParams(C.Structure):
    """scpp_b200_model_params == RocketQuat::Parameters / Rocket2d::Parameters (rocketQuat.hpp:50-85, rocket2d.hpp:51-84)"""
    _fields_ = [("g_I", C.c_double * 3), ("J_B", C.c_double * 3), ("r_T_B", C.c_double * 3), ("alpha_m", C.c_double),
                ("m", C.c_double), ("T_min", C.c_double), ("T_max", C.c_double), ("t_max", C.c_double),
                ("gimbal_max", C.c_double), ("theta_max", C.c_double), ("gamma_gs", C.c_double), ("w_B_max", C.c_double),
                ("final_time", C.c_double), ("exact_minimum_thrust", C.c_int), ("enable_roll_control", C.c_int),
                ("constrain_initial_final", C.c_int), ("pad_", C.c_int)]


class IpmSettings(C.Structure):
    _fields_ = [("feastol", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double), ("maxit", C.c_int), ("stalled_step", C.c_int), ("warm", C.c_double)]


class SCConfig(C.Structure):
    """scpp_b200_sc_config == SC.info (SCAlgorithm.cpp:22-46) + engine knobs"""
    _fields_ = [("K", C.c_int), ("free_final_time", C.c_int), ("interpolate_input", C.c_int), ("nondimensionalize", C.c_int),
                ("weight_time", C.c_double), ("weight_trust_region_time", C.c_double),
                ("weight_trust_region_trajectory", C.c_double), ("weight_virtual_control", C.c_double),
                ("nu_tol", C.c_double), ("delta_tol", C.c_double), ("max_iterations", C.c_int), ("nsub", C.c_int),
                ("keep_history", C.c_int), ("ipm_slice", C.c_int), ("ipm", IpmSettings),
                ("algorithm", C.c_int), ("solver", C.c_int), ("scvx_rho_0", C.c_double), ("scvx_rho_1", C.c_double), ("scvx_rho_2", C.c_double),
                ("scvx_alpha", C.c_double), ("scvx_beta", C.c_double), ("scvx_change_threshold", C.c_double), ("scvx_trust_region", C.c_double),
                ("jacobian", C.c_int), ("pad3_", C.c_int)]


class ScppError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libscpp_b200.so; raises loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScppError(f"{LIB_PATH} is missing: build it with `python -m scpp_b200.build` (nvcc, sm_100a)")
        L = C.CDLL(LIB_PATH)
        L.scpp_b200_last_error.restype = C.c_char_p
        L.scpp_b200_device_bytes.restype = C.c_size_t
        L.scpp_b200_global_active.restype = C.c_longlong
        L.scpp_b200_device_bytes.argtypes = [C.c_void_p]
        L.scpp_b200_global_active.argtypes = [C.c_void_p]
        L.scpp_b200_destroy.argtypes = [C.c_void_p]
        L.scpp_b200_destroy.restype = None
        L.scpp_b200_default_config.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ScppError(f"scpp_b200 error {rc}: {lib().scpp_b200_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().scpp_b200_device_count()


def model_dims(model):
    nx, nu, npar = C.c_int(), C.c_int(), C.c_int()
    _check(lib().scpp_b200_model_dims(model, C.byref(nx), C.byref(nu), C.byref(npar)))
    return nx.value, nu.value, npar.value


def default_config(model, **overrides):
    cfg = SCConfig()
    lib().scpp_b200_default_config(model, C.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def load_model_info(path, model):
    """ParameterServer + Parameters::loadFromFile: returns (ModelParams, x_init, x_final)"""
    nx, _, _ = model_dims(model)
    p = ModelParams(); xi = np.zeros(nx); xf = np.zeros(nx)
    _check(lib().scpp_b200_load_model_info(path.encode(), model, C.byref(p), _p(xi), _p(xf)))
    return p, xi, xf


def load_sc_info(path, model):
    cfg = default_config(model)
    _check(lib().scpp_b200_load_sc_info(path.encode(), C.byref(cfg)))
    return cfg


def load_scvx_info(path, model):
    """SCvxAlgorithm::loadParameters (SCvxAlgorithm.cpp:23-44): SCConfig with algorithm = 1"""
    cfg = default_config(model)
    _check(lib().scpp_b200_load_scvx_info(path.encode(), C.byref(cfg)))
    return cfg


def model_rows(model, params, x_init, x_final):
    """scpp_b200_model_rows: the stage-wise constraint table the engine uses, evaluated for params: (LP rows, [cone rows...]) with every row
    = (dict index -> coefficient, h) meaning s = h - sum coef * xi[index]"""
    rows = np.zeros((64, 8)); nlp = C.c_int(); nc = C.c_int(); dims = (C.c_int * 16)()
    xi = np.ascontiguousarray(x_init, float); xf = np.ascontiguousarray(x_final, float)
    _check(lib().scpp_b200_model_rows(model, C.byref(params), _p(xi), _p(xf), 64, _p(rows), C.byref(nlp), C.byref(nc), dims))
    conv = lambda r: ({int(r[1 + q]): float(r[4 + q]) for q in range(int(r[0]))}, float(r[7]))
    lp = [conv(rows[r]) for r in range(nlp.value)]
    cones, at = [], nlp.value
    for c in range(nc.value):
        cones.append([conv(rows[at + i]) for i in range(dims[c])]); at += dims[c]
    return lp, cones


def load_model(name, K=None, algorithm="SC", **overrides):
    """convenience: configs/<name>/{model,SC|SCvx}.info -> (model id, ModelParams, x_init, x_final, SCConfig); name "Rocket2DPlugin" = the
    Rocket2D files with the plugin-surface model"""
    model = {"Rocket2DPlugin": ROCKET2D_PLUGIN, "Rocket2D": ROCKET2D, "RocketQuatRoll": ROCKETQUAT_ROLL}.get(name, ROCKETQUAT)
    name = MODEL_NAMES[model] if model in (ROCKET2D_PLUGIN, ROCKETQUAT_ROLL) else name
    folder = os.path.join(CONFIG_DIR, name)
    p, xi, xf = load_model_info(os.path.join(folder, "model.info"), model)
    cfg = load_scvx_info(os.path.join(folder, "SCvx.info"), model) if algorithm == "SCvx" else load_sc_info(os.path.join(folder, "SC.info"), model)
    if K is not None:
        cfg.K = K
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return model, p, xi, xf, cfg


# ---- the reference's Monte-Carlo recipe (rocketQuat.cpp:203-227, commented out there) with a counter-based generator ----
_M64 = (1 << 64) - 1


def _splitmix64(z):
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def uniform_pm1(seed, instance, draw):
    h = _splitmix64(seed ^ _splitmix64((instance * 0x100000001B3 + draw) & _M64))
    return (h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0


def _quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def euler_to_quat_xyz(e):
    """eulerToQuaternionXYZ, scpp_models/include/common.hpp:29-38"""
    qx = np.array([np.cos(e[0] / 2), np.sin(e[0] / 2), 0, 0]); qy = np.array([np.cos(e[1] / 2), 0, np.sin(e[1] / 2), 0])
    qz = np.array([np.cos(e[2] / 2), 0, 0, np.sin(e[2] / 2)])
    return _quat_mul(_quat_mul(qx, qy), qz)


def perturbed_initial_states(x_init, rpy_init, n, seed=0x5C99, first=0):
    """RocketQuat batch: r_x,r_y *= U(-1,1); v_x,v_y *= U(-1,1); v_z *= 1+0.2U; roll,pitch = U*rpy_init; rest unchanged.
    rpy_init in radians.  Instance i of the batch uses counter `first + i` (so shards of one batch are disjoint)."""
    out = np.tile(np.asarray(x_init, float), (n, 1))
    for i in range(n):
        g = first + i
        out[i, 1] *= uniform_pm1(seed, g, 0); out[i, 2] *= uniform_pm1(seed, g, 1)
        out[i, 4] *= uniform_pm1(seed, g, 2); out[i, 5] *= uniform_pm1(seed, g, 3)
        out[i, 6] *= 1.0 + 0.2 * uniform_pm1(seed, g, 4)
        e = np.array([uniform_pm1(seed, g, 5) * rpy_init[0], uniform_pm1(seed, g, 6) * rpy_init[1], rpy_init[2]])
        out[i, 7:11] = euler_to_quat_xyz(e)
    return out


class SCAlgorithm:
    """Batched counterpart of scpp::SCAlgorithm (scpp_core/include/SCAlgorithm.hpp:9-45): constructor + initialize()
    allocate the device engine; solve(warm_start), getSolution, getAllSolutions keep their meaning for every instance."""

    def __init__(self, model, params, config, n_instances, device=0):
        self.model, self.params, self.config, self.N = model, params, config, int(n_instances)
        self.nx, self.nu, self.np_ = model_dims(model)
        self._h = C.c_void_p()
        _check(lib().scpp_b200_create(model, C.byref(params), C.byref(config), self.N, device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().scpp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_boundary_states(self, x_init, x_final):
        xi = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(x_init, float)), (self.N, self.nx)))
        xf = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(x_final, float)), (self.N, self.nx)))
        _check(lib().scpp_b200_set_boundary_states(self._h, _p(xi), _p(xf)))

    def solve(self, warm_start=False):
        _check(lib().scpp_b200_solve(self._h, int(warm_start)))

    def get_solution(self, out=None):
        K = self.config.K
        if out is None:
            out = dict(X=np.empty((self.N, K, self.nx)), U=np.empty((self.N, K, self.nu)), t=np.empty(self.N),
                       iterations=np.empty(self.N, np.int32), flags=np.empty(self.N, np.int32))
        _check(lib().scpp_b200_get_solution(self._h, _p(out["X"]), _p(out["U"]), _p(out["t"]), _p(out["iterations"]), _p(out["flags"])))
        return out

    def get_iterate_dimensional(self, it):
        """one iterate, redimensionalised like SCAlgorithm::getAllSolutions (SCAlgorithm.cpp:217-232)"""
        X = np.empty((self.N, self.config.K, self.nx)); U = np.empty((self.N, self.config.K, self.nu)); t = np.empty(self.N)
        _check(lib().scpp_b200_get_iterate_dimensional(self._h, it, _p(X), _p(U), _p(t)))
        return X, U, t

    def get_iterate(self, it):
        K = self.config.K
        X = np.empty((self.N, K, self.nx)); U = np.empty((self.N, K, self.nu)); t = np.empty(self.N)
        _check(lib().scpp_b200_get_iterate(self._h, it, _p(X), _p(U), _p(t)))
        return X, U, t

    def get_all_solutions(self):
        its = [self.get_iterate(i) for i in range(self.config.max_iterations + 1)]
        return np.stack([a[0] for a in its], 1), np.stack([a[1] for a in its], 1), np.stack([a[2] for a in its], 1)

    def get_info(self):
        info = np.empty((self.N, self.config.max_iterations, INFO_STRIDE))
        _check(lib().scpp_b200_get_info(self._h, _p(info)))
        return info

    def set_instance_params(self, params_list):
        """one ModelParams per instance (Monte-Carlo over the vehicle); None: back to the shared parameters"""
        if params_list is None:
            _check(lib().scpp_b200_set_instance_params(self._h, None))
            return
        assert len(params_list) == self.N
        arr = (ModelParams * self.N)(*params_list)
        _check(lib().scpp_b200_set_instance_params(self._h, C.byref(arr)))

    def last_timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        l, o, ii = C.c_int(), C.c_int(), C.c_longlong()
        _check(lib().scpp_b200_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(l), C.byref(o), C.byref(ii)))
        return dict(ms_discretize=a.value, ms_socp=b.value, ms_total=c.value, kernel_launches=l.value, outer_iterations=o.value,
                    instance_iterations=ii.value)

    def sim_step(self, time_step=0.05):
        """one closed-loop step (scpp/src/SC_sim.cpp:47-61): x_init <- simulate(x_init, u0, u1, time_step) on the device"""
        x = np.empty((self.N, self.nx)); u = np.empty((self.N, self.nu)); r = np.empty(self.N, np.int32)
        _check(lib().scpp_b200_sim_step(self._h, C.c_double(time_step), _p(x), _p(u), _p(r)))
        return dict(x=x, u0=u, reached=r)

    def lqr_gains(self, q_diag, r_diag):
        """LQRTracker gains for every node of every instance's current solution: (gains [N][K][nu][nx], ok [N][K])"""
        K = self.config.K
        g = np.empty((self.N, K, self.nu, self.nx)); ok = np.empty((self.N, K), np.int32)
        q = np.ascontiguousarray(q_diag, float); r = np.ascontiguousarray(r_diag, float)
        assert q.shape == (self.nx,) and r.shape == (self.nu,)
        _check(lib().scpp_b200_lqr_gains(self._h, _p(q), _p(r), _p(g), _p(ok)))
        return g, ok

    def last_rounds(self):
        r, ir = C.c_int(), C.c_longlong()
        _check(lib().scpp_b200_last_rounds(self._h, C.byref(r), C.byref(ir)))
        return dict(rounds=r.value, instance_rounds=ir.value)

    def device_bytes(self):
        return lib().scpp_b200_device_bytes(self._h)

    def comm_init(self, nranks, rank, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(lib().scpp_b200_comm_init(self._h, nranks, rank, buf))

    def global_active(self):
        return lib().scpp_b200_global_active(self._h)


class MPCConfig(C.Structure):
    """scpp_b200_mpc_config == MPC.info (MPCAlgorithm.cpp:17-32)"""
    _fields_ = [("K", C.c_int), ("nondimensionalize", C.c_int), ("constant_dynamics", C.c_int), ("intermediate_cost_active", C.c_int),
                ("time_horizon", C.c_double), ("state_weights_intermediate", C.c_double * 16), ("state_weights_terminal", C.c_double * 16),
                ("input_weights", C.c_double * 8), ("ipm", IpmSettings)]


def load_mpc_info(path, model):
    cfg = MPCConfig()
    _check(lib().scpp_b200_load_mpc_info(path.encode(), model, C.byref(cfg)))
    return cfg


class MPCAlgorithm:
    """Batched counterpart of scpp::MPCAlgorithm (scpp_core/include/MPCAlgorithm.hpp:9-98): one linear receding-horizon problem per
    instance, all sharing the model, the operating point and the weights"""

    def __init__(self, model, params, config, n_instances, device=0):
        self.model, self.params, self.config, self.N = model, params, config, int(n_instances)
        self.nx, self.nu, _ = model_dims(model)
        self._h = C.c_void_p()
        _check(lib().scpp_b200_mpc_create(model, C.byref(params), C.byref(config), self.N, device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().scpp_b200_mpc_destroy.argtypes = [C.c_void_p]
            lib().scpp_b200_mpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_states(self, x_init=None, x_final=None):
        """setInitialState / setFinalState for every instance"""
        b = lambda a: None if a is None else np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(a, float)), (self.N, self.nx)))
        xi, xf = b(x_init), b(x_final)
        _check(lib().scpp_b200_mpc_set_states(self._h, _p(xi), _p(xf)))

    def solve(self):
        _check(lib().scpp_b200_mpc_solve(self._h))

    def get_solution(self):
        K = self.config.K
        X = np.empty((self.N, K, self.nx)); U = np.empty((self.N, K - 1, self.nu)); st = np.empty(self.N, np.int32); it = np.empty(self.N, np.int32)
        _check(lib().scpp_b200_mpc_get_solution(self._h, _p(X), _p(U), _p(st), _p(it)))
        return dict(X=X, U=U, status=st, iterations=it)

    def sim_step(self, dt):
        x = np.empty((self.N, self.nx))
        _check(lib().scpp_b200_mpc_sim_step(self._h, C.c_double(dt), _p(x)))
        return x

    def discretization(self):
        A = np.empty((self.nx, self.nx)); B = np.empty((self.nx, self.nu)); z = np.empty(self.nx)
        _check(lib().scpp_b200_mpc_get_discretization(self._h, _p(A), _p(B), _p(z)))
        return A, B, z

    def last_ms(self):
        lib().scpp_b200_mpc_last_ms.restype = C.c_double
        return lib().scpp_b200_mpc_last_ms(self._h)


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().scpp_b200_comm_unique_id(buf))
    return buf.raw


def selftest_blockops(device=0):
    """largest |DMMA block product - scalar loops| over the shapes K2 uses (device self-test)"""
    err = C.c_double()
    _check(lib().scpp_b200_selftest_blockops(device, C.byref(err)))
    return err.value


def lqr_input(t, x, X, U, t_total, gains):
    """LQRTracker::getInput (scpp_core/src/LQRTracker.cpp:41-65) with TrajectoryData::inputAtTime / approxStateAtTime
    (trajectoryData.hpp:41-78), first-order hold: u = -K(t) (x - x_target(t)) + u_target(t)"""
    Kn = X.shape[0]
    t = min(max(t, 0.0), t_total)
    dt = t_total / (Kn - 1)
    if t == t_total:
        xt, ut = X[-1], U[-1]
    else:
        i = int(t / dt); a = np.fmod(t, dt) / dt
        xt = X[i] + a * (X[i + 1] - X[i]); ut = U[i] + a * (U[i + 1] - U[i])
    i = int(t / dt); a = np.fmod(t, dt) / dt
    K0 = gains[min(i, Kn - 1)]; K1 = gains[min(Kn - 1, i + 1)]
    return -(K0 + a * (K1 - K0)) @ (np.asarray(x, float) - xt) + ut


def simulate(model, x, u0, u1, par, dt, device=0):
    """K4 alone: scpp::simulate for n states (x [n][nx], u0/u1 [n][nu], par [n][np] or one row)"""
    x = np.ascontiguousarray(np.atleast_2d(x), float).copy(); n = x.shape[0]
    bc = lambda a: np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(a, float)), (n, np.atleast_2d(a).shape[1])))
    _check(lib().scpp_b200_simulate(model, n, C.c_double(dt), device, _p(x), _p(bc(u0)), _p(bc(u1)), _p(bc(par))))
    return x


def discretize(model, X, U, sigma, par, nsub=20, device=0, jacobian=1):
    """test hook on hot path 1 (discretization::multipleShooting): returns dict of [n][K-1][row][col] arrays"""
    nx, nu, npar = model_dims(model)
    X = np.ascontiguousarray(X, float); U = np.ascontiguousarray(U, float)
    if X.ndim == 2:
        X, U = X[None], U[None]
    n, K = X.shape[0], X.shape[1]
    sigma = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(np.asarray(sigma, float)), (n,)))
    par = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(np.asarray(par, float)), (n, npar)))
    A = np.empty((n, K - 1, nx, nx)); B = np.empty((n, K - 1, nu, nx)); Cc = np.empty((n, K - 1, nu, nx))
    s = np.empty((n, K - 1, nx)); z = np.empty((n, K - 1, nx))
    _check(lib().scpp_b200_discretize2(model, K, n, nsub, jacobian, device, _p(X), _p(U), _p(sigma), _p(par), _p(A), _p(B), _p(Cc), _p(s), _p(z)))
    return dict(A=A.transpose(0, 1, 3, 2), B=B.transpose(0, 1, 3, 2), C=Cc.transpose(0, 1, 3, 2), s=s, z=z)
