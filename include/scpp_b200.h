/*
 * include/scpp_b200.h — C-ABI of libscpp_b200.so: the drop-in boundary of the B200-native batched
 * successive-convexification engine.  Plain pointers and sizes only; all arrays are HOST memory owned by the
 * caller, instance-major and C-contiguous; the opaque engine owns every device buffer.
 *
 * Each entry point names the reference interface (EmbersArc/SCpp, commit d45d2c8) it replaces.  The reference is a
 * single-instance C++ API; here every call acts on a batch of N independent problem instances that share the
 * model/algorithm parameter files and differ in their boundary states (the reference's own, commented-out,
 * Monte-Carlo recipe: scpp_models/src/rocketQuat.cpp:203-227).
 *
 * There is no CPU execution path: every compute entry point returns SCPP_B200_ERR_CUDA if no CUDA device is usable.
 */
#ifndef SCPP_B200_H
#define SCPP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCPP_B200_OK 0
#define SCPP_B200_ERR_ARG 1
#define SCPP_B200_ERR_CUDA 2
#define SCPP_B200_ERR_IO 3
#define SCPP_B200_ERR_NCCL 4
#define SCPP_B200_ERR_UNSUPPORTED 5

/* model selector: the reference picks the model at compile time, scpp_core/include/activeModel.hpp:6-10 */
#define SCPP_B200_MODEL_ROCKETQUAT 0 /* scpp_models/src/rocketQuat.cpp, nx=14 nu=4 np=10 */
#define SCPP_B200_MODEL_ROCKET2D 1   /* scpp_models/src/rocket2d.cpp,  nx=6  nu=2 np=6  */
/* the same planar rocket written ONLY against the reference's plugin surface (scpp_core/include/systemModel.hpp:64-137): generic-scalar
 * systemFlowMap, getInitializedTrajectory, (non/re)dimensionalisation and addApplicationConstraints in the cvx:: DSL (include/scpp_cvx.hpp).
 * Jacobians by forward-mode dual numbers; the constraint table is generated at build time from the recorded constraints
 * (include/scpp_plugin.hpp, tools/gen_plugin.cpp, scpp_b200/plugins/rocket2d_plugin.hpp).  Needs constrain_initial_final = true. */
#define SCPP_B200_MODEL_ROCKET2D_PLUGIN 2
/* RocketQuat written the same way WITH roll control (enable_roll_control = true, rocketQuat.cpp:135-138: |roll torque| <= t_max, w_z and the
 * torque free): scpp_b200/plugins/rocketquat_plugin.hpp.  Requires params.enable_roll_control = 1; models 0 requires it to be 0. */
#define SCPP_B200_MODEL_ROCKETQUAT_ROLL 3

/* RocketQuat::Parameters (scpp_models/include/rocketQuat.hpp:50-85) and Rocket2d::Parameters
 * (scpp_models/include/rocket2d.hpp:51-84) as loaded from model.info; angles in radians.
 * Rocket2d uses g_I[0..1], J_B[0], r_T_B[0..1], m. */
typedef struct {
    double g_I[3];
    double J_B[3];
    double r_T_B[3];
    double alpha_m;
    double m;
    double T_min, T_max, t_max;
    double gimbal_max, theta_max, gamma_gs, w_B_max;
    double final_time;
    int exact_minimum_thrust;
    int enable_roll_control;
    int constrain_initial_final;
    int pad_;
} scpp_b200_model_params;

typedef struct {
    double feastol, abstol, reltol; /* ECOS defaults: 1e-8 */
    int maxit;                      /* ECOS default: 100 */
    int stalled_step;               /* engine knob, with warm > 0 only: the sub-problem after a stalled outer iteration steps 1 - 10^-stalled_step of the way
                                       to the cone boundary instead of 0.99 (0 = 0.99 everywhere, the default).  3 was measured to shorten the AVERAGE
                                       solve and to lengthen the slowest one of a batch, which is what a round waits for (profiles/README.md) */
    double warm;                    /* 0 = cold start of every sub-problem (what ECOS does); 0<warm<1: from the second outer iteration on,
                                       start from the instance's previous interior solution pulled back from the cone boundary,
                                       (s,z) <- warm*(s,z) + (1-warm)*e (same optimum, ~2.8x fewer interior-point iterations); the sub-problem after
                                       a stalled outer iteration (sum of the trust-region radii <= delta_tol) starts 50x closer, 1-w' = (1-warm)/50 */
} scpp_b200_ipm_settings;

/* SC.info as read by SCAlgorithm::loadParameters (scpp_core/src/SCAlgorithm.cpp:22-46) + engine knobs */
typedef struct {
    int K;
    int free_final_time, interpolate_input, nondimensionalize;   /* free_final_time = 0: sigma is not a variable (SCProblem.cpp:27-35), the final time stays
                                                                     model.info's final_time; interpolate_input = 0 (zero-order hold, discretizationImplementation.hpp:41-50,
                                                                     SCProblem.cpp:49-56,114-121): SC on RocketQuat / Rocket2D; the reference's U then has K - 1 columns, here
                                                                     column K - 1 of every returned U is a pinned placeholder that is in no dynamics row (sc.cuh: sc_zoh_pins);
                                                                     SCvx or a plugin model with it returns UNSUPPORTED */
    double weight_time, weight_trust_region_time, weight_trust_region_trajectory, weight_virtual_control;
    double nu_tol, delta_tol;
    int max_iterations;
    int nsub;         /* RK4 sub-steps per shooting interval (reference: RKF78 x 5, discretizationImplementation.hpp:154);
                         nsub < 0: RK4 with -nsub and with -2 nsub sub-steps, Richardson-extrapolated (default -5) */
    int keep_history; /* keep every iterate for scpp_b200_get_iterate (SCAlgorithm::getAllSolutions) */
    int ipm_slice;    /* engine knob: interior-point iterations per K2 launch (default 1).  Between launches the engine re-forms
                         the batch, so an instance that needs 20 iterations does not hold back one that needs 5; 0 = run each
                         sub-problem to the end in one launch (lock-step outer iterations).  Same arithmetic either way.
                         -1 = split pipeline (one kernel per step of the interior-point iteration, several warps per instance);
                         -T (T > 1) = split pipeline only in rounds that advance fewer than T instances.  Same iterates up to the
                         rounding of re-ordered sums. */
    scpp_b200_ipm_settings ipm;
    /* SCvx variant (scpp_core/src/SCvxAlgorithm.cpp, SCvxProblem.cpp; parameters of SCvx.info, SCvxAlgorithm.cpp:23-44).
     * algorithm = 0: SC (the fields above), 1: SCvx -- fixed final time, hard trust region of radius scvx_trust_region on the inputs,
     * ratio test (rho_0/1/2, alpha, beta) against the simulated nonlinear cost, converged when |predicted change| < change_threshold.
     * For SCvx, scpp_b200_get_info returns per outer iteration: norm1_nu, nonlinear cost, rho, trust region used, sub-problem solves,
     * ipm_iterations, ipm_status, pres, dres, relgap. */
    int algorithm;
    int solver;       /* K2 mapping.  0 (default): one warp per instance, one interior-point iteration per launch, the batch re-formed in between.
                         1: one CTA per instance, whole sub-problem per launch, factor in shared memory (K <= ~120).  2: as 0, and in the TAIL of a
                         solve (fewer unfinished instances than resident CTAs) the sub-problems that start run on mapping 1 -- results then agree
                         with mapping 0 to solver accuracy, not bit for bit, and depend on the batch composition */
    double scvx_rho_0, scvx_rho_1, scvx_rho_2, scvx_alpha, scvx_beta, scvx_change_threshold, scvx_trust_region;
    int jacobian;     /* how K1 obtains the Jacobian products (computeJacobians, systemDynamics.hpp:206-235): 1 (default) forward-mode dual
                         numbers over the model's generic-scalar flow map -- what CppAD gives the reference, nothing model-specific beyond
                         systemFlowMap; 0 the hand-derived sparse Jacobian of models.cuh (optional fast path, tested equal); 2 the hand-derived
                         Jacobian evaluated once per right-hand side and shared by the columns of an interval (discretize_shared.cuh: the
                         fastest K1; models without a hand-derived Jacobian take path 1) */
    int pad3_;
} scpp_b200_sc_config;

typedef struct scpp_b200_engine scpp_b200_engine;

/* per (instance, iteration) record returned by scpp_b200_get_info: what SCAlgorithm::iterate prints
 * (SCAlgorithm.cpp:117-128) plus the solver certificate */
#define SCPP_B200_INFO_STRIDE 10 /* norm1_nu, sum_delta, delta_sigma, sigma, weight_tr_used, ipm_iterations, ipm_status, pres, dres, relgap */

int scpp_b200_version(void);
const char *scpp_b200_last_error(void);
int scpp_b200_device_count(void); /* 0 if no usable CUDA device */
int scpp_b200_model_dims(int model, int *nx, int *nu, int *np);
/* the stage-wise constraint table the engine uses for `model`, evaluated for `params` (dimensional, nondimensionalize = 0): one record of
 * 8 doubles per row {n, idx0, idx1, idx2, coef0, coef1, coef2, h} meaning  s = h - sum_j coef_j * xi[idx_j]  (xi = [x ; u] of a node; a
 * coefficient taken from the per-node minimum-thrust direction is reported as NaN).  rows receives n_lp LP rows, then the cone rows;
 * cone_dims [n_cones].  What a maintainer compares against a model's addApplicationConstraints (tests/cvx_shim_test.cpp does). */
int scpp_b200_model_rows(int model, const scpp_b200_model_params *params, const double *x_init, const double *x_final, int max_rows,
                         double *rows, int *n_lp, int *n_cones, int *cone_dims);
void scpp_b200_default_config(int model, scpp_b200_sc_config *cfg); /* values of scpp_models/config/<Model>/SC.info */

/* ---- parameter files (host logic; usable without a GPU) -------------------------------------------------------
 * replaces ParameterServer (scpp_core/utils/include/parameterServer.hpp:34-127: Boost INFO files) and
 * RocketQuat::Parameters::loadFromFile (rocketQuat.cpp:234-289) / Rocket2d::Parameters::loadFromFile (rocket2d.cpp:150-196).
 * x_init / x_final receive the boundary states built there (nx doubles each). */
int scpp_b200_load_model_info(const char *path, int model, scpp_b200_model_params *params, double *x_init, double *x_final);
int scpp_b200_load_sc_info(const char *path, scpp_b200_sc_config *cfg); /* SCAlgorithm::loadParameters */
int scpp_b200_load_scvx_info(const char *path, scpp_b200_sc_config *cfg); /* SCvxAlgorithm::loadParameters (SCvxAlgorithm.cpp:23-44); sets algorithm = 1 */

/* ---- engine life cycle ------------------------------------------------------------------------------------------
 * replaces SCAlgorithm::SCAlgorithm(Model::ptr_t) + SCAlgorithm::initialize() (SCAlgorithm.cpp:14-20,48-64):
 * allocates all device state for n_instances problem instances on CUDA device `device`. */
int scpp_b200_create(int model, const scpp_b200_model_params *params, const scpp_b200_sc_config *cfg,
                     int n_instances, int device, scpp_b200_engine **out);
void scpp_b200_destroy(scpp_b200_engine *e);

/* boundary states of every instance, dimensional: x_init [N][nx], x_final [N][nx] (model->p.x_init / x_final,
 * mutated by callers such as scpp/src/SC_sim.cpp:36).  Host -> device copy. */
int scpp_b200_set_boundary_states(scpp_b200_engine *e, const double *x_init, const double *x_final);

/* SCAlgorithm::solve(bool warm_start) (SCAlgorithm.cpp:134-189): the whole outer loop on the device —
 * per iteration K1 multiple shooting (discretization::multipleShooting, discretization.cpp:42-55),
 * K2 SOCP solve (ECOSSolver::solve, SCAlgorithm.cpp:78) fused with readSolution and the convergence logic
 * (SCAlgorithm.cpp:100-131).  A failed instance is flagged, never aborts the batch (reference: std::terminate, :94-98). */
/* per-instance model parameters (optional): Pn [N] in the layout of scpp_b200_model_params, i.e. one RocketQuat::Parameters /
 * Rocket2d::Parameters (scpp_models/include/rocketQuat.hpp:50-85) per instance, so a Monte-Carlo batch can vary the vehicle (inertia, I_sp,
 * thrust limits, constraint angles ...) and not only the boundary states.  NULL: every instance uses the parameters given at creation.
 * Takes effect at the next scpp_b200_solve. */
int scpp_b200_set_instance_params(scpp_b200_engine *e, const scpp_b200_model_params *Pn);
int scpp_b200_solve(scpp_b200_engine *e, int warm_start);

/* SCAlgorithm::getSolution (SCAlgorithm.cpp:212-215): final trajectories REDIMENSIONALISED (:182-187).
 * X [N][K][nx], U [N][K][nu], t [N]; iterations [N]; flags [N]: 0 not converged, 1 converged, 2 solver failure.
 * Any pointer may be NULL.  Device -> host copy. */
int scpp_b200_get_solution(scpp_b200_engine *e, double *X, double *U, double *t, int *iterations, int *flags);

/* SCAlgorithm::getAllSolutions (SCAlgorithm.cpp:217-232): iterate `it` (0 = initial guess) of every instance in the
 * units the algorithm iterates on (nondimensional when cfg.nondimensionalize).  Needs cfg.keep_history. */
int scpp_b200_get_iterate(scpp_b200_engine *e, int it, double *X, double *U, double *t);
/* the same iterate redimensionalised, as SCAlgorithm::getAllSolutions returns it (SCAlgorithm.cpp:217-232: model->redimensionalizeTrajectory
 * on a copy of every iterate); scpp_b200_get_iterate returns the units the algorithm iterates on (what the parity tests compare) */
int scpp_b200_get_iterate_dimensional(scpp_b200_engine *e, int it, double *X, double *U, double *t);
int scpp_b200_get_info(scpp_b200_engine *e, double *info /* [N][max_iterations][SCPP_B200_INFO_STRIDE] */);

/* One step of the closed loop of scpp/src/SC_sim.cpp:47-61 for every instance (kernel K4): the first input of the current solution
 * (u0 = U[0], u1 = interpolatedInput(U, time_step, t), scpp/src/commonFunctions.cpp:6-19) is applied to the nonlinear model for
 * time_step seconds (scpp::simulate, scpp_core/src/simulation.cpp:31-42: RKF78, time_step/20), x_init <- simulated state ON THE DEVICE.
 * The next scpp_b200_solve(e, 1) continues from it.  Outputs (host, optional): x_new [N][nx], u0 [N][nu] (dimensional),
 * reached [N] = |x - x_final| < 0.02 or t < 0.25 (SC_sim.cpp:58); an instance that reached the end is frozen: later solves and
 * steps skip it (flag 8) until scpp_b200_set_boundary_states is called again. */
int scpp_b200_sim_step(scpp_b200_engine *e, double time_step, double *x_new, double *u0, int *reached);

/* LQR tracking gains along the current solution of every instance (kernel K5): LQRTracker::LQRTracker scpp_core/src/LQRTracker.cpp:6-28 with
 * ComputeLQR / careSolve / solveSchurIterative scpp_core/src/LQR.cpp:7-109 (matrix sign iteration on the Hamiltonian, eps 1e-8, <= 100
 * steps; FullPivLU solve).  q_diag [nx], r_diag [nu] = state_weights / input_weights of LQR.info (LQRTracker.cpp:30-40).
 * gains [N][K][nu][nx] (row-major, dimensional units), ok [N][K] (optional) = success flag of careSolve. */
int scpp_b200_lqr_gains(scpp_b200_engine *e, const double *q_diag, const double *r_diag, double *gains, int *ok);

/* device timing of the last solve (CUDA events on the engine stream): ms in K1, ms in K2, ms total, kernel launches,
 * outer iterations executed, sum over instances of iterations executed */
int scpp_b200_last_timing(scpp_b200_engine *e, double *ms_discretize, double *ms_socp, double *ms_total,
                          int *kernel_launches, int *outer_iterations, long long *instance_iterations);
/* rounds of the last solve (one K2 launch, or one split-pipeline kernel sequence, each) and the sum over the rounds of the
 * instances they advanced */
int scpp_b200_last_rounds(scpp_b200_engine *e, int *rounds, long long *instance_rounds);
size_t scpp_b200_device_bytes(scpp_b200_engine *e);

/* ---- test hooks on the two hot paths ----------------------------------------------------------------------------
 * K1 alone: discretization::multipleShooting for n trajectories.  X [n][K][nx], U [n][K][nu], sigma [n], par [n][np];
 * outputs in the reference's layout (column-major Eigen blocks per interval, DiscretizationData
 * scpp_core/include/discretizationData.hpp:8-20): A [n][K-1][nx*nx], B,C [n][K-1][nx*nu], s,z [n][K-1][nx]. */
int scpp_b200_discretize(int model, int K, int n, int nsub, int device, const double *X, const double *U, const double *sigma,
                         const double *par, double *A, double *B, double *C, double *s, double *z);
/* the same with the Jacobian path chosen explicitly (scpp_b200_sc_config.jacobian: 1 dual numbers over the flow map, 0 hand-derived) */
int scpp_b200_discretize2(int model, int K, int n, int nsub, int jacobian, int device, const double *X, const double *U, const double *sigma,
                          const double *par, double *A, double *B, double *C, double *s, double *z);

/* the tensor-core block products of K2 (mma.sync.m8n8k4.f64, scpp_b200/csrc/blockops.cuh) checked on the device against scalar
 * loops for the shapes the factorisation uses; returns the largest absolute deviation (no reference counterpart) */
/* K4 alone: scpp::simulate (simulation.cpp:31-42) for n states.  x [n][nx] in/out, u0, u1 [n][nu], par [n][np] */
int scpp_b200_simulate(int model, int n, double dt, int device, double *x, const double *u0, const double *u1, const double *par);

int scpp_b200_selftest_blockops(int device, double *max_abs_err);

/* ---- multi-GPU: one process (rank) per GPU, the batch is sharded by the caller ----------------------------------
 * the only data-path collective is one ncclAllGather of the per-instance convergence flags per outer iteration.
 * Shards may differ in size (N_total % nranks != 0): comm_init all-gathers the shard sizes and every rank contributes max(N) flag bytes,
 * its padding marked 'done'.  comm_init is collective: every rank must call it. */
int scpp_b200_comm_unique_id(char id[128]);
int scpp_b200_comm_init(scpp_b200_engine *e, int nranks, int rank, const char id[128]);
long long scpp_b200_global_active(scpp_b200_engine *e); /* instances still iterating over all ranks after the last solve */

/* ---- receding-horizon MPC ------------------------------------------------------------------------------------
 * replaces MPCAlgorithm (scpp_core/include/MPCAlgorithm.hpp:9-98, src/MPCAlgorithm.cpp:11-140), buildMPCProblem (src/MPCProblem.cpp:6-87)
 * and exactLinearDiscretization (src/discretization.cpp:9-40) for a Monte-Carlo batch: N instances that share the model, the operating
 * point and the weights and differ in x_init / x_final.  The linear time-invariant dynamics are eliminated once on the host; kernel K6
 * solves the dense conic program of every instance (one thread each).  Models: Rocket2D (the only reference model with getOperatingPoint
 * and an MPC.info).  Not built: intermediate_cost_active, constant_dynamics = false, nondimensionalize (UNSUPPORTED); the state rows of
 * addApplicationConstraints act on nodes 1..K-1 (node 0 is the given state). */
typedef struct {
    int K;
    int nondimensionalize, constant_dynamics, intermediate_cost_active;
    double time_horizon;
    double state_weights_intermediate[16], state_weights_terminal[16], input_weights[8];
    scpp_b200_ipm_settings ipm;
} scpp_b200_mpc_config;
typedef struct scpp_b200_mpc scpp_b200_mpc;
int scpp_b200_load_mpc_info(const char *path, int model, scpp_b200_mpc_config *cfg);                  /* MPCAlgorithm::loadParameters */
int scpp_b200_mpc_create(int model, const scpp_b200_model_params *params, const scpp_b200_mpc_config *cfg, int n_instances, int device,
                         scpp_b200_mpc **out);                                                         /* MPCAlgorithm ctor + initialize() */
void scpp_b200_mpc_destroy(scpp_b200_mpc *m);
int scpp_b200_mpc_set_states(scpp_b200_mpc *m, const double *x_init, const double *x_final);           /* setInitialState / setFinalState; [N][nx], NULL keeps */
int scpp_b200_mpc_solve(scpp_b200_mpc *m);                                                             /* MPCAlgorithm::solve for every instance */
/* getSolution: X [N][K][nx], U [N][K-1][nu]; status [N]: 0 optimal, 3 reduced accuracy, 1/2 failed; iters [N] interior-point iterations */
int scpp_b200_mpc_get_solution(scpp_b200_mpc *m, double *X, double *U, int *status, int *iters);
/* one step of scpp/src/MPC_sim.cpp:64-70 on the device: simulate(model, dt, u, u, x) with u = U[0], x_init <- x; x_new [N][nx] or NULL */
int scpp_b200_mpc_sim_step(scpp_b200_mpc *m, double dt, double *x_new);
/* test hook: A [nx][nx], B [nx][nu] (row-major), z [nx] of exactLinearDiscretization at the operating point */
int scpp_b200_mpc_get_discretization(scpp_b200_mpc *m, double *A, double *B, double *z);
double scpp_b200_mpc_last_ms(scpp_b200_mpc *m);                                                        /* device time of the last solve */

#ifdef __cplusplus
}
#endif
#endif
