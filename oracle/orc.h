/*
 * oracle/orc.h — CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C double-precision restatement of the successive-convexification hot path of
 * EmbersArc/SCpp (reference tree: /root/reference, commit d45d2c8).  Every function cites
 * the reference file:line it follows.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and it cannot
 * be built here (CppAD / CppADCodeGen / Epigraph+ECOS submodules are empty, Eigen and Boost
 * are absent).  Trust in this oracle therefore comes from certificates (KKT residuals of every
 * SOCP solve, exact linearisation identities of the discretisation, finite-difference and
 * dual-number checks of the Jacobians), not from comparison with reference output.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (scpp_b200/) never links or calls it.
 */
#ifndef ORC_H
#define ORC_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODEL_ROCKETQUAT = 0, ORC_MODEL_ROCKET2D = 1 };

#define ORC_MAX_NX 14
#define ORC_MAX_NU 4
#define ORC_MAX_NP 10

/* ---- model dimensions: scpp_models/include/rocketQuatDefinitions.hpp:6-11, rocket2dDefinitions.hpp */
int orc_model_dims(int model, int *nx, int *nu, int *np);

/* ---- dynamics f(x,u,par): rocketQuat.cpp:7-37 / rocket2d.cpp:7-40 */
void orc_f(int model, const double *x, const double *u, const double *par, double *f);
/* ---- Jacobians A=df/dx (nx*nx col-major), B=df/du (nx*nu col-major): what
 *      systemDynamics.hpp:206-235 returns (exact derivatives of the flow map) */
void orc_jac(int model, const double *x, const double *u, const double *par, double *A, double *B);

/* ---- multiple shooting: discretizationImplementation.hpp:38-181 (RKF78, 5 fixed steps / interval,
 *      Phi^-1 form).  X:[K][nx], U:[K][nu] (FOH) ; outputs per interval k=0..K-2, column-major:
 *      A:[K-1][nx*nx] B:[K-1][nx*nu] C:[K-1][nx*nu] s:[K-1][nx] z:[K-1][nx].
 *      foh / free_time select the template instantiation (interpolate_input / variable time). */
void orc_discretize(int model, int K, const double *X, const double *U, double t_or_sigma,
                    const double *par, int foh, int free_time,
                    double *A, double *B, double *C, double *s, double *z);

/* ---- forward simulation: simulation.cpp:31-42 (RKF78, dt/20 => 20 steps), FOH input */
void orc_simulate(int model, double dt, const double *u0, const double *u1, const double *par,
                  double *x /* in/out */);

/* ---- generic RKF78 step count hook (tests) */
void orc_rkf78_tableau(double *c13, double *a13x13, double *b13);

/* =============================== SC problem ========================================== */

/* RocketQuat parameters as loaded from model.info (rocketQuat.cpp:234-289), angles in rad */
typedef struct {
    double g_I[3], J_B[3], r_T_B[3];
    double alpha_m;
    double T_min, T_max, t_max;
    double gimbal_max, theta_max, gamma_gs, w_B_max;
    double x_init[14], x_final[14];
    double final_time;
    int exact_minimum_thrust;
    int enable_roll_control;
    double m_scale, r_scale; /* set by nondimensionalize */
} orc_rq_params;

/* Rocket2d parameters (rocket2d.cpp:152-196) */
typedef struct {
    double g_I[2], J_B, r_T_B[2], m;
    double T_min, T_max;
    double gimbal_max, theta_max, gamma_gs, w_B_max;
    double x_init[6], x_final[6];
    double final_time;
    int constrain_initial_final;
    double m_scale, r_scale;
} orc_r2d_params;

/* SC.info (SCAlgorithm.cpp:22-46) */
typedef struct {
    int K;
    int free_final_time, interpolate_input, nondimensionalize;
    double weight_time, weight_trust_region_time, weight_trust_region_trajectory, weight_virtual_control;
    double nu_tol, delta_tol;
    int max_iterations;
} orc_sc_config;

/* per SOCP-solve certificate (relative residuals, ECOS-style) */
typedef struct {
    int status;       /* 0 optimal, 1 max iterations, 2 numerical failure */
    int iterations;
    double pres, dres, gap, relgap; /* final primal/dual residual (relative), s'z, relative gap */
    double pcost, dcost;
    double cone_viol; /* max cone violation of (s,z): 0 if strictly interior */
    double kkt_resid; /* max abs linear-solve residual seen after refinement */
} orc_ipm_info;

/* per outer-iteration record */
typedef struct {
    double norm1_nu, sum_delta, delta_sigma, sigma;
    double weight_tr_used;
    orc_ipm_info ipm;
    double t_discretize_ms, t_solve_ms;
} orc_iter_info;

/* literal SCAlgorithm::solve() (SCAlgorithm.cpp:134-189), cold start.
 * params: orc_rq_params* or orc_r2d_params* (dimensional; left unchanged on return).
 * X_all: [(max_iterations+1)][K][nx]  U_all: [(max_iterations+1)][K][nu]  t_all: [max_iterations+1]
 *        iterate 0 = initial guess; all NONDIMENSIONAL when cfg->nondimensionalize (that is what the
 *        algorithm iterates on); X_out/U_out/t_out = final trajectory REDIMENSIONALISED
 *        (SCAlgorithm.cpp:182-187).
 * returns number of iterations performed (>=1), negative on solver failure (-iteration). */
int orc_sc_solve(int model, const void *params, const orc_sc_config *cfg,
                 double *X_all, double *U_all, double *t_all, orc_iter_info *info,
                 double *X_out, double *U_out, double *t_out, int *converged);

/* SCAlgorithm::solve(warm_start): as orc_sc_solve, optionally warm-started from a DIMENSIONAL trajectory (X_warm != NULL) with the
 * trust-region weight carried in *weight_tr_io (in/out; NULL = SC.info value) */
int orc_sc_solve2(int model, const void *params, const orc_sc_config *cfg,
                  const double *X_warm, const double *U_warm, double t_warm, double *weight_tr_io,
                  double *X_all, double *U_all, double *t_all, orc_iter_info *info,
                  double *X_out, double *U_out, double *t_out, int *converged);
/* SC_sim closed loop (scpp/src/SC_sim.cpp:28-65) */
int orc_sc_sim(int model, const void *params, const orc_sc_config *cfg, double time_step, int max_steps,
               double *X_sim, double *U_sim, int *iters, int *reached_end);

/* One SOCP sub-problem (buildSCProblem SCProblem.cpp:6-138 + addApplicationConstraints
 * rocketQuat.cpp:70-144 / rocket2d.cpp:46-84) around (Xbar,Ubar,sigmabar) with given dd.
 * params must already be in the units of Xbar (i.e. nondimensional if the trajectory is).
 * thrust_dir: [K][3] linearised-min-thrust directions (RocketQuat only, may be NULL => (0,0,1)).
 * Outputs may be NULL. */
int orc_sc_subproblem(int model, const void *params, const orc_sc_config *cfg, double weight_tr,
                      const double *Xbar, const double *Ubar, double sigmabar,
                      const double *A, const double *B, const double *C, const double *s, const double *z,
                      const double *thrust_dir,
                      double *X, double *U, double *sigma, double *nu, double *delta,
                      double *norm1_nu, double *delta_sigma, orc_ipm_info *info);

/* ---- LQR tracking gains (scpp_core/src/LQR.cpp, LQRTracker.cpp:6-28); matrices row-major */
int orc_lqr_gain(int nx, int nu, const double *q_diag, const double *r_diag, const double *A, const double *B, double *K);
void orc_lqr_tracker_gains(int model, int K, const double *X, const double *U, const double *par,
                           const double *q_diag, const double *r_diag, double *gains, int *ok);

/* ---- SCvx variant (SCvx.info: SCvxAlgorithm.cpp:23-44) ---- */
typedef struct {
    int K;
    int interpolate_input, nondimensionalize;
    double rho_0, rho_1, rho_2, alpha, beta;
    double change_threshold, weight_virtual_control, trust_region;
    int max_iterations;
} orc_scvx_config;

/* per outer-iteration record (values of the LAST solve of the iteration) */
typedef struct {
    double norm1_nu, nonlinear_cost, actual_change, predicted_change, rho;
    double trust_region_used, trust_region_next;
    int solves;                 /* sub-problem solves in this iteration (1 + rejected steps) */
    int pad_;
    orc_ipm_info ipm;
} orc_scvx_info;

/* literal SCvxAlgorithm::solve() (SCvxAlgorithm.cpp:166-216) + iterate() (:61-164), cold start; same conventions as orc_sc_solve */
int orc_scvx_solve(int model, const void *params, const orc_scvx_config *cfg,
                   double *X_all, double *U_all, orc_scvx_info *info,
                   double *X_out, double *U_out, double *t_out, int *converged);
/* one SCvx sub-problem (buildSCvxProblem SCvxProblem.cpp:6-71 + addApplicationConstraints) around Ubar with trust-region radius */
int orc_scvx_subproblem(int model, const void *params, int K, double weight_vc, double trust_region,
                        const double *Ubar, const double *A, const double *B, const double *C, const double *z,
                        const double *thrust_dir, double *X, double *U, double *nu, double *norm1_nu, orc_ipm_info *info);
/* SCvxAlgorithm::getNonlinearCost (SCvxAlgorithm.cpp:262-278): sum_k |simulate(x_k; u_k, u_k+1) - x_k+1|_1 */
double orc_scvx_nonlinear_cost(int model, int K, const double *X, const double *U, double t, const double *par);

/* Export the ECOS standard form  min c'x  s.t. Ax=b, h-Gx in R+^l x Q...  of the same sub-problem
 * as COO triplets so a test can verify certificates independently (numpy).  Call with NULL arrays
 * first to get sizes.  Returns 0. */
typedef struct {
    int n, p, m, l, ncones;
    int nnzA, nnzG;
} orc_socp_dims;
int orc_sc_export(int model, const void *params, const orc_sc_config *cfg, double weight_tr,
                  const double *Xbar, const double *Ubar, double sigmabar,
                  const double *A, const double *B, const double *C, const double *s, const double *z,
                  const double *thrust_dir,
                  orc_socp_dims *dims, double *c, double *b, double *h, int *q /* cone dims */,
                  int *Ai, int *Aj, double *Av, int *Gi, int *Gj, double *Gv,
                  int *idx_X, int *idx_U, int *idx_sigma);

/* Generic conic solve of the exported standard form (used by tests on hand-made problems).
 * Triplets may contain duplicates (summed). x:[n] y:[p] s,z:[m]. */
int orc_conic_solve(int n, int p, int m, int l, int ncones, const int *q,
                    const double *c, const double *b, const double *h,
                    int nnzA, const int *Ai, const int *Aj, const double *Av,
                    int nnzG, const int *Gi, const int *Gj, const double *Gv,
                    double *x, double *y, double *s, double *z, orc_ipm_info *info);

/* helpers mirrored from the reference, exposed for tests */
void orc_rq_nondimensionalize(orc_rq_params *p);                 /* rocketQuat.cpp:291-312 */
void orc_rq_redimensionalize(orc_rq_params *p);                  /* rocketQuat.cpp:314-332 */
void orc_rq_initial_trajectory(const orc_rq_params *p, int K, double *X, double *U, double *t); /* :39-68 */
void orc_rq_model_par(const orc_rq_params *p, double *par10);    /* :168-173 */
void orc_r2d_nondimensionalize(orc_r2d_params *p);               /* rocket2d.cpp:198-214 */
void orc_r2d_initial_trajectory(const orc_r2d_params *p, int K, double *X, double *U, double *t); /* :121-136 */
void orc_r2d_model_par(const orc_r2d_params *p, double *par6);   /* :143-148 */
void orc_euler_to_quat_xyz(const double *rpy, double *q_wxyz);   /* common.hpp:29-38 */
/* the reference's (commented-out) Monte-Carlo recipe, rocketQuat.cpp:203-227, with a counter-based
 * generator so CPU and GPU build identical batches: u = 6 uniforms in [-1,1) for instance i */
void orc_rq_perturb(const orc_rq_params *nominal, const double *rpy_init, unsigned long long seed,
                    unsigned long long instance, orc_rq_params *out);
double orc_uniform_pm1(unsigned long long seed, unsigned long long instance, unsigned draw);

int orc_num_threads(void);
/* CPU-baseline driver: n_inst independent cold SC solves, one instance per OpenMP thread (the reference itself is
 * single-threaded); params_array = n_inst parameter structs, params_stride bytes apart.  Returns the total number of
 * SC iterations executed.  Output pointers may be NULL. */
#include <stddef.h>
int orc_sc_solve_batch(int model, int n_inst, const void *params_array, size_t params_stride, const orc_sc_config *cfg,
                       int *iters_out, int *conv_out, double *X_out, double *U_out, double *t_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
