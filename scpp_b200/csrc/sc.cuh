// scpp_b200/csrc/sc.cuh — per-instance successive-convexification state and the glue around the two hot paths
// (kernel K0 = setup / initial guess, K3 = read solution + convergence logic; K3 is fused into the K2 epilogue).
//
// Reference: scpp_core/src/SCAlgorithm.cpp:66-210 (iterate / solve / readSolution), scpp_models (setup).
#pragma once
#include "ipm.cuh"
#include "discretize.cuh"
#include "simulate.cuh"

namespace scpp {

// SC.info, scpp_core/src/SCAlgorithm.cpp:22-46, plus engine knobs
struct ScConfig {
    int K;
    int free_final_time, interpolate_input, nondimensionalize;
    double weight_time, weight_trust_region_time, weight_trust_region_trajectory, weight_virtual_control;
    double nu_tol, delta_tol;
    int max_iterations;
    int nsub;                 // RK4 sub-steps per shooting interval (K1)
    int keep_history;         // store every iterate (getAllSolutions)
    int ipm_slice;            // interior-point iterations per K2 launch (0: run every sub-problem to the end in one launch)
    IpmSettings ipm;
    // SCvx variant (SCvx.info, scpp_core/src/SCvxAlgorithm.cpp:23-44); algorithm == 0: SC (everything above), 1: SCvx
    int algorithm;
    int solver;               // K2 mapping.  0: one warp per instance, one interior-point slice per launch (round 1).  1: one CTA per instance,
                              // whole sub-problem per launch, factor in shared memory, persistent CTAs pulling from a queue (round 2, ipm_cta.inl)
    double scvx_rho_0, scvx_rho_1, scvx_rho_2, scvx_alpha, scvx_beta, scvx_change_threshold, scvx_trust_region;
    int jacobian;             // K1: 1 = forward-mode dual numbers over the model's flow map (default), 0 = the hand-derived sparse Jacobian (models.cuh),
                              // 2 = hand-derived, linearisation shared by the columns of an interval (discretize_shared.cuh)
    int pad3_;
};
constexpr int SCVX_MAX_RESOLVE = 40;   // the reference's re-solve loop of a rejected step has no bound; the engine fails the instance after this many

constexpr int INFO_STRIDE = 10;   // per (instance, iteration): norm1_nu, sum_delta, delta_sigma, sigma, w_tr_used,
                                  //                            ipm_iterations, ipm_status, pres, dres, relgap

// batch-wide device arrays, instance-major
template <class M>
struct ScArrays {
    static constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2;
    int N, K, max_it;
    double *x_init, *x_final;      // [N][NX] as uploaded (dimensional)
    double *xi, *xf;               // [N][NX] scaled
    const ModelParamsHost *Pn;     // [N] per-instance model parameters (Monte-Carlo over the vehicle), or null: every instance uses the engine's
    double *par;                   // [N][NP]
    double *cst;                   // [N][MAX_CST]
    double *scale;                 // [N][2]
    double *X, *U, *sigma;         // [N][K][NX], [N][K][NU], [N]
    double *tdir;                  // [N][K][3]
    uint32_t *fixm;                // [N][K]
    double *fixv;                  // [N][K][NB]
    double *w_tr;                  // [N] current trust-region weight
    int *iters, *status, *converged;   // [N]
    double *dd;                    // [N][K-1][NX][NC]
    double *ddT;                   // [N][NX*NC][KS]  the same tiles, stage-minor
    double *ws;                    // [N][ws_doubles]
    double *ipm_state;             // [N][Ipm::IPM_STATE]  solver state parked between K2 launches ([0] != 0: mid-solve)
    int *frozen;                   // [N] or null: instances whose closed loop has reached the end (skipped by solve and sim_step)
    // SCvx state (null for SC): trust-region radius, last nonlinear cost (+ flag), solves of the running outer iteration, 1 = a solved
    // sub-problem waits for its nonlinear cost and the ratio test, candidate iterate, its norm1_nu, per-interval simulated defects
    double *trust, *last_cost, *n1c, *Xc, *Uc, *costp;
    int *have_last, *solves, *phase;
    double *hist;                  // [N][max_it+1][K*NB+1] or null
    double *info;                  // [N][max_it][INFO_STRIDE]
    size_t ws_stride;

    SCPP_HD size_t hist_stride() const { return (size_t)K * NB + 1; }
};

// ---- zero-order-hold inputs (interpolate_input = false; discretizationImplementation.hpp:41-50, SCProblem.cpp:49-56,114-121) --------
// The reference then has K - 1 input columns: u_k acts on interval k alone (no C_k), the model's "final input" constraints bind column
// K - 2 (v_U.col(v_U.cols() - 1), rocketQuat.cpp:109-111) and the trust region of node K - 1 has no input part.  The engine keeps its K-column
// layout: K1 holds the input over the interval (C_k == 0 exactly, B_k is the whole input matrix), the final-input pins of the model move to
// node K - 2, and column K - 1 becomes a placeholder -- pinned to a strictly feasible constant, so it is no variable, its constraint rows are
// constants and its trust-region term is identically zero.  What remains is the reference's problem.
template <class M>
SCPP_HD void sc_zoh_pins(const ScArrays<M> &a, const ScConfig &cfg, int n)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    if (cfg.interpolate_input) return;
    const int K = a.K;
    uint32_t *fm = a.fixm + (size_t)n * K;
    double *fv = a.fixv + (size_t)n * K * NB, *U = a.U + (size_t)n * K * NU;
    const uint32_t in_bits = ((1u << NU) - 1u) << NX;
    fm[K - 2] |= fm[K - 1] & in_bits;
    for (int j = 0; j < NU; j++) if (fm[K - 1] >> (NX + j) & 1u) fv[(K - 2) * NB + NX + j] = fv[(K - 1) * NB + NX + j];
    double ph[NU];
    M::zoh_placeholder_input(a.cst + (size_t)n * MAX_CST, ph);
    fm[K - 1] |= in_bits;
    for (int j = 0; j < NU; j++) { fv[(K - 1) * NB + NX + j] = ph[j]; U[(K - 1) * NU + j] = ph[j]; }
}

// ---- K0: nondimensionalise, model parameters, constraint constants, pinned variables, initial guess -------------
// SCAlgorithm::solve cold start, SCAlgorithm.cpp:138-159
template <class M>
SCPP_HD void sc_setup_instance(const ScArrays<M> &a, const ModelParamsHost &P, const ScConfig &cfg, int n)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    const int K = a.K;
    double *xi = a.xi + (size_t)n * NX, *xf = a.xf + (size_t)n * NX;
    for (int i = 0; i < NX; i++) { xi[i] = a.x_init[(size_t)n * NX + i]; xf[i] = a.x_final[(size_t)n * NX + i]; }
    double *cst = a.cst + (size_t)n * MAX_CST;
    M::setup(P, cfg.nondimensionalize, xi, xf, a.par + (size_t)n * M::NP, cst, a.scale + (size_t)n * 2);
    double *X = a.X + (size_t)n * K * NX, *U = a.U + (size_t)n * K * NU;
    for (int k = 0; k < K; k++) {
        M::initial_guess(xi, xf, cst, K, k, X + k * NX, U + k * NU);
        a.fixm[(size_t)n * K + k] = M::fixed(P, xi, xf, K, k, a.fixv + ((size_t)n * K + k) * NB);
        double *td = a.tdir + ((size_t)n * K + k) * 3;
        if (P.exact_minimum_thrust) M::thrust_dir(U + k * NU, td);            // rocketQuat.cpp:162-165
        else { td[0] = 0.; td[1] = 0.; td[2] = 1.; }
    }
    sc_zoh_pins<M>(a, cfg, n);
    if (!cfg.interpolate_input && P.exact_minimum_thrust) M::thrust_dir(U + (K - 1) * NU, a.tdir + ((size_t)n * K + K - 1) * 3);
    a.sigma[n] = P.final_time;
    a.w_tr[n] = cfg.weight_trust_region_trajectory;                            // loadParameters(), SCAlgorithm.cpp:148
    a.iters[n] = 0; a.status[n] = 0; a.converged[n] = 0;
    a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE] = 0.; a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE + Ipm<M>::ST_NOPOINT] = 0.;
    if (cfg.algorithm == 1) { a.trust[n] = cfg.scvx_trust_region; a.have_last[n] = 0; a.last_cost[n] = 0.; a.solves[n] = 0; a.phase[n] = 0; }   // loadParameters(), SCvxAlgorithm.cpp:182
    if (a.hist) {
        double *h = a.hist + (size_t)n * (a.max_it + 1) * a.hist_stride();
        for (int k = 0; k < K; k++) { for (int i = 0; i < NX; i++) h[k * NB + i] = X[k * NX + i]; for (int i = 0; i < NU; i++) h[k * NB + NX + i] = U[k * NU + i]; }
        h[K * NB] = a.sigma[n];
    }
}

// ---- K0 (warm): SCAlgorithm::solve(warm_start = true), SCAlgorithm.cpp:141-145,152: keep the trajectory (re-nondimensionalised with
// the scales of the NEW x_init), keep the trust-region weight and sigma, refresh parameters and the minimum-thrust directions
template <class M>
SCPP_HD void sc_warm_instance(const ScArrays<M> &a, const ModelParamsHost &P, const ScConfig &cfg, int n)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    const int K = a.K;
    double *X = a.X + (size_t)n * K * NX, *U = a.U + (size_t)n * K * NU;
    double old_scale[2] = {a.scale[2 * n], a.scale[2 * n + 1]};
    for (int k = 0; k < K; k++) M::redim(old_scale, X + k * NX, U + k * NU);
    double *xi = a.xi + (size_t)n * NX, *xf = a.xf + (size_t)n * NX;
    for (int i = 0; i < NX; i++) { xi[i] = a.x_init[(size_t)n * NX + i]; xf[i] = a.x_final[(size_t)n * NX + i]; }
    M::setup(P, cfg.nondimensionalize, xi, xf, a.par + (size_t)n * M::NP, a.cst + (size_t)n * MAX_CST, a.scale + (size_t)n * 2);
    for (int k = 0; k < K; k++) {
        M::nondim(a.scale + 2 * n, X + k * NX, U + k * NU);
        a.fixm[(size_t)n * K + k] = M::fixed(P, xi, xf, K, k, a.fixv + ((size_t)n * K + k) * NB);
        double *td = a.tdir + ((size_t)n * K + k) * 3;
        if (P.exact_minimum_thrust) M::thrust_dir(U + k * NU, td); else { td[0] = 0.; td[1] = 0.; td[2] = 1.; }
    }
    sc_zoh_pins<M>(a, cfg, n);
    if (!cfg.interpolate_input && P.exact_minimum_thrust) M::thrust_dir(U + (K - 1) * NU, a.tdir + ((size_t)n * K + K - 1) * 3);
    a.iters[n] = 0; a.status[n] = 0; a.converged[n] = 0;
    a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE] = 0.; a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE + Ipm<M>::ST_NOPOINT] = 0.;
    if (cfg.algorithm == 1) { a.solves[n] = 0; a.phase[n] = 0; }   // the radius and last_nonlinear_cost are members the reference does not reset on a warm start
    if (a.hist) {
        double *h = a.hist + (size_t)n * (a.max_it + 1) * a.hist_stride();
        for (int k = 0; k < K; k++) { for (int i = 0; i < NX; i++) h[k * NB + i] = X[k * NX + i]; for (int i = 0; i < NU; i++) h[k * NB + NX + i] = U[k * NU + i]; }
        h[K * NB] = a.sigma[n];
    }
}

// ---- K4: one step of the closed loop of scpp/src/SC_sim.cpp:47-61 for instance n: apply the first input of the current solution to
// the nonlinear model for time_step (dimensional units, model parameters redimensionalised as after SCAlgorithm.cpp:182-186),
// x_init <- simulated state.  u1 = scpp::interpolatedInput(td.U, time_step, td.t, foh) (scpp/src/commonFunctions.cpp:6-19).
template <class M>
SCPP_HD void sc_sim_step_instance(const ScArrays<M> &a, const ModelParamsHost &P, const ScConfig &cfg, int n, double time_step,
                                  double *x_out /* [NX] or null */, double *u_out /* [NU] or null */, int *reached /* or null */)
{
    constexpr int NX = M::NX, NU = M::NU;
    const int K = a.K;
    if (a.frozen && a.frozen[n]) {
        if (x_out) for (int i = 0; i < NX; i++) x_out[i] = a.x_init[(size_t)n * NX + i];
        if (u_out) for (int j = 0; j < NU; j++) u_out[j] = 0.;
        if (reached) *reached = 1;
        return;
    }
    const double *U = a.U + (size_t)n * K * NU, *scale = a.scale + 2 * n;
    const double t = a.sigma[n];
    const double node_dt = t / (K - 1);
    long long i = (long long)(time_step / node_dt);
    if (i > K - 2) i = K - 2;
    const double ti = fmod(time_step, node_dt) / node_dt;
    double xd[NX], u0[NU], ua[NU], ub[NU], u1[NU];
    for (int j = 0; j < NU; j++) { u0[j] = U[j]; ua[j] = U[NU * i + j]; ub[j] = cfg.interpolate_input ? U[NU * (i + 1) + j] : U[NU * i + j]; }
    for (int e = 0; e < NX; e++) xd[e] = 0.;
    M::redim(scale, xd, u0); M::redim(scale, xd, ua); M::redim(scale, xd, ub);      // redimensionalizeTrajectory (SCAlgorithm.cpp:186)
    for (int j = 0; j < NU; j++) u1[j] = ua[j] + (ub[j] - ua[j]) * ti;
    // dimensional model parameters (model->redimensionalize(); updateModelParameters(), :184-185)
    double xi[NX], xf[NX], par[M::NP], cst[MAX_CST], sc2[2];
    for (int e = 0; e < NX; e++) { xi[e] = a.x_init[(size_t)n * NX + e]; xf[e] = a.x_final[(size_t)n * NX + e]; }
    M::setup(P, 0, xi, xf, par, cst, sc2);
    double x[NX];
    for (int e = 0; e < NX; e++) x[e] = a.x_init[(size_t)n * NX + e];
    rkf78_simulate<M>(x, u0, u1, par, time_step, 20);                               // simulate(model, time_step, u0, u1, x)  SC_sim.cpp:53
    double d2 = 0.;
    for (int e = 0; e < NX; e++) { a.x_init[(size_t)n * NX + e] = x[e]; const double d = x[e] - a.x_final[(size_t)n * NX + e]; d2 += d * d; }
    const int end = (sqrt(d2) < 0.02) || (t < 0.25);                                // :58
    if (a.frozen && end) a.frozen[n] = 1;
    if (x_out) for (int e = 0; e < NX; e++) x_out[e] = x[e];
    if (u_out) for (int j = 0; j < NU; j++) u_out[j] = u0[j];
    if (reached) *reached = end;
}

// ---- binding of the solver object to instance n ------------------------------------------------------------------
template <class M>
SCPP_HD void sc_bind(const ScArrays<M> &a, const ScConfig &cfg, int n, double *smem, Ipm<M> &ipm, bool cta = false)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2;
    const int K = a.K;
    ipm.K = K;
    ipm.dd = a.dd + (size_t)n * (K - 1) * NX * NC;
    ipm.ddT = a.ddT + (size_t)n * Ipm<M>::ddt_doubles(K);
    ipm.Xbar = a.X + (size_t)n * K * NX; ipm.Ubar = a.U + (size_t)n * K * NU; ipm.sigbar = a.sigma[n];
    ipm.cst = a.cst + (size_t)n * MAX_CST;
    ipm.tdir = a.tdir + (size_t)n * K * 3;
    ipm.fixm = a.fixm + (size_t)n * K;
    ipm.fixv = a.fixv + (size_t)n * K * NB;
    ipm.w_time = cfg.weight_time; ipm.w_trs = cfg.weight_trust_region_time; ipm.w_vc = cfg.weight_virtual_control;
    ipm.w_tr = a.w_tr[n];
    if (cfg.algorithm == 1) { ipm.scvx = true; ipm.sig_fixed = true; ipm.tr_rad = a.trust[n]; ipm.w_time = 0.; ipm.w_trs = 0.; ipm.w_tr = 0.; }
    else if (!cfg.free_final_time) { ipm.sig_fixed = true; ipm.w_time = 0.; ipm.w_trs = 0.; }      // SC with a fixed final time: no sigma / delta_sigma (SCProblem.cpp:27-35,82-100)
    if (cta) ipm.bind_cta(a.ws + (size_t)n * a.ws_stride, smem); else ipm.bind(a.ws + (size_t)n * a.ws_stride, smem);
}

// ---- K3: readSolution and the convergence logic of SCAlgorithm::iterate (SCAlgorithm.cpp:100-131, 191-210) for a solved sub-problem
template <class M>
SCPP_HD void sc_finish_instance(const ScArrays<M> &a, const ScConfig &cfg, int n, const Ipm<M> &ipm, const IpmResult &r)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    const int K = a.K, it = a.iters[n];
    double *X = a.X + (size_t)n * K * NX, *U = a.U + (size_t)n * K * NU;
    const double w_tr = ipm.w_tr;
    double *inf = a.info + ((size_t)n * a.max_it + it) * INFO_STRIDE;
    const bool ok = (r.status == 0 || r.status == 3);
    warp_sync();
    double n1 = 0, sd = 0;
    if (ok) {
        // readSolution (SCAlgorithm.cpp:191-210): X, U, sigma <- solver variables
        FOR_LANE(e, K * NB) { const int k = e / NB, i = e - k * NB; const double v = ipm.xi_at(k, i); if (i < NX) X[k * NX + i] = v; else U[k * NU + (i - NX)] = v; }
        FOR_LANE(e, (K - 1) * NX) n1 += ipm.t_at(e / NX, e % NX);            // norm1_nu  (:102-103)
        FOR_LANE(k, K) sd += ipm.delta_at(k);                       // delta.sum() (:105-107)
    }
    n1 = warp_sum(n1); sd = warp_sum(sd);
    warp_sync();
    if (lane_id() == 0) {
        const double sg = ok ? ipm.sigma_val() : a.sigma[n];
        const double dsg = ok ? ipm.dsigma_val() : 0.;
        inf[0] = n1; inf[1] = sd; inf[2] = dsg; inf[3] = sg; inf[4] = w_tr;
        inf[5] = r.iterations; inf[6] = r.status; inf[7] = r.pres; inf[8] = r.dres; inf[9] = r.relgap;
        a.iters[n] = it + 1;
        if (!ok) { a.status[n] = r.status; a.converged[n] = 2; }             // reference: std::terminate (:94-98); here: flag the instance
        else {
            a.sigma[n] = sg;
            if (n1 < cfg.nu_tol) a.w_tr[n] = w_tr * 2.;                      // :112-115
            if (sd < cfg.delta_tol && n1 < cfg.nu_tol) a.converged[n] = 1;   // :131
            if (r.status == 3) a.status[n] = 3;
        }
    }
    if (a.hist && ok) {
        double *h = a.hist + ((size_t)n * (a.max_it + 1) + it + 1) * a.hist_stride();
        FOR_LANE(e, K * NB) { const int k = e / NB, i = e - k * NB; h[e] = ipm.xi_at(k, i); }
        if (lane_id() == 0) h[K * NB] = ipm.sigma_val();
    }
    warp_sync();
}

// ---- SCvx, stage 1 of an outer iteration's end (warp, K2 epilogue): readSolution into the CANDIDATE iterate (SCvxAlgorithm.cpp:94,
// :218-232) and norm1_nu; the nonlinear cost needs a simulation of every interval and runs as its own kernel (sc_scvx_cost), the ratio
// test after it (sc_scvx_decide)
template <class M>
SCPP_HD void sc_scvx_candidate(const ScArrays<M> &a, const ScConfig &cfg, int n, const Ipm<M> &ipm, const IpmResult &r)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    const int K = a.K;
    double *Xc = a.Xc + (size_t)n * K * NX, *Uc = a.Uc + (size_t)n * K * NU;
    const bool ok = (r.status == 0 || r.status == 3);
    warp_sync();
    double n1 = 0;
    if (ok) {
        FOR_LANE(e, K * NB) { const int k = e / NB, i = e - k * NB; const double v = ipm.xi_at(k, i); if (i < NX) Xc[k * NX + i] = v; else Uc[k * NU + (i - NX)] = v; }
        FOR_LANE(e, (K - 1) * NX) n1 += ipm.t_at(e / NX, e % NX);            // norm1_nu (:100-101)
    }
    n1 = warp_sum(n1);
    if (lane_id() == 0) {
        double *inf = a.info + ((size_t)n * a.max_it + a.iters[n]) * INFO_STRIDE;
        inf[5] = r.iterations; inf[6] = r.status; inf[7] = r.pres; inf[8] = r.dres; inf[9] = r.relgap;
        a.solves[n] += 1;
        if (!ok || a.solves[n] > SCVX_MAX_RESOLVE) { a.status[n] = ok ? 1 : r.status; a.converged[n] = 2; a.iters[n] += 1; }   // :87-91 (terminate)
        else { a.n1c[n] = n1; a.phase[n] = 1; if (r.status == 3) a.status[n] = 3; }
    }
    warp_sync();
}
// ---- SCvx, stage 2 (thread per (instance, interval)): one term of getNonlinearCost (SCvxAlgorithm.cpp:262-278) for the candidate, in the
// units the algorithm iterates on
template <class M>
SCPP_HD void sc_scvx_cost(const ScArrays<M> &a, const ScConfig &cfg, int n, int k)
{
    constexpr int NX = M::NX, NU = M::NU;
    const int K = a.K;
    if (a.phase[n] != 1 || k >= K - 1) return;
    const double *Xc = a.Xc + (size_t)n * K * NX, *Uc = a.Uc + (size_t)n * K * NU;
    double x[NX];
    for (int i = 0; i < NX; i++) x[i] = Xc[k * NX + i];
    rkf78_simulate<M>(x, Uc + k * NU, Uc + (k + 1) * NU, a.par + (size_t)n * M::NP, a.sigma[n] / (K - 1), 20);
    double c = 0.;
    for (int i = 0; i < NX; i++) c += fabs(x[i] - Xc[(k + 1) * NX + i]);
    a.costp[(size_t)n * K + k] = c;
}
// ---- SCvx, stage 3 (thread per instance): the ratio test and trust-region update of SCvxAlgorithm::iterate (:96-155)
template <class M>
SCPP_HD void sc_scvx_decide(const ScArrays<M> &a, const ScConfig &cfg, int n)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU;
    const int K = a.K;
    if (a.phase[n] != 1) return;
    a.phase[n] = 0;
    double J = 0.;
    for (int k = 0; k < K - 1; k++) J += a.costp[(size_t)n * K + k];               // nonlinear_cost (:98)
    const double L = a.n1c[n];                                                    // linear_cost   (:107-108)
    const int it = a.iters[n];
    double *inf = a.info + ((size_t)n * a.max_it + it) * INFO_STRIDE;
    inf[0] = L; inf[1] = J; inf[2] = 0.; inf[3] = a.trust[n]; inf[4] = a.solves[n];
    bool accept = true, converged = false;
    if (!a.have_last[n]) { a.last_cost[n] = J; a.have_last[n] = 1; }              // :110-114
    else {
        const double actual = a.last_cost[n] - J, predicted = a.last_cost[n] - L;   // :116-117
        a.last_cost[n] = J;                                                       // :119 (also when the step is rejected below)
        if (fabs(predicted) < cfg.scvx_change_threshold) converged = true;        // :126-130
        else {
            const double rho = actual / predicted;                                // :132
            inf[2] = rho;
            if (rho < cfg.scvx_rho_0) { a.trust[n] /= cfg.scvx_alpha; accept = false; }            // :133-139  td = old_td
            else if (rho < cfg.scvx_rho_1) a.trust[n] /= cfg.scvx_alpha;          // :144-148
            else if (rho >= cfg.scvx_rho_2) a.trust[n] *= cfg.scvx_beta;          // :149-153
        }
    }
    if (!accept) return;                                                          // same outer iteration: solve again with the smaller radius
    double *X = a.X + (size_t)n * K * NX, *U = a.U + (size_t)n * K * NU;
    const double *Xc = a.Xc + (size_t)n * K * NX, *Uc = a.Uc + (size_t)n * K * NU;
    for (int e = 0; e < K * NX; e++) X[e] = Xc[e];
    for (int e = 0; e < K * NU; e++) U[e] = Uc[e];
    a.iters[n] = it + 1; a.solves[n] = 0;
    if (converged) a.converged[n] = 1;
    if (a.hist) {
        double *h = a.hist + ((size_t)n * (a.max_it + 1) + it + 1) * a.hist_stride();
        for (int k = 0; k < K; k++) { for (int i = 0; i < NX; i++) h[k * NB + i] = X[k * NX + i]; for (int i = 0; i < NU; i++) h[k * NB + NX + i] = U[k * NU + i]; }
        h[K * NB] = a.sigma[n];
    }
}

// Interior warm start (ipm.warm in (0,1), an engine knob: ECOS has none).  The sub-problem that follows a STALLED outer iteration -- the sum of
// the trust-region radii of the previous solution is below delta_tol, the trajectory did not move -- is almost the sub-problem just solved, so
// it starts 50x closer to that solution: (s,z) <- w'(s,z) + (1-w')e with 1-w' = (1-warm)/50.  Measured on the kernel source (4 instances x 15
// iterations): 4 instead of 5 interior-point iterations per stalled sub-problem, while a uniformly closer start costs the moving sub-problems
// 30 % more iterations.  Same optimum, same tolerances; only the starting point changes.
constexpr double WARM_STALLED_GAIN = 0.02;
template <class M>
SCPP_HD IpmSettings sc_ipm_settings(const ScArrays<M> &a, const ScConfig &cfg, int n)
{
    IpmSettings st = cfg.ipm;
    const int it = a.iters[n];
    if (cfg.algorithm == 0 && st.warm > 0. && st.warm < 1. && it > 0) {
        const double sum_delta = a.info[((size_t)n * a.max_it + it - 1) * INFO_STRIDE + 1];
        if (sum_delta <= cfg.delta_tol) st.warm = 1. - (1. - st.warm) * WARM_STALLED_GAIN;
    }
    return st;
}
// the step fraction that goes with the starting point (Ipm::step_frac): 0.99 (ECOS / CVXOPT), or, as an experiment knob (ipm.stalled_step = d > 0),
// 1 - 10^-d for the close start after a stalled iteration.  Measured at batch 1024 (profiles/README.md): d = 4 takes the stalled sub-problems from
// 4 to 2 interior-point iterations on average (7.8 instead of 8.45 per instance-iteration) but a few instances per sub-problem then need 20-40, and
// a round waits for the slowest instance: 383 instead of 157 rounds per step, 26.6 k instead of 46.2 k instance-iterations/s.  Default 0.
template <class M>
SCPP_HD double sc_step_fraction(const ScArrays<M> &a, const ScConfig &cfg, int n)
{
    if (cfg.ipm.stalled_step <= 0 || !(cfg.ipm.warm > 0.) || sc_ipm_settings(a, cfg, n).warm == cfg.ipm.warm) return 0.99;
    double f = 1.;
    for (int d = 0; d < cfg.ipm.stalled_step && d < 8; d++) f *= 0.1;
    return 1. - f;
}

// ---- K2 + K3, monolithic: advance the sub-problem of instance n by one slice; when it is solved: K3 ----
// SCAlgorithm::iterate, SCAlgorithm.cpp:78-131 (the defect print :85-92 is diagnostic only and not computed)
template <class M>
SCPP_HD void sc_solve_instance(const ScArrays<M> &a, const ScConfig &cfg, int n, double *smem)
{
    Ipm<M> ipm;
    sc_bind(a, cfg, n, smem, ipm);
    ipm.step_frac = sc_step_fraction(a, cfg, n);
    bool finished;
    double *state = a.ipm_state + (size_t)n * Ipm<M>::IPM_STATE;
    const bool have_prev = (a.iters[n] > 0 || (cfg.algorithm == 1 && a.solves[n] > 0)) && a.status[n] != 2 && state[Ipm<M>::ST_NOPOINT] == 0.;
    const IpmResult r = ipm.solve(sc_ipm_settings(a, cfg, n), have_prev, cfg.ipm_slice > 0 ? cfg.ipm_slice : (cfg.ipm_slice < 0 ? 1 : (1 << 30)), state, finished);
    if (!finished) return;                                                   // continues in the next launch
    if (lane_id() == 0) state[Ipm<M>::ST_NOPOINT] = r.point_ok ? 0. : 1.;
    if (cfg.algorithm == 1) sc_scvx_candidate(a, cfg, n, ipm, r); else sc_finish_instance(a, cfg, n, ipm, r);
}

// ---- K2 + K3, one CTA per instance (cfg.solver == 1): the whole sub-problem of instance n, then K3 on warp 0 ----
template <class M>
SCPP_HD void sc_solve_instance_cta(const ScArrays<M> &a, const ScConfig &cfg, int n, double *smem)
{
    Ipm<M> ipm;
    sc_bind(a, cfg, n, smem, ipm, true);
    ipm.step_frac = sc_step_fraction(a, cfg, n);
    double *state = a.ipm_state + (size_t)n * Ipm<M>::IPM_STATE;
    const bool have_prev = (a.iters[n] > 0 || (cfg.algorithm == 1 && a.solves[n] > 0)) && a.status[n] != 2 && state[Ipm<M>::ST_NOPOINT] == 0.;
    cta_sync();                                                              // every thread has read the flags before warp 0 updates them below
    const IpmResult r = ipm.cp_solve_subproblem(sc_ipm_settings(a, cfg, n), have_prev);
    if (cta_tid() < LANES) {
        if (lane_id() == 0) state[Ipm<M>::ST_NOPOINT] = r.point_ok ? 0. : 1.;
        if (cfg.algorithm == 1) sc_scvx_candidate(a, cfg, n, ipm, r); else sc_finish_instance(a, cfg, n, ipm, r);
    }
    cta_sync();
}

// ---- K2 + K3, split pipeline (cfg.ipm_slice < 0): the steps of one interior-point iteration, each called by its own kernel ----
enum ScStep { SP_START = 0, SP_ASSEMBLE, SP_FACTOR, SP_RHS, SP_CHAIN, SP_RECOVER, SP_UPDATE, SP_RESIDUALS, SP_TEST };
// `sub`: stage (SP_ASSEMBLE) or 32-stage part (SP_RHS, SP_RECOVER, SP_UPDATE, SP_RESIDUALS); `mode`: 1 affine, 2 combined solve.
// Every step except SP_START acts only on instances that are mid-solve and whose factorisation did not fail.
template <class M, int STEP>
SCPP_HD void sc_split_step(const ScArrays<M> &a, const ScConfig &cfg, int n, double *smem, int mode, int sub)
{
    double *state = a.ipm_state + (size_t)n * Ipm<M>::IPM_STATE;
    if (STEP == SP_START) { if (state[0] != 0.) return; }
    else {
        if (state[0] != 1.) return;
        if (STEP != SP_ASSEMBLE && STEP != SP_FACTOR && STEP != SP_TEST && state[Ipm<M>::ST_FAIL] != 0.) return;
    }
    Ipm<M> ipm;
    sc_bind(a, cfg, n, smem, ipm);
    ipm.step_frac = sc_step_fraction(a, cfg, n);
    if (STEP != SP_START) ipm.dcap = state[Ipm<M>::ST_DCAP];
    IpmResult r;
    if (STEP == SP_START) {
        if (ipm.sp_start(sc_ipm_settings(a, cfg, n), (a.iters[n] > 0 || (cfg.algorithm == 1 && a.solves[n] > 0)) && a.status[n] != 2 && state[Ipm<M>::ST_NOPOINT] == 0., state, r)) { if (cfg.algorithm == 1) sc_scvx_candidate(a, cfg, n, ipm, r); else sc_finish_instance(a, cfg, n, ipm, r); }
    } else if (STEP == SP_ASSEMBLE) {
        if (sub >= a.K) return;
        ipm.tables_init();
        ipm.assemble_stage(sub);
    } else if (STEP == SP_FACTOR) ipm.sp_factor(state);
    else if (STEP == SP_CHAIN) ipm.sp_chain(mode, state);
    else if (STEP == SP_TEST) { if (ipm.sp_test(cfg.ipm, state, r)) { if (cfg.algorithm == 1) sc_scvx_candidate(a, cfg, n, ipm, r); else sc_finish_instance(a, cfg, n, ipm, r); } }
    else {
        if (sub >= ipm.nparts()) return;
        ipm.cst_init();
        if (STEP == SP_RHS) ipm.sp_rhs(mode, state, sub);
        else if (STEP == SP_RECOVER) ipm.sp_recover(mode, state, sub);
        else if (STEP == SP_UPDATE) ipm.sp_update(sub);
        else if (STEP == SP_RESIDUALS) ipm.sp_residuals(sub);
    }
}

} // namespace scpp
