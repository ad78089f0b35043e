// scpp_b200/csrc/hostsim.cpp — TEST-ONLY host-simulation build of the kernel bodies (LANES == 1).
// Built into tests/_hostsim/libhostsim.so by tests/hostsim.py; never part of libscpp_b200.so and never
// loaded by the scpp_b200 package: the product has no CPU execution path.
#include "sc.cuh"
#include "discretize_shared.cuh"
#include "lqr.cuh"
#include <cstddef>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cmath>

using namespace scpp;

template <class M>
static void run_discretize(int K, const double *X, const double *U, double sigma, const double *par, int nsub, double *dd, double *ddT = nullptr, int jacobian = 0, bool zoh = false)
{
    constexpr int NC = M::NX + 2 * M::NU + 2;
    if (jacobian == 2) {      // the roles of k_discretize_shared run one after the other, step by step
        constexpr int NX = M::NX, NU = M::NU;
        const int S = k1s_steps(nsub);
        double h0, h1, rdtau;
        k1s_step_sizes(nsub, K, h0, h1, rdtau);
        for (int k = 0; k < K - 1; k++) {
            std::vector<double> stash(k1s_stride<M>()), xs(k1s_xstride<M>()), cols(NC * NX, 0.), colc(NC * NX, 0.);
            double x[NX], x0[NX], u0[NU], du[NU];
            for (int i = 0; i < NX; i++) x0[i] = X[NX * k + i];
            for (int j = 0; j < NU; j++) { u0[j] = U[NU * k + j]; du[j] = zoh ? 0. : U[NU * (k + 1) + j] - u0[j]; }
            StageLin<M> *rec = reinterpret_cast<StageLin<M> *>(stash.data());
            for (int s = 0; s < S; s++) {
                int ns, st;
                k1s_schedule(nsub, s, ns, st);
                k1s_chain<M>(x, x0, u0, du, par, sigma, (nsub < 0 && s >= -nsub) ? h1 : h0, rdtau, st, xs.data());
                for (int sgi = 0; sgi < 4; sgi++) k1s_linearize<M>(xs.data() + sgi * (NX + NU), par, rec[sgi]);
                for (int c = 0; c < NC; c++) {
                    int ctype, cidx;
                    k1s_column_type(NX, NU, c, ctype, cidx);
                    k1s_consumer_step<M>(&cols[c * NX], &colc[c * NX], 1, ctype, cidx, sigma, 1. / sigma, h0, h1, rdtau, nsub, s, rec, zoh);
                }
            }
            for (int c = 0; c < NC; c++)
                for (int i = 0; i < NX; i++) {
                    dd[(size_t)k * NX * NC + i * NC + c] = cols[c * NX + i];
                    if (ddT) ddT[(size_t)(i * NC + c) * Ipm<M>::ks(K) + k] = cols[c * NX + i];
                }
        }
        return;
    }
    for (int k = 0; k < K - 1; k++)
        for (int c = 0; c < NC; c++) {
            if (jacobian) discretize_column<M, true>(X, U, sigma, par, K, k, c, nsub, zoh ? 3 : 1, dd + (size_t)k * M::NX * NC, ddT, Ipm<M>::ks(K));
            else discretize_column<M, false>(X, U, sigma, par, K, k, c, nsub, zoh ? 3 : 1, dd + (size_t)k * M::NX * NC, ddT, Ipm<M>::ks(K));
        }
}
extern "C" void hs_discretize2(int model, int K, const double *X, const double *U, double sigma, const double *par, int nsub, int jacobian, double *dd)
{
    if (model == 0) run_discretize<RocketQuat>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian);
    else if (model == 2) run_discretize<Rocket2dPlugin>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian);
    else if (model == 3) run_discretize<RocketQuatRollPlugin>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian);
    else run_discretize<Rocket2d>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian);
}

extern "C" void hs_discretize_zoh(int model, int K, const double *X, const double *U, double sigma, const double *par, int nsub, int jacobian, double *dd)
{
    if (model == 0) run_discretize<RocketQuat>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian, true);
    else run_discretize<Rocket2d>(K, X, U, sigma, par, nsub, dd, nullptr, jacobian, true);
}

extern "C" void hs_discretize(int model, int K, const double *X, const double *U, double sigma, const double *par, int nsub, double *dd)
{
    if (model == 0) run_discretize<RocketQuat>(K, X, U, sigma, par, nsub, dd);
    else run_discretize<Rocket2d>(K, X, U, sigma, par, nsub, dd);
}

// host stand-in for EngineT: the per-instance arrays and the round loop of solve(); one object can run a closed loop
template <class M>
struct HostEngine {
    static constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2;
    ScArrays<M> a;
    ModelParamsHost P;
    ScConfig cfg;
    int N, K;
    std::vector<double> x_init, x_final, xi, xf, par, cst, scale, tdir, fixv, wtr, dd, ddT, ws, smem, ist, X, U, sigma, hist, info;
    std::vector<uint32_t> fixm;
    std::vector<int> iters, status, converged, frozen, have_last, solves, phase;
    std::vector<double> trust, last_cost, n1c, Xc, Uc, costp;
    HostEngine(const ModelParamsHost &P_, const ScConfig &cfg_, int N_, const double *xi_in, const double *xf_in) : P(P_), cfg(cfg_), N(N_), K(cfg_.K)
    {
        x_init.assign(xi_in, xi_in + (size_t)N * NX); x_final.assign(xf_in, xf_in + (size_t)N * NX);
        xi.resize(N * NX); xf.resize(N * NX); par.resize(N * M::NP); cst.resize(N * MAX_CST); scale.resize(N * 2); tdir.resize((size_t)N * K * 3);
        fixv.resize((size_t)N * K * NB); wtr.resize(N); fixm.resize((size_t)N * K);
        dd.resize((size_t)N * (K - 1) * NX * NC); ddT.resize((size_t)N * Ipm<M>::ddt_doubles(K));
        a.N = N; a.K = K; a.max_it = cfg.max_iterations; a.Pn = nullptr;
        a.ws_stride = Ipm<M>::ws_doubles(K);
        ws.resize((size_t)N * a.ws_stride); smem.resize(Ipm<M>::sm_doubles() + Ipm<M>::cta_sm_doubles(K)); ist.resize((size_t)N * Ipm<M>::IPM_STATE);
        X.resize((size_t)N * K * NX); U.resize((size_t)N * K * NU); sigma.resize(N);
        hist.resize((size_t)N * (cfg.max_iterations + 1) * (K * NB + 1)); info.resize((size_t)N * cfg.max_iterations * INFO_STRIDE);
        iters.resize(N); status.resize(N); converged.resize(N); frozen.assign(N, 0);
        a.ipm_state = ist.data(); a.x_init = x_init.data(); a.x_final = x_final.data();
        a.xi = xi.data(); a.xf = xf.data(); a.par = par.data(); a.cst = cst.data(); a.scale = scale.data();
        a.X = X.data(); a.U = U.data(); a.sigma = sigma.data(); a.tdir = tdir.data(); a.fixm = fixm.data(); a.fixv = fixv.data(); a.w_tr = wtr.data();
        a.iters = iters.data(); a.status = status.data(); a.converged = converged.data(); a.dd = dd.data(); a.ddT = ddT.data(); a.ws = ws.data();
        a.hist = cfg.keep_history ? hist.data() : nullptr; a.info = info.data(); a.frozen = frozen.data();
        trust.resize(N); last_cost.resize(N); n1c.resize(N); Xc.resize((size_t)N * K * NX); Uc.resize((size_t)N * K * NU); costp.resize((size_t)N * K);
        have_last.resize(N); solves.resize(N); phase.resize(N);
        a.trust = trust.data(); a.last_cost = last_cost.data(); a.n1c = n1c.data(); a.Xc = Xc.data(); a.Uc = Uc.data(); a.costp = costp.data();
        a.have_last = have_last.data(); a.solves = solves.data(); a.phase = phase.data();
    }
    void solve(bool warm)
    {
        for (int n = 0; n < N; n++) {
            if (warm) { if (frozen[n]) { converged[n] = 8; continue; } sc_warm_instance<M>(a, P, cfg, n); }
            else { frozen[n] = 0; sc_setup_instance<M>(a, P, cfg, n); }
        }
        // the engine's round structure: every round discretises the instances that start a new sub-problem and advances every
        // unfinished instance by one slice of interior-point iterations
        long rounds = 0;
        for (;;) {
            int active = 0; rounds++;
            for (int n = 0; n < N; n++) {
                if (converged[n] || iters[n] >= cfg.max_iterations) continue;
                active++;
                if (a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE] == 0.)
                    run_discretize<M>(K, a.X + (size_t)n * K * NX, a.U + (size_t)n * K * NU, sigma[n], a.par + (size_t)n * M::NP, cfg.nsub,
                                      a.dd + (size_t)n * (K - 1) * NX * NC, a.ddT + (size_t)n * Ipm<M>::ddt_doubles(K), cfg.jacobian, !cfg.interpolate_input);
                if (cfg.solver == 1) {
                    sc_solve_instance_cta<M>(a, cfg, n, smem.data());
                    if (cfg.algorithm == 1) { for (int k = 0; k < K - 1; k++) sc_scvx_cost<M>(a, cfg, n, k); sc_scvx_decide<M>(a, cfg, n); }
                    continue;
                }
                if (cfg.ipm_slice >= 0) {
                    sc_solve_instance<M>(a, cfg, n, smem.data());
                    if (cfg.algorithm == 1) { for (int k = 0; k < K - 1; k++) sc_scvx_cost<M>(a, cfg, n, k); sc_scvx_decide<M>(a, cfg, n); }   // the cost / decide kernels
                    continue;
                }
                // split pipeline: the kernel sequence of one round, run here for one instance after the other
                double *w = smem.data();
                const int Pn = (K + 31) / 32;
                sc_split_step<M, SP_START>(a, cfg, n, w, 0, 0);
                for (int k = 0; k < K; k++) sc_split_step<M, SP_ASSEMBLE>(a, cfg, n, w, 0, k);
                sc_split_step<M, SP_FACTOR>(a, cfg, n, w, 0, 0);
                for (int mode = 1; mode <= 2; mode++) {
                    for (int p = 0; p < Pn; p++) sc_split_step<M, SP_RHS>(a, cfg, n, w, mode, p);
                    sc_split_step<M, SP_CHAIN>(a, cfg, n, w, mode, 0);
                    for (int p = 0; p < Pn; p++) sc_split_step<M, SP_RECOVER>(a, cfg, n, w, mode, p);
                }
                for (int p = 0; p < Pn; p++) sc_split_step<M, SP_UPDATE>(a, cfg, n, w, 0, p);
                for (int p = 0; p < Pn; p++) sc_split_step<M, SP_RESIDUALS>(a, cfg, n, w, 0, p);
                sc_split_step<M, SP_TEST>(a, cfg, n, w, 0, 0);
                if (cfg.algorithm == 1) { for (int k = 0; k < K - 1; k++) sc_scvx_cost<M>(a, cfg, n, k); sc_scvx_decide<M>(a, cfg, n); }
            }
            if (!active) break;
        }
        if (getenv("SCPP_DEBUG_ROUNDS")) fprintf(stderr, "rounds %ld\n", rounds);
    }
    // redimensionalised copy of the final trajectories (SCAlgorithm.cpp:182-187)
    void export_solution(double *Xo, double *Uo, double *so) const
    {
        for (int n = 0; n < N; n++) {
            for (int k = 0; k < K; k++) {
                double *x = Xo + ((size_t)n * K + k) * NX, *u = Uo + ((size_t)n * K + k) * NU;
                for (int i = 0; i < NX; i++) x[i] = X[((size_t)n * K + k) * NX + i];
                for (int i = 0; i < NU; i++) u[i] = U[((size_t)n * K + k) * NU + i];
                M::redim(scale.data() + 2 * n, x, u);
            }
            so[n] = sigma[n];
        }
    }
};

template <class M>
static int run_sc(const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final,
                  double *X, double *U, double *sigma, int *iters, int *status, int *converged, double *hist, double *info)
{
    ScConfig c = *cfg; c.keep_history = 1;
    HostEngine<M> e(*P, c, N, x_init, x_final);
    e.solve(false);
    e.export_solution(X, U, sigma);
    memcpy(iters, e.iters.data(), sizeof(int) * N); memcpy(status, e.status.data(), sizeof(int) * N); memcpy(converged, e.converged.data(), sizeof(int) * N);
    if (hist) memcpy(hist, e.hist.data(), sizeof(double) * e.hist.size());
    if (info) memcpy(info, e.info.data(), sizeof(double) * e.info.size());
    return 0;
}

extern "C" int hs_sc_solve(int model, const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final,
                           double *X, double *U, double *sigma, int *iters, int *status, int *converged, double *hist, double *info)
{
    if (model == 0) return run_sc<RocketQuat>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
    if (model == 2) return run_sc<Rocket2dPlugin>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
    if (model == 3) return run_sc<RocketQuatRollPlugin>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
    return run_sc<Rocket2d>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
}

// closed loop of scpp/src/SC_sim.cpp for N instances: X_sim [steps][N][nx], U_sim [steps][N][nu], iters [steps][N], reached [N]
template <class M>
static int run_sim(const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final, double time_step, int steps,
                   double *X_sim, double *U_sim, int *iters, int *reached)
{
    HostEngine<M> e(*P, *cfg, N, x_init, x_final);
    for (int n = 0; n < N; n++) reached[n] = 0;
    for (int s = 0; s < steps; s++) {
        e.solve(s > 0);
        for (int n = 0; n < N; n++) {
            iters[(size_t)s * N + n] = e.frozen[n] ? 0 : e.iters[n];
            int r = 0;
            sc_sim_step_instance<M>(e.a, e.P, e.cfg, n, time_step, X_sim + ((size_t)s * N + n) * M::NX, U_sim + ((size_t)s * N + n) * M::NU, &r);
            reached[n] = r;
        }
    }
    return 0;
}
extern "C" int hs_sc_sim(int model, const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final, double time_step,
                         int steps, double *X_sim, double *U_sim, int *iters, int *reached)
{
    if (model == 0) return run_sim<RocketQuat>(P, cfg, N, x_init, x_final, time_step, steps, X_sim, U_sim, iters, reached);
    return run_sim<Rocket2d>(P, cfg, N, x_init, x_final, time_step, steps, X_sim, U_sim, iters, reached);
}
extern "C" void hs_simulate(int model, double dt, double *x, const double *u0, const double *u1, const double *par)
{
    if (model == 0) rkf78_simulate<RocketQuat>(x, u0, u1, par, dt, 20); else rkf78_simulate<Rocket2d>(x, u0, u1, par, dt, 20);
}

// forward-mode dual number: instantiates the generic-scalar flow map (the plugin surface, systemFlowMap) to obtain the exact
// Jacobian the reference gets from CppAD (systemDynamics.hpp:206-235); compared in tests with the hand-derived sparse one
struct HDual {
    double v, d;
    HDual(double v_ = 0., double d_ = 0.) : v(v_), d(d_) {}
};
static inline HDual operator+(HDual a, HDual b) { return {a.v + b.v, a.d + b.d}; }
static inline HDual operator-(HDual a, HDual b) { return {a.v - b.v, a.d - b.d}; }
static inline HDual operator-(HDual a) { return {-a.v, -a.d}; }
static inline HDual operator*(HDual a, HDual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
static inline HDual operator/(HDual a, HDual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
static inline HDual sqrt(HDual a) { double r = std::sqrt(a.v); return {r, a.d / (2. * r)}; }
static inline HDual sin(HDual a) { return {std::sin(a.v), std::cos(a.v) * a.d}; }
static inline HDual cos(HDual a) { return {std::cos(a.v), -std::sin(a.v) * a.d}; }

template <class M>
static void run_jac(const double *x, const double *u, const double *par, double *f, double *A_ad, double *B_ad, double *A_lin, double *B_lin)
{
    constexpr int NX = M::NX, NU = M::NU;
    M::template flow_map<double>(x, u, par, f);
    for (int j = 0; j < NX + NU; j++) {
        HDual xd[NX], ud[NU], fd[NX];
        for (int i = 0; i < NX; i++) xd[i] = HDual(x[i], i == j ? 1. : 0.);
        for (int i = 0; i < NU; i++) ud[i] = HDual(u[i], NX + i == j ? 1. : 0.);
        M::template flow_map<HDual>(xd, ud, par, fd);
        for (int i = 0; i < NX; i++) { if (j < NX) A_ad[i * NX + j] = fd[i].d; else B_ad[i * NU + (j - NX)] = fd[i].d; }
    }
    typename M::Lin L;
    M::linearize(x, u, par, L);
    for (int j = 0; j < NX; j++) { double e[NX] = {0}, o[NX]; e[j] = 1.; M::A_apply(L, e, o); for (int i = 0; i < NX; i++) A_lin[i * NX + j] = o[i]; }
    for (int j = 0; j < NU; j++) { double e[NU] = {0}, o[NX]; e[j] = 1.; M::B_apply(L, e, o); for (int i = 0; i < NX; i++) B_lin[i * NU + j] = o[i]; }
    for (int i = 0; i < NX; i++) if (std::fabs(L.f[i] - f[i]) > 1e-13 * (1. + std::fabs(f[i]))) f[i] = NAN;   // Lin.f must equal flow_map
}
extern "C" void hs_jacobians(int model, const double *x, const double *u, const double *par, double *f, double *A_ad, double *B_ad, double *A_lin, double *B_lin)
{
    if (model == 0) run_jac<RocketQuat>(x, u, par, f, A_ad, B_ad, A_lin, B_lin); else if (model == 2) run_jac<Rocket2dPlugin>(x, u, par, f, A_ad, B_ad, A_lin, B_lin); else if (model == 3) run_jac<RocketQuatRollPlugin>(x, u, par, f, A_ad, B_ad, A_lin, B_lin); else run_jac<Rocket2d>(x, u, par, f, A_ad, B_ad, A_lin, B_lin);
}

// K5 body on the host: gains [K][nu][nx], ok [K]
template <class M>
static void run_lqr(int K, const double *X, const double *U, const double *par, const double *qd, const double *rd, double *gains, int *ok)
{
    std::vector<double> sm(Lqr<M>::sm_doubles());
    for (int k = 0; k < K; k++) ok[k] = Lqr<M>::gain(X + (size_t)k * M::NX, U + (size_t)k * M::NU, par, qd, rd, gains + (size_t)k * M::NU * M::NX, sm.data());
}
extern "C" void hs_lqr_gains(int model, int K, const double *X, const double *U, const double *par, const double *qd, const double *rd, double *gains, int *ok)
{
    if (model == 0) run_lqr<RocketQuat>(K, X, U, par, qd, rd, gains, ok); else run_lqr<Rocket2d>(K, X, U, par, qd, rd, gains, ok);
}

// K6 body on the host: the dense conic solver of mpc.cuh on a problem given in full (G row-major [nr][nv])
#include "mpc.cuh"
extern "C" int hs_dense_conic(int nv, int nl, int ncones, const int *cdim, const double *G, const double *c, const double *h, double tol, double *y, int *iters)
{
    static DenseConic<48, 320, 48> S;
    int nr = nl;
    for (int k = 0; k < ncones; k++) nr += cdim[k];
    if (nv > 48 || nr > 320 || ncones > 48) return -1;
    S.nv = nv; S.nl = nl; S.ncones = ncones; S.nr = nr; S.cdim = cdim; S.G = G; S.c = c;
    IpmSettings st; st.feastol = st.abstol = st.reltol = tol; st.maxit = 100; st.stalled_step = 0; st.warm = 0.;
    const IpmResult r = S.solve(h, st);
    for (int i = 0; i < nv; i++) y[i] = S.y[i];
    *iters = r.iterations;
    return r.status;
}

extern "C" int hs_sizes(int which) { return which == 0 ? (int)sizeof(ModelParamsHost) : (int)sizeof(ScConfig); }
