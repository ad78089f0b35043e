// scpp_b200/csrc/hostsim.cpp — TEST-ONLY host-simulation build of the kernel bodies (LANES == 1).
// Built into tests/_hostsim/libhostsim.so by tests/hostsim.py; never part of libscpp_b200.so and never
// loaded by the scpp_b200 package: the product has no CPU execution path.
#include "sc.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cmath>

using namespace scpp;

template <class M>
static void run_discretize(int K, const double *X, const double *U, double sigma, const double *par, int nsub, double *dd, double *ddT = nullptr)
{
    constexpr int NC = M::NX + 2 * M::NU + 2;
    for (int k = 0; k < K - 1; k++)
        for (int c = 0; c < NC; c++)
            discretize_column<M>(X, U, sigma, par, K, k, c, nsub, 1, dd + (size_t)k * M::NX * NC, ddT, Ipm<M>::ks(K));
}

extern "C" void hs_discretize(int model, int K, const double *X, const double *U, double sigma, const double *par, int nsub, double *dd)
{
    if (model == 0) run_discretize<RocketQuat>(K, X, U, sigma, par, nsub, dd);
    else run_discretize<Rocket2d>(K, X, U, sigma, par, nsub, dd);
}

template <class M>
static int run_sc(const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final,
                  double *X, double *U, double *sigma, int *iters, int *status, int *converged, double *hist, double *info)
{
    constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2;
    const int K = cfg->K;
    ScArrays<M> a;
    a.N = N; a.K = K; a.max_it = cfg->max_iterations;
    std::vector<double> xi(N * NX), xf(N * NX), par(N * M::NP), cst(N * MAX_CST), scale(N * 2), tdir((size_t)N * K * 3), fixv((size_t)N * K * NB), wtr(N);
    std::vector<uint32_t> fixm((size_t)N * K);
    std::vector<double> dd((size_t)N * (K - 1) * NX * NC), ddT((size_t)N * Ipm<M>::ddt_doubles(K));
    a.ws_stride = Ipm<M>::ws_doubles(K);
    std::vector<double> ws((size_t)N * a.ws_stride), smem(Ipm<M>::sm_doubles()), ist((size_t)N * Ipm<M>::IPM_STATE);
    a.ipm_state = ist.data();
    a.x_init = const_cast<double *>(x_init); a.x_final = const_cast<double *>(x_final);
    a.xi = xi.data(); a.xf = xf.data(); a.par = par.data(); a.cst = cst.data(); a.scale = scale.data();
    a.X = X; a.U = U; a.sigma = sigma; a.tdir = tdir.data(); a.fixm = fixm.data(); a.fixv = fixv.data(); a.w_tr = wtr.data();
    a.iters = iters; a.status = status; a.converged = converged; a.dd = dd.data(); a.ddT = ddT.data(); a.ws = ws.data(); a.hist = hist; a.info = info;
    for (int n = 0; n < N; n++) sc_setup_instance<M>(a, *P, *cfg, n);
    // the engine's round structure: every round discretises the instances that start a new sub-problem and advances every
    // unfinished instance by one slice of interior-point iterations
    long rounds = 0;
    for (;;) {
        int active = 0; rounds++;
        for (int n = 0; n < N; n++) {
            if (converged[n] || iters[n] >= cfg->max_iterations) continue;
            active++;
            if (a.ipm_state[(size_t)n * Ipm<M>::IPM_STATE] == 0.)
                run_discretize<M>(K, X + (size_t)n * K * NX, U + (size_t)n * K * NU, sigma[n], a.par + (size_t)n * M::NP, cfg->nsub,
                                  a.dd + (size_t)n * (K - 1) * NX * NC, a.ddT + (size_t)n * Ipm<M>::ddt_doubles(K));
            if (cfg->ipm_slice >= 0) { sc_solve_instance<M>(a, *cfg, n, smem.data()); continue; }
            // split pipeline: the kernel sequence of one round, run here for one instance after the other
            double *w = smem.data();
            const int P = (K + 31) / 32;
            sc_split_step<M, SP_START>(a, *cfg, n, w, 0, 0);
            for (int k = 0; k < K; k++) sc_split_step<M, SP_ASSEMBLE>(a, *cfg, n, w, 0, k);
            sc_split_step<M, SP_FACTOR>(a, *cfg, n, w, 0, 0);
            for (int mode = 1; mode <= 2; mode++) {
                for (int p = 0; p < P; p++) sc_split_step<M, SP_RHS>(a, *cfg, n, w, mode, p);
                sc_split_step<M, SP_CHAIN>(a, *cfg, n, w, mode, 0);
                for (int p = 0; p < P; p++) sc_split_step<M, SP_RECOVER>(a, *cfg, n, w, mode, p);
            }
            for (int p = 0; p < P; p++) sc_split_step<M, SP_UPDATE>(a, *cfg, n, w, 0, p);
            for (int p = 0; p < P; p++) sc_split_step<M, SP_RESIDUALS>(a, *cfg, n, w, 0, p);
            sc_split_step<M, SP_TEST>(a, *cfg, n, w, 0, 0);
        }
        if (!active) break;
    }
    if (getenv("SCPP_DEBUG_ROUNDS")) fprintf(stderr, "rounds %ld\n", rounds);
    // redimensionalise the final trajectories (SCAlgorithm.cpp:182-187)
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) M::redim(a.scale + 2 * n, X + ((size_t)n * K + k) * NX, U + ((size_t)n * K + k) * NU);
    return 0;
}

extern "C" int hs_sc_solve(int model, const ModelParamsHost *P, const ScConfig *cfg, int N, const double *x_init, const double *x_final,
                           double *X, double *U, double *sigma, int *iters, int *status, int *converged, double *hist, double *info)
{
    if (model == 0) return run_sc<RocketQuat>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
    return run_sc<Rocket2d>(P, cfg, N, x_init, x_final, X, U, sigma, iters, status, converged, hist, info);
}

// forward-mode dual number: instantiates the generic-scalar flow map (the plugin surface, systemFlowMap) to obtain the exact
// Jacobian the reference gets from CppAD (systemDynamics.hpp:206-235); compared in tests with the hand-derived sparse one
struct Dual {
    double v, d;
    Dual(double v_ = 0., double d_ = 0.) : v(v_), d(d_) {}
};
static inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
static inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
static inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
static inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
static inline Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
static inline Dual sqrt(Dual a) { double r = std::sqrt(a.v); return {r, a.d / (2. * r)}; }
static inline Dual sin(Dual a) { return {std::sin(a.v), std::cos(a.v) * a.d}; }
static inline Dual cos(Dual a) { return {std::cos(a.v), -std::sin(a.v) * a.d}; }

template <class M>
static void run_jac(const double *x, const double *u, const double *par, double *f, double *A_ad, double *B_ad, double *A_lin, double *B_lin)
{
    constexpr int NX = M::NX, NU = M::NU;
    M::template flow_map<double>(x, u, par, f);
    for (int j = 0; j < NX + NU; j++) {
        Dual xd[NX], ud[NU], fd[NX];
        for (int i = 0; i < NX; i++) xd[i] = Dual(x[i], i == j ? 1. : 0.);
        for (int i = 0; i < NU; i++) ud[i] = Dual(u[i], NX + i == j ? 1. : 0.);
        M::template flow_map<Dual>(xd, ud, par, fd);
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
    if (model == 0) run_jac<RocketQuat>(x, u, par, f, A_ad, B_ad, A_lin, B_lin); else run_jac<Rocket2d>(x, u, par, f, A_ad, B_ad, A_lin, B_lin);
}

extern "C" int hs_sizes(int which) { return which == 0 ? (int)sizeof(ModelParamsHost) : (int)sizeof(ScConfig); }
