// scpp_b200/csrc/mpc.cu — batched receding-horizon MPC: host set-up, kernel K6, C-ABI.
//
// Reference: scpp_core/src/MPCAlgorithm.cpp:11-140 (loadParameters, initialize, solve, readSolution), scpp_core/src/MPCProblem.cpp:6-87
// (buildMPCProblem), scpp_core/src/discretization.cpp:9-40 (exactLinearDiscretization: zero-order hold through the matrix exponential),
// scpp_models/src/rocket2d.cpp:40-84 (getOperatingPoint, addApplicationConstraints), scpp/src/MPC_sim.cpp:40-84 (closed loop).
// Built for sm_100a only; no CPU execution path (the host code below only prepares the problem data).
#include "../../include/scpp_b200.h"
#include "mpc.cuh"
#include "simulate.cuh"
#include "info_parser.hpp"

#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <vector>

using namespace scpp;

static_assert(sizeof(scpp_b200_mpc_config) == sizeof(MpcConfig), "ABI struct mismatch");

extern int scpp_b200_fail(int code, const std::string &msg);      // engine.cu: sets scpp_b200_last_error
#define MCU(call)                                                                                                            \
    do {                                                                                                                     \
        cudaError_t e_ = (call);                                                                                             \
        if (e_ != cudaSuccess) return scpp_b200_fail(SCPP_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---- dense helpers (row-major, host) -----------------------------------------------------------------------------------------------
typedef std::vector<double> Mat;
static Mat matmul(const Mat &a, const Mat &b, int n, int m, int p)      // (n x m)(m x p)
{
    Mat c((size_t)n * p, 0.);
    for (int i = 0; i < n; i++) for (int k = 0; k < m; k++) { const double v = a[(size_t)i * m + k]; if (v != 0.) for (int j = 0; j < p; j++) c[(size_t)i * p + j] += v * b[(size_t)k * p + j]; }
    return c;
}
// matrix exponential (what Eigen's MatrixBase::exp() returns, discretization.cpp:29,37): scaling and squaring with a Taylor series summed to
// rounding (the scaled norm is below 1/2, so 20 terms reach 1e-19)
static Mat expm(Mat a, int n)
{
    double nrm = 0;
    for (int i = 0; i < n; i++) { double r = 0; for (int j = 0; j < n; j++) r += fabs(a[(size_t)i * n + j]); nrm = fmax(nrm, r); }
    int s = 0;
    while (nrm > 0.5) { nrm *= 0.5; s++; }
    for (double &v : a) v = ldexp(v, -s);
    Mat e((size_t)n * n, 0.), term((size_t)n * n, 0.);
    for (int i = 0; i < n; i++) { e[(size_t)i * n + i] = 1.; term[(size_t)i * n + i] = 1.; }
    for (int k = 1; k <= 24; k++) {
        term = matmul(term, a, n, n, n);
        for (double &v : term) v /= k;
        for (size_t i = 0; i < e.size(); i++) e[i] += term[i];
    }
    for (int i = 0; i < s; i++) e = matmul(e, e, n, n, n);
    return e;
}

// kernel K6: one thread per instance
template <class M, int KM>
__global__ void __launch_bounds__(64) k_mpc(MpcProblem P, int K, IpmSettings st, int n, const double *__restrict__ x0, const double *__restrict__ xf,
                                            double *X, double *U, int *status, int *iters)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mpc_solve_instance<M, KM>(P, K, st, x0 + (size_t)i * M::NX, xf + (size_t)i * M::NX, X + (size_t)i * K * M::NX, U + (size_t)i * (K - 1) * M::NU, status + i, iters + i);
}
// closed loop (MPC_sim.cpp:64-70): apply the first input for dt with scpp::simulate (u0 = u1 = u), x_init <- simulated state
template <class M>
__global__ void k_mpc_step(int n, int K, double dt, double *x0, const double *U, const double *par, double *x_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[M::NX];
    for (int e = 0; e < M::NX; e++) x[e] = x0[(size_t)i * M::NX + e];
    const double *u = U + (size_t)i * (K - 1) * M::NU;
    rkf78_simulate<M>(x, u, u, par, dt, 20);
    for (int e = 0; e < M::NX; e++) { x0[(size_t)i * M::NX + e] = x[e]; if (x_out) x_out[(size_t)i * M::NX + e] = x[e]; }
}

struct scpp_b200_mpc {
    virtual ~scpp_b200_mpc() {}
    virtual int init() = 0;
    virtual int set_states(const double *xi, const double *xf) = 0;
    virtual int solve() = 0;
    virtual int get_solution(double *X, double *U, int *status, int *iters) = 0;
    virtual int sim_step(double dt, double *x_new) = 0;
    int model = 0, N = 0, device = 0;
    ModelParamsHost P;
    MpcConfig cfg;
    std::vector<double> A, B, z;      // exactLinearDiscretization (row-major)
    double ms_solve = 0;
};

template <class M>
struct MpcEngineT : scpp_b200_mpc {
    static constexpr int NX = M::NX, NU = M::NU, NROW = M::NLP + M::NCR;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    std::vector<void *> allocs;
    MpcProblem dp{};
    double *d_x0 = nullptr, *d_xf = nullptr, *d_X = nullptr, *d_U = nullptr, *d_par = nullptr, *d_xs = nullptr;
    int *d_status = nullptr, *d_iters = nullptr;
    bool have_states = false, solved = false;

    ~MpcEngineT() override
    {
        cudaSetDevice(device);
        for (void *p : allocs) cudaFree(p);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
    template <class T>
    int upload(const std::vector<T> &h, const T **d)
    {
        T *p = nullptr;
        MCU(cudaMalloc((void **)&p, h.size() * sizeof(T)));
        allocs.push_back(p);
        MCU(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        *d = p;
        return 0;
    }
    template <class T>
    int dalloc(T **p, size_t n) { MCU(cudaMalloc((void **)p, n * sizeof(T))); MCU(cudaMemset(*p, 0, n * sizeof(T))); allocs.push_back(*p); return 0; }

    // exactLinearDiscretization at the operating point + the condensed conic program
    int build(std::string &err)
    {
        const int K = cfg.K, NS = K - 1, nuu = NU * NS;
        double xi[NX] = {0}, xfz[NX] = {0}, par[M::NP], cst[MAX_CST], sc2[2];
        M::setup(P, 0, xi, xfz, par, cst, sc2);
        // getOperatingPoint (rocket2d.cpp:40-44): x = 0, u = (0, -g_I m).  The reference writes `u << 0, -p.g_I * p.m` into a 2-vector, i.e.
        // three coefficients: the intended hover point (gimbal 0, thrust |g| m) is used here
        double xe[NX] = {0}, ue[NU] = {0};
        if (!M::operating_point(P, xe, ue)) { err = "this model has no operating point (getOperatingPoint throws in the reference, systemModel.hpp:119)"; return SCPP_B200_ERR_UNSUPPORTED; }
        // A_c, B_c, f by dual numbers over the flow map (computeJacobians / computef, discretization.cpp:19-20)
        Mat Ac((size_t)NX * NX), Bc((size_t)NX * NU), f(NX);
        for (int j = 0; j < NX + NU; j++) {
            Dual xd[NX], ud[NU], fd[NX];
            for (int i = 0; i < NX; i++) xd[i] = Dual(xe[i], i == j ? 1. : 0.);
            for (int i = 0; i < NU; i++) ud[i] = Dual(ue[i], NX + i == j ? 1. : 0.);
            M::template flow_map<Dual>(xd, ud, par, fd);
            for (int i = 0; i < NX; i++) { if (j < NX) Ac[(size_t)i * NX + j] = fd[i].d; else Bc[(size_t)i * NU + (j - NX)] = fd[i].d; f[i] = fd[i].v; }
        }
        const double ts = cfg.time_horizon / (K - 1);
        {
            const int n = NX + NU;
            Mat E((size_t)n * n, 0.);
            for (int i = 0; i < NX; i++) { for (int j = 0; j < NX; j++) E[(size_t)i * n + j] = Ac[(size_t)i * NX + j] * ts; for (int j = 0; j < NU; j++) E[(size_t)i * n + NX + j] = Bc[(size_t)i * NU + j] * ts; }
            const Mat X = expm(E, n);
            A.assign((size_t)NX * NX, 0.); B.assign((size_t)NX * NU, 0.);
            for (int i = 0; i < NX; i++) { for (int j = 0; j < NX; j++) A[(size_t)i * NX + j] = X[(size_t)i * n + j]; for (int j = 0; j < NU; j++) B[(size_t)i * NU + j] = X[(size_t)i * n + NX + j]; }
        }
        {
            const int n = NX + 1;
            Mat E((size_t)n * n, 0.);
            for (int i = 0; i < NX; i++) {
                double r = f[i];
                for (int j = 0; j < NX; j++) { E[(size_t)i * n + j] = Ac[(size_t)i * NX + j] * ts; r -= Ac[(size_t)i * NX + j] * xe[j]; }
                for (int j = 0; j < NU; j++) r -= Bc[(size_t)i * NU + j] * ue[j];
                E[(size_t)i * n + NX] = r * ts;
            }
            const Mat X = expm(E, n);
            z.assign(NX, 0.);
            for (int i = 0; i < NX; i++) z[i] = X[(size_t)i * n + NX];
        }
        // ---- condensation: x_k = Phi_k x0 + S_k U + zh_k
        std::vector<double> Phi((size_t)K * NX * NX, 0.), S((size_t)K * NX * nuu, 0.), zh((size_t)K * NX, 0.);
        for (int i = 0; i < NX; i++) Phi[(size_t)i * NX + i] = 1.;
        for (int k = 0; k + 1 < K; k++) {
            for (int i = 0; i < NX; i++) {
                for (int j = 0; j < NX; j++) { double a = 0; for (int q = 0; q < NX; q++) a += A[(size_t)i * NX + q] * Phi[((size_t)k * NX + q) * NX + j]; Phi[((size_t)(k + 1) * NX + i) * NX + j] = a; }
                for (int j = 0; j < nuu; j++) { double a = 0; for (int q = 0; q < NX; q++) a += A[(size_t)i * NX + q] * S[((size_t)k * NX + q) * nuu + j]; S[((size_t)(k + 1) * NX + i) * nuu + j] = a; }
                for (int j = 0; j < NU; j++) S[((size_t)(k + 1) * NX + i) * nuu + k * NU + j] += B[(size_t)i * NU + j];
                double a = z[i];
                for (int q = 0; q < NX; q++) a += A[(size_t)i * NX + q] * zh[(size_t)k * NX + q];
                zh[(size_t)(k + 1) * NX + i] = a;
            }
        }
        // ---- rows.  y = (U, error_cost, input_cost); a row is  s = h - G y  with  h = hc + Hx x_init + Hf x_final
        const int nv = nuu + 2, ie = nuu, ic = nuu + 1;
        std::vector<double> G, hc, Hx, Hf;
        std::vector<int> cdim;
        int nr = 0;
        auto add_row = [&]() { G.resize((size_t)(nr + 1) * nv, 0.); hc.push_back(0.); Hx.resize((size_t)(nr + 1) * NX, 0.); Hf.resize((size_t)(nr + 1) * NX, 0.); return nr++; };
        // one row of the model's table (models.cuh: s = h - sum coef xi[idx], xi = (x_k, u_k)) at node / step k; returns false for a row that
        // mixes states and inputs (none in the models with an operating point)
        auto model_row = [&](int r, int k, bool state_rows) -> int {
            const RowDesc rd = M::row(r);
            bool has_x = false, has_u = false;
            for (int q = 0; q < rd.n; q++) (rd.idx[q] < NX ? has_x : has_u) = true;
            if (has_x && has_u) return -1;
            if (rd.n == 0 || has_x != state_rows) return 0;      // (constant rows, e.g. the head of |T| <= T_max, are added by the caller where their cone lives)
            const int row = add_row();
            hc[row] = cst[rd.hs];
            for (int q = 0; q < rd.n; q++) {
                const double cf = cst[rd.cs[q]];
                const int idx = rd.idx[q];
                if (idx < NX) {
                    for (int j = 0; j < nuu; j++) G[(size_t)row * nv + j] += cf * S[((size_t)k * NX + idx) * nuu + j];
                    hc[row] -= cf * zh[(size_t)k * NX + idx];
                    for (int j = 0; j < NX; j++) Hx[(size_t)row * NX + j] -= cf * Phi[((size_t)k * NX + idx) * NX + j];
                } else G[(size_t)row * nv + k * NU + (idx - NX)] += cf;
            }
            return 1;
        };
        for (int r = 0; r < M::NLP; r++) { const RowDesc rd = M::row(r); for (int q = 0; q < rd.n; q++) if (rd.cs[q] < 0) { err = "linearised rows (thrust direction) are not part of the MPC path"; return SCPP_B200_ERR_UNSUPPORTED; } }
        // LP rows: states at nodes 1..K-1 (x_0 is the given state), inputs at steps 0..K-2
        for (int k = 1; k < K; k++) for (int r = 0; r < M::NLP; r++) if (model_row(r, k, true) < 0) { err = "mixed state/input row"; return SCPP_B200_ERR_UNSUPPORTED; }
        for (int k = 0; k < NS; k++) for (int r = 0; r < M::NLP; r++) if (model_row(r, k, false) < 0) { err = "mixed state/input row"; return SCPP_B200_ERR_UNSUPPORTED; }
        const int nl = nr;
        // model cones: a cone is a state cone or an input cone by the rows that have entries
        for (int c = 0; c < M::NCONE; c++) {
            const int o = M::NLP + M::cone_off(c), d = M::cone_dim(c);
            bool is_state = false, is_input = false;
            for (int t = 0; t < d; t++) { const RowDesc rd = M::row(o + t); for (int q = 0; q < rd.n; q++) (rd.idx[q] < NX ? is_state : is_input) = true; }
            if (is_state && is_input) { err = "mixed state/input cone"; return SCPP_B200_ERR_UNSUPPORTED; }
            for (int k = is_state ? 1 : 0; k < (is_state ? K : NS); k++) {
                for (int t = 0; t < d; t++) {
                    const RowDesc rd = M::row(o + t);
                    if (rd.n == 0) { const int row = add_row(); hc[row] = cst[rd.hs]; }
                    else if (model_row(o + t, k, is_state) != 1) { err = "cone row"; return SCPP_B200_ERR_UNSUPPORTED; }
                }
                cdim.push_back(d);
            }
        }
        // error cost (MPCProblem.cpp:61-73, intermediate_cost_active = false):  | W_T (X_{K-1} - x_final) | <= error_cost
        {
            int row = add_row();
            G[(size_t)row * nv + ie] = -1.;
            for (int i = 0; i < NX; i++) {
                row = add_row();
                const double wt = cfg.state_weights_terminal[i];
                for (int j = 0; j < nuu; j++) G[(size_t)row * nv + j] = -wt * S[((size_t)(K - 1) * NX + i) * nuu + j];
                hc[row] = wt * zh[(size_t)(K - 1) * NX + i];
                for (int j = 0; j < NX; j++) Hx[(size_t)row * NX + j] = wt * Phi[((size_t)(K - 1) * NX + i) * NX + j];
                Hf[(size_t)row * NX + i] = -wt;
            }
            cdim.push_back(1 + NX);
        }
        // input cost (:79-86):  | (W_u u_0, ..., W_u u_{K-2}) | <= input_cost
        {
            int row = add_row();
            G[(size_t)row * nv + ic] = -1.;
            for (int k = 0; k < NS; k++) for (int j = 0; j < NU; j++) { row = add_row(); G[(size_t)row * nv + k * NU + j] = -cfg.input_weights[j]; }
            cdim.push_back(1 + nuu);
        }
        std::vector<double> c(nv, 0.);
        c[ie] = 1.; c[ic] = 1.;                                                   // socp->addCostTerm(v_error_cost), (v_input_cost)
        dp.nv = nv; dp.nl = nl; dp.ncones = (int)cdim.size(); dp.nr = nr;
        int rc;
        if ((rc = upload(G, &dp.G)) || (rc = upload(c, &dp.c)) || (rc = upload(hc, &dp.hc)) || (rc = upload(Hx, &dp.Hx)) || (rc = upload(Hf, &dp.Hf)) ||
            (rc = upload(Phi, &dp.Phi)) || (rc = upload(S, &dp.S)) || (rc = upload(zh, &dp.zh)) || (rc = upload(cdim, &dp.cdim))) return rc;
        std::vector<double> hp(par, par + M::NP);
        const double *dpar = nullptr;
        if ((rc = upload(hp, &dpar))) return rc;
        d_par = const_cast<double *>(dpar);
        return 0;
    }
    int init() override
    {
        MCU(cudaSetDevice(device));
        MCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto &e : ev) MCU(cudaEventCreate(&e));
        std::string err;
        int rc = build(err);
        if (rc) return err.empty() ? rc : scpp_b200_fail(rc, "scpp_b200_mpc_create: " + err);
        const int K = cfg.K;
        if ((rc = dalloc(&d_x0, (size_t)N * NX)) || (rc = dalloc(&d_xf, (size_t)N * NX)) || (rc = dalloc(&d_X, (size_t)N * K * NX)) || (rc = dalloc(&d_U, (size_t)N * (K - 1) * NU)) ||
            (rc = dalloc(&d_status, (size_t)N)) || (rc = dalloc(&d_iters, (size_t)N)) || (rc = dalloc(&d_xs, (size_t)N * NX))) return rc;
        return 0;
    }
    int set_states(const double *xi, const double *xf) override
    {
        MCU(cudaSetDevice(device));
        if (xi) MCU(cudaMemcpyAsync(d_x0, xi, (size_t)N * NX * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (xf) MCU(cudaMemcpyAsync(d_xf, xf, (size_t)N * NX * sizeof(double), cudaMemcpyHostToDevice, stream));
        MCU(cudaStreamSynchronize(stream));
        have_states = true;
        return 0;
    }
    int solve() override
    {
        if (!have_states) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_solve: states not set");
        MCU(cudaSetDevice(device));
        MCU(cudaEventRecord(ev[0], stream));
        const int T = 64, grid = (N + T - 1) / T;
        if (cfg.K <= 8) k_mpc<M, 8><<<grid, T, 0, stream>>>(dp, cfg.K, cfg.ipm, N, d_x0, d_xf, d_X, d_U, d_status, d_iters);
        else k_mpc<M, 21><<<grid, T, 0, stream>>>(dp, cfg.K, cfg.ipm, N, d_x0, d_xf, d_X, d_U, d_status, d_iters);
        MCU(cudaEventRecord(ev[1], stream));
        MCU(cudaStreamSynchronize(stream));
        MCU(cudaGetLastError());
        float ms = 0;
        MCU(cudaEventElapsedTime(&ms, ev[0], ev[1]));
        ms_solve = ms;
        solved = true;
        return 0;
    }
    int get_solution(double *X, double *U, int *status, int *iters) override
    {
        if (!solved) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_get_solution: no solution yet");
        MCU(cudaSetDevice(device));
        const int K = cfg.K;
        if (X) MCU(cudaMemcpyAsync(X, d_X, (size_t)N * K * NX * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (U) MCU(cudaMemcpyAsync(U, d_U, (size_t)N * (K - 1) * NU * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (status) MCU(cudaMemcpyAsync(status, d_status, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (iters) MCU(cudaMemcpyAsync(iters, d_iters, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        MCU(cudaStreamSynchronize(stream));
        return 0;
    }
    int sim_step(double dt, double *x_new) override
    {
        if (!solved) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_sim_step: no solution yet");
        if (!(dt > 0.)) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_sim_step: dt must be positive");
        MCU(cudaSetDevice(device));
        k_mpc_step<M><<<(N + 63) / 64, 64, 0, stream>>>(N, cfg.K, dt, d_x0, d_U, d_par, d_xs);
        if (x_new) MCU(cudaMemcpyAsync(x_new, d_xs, (size_t)N * NX * sizeof(double), cudaMemcpyDeviceToHost, stream));
        MCU(cudaStreamSynchronize(stream));
        MCU(cudaGetLastError());
        return 0;
    }
};

extern "C" {

int scpp_b200_load_mpc_info(const char *path, int model, scpp_b200_mpc_config *c)
{
    int nx, nu, np_;
    if (scpp_b200_model_dims(model, &nx, &nu, &np_)) return SCPP_B200_ERR_ARG;
    try {      // MPCAlgorithm::loadParameters, MPCAlgorithm.cpp:17-32
        ParameterServer ps(path);
        memset(c, 0, sizeof(*c));
        bool nd, cd, ic;
        ps.loadScalar("K", c->K);
        ps.loadScalar("nondimensionalize", nd); ps.loadScalar("constant_dynamics", cd); ps.loadScalar("intermediate_cost_active", ic);
        ps.loadScalar("time_horizon", c->time_horizon);
        ps.loadMatrix("state_weights_intermediate", c->state_weights_intermediate, nx);
        ps.loadMatrix("state_weights_terminal", c->state_weights_terminal, nx);
        ps.loadMatrix("input_weights", c->input_weights, nu);
        c->nondimensionalize = nd; c->constant_dynamics = cd; c->intermediate_cost_active = ic;
        c->ipm.feastol = 1e-8; c->ipm.abstol = 1e-8; c->ipm.reltol = 1e-8; c->ipm.maxit = 100; c->ipm.warm = 0.;
    } catch (const std::exception &ex) { return scpp_b200_fail(SCPP_B200_ERR_IO, ex.what()); }
    return 0;
}

int scpp_b200_mpc_create(int model, const scpp_b200_model_params *params, const scpp_b200_mpc_config *cfg, int n, int device, scpp_b200_mpc **out)
{
    if (!params || !cfg || !out || n <= 0) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_create: bad argument");
    if (cfg->K < 2 || cfg->K > 21 || !(cfg->time_horizon > 0.)) return scpp_b200_fail(SCPP_B200_ERR_ARG, "scpp_b200_mpc_create: 2 <= K <= 21 and time_horizon > 0 required");
    if (cfg->intermediate_cost_active || !cfg->constant_dynamics || cfg->nondimensionalize)
        return scpp_b200_fail(SCPP_B200_ERR_UNSUPPORTED, "only intermediate_cost_active = false, constant_dynamics = true, nondimensionalize = false (the shipped MPC.info) are built");
    if (params->constrain_initial_final)
        return scpp_b200_fail(SCPP_B200_ERR_UNSUPPORTED, "constrain_initial_final = true pins X_0 to the MODEL's x_init and X_{K-1} to x_final (rocket2d.cpp:54-59): "
                                                         "with the MPC horizon that problem is infeasible after the first step; run MPC with it disabled");
    if (scpp_b200_device_count() <= 0) return scpp_b200_fail(SCPP_B200_ERR_CUDA, "no CUDA device: libscpp_b200 has no CPU execution path");
    scpp_b200_mpc *e = nullptr;
    if (model == SCPP_B200_MODEL_ROCKET2D) e = new MpcEngineT<Rocket2d>();
    else if (model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) e = new MpcEngineT<Rocket2dPlugin>();
    else if (model == SCPP_B200_MODEL_ROCKETQUAT) return scpp_b200_fail(SCPP_B200_ERR_UNSUPPORTED, "RocketQuat has no operating point (getOperatingPoint is not overridden in the reference: it throws)");
    else return scpp_b200_fail(SCPP_B200_ERR_ARG, "unknown model");
    e->model = model; e->N = n; e->device = device;
    memcpy(&e->P, params, sizeof(ModelParamsHost));
    memcpy(&e->cfg, cfg, sizeof(MpcConfig));
    int rc = e->init();
    if (rc) { delete e; return rc; }
    *out = e;
    return 0;
}
void scpp_b200_mpc_destroy(scpp_b200_mpc *e) { delete e; }
int scpp_b200_mpc_set_states(scpp_b200_mpc *e, const double *xi, const double *xf) { return e ? e->set_states(xi, xf) : scpp_b200_fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_mpc_solve(scpp_b200_mpc *e) { return e ? e->solve() : scpp_b200_fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_mpc_get_solution(scpp_b200_mpc *e, double *X, double *U, int *status, int *iters) { return e ? e->get_solution(X, U, status, iters) : scpp_b200_fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_mpc_sim_step(scpp_b200_mpc *e, double dt, double *x_new) { return e ? e->sim_step(dt, x_new) : scpp_b200_fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_mpc_get_discretization(scpp_b200_mpc *e, double *A, double *B, double *z)
{
    if (!e) return scpp_b200_fail(SCPP_B200_ERR_ARG, "null engine");
    if (A) memcpy(A, e->A.data(), e->A.size() * sizeof(double));
    if (B) memcpy(B, e->B.data(), e->B.size() * sizeof(double));
    if (z) memcpy(z, e->z.data(), e->z.size() * sizeof(double));
    return 0;
}
double scpp_b200_mpc_last_ms(scpp_b200_mpc *e) { return e ? e->ms_solve : -1.; }

} // extern "C"
