// tests/cvx_shim_test.cpp — TEST: the cvx:: recording shim (include/scpp_cvx.hpp) and the lowering (include/scpp_plugin.hpp).
//  1. RocketQuat's and Rocket2d's addApplicationConstraints, written here in the reference's DSL call by call
//     (scpp_models/src/rocketQuat.cpp:70-144, rocket2d.cpp:46-84), are recorded, lowered to the stage-wise table and compared with the
//     table the engine really uses (scpp_b200_model_rows through the C-ABI): same rows (as sets, per kind), same cone dimensions, same
//     pinned variables.  This is the check a maintainer runs when a model file changes.
//  2. the problem-builder subset (matrix product of dynpar blocks with variable columns, comma initialiser, norm, sum, cost terms:
//     scpp_core/src/SCProblem.cpp:16-134) records a small SC problem; violation() is ~0 at a point built to satisfy it.
// Built and run by tests/test_host.py::test_cvx_shim_lowers_the_reference_constraints.  Prints "ok" on success.
#include "scpp_b200.h"
#include "scpp_plugin.hpp"
#include <cstdio>
#include <cstring>
#include <map>
#include <set>

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

// ---- RocketQuat, as the reference writes it (member names of RocketQuat::Parameters / p_dyn kept) ----------------------------
struct RQ {
    static constexpr int NX = 14, NU = 4;
    struct { double x_init[NX], x_final[NX], T_min, T_max, t_max, w_B_max; bool exact_minimum_thrust = true, enable_roll_control = false; } p;
    struct { double gimbal_const, gs_const, tilt_const; std::vector<double> thrust_const; /* 3 x K */ } p_dyn;
    void addApplicationConstraints(std::shared_ptr<cvx::OptimizationProblem> socp, int K)
    {
        cvx::MatrixX v_X, v_U;
        socp->getVariable("X", v_X);
        socp->getVariable("U", v_U);
        socp->addConstraint(cvx::equalTo(v_X.col(0), cvx::dynpar(p.x_init)));                                             // initial state
        for (size_t i : {1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13})                                                            // final state: mass and roll free
            socp->addConstraint(cvx::equalTo(v_X(i, v_X.cols() - 1), cvx::dynpar(p.x_final[i])));
        socp->addConstraint(cvx::greaterThan(v_X.row(0), cvx::dynpar(p.x_final[0])));                                     // mass
        socp->addConstraint(cvx::lessThan(v_X.block(1, 0, 2, v_X.cols()).colwise().norm(),                               // glide slope
                                          cvx::dynpar(p_dyn.gs_const) * v_X.block(3, 0, 1, v_X.cols())));
        socp->addConstraint(cvx::lessThan(v_X.block(8, 0, 2, v_X.cols()).colwise().norm(), cvx::dynpar(p_dyn.tilt_const)));   // max tilt
        socp->addConstraint(cvx::lessThan(v_X.block(11, 0, 3, v_X.cols()).colwise().norm(), cvx::dynpar(p.w_B_max)));         // max rate
        socp->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(0), 0.));                                                // final input
        socp->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(1), 0.));
        socp->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(3), 0.));
        if (p.exact_minimum_thrust) {
            p_dyn.thrust_const.assign(3 * K, 0.);
            socp->addConstraint(cvx::greaterThan(cvx::dynpar(p_dyn.thrust_const.data(), 3, K).cwiseProduct(v_U.topRows(3)).colwise().sum(),   // linearised minimum thrust
                                                 cvx::dynpar(p.T_min)));
        } else
            socp->addConstraint(cvx::greaterThan(v_U.row(2), cvx::dynpar(p.T_min)));
        socp->addConstraint(cvx::lessThan(v_U.topRows(3).colwise().norm(), cvx::dynpar(p.T_max)));                        // maximum thrust
        socp->addConstraint(cvx::lessThan(v_U.topRows(2).colwise().norm(), cvx::dynpar(p_dyn.gimbal_const) * v_U.row(2))); // gimbal
        if (p.enable_roll_control)
            socp->addConstraint(cvx::box(-cvx::dynpar(p.t_max), v_U.row(3), cvx::dynpar(p.t_max)));
        else {
            socp->addConstraint(cvx::equalTo(v_X.row(13), 0.));
            socp->addConstraint(cvx::equalTo(v_U.row(3), 0.));
        }
    }
};

// ---- Rocket2d ------------------------------------------------------------------------------------------------------------------
struct R2D {
    static constexpr int NX = 6, NU = 2;
    struct { double x_init[NX], x_final[NX], T_min, T_max, gimbal_max, theta_max, w_B_max, tan_gamma_gs; bool constrain_initial_final = true; } p;
    void addApplicationConstraints(std::shared_ptr<cvx::OptimizationProblem> socp)
    {
        cvx::MatrixX v_X, v_U;
        socp->getVariable("X", v_X);
        socp->getVariable("U", v_U);
        if (p.constrain_initial_final) {
            socp->addConstraint(cvx::equalTo(cvx::dynpar(p.x_init), v_X.col(0)));
            socp->addConstraint(cvx::equalTo(cvx::dynpar(p.x_final), v_X.rightCols(1)));
            socp->addConstraint(cvx::equalTo(v_U(0, v_U.cols() - 1), 0.));
        }
        socp->addConstraint(cvx::lessThan(v_X.row(0).colwise().norm(), cvx::dynpar(p.tan_gamma_gs) * v_X.row(1)));
        socp->addConstraint(cvx::box(-cvx::dynpar(p.theta_max), v_X.row(4), cvx::dynpar(p.theta_max)));
        socp->addConstraint(cvx::box(-cvx::dynpar(p.w_B_max), v_X.row(5), cvx::dynpar(p.w_B_max)));
        socp->addConstraint(cvx::box(-cvx::dynpar(p.gimbal_max), v_U.row(0), cvx::dynpar(p.gimbal_max)));
        socp->addConstraint(cvx::box(cvx::dynpar(p.T_min), v_U.row(1), cvx::dynpar(p.T_max)));
    }
};

// a lowered row evaluated to numbers: (index -> coefficient, h); the minimum-thrust direction coefficients become NaN like in the C-ABI
typedef std::pair<std::map<int, double>, double> NumRow;
static NumRow numeric(const scpp_plugin::Row &r, const double *constants, const double *xi, const double *xf)
{
    NumRow n;
    for (auto &e : r.c) n.first[e.first] = e.second.kind == scpp_plugin::Source::NODE_ARRAY ? NAN : e.second.value(constants, xi, xf, nullptr, 0, 0);
    n.second = r.h.value(constants, xi, xf, nullptr, 0, 0);
    return n;
}
static bool same_row(const NumRow &a, const NumRow &b)
{
    if (a.first.size() != b.first.size() || std::fabs(a.second - b.second) > 1e-12 * (1 + std::fabs(b.second))) return false;
    for (auto &e : a.first) {
        auto it = b.first.find(e.first);
        if (it == b.first.end()) return false;
        if (std::isnan(e.second) != std::isnan(it->second)) return false;
        if (!std::isnan(e.second) && std::fabs(e.second - it->second) > 1e-12 * (1 + std::fabs(e.second))) return false;
    }
    return true;
}
// engine table (C-ABI) vs lowered table: LP rows as a set, cones in order with rows in order
static void compare_with_engine(int model, const scpp_b200_model_params &P, const double *xi, const double *xf, const scpp_plugin::StageTable &t,
                                const double *constants)
{
    double rows[64 * 8]; int nlp = 0, nc = 0, dims[16];
    CHECK(scpp_b200_model_rows(model, &P, xi, xf, 64, rows, &nlp, &nc, dims) == 0);
    auto eng = [&](int r) { NumRow n; const double *o = rows + 8 * r; for (int q = 0; q < (int)o[0]; q++) n.first[(int)o[1 + q]] = o[4 + q]; n.second = o[7]; return n; };
    CHECK(nlp == (int)t.lp.size() && nc == (int)t.cones.size());
    std::vector<bool> used(nlp, false);
    for (auto &r : t.lp) {
        bool found = false;
        for (int e = 0; e < nlp && !found; e++) if (!used[e] && same_row(numeric(r, constants, xi, xf), eng(e))) { used[e] = true; found = true; }
        CHECK(found);
    }
    int at = nlp;
    for (int c = 0; c < nc && c < (int)t.cones.size(); c++) {
        CHECK(dims[c] == (int)t.cones[c].size());
        for (int i = 0; i < dims[c] && i < (int)t.cones[c].size(); i++) CHECK(same_row(numeric(t.cones[c][i], constants, xi, xf), eng(at + i)));
        at += dims[c];
    }
}

int main()
{
    const int K = 6;
    // ---------------- RocketQuat ----------------
    {
        RQ m;
        scpp_b200_model_params P; memset(&P, 0, sizeof(P));
        P.T_min = 2e5; P.T_max = 4.2e5; P.t_max = 0.; P.w_B_max = 1.05; P.gimbal_max = 0.26; P.theta_max = 1.57; P.gamma_gs = 0.52; P.alpha_m = 3.7e-4;
        P.g_I[2] = -9.81; P.J_B[0] = P.J_B[1] = 5e6; P.J_B[2] = 7e4; P.r_T_B[2] = -15.; P.exact_minimum_thrust = 1;
        double xi[14] = {24000, 200, 200, 800, -40, -40, -80, 0.97, -0.17, 0.17, 0.03, 0, 0, 0}, xf[14] = {22000, 0, 0, 0, 0, 0, -1, 1, 0, 0, 0, 0, 0, 0};
        memcpy(m.p.x_init, xi, sizeof(xi)); memcpy(m.p.x_final, xf, sizeof(xf));
        m.p.T_min = P.T_min; m.p.T_max = P.T_max; m.p.t_max = P.t_max; m.p.w_B_max = P.w_B_max;
        m.p_dyn.gimbal_const = std::tan(P.gimbal_max); m.p_dyn.gs_const = std::tan(P.gamma_gs); m.p_dyn.tilt_const = std::sqrt((1. - std::cos(P.theta_max)) / 2.);   // updateProblemParameters, :156-160
        auto socp = std::make_shared<cvx::OptimizationProblem>();
        socp->addVariable("X", 14, K); socp->addVariable("U", 4, K);
        m.addApplicationConstraints(socp, K);
        // the dynpars point into two structs: register them as one constant region each side by treating the struct as an array of doubles
        scpp_plugin::Lowering L;
        L.constants = {&m.p.T_min, 4};                                   // T_min, T_max, t_max, w_B_max
        L.x_init = {m.p.x_init, 14}; L.x_final = {m.p.x_final, 14};
        L.node_array = {m.p_dyn.thrust_const.data(), 3 * K}; L.node_rows = 3;
        scpp_plugin::StageTable t;
        bool threw = false;
        try { t = L.lower(*socp); } catch (const std::exception &e) { threw = true; }
        CHECK(threw);                                                    // p_dyn.* is outside the registered regions: reported, not guessed
        // the usual arrangement: one constant block
        struct { double T_min, T_max, t_max, w_B_max, gimbal_const, gs_const, tilt_const; } c = {m.p.T_min, m.p.T_max, m.p.t_max, m.p.w_B_max, m.p_dyn.gimbal_const, m.p_dyn.gs_const, m.p_dyn.tilt_const};
        RQ m2 = m;
        auto socp2 = std::make_shared<cvx::OptimizationProblem>();
        socp2->addVariable("X", 14, K); socp2->addVariable("U", 4, K);
        {   // same calls with the parameters read from the block c
            cvx::MatrixX v_X, v_U; socp2->getVariable("X", v_X); socp2->getVariable("U", v_U);
            socp2->addConstraint(cvx::equalTo(v_X.col(0), cvx::dynpar(m2.p.x_init)));
            for (size_t i : {1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13}) socp2->addConstraint(cvx::equalTo(v_X(i, v_X.cols() - 1), cvx::dynpar(m2.p.x_final[i])));
            socp2->addConstraint(cvx::greaterThan(v_X.row(0), cvx::dynpar(m2.p.x_final[0])));
            socp2->addConstraint(cvx::lessThan(v_X.block(1, 0, 2, v_X.cols()).colwise().norm(), cvx::dynpar(c.gs_const) * v_X.block(3, 0, 1, v_X.cols())));
            socp2->addConstraint(cvx::lessThan(v_X.block(8, 0, 2, v_X.cols()).colwise().norm(), cvx::dynpar(c.tilt_const)));
            socp2->addConstraint(cvx::lessThan(v_X.block(11, 0, 3, v_X.cols()).colwise().norm(), cvx::dynpar(c.w_B_max)));
            socp2->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(0), 0.));
            socp2->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(1), 0.));
            socp2->addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(3), 0.));
            m2.p_dyn.thrust_const.assign(3 * K, 0.);
            socp2->addConstraint(cvx::greaterThan(cvx::dynpar(m2.p_dyn.thrust_const.data(), 3, K).cwiseProduct(v_U.topRows(3)).colwise().sum(), cvx::dynpar(c.T_min)));
            socp2->addConstraint(cvx::lessThan(v_U.topRows(3).colwise().norm(), cvx::dynpar(c.T_max)));
            socp2->addConstraint(cvx::lessThan(v_U.topRows(2).colwise().norm(), cvx::dynpar(c.gimbal_const) * v_U.row(2)));
            socp2->addConstraint(cvx::equalTo(v_X.row(13), 0.));
            socp2->addConstraint(cvx::equalTo(v_U.row(3), 0.));
        }
        scpp_plugin::Lowering L2;
        L2.constants = {&c.T_min, 7}; L2.x_init = {m2.p.x_init, 14}; L2.x_final = {m2.p.x_final, 14};
        L2.node_array = {m2.p_dyn.thrust_const.data(), 3 * K}; L2.node_rows = 3;
        t = L2.lower(*socp2);
        CHECK(t.lp.size() == 2 && t.cones.size() == 5 && t.cone_rows() == 17 && t.max_cone_dim() == 4);
        // m_dry is x_final(0): the engine keeps it as a constant slot, the recorded row reads it from x_final -> both evaluate to the same number
        compare_with_engine(SCPP_B200_MODEL_ROCKETQUAT, P, xi, xf, t, &c.T_min);
        // pinned variables: w_z and the roll torque everywhere; the whole state at node 0; 11 final states + 2 more inputs at the last node
        auto idx = [](const std::vector<scpp_plugin::Pin> &l) { std::set<int> s; for (auto &q : l) s.insert(q.idx); return s; };
        CHECK(idx(t.pin_all) == (std::set<int>{13, 17}));
        CHECK(idx(t.pin_first).size() == 14);
        CHECK(idx(t.pin_last) == (std::set<int>{1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13, 14, 15}));      // the final roll torque == 0 is the same pin as U.row(3) == 0: listed once, under 'all'
        // the emitted table fits RowDesc and MAX_CST
        const scpp_plugin::Emitted e = scpp_plugin::emit_inc(t, "RQ_CHECK");
        CHECK(e.nlp == 2 && e.ncone == 5 && e.ncr == 17 && e.ncst <= 12);
    }
    // ---------------- Rocket2d ----------------
    {
        R2D m;
        scpp_b200_model_params P; memset(&P, 0, sizeof(P));
        P.m = 24000.; P.J_B[0] = 5e6; P.g_I[1] = -9.81; P.r_T_B[1] = -15.; P.T_min = 1e4; P.T_max = 4.2e5; P.gimbal_max = 0.26; P.theta_max = 1.05; P.gamma_gs = 0.79; P.w_B_max = 0.35;
        P.constrain_initial_final = 1;
        double xi[6] = {-200, 800, 0, -100, -0.35, 0}, xf[6] = {0, 0, 0, -1, 0, 0};
        memcpy(m.p.x_init, xi, sizeof(xi)); memcpy(m.p.x_final, xf, sizeof(xf));
        m.p.T_min = P.T_min; m.p.T_max = P.T_max; m.p.gimbal_max = P.gimbal_max; m.p.theta_max = P.theta_max; m.p.w_B_max = P.w_B_max; m.p.tan_gamma_gs = std::tan(P.gamma_gs);
        auto socp = std::make_shared<cvx::OptimizationProblem>();
        socp->addVariable("X", 6, K); socp->addVariable("U", 2, K);
        m.addApplicationConstraints(socp);
        scpp_plugin::Lowering L;
        L.constants = {&m.p.T_min, 6}; L.x_init = {m.p.x_init, 6}; L.x_final = {m.p.x_final, 6};
        const scpp_plugin::StageTable t = L.lower(*socp);
        CHECK(t.lp.size() == 8 && t.cones.size() == 1 && t.cones[0].size() == 2);
        compare_with_engine(SCPP_B200_MODEL_ROCKET2D, P, xi, xf, t, &m.p.T_min);
        compare_with_engine(SCPP_B200_MODEL_ROCKET2D_PLUGIN, P, xi, xf, t, &m.p.T_min);      // the generated table of the plugin model
        CHECK(t.pin_all.empty() && t.pin_first.size() == 6 && t.pin_last.size() == 7);
        // a feasible point satisfies the recorded problem, an infeasible one is measured
        std::vector<double> x(socp->numVariables(), 0.);
        for (int k = 0; k < K; k++) {
            for (int i = 0; i < 6; i++) x[socp->var("X").offset + 6 * k + i] = k == 0 ? xi[i] : (k == K - 1 ? xf[i] : 0.);
            x[socp->var("X").offset + 6 * k + 1] = k == 0 ? xi[1] : (k == K - 1 ? 0. : 500.);
            x[socp->var("U").offset + 2 * k + 1] = 2e5;
        }
        CHECK(socp->violation(x) < 1e-12);
        x[socp->var("U").offset + 1] = 5e5;                                                      // above T_max
        CHECK(std::fabs(socp->violation(x) - 8e4) < 1e-6);
        // something the stage-wise tables cannot express is reported
        socp->addConstraint(cvx::lessThan(socp->var("X").offset == 0 ? cvx::MatrixX(cvx::Affine::variable(0)) + cvx::MatrixX(cvx::Affine::variable(6)) : cvx::MatrixX(), 1.));
        bool threw = false;
        try { scpp_plugin::Lowering L3 = L; L3.lower(*socp); } catch (const std::exception &) { threw = true; }
        CHECK(threw);
    }
    // ---------------- the builder subset: dynamics rows, trust region, virtual control (SCProblem.cpp:16-134) ----------------
    {
        const int nx = 2, nu = 1, Kb = 3;
        auto socp = std::make_shared<cvx::OptimizationProblem>();
        cvx::MatrixX v_X = socp->addVariable("X", nx, Kb), v_U = socp->addVariable("U", nu, Kb), v_nu = socp->addVariable("nu", nx, Kb - 1);
        cvx::MatrixX v_nu_bound = socp->addVariable("nu_bound", nx, Kb - 1);
        cvx::Scalar v_norm1_nu = socp->addVariable("norm1_nu"), v_sigma = socp->addVariable("sigma");
        cvx::VectorX v_delta = socp->addVariable("delta", Kb);
        double weight_time = 1., weight_vc = 1e3, weight_tr = 2.;
        std::vector<std::vector<double>> A(Kb - 1, {1, 0, 0.5, 1}), B(Kb - 1, {0.125, 0.5}), C(Kb - 1, {0.1, 0.2}), s(Kb - 1, {0.3, -0.1}), z(Kb - 1, {0.01, 0.02});   // column-major
        std::vector<std::vector<double>> Xb(Kb, {1., -1.}), Ub(Kb, {0.5});
        socp->addCostTerm(cvx::dynpar(weight_time) * v_sigma);
        socp->addConstraint(cvx::greaterThan(v_sigma, 0.001));
        for (int k = 0; k < Kb - 1; k++) {
            cvx::VectorX lhs = cvx::dynpar(A[k].data(), nx, nx) * v_X.col(k) + cvx::dynpar(B[k].data(), nx, nu) * v_U.col(k) + cvx::dynpar(z[k]);
            lhs += cvx::dynpar(C[k].data(), nx, nu) * v_U.col(k + 1);
            lhs += cvx::dynpar(s[k]) * v_sigma;
            lhs += v_nu.col(k);
            socp->addConstraint(cvx::equalTo(lhs, v_X.col(k + 1)));
        }
        socp->addConstraint(cvx::box(-v_nu_bound, v_nu, v_nu_bound));
        socp->addConstraint(cvx::lessThan(v_nu_bound.sum(), v_norm1_nu));
        socp->addCostTerm(cvx::dynpar(weight_vc) * v_norm1_nu);
        for (int k = 0; k < Kb; k++) {
            cvx::VectorX norm2_terms(nx + nu);
            norm2_terms << cvx::dynpar(Xb[k]) - v_X.col(k), cvx::dynpar(Ub[k]) - v_U.col(k);
            socp->addConstraint(cvx::lessThan(norm2_terms.norm(), v_delta(k)));
        }
        socp->addCostTerm(cvx::dynpar(weight_tr) * v_delta.sum());
        CHECK(socp->numVariables() == nx * Kb + nu * Kb + 2 * nx * (Kb - 1) + 2 + Kb);
        // a point: X, U = linearisation point, sigma = 2, nu = the defect, bounds tight, delta = 0
        std::vector<double> x(socp->numVariables(), 0.);
        const double sigma = 2.;
        auto X = [&](int k, int i) -> double & { return x[socp->var("X").offset + nx * k + i]; };
        auto U = [&](int k) -> double & { return x[socp->var("U").offset + k]; };
        for (int k = 0; k < Kb; k++) { X(k, 0) = 1.; X(k, 1) = -1.; U(k) = 0.5; }
        x[socp->var("sigma").offset] = sigma;
        double n1 = 0;
        for (int k = 0; k < Kb - 1; k++)
            for (int i = 0; i < nx; i++) {
                double lhs = z[k][i] + s[k][i] * sigma + B[k][i] * U(k) + C[k][i] * U(k + 1);
                for (int j = 0; j < nx; j++) lhs += A[k][j * nx + i] * X(k, j);
                const double nu_ki = X(k + 1, i) - lhs;
                x[socp->var("nu").offset + nx * k + i] = nu_ki; x[socp->var("nu_bound").offset + nx * k + i] = std::fabs(nu_ki); n1 += std::fabs(nu_ki);
            }
        x[socp->var("norm1_nu").offset] = n1;
        CHECK(socp->violation(x) < 1e-12);
        CHECK(std::fabs(socp->cost.evaluate(x) - (sigma + 1e3 * n1)) < 1e-9);
        weight_vc = 10.;                                                     // dynpar: re-read at evaluation, no re-recording
        CHECK(std::fabs(socp->cost.evaluate(x) - (sigma + 10. * n1)) < 1e-9);
        X(1, 0) += 0.25;                                                     // breaks two dynamics rows and leaves the trust region
        CHECK(socp->violation(x) > 0.2);
        CHECK(socp->getVariableValue("sigma", x)[0] == sigma);
    }
    if (!fails) printf("ok\n");
    return fails ? 1 : 0;
}
