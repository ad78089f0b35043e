// scpp_b200/csrc/models.cuh — the model plugins (reference: scpp_models/) as compile-time traits.
//
// The reference selects ONE model at compile time (scpp_core/include/activeModel.hpp:6-10) and obtains
// f / A / B through CppAD code generation (scpp_core/include/systemDynamics.hpp:110-235).  Here each model is a
// struct with (1) the flow map written once for a generic scalar (the plugin surface, systemFlowMap),
// (2) a hand-derived sparse Jacobian used by the device code (tests check it against forward-mode dual numbers
// of (1) and against the oracle), (3) the application constraints (addApplicationConstraints) as a row table,
// (4) nondimensionalisation and the initial guess.
#pragma once
#include "portable.cuh"

namespace scpp {

// ------------------------------------------------------------------------------------------------
// Forward-mode dual number (value + ONE tangent): the device counterpart of the reference's AD scalar
// (CppAD::AD<CppAD::cg::CG<double>>, scpp_core/include/systemDynamics.hpp:34-39).  Instantiating a model's generic-scalar flow map
// (systemFlowMap, :69-73) with it gives f(x,u) and the directional derivative J (dx, du) in one pass; the multiple-shooting kernel needs
// exactly one such product per right-hand-side evaluation and column (discretize.cuh), so no Jacobian is ever formed and a model
// plugs in with nothing but its flow map.
// ------------------------------------------------------------------------------------------------
struct Dual {
    double v, d;
    SCPP_HD Dual() : v(0.), d(0.) {}
    SCPP_HD explicit Dual(double v_) : v(v_), d(0.) {}
    SCPP_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};
SCPP_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
SCPP_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
SCPP_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
SCPP_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
SCPP_HD Dual operator/(Dual a, Dual b) { const double ib = 1. / b.v, q = a.v * ib; return Dual(q, (a.d - q * b.d) * ib); }   // one division
SCPP_HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
SCPP_HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
SCPP_HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
SCPP_HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
SCPP_HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
SCPP_HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
SCPP_HD Dual operator/(Dual a, double b) { const double ib = 1. / b; return Dual(a.v * ib, a.d * ib); }
SCPP_HD Dual operator/(double a, Dual b) { const double ib = 1. / b.v, q = a * ib; return Dual(q, -q * b.d * ib); }
// (the overloads below would hide ::sqrt / ::sin / ::cos for unqualified calls inside this namespace: re-declare the double versions here)
using ::sqrt;
using ::sin;
using ::cos;
SCPP_HD Dual sqrt(Dual a) { const double r = ::sqrt(a.v); return Dual(r, a.d / (2. * r)); }
SCPP_HD Dual sin(Dual a) { return Dual(::sin(a.v), ::cos(a.v) * a.d); }
SCPP_HD Dual cos(Dual a) { return Dual(::cos(a.v), -::sin(a.v) * a.d); }

// ------------------------------------------------------------------------------------------------
// per-instance problem description as uploaded by the host (dimensional, angles in rad)
// mirrors RocketQuat::Parameters (scpp_models/include/rocketQuat.hpp:50-85) /
//         Rocket2d::Parameters  (scpp_models/include/rocket2d.hpp:51-84)
// ------------------------------------------------------------------------------------------------
struct ModelParamsHost {
    double g_I[3];
    double J_B[3];
    double r_T_B[3];
    double alpha_m;      // RocketQuat: 1/(I_sp |g_z|)         Rocket2d: unused
    double m;            // Rocket2d: vehicle mass             RocketQuat: unused
    double T_min, T_max, t_max;
    double gimbal_max, theta_max, gamma_gs, w_B_max;
    double final_time;
    int exact_minimum_thrust;
    int enable_roll_control;
    int constrain_initial_final;
    int pad_;
};

constexpr int MAX_CST = 12;

// one row of a constraint  s_r = h_r - sum_j coef_j * xi[idx_j]   (xi = [x ; u] of one node)
// coef_j = cst[cs_j] for cs_j >= 0,  = -tdir[-cs_j-1] for cs_j < 0 (linearised minimum-thrust direction)
struct RowDesc {
    signed char n;
    signed char idx[3];
    signed char cs[3];
    signed char hs;
};

// ---- constraint row tables (real constant data: a function-local table would be rebuilt on the stack at every call) ----
#define SCPP_RQ_ROWS {                                                                                          \
        {1, {0, 0, 0}, {2, 0, 0}, 3},          /* m_k - m_dry >= 0                          rocketQuat.cpp:93      */ \
        {3, {14, 15, 16}, {-1, -2, -3}, 4},    /* n_k' T_k - T_min >= 0                     :113-121               */ \
        {1, {3, 0, 0}, {5, 0, 0}, 0},          /* glide slope: tan(gamma) r_z               :96-97                 */ \
        {1, {1, 0, 0}, {2, 0, 0}, 0},                                                                             \
        {1, {2, 0, 0}, {2, 0, 0}, 0},                                                                             \
        {0, {0, 0, 0}, {0, 0, 0}, 6},          /* tilt: sqrt((1-cos theta_max)/2)           :100-101               */ \
        {1, {8, 0, 0}, {2, 0, 0}, 0},                                                                             \
        {1, {9, 0, 0}, {2, 0, 0}, 0},                                                                             \
        {0, {0, 0, 0}, {0, 0, 0}, 7},          /* |w| <= w_B_max                            :104-105               */ \
        {1, {11, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {12, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {13, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {0, {0, 0, 0}, {0, 0, 0}, 8},          /* |T| <= T_max                              :129                   */ \
        {1, {14, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {15, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {16, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {16, 0, 0}, {9, 0, 0}, 0},         /* gimbal: tan(gimbal_max) T_z               :132-133               */ \
        {1, {14, 0, 0}, {2, 0, 0}, 0},                                                                            \
        {1, {15, 0, 0}, {2, 0, 0}, 0},                                                                            \
    }
#define SCPP_R2D_ROWS {                                                                                         \
        {1, {4, 0, 0}, {1, 0, 0}, 3}, {1, {4, 0, 0}, {2, 0, 0}, 3},    /* |eta| <= theta_max          rocket2d.cpp:66-68 */ \
        {1, {5, 0, 0}, {1, 0, 0}, 4}, {1, {5, 0, 0}, {2, 0, 0}, 4},    /* |w| <= w_B_max              :70-72 */ \
        {1, {6, 0, 0}, {1, 0, 0}, 5}, {1, {6, 0, 0}, {2, 0, 0}, 5},    /* |gimbal| <= gimbal_max      :76-78 */ \
        {1, {7, 0, 0}, {2, 0, 0}, 6}, {1, {7, 0, 0}, {1, 0, 0}, 7},    /* T_min <= T <= T_max         :80-82 */ \
        {1, {1, 0, 0}, {8, 0, 0}, 0}, {1, {0, 0, 0}, {2, 0, 0}, 0},    /* |r_x| <= tan(gamma) r_y     :63-64 */ \
    }

// ================================= RocketQuat =====================================================
struct RocketQuat {
    static constexpr int NX = 14, NU = 4, NP = 10;
    static constexpr int NLP = 2;                 // mass, (linearised) minimum thrust
    static constexpr int NCONE = 5;               // glide slope, tilt, angular rate, max thrust, gimbal
    static constexpr int NCR = 17;                // their total dimension
    static constexpr int MAXDIM = 4;
    static constexpr const char *name = "RocketQuat";

    SCPP_HD static int cone_dim(int c) { return (c == 2 || c == 3) ? 4 : 3; }                       // {3,3,4,4,3}
    SCPP_HD static int cone_off(int c) { return 3 * c + (c > 2 ? 1 : 0) + (c > 3 ? 1 : 0); }         // {0,3,6,10,14}

    // rocketQuat.cpp:70-144 (exact_minimum_thrust handled through tdir: (0,0,1) reproduces T_z >= T_min, :125;
    // enable_roll_control == false: X.row(13)==0 and U.row(3)==0 become fixed variables, :141-142)
    SCPP_HD static RowDesc row(int r);
    // the same table as a constant expression: with a compile-time row index (unrolled loops) it folds into immediates
    SCPP_HD static constexpr RowDesc crow(int r) { constexpr RowDesc t[NLP + NCR] = SCPP_RQ_ROWS; return t[r]; }

    // ---- systemFlowMap for a generic scalar (rocketQuat.cpp:7-37); par = [alpha_m, g_I, J_B, r_T_B]
    template <class T>
    SCPP_HD static void flow_map(const T *x, const T *u, const double *par, T *f)
    {
        const T m = x[0];
        const T qw = x[7], qx = x[8], qy = x[9], qz = x[10];
        const T wx = x[11], wy = x[12], wz = x[13];
        const T Tx = u[0], Ty = u[1], Tz = u[2];
        const T R00 = 1. - 2. * (qy * qy + qz * qz), R01 = 2. * (qx * qy - qw * qz), R02 = 2. * (qx * qz + qw * qy);
        const T R10 = 2. * (qx * qy + qw * qz), R11 = 1. - 2. * (qx * qx + qz * qz), R12 = 2. * (qy * qz - qw * qx);
        const T R20 = 2. * (qx * qz - qw * qy), R21 = 2. * (qy * qz + qw * qx), R22 = 1. - 2. * (qx * qx + qy * qy);
        f[0] = -par[0] * sqrt(Tx * Tx + Ty * Ty + Tz * Tz);
        f[1] = x[4]; f[2] = x[5]; f[3] = x[6];
        f[4] = (R00 * Tx + R01 * Ty + R02 * Tz) / m + par[1];
        f[5] = (R10 * Tx + R11 * Ty + R12 * Tz) / m + par[2];
        f[6] = (R20 * Tx + R21 * Ty + R22 * Tz) / m + par[3];
        f[7] = 0.5 * (-wx * qx - wy * qy - wz * qz);
        f[8] = 0.5 * (wx * qw + wz * qy - wy * qz);
        f[9] = 0.5 * (wy * qw - wz * qx + wx * qz);
        f[10] = 0.5 * (wz * qw + wy * qx - wx * qy);
        // J^-1 (r_T x T + torque) - w x w ; the last term is identically zero (rocketQuat.cpp:36)
        f[11] = (par[8] * Tz - par[9] * Ty) / par[4];
        f[12] = (par[9] * Tx - par[7] * Tz) / par[5];
        f[13] = (par[7] * Ty - par[8] * Tx + u[3]) / par[6];
    }

    // ---- evaluated dynamics + sparse Jacobian: what the variational equations need
    struct Lin {
        double f[NX];
        double am[3];     // d vdot / d m     = -R T / m^2
        double Dq[3][4];  // d vdot / d q     = (1/m) d(R T)/dq
        double Rm[3][3];  // d vdot / d T     = R / m
        double hw[3];     // 0.5 * w
        double hq[4];     // 0.5 * q
        double b0[3];     // d mdot / d T     = -alpha_m T / |T|
        double iJ[3];     // 1 / J_B
        double rT[3];
    };
    SCPP_HD static void linearize(const double *x, const double *u, const double *par, Lin &L)
    {
        const double m = x[0], im = 1. / m;
        const double w = x[7], qx = x[8], qy = x[9], qz = x[10];
        const double Tx = u[0], Ty = u[1], Tz = u[2];
        const double R00 = 1. - 2. * (qy * qy + qz * qz), R01 = 2. * (qx * qy - w * qz), R02 = 2. * (qx * qz + w * qy);
        const double R10 = 2. * (qx * qy + w * qz), R11 = 1. - 2. * (qx * qx + qz * qz), R12 = 2. * (qy * qz - w * qx);
        const double R20 = 2. * (qx * qz - w * qy), R21 = 2. * (qy * qz + w * qx), R22 = 1. - 2. * (qx * qx + qy * qy);
        const double RT0 = R00 * Tx + R01 * Ty + R02 * Tz, RT1 = R10 * Tx + R11 * Ty + R12 * Tz, RT2 = R20 * Tx + R21 * Ty + R22 * Tz;
        const double nT = sqrt(Tx * Tx + Ty * Ty + Tz * Tz);
        L.f[0] = -par[0] * nT;
        L.f[1] = x[4]; L.f[2] = x[5]; L.f[3] = x[6];
        L.f[4] = RT0 * im + par[1]; L.f[5] = RT1 * im + par[2]; L.f[6] = RT2 * im + par[3];
        L.hw[0] = 0.5 * x[11]; L.hw[1] = 0.5 * x[12]; L.hw[2] = 0.5 * x[13];
        L.hq[0] = 0.5 * w; L.hq[1] = 0.5 * qx; L.hq[2] = 0.5 * qy; L.hq[3] = 0.5 * qz;
        L.f[7] = -L.hw[0] * qx - L.hw[1] * qy - L.hw[2] * qz;
        L.f[8] = L.hw[0] * w + L.hw[2] * qy - L.hw[1] * qz;
        L.f[9] = L.hw[1] * w - L.hw[2] * qx + L.hw[0] * qz;
        L.f[10] = L.hw[2] * w + L.hw[1] * qx - L.hw[0] * qy;
        L.iJ[0] = 1. / par[4]; L.iJ[1] = 1. / par[5]; L.iJ[2] = 1. / par[6];
        L.rT[0] = par[7]; L.rT[1] = par[8]; L.rT[2] = par[9];
        L.f[11] = (par[8] * Tz - par[9] * Ty) * L.iJ[0];
        L.f[12] = (par[9] * Tx - par[7] * Tz) * L.iJ[1];
        L.f[13] = (par[7] * Ty - par[8] * Tx + u[3]) * L.iJ[2];
        const double im2 = im * im;
        L.am[0] = -RT0 * im2; L.am[1] = -RT1 * im2; L.am[2] = -RT2 * im2;
        const double s = 2. * im;
        L.Dq[0][0] = s * (-qz * Ty + qy * Tz);           L.Dq[1][0] = s * (qz * Tx - qx * Tz);            L.Dq[2][0] = s * (-qy * Tx + qx * Ty);
        L.Dq[0][1] = s * (qy * Ty + qz * Tz);            L.Dq[1][1] = s * (qy * Tx - 2 * qx * Ty - w * Tz); L.Dq[2][1] = s * (qz * Tx + w * Ty - 2 * qx * Tz);
        L.Dq[0][2] = s * (-2 * qy * Tx + qx * Ty + w * Tz); L.Dq[1][2] = s * (qx * Tx + qz * Tz);          L.Dq[2][2] = s * (-w * Tx + qz * Ty - 2 * qy * Tz);
        L.Dq[0][3] = s * (-2 * qz * Tx - w * Ty + qx * Tz); L.Dq[1][3] = s * (w * Tx - 2 * qz * Ty + qy * Tz); L.Dq[2][3] = s * (qx * Tx + qy * Ty);
        L.Rm[0][0] = R00 * im; L.Rm[0][1] = R01 * im; L.Rm[0][2] = R02 * im;
        L.Rm[1][0] = R10 * im; L.Rm[1][1] = R11 * im; L.Rm[1][2] = R12 * im;
        L.Rm[2][0] = R20 * im; L.Rm[2][1] = R21 * im; L.Rm[2][2] = R22 * im;
        const double a = -par[0] / nT;
        L.b0[0] = a * Tx; L.b0[1] = a * Ty; L.b0[2] = a * Tz;
    }
    // out = A v  (42 structural non-zeros)
    SCPP_HD static void A_apply(const Lin &L, const double *v, double *o)
    {
        o[0] = 0.;
        o[1] = v[4]; o[2] = v[5]; o[3] = v[6];
#pragma unroll
        for (int r = 0; r < 3; r++)
            o[4 + r] = L.am[r] * v[0] + L.Dq[r][0] * v[7] + L.Dq[r][1] * v[8] + L.Dq[r][2] * v[9] + L.Dq[r][3] * v[10];
        const double qw = L.hq[0], qx = L.hq[1], qy = L.hq[2], qz = L.hq[3];
        const double wx = L.hw[0], wy = L.hw[1], wz = L.hw[2];
        o[7]  = -wx * v[8] - wy * v[9] - wz * v[10] - qx * v[11] - qy * v[12] - qz * v[13];
        o[8]  =  wx * v[7] + wz * v[9] - wy * v[10] + qw * v[11] - qz * v[12] + qy * v[13];
        o[9]  =  wy * v[7] - wz * v[8] + wx * v[10] + qz * v[11] + qw * v[12] - qx * v[13];
        o[10] =  wz * v[7] + wy * v[8] - wx * v[9]  - qy * v[11] + qx * v[12] + qw * v[13];
        o[11] = 0.; o[12] = 0.; o[13] = 0.;
    }
    // out = B w  (w in R^4)
    SCPP_HD static void B_apply(const Lin &L, const double *w, double *o)
    {
        o[0] = L.b0[0] * w[0] + L.b0[1] * w[1] + L.b0[2] * w[2];
        o[1] = 0.; o[2] = 0.; o[3] = 0.;
#pragma unroll
        for (int r = 0; r < 3; r++) o[4 + r] = L.Rm[r][0] * w[0] + L.Rm[r][1] * w[1] + L.Rm[r][2] * w[2];
        o[7] = 0.; o[8] = 0.; o[9] = 0.; o[10] = 0.;
        o[11] = (L.rT[1] * w[2] - L.rT[2] * w[1]) * L.iJ[0];
        o[12] = (L.rT[2] * w[0] - L.rT[0] * w[2]) * L.iJ[1];
        o[13] = (L.rT[0] * w[1] - L.rT[1] * w[0] + w[3]) * L.iJ[2];
    }

    // ---- per-instance setup (K0): Parameters::nondimensionalize rocketQuat.cpp:291-312,
    //      getNewModelParameters :168-173, updateProblemParameters :156-166
    SCPP_HD static void setup(const ModelParamsHost &P, int nondim, double *xi /*in: x_init, out: scaled*/,
                              double *xf, double *par, double *cst, double *scale /*[m_scale, r_scale]*/)
    {
        double ms = 1., rs = 1.;
        if (nondim) { ms = xi[0]; rs = sqrt(xi[1] * xi[1] + xi[2] * xi[2] + xi[3] * xi[3]); }
        scale[0] = ms; scale[1] = rs;
        par[0] = P.alpha_m * rs;
        for (int i = 0; i < 3; i++) { par[1 + i] = P.g_I[i] / rs; par[4 + i] = P.J_B[i] / (ms * rs * rs); par[7 + i] = P.r_T_B[i] / rs; }
        xi[0] /= ms; xf[0] /= ms;
        for (int i = 1; i < 7; i++) { xi[i] /= rs; xf[i] /= rs; }
        const double T_min = P.T_min / (ms * rs), T_max = P.T_max / (ms * rs);
        cst[0] = 0.; cst[1] = 1.; cst[2] = -1.;
        cst[3] = -xf[0];                                    // -m_dry (x_final(0), rocketQuat.cpp:93)
        cst[4] = -T_min;
        cst[5] = -tan(P.gamma_gs);
        cst[6] = sqrt((1. - cos(P.theta_max)) / 2.);
        cst[7] = P.w_B_max;
        cst[8] = T_max;
        cst[9] = -tan(P.gimbal_max);
        cst[10] = P.t_max / (ms * rs * rs);
        cst[11] = T_min;
    }
    // getInitializedTrajectory rocketQuat.cpp:39-68
    SCPP_HD static void initial_guess(const double *xi, const double *xf, const double *cst, int K, int k, double *x, double *u)
    {
        const double a1 = double(K - k) / K, a2 = double(k) / K;
        x[0] = a1 * xi[0] + a2 * xf[0];
        for (int i = 1; i < 7; i++) x[i] = a1 * xi[i] + a2 * xf[i];
        // Eigen slerp semantics
        const double one = 1.0 - 2.220446049250313e-16;
        double d = xi[7] * xf[7] + xi[8] * xf[8] + xi[9] * xf[9] + xi[10] * xf[10];
        double ad = fabs(d), s0, s1;
        if (ad >= one) { s0 = 1. - a2; s1 = a2; }
        else { double th = acos(ad), st = sin(th); s0 = sin((1. - a2) * th) / st; s1 = sin(a2 * th) / st; }
        if (d < 0) s1 = -s1;
        for (int i = 7; i < 11; i++) x[i] = s0 * xi[i] + s1 * xf[i];
        for (int i = 11; i < 14; i++) x[i] = a1 * xi[i] + a2 * xf[i];
        u[0] = 0.; u[1] = 0.; u[2] = (cst[8] - cst[11]) / 2.; u[3] = 0.;   // (T_max - T_min)/2  :64
    }
    // zero-order hold (interpolate_input = false): the reference's U has K - 1 columns; the engine keeps K and pins column K - 1 to this
    // strictly feasible placeholder, which appears in no dynamics row (sc.cuh: sc_zoh_pins)
    static constexpr bool ZOH = true;
    SCPP_HD static void zoh_placeholder_input(const double *cst, double *u) { u[0] = 0.; u[1] = 0.; u[2] = 0.5 * (cst[8] + cst[11]); u[3] = 0.; }
    // fixed variables of node k: bit i set => xi[i] is pinned to val[i]  (rocketQuat.cpp:79,83-89,109-111,141-142)
    SCPP_HD static uint32_t fixed(const ModelParamsHost &, const double *xi, const double *xf, int K, int k, double *val)
    {
        uint32_t mask = 0;
        for (int i = 0; i < NX + NU; i++) val[i] = 0.;
        mask |= 1u << 13; mask |= 1u << 17;                                   // w_z == 0, roll torque == 0
        if (k == 0) for (int i = 0; i < NX; i++) { mask |= 1u << i; val[i] = xi[i]; }
        if (k == K - 1) {
            const int fin[11] = {1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13};
            for (int q = 0; q < 11; q++) { mask |= 1u << fin[q]; val[fin[q]] = xf[fin[q]]; }
            mask |= (1u << 14) | (1u << 15) | (1u << 17);
        }
        return mask;
    }
    // redimensionalizeTrajectory rocketQuat.cpp:188-201
    SCPP_HD static void redim(const double *scale, double *x, double *u)
    {
        x[0] *= scale[0];
        for (int i = 1; i < 7; i++) x[i] *= scale[1];
        for (int i = 0; i < 3; i++) u[i] *= scale[0] * scale[1];
        u[3] *= scale[0] * scale[1] * scale[1];
    }
    // nondimensionalizeTrajectory rocketQuat.cpp:175-186
    SCPP_HD static void nondim(const double *scale, double *x, double *u)
    {
        x[0] /= scale[0];
        for (int i = 1; i < 7; i++) x[i] /= scale[1];
        for (int i = 0; i < 3; i++) u[i] /= scale[0] * scale[1];
        u[3] /= scale[0] * scale[1] * scale[1];
    }
    // getOperatingPoint is not overridden by RocketQuat: the base class throws (scpp_core/include/systemModel.hpp:119)
    SCPP_HD static bool operating_point(const ModelParamsHost &, double *, double *) { return false; }
    // linearised minimum-thrust direction: U0.head<3>().normalized() (rocketQuat.cpp:162-165)
    SCPP_HD static void thrust_dir(const double *u, double *d)
    {
        double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int i = 0; i < 3; i++) d[i] = n > 0 ? u[i] / n : u[i];
    }
};

// ================================= Rocket2d =======================================================
struct Rocket2d {
    static constexpr int NX = 6, NU = 2, NP = 6;
    static constexpr int NLP = 8;    // boxes on eta, omega, gimbal, thrust
    static constexpr int NCONE = 1;  // glide slope (norm of a 1-vector)
    static constexpr int NCR = 2;
    static constexpr int MAXDIM = 2;
    static constexpr const char *name = "Rocket2D";

    SCPP_HD static int cone_dim(int) { return 2; }
    SCPP_HD static int cone_off(int) { return 0; }
    // rocket2d.cpp:46-84
    SCPP_HD static RowDesc row(int r);
    // the same table as a constant expression: with a compile-time row index (unrolled loops) it folds into immediates
    SCPP_HD static constexpr RowDesc crow(int r) { constexpr RowDesc t[NLP + NCR] = SCPP_R2D_ROWS; return t[r]; }
    // rocket2d.cpp:7-40 ; par = [m, J_B, g_I(2), r_T_B(2)]
    template <class T>
    SCPP_HD static void flow_map(const T *x, const T *u, const double *par, T *f)
    {
        const T TBx = -sin(u[0]) * u[1], TBy = cos(u[0]) * u[1];
        const T ce = cos(x[4]), se = sin(x[4]);
        f[0] = x[2]; f[1] = x[3];
        f[2] = (ce * TBx - se * TBy) / par[0] + par[2];
        f[3] = (se * TBx + ce * TBy) / par[0] + par[3];
        f[4] = x[5];
        f[5] = (par[4] * TBy - par[5] * TBx) / par[1];
    }
    struct Lin {
        double f[NX];
        double a24, a34;             // d vdot / d eta
        double b[NX][NU];
    };
    SCPP_HD static void linearize(const double *x, const double *u, const double *par, Lin &L)
    {
        const double im = 1. / par[0], iJ = 1. / par[1];
        const double sa = sin(u[0]), ca = cos(u[0]), se = sin(x[4]), ce = cos(x[4]);
        const double TBx = -sa * u[1], TBy = ca * u[1];
        const double dxa = -ca * u[1], dya = -sa * u[1], dxm = -sa, dym = ca;
        L.f[0] = x[2]; L.f[1] = x[3];
        L.f[2] = (ce * TBx - se * TBy) * im + par[2];
        L.f[3] = (se * TBx + ce * TBy) * im + par[3];
        L.f[4] = x[5];
        L.f[5] = (par[4] * TBy - par[5] * TBx) * iJ;
        L.a24 = (-se * TBx - ce * TBy) * im;
        L.a34 = (ce * TBx - se * TBy) * im;
        for (int i = 0; i < NX; i++) { L.b[i][0] = 0.; L.b[i][1] = 0.; }
        L.b[2][0] = (ce * dxa - se * dya) * im; L.b[2][1] = (ce * dxm - se * dym) * im;
        L.b[3][0] = (se * dxa + ce * dya) * im; L.b[3][1] = (se * dxm + ce * dym) * im;
        L.b[5][0] = (par[4] * dya - par[5] * dxa) * iJ; L.b[5][1] = (par[4] * dym - par[5] * dxm) * iJ;
    }
    SCPP_HD static void A_apply(const Lin &L, const double *v, double *o)
    {
        o[0] = v[2]; o[1] = v[3]; o[2] = L.a24 * v[4]; o[3] = L.a34 * v[4]; o[4] = v[5]; o[5] = 0.;
    }
    SCPP_HD static void B_apply(const Lin &L, const double *w, double *o)
    {
        for (int i = 0; i < NX; i++) o[i] = L.b[i][0] * w[0] + L.b[i][1] * w[1];
    }
    // Rocket2d::Parameters::nondimensionalize rocket2d.cpp:198-214 ; getNewModelParameters :143-148
    SCPP_HD static void setup(const ModelParamsHost &P, int nondim, double *xi, double *xf, double *par, double *cst, double *scale)
    {
        double ms = 1., rs = 1.;
        if (nondim) { rs = sqrt(xi[0] * xi[0] + xi[1] * xi[1]); ms = P.m; }
        scale[0] = ms; scale[1] = rs;
        par[0] = P.m / ms; par[1] = P.J_B[0] / (ms * rs * rs);
        par[2] = P.g_I[0] / rs; par[3] = P.g_I[1] / rs; par[4] = P.r_T_B[0] / rs; par[5] = P.r_T_B[1] / rs;
        for (int i = 0; i < 4; i++) { xi[i] /= rs; xf[i] /= rs; }
        cst[0] = 0.; cst[1] = 1.; cst[2] = -1.;
        cst[3] = P.theta_max; cst[4] = P.w_B_max; cst[5] = P.gimbal_max;
        cst[6] = -P.T_min / (ms * rs); cst[7] = P.T_max / (ms * rs);
        cst[8] = -tan(P.gamma_gs);
        cst[9] = 0.; cst[10] = 0.; cst[11] = 0.;
    }
    // rocket2d.cpp:121-136 : (T_max + T_min)/2
    SCPP_HD static void initial_guess(const double *xi, const double *xf, const double *cst, int K, int k, double *x, double *u)
    {
        const double a1 = double(K - k) / K, a2 = double(k) / K;
        for (int i = 0; i < NX; i++) x[i] = a1 * xi[i] + a2 * xf[i];
        u[0] = 0.; u[1] = (cst[7] + (-cst[6])) / 2.;
    }
    static constexpr bool ZOH = true;
    SCPP_HD static void zoh_placeholder_input(const double *cst, double *u) { u[0] = 0.; u[1] = 0.5 * (cst[7] - cst[6]); }
    SCPP_HD static uint32_t fixed(const ModelParamsHost &P, const double *xi, const double *xf, int K, int k, double *val)
    {
        uint32_t mask = 0;
        for (int i = 0; i < NX + NU; i++) val[i] = 0.;
        if (P.constrain_initial_final) {                                        // rocket2d.cpp:53-59
            if (k == 0) for (int i = 0; i < NX; i++) { mask |= 1u << i; val[i] = xi[i]; }
            if (k == K - 1) { for (int i = 0; i < NX; i++) { mask |= 1u << i; val[i] = xf[i]; } mask |= 1u << 6; }
        }
        return mask;
    }
    SCPP_HD static void redim(const double *scale, double *x, double *u)   // rocket2d.cpp:109-119
    {
        for (int i = 0; i < 4; i++) x[i] *= scale[1];
        u[1] *= scale[0] * scale[1];
    }
    SCPP_HD static void nondim(const double *scale, double *x, double *u)  // rocket2d.cpp:97-107
    {
        for (int i = 0; i < 4; i++) x[i] /= scale[1];
        u[1] /= scale[0] * scale[1];
    }
    SCPP_HD static void thrust_dir(const double *, double *d) { d[0] = 0.; d[1] = 0.; d[2] = 1.; }
    // Rocket2d::getOperatingPoint (rocket2d.cpp:40-44): hover, x = 0, u = (gimbal 0, thrust |g| m)
    SCPP_HD static bool operating_point(const ModelParamsHost &P, double *x, double *u)
    {
        for (int i = 0; i < NX; i++) x[i] = 0.;
        u[0] = 0.; u[1] = -P.g_I[1] * P.m;
        return true;
    }
};

} // namespace scpp
// a model written only against the plugin surface (flow map + cvx:: constraints): Jacobians by dual numbers, tables generated at build time
#include "../plugins/rocket2d_plugin.hpp"
#include "../plugins/rocketquat_plugin.hpp"
namespace scpp {

static const RowDesc rq_rows_host[RocketQuat::NLP + RocketQuat::NCR] = SCPP_RQ_ROWS;
static const RowDesc r2d_rows_host[Rocket2d::NLP + Rocket2d::NCR] = SCPP_R2D_ROWS;
#if defined(__CUDACC__)
static __constant__ RowDesc rq_rows_dev[RocketQuat::NLP + RocketQuat::NCR] = SCPP_RQ_ROWS;
static __constant__ RowDesc r2d_rows_dev[Rocket2d::NLP + Rocket2d::NCR] = SCPP_R2D_ROWS;
#endif
SCPP_HD RowDesc RocketQuat::row(int r)
{
#if defined(__CUDA_ARCH__)
    return rq_rows_dev[r];
#else
    return rq_rows_host[r];
#endif
}
SCPP_HD RowDesc Rocket2d::row(int r)
{
#if defined(__CUDA_ARCH__)
    return r2d_rows_dev[r];
#else
    return r2d_rows_host[r];
#endif
}
#if !defined(SCPP_PLUGIN_GENERATE)
static const RowDesc r2dp_rows_host[Rocket2dPlugin::NLP + Rocket2dPlugin::NCR] = ROCKET2D_PLUGIN_ROWS;
#if defined(__CUDACC__)
static __constant__ RowDesc r2dp_rows_dev[Rocket2dPlugin::NLP + Rocket2dPlugin::NCR] = ROCKET2D_PLUGIN_ROWS;
#endif
SCPP_HD RowDesc Rocket2dPlugin::row(int r)
{
#if defined(__CUDA_ARCH__)
    return r2dp_rows_dev[r];
#else
    return r2dp_rows_host[r];
#endif
}
static const RowDesc rqrp_rows_host[RocketQuatRollPlugin::NLP + RocketQuatRollPlugin::NCR] = ROCKETQUAT_ROLL_PLUGIN_ROWS;
#if defined(__CUDACC__)
static __constant__ RowDesc rqrp_rows_dev[RocketQuatRollPlugin::NLP + RocketQuatRollPlugin::NCR] = ROCKETQUAT_ROLL_PLUGIN_ROWS;
#endif
SCPP_HD RowDesc RocketQuatRollPlugin::row(int r)
{
#if defined(__CUDA_ARCH__)
    return rqrp_rows_dev[r];
#else
    return rqrp_rows_host[r];
#endif
}
#endif

} // namespace scpp
