// scpp_b200/plugins/rocketquat_plugin.hpp — the reference's 6-DoF rocket (scpp_models/src/rocketQuat.cpp) written ONLY against the plugin
// surface, with ROLL CONTROL ENABLED (enable_roll_control = true, rocketQuat.cpp:135-138: box on the roll torque, w_z and the
// torque free) — the variant the hand-written table of models.cuh does not cover.  As for rocket2d_plugin.hpp nothing is derived by hand:
// Jacobians by dual numbers over systemFlowMap, the constraint table / pins / constant slots generated at build time from
// addApplicationConstraints recorded through the cvx:: shim (csrc/gen/rocketquat_roll_plugin.inc).
// Model id SCPP_B200_MODEL_ROCKETQUAT_ROLL; parity: the oracle's RocketQuat with enable_roll_control = 1 (tests).
#pragma once
#include "plugin_support.hpp"
#if !defined(SCPP_PLUGIN_GENERATE)
#include "../csrc/gen/rocketquat_roll_plugin.inc"
#endif

namespace scpp {

struct RocketQuatRollPlugin : AutoJacobian<RocketQuatRollPlugin, 14, 4> {
    static constexpr int NX = 14, NU = 4, NP = 10;
    static constexpr const char *name = "RocketQuatRollPlugin";
    // the constants addApplicationConstraints refers to: RocketQuat::Parameters fields and p_dyn (rocketQuat.hpp:50-95)
    enum { C_T_MIN, C_T_MAX, C_TORQUE_MAX, C_W_B_MAX, C_GIMBAL_CONST, C_GS_CONST, C_TILT_CONST, NCONST };

    // RocketQuat::systemFlowMap (rocketQuat.cpp:7-37); par = [alpha_m, g_I(3), J_B(3), r_T_B(3)] (getNewModelParameters, :168-173)
    template <class T>
    SCPP_HD static void flow_map(const T *x, const T *u, const double *par, T *f)
    {
        const T m = x[0];
        const T q_w = x[7], q_x = x[8], q_y = x[9], q_z = x[10];
        const T w_x = x[11], w_y = x[12], w_z = x[13];
        const T T_x = u[0], T_y = u[1], T_z = u[2], torque = u[3];
        // Quaternion(w, x, y, z).toRotationMatrix() without normalisation (:29-30)
        const T R00 = 1. - 2. * (q_y * q_y + q_z * q_z), R01 = 2. * (q_x * q_y - q_w * q_z), R02 = 2. * (q_x * q_z + q_w * q_y);
        const T R10 = 2. * (q_x * q_y + q_w * q_z), R11 = 1. - 2. * (q_x * q_x + q_z * q_z), R12 = 2. * (q_y * q_z - q_w * q_x);
        const T R20 = 2. * (q_x * q_z - q_w * q_y), R21 = 2. * (q_y * q_z + q_w * q_x), R22 = 1. - 2. * (q_x * q_x + q_y * q_y);
        f[0] = -par[0] * sqrt(T_x * T_x + T_y * T_y + T_z * T_z);                       // mass depletion
        f[1] = x[4]; f[2] = x[5]; f[3] = x[6];                                          // position
        f[4] = (R00 * T_x + R01 * T_y + R02 * T_z) / m + par[1];                        // velocity
        f[5] = (R10 * T_x + R11 * T_y + R12 * T_z) / m + par[2];
        f[6] = (R20 * T_x + R21 * T_y + R22 * T_z) / m + par[3];
        f[7] = 0.5 * (-w_x * q_x - w_y * q_y - w_z * q_z);                              // 0.5 * Omega(w) * q  (common.hpp:124-134)
        f[8] = 0.5 * (w_x * q_w + w_z * q_y - w_y * q_z);
        f[9] = 0.5 * (w_y * q_w - w_z * q_x + w_x * q_z);
        f[10] = 0.5 * (w_z * q_w + w_y * q_x - w_x * q_y);
        f[11] = (par[8] * T_z - par[9] * T_y) / par[4];                                 // J^-1 (r_T x T + torque) - w x w, the last term == 0 (:36)
        f[12] = (par[9] * T_x - par[7] * T_z) / par[5];
        f[13] = (par[7] * T_y - par[8] * T_x + torque) / par[6];
    }
    // Parameters::nondimensionalize (rocketQuat.cpp:291-312) + getNewModelParameters (:168-173) + updateProblemParameters (:156-160)
    SCPP_HD static void parameters(const ModelParamsHost &P, int nondim, double *xi, double *xf, double *par, double *constants, double *scale)
    {
        double m_scale = 1., r_scale = 1.;
        if (nondim) { m_scale = xi[0]; r_scale = sqrt(xi[1] * xi[1] + xi[2] * xi[2] + xi[3] * xi[3]); }
        scale[0] = m_scale; scale[1] = r_scale;
        par[0] = P.alpha_m * r_scale;
        for (int i = 0; i < 3; i++) { par[1 + i] = P.g_I[i] / r_scale; par[4 + i] = P.J_B[i] / (m_scale * r_scale * r_scale); par[7 + i] = P.r_T_B[i] / r_scale; }
        xi[0] /= m_scale; xf[0] /= m_scale;
        for (int i = 1; i < 7; i++) { xi[i] /= r_scale; xf[i] /= r_scale; }
        constants[C_T_MIN] = P.T_min / (m_scale * r_scale); constants[C_T_MAX] = P.T_max / (m_scale * r_scale);
        constants[C_TORQUE_MAX] = P.t_max / (m_scale * r_scale * r_scale);
        constants[C_W_B_MAX] = P.w_B_max;
        constants[C_GIMBAL_CONST] = tan(P.gimbal_max); constants[C_GS_CONST] = tan(P.gamma_gs);
        constants[C_TILT_CONST] = sqrt((1. - cos(P.theta_max)) / 2.);
    }
    // getInitializedTrajectory (rocketQuat.cpp:39-68): alpha2 = k / K, Eigen slerp on the quaternion, U = (0, 0, (T_max - T_min) / 2, 0)
    SCPP_HD static void initial_trajectory(const double *xi, const double *xf, const double *constants, int K, int k, double *x, double *u)
    {
        const double alpha1 = double(K - k) / K, alpha2 = double(k) / K;
        for (int i = 0; i < 7; i++) x[i] = alpha1 * xi[i] + alpha2 * xf[i];
        const double one = 1.0 - 2.220446049250313e-16;
        const double d = xi[7] * xf[7] + xi[8] * xf[8] + xi[9] * xf[9] + xi[10] * xf[10], ad = fabs(d);
        double s0, s1;
        if (ad >= one) { s0 = 1. - alpha2; s1 = alpha2; }
        else { const double theta = acos(ad), st = sin(theta); s0 = sin((1. - alpha2) * theta) / st; s1 = sin(alpha2 * theta) / st; }
        if (d < 0) s1 = -s1;
        for (int i = 7; i < 11; i++) x[i] = s0 * xi[i] + s1 * xf[i];
        for (int i = 11; i < 14; i++) x[i] = alpha1 * xi[i] + alpha2 * xf[i];
        u[0] = 0.; u[1] = 0.; u[2] = (constants[C_T_MAX] - constants[C_T_MIN]) / 2.; u[3] = 0.;
    }
    // redimensionalizeTrajectory / nondimensionalizeTrajectory (rocketQuat.cpp:175-201)
    SCPP_HD static void redim(const double *scale, double *x, double *u)
    {
        x[0] *= scale[0];
        for (int i = 1; i < 7; i++) x[i] *= scale[1];
        for (int i = 0; i < 3; i++) u[i] *= scale[0] * scale[1];
        u[3] *= scale[0] * scale[1] * scale[1];
    }
    SCPP_HD static void nondim(const double *scale, double *x, double *u)
    {
        x[0] /= scale[0];
        for (int i = 1; i < 7; i++) x[i] /= scale[1];
        for (int i = 0; i < 3; i++) u[i] /= scale[0] * scale[1];
        u[3] /= scale[0] * scale[1] * scale[1];
    }
    SCPP_HD static bool operating_point(const ModelParamsHost &, double *, double *) { return false; }      // not overridden by RocketQuat: the base class throws
    // updateProblemParameters (:162-165): thrust_const.col(k) = U0[k].head<3>().normalized()
    SCPP_HD static void thrust_dir(const double *u, double *d)
    {
        const double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int i = 0; i < 3; i++) d[i] = n > 0 ? u[i] / n : u[i];
    }

#if defined(SCPP_PLUGIN_HOST)
    // addApplicationConstraints in the reference's DSL (rocketQuat.cpp:70-144 with exact_minimum_thrust = true, enable_roll_control = true);
    // c = the constant block, thrust_const = the 3 x K array updateProblemParameters refreshes
    static constexpr bool uses_node_array = true;
    static void addApplicationConstraints(cvx::OptimizationProblem &socp, const double *c, const double *x_init, const double *x_final, const double *thrust_const)
    {
        cvx::MatrixX v_X, v_U;
        socp.getVariable("X", v_X);
        socp.getVariable("U", v_U);
        socp.addConstraint(cvx::equalTo(v_X.col(0), cvx::dynpar(x_init, NX)));                                           // initial state
        for (size_t i : {1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13})                                                           // final state: mass and roll free
            socp.addConstraint(cvx::equalTo(v_X(i, v_X.cols() - 1), cvx::dynpar(x_final[i])));
        socp.addConstraint(cvx::greaterThan(v_X.row(0), cvx::dynpar(x_final[0])));                                       // mass
        socp.addConstraint(cvx::lessThan(v_X.block(1, 0, 2, v_X.cols()).colwise().norm(),                              // glide slope
                                         cvx::dynpar(c[C_GS_CONST]) * v_X.block(3, 0, 1, v_X.cols())));
        socp.addConstraint(cvx::lessThan(v_X.block(8, 0, 2, v_X.cols()).colwise().norm(), cvx::dynpar(c[C_TILT_CONST])));    // max tilt
        socp.addConstraint(cvx::lessThan(v_X.block(11, 0, 3, v_X.cols()).colwise().norm(), cvx::dynpar(c[C_W_B_MAX])));     // max rate
        socp.addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(0), 0.));                                               // final input
        socp.addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(1), 0.));
        socp.addConstraint(cvx::equalTo(v_U.col(v_U.cols() - 1)(3), 0.));
        socp.addConstraint(cvx::greaterThan(cvx::dynpar(thrust_const, 3, v_U.cols()).cwiseProduct(v_U.topRows(3)).colwise().sum(),   // linearised minimum thrust
                                            cvx::dynpar(c[C_T_MIN])));
        socp.addConstraint(cvx::lessThan(v_U.topRows(3).colwise().norm(), cvx::dynpar(c[C_T_MAX])));                     // maximum thrust
        socp.addConstraint(cvx::lessThan(v_U.topRows(2).colwise().norm(), cvx::dynpar(c[C_GIMBAL_CONST]) * v_U.row(2)));  // gimbal
        socp.addConstraint(cvx::box(-cvx::dynpar(c[C_TORQUE_MAX]), v_U.row(3), cvx::dynpar(c[C_TORQUE_MAX])));           // roll control (:135-138)
    }
#endif

#if !defined(SCPP_PLUGIN_GENERATE)
    SCPP_PLUGIN_MEMBERS(ROCKETQUAT_ROLL_PLUGIN)
#endif
};

} // namespace scpp
