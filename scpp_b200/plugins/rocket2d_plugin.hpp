// scpp_b200/plugins/rocket2d_plugin.hpp — the reference's planar rocket (scpp_models/src/rocket2d.cpp) written ONLY against the plugin
// surface of the reference: systemFlowMap for a generic scalar, getInitializedTrajectory, (non/re)dimensionalisation,
// getOperatingPoint, and addApplicationConstraints in the cvx:: DSL.  No Jacobian and no row table are written by hand:
//   * A, B come from forward-mode dual numbers over the flow map (AutoJacobian below; K1 uses the same duals directly),
//     the role CppAD plays in the reference (scpp_core/include/systemDynamics.hpp:110-235);
//   * the stage-wise constraint table, the pinned-variable lists and the constant slots are GENERATED at build time:
//     tools/gen_plugin.cpp records addApplicationConstraints through include/scpp_cvx.hpp, include/scpp_plugin.hpp lowers it,
//     the result is csrc/gen/rocket2d_plugin.inc (scpp_b200/build.py runs the generator before nvcc).
// Model id SCPP_B200_MODEL_ROCKET2D_PLUGIN; tests run it against the oracle's Rocket2D and against the hand-written Rocket2d traits.
#pragma once
#include "plugin_support.hpp"
#if !defined(SCPP_PLUGIN_GENERATE)
#include "../csrc/gen/rocket2d_plugin.inc"
#endif

namespace scpp {

struct Rocket2dPlugin : AutoJacobian<Rocket2dPlugin, 6, 2> {
    static constexpr int NX = 6, NU = 2, NP = 6;
    static constexpr const char *name = "Rocket2DPlugin";
    // the constants addApplicationConstraints refers to: fields of Rocket2d::Parameters (scpp_models/include/rocket2d.hpp:51-84)
    enum { C_THETA_MAX, C_W_B_MAX, C_GIMBAL_MAX, C_T_MIN, C_T_MAX, C_TAN_GAMMA_GS, NCONST };

    // Rocket2d::systemFlowMap (rocket2d.cpp:7-40); par = [m, J_B, g_I(2), r_T_B(2)] (getNewModelParameters, :143-148)
    template <class T>
    SCPP_HD static void flow_map(const T *x, const T *u, const double *par, T *f)
    {
        const T gimbal = u[0], thrust = u[1], eta = x[4];
        const T T_Bx = -sin(gimbal) * thrust, T_By = cos(gimbal) * thrust;       // thrust in the body frame
        const T c = cos(eta), s = sin(eta);
        f[0] = x[2];
        f[1] = x[3];
        f[2] = (c * T_Bx - s * T_By) / par[0] + par[2];
        f[3] = (s * T_Bx + c * T_By) / par[0] + par[3];
        f[4] = x[5];
        f[5] = (par[4] * T_By - par[5] * T_Bx) / par[1];
    }
    // Parameters::nondimensionalize (rocket2d.cpp:198-214) + getNewModelParameters (:143-148): scaled boundary states, the dynamics
    // parameters and the constant block of the constraints
    SCPP_HD static void parameters(const ModelParamsHost &P, int nondim, double *xi, double *xf, double *par, double *constants, double *scale)
    {
        double m_scale = 1., r_scale = 1.;
        if (nondim) { r_scale = sqrt(xi[0] * xi[0] + xi[1] * xi[1]); m_scale = P.m; }
        scale[0] = m_scale; scale[1] = r_scale;
        par[0] = P.m / m_scale; par[1] = P.J_B[0] / (m_scale * r_scale * r_scale);
        par[2] = P.g_I[0] / r_scale; par[3] = P.g_I[1] / r_scale;
        par[4] = P.r_T_B[0] / r_scale; par[5] = P.r_T_B[1] / r_scale;
        for (int i = 0; i < 4; i++) { xi[i] /= r_scale; xf[i] /= r_scale; }
        constants[C_THETA_MAX] = P.theta_max; constants[C_W_B_MAX] = P.w_B_max; constants[C_GIMBAL_MAX] = P.gimbal_max;
        constants[C_T_MIN] = P.T_min / (m_scale * r_scale); constants[C_T_MAX] = P.T_max / (m_scale * r_scale);
        constants[C_TAN_GAMMA_GS] = tan(P.gamma_gs);
    }
    // getInitializedTrajectory (rocket2d.cpp:121-136)
    SCPP_HD static void initial_trajectory(const double *xi, const double *xf, const double *constants, int K, int k, double *x, double *u)
    {
        const double alpha1 = double(K - k) / K, alpha2 = double(k) / K;
        for (int i = 0; i < NX; i++) x[i] = alpha1 * xi[i] + alpha2 * xf[i];
        u[0] = 0.; u[1] = (constants[C_T_MAX] + constants[C_T_MIN]) / 2.;
    }
    // redimensionalizeTrajectory / nondimensionalizeTrajectory (rocket2d.cpp:97-119)
    SCPP_HD static void redim(const double *scale, double *x, double *u) { for (int i = 0; i < 4; i++) x[i] *= scale[1]; u[1] *= scale[0] * scale[1]; }
    SCPP_HD static void nondim(const double *scale, double *x, double *u) { for (int i = 0; i < 4; i++) x[i] /= scale[1]; u[1] /= scale[0] * scale[1]; }
    // getOperatingPoint (rocket2d.cpp:40-44)
    SCPP_HD static bool operating_point(const ModelParamsHost &P, double *x, double *u)
    {
        for (int i = 0; i < NX; i++) x[i] = 0.;
        u[0] = 0.; u[1] = -P.g_I[1] * P.m;
        return true;
    }
    SCPP_HD static void thrust_dir(const double *, double *d) { d[0] = 0.; d[1] = 0.; d[2] = 1.; }

#if defined(SCPP_PLUGIN_HOST)
    // addApplicationConstraints in the reference's DSL (rocket2d.cpp:46-84 with constrain_initial_final = true, the setting of the shipped
    // model.info); c = the constant block, x_init / x_final = the boundary states the dynpars point at
    static constexpr bool uses_node_array = false;
    static void addApplicationConstraints(cvx::OptimizationProblem &socp, const double *c, const double *x_init, const double *x_final)
    {
        cvx::MatrixX v_X, v_U;
        socp.getVariable("X", v_X);
        socp.getVariable("U", v_U);
        // initial and final state, final gimbal angle
        socp.addConstraint(cvx::equalTo(cvx::dynpar(x_init, NX), v_X.col(0)));
        socp.addConstraint(cvx::equalTo(cvx::dynpar(x_final, NX), v_X.rightCols(1)));
        socp.addConstraint(cvx::equalTo(v_U(0, v_U.cols() - 1), 0.));
        // glide slope
        socp.addConstraint(cvx::lessThan(v_X.row(0).colwise().norm(), cvx::dynpar(c[C_TAN_GAMMA_GS]) * v_X.row(1)));
        // tilt angle, angular rate
        socp.addConstraint(cvx::box(-cvx::dynpar(c[C_THETA_MAX]), v_X.row(4), cvx::dynpar(c[C_THETA_MAX])));
        socp.addConstraint(cvx::box(-cvx::dynpar(c[C_W_B_MAX]), v_X.row(5), cvx::dynpar(c[C_W_B_MAX])));
        // gimbal and thrust ranges
        socp.addConstraint(cvx::box(-cvx::dynpar(c[C_GIMBAL_MAX]), v_U.row(0), cvx::dynpar(c[C_GIMBAL_MAX])));
        socp.addConstraint(cvx::box(cvx::dynpar(c[C_T_MIN]), v_U.row(1), cvx::dynpar(c[C_T_MAX])));
    }
#endif

#if !defined(SCPP_PLUGIN_GENERATE)
    SCPP_PLUGIN_MEMBERS(ROCKET2D_PLUGIN)      // NLP, NCONE, NCR, MAXDIM, cone_dim/off, crow/row, setup, initial_guess, fixed: from the generated table
#endif
};

} // namespace scpp
