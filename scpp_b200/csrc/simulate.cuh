// scpp_b200/csrc/simulate.cuh — kernel K4 body: forward simulation of the nonlinear model over one control step.
//
// Replaces scpp::simulate (scpp_core/src/simulation.cpp:10-42): Boost.odeint runge_kutta_fehlberg78 driven by
// integrate_adaptive with a plain stepper, i.e. n equal steps with the 8th-order weights (the
// reference uses dt/20 => 20 steps), input interpolated linearly between u0 and u1 (:25-29).  One THREAD per instance:
// the state is 14 doubles and a control step is 260 evaluations of the flow map, so the whole closed-loop step of a batch
// is one small launch; the 13 stage derivatives live in local memory.
#pragma once
#include "models.cuh"

namespace scpp {

// Fehlberg 7(8): nodes c, 8th-order weights b, coupling rows a (13 stages).  Real constant data: host copy + __constant__ copy (a
// function-local table would be rebuilt on the stack at every call; measured: 28 ms instead of < 1 ms for a 1024-instance step)
#define SCPP_RKF78_C {0., 2. / 27, 1. / 9, 1. / 6, 5. / 12, 1. / 2, 5. / 6, 1. / 6, 2. / 3, 1. / 3, 1., 0., 1.}
#define SCPP_RKF78_B {0., 0., 0., 0., 0., 34. / 105, 9. / 35, 9. / 35, 9. / 280, 9. / 280, 0., 41. / 840, 41. / 840}
#define SCPP_RKF78_A {                                                                                                       \
        {0.},                                                                                                                \
        {2. / 27},                                                                                                           \
        {1. / 36, 1. / 12},                                                                                                  \
        {1. / 24, 0., 1. / 8},                                                                                               \
        {5. / 12, 0., -25. / 16, 25. / 16},                                                                                  \
        {1. / 20, 0., 0., 1. / 4, 1. / 5},                                                                                   \
        {-25. / 108, 0., 0., 125. / 108, -65. / 27, 125. / 54},                                                              \
        {31. / 300, 0., 0., 0., 61. / 225, -2. / 9, 13. / 900},                                                              \
        {2., 0., 0., -53. / 6, 704. / 45, -107. / 9, 67. / 90, 3.},                                                          \
        {-91. / 108, 0., 0., 23. / 108, -976. / 135, 311. / 54, -19. / 60, 17. / 6, -1. / 12},                               \
        {2383. / 4100, 0., 0., -341. / 164, 4496. / 1025, -301. / 82, 2133. / 4100, 45. / 82, 45. / 164, 18. / 41},           \
        {3. / 205, 0., 0., 0., 0., -6. / 41, -3. / 205, -3. / 41, 3. / 41, 6. / 41, 0.},                                      \
        {-1777. / 4100, 0., 0., -341. / 164, 4496. / 1025, -289. / 82, 2193. / 4100, 51. / 82, 33. / 164, 12. / 41, 0., 1.}}
static const double rkf78_c_host[13] = SCPP_RKF78_C, rkf78_b_host[13] = SCPP_RKF78_B, rkf78_a_host[13][12] = SCPP_RKF78_A;
#if defined(__CUDACC__)
static __constant__ double rkf78_c_dev[13] = SCPP_RKF78_C, rkf78_b_dev[13] = SCPP_RKF78_B, rkf78_a_dev[13][12] = SCPP_RKF78_A;
#endif
SCPP_HD double rkf78_c(int i)
{
#if defined(__CUDA_ARCH__)
    return rkf78_c_dev[i];
#else
    return rkf78_c_host[i];
#endif
}
SCPP_HD double rkf78_b(int i)
{
#if defined(__CUDA_ARCH__)
    return rkf78_b_dev[i];
#else
    return rkf78_b_host[i];
#endif
}
SCPP_HD double rkf78_a(int i, int j)
{
#if defined(__CUDA_ARCH__)
    return rkf78_a_dev[i][j];
#else
    return rkf78_a_host[i][j];
#endif
}

// x <- x(dt) under  x' = f(x, u0 + t/dt (u1 - u0)),  nsteps equal RKF78 steps (time of step n = n*h, as integrate_const)
template <class M>
SCPP_HD void rkf78_simulate(double *x, const double *u0, const double *u1, const double *par, double dt, int nsteps)
{
    constexpr int NX = M::NX, NU = M::NU;
    const double h = dt / nsteps;
    double k[13][NX];
    for (int st = 0; st < nsteps; st++) {
        const double t = st * h;
        for (int i = 0; i < 13; i++) {
            double xs[NX], u[NU];
            for (int e = 0; e < NX; e++) {
                double acc = 0.;
                for (int j = 0; j < i; j++) { const double aij = rkf78_a(i, j); if (aij != 0.) acc += aij * k[j][e]; }
                xs[e] = x[e] + h * acc;
            }
            const double ts = t + rkf78_c(i) * h;
            for (int j = 0; j < NU; j++) u[j] = u0[j] + ts / dt * (u1[j] - u0[j]);
            M::template flow_map<double>(xs, u, par, k[i]);
        }
        for (int e = 0; e < NX; e++) {
            double acc = 0.;
            for (int i = 0; i < 13; i++) { const double bi = rkf78_b(i); if (bi != 0.) acc += bi * k[i][e]; }
            x[e] += h * acc;
        }
    }
}

} // namespace scpp
