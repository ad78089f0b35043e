/*
 * oracle/orc_models.c — CPU ORACLE (test infrastructure): model plugins restated in plain C.
 * PARITY UNPINNED (see orc.h).  Reference: scpp_models/src/rocketQuat.cpp, rocket2d.cpp,
 * scpp_models/include/common.hpp.
 */
#include "orc.h"
#include <math.h>
#include <string.h>

int orc_model_dims(int model, int *nx, int *nu, int *np)
{
    if (model == ORC_MODEL_ROCKETQUAT) { *nx = 14; *nu = 4; *np = 10; return 0; }   /* rocketQuatDefinitions.hpp:6-11 */
    if (model == ORC_MODEL_ROCKET2D)   { *nx = 6;  *nu = 2; *np = 6;  return 0; }   /* rocket2dDefinitions.hpp */
    return -1;
}

/* ---- RocketQuat::systemFlowMap, rocketQuat.cpp:7-37.
 * par = [alpha_m, g_I(3), J_B(3), r_T_B(3)] (packed at :170).
 * R(q) is Eigen's Quaternion(w,x,y,z).toRotationMatrix() WITHOUT normalisation (:29-30).
 * The last term is w.cross(w) == 0 (:36) — kept literally. */
static void rq_f(const double *x, const double *u, const double *par, double *f)
{
    const double alpha_m = par[0];
    const double *g = par + 1, *J = par + 4, *rT = par + 7;
    const double m = x[0];
    const double qw = x[7], qx = x[8], qy = x[9], qz = x[10];
    const double wx = x[11], wy = x[12], wz = x[13];
    const double Tx = u[0], Ty = u[1], Tz = u[2], tau = u[3];

    /* Eigen toRotationMatrix: tx=2x, ty=2y, tz=2z, twx=tx*w ... */
    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    const double R00 = 1 - (tyy + tzz), R01 = txy - twz, R02 = txz + twy;
    const double R10 = txy + twz, R11 = 1 - (txx + tzz), R12 = tyz - twx;
    const double R20 = txz - twy, R21 = tyz + twx, R22 = 1 - (txx + tyy);

    f[0] = -alpha_m * sqrt(Tx * Tx + Ty * Ty + Tz * Tz);
    f[1] = x[4]; f[2] = x[5]; f[3] = x[6];
    f[4] = 1. / m * (R00 * Tx + R01 * Ty + R02 * Tz) + g[0];
    f[5] = 1. / m * (R10 * Tx + R11 * Ty + R12 * Tz) + g[1];
    f[6] = 1. / m * (R20 * Tx + R21 * Ty + R22 * Tz) + g[2];
    /* 0.5 * omegaMatrix(w) * q, common.hpp:124-134 */
    f[7]  = 0.5 * (-wx * qx - wy * qy - wz * qz);
    f[8]  = 0.5 * ( wx * qw + wz * qy - wy * qz);
    f[9]  = 0.5 * ( wy * qw - wz * qx + wx * qz);
    f[10] = 0.5 * ( wz * qw + wy * qx - wx * qy);
    /* J^-1 (r_T x T + torque) - w x w */
    const double cx = rT[1] * Tz - rT[2] * Ty;
    const double cy = rT[2] * Tx - rT[0] * Tz;
    const double cz = rT[0] * Ty - rT[1] * Tx;
    f[11] = (cx + 0.) / J[0] - (wy * wz - wz * wy);
    f[12] = (cy + 0.) / J[1] - (wz * wx - wx * wz);
    f[13] = (cz + tau) / J[2] - (wx * wy - wy * wx);
}

/* exact derivatives of rq_f (what CppAD's Jacobian returns, systemDynamics.hpp:206-235);
 * hand-derived, checked in tests against central differences and dual numbers */
static void rq_jac(const double *x, const double *u, const double *par, double *A, double *B)
{
    enum { NX = 14 };
    const double alpha_m = par[0];
    const double *J = par + 4, *rT = par + 7;
    const double m = x[0];
    const double w = x[7], qx = x[8], qy = x[9], qz = x[10];
    const double wx = x[11], wy = x[12], wz = x[13];
    const double Tx = u[0], Ty = u[1], Tz = u[2];
    memset(A, 0, sizeof(double) * NX * NX);
    memset(B, 0, sizeof(double) * NX * 4);
#define A_(r, c) A[(r) + NX * (c)]
#define B_(r, c) B[(r) + NX * (c)]
    const double R00 = 1 - 2 * (qy * qy + qz * qz), R01 = 2 * (qx * qy - w * qz), R02 = 2 * (qx * qz + w * qy);
    const double R10 = 2 * (qx * qy + w * qz), R11 = 1 - 2 * (qx * qx + qz * qz), R12 = 2 * (qy * qz - w * qx);
    const double R20 = 2 * (qx * qz - w * qy), R21 = 2 * (qy * qz + w * qx), R22 = 1 - 2 * (qx * qx + qy * qy);
    const double RT0 = R00 * Tx + R01 * Ty + R02 * Tz;
    const double RT1 = R10 * Tx + R11 * Ty + R12 * Tz;
    const double RT2 = R20 * Tx + R21 * Ty + R22 * Tz;
    /* r_dot = v */
    A_(1, 4) = 1; A_(2, 5) = 1; A_(3, 6) = 1;
    /* v_dot = R T / m + g */
    A_(4, 0) = -RT0 / (m * m); A_(5, 0) = -RT1 / (m * m); A_(6, 0) = -RT2 / (m * m);
    const double im = 1. / m;
    A_(4, 7)  = im * 2 * (-qz * Ty + qy * Tz);
    A_(5, 7)  = im * 2 * ( qz * Tx - qx * Tz);
    A_(6, 7)  = im * 2 * (-qy * Tx + qx * Ty);
    A_(4, 8)  = im * 2 * ( qy * Ty + qz * Tz);
    A_(5, 8)  = im * 2 * ( qy * Tx - 2 * qx * Ty - w * Tz);
    A_(6, 8)  = im * 2 * ( qz * Tx + w * Ty - 2 * qx * Tz);
    A_(4, 9)  = im * 2 * (-2 * qy * Tx + qx * Ty + w * Tz);
    A_(5, 9)  = im * 2 * ( qx * Tx + qz * Tz);
    A_(6, 9)  = im * 2 * (-w * Tx + qz * Ty - 2 * qy * Tz);
    A_(4, 10) = im * 2 * (-2 * qz * Tx - w * Ty + qx * Tz);
    A_(5, 10) = im * 2 * ( w * Tx - 2 * qz * Ty + qy * Tz);
    A_(6, 10) = im * 2 * ( qx * Tx + qy * Ty);
    /* q_dot = 0.5 Omega(w) q */
    A_(7, 8) = -0.5 * wx;  A_(7, 9) = -0.5 * wy;  A_(7, 10) = -0.5 * wz;
    A_(8, 7) =  0.5 * wx;  A_(8, 9) =  0.5 * wz;  A_(8, 10) = -0.5 * wy;
    A_(9, 7) =  0.5 * wy;  A_(9, 8) = -0.5 * wz;  A_(9, 10) =  0.5 * wx;
    A_(10, 7) = 0.5 * wz;  A_(10, 8) = 0.5 * wy;  A_(10, 9) = -0.5 * wx;
    A_(7, 11) = -0.5 * qx; A_(7, 12) = -0.5 * qy; A_(7, 13) = -0.5 * qz;
    A_(8, 11) =  0.5 * w;  A_(8, 12) = -0.5 * qz; A_(8, 13) =  0.5 * qy;
    A_(9, 11) =  0.5 * qz; A_(9, 12) =  0.5 * w;  A_(9, 13) = -0.5 * qx;
    A_(10, 11) = -0.5 * qy; A_(10, 12) = 0.5 * qx; A_(10, 13) = 0.5 * w;
    /* w_dot: d(w x w)/dw == 0 exactly */
    /* B */
    const double nT = sqrt(Tx * Tx + Ty * Ty + Tz * Tz);
    B_(0, 0) = -alpha_m * Tx / nT; B_(0, 1) = -alpha_m * Ty / nT; B_(0, 2) = -alpha_m * Tz / nT;
    B_(4, 0) = im * R00; B_(4, 1) = im * R01; B_(4, 2) = im * R02;
    B_(5, 0) = im * R10; B_(5, 1) = im * R11; B_(5, 2) = im * R12;
    B_(6, 0) = im * R20; B_(6, 1) = im * R21; B_(6, 2) = im * R22;
    /* r x T = (ry Tz - rz Ty, rz Tx - rx Tz, rx Ty - ry Tx) */
    B_(11, 1) = -rT[2] / J[0]; B_(11, 2) =  rT[1] / J[0];
    B_(12, 0) =  rT[2] / J[1]; B_(12, 2) = -rT[0] / J[1];
    B_(13, 0) = -rT[1] / J[2]; B_(13, 1) =  rT[0] / J[2];
    B_(13, 3) = 1. / J[2];
#undef A_
#undef B_
}

/* ---- Rocket2d::systemFlowMap, rocket2d.cpp:7-40. par = [m, J_B, g_I(2), r_T_B(2)] (:145).
 * T_B = Rotation2D(angle) * (0, magnitude) = (-sin(a) mag, cos(a) mag). */
static void r2d_f(const double *x, const double *u, const double *par, double *f)
{
    const double m = par[0], JB = par[1];
    const double *g = par + 2, *rT = par + 4;
    const double eta = x[4], w = x[5];
    const double ang = u[0], mag = u[1];
    const double TBx = cos(ang) * 0. - sin(ang) * mag;
    const double TBy = sin(ang) * 0. + cos(ang) * mag;
    const double ce = cos(eta), se = sin(eta);
    f[0] = x[2]; f[1] = x[3];
    f[2] = 1. / m * (ce * TBx - se * TBy) + g[0];
    f[3] = 1. / m * (se * TBx + ce * TBy) + g[1];
    f[4] = w;
    f[5] = 1. / JB * (rT[0] * TBy - rT[1] * TBx);
}

static void r2d_jac(const double *x, const double *u, const double *par, double *A, double *B)
{
    enum { NX = 6 };
    const double m = par[0], JB = par[1];
    const double *rT = par + 4;
    const double eta = x[4];
    const double ang = u[0], mag = u[1];
    const double sa = sin(ang), ca = cos(ang), ce = cos(eta), se = sin(eta);
    const double TBx = -sa * mag, TBy = ca * mag;
    /* dTB/dang = (-ca mag, -sa mag), dTB/dmag = (-sa, ca) */
    const double dTBx_a = -ca * mag, dTBy_a = -sa * mag, dTBx_m = -sa, dTBy_m = ca;
    memset(A, 0, sizeof(double) * NX * NX);
    memset(B, 0, sizeof(double) * NX * 2);
#define A_(r, c) A[(r) + NX * (c)]
#define B_(r, c) B[(r) + NX * (c)]
    A_(0, 2) = 1; A_(1, 3) = 1;
    A_(2, 4) = 1. / m * (-se * TBx - ce * TBy);
    A_(3, 4) = 1. / m * ( ce * TBx - se * TBy);
    A_(4, 5) = 1;
    B_(2, 0) = 1. / m * (ce * dTBx_a - se * dTBy_a);
    B_(3, 0) = 1. / m * (se * dTBx_a + ce * dTBy_a);
    B_(2, 1) = 1. / m * (ce * dTBx_m - se * dTBy_m);
    B_(3, 1) = 1. / m * (se * dTBx_m + ce * dTBy_m);
    B_(5, 0) = 1. / JB * (rT[0] * dTBy_a - rT[1] * dTBx_a);
    B_(5, 1) = 1. / JB * (rT[0] * dTBy_m - rT[1] * dTBx_m);
#undef A_
#undef B_
}

void orc_f(int model, const double *x, const double *u, const double *par, double *f)
{
    if (model == ORC_MODEL_ROCKETQUAT) rq_f(x, u, par, f); else r2d_f(x, u, par, f);
}
void orc_jac(int model, const double *x, const double *u, const double *par, double *A, double *B)
{
    if (model == ORC_MODEL_ROCKETQUAT) rq_jac(x, u, par, A, B); else r2d_jac(x, u, par, A, B);
}

/* ---- eulerToQuaternionXYZ, common.hpp:29-38: AngleAxis(x,X)*AngleAxis(y,Y)*AngleAxis(z,Z) */
static void quat_mul(const double *a, const double *b, double *o) /* (w,x,y,z) Hamilton */
{
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}
void orc_euler_to_quat_xyz(const double *rpy, double *q)
{
    double qx[4] = {cos(rpy[0] / 2), sin(rpy[0] / 2), 0, 0};
    double qy[4] = {cos(rpy[1] / 2), 0, sin(rpy[1] / 2), 0};
    double qz[4] = {cos(rpy[2] / 2), 0, 0, sin(rpy[2] / 2)};
    double t[4];
    quat_mul(qx, qy, t);
    quat_mul(t, qz, q);
}

/* ---- Parameters::nondimensionalize, rocketQuat.cpp:291-312 (time and angles unscaled) */
void orc_rq_nondimensionalize(orc_rq_params *p)
{
    p->m_scale = p->x_init[0];
    p->r_scale = sqrt(p->x_init[1] * p->x_init[1] + p->x_init[2] * p->x_init[2] + p->x_init[3] * p->x_init[3]);
    const double ms = p->m_scale, rs = p->r_scale;
    p->alpha_m *= rs;
    for (int i = 0; i < 3; i++) { p->r_T_B[i] /= rs; p->g_I[i] /= rs; p->J_B[i] /= ms * rs * rs; }
    p->x_init[0] /= ms;
    for (int i = 1; i < 7; i++) p->x_init[i] /= rs;
    p->x_final[0] /= ms;
    for (int i = 1; i < 7; i++) p->x_final[i] /= rs;
    p->T_min /= ms * rs; p->T_max /= ms * rs; p->t_max /= ms * rs * rs;
}
/* ---- Parameters::redimensionalize, rocketQuat.cpp:314-332 */
void orc_rq_redimensionalize(orc_rq_params *p)
{
    const double ms = p->m_scale, rs = p->r_scale;
    p->alpha_m /= rs;
    for (int i = 0; i < 3; i++) { p->r_T_B[i] *= rs; p->g_I[i] *= rs; p->J_B[i] *= ms * rs * rs; }
    p->x_init[0] *= ms;
    for (int i = 1; i < 7; i++) p->x_init[i] *= rs;
    p->x_final[0] *= ms;
    for (int i = 1; i < 7; i++) p->x_final[i] *= rs;
    p->T_min *= ms * rs; p->T_max *= ms * rs; p->t_max *= ms * rs * rs;
}

/* Eigen Quaternion::slerp semantics */
static void slerp(const double *q0, const double *q1, double t, double *o)
{
    const double one = 1.0 - 2.220446049250313e-16;
    double d = q0[0] * q1[0] + q0[1] * q1[1] + q0[2] * q1[2] + q0[3] * q1[3];
    double ad = fabs(d), s0, s1;
    if (ad >= one) { s0 = 1 - t; s1 = t; }
    else {
        double th = acos(ad), st = sin(th);
        s0 = sin((1 - t) * th) / st; s1 = sin(t * th) / st;
    }
    if (d < 0) s1 = -s1;
    for (int i = 0; i < 4; i++) o[i] = s0 * q0[i] + s1 * q1[i];
}

/* ---- RocketQuat::getInitializedTrajectory, rocketQuat.cpp:39-68 (alpha2 = k/K, NOT k/(K-1);
 * U = (0,0,(T_max - T_min)/2,0)) */
void orc_rq_initial_trajectory(const orc_rq_params *p, int K, double *X, double *U, double *t)
{
    for (int k = 0; k < K; k++) {
        const double a1 = (double)(K - k) / K, a2 = (double)k / K;
        double *x = X + 14 * k;
        x[0] = a1 * p->x_init[0] + a2 * p->x_final[0];
        for (int i = 1; i < 7; i++) x[i] = a1 * p->x_init[i] + a2 * p->x_final[i];
        slerp(p->x_init + 7, p->x_final + 7, a2, x + 7);
        for (int i = 11; i < 14; i++) x[i] = a1 * p->x_init[i] + a2 * p->x_final[i];
        double *u = U + 4 * k;
        u[0] = 0; u[1] = 0; u[2] = (p->T_max - p->T_min) / 2.; u[3] = 0;
    }
    *t = p->final_time;
}
/* ---- getNewModelParameters, rocketQuat.cpp:168-173 */
void orc_rq_model_par(const orc_rq_params *p, double *par)
{
    par[0] = p->alpha_m;
    for (int i = 0; i < 3; i++) { par[1 + i] = p->g_I[i]; par[4 + i] = p->J_B[i]; par[7 + i] = p->r_T_B[i]; }
}

/* ---- Rocket2d::Parameters::nondimensionalize, rocket2d.cpp:198-214 */
void orc_r2d_nondimensionalize(orc_r2d_params *p)
{
    p->r_scale = sqrt(p->x_init[0] * p->x_init[0] + p->x_init[1] * p->x_init[1]);
    p->m_scale = p->m;
    const double ms = p->m_scale, rs = p->r_scale;
    p->m /= ms;
    for (int i = 0; i < 2; i++) { p->r_T_B[i] /= rs; p->g_I[i] /= rs; }
    p->J_B /= ms * rs * rs;
    for (int i = 0; i < 4; i++) { p->x_init[i] /= rs; p->x_final[i] /= rs; }
    p->T_min /= ms * rs; p->T_max /= ms * rs;
}
/* ---- Rocket2d::getInitializedTrajectory, rocket2d.cpp:121-136 ((T_max + T_min)/2) */
void orc_r2d_initial_trajectory(const orc_r2d_params *p, int K, double *X, double *U, double *t)
{
    for (int k = 0; k < K; k++) {
        const double a1 = (double)(K - k) / K, a2 = (double)k / K;
        for (int i = 0; i < 6; i++) X[6 * k + i] = a1 * p->x_init[i] + a2 * p->x_final[i];
        U[2 * k] = 0; U[2 * k + 1] = (p->T_max + p->T_min) / 2;
    }
    *t = p->final_time;
}
void orc_r2d_model_par(const orc_r2d_params *p, double *par)
{
    par[0] = p->m; par[1] = p->J_B; par[2] = p->g_I[0]; par[3] = p->g_I[1]; par[4] = p->r_T_B[0]; par[5] = p->r_T_B[1];
}

/* ---- counter-based uniform in [-1,1): splitmix64 of (seed, instance, draw) */
static unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
double orc_uniform_pm1(unsigned long long seed, unsigned long long instance, unsigned draw)
{
    unsigned long long h = splitmix64(seed ^ splitmix64(instance * 0x100000001B3ULL + draw));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

/* ---- Parameters::randomizeInitialState, rocketQuat.cpp:203-227 (commented out in the reference):
 * r_x,r_y *= U(-1,1); v_x,v_y *= U(-1,1); v_z *= 1+0.2U; roll,pitch = U*rpy_init; mass, r_z, yaw unchanged.
 * NOTE the reference writes only 3 of the 4 quaternion entries (x_init.segment(7,3)); here the full
 * quaternion is written (a 3-entry write of (w,x,y) would leave a stale z and a non-unit quaternion). */
void orc_rq_perturb(const orc_rq_params *nom, const double *rpy_init, unsigned long long seed,
                    unsigned long long inst, orc_rq_params *out)
{
    *out = *nom;
    out->x_init[1] *= orc_uniform_pm1(seed, inst, 0);
    out->x_init[2] *= orc_uniform_pm1(seed, inst, 1);
    out->x_init[4] *= orc_uniform_pm1(seed, inst, 2);
    out->x_init[5] *= orc_uniform_pm1(seed, inst, 3);
    out->x_init[6] *= 1. + 0.2 * orc_uniform_pm1(seed, inst, 4);
    double e[3] = {orc_uniform_pm1(seed, inst, 5) * rpy_init[0], orc_uniform_pm1(seed, inst, 6) * rpy_init[1], rpy_init[2]};
    orc_euler_to_quat_xyz(e, out->x_init + 7);
}
