/*
 * oracle/orc_rkf78.c — CPU ORACLE (test infrastructure): multiple-shooting discretisation and
 * forward simulation, restated literally.  PARITY UNPINNED (see orc.h).
 *
 * Reference: scpp_core/include/discretizationImplementation.hpp:38-181, scpp_core/src/simulation.cpp:10-42.
 * Third-party piece restated: Boost.odeint runge_kutta_fehlberg78 driven by integrate_adaptive with a
 * plain stepper => integrate_const: n equal steps, t_n = t0 + n*h, do_step with the 8th-order weights
 * (Boost is absent from this image and not pinned by the reference).
 */
#include "orc.h"
#include <math.h>
#include <string.h>

/* Fehlberg 7(8) tableau, 13 stages */
static const double RK_C[13] = {0, 2. / 27, 1. / 9, 1. / 6, 5. / 12, 1. / 2, 5. / 6, 1. / 6, 2. / 3, 1. / 3, 1, 0, 1};
static const double RK_A[13][12] = {
    {0},
    {2. / 27},
    {1. / 36, 1. / 12},
    {1. / 24, 0, 1. / 8},
    {5. / 12, 0, -25. / 16, 25. / 16},
    {1. / 20, 0, 0, 1. / 4, 1. / 5},
    {-25. / 108, 0, 0, 125. / 108, -65. / 27, 125. / 54},
    {31. / 300, 0, 0, 0, 61. / 225, -2. / 9, 13. / 900},
    {2, 0, 0, -53. / 6, 704. / 45, -107. / 9, 67. / 90, 3},
    {-91. / 108, 0, 0, 23. / 108, -976. / 135, 311. / 54, -19. / 60, 17. / 6, -1. / 12},
    {2383. / 4100, 0, 0, -341. / 164, 4496. / 1025, -301. / 82, 2133. / 4100, 45. / 82, 45. / 164, 18. / 41},
    {3. / 205, 0, 0, 0, 0, -6. / 41, -3. / 205, -3. / 41, 3. / 41, 6. / 41, 0},
    {-1777. / 4100, 0, 0, -341. / 164, 4496. / 1025, -289. / 82, 2193. / 4100, 51. / 82, 33. / 164, 12. / 41, 0, 1}};
static const double RK_B[13] = {0, 0, 0, 0, 0, 34. / 105, 9. / 35, 9. / 35, 9. / 280, 9. / 280, 0, 41. / 840, 41. / 840};

void orc_rkf78_tableau(double *c, double *a, double *b)
{
    for (int i = 0; i < 13; i++) { c[i] = RK_C[i]; b[i] = RK_B[i]; for (int j = 0; j < 13; j++) a[13 * i + j] = (j < 12) ? RK_A[i][j] : 0.; }
}

typedef void (*rhs_fn)(const double *V, double *dV, double t, void *ctx);

#define MAXV (ORC_MAX_NX * (1 + ORC_MAX_NX + 2 * ORC_MAX_NU + 2))

/* integrate_const semantics: nsteps equal steps of h, time of step n = t0 + n*h */
static void rkf78_integrate(rhs_fn f, void *ctx, double *V, int n, double t0, double h, int nsteps)
{
    static _Thread_local double k[13][MAXV];
    double tmp[MAXV];
    for (int st = 0; st < nsteps; st++) {
        const double t = t0 + st * h;
        for (int i = 0; i < 13; i++) {
            for (int e = 0; e < n; e++) {
                double acc = 0;
                for (int j = 0; j < i; j++) if (RK_A[i][j] != 0.) acc += RK_A[i][j] * k[j][e];
                tmp[e] = V[e] + h * acc;
            }
            f(tmp, k[i], t + RK_C[i] * h, ctx);
        }
        for (int e = 0; e < n; e++) {
            double acc = 0;
            for (int i = 0; i < 13; i++) if (RK_B[i] != 0.) acc += RK_B[i] * k[i][e];
            V[e] += h * acc;
        }
    }
}

/* dense inverse by LU with partial pivoting (Eigen's .inverse() for fixed sizes > 4 goes through
 * PartialPivLU; discretizationImplementation.hpp:65) */
static void mat_inverse(int n, const double *M /* col-major */, double *Minv)
{
    double a[ORC_MAX_NX * ORC_MAX_NX], b[ORC_MAX_NX * ORC_MAX_NX];
    memcpy(a, M, sizeof(double) * n * n);
    for (int i = 0; i < n * n; i++) b[i] = 0;
    for (int i = 0; i < n; i++) b[i + n * i] = 1;
    for (int c = 0; c < n; c++) {
        int piv = c; double best = fabs(a[c + n * c]);
        for (int r = c + 1; r < n; r++) if (fabs(a[r + n * c]) > best) { best = fabs(a[r + n * c]); piv = r; }
        if (piv != c) for (int j = 0; j < n; j++) {
            double t = a[c + n * j]; a[c + n * j] = a[piv + n * j]; a[piv + n * j] = t;
            t = b[c + n * j]; b[c + n * j] = b[piv + n * j]; b[piv + n * j] = t;
        }
        const double d = 1. / a[c + n * c];
        for (int r = c + 1; r < n; r++) {
            const double l = a[r + n * c] * d;
            if (l == 0.) continue;
            for (int j = c; j < n; j++) a[r + n * j] -= l * a[c + n * j];
            for (int j = 0; j < n; j++) b[r + n * j] -= l * b[c + n * j];
        }
    }
    for (int j = 0; j < n; j++)
        for (int r = n - 1; r >= 0; r--) {
            double acc = b[r + n * j];
            for (int c = r + 1; c < n; c++) acc -= a[r + n * c] * Minv[c + n * j];
            Minv[r + n * j] = acc / a[r + n * r];
        }
}

typedef struct {
    int model, nx, nu, foh, vartime;
    const double *u0, *u1, *par;
    double time; /* td.t (sigma) */
    double dt;
} ms_ctx;

/* ODE<INTERPOLATE_INPUT, VARIABLE_TIME>::operator(), discretizationImplementation.hpp:38-120.
 * V is nx x ncols column-major: [x | Phi | Bbar | (Cbar) | (sbar) | zbar]. */
static void ms_rhs(const double *V, double *dV, double t, void *vctx)
{
    const ms_ctx *c = (const ms_ctx *)vctx;
    const int nx = c->nx, nu = c->nu;
    const double *x = V;
    double u[ORC_MAX_NU], f[ORC_MAX_NX], A[ORC_MAX_NX * ORC_MAX_NX], B[ORC_MAX_NX * ORC_MAX_NU];
    for (int i = 0; i < nu; i++) u[i] = c->foh ? c->u0[i] + t / c->dt * (c->u1[i] - c->u0[i]) : c->u0[i]; /* :45 */
    orc_f(c->model, x, u, c->par, f);
    orc_jac(c->model, x, u, c->par, A, B);
    if (c->vartime) { /* :58-62 */
        for (int i = 0; i < nx * nx; i++) A[i] *= c->time;
        for (int i = 0; i < nx * nu; i++) B[i] *= c->time;
    }
    const double *Phi = V + nx;
    double Pinv[ORC_MAX_NX * ORC_MAX_NX];
    mat_inverse(nx, Phi, Pinv); /* :65 */
    int col = 0;
    for (int i = 0; i < nx; i++) dV[i] = c->vartime ? c->time * f[i] : f[i]; /* :70-77 */
    col += 1;
    /* A * Phi  :81 */
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nx; i++) {
            double acc = 0;
            for (int l = 0; l < nx; l++) acc += A[i + nx * l] * Phi[l + nx * j];
            dV[nx * (col + j) + i] = acc;
        }
    col += nx;
    /* Phi^-1 * B */
    double PB[ORC_MAX_NX * ORC_MAX_NU];
    for (int j = 0; j < nu; j++)
        for (int i = 0; i < nx; i++) {
            double acc = 0;
            for (int l = 0; l < nx; l++) acc += Pinv[i + nx * l] * B[l + nx * j];
            PB[i + nx * j] = acc;
        }
    if (c->foh) { /* :84-96 */
        const double alpha = (c->dt - t) / c->dt, beta = t / c->dt;
        for (int e = 0; e < nx * nu; e++) dV[nx * col + e] = PB[e] * alpha;
        col += nu;
        for (int e = 0; e < nx * nu; e++) dV[nx * col + e] = PB[e] * beta;
        col += nu;
    } else { /* :99-101 */
        for (int e = 0; e < nx * nu; e++) dV[nx * col + e] = PB[e];
        col += nu;
    }
    /* -A x - B u */
    double r[ORC_MAX_NX];
    for (int i = 0; i < nx; i++) {
        double acc = 0;
        for (int l = 0; l < nx; l++) acc -= A[i + nx * l] * x[l];
        for (int l = 0; l < nu; l++) acc -= B[i + nx * l] * u[l];
        r[i] = acc;
    }
    if (c->vartime) { /* :104-111 */
        for (int i = 0; i < nx; i++) {
            double acc = 0;
            for (int l = 0; l < nx; l++) acc += Pinv[i + nx * l] * f[l];
            dV[nx * col + i] = acc;
        }
        col += 1;
        for (int i = 0; i < nx; i++) {
            double acc = 0;
            for (int l = 0; l < nx; l++) acc += Pinv[i + nx * l] * r[l];
            dV[nx * col + i] = acc;
        }
        col += 1;
    } else { /* :113-117 */
        for (int i = 0; i < nx; i++) {
            double acc = 0;
            for (int l = 0; l < nx; l++) acc += Pinv[i + nx * l] * (f[l] + r[l]);
            dV[nx * col + i] = acc;
        }
        col += 1;
    }
}

/* multipleShootingImplementation, discretizationImplementation.hpp:122-181 */
void orc_discretize(int model, int K, const double *X, const double *U, double t_or_sigma,
                    const double *par, int foh, int free_time,
                    double *A, double *B, double *C, double *s, double *z)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    double dt = 1. / (double)(K - 1); /* :133 */
    if (!free_time) dt *= t_or_sigma; /* :135-138 */
    const int ncols = 1 + nx + nu + (foh ? nu : 0) + (free_time ? 1 : 0) + 1;
    for (int k = 0; k < K - 1; k++) {
        double V[MAXV];
        memset(V, 0, sizeof(V));
        for (int i = 0; i < nx; i++) V[i] = X[nx * k + i];          /* :145 */
        for (int i = 0; i < nx; i++) V[nx * (1 + i) + i] = 1.;      /* :146 */
        ms_ctx c = {model, nx, nu, foh, free_time, U + nu * k, foh ? U + nu * (k + 1) : U + nu * k, par, t_or_sigma, dt};
        rkf78_integrate(ms_rhs, &c, V, nx * ncols, 0., dt / 5., 5); /* :154 */
        const double *Phi = V + nx;
        double *Ak = A + nx * nx * k;
        memcpy(Ak, Phi, sizeof(double) * nx * nx);                  /* :158 */
        int col = 1 + nx;
#define MUL(dst, ncol)                                                          \
    for (int j = 0; j < (ncol); j++)                                            \
        for (int i = 0; i < nx; i++) {                                          \
            double acc = 0;                                                     \
            for (int l = 0; l < nx; l++) acc += Ak[i + nx * l] * V[nx * (col + j) + l]; \
            (dst)[i + nx * j] = acc;                                            \
        }                                                                       \
    col += (ncol);
        MUL(B + nx * nu * k, nu)                                    /* :161 */
        if (foh) { MUL(C + nx * nu * k, nu) }                       /* :164-168 */
        if (free_time) { MUL(s + nx * k, 1) }                       /* :170-174 */
        MUL(z + nx * k, 1)                                          /* :176 */
#undef MUL
    }
}

typedef struct { int model, nu; const double *u0, *u1, *par; double dt; } sim_ctx;
static void sim_rhs(const double *x, double *dx, double t, void *vctx) /* simulation.cpp:25-29 */
{
    const sim_ctx *c = (const sim_ctx *)vctx;
    double u[ORC_MAX_NU];
    for (int i = 0; i < c->nu; i++) u[i] = c->u0[i] + t / c->dt * (c->u1[i] - c->u0[i]);
    orc_f(c->model, x, u, c->par, dx);
}
/* simulate(), simulation.cpp:31-42: integrate_adaptive(stepper, ode, x, 0, dt, dt/20) */
void orc_simulate(int model, double dt, const double *u0, const double *u1, const double *par, double *x)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    sim_ctx c = {model, nu, u0, u1, par, dt};
    rkf78_integrate(sim_rhs, &c, x, nx, 0., dt / 20., 20);
}
