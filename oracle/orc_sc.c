/*
 * oracle/orc_sc.c — CPU ORACLE (test infrastructure): the SC convex sub-problem in ECOS standard form
 * and the outer successive-convexification loop, restated literally.  PARITY UNPINNED (see orc.h).
 *
 * Reference: scpp_core/src/SCProblem.cpp:6-138 (buildSCProblem), scpp_models/src/rocketQuat.cpp:70-166
 * (addApplicationConstraints / updateProblemParameters), scpp_models/src/rocket2d.cpp:46-84,
 * scpp_core/src/SCAlgorithm.cpp:66-210 (iterate / solve / readSolution).
 * Epigraph (absent) only canonicalises these expressions into c,A,b,G,h; row order is irrelevant to the optimum.
 */
#include "orc.h"
#include "orc_ipm.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

typedef struct { int *i, *j; double *v; int nnz, cap; } coo_t;
static void coo_push(coo_t *c, int i, int j, double v)
{
    if (c->nnz == c->cap) {
        c->cap = c->cap ? 2 * c->cap : 4096;
        c->i = (int *)realloc(c->i, sizeof(int) * c->cap);
        c->j = (int *)realloc(c->j, sizeof(int) * c->cap);
        c->v = (double *)realloc(c->v, sizeof(double) * c->cap);
    }
    c->i[c->nnz] = i; c->j[c->nnz] = j; c->v[c->nnz] = v; c->nnz++;
}
typedef struct { double *v; int n, cap; } dvec_t;
static void dv_push(dvec_t *d, double v)
{
    if (d->n == d->cap) { d->cap = d->cap ? 2 * d->cap : 1024; d->v = (double *)realloc(d->v, sizeof(double) * d->cap); }
    d->v[d->n++] = v;
}

typedef struct {
    int model, nx, nu, K, free_time;
    int KU;                     /* input columns: K (first-order hold) or K - 1 (zero-order hold, trajectoryData.hpp:27-32) */
    int has_delta;              /* 1: SC (trust-region epigraph variable delta_k per stage), 0: SCvx */
    int n;                      /* variables */
    int stage_stride, stage_last_sz;
    int iN1, iSigma, iDsigma;
    coo_t A, GL, GQ;            /* equalities, LP rows, SOC rows */
    dvec_t b, hL, hQ, keq;
    int q[4096]; int ncones;
    double *c;
} socp_t;

/* variable layout: stage k = [x(nx) u(nu) delta(1) nu(nx) nu_bound(nx)], last stage without nu/nu_bound;
 * then norm1_nu, sigma, delta_sigma */
static int iX(const socp_t *P, int k, int i) { return k * P->stage_stride + i; }
static int iU(const socp_t *P, int k, int i) { return k * P->stage_stride + P->nx + i; }
static int iDelta(const socp_t *P, int k) { return k * P->stage_stride + P->nx + P->nu; }
static int iNu(const socp_t *P, int k, int i) { return k * P->stage_stride + P->nx + P->nu + P->has_delta + i; }
static int iNub(const socp_t *P, int k, int i) { return k * P->stage_stride + 2 * P->nx + P->nu + P->has_delta + i; }

/* equality  sum coef*x = rhs ; key places the multiplier in the KKT ordering */
static int eq_begin(socp_t *P, double rhs, double key) { dv_push(&P->b, rhs); dv_push(&P->keq, key); return P->b.n - 1; }
/* s = a'x + a0 >= 0 rows:  G = -a, h = a0 */
static int lp_begin(socp_t *P, double a0) { dv_push(&P->hL, a0); return P->hL.n - 1; }
static void lp_coef(socp_t *P, int row, int col, double a) { coo_push(&P->GL, row, col, -a); }
static int soc_begin(socp_t *P, int dim) { int r0 = P->hQ.n; P->q[P->ncones++] = dim; for (int i = 0; i < dim; i++) dv_push(&P->hQ, 0.); return r0; }
static void soc_const(socp_t *P, int row, double a0) { P->hQ.v[row] += a0; }
static void soc_coef(socp_t *P, int row, int col, double a) { coo_push(&P->GQ, row, col, -a); }

static void socp_free(socp_t *P)
{
    free(P->A.i); free(P->A.j); free(P->A.v); free(P->GL.i); free(P->GL.j); free(P->GL.v); free(P->GQ.i); free(P->GQ.j); free(P->GQ.v);
    free(P->b.v); free(P->hL.v); free(P->hQ.v); free(P->keq.v); free(P->c);
}

/* ---- buildSCProblem, SCProblem.cpp:6-138.  Zero-order hold (dd.interpolatedInput() false): no C_k term (:49-52), the trust region of the
 * last node has no input part (:116-121); the variable layout keeps a column U(:, K-1), which then appears in no row but the equalities that
 * pin it to zero in build_full (a variable the reference does not have == a variable fixed to a constant that enters nothing) ---- */
static void build_sc_problem(socp_t *P, const orc_sc_config *cfg, double weight_tr,
                             const double *Xbar, const double *Ubar, double sigmabar,
                             const double *A, const double *B, const double *C, const double *s, const double *z)
{
    const int nx = P->nx, nu = P->nu, K = P->K;
    P->has_delta = 1;
    P->stage_stride = 2 * nx + nu + 1 + nx;
    P->n = (K - 1) * P->stage_stride + (nx + nu + 1);
    P->iN1 = P->n++;
    if (P->free_time) { P->iSigma = P->n++; P->iDsigma = P->n++; } else { P->iSigma = P->iDsigma = -1; }
    P->c = (double *)calloc(P->n, sizeof(double));
    if (P->free_time) {
        P->c[P->iSigma] += cfg->weight_time;                     /* :32 */
        int r = lp_begin(P, -0.001); lp_coef(P, r, P->iSigma, 1.); /* sigma >= 0.001  :34 */
    }
    for (int k = 0; k < K - 1; k++) {                            /* :37-58 */
        const double *Ak = A + nx * nx * k, *Bk = B + nx * nu * k, *Ck = C + nx * nu * k, *zk = z + nx * k;
        for (int i = 0; i < nx; i++) {
            /* A x_k + B u_k + C u_k+1 + s sigma + z + nu - x_k+1 = 0 */
            int r = eq_begin(P, -zk[i], iNub(P, k, nx - 1) + 0.5);
            for (int j = 0; j < nx; j++) if (Ak[i + nx * j] != 0.) coo_push(&P->A, r, iX(P, k, j), Ak[i + nx * j]);
            for (int j = 0; j < nu; j++) if (Bk[i + nx * j] != 0.) coo_push(&P->A, r, iU(P, k, j), Bk[i + nx * j]);
            if (P->KU == K) for (int j = 0; j < nu; j++) if (Ck[i + nx * j] != 0.) coo_push(&P->A, r, iU(P, k + 1, j), Ck[i + nx * j]);
            if (P->free_time && s[nx * k + i] != 0.) coo_push(&P->A, r, P->iSigma, s[nx * k + i]);
            coo_push(&P->A, r, iNu(P, k, i), 1.);
            coo_push(&P->A, r, iX(P, k + 1, i), -1.);
        }
    }
    /* virtual control norm  :68-77 */
    for (int k = 0; k < K - 1; k++)
        for (int i = 0; i < nx; i++) {
            int r = lp_begin(P, 0.); lp_coef(P, r, iNub(P, k, i), 1.); lp_coef(P, r, iNu(P, k, i), 1.);   /* nu >= -nu_bound */
            r = lp_begin(P, 0.);     lp_coef(P, r, iNub(P, k, i), 1.); lp_coef(P, r, iNu(P, k, i), -1.);  /* nu <=  nu_bound */
        }
    {
        int r = lp_begin(P, 0.); /* norm1_nu - sum(nu_bound) >= 0  :73 */
        lp_coef(P, r, P->iN1, 1.);
        for (int k = 0; k < K - 1; k++) for (int i = 0; i < nx; i++) lp_coef(P, r, iNub(P, k, i), -1.);
        P->c[P->iN1] += cfg->weight_virtual_control;             /* :76 */
    }
    if (P->free_time) {                                          /* :79-100 */
        int r0 = soc_begin(P, 3);
        soc_const(P, r0, 0.5);     soc_coef(P, r0, P->iDsigma, 0.5);       /* 0.5 + 0.5 delta_sigma */
        soc_const(P, r0 + 1, 0.5); soc_coef(P, r0 + 1, P->iDsigma, -0.5);  /* 0.5 - 0.5 delta_sigma */
        soc_const(P, r0 + 2, -sigmabar); soc_coef(P, r0 + 2, P->iSigma, 1.); /* -t + sigma */
        P->c[P->iDsigma] += cfg->weight_trust_region_time;       /* :99 */
    }
    for (int k = 0; k < K; k++) {                                /* :102-126 */
        const int with_u = k < P->KU;                            /* :116 */
        int r0 = soc_begin(P, 1 + nx + (with_u ? nu : 0));
        soc_coef(P, r0, iDelta(P, k), 1.);
        for (int i = 0; i < nx; i++) { soc_const(P, r0 + 1 + i, Xbar[nx * k + i]); soc_coef(P, r0 + 1 + i, iX(P, k, i), -1.); }
        if (with_u) for (int i = 0; i < nu; i++) { soc_const(P, r0 + 1 + nx + i, Ubar[nu * k + i]); soc_coef(P, r0 + 1 + nx + i, iU(P, k, i), -1.); }
        P->c[iDelta(P, k)] += weight_tr;                         /* :134 */
    }
}

static void fix_var(socp_t *P, int col, double val) { int r = eq_begin(P, val, col + 0.25); coo_push(&P->A, r, col, 1.); }

/* ---- RocketQuat::addApplicationConstraints, rocketQuat.cpp:70-144 ---- */
static void add_rq_constraints(socp_t *P, const orc_rq_params *p, const double *thrust_dir)
{
    const int K = P->K, KU = P->KU;      /* KU = v_U.cols() */
    const double gimbal_const = tan(p->gimbal_max), gs_const = tan(p->gamma_gs);      /* :158-160 */
    const double tilt_const = sqrt((1. - cos(p->theta_max)) / 2.);
    for (int i = 0; i < 14; i++) fix_var(P, iX(P, 0, i), p->x_init[i]);              /* :79 */
    static const int fin[11] = {1, 2, 3, 4, 5, 6, 8, 9, 11, 12, 13};
    for (int q = 0; q < 11; q++) fix_var(P, iX(P, K - 1, fin[q]), p->x_final[fin[q]]); /* :83-89 */
    for (int k = 0; k < K; k++) { int r = lp_begin(P, -p->x_final[0]); lp_coef(P, r, iX(P, k, 0), 1.); } /* mass :93 */
    for (int k = 0; k < K; k++) {                                                     /* glide slope :96-97 */
        int r0 = soc_begin(P, 3);
        soc_coef(P, r0, iX(P, k, 3), gs_const); soc_coef(P, r0 + 1, iX(P, k, 1), 1.); soc_coef(P, r0 + 2, iX(P, k, 2), 1.);
    }
    for (int k = 0; k < K; k++) {                                                     /* tilt :100-101 */
        int r0 = soc_begin(P, 3);
        soc_const(P, r0, tilt_const); soc_coef(P, r0 + 1, iX(P, k, 8), 1.); soc_coef(P, r0 + 2, iX(P, k, 9), 1.);
    }
    for (int k = 0; k < K; k++) {                                                     /* angular rate :104-105 */
        int r0 = soc_begin(P, 4);
        soc_const(P, r0, p->w_B_max);
        for (int i = 0; i < 3; i++) soc_coef(P, r0 + 1 + i, iX(P, k, 11 + i), 1.);
    }
    fix_var(P, iU(P, KU - 1, 0), 0.); fix_var(P, iU(P, KU - 1, 1), 0.); fix_var(P, iU(P, KU - 1, 3), 0.); /* :109-111 */
    if (p->exact_minimum_thrust) {                                                    /* :113-121 */
        for (int k = 0; k < KU; k++) {
            int r = lp_begin(P, -p->T_min);
            for (int i = 0; i < 3; i++) { double d = thrust_dir ? thrust_dir[3 * k + i] : (i == 2 ? 1. : 0.); lp_coef(P, r, iU(P, k, i), d); }
        }
    } else {
        for (int k = 0; k < KU; k++) { int r = lp_begin(P, -p->T_min); lp_coef(P, r, iU(P, k, 2), 1.); } /* :125 */
    }
    for (int k = 0; k < KU; k++) {                                                     /* max thrust :129 */
        int r0 = soc_begin(P, 4);
        soc_const(P, r0, p->T_max);
        for (int i = 0; i < 3; i++) soc_coef(P, r0 + 1 + i, iU(P, k, i), 1.);
    }
    for (int k = 0; k < KU; k++) {                                                     /* gimbal :132-133 */
        int r0 = soc_begin(P, 3);
        soc_coef(P, r0, iU(P, k, 2), gimbal_const); soc_coef(P, r0 + 1, iU(P, k, 0), 1.); soc_coef(P, r0 + 2, iU(P, k, 1), 1.);
    }
    if (p->enable_roll_control) {                                                     /* :135-138 */
        for (int k = 0; k < KU; k++) {
            int r = lp_begin(P, p->t_max); lp_coef(P, r, iU(P, k, 3), 1.);
            r = lp_begin(P, p->t_max);     lp_coef(P, r, iU(P, k, 3), -1.);
        }
    } else {                                                                          /* :139-143 */
        for (int k = 0; k < K; k++) fix_var(P, iX(P, k, 13), 0.);
        for (int k = 0; k < KU; k++) fix_var(P, iU(P, k, 3), 0.);
    }
}

/* ---- Rocket2d::addApplicationConstraints, rocket2d.cpp:46-84 ---- */
static void add_r2d_constraints(socp_t *P, const orc_r2d_params *p)
{
    const int K = P->K, KU = P->KU;
    const double tan_gs = tan(p->gamma_gs);                                           /* :147 */
    if (p->constrain_initial_final) {                                                 /* :53-59 */
        for (int i = 0; i < 6; i++) fix_var(P, iX(P, 0, i), p->x_init[i]);
        for (int i = 0; i < 6; i++) fix_var(P, iX(P, K - 1, i), p->x_final[i]);
        fix_var(P, iU(P, KU - 1, 0), 0.);
    }
    for (int k = 0; k < K; k++) {                                                     /* glideslope :63-64 (norm of a 1-vector) */
        int r0 = soc_begin(P, 2);
        soc_coef(P, r0, iX(P, k, 1), tan_gs); soc_coef(P, r0 + 1, iX(P, k, 0), 1.);
    }
#define BOX(col, lo, hi) { int r = lp_begin(P, -(lo)); lp_coef(P, r, (col), 1.); r = lp_begin(P, (hi)); lp_coef(P, r, (col), -1.); }
    for (int k = 0; k < K; k++) BOX(iX(P, k, 4), -p->theta_max, p->theta_max)        /* :66-68 */
    for (int k = 0; k < K; k++) BOX(iX(P, k, 5), -p->w_B_max, p->w_B_max)            /* :70-72 */
    for (int k = 0; k < KU; k++) BOX(iU(P, k, 0), -p->gimbal_max, p->gimbal_max)      /* :76-78 */
    for (int k = 0; k < KU; k++) BOX(iU(P, k, 1), p->T_min, p->T_max)                 /* :80-82 */
#undef BOX
}

static void build_full(socp_t *P, int model, const void *params, const orc_sc_config *cfg, double weight_tr,
                       const double *Xbar, const double *Ubar, double sigmabar,
                       const double *A, const double *B, const double *C, const double *s, const double *z,
                       const double *thrust_dir)
{
    int np;
    memset(P, 0, sizeof(*P));
    P->model = model; P->K = cfg->K; P->free_time = cfg->free_final_time;
    orc_model_dims(model, &P->nx, &P->nu, &np);
    P->KU = cfg->interpolate_input ? P->K : P->K - 1;
    build_sc_problem(P, cfg, weight_tr, Xbar, Ubar, sigmabar, A, B, C, s, z);
    if (model == ORC_MODEL_ROCKETQUAT) add_rq_constraints(P, (const orc_rq_params *)params, thrust_dir);
    else add_r2d_constraints(P, (const orc_r2d_params *)params);
    for (int k = P->KU; k < P->K; k++) for (int j = 0; j < P->nu; j++) fix_var(P, iU(P, k, j), 0.);   /* the column the reference does not have */
}

/* merge LP + SOC rows into one G/h */
static void merged_G(const socp_t *P, int **Gi, int **Gj, double **Gv, double **h, int *nnz, int *m, int *l)
{
    *l = P->hL.n; *m = P->hL.n + P->hQ.n; *nnz = P->GL.nnz + P->GQ.nnz;
    *Gi = (int *)malloc(sizeof(int) * (*nnz)); *Gj = (int *)malloc(sizeof(int) * (*nnz)); *Gv = (double *)malloc(sizeof(double) * (*nnz));
    *h = (double *)malloc(sizeof(double) * (*m));
    memcpy(*Gi, P->GL.i, sizeof(int) * P->GL.nnz); memcpy(*Gj, P->GL.j, sizeof(int) * P->GL.nnz); memcpy(*Gv, P->GL.v, sizeof(double) * P->GL.nnz);
    for (int e = 0; e < P->GQ.nnz; e++) { (*Gi)[P->GL.nnz + e] = P->GQ.i[e] + *l; (*Gj)[P->GL.nnz + e] = P->GQ.j[e]; (*Gv)[P->GL.nnz + e] = P->GQ.v[e]; }
    memcpy(*h, P->hL.v, sizeof(double) * P->hL.n); memcpy(*h + *l, P->hQ.v, sizeof(double) * P->hQ.n);
}

int orc_sc_subproblem(int model, const void *params, const orc_sc_config *cfg, double weight_tr,
                      const double *Xbar, const double *Ubar, double sigmabar,
                      const double *A, const double *B, const double *C, const double *s, const double *z,
                      const double *thrust_dir,
                      double *X, double *U, double *sigma, double *nu, double *delta,
                      double *norm1_nu, double *delta_sigma, orc_ipm_info *info)
{
    socp_t P;
    build_full(&P, model, params, cfg, weight_tr, Xbar, Ubar, sigmabar, A, B, C, s, z, thrust_dir);
    int *Gi, *Gj, nnzG, m, l; double *Gv, *h;
    merged_G(&P, &Gi, &Gj, &Gv, &h, &nnzG, &m, &l);
    double *x = (double *)calloc(P.n, sizeof(double)), *y = (double *)calloc(P.b.n, sizeof(double));
    double *sv = (double *)calloc(m, sizeof(double)), *zv = (double *)calloc(m, sizeof(double));
    double *kv = (double *)malloc(sizeof(double) * P.n);
    for (int j = 0; j < P.n; j++) kv[j] = j;
    int st = orc_conic_solve_keys(P.n, P.b.n, m, l, P.ncones, P.q, P.c, P.b.v, h, P.A.nnz, P.A.i, P.A.j, P.A.v,
                                  nnzG, Gi, Gj, Gv, kv, P.keq.v, x, y, sv, zv, info);
    const int nx = P.nx, nuu = P.nu, K = P.K;
    for (int k = 0; k < K; k++) {
        if (X) for (int i = 0; i < nx; i++) X[nx * k + i] = x[iX(&P, k, i)];
        if (U) for (int i = 0; i < nuu; i++) U[nuu * k + i] = x[iU(&P, k, i)];
        if (delta) delta[k] = x[iDelta(&P, k)];
        if (nu && k < K - 1) for (int i = 0; i < nx; i++) nu[nx * k + i] = x[iNu(&P, k, i)];
    }
    if (sigma) *sigma = P.free_time ? x[P.iSigma] : sigmabar;
    if (norm1_nu) *norm1_nu = x[P.iN1];
    if (delta_sigma) *delta_sigma = P.free_time ? x[P.iDsigma] : 0.;
    free(x); free(y); free(sv); free(zv); free(kv); free(Gi); free(Gj); free(Gv); free(h);
    socp_free(&P);
    return st;
}

int orc_sc_export(int model, const void *params, const orc_sc_config *cfg, double weight_tr,
                  const double *Xbar, const double *Ubar, double sigmabar,
                  const double *A, const double *B, const double *C, const double *s, const double *z,
                  const double *thrust_dir,
                  orc_socp_dims *dims, double *c, double *b, double *h, int *q,
                  int *Ai, int *Aj, double *Av, int *Gi, int *Gj, double *Gv,
                  int *idx_X, int *idx_U, int *idx_sigma)
{
    socp_t P;
    build_full(&P, model, params, cfg, weight_tr, Xbar, Ubar, sigmabar, A, B, C, s, z, thrust_dir);
    int *gi, *gj, nnzG, m, l; double *gv, *hh;
    merged_G(&P, &gi, &gj, &gv, &hh, &nnzG, &m, &l);
    dims->n = P.n; dims->p = P.b.n; dims->m = m; dims->l = l; dims->ncones = P.ncones; dims->nnzA = P.A.nnz; dims->nnzG = nnzG;
    if (c) memcpy(c, P.c, sizeof(double) * P.n);
    if (b) memcpy(b, P.b.v, sizeof(double) * P.b.n);
    if (h) memcpy(h, hh, sizeof(double) * m);
    if (q) memcpy(q, P.q, sizeof(int) * P.ncones);
    if (Ai) { memcpy(Ai, P.A.i, sizeof(int) * P.A.nnz); memcpy(Aj, P.A.j, sizeof(int) * P.A.nnz); memcpy(Av, P.A.v, sizeof(double) * P.A.nnz); }
    if (Gi) { memcpy(Gi, gi, sizeof(int) * nnzG); memcpy(Gj, gj, sizeof(int) * nnzG); memcpy(Gv, gv, sizeof(double) * nnzG); }
    if (idx_X) for (int k = 0; k < P.K; k++) for (int i = 0; i < P.nx; i++) idx_X[P.nx * k + i] = iX(&P, k, i);
    if (idx_U) for (int k = 0; k < P.K; k++) for (int i = 0; i < P.nu; i++) idx_U[P.nu * k + i] = iU(&P, k, i);
    if (idx_sigma) *idx_sigma = P.iSigma;
    free(gi); free(gj); free(gv); free(hh);
    socp_free(&P);
    return 0;
}

/* ---- SCAlgorithm::solve (cold start) + iterate, SCAlgorithm.cpp:66-189 ---- */
int orc_sc_solve(int model, const void *params, const orc_sc_config *cfg,
                 double *X_all, double *U_all, double *t_all, orc_iter_info *info,
                 double *X_out, double *U_out, double *t_out, int *converged_out)
{
    return orc_sc_solve2(model, params, cfg, NULL, NULL, 0., NULL, X_all, U_all, t_all, info, X_out, U_out, t_out, converged_out);
}

/* SCAlgorithm::solve(warm_start) (SCAlgorithm.cpp:134-189).  X_warm != NULL: warm start from the DIMENSIONAL trajectory
 * (X_warm, U_warm, t_warm) of a previous solve -- it is nondimensionalised with the scales of the CURRENT x_init (:141-145) and
 * loadParameters() is not called, so the trust-region weight carries over: *weight_tr_io in/out (NULL: SC.info value, not returned). */
int orc_sc_solve2(int model, const void *params, const orc_sc_config *cfg,
                  const double *X_warm, const double *U_warm, double t_warm, double *weight_tr_io,
                  double *X_all, double *U_all, double *t_all, orc_iter_info *info,
                  double *X_out, double *U_out, double *t_out, int *converged_out)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    const int K = cfg->K;
    orc_rq_params rq; orc_r2d_params r2;
    const void *pp;
    double par[ORC_MAX_NP];
    double *X = (double *)malloc(sizeof(double) * K * nx), *U = (double *)malloc(sizeof(double) * K * nu), t;
    double *tdir = (double *)calloc(3 * K, sizeof(double));
    if (model == ORC_MODEL_ROCKETQUAT) {
        rq = *(const orc_rq_params *)params;
        if (cfg->nondimensionalize) orc_rq_nondimensionalize(&rq);                  /* :138-139 */
        if (X_warm) {                                                               /* :141-145, rocketQuat.cpp:175-186 */
            memcpy(X, X_warm, sizeof(double) * K * nx); memcpy(U, U_warm, sizeof(double) * K * nu); t = t_warm;
            if (cfg->nondimensionalize)
                for (int k = 0; k < K; k++) {
                    X[14 * k] /= rq.m_scale;
                    for (int i = 1; i < 7; i++) X[14 * k + i] /= rq.r_scale;
                    for (int i = 0; i < 3; i++) U[4 * k + i] /= rq.m_scale * rq.r_scale;
                    U[4 * k + 3] /= rq.m_scale * rq.r_scale * rq.r_scale;
                }
        } else orc_rq_initial_trajectory(&rq, K, X, U, &t);                         /* :149 */
        orc_rq_model_par(&rq, par);                                                 /* :152 */
        for (int k = 0; k < K; k++) {                                               /* rocketQuat.cpp:162-165 */
            const double *u = U + 4 * k; double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            for (int i = 0; i < 3; i++) tdir[3 * k + i] = n > 0 ? u[i] / n : u[i];
        }
        pp = &rq;
    } else {
        r2 = *(const orc_r2d_params *)params;
        if (cfg->nondimensionalize) orc_r2d_nondimensionalize(&r2);
        if (X_warm) {                                                               /* rocket2d.cpp:97-107 */
            memcpy(X, X_warm, sizeof(double) * K * nx); memcpy(U, U_warm, sizeof(double) * K * nu); t = t_warm;
            if (cfg->nondimensionalize)
                for (int k = 0; k < K; k++) {
                    for (int i = 0; i < 4; i++) X[6 * k + i] /= r2.r_scale;
                    U[2 * k + 1] /= r2.m_scale * r2.r_scale;
                }
        } else orc_r2d_initial_trajectory(&r2, K, X, U, &t);
        orc_r2d_model_par(&r2, par);
        pp = &r2;
    }
    double weight_tr = (X_warm && weight_tr_io) ? *weight_tr_io : cfg->weight_trust_region_trajectory;   /* :148 loadParameters() only on a cold start */
    double *A = (double *)malloc(sizeof(double) * (K - 1) * nx * nx), *B = (double *)malloc(sizeof(double) * (K - 1) * nx * nu);
    double *C = (double *)malloc(sizeof(double) * (K - 1) * nx * nu), *s = (double *)malloc(sizeof(double) * (K - 1) * nx), *z = (double *)malloc(sizeof(double) * (K - 1) * nx);
    double *delta = (double *)malloc(sizeof(double) * K);
    memcpy(X_all, X, sizeof(double) * K * nx); memcpy(U_all, U, sizeof(double) * K * nu); t_all[0] = t; /* :159 */
    int iteration = 0, converged = 0, failed = 0;
    while (iteration < cfg->max_iterations && !converged) {                         /* :161-169 */
        iteration++;
        orc_iter_info *inf = info ? info + (iteration - 1) : NULL;
        double t0 = now_ms();
        orc_discretize(model, K, X, U, t, par, cfg->interpolate_input, cfg->free_final_time, A, B, C, s, z); /* :71 */
        double t1 = now_ms();
        double norm1_nu, delta_sigma, sigma;
        orc_ipm_info ipm;
        int st = orc_sc_subproblem(model, pp, cfg, weight_tr, X, U, t, A, B, C, s, z,
                                   model == ORC_MODEL_ROCKETQUAT ? tdir : NULL,
                                   X, U, &sigma, NULL, delta, &norm1_nu, &delta_sigma, &ipm);  /* :78, readSolution :100 */
        double t2 = now_ms();
        /* status 3 = reduced accuracy (ECOS: ECOS_OPTIMAL + ECOS_INACC_OFFSET, "close to optimal"): the iteration continues, as in the engine;
         * how Epigraph's ECOSSolver::solve maps that exit code to its bool result cannot be checked here (submodule absent) */
        if (st != 0 && st != 3) { failed = 1; if (inf) { memset(inf, 0, sizeof(*inf)); inf->ipm = ipm; } break; } /* :94-98 (terminate) */
        t = cfg->free_final_time ? sigma : t;                                       /* :193-196 */
        double sum_delta = 0;
        for (int k = 0; k < K; k++) sum_delta += delta[k];                          /* :105-107 */
        if (inf) { inf->norm1_nu = norm1_nu; inf->sum_delta = sum_delta; inf->delta_sigma = delta_sigma; inf->sigma = t;
                   inf->weight_tr_used = weight_tr; inf->ipm = ipm; inf->t_discretize_ms = t1 - t0; inf->t_solve_ms = t2 - t1; }
        if (norm1_nu < cfg->nu_tol) weight_tr *= 2.;                                /* :112-115 */
        converged = sum_delta < cfg->delta_tol && norm1_nu < cfg->nu_tol;           /* :131 */
        memcpy(X_all + (size_t)iteration * K * nx, X, sizeof(double) * K * nx);     /* :168 */
        memcpy(U_all + (size_t)iteration * K * nu, U, sizeof(double) * K * nu);
        t_all[iteration] = t;
    }
    /* :182-187 redimensionalise the final trajectory */
    if (X_out) {
        memcpy(X_out, X, sizeof(double) * K * nx); memcpy(U_out, U, sizeof(double) * K * nu); *t_out = t;
        if (cfg->nondimensionalize) {
            if (model == ORC_MODEL_ROCKETQUAT) {                                    /* rocketQuat.cpp:188-201 */
                for (int k = 0; k < K; k++) {
                    X_out[14 * k] *= rq.m_scale;
                    for (int i = 1; i < 7; i++) X_out[14 * k + i] *= rq.r_scale;
                    for (int i = 0; i < 3; i++) U_out[4 * k + i] *= rq.m_scale * rq.r_scale;
                    U_out[4 * k + 3] *= rq.m_scale * rq.r_scale * rq.r_scale;
                }
            } else {                                                                /* rocket2d.cpp:109-119 */
                for (int k = 0; k < K; k++) {
                    for (int i = 0; i < 4; i++) X_out[6 * k + i] *= r2.r_scale;
                    U_out[2 * k + 1] *= r2.m_scale * r2.r_scale;
                }
            }
        }
    }
    if (converged_out) *converged_out = converged;
    if (weight_tr_io) *weight_tr_io = weight_tr;
    free(X); free(U); free(tdir); free(A); free(B); free(C); free(s); free(z); free(delta);
    return failed ? -iteration : iteration;
}

/* scpp::interpolatedInput, scpp/src/commonFunctions.cpp:6-19 */
static void interpolated_input(const double *U, int K, int nu, double t, double total_time, int foh, double *u)
{
    const double time_step = total_time / (K - 1);
    size_t i = (size_t)(t / time_step);
    if (i > (size_t)(K - 2)) i = K - 2;
    const double *u0 = U + nu * i, *u1 = foh ? U + nu * (i + 1) : u0;
    const double ti = fmod(t, time_step) / time_step;
    for (int j = 0; j < nu; j++) u[j] = u0[j] + (u1[j] - u0[j]) * ti;
}

/* SC_sim (scpp/src/SC_sim.cpp:28-65): closed loop  solve(warm) -> apply the first input for time_step on the nonlinear model -> new x_init.
 * X_sim [max_steps][nx], U_sim [max_steps][nu] (dimensional), iters [max_steps] SC iterations of each solve.  Returns the number of
 * simulated steps written; *reached_end as in :56-61; negative on solver failure. */
int orc_sc_sim(int model, const void *params, const orc_sc_config *cfg, double time_step, int max_steps,
               double *X_sim, double *U_sim, int *iters, int *reached_end_out)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    const int K = cfg->K;
    orc_rq_params rq; orc_r2d_params r2;
    void *pp; double *x_init; const double *x_final;
    if (model == ORC_MODEL_ROCKETQUAT) { rq = *(const orc_rq_params *)params; pp = &rq; x_init = rq.x_init; x_final = rq.x_final; }
    else { r2 = *(const orc_r2d_params *)params; pp = &r2; x_init = r2.x_init; x_final = r2.x_final; }
    double *X = (double *)malloc(sizeof(double) * K * nx), *U = (double *)malloc(sizeof(double) * K * nu), t = 0.;
    double *Xa = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1) * K * nx), *Ua = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1) * K * nu);
    double *ta = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1));
    double weight_tr = cfg->weight_trust_region_trajectory;
    int sim_step = 0, written = 0, reached_end = 0, failed = 0;
    while (sim_step < max_steps) {                                                  /* :41 */
        const int warm = sim_step > 0;                                              /* :45 */
        int conv;
        int it = orc_sc_solve2(model, pp, cfg, warm ? X : NULL, warm ? U : NULL, t, &weight_tr, Xa, Ua, ta, NULL, X, U, &t, &conv);   /* :46-47 */
        if (iters) iters[sim_step] = it;
        if (it < 0) { failed = 1; break; }
        double u1[ORC_MAX_NU], par[ORC_MAX_NP];
        interpolated_input(U, K, nu, time_step, t, cfg->interpolate_input, u1);     /* :49-51 */
        if (model == ORC_MODEL_ROCKETQUAT) orc_rq_model_par(&rq, par); else orc_r2d_model_par(&r2, par);   /* dimensional again after :182-186 */
        orc_simulate(model, time_step, U, u1, par, x_init);                         /* :53  (x aliases model->p.x_init, :37) */
        memcpy(X_sim + (size_t)written * nx, x_init, sizeof(double) * nx);          /* :55-56 */
        memcpy(U_sim + (size_t)written * nu, U, sizeof(double) * nu);
        written++;
        double d2 = 0;
        for (int i = 0; i < nx; i++) d2 += (x_init[i] - x_final[i]) * (x_init[i] - x_final[i]);
        reached_end = sqrt(d2) < 0.02 || t < 0.25;                                  /* :58 */
        if (reached_end) break;
        sim_step++;
    }
    if (reached_end_out) *reached_end_out = reached_end;
    free(X); free(U); free(Xa); free(Ua); free(ta);
    return failed ? -written - 1 : written;
}


/* ================================================================================================================
 * SCvx variant: buildSCvxProblem (scpp_core/src/SCvxProblem.cpp:6-71) + SCvxAlgorithm (scpp_core/src/SCvxAlgorithm.cpp)
 * fixed final time, hard trust region on the inputs, ratio test against the nonlinear (simulated) cost
 * ================================================================================================================ */
double orc_scvx_static_reg = 2e-7;   /* test knob: tests/test_oracle.py varies it to show which parts of the optimum are unique */
static void build_scvx_problem(socp_t *P, double weight_vc, double trust_region, const double *Ubar,
                               const double *A, const double *B, const double *C, const double *z)
{
    const int nx = P->nx, nu = P->nu, K = P->K;
    P->has_delta = 0;
    P->stage_stride = nx + nu + nx + nx;                         /* x, u, nu, nu_bound  (:14-17) */
    P->n = (K - 1) * P->stage_stride + (nx + nu);
    P->iN1 = P->n++;                                             /* norm1_nu  :18 */
    P->iSigma = P->iDsigma = -1;
    P->c = (double *)calloc(P->n, sizeof(double));
    for (int k = 0; k < K - 1; k++) {                            /* :20-40 */
        const double *Ak = A + nx * nx * k, *Bk = B + nx * nu * k, *Ck = C + nx * nu * k, *zk = z + nx * k;
        for (int i = 0; i < nx; i++) {
            /* A x_k + B u_k + z + nu + C u_k+1 - x_k+1 = 0 */
            int r = eq_begin(P, -zk[i], iNub(P, k, nx - 1) + 0.5);
            for (int j = 0; j < nx; j++) if (Ak[i + nx * j] != 0.) coo_push(&P->A, r, iX(P, k, j), Ak[i + nx * j]);
            for (int j = 0; j < nu; j++) if (Bk[i + nx * j] != 0.) coo_push(&P->A, r, iU(P, k, j), Bk[i + nx * j]);
            for (int j = 0; j < nu; j++) if (Ck[i + nx * j] != 0.) coo_push(&P->A, r, iU(P, k + 1, j), Ck[i + nx * j]);
            coo_push(&P->A, r, iNu(P, k, i), 1.);
            coo_push(&P->A, r, iX(P, k + 1, i), -1.);
        }
    }
    for (int k = 0; k < K - 1; k++)                              /* :49-50 */
        for (int i = 0; i < nx; i++) {
            int r = lp_begin(P, 0.); lp_coef(P, r, iNub(P, k, i), 1.); lp_coef(P, r, iNu(P, k, i), 1.);
            r = lp_begin(P, 0.);     lp_coef(P, r, iNub(P, k, i), 1.); lp_coef(P, r, iNu(P, k, i), -1.);
        }
    {
        int r = lp_begin(P, 0.);                                 /* sum(nu_bound) <= norm1_nu  :53 */
        lp_coef(P, r, P->iN1, 1.);
        for (int k = 0; k < K - 1; k++) for (int i = 0; i < nx; i++) lp_coef(P, r, iNub(P, k, i), -1.);
        P->c[P->iN1] += weight_vc;                               /* :56 */
    }
    for (int k = 0; k < K; k++) {                                /* norm2(Ubar_k - u_k) <= trust_region  :59-68 (n_U = K with FOH) */
        int r0 = soc_begin(P, 1 + nu);
        soc_const(P, r0, trust_region);
        for (int i = 0; i < nu; i++) { soc_const(P, r0 + 1 + i, Ubar[nu * k + i]); soc_coef(P, r0 + 1 + i, iU(P, k, i), -1.); }
    }
}

/* Ruiz equilibration of the problem data as ECOS does before it factors (its set_equilibration: a few sweeps of square-rooted
 * row / column infinity norms, one common factor per second-order cone), solve, then scale the solution back.  Used for the SCvx
 * sub-problem, whose KKT system has no trust-region block to hide the 1/J_z ~ 2.5e5 entries of the nondimensional B matrices. */
static int conic_solve_equilibrated(int n, int p, int m, int l, int ncones, const int *q, const double *c, const double *b, const double *h,
                                    int nnzA, const int *Ai, const int *Aj, const double *Av, int nnzG, const int *Gi, const int *Gj, const double *Gv,
                                    const double *kv, const double *keq, double *x, double *y, double *sv, double *zv, orc_ipm_info *info)
{
    double *xe = (double *)malloc(sizeof(double) * n), *ae = (double *)malloc(sizeof(double) * (p + 1)), *ge = (double *)malloc(sizeof(double) * m);
    double *xt = (double *)malloc(sizeof(double) * n), *at = (double *)malloc(sizeof(double) * (p + 1)), *gt = (double *)malloc(sizeof(double) * m);
    double *As = (double *)malloc(sizeof(double) * (nnzA + 1)), *Gs = (double *)malloc(sizeof(double) * (nnzG + 1));
    double *cs = (double *)malloc(sizeof(double) * n), *bs = (double *)malloc(sizeof(double) * (p + 1)), *hs = (double *)malloc(sizeof(double) * m);
    memcpy(As, Av, sizeof(double) * nnzA); memcpy(Gs, Gv, sizeof(double) * nnzG);
    for (int j = 0; j < n; j++) xe[j] = 1.;
    for (int i = 0; i < p; i++) ae[i] = 1.;
    for (int i = 0; i < m; i++) ge[i] = 1.;
    for (int sweep = 0; sweep < 3; sweep++) {
        for (int j = 0; j < n; j++) xt[j] = 0.;
        for (int i = 0; i < p; i++) at[i] = 0.;
        for (int i = 0; i < m; i++) gt[i] = 0.;
        for (int e = 0; e < nnzA; e++) { const double a = fabs(As[e]); if (a > xt[Aj[e]]) xt[Aj[e]] = a; if (a > at[Ai[e]]) at[Ai[e]] = a; }
        for (int e = 0; e < nnzG; e++) { const double a = fabs(Gs[e]); if (a > xt[Gj[e]]) xt[Gj[e]] = a; if (a > gt[Gi[e]]) gt[Gi[e]] = a; }
        int o = l;
        for (int k = 0; k < ncones; k++) { double sum = 0.; for (int r = 0; r < q[k]; r++) sum += gt[o + r]; for (int r = 0; r < q[k]; r++) gt[o + r] = sum / q[k]; o += q[k]; }
        for (int j = 0; j < n; j++) xt[j] = xt[j] < 1e-6 ? 1. : sqrt(xt[j]);
        for (int i = 0; i < p; i++) at[i] = at[i] < 1e-6 ? 1. : sqrt(at[i]);
        for (int i = 0; i < m; i++) gt[i] = gt[i] < 1e-6 ? 1. : sqrt(gt[i]);
        for (int e = 0; e < nnzA; e++) As[e] /= at[Ai[e]] * xt[Aj[e]];
        for (int e = 0; e < nnzG; e++) Gs[e] /= gt[Gi[e]] * xt[Gj[e]];
        for (int j = 0; j < n; j++) xe[j] *= xt[j];
        for (int i = 0; i < p; i++) ae[i] *= at[i];
        for (int i = 0; i < m; i++) ge[i] *= gt[i];
    }
    for (int j = 0; j < n; j++) cs[j] = c[j] / xe[j];
    for (int i = 0; i < p; i++) bs[i] = b[i] / ae[i];
    for (int i = 0; i < m; i++) hs[i] = h[i] / ge[i];
    const int st = orc_conic_solve_keys(n, p, m, l, ncones, q, cs, bs, hs, nnzA, Ai, Aj, As, nnzG, Gi, Gj, Gs, kv, keq, x, y, sv, zv, info);
    for (int j = 0; j < n; j++) x[j] /= xe[j];
    for (int i = 0; i < p; i++) y[i] /= ae[i];
    for (int i = 0; i < m; i++) { zv[i] /= ge[i]; sv[i] *= ge[i]; }
    free(xe); free(ae); free(ge); free(xt); free(at); free(gt); free(As); free(Gs); free(cs); free(bs); free(hs);
    return st;
}

int orc_scvx_subproblem(int model, const void *params, int K, double weight_vc, double trust_region,
                        const double *Ubar, const double *A, const double *B, const double *C, const double *z,
                        const double *thrust_dir, double *X, double *U, double *nu, double *norm1_nu, orc_ipm_info *info)
{
    socp_t P;
    int np;
    memset(&P, 0, sizeof(P));
    P.model = model; P.K = K; P.KU = K; P.free_time = 0;
    orc_model_dims(model, &P.nx, &P.nu, &np);
    build_scvx_problem(&P, weight_vc, trust_region, Ubar, A, B, C, z);
    if (model == ORC_MODEL_ROCKETQUAT) add_rq_constraints(&P, (const orc_rq_params *)params, thrust_dir);   /* SCvxAlgorithm.cpp:56 */
    else add_r2d_constraints(&P, (const orc_r2d_params *)params);
    int *Gi, *Gj, nnzG, m, l; double *Gv, *h;
    merged_G(&P, &Gi, &Gj, &Gv, &h, &nnzG, &m, &l);
    double *x = (double *)calloc(P.n, sizeof(double)), *y = (double *)calloc(P.b.n, sizeof(double));
    double *sv = (double *)calloc(m, sizeof(double)), *zv = (double *)calloc(m, sizeof(double));
    double *kv = (double *)malloc(sizeof(double) * P.n);
    for (int j = 0; j < P.n; j++) kv[j] = j;
    /* elimination order of the quasi-definite KKT system (solver detail, not part of the problem): x_0 has no cone row in the SCvx
     * problem, so it is eliminated AFTER the dynamics rows of interval 0, from which it inherits a positive definite block
     * (every later x_k inherits one from the rows of interval k-1); its pinning equalities follow it */
    {
        const double k0 = iNub(&P, 0, P.nx - 1) + 0.5;
        for (int r = 0; r < P.b.n; r++) {
            const double key = P.keq.v[r];
            if (key < P.nx && key != floor(key)) P.keq.v[r] = k0 + 0.001 * (floor(key) + 1.) + 0.0001 * (key - floor(key));
        }
        for (int j = 0; j < P.nx; j++) kv[j] = k0 + 0.001 * (j + 1.);
    }
    /* most states have no cone row here, so the KKT system is only quasi-definite through its static regularisation: use the
     * size ECOS uses (2e-7, with iterative refinement against the unregularised system) instead of the SC path's 1e-13 */
    const double reg_saved = orc_get_static_reg();
    orc_set_static_reg(orc_scvx_static_reg);
    int st = conic_solve_equilibrated(P.n, P.b.n, m, l, P.ncones, P.q, P.c, P.b.v, h, P.A.nnz, P.A.i, P.A.j, P.A.v,
                                      nnzG, Gi, Gj, Gv, kv, P.keq.v, x, y, sv, zv, info);
    orc_set_static_reg(reg_saved);
    const int nx = P.nx, nuu = P.nu;
    for (int k = 0; k < K; k++) {
        if (X) for (int i = 0; i < nx; i++) X[nx * k + i] = x[iX(&P, k, i)];
        if (U) for (int i = 0; i < nuu; i++) U[nuu * k + i] = x[iU(&P, k, i)];
        if (nu && k < K - 1) for (int i = 0; i < nx; i++) nu[nx * k + i] = x[iNu(&P, k, i)];
    }
    if (norm1_nu) *norm1_nu = x[P.iN1];
    free(x); free(y); free(sv); free(zv); free(kv); free(Gi); free(Gj); free(Gv); free(h);
    socp_free(&P);
    return st;
}

/* SCvxAlgorithm::getNonlinearCost, SCvxAlgorithm.cpp:262-278 */
double orc_scvx_nonlinear_cost(int model, int K, const double *X, const double *U, double t, const double *par)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    double cost = 0.;
    for (int k = 0; k < K - 1; k++) {
        double x[ORC_MAX_NX];
        memcpy(x, X + nx * k, sizeof(double) * nx);
        orc_simulate(model, t / (K - 1), U + nu * k, U + nu * (k + 1), par, x);
        for (int i = 0; i < nx; i++) cost += fabs(x[i] - X[nx * (k + 1) + i]);
    }
    return cost;
}

/* SCvxAlgorithm::solve (cold start) + iterate, SCvxAlgorithm.cpp:61-216.  The re-solve loop of a rejected step has no bound in the
 * reference; the restatement stops an outer iteration after ORC_SCVX_MAX_RESOLVE solves and reports failure. */
#define ORC_SCVX_MAX_RESOLVE 40
int orc_scvx_solve(int model, const void *params, const orc_scvx_config *cfg,
                   double *X_all, double *U_all, orc_scvx_info *info,
                   double *X_out, double *U_out, double *t_out, int *converged_out)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    const int K = cfg->K;
    orc_rq_params rq; orc_r2d_params r2;
    const void *pp;
    double par[ORC_MAX_NP];
    double *X = (double *)malloc(sizeof(double) * K * nx), *U = (double *)malloc(sizeof(double) * K * nu), t;
    double *Xo = (double *)malloc(sizeof(double) * K * nx), *Uo = (double *)malloc(sizeof(double) * K * nu);
    double *tdir = (double *)calloc(3 * K, sizeof(double));
    if (!cfg->interpolate_input) { fprintf(stderr, "orc: only interpolate_input=true is restated\n"); abort(); }
    if (model == ORC_MODEL_ROCKETQUAT) {
        rq = *(const orc_rq_params *)params;
        if (cfg->nondimensionalize) orc_rq_nondimensionalize(&rq);                  /* :172-173 */
        orc_rq_initial_trajectory(&rq, K, X, U, &t);                                /* :183 */
        orc_rq_model_par(&rq, par);                                                 /* :186 */
        for (int k = 0; k < K; k++) {                                               /* rocketQuat.cpp:162-165 */
            const double *u = U + 4 * k; double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            for (int i = 0; i < 3; i++) tdir[3 * k + i] = n > 0 ? u[i] / n : u[i];
        }
        pp = &rq;
    } else {
        r2 = *(const orc_r2d_params *)params;
        if (cfg->nondimensionalize) orc_r2d_nondimensionalize(&r2);
        orc_r2d_initial_trajectory(&r2, K, X, U, &t);
        orc_r2d_model_par(&r2, par);
        pp = &r2;
    }
    double trust_region = cfg->trust_region;                                        /* loadParameters() :182 */
    double *A = (double *)malloc(sizeof(double) * (K - 1) * nx * nx), *B = (double *)malloc(sizeof(double) * (K - 1) * nx * nu);
    double *C = (double *)malloc(sizeof(double) * (K - 1) * nx * nu), *z = (double *)malloc(sizeof(double) * (K - 1) * nx);
    memcpy(X_all, X, sizeof(double) * K * nx); memcpy(U_all, U, sizeof(double) * K * nu);   /* :190 */
    int iteration = 0, converged = 0, failed = 0, have_last = 0;
    double last_nonlinear_cost = 0.;
    while (iteration < cfg->max_iterations && !converged && !failed) {              /* :194-201 */
        iteration++;
        orc_scvx_info *inf = info ? info + (iteration - 1) : NULL;
        if (inf) memset(inf, 0, sizeof(*inf));
        orc_discretize(model, K, X, U, t, par, 1, 0, A, B, C, NULL, z);             /* :67 (fixed final time) */
        int solves = 0;
        for (;;) {                                                                  /* :73-146 */
            memcpy(Xo, X, sizeof(double) * K * nx); memcpy(Uo, U, sizeof(double) * K * nu);     /* old_td :76 */
            double norm1_nu;
            orc_ipm_info ipm;
            int st = orc_scvx_subproblem(model, pp, K, cfg->weight_virtual_control, trust_region, Uo, A, B, C, z,
                                         model == ORC_MODEL_ROCKETQUAT ? tdir : NULL, X, U, NULL, &norm1_nu, &ipm);   /* :81, readSolution :94 */
            solves++;
            if (inf) { inf->ipm = ipm; inf->solves = solves; }
            if ((st != 0 && st != 3) || solves > ORC_SCVX_MAX_RESOLVE) { failed = 1; break; }    /* :87-91 (terminate) */
            const double J = orc_scvx_nonlinear_cost(model, K, X, U, t, par);       /* :98 */
            const double L = norm1_nu;                                              /* :100-108 */
            if (inf) { inf->norm1_nu = norm1_nu; inf->nonlinear_cost = J; inf->trust_region_used = trust_region; }
            if (!have_last) { last_nonlinear_cost = J; have_last = 1; break; }      /* :110-114 */
            const double actual_change = last_nonlinear_cost - J, predicted_change = last_nonlinear_cost - L;   /* :116-117 */
            last_nonlinear_cost = J;                                                /* :119 (also when the step is rejected below) */
            if (inf) { inf->actual_change = actual_change; inf->predicted_change = predicted_change; }
            if (fabs(predicted_change) < cfg->change_threshold) { converged = 1; break; }       /* :126-130 */
            const double rho = actual_change / predicted_change;                    /* :132 */
            if (inf) inf->rho = rho;
            if (rho < cfg->rho_0) {                                                 /* :133-139 */
                trust_region /= cfg->alpha;
                memcpy(X, Xo, sizeof(double) * K * nx); memcpy(U, Uo, sizeof(double) * K * nu);
            } else {                                                                /* :140-155 */
                if (rho < cfg->rho_1) trust_region /= cfg->alpha;
                else if (rho >= cfg->rho_2) trust_region *= cfg->beta;
                break;
            }
        }
        if (inf) inf->trust_region_next = trust_region;
        memcpy(X_all + (size_t)iteration * K * nx, X, sizeof(double) * K * nx);     /* :200 */
        memcpy(U_all + (size_t)iteration * K * nu, U, sizeof(double) * K * nu);
    }
    if (X_out) {                                                                    /* :214-219 */
        memcpy(X_out, X, sizeof(double) * K * nx); memcpy(U_out, U, sizeof(double) * K * nu); *t_out = t;
        if (cfg->nondimensionalize) {
            if (model == ORC_MODEL_ROCKETQUAT) {
                for (int k = 0; k < K; k++) {
                    X_out[14 * k] *= rq.m_scale;
                    for (int i = 1; i < 7; i++) X_out[14 * k + i] *= rq.r_scale;
                    for (int i = 0; i < 3; i++) U_out[4 * k + i] *= rq.m_scale * rq.r_scale;
                    U_out[4 * k + 3] *= rq.m_scale * rq.r_scale * rq.r_scale;
                }
            } else {
                for (int k = 0; k < K; k++) {
                    for (int i = 0; i < 4; i++) X_out[6 * k + i] *= r2.r_scale;
                    U_out[2 * k + 1] *= r2.m_scale * r2.r_scale;
                }
            }
        }
    }
    if (converged_out) *converged_out = converged;
    free(X); free(U); free(Xo); free(Uo); free(tdir); free(A); free(B); free(C); free(z);
    return failed ? -iteration : iteration;
}

/* batch driver for the CPU baseline: one instance per thread (the reference itself is single-threaded) */
int orc_sc_solve_batch(int model, int n_inst, const void *params_array, size_t params_stride, const orc_sc_config *cfg,
                       int *iters_out, int *conv_out, double *X_out, double *U_out, double *t_out, int nthreads)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    const int K = cfg->K;
    int total = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic) reduction(+ : total)
    for (int i = 0; i < n_inst; i++) {
        double *Xa = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1) * K * nx);
        double *Ua = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1) * K * nu);
        double *ta = (double *)malloc(sizeof(double) * (cfg->max_iterations + 1));
        int conv = 0;
        int it = orc_sc_solve(model, (const char *)params_array + params_stride * i, cfg, Xa, Ua, ta, NULL,
                              X_out ? X_out + (size_t)i * K * nx : NULL, U_out ? U_out + (size_t)i * K * nu : NULL,
                              t_out ? t_out + i : NULL, &conv);
        if (iters_out) iters_out[i] = it;
        if (conv_out) conv_out[i] = conv;
        total += it > 0 ? it : -it;
        free(Xa); free(Ua); free(ta);
    }
    return total;
}
