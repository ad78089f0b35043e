/*
 * oracle/orc_ipm.c — CPU ORACLE (test infrastructure): generic second-order-cone solver for the
 * ECOS standard form
 *        min c'x   s.t.  A x = b,   G x + s = h,   s in R+^l x Q^{q_1} x ... x Q^{q_N}
 * PARITY UNPINNED (see orc.h): ECOS itself (vendored by the absent Epigraph submodule,
 * lib/Epigraph @ eabeed5fe898, version not determinable) cannot run here.  This file restates the
 * published algorithm family ECOS implements — Mehrotra predictor–corrector primal–dual interior point
 * with Nesterov–Todd scaling, a regularised quasi-definite reduced KKT system factored by LDL' and
 * iterative refinement — without the homogeneous self-dual embedding (the SC sub-problems are always
 * feasible thanks to the virtual control).  Because the SOCP optimality conditions are necessary and
 * sufficient, any point that satisfies them to tolerance is ECOS's answer up to that tolerance when the
 * optimum is unique; every solve returns its certificate (orc_ipm_info).
 * Reference call sites: scpp_core/src/SCAlgorithm.cpp:63,78 (ECOSSolver ctor / solve(false)).
 */
#include "orc.h"
#include "orc_ipm.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define FEASTOL 1e-9
#define ABSTOL 1e-9
#define RELTOL 1e-9
#define MAXIT 100
static __thread double g_static_reg = 1e-13;   /* per thread: the tests run the oracle on several host threads and the SCvx solve sets and restores it (a shared global raced and leaked 2e-7 into later SC solves).  orc_set_static_reg(): the SCvx sub-problem has variables without any cone row and needs ECOS-sized regularisation */
#define STATIC_REG g_static_reg
void orc_set_static_reg(double v) { g_static_reg = v; }
double orc_get_static_reg(void) { return g_static_reg; }
#define STEP_FRAC 0.99
#define EXPAND_THRESHOLD 48 /* LP rows with more nonzeros are kept as explicit KKT rows */

static void *xcalloc(size_t n, size_t sz) { void *p = calloc(n ? n : 1, sz); if (!p) { fprintf(stderr, "orc: out of memory\n"); abort(); } return p; }

/* ---------- triplets -> CSR (duplicates summed) ---------- */
static void coo_to_csr(int nrows, int nnz, const int *I, const int *J, const double *V, int **rp, int **ci, double **cv)
{
    int *ptr = (int *)xcalloc(nrows + 1, sizeof(int));
    for (int e = 0; e < nnz; e++) ptr[I[e] + 1]++;
    for (int i = 0; i < nrows; i++) ptr[i + 1] += ptr[i];
    int *col = (int *)xcalloc(nnz, sizeof(int));
    double *val = (double *)xcalloc(nnz, sizeof(double));
    int *fill = (int *)xcalloc(nrows, sizeof(int));
    for (int e = 0; e < nnz; e++) { int r = I[e], q = ptr[r] + fill[r]++; col[q] = J[e]; val[q] = V[e]; }
    /* sort each row by column and merge duplicates */
    int w = 0;
    int *nptr = (int *)xcalloc(nrows + 1, sizeof(int));
    for (int r = 0; r < nrows; r++) {
        int a = ptr[r], b = ptr[r + 1];
        for (int i = a + 1; i < b; i++) { /* insertion sort */
            int cj = col[i]; double vj = val[i]; int k = i - 1;
            while (k >= a && col[k] > cj) { col[k + 1] = col[k]; val[k + 1] = val[k]; k--; }
            col[k + 1] = cj; val[k + 1] = vj;
        }
        nptr[r] = w;
        for (int i = a; i < b; i++) {
            if (w > nptr[r] && col[w - 1] == col[i]) val[w - 1] += val[i];
            else { col[w] = col[i]; val[w] = val[i]; w++; }
        }
    }
    nptr[nrows] = w;
    free(ptr); free(fill);
    *rp = nptr; *ci = col; *cv = val;
}

typedef struct {
    int n, p, m, l, ncones;
    const int *q;
    int *coff; /* cone row offsets (ncones+1), absolute row index in G */
    int *Ap, *Aj; double *Av;
    int *Gp, *Gj; double *Gv;
    const double *c, *b, *h;
    /* KKT */
    int ne;        /* expanded LP rows */
    int *exp_rows; /* their G-row indices */
    int *exp_of_row; /* G row -> expanded index or -1 */
    int N;         /* n + p + ne */
    int *pos;      /* KKT index (0..N-1: vars, eqs, expanded) -> permuted position */
    int *first, *rowptr;
    double *L, *D; /* envelope factor */
    double *Kenv;  /* assembled matrix (same layout) */
    int *sign;     /* expected pivot sign per permuted position */
    /* NT scaling */
    double *wl;    /* LP: w_i^2 = s_i/z_i  (size l) */
    double *eta;   /* per cone */
    double *wbar;  /* per cone row (normalised NT point, w0^2-|w1|^2=1) */
    double *lambda;
    double kkt_resid;
    /* work arrays allocated once per solve (the CPU baseline runs many solves in parallel: no malloc in the hot loop) */
    double *w_rhs, *w_sol, *w_wrk, *w_res, *w_t1, *w_t2, *w_ex, *w_ey, *w_ez, *w_cx, *w_cy, *w_cz, *w_w1, *w_w2;
} ipm_t;

/* ---------- cone helpers ---------- */
static double soc_res(const double *u, int d) { double n = 0; for (int i = 1; i < d; i++) n += u[i] * u[i]; return u[0] - sqrt(n); }
static double jnorm2(const double *u, int d) { double n = 0; for (int i = 1; i < d; i++) n += u[i] * u[i]; return u[0] * u[0] - n; }

static double cone_min_margin(const ipm_t *S, const double *u)
{
    double mn = 1e300;
    for (int i = 0; i < S->l; i++) if (u[i] < mn) mn = u[i];
    for (int k = 0; k < S->ncones; k++) { double r = soc_res(u + S->coff[k], S->q[k]); if (r < mn) mn = r; }
    return mn;
}
static void cone_add_e(const ipm_t *S, double *u, double a)
{
    for (int i = 0; i < S->l; i++) u[i] += a;
    for (int k = 0; k < S->ncones; k++) u[S->coff[k]] += a;
}

/* NT scaling from (s,z); lambda = W z = W^-1 s */
static int compute_scaling(ipm_t *S, const double *s, const double *z)
{
    for (int i = 0; i < S->l; i++) {
        if (!(s[i] > 0) || !(z[i] > 0)) return -1;
        S->wl[i] = s[i] / z[i];
        S->lambda[i] = sqrt(s[i] * z[i]);
    }
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        const double *sk = s + o, *zk = z + o;
        double ss = jnorm2(sk, d), zz = jnorm2(zk, d);
        if (!(ss > 0) || !(zz > 0) || !(sk[0] > 0) || !(zk[0] > 0)) return -1;
        double sn = sqrt(ss), zn = sqrt(zz);
        double sz = 0;
        for (int i = 0; i < d; i++) sz += sk[i] * zk[i];
        double gamma = sqrt((1. + sz / (sn * zn)) / 2.);
        double *w = S->wbar + o;
        /* wbar = (sbar + J zbar) / (2 gamma) */
        w[0] = (sk[0] / sn + zk[0] / zn) / (2 * gamma);
        for (int i = 1; i < d; i++) w[i] = (sk[i] / sn - zk[i] / zn) / (2 * gamma);
        S->eta[k] = sqrt(sn / zn);
        /* lambda = W z */
        double eta = S->eta[k];
        double w1z1 = 0;
        for (int i = 1; i < d; i++) w1z1 += w[i] * zk[i];
        double *lam = S->lambda + o;
        lam[0] = eta * (w[0] * zk[0] + w1z1);
        double f = (zk[0] + w1z1 / (1. + w[0]));
        for (int i = 1; i < d; i++) lam[i] = eta * (zk[i] + f * w[i]);
    }
    return 0;
}
/* out = W v (inv=0) or W^-1 v (inv=1), cone-wise */
static void apply_W(const ipm_t *S, const double *v, double *out, int inv)
{
    for (int i = 0; i < S->l; i++) { double w = sqrt(S->wl[i]); out[i] = inv ? v[i] / w : v[i] * w; }
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        const double *w = S->wbar + o, *vk = v + o;
        double eta = S->eta[k], sg = inv ? -1. : 1., sc = inv ? 1. / eta : eta;
        double w1v1 = 0;
        for (int i = 1; i < d; i++) w1v1 += w[i] * vk[i];
        double o0 = w[0] * vk[0] + sg * w1v1;
        double f = sg * vk[0] + w1v1 / (1. + w[0]);
        for (int i = 1; i < d; i++) out[o + i] = sc * (vk[i] + f * w[i]);
        out[o] = sc * o0;
    }
}
/* out = W^-2 v */
static void apply_Winv2(const ipm_t *S, const double *v, double *out)
{
    for (int i = 0; i < S->l; i++) out[i] = v[i] / S->wl[i];
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        const double *w = S->wbar + o, *vk = v + o;
        double e2 = 1. / (S->eta[k] * S->eta[k]);
        /* W^-2 = eta^-2 (2 what what' - J), what = (w0,-w1) */
        double dot = w[0] * vk[0];
        for (int i = 1; i < d; i++) dot -= w[i] * vk[i];
        out[o] = e2 * (2 * dot * w[0] - vk[0]);
        for (int i = 1; i < d; i++) out[o + i] = e2 * (-2 * dot * w[i] + vk[i]);
    }
}
/* Jordan product and inverse */
static void jprod(const ipm_t *S, const double *u, const double *v, double *out)
{
    for (int i = 0; i < S->l; i++) out[i] = u[i] * v[i];
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        double dot = 0;
        for (int i = 0; i < d; i++) dot += u[o + i] * v[o + i];
        double u0 = u[o], v0 = v[o];
        for (int i = 1; i < d; i++) out[o + i] = u0 * v[o + i] + v0 * u[o + i];
        out[o] = dot;
    }
}
static void jdiv(const ipm_t *S, const double *lam, const double *dv, double *out) /* out = lam \ d */
{
    for (int i = 0; i < S->l; i++) out[i] = dv[i] / lam[i];
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        double den = jnorm2(lam + o, d), l1d1 = 0;
        for (int i = 1; i < d; i++) l1d1 += lam[o + i] * dv[o + i];
        double x0 = (lam[o] * dv[o] - l1d1) / den;
        for (int i = 1; i < d; i++) out[o + i] = (dv[o + i] - x0 * lam[o + i]) / lam[o];
        out[o] = x0;
    }
}
/* largest t such that lambda + (1/t) d hits the boundary (scaled space); returns max(0, ...) */
static double max_step_t(const ipm_t *S, const double *d)
{
    double t = 0;
    for (int i = 0; i < S->l; i++) { double r = -d[i] / S->lambda[i]; if (r > t) t = r; }
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], dd = S->q[k];
        const double *lam = S->lambda + o, *dk = d + o;
        double a = sqrt(jnorm2(lam, dd));
        double l0 = lam[0] / a;
        double ld = l0 * dk[0];
        for (int i = 1; i < dd; i++) ld -= lam[i] / a * dk[i];
        double rho0 = ld / a;
        double f = (ld + dk[0]) / (l0 + 1.);
        double n1 = 0;
        for (int i = 1; i < dd; i++) { double r = (dk[i] - f * lam[i] / a) / a; n1 += r * r; }
        double r = sqrt(n1) - rho0;
        if (r > t) t = r;
    }
    return t;
}

/* ---------- KKT structure ---------- */
static inline double *ent(ipm_t *S, double *M, int i, int j) /* permuted positions, i>=j */
{
    return M + S->rowptr[i] + (j - S->first[i]);
}
static inline void touch(ipm_t *S, int a, int b) { int i = a > b ? a : b, j = a > b ? b : a; if (j < S->first[i]) S->first[i] = j; }

static void build_structure(ipm_t *S, const double *keys_var, const double *keys_eq)
{
    const int n = S->n, p = S->p;
    /* expanded rows */
    S->exp_of_row = (int *)xcalloc(S->m, sizeof(int));
    S->exp_rows = (int *)xcalloc(S->l, sizeof(int));
    S->ne = 0;
    for (int i = 0; i < S->m; i++) S->exp_of_row[i] = -1;
    for (int i = 0; i < S->l; i++)
        if (S->Gp[i + 1] - S->Gp[i] > EXPAND_THRESHOLD) { S->exp_of_row[i] = S->ne; S->exp_rows[S->ne++] = i; }
    S->N = n + p + S->ne;
    const int N = S->N;
    /* ordering keys */
    double *key = (double *)xcalloc(N, sizeof(double));
    for (int j = 0; j < n; j++) key[j] = keys_var ? keys_var[j] : (double)j;
    for (int i = 0; i < p; i++) {
        if (keys_eq) key[n + i] = keys_eq[i];
        else { int mx = 0; for (int e = S->Ap[i]; e < S->Ap[i + 1]; e++) if (S->Aj[e] > mx) mx = S->Aj[e]; key[n + i] = mx + 0.5; }
    }
    for (int e = 0; e < S->ne; e++) key[n + p + e] = 1e300;
    int *order = (int *)xcalloc(N, sizeof(int));
    for (int i = 0; i < N; i++) order[i] = i;
    /* stable merge sort by key */
    {
        int *tmp = (int *)xcalloc(N, sizeof(int));
        for (int w = 1; w < N; w *= 2) {
            for (int lo = 0; lo < N; lo += 2 * w) {
                int mid = lo + w < N ? lo + w : N, hi = lo + 2 * w < N ? lo + 2 * w : N;
                int a = lo, b = mid, o = lo;
                while (a < mid && b < hi) tmp[o++] = (key[order[b]] < key[order[a]]) ? order[b++] : order[a++];
                while (a < mid) tmp[o++] = order[a++];
                while (b < hi) tmp[o++] = order[b++];
            }
            memcpy(order, tmp, sizeof(int) * N);
        }
        free(tmp);
    }
    S->pos = (int *)xcalloc(N, sizeof(int));
    S->sign = (int *)xcalloc(N, sizeof(int));
    for (int r = 0; r < N; r++) { S->pos[order[r]] = r; S->sign[r] = order[r] < n ? 1 : -1; }
    free(order); free(key);
    /* envelope */
    S->first = (int *)xcalloc(N, sizeof(int));
    for (int i = 0; i < N; i++) S->first[i] = i;
    for (int i = 0; i < S->l; i++) {
        int a = S->Gp[i], b = S->Gp[i + 1];
        if (S->exp_of_row[i] >= 0) {
            int pr = S->pos[n + p + S->exp_of_row[i]];
            for (int e = a; e < b; e++) touch(S, pr, S->pos[S->Gj[e]]);
        } else {
            int mn = N;
            for (int e = a; e < b; e++) if (S->pos[S->Gj[e]] < mn) mn = S->pos[S->Gj[e]];
            for (int e = a; e < b; e++) touch(S, S->pos[S->Gj[e]], mn);
        }
    }
    for (int k = 0; k < S->ncones; k++) {
        int a = S->Gp[S->coff[k]], b = S->Gp[S->coff[k + 1]];
        int mn = N;
        for (int e = a; e < b; e++) if (S->pos[S->Gj[e]] < mn) mn = S->pos[S->Gj[e]];
        for (int e = a; e < b; e++) touch(S, S->pos[S->Gj[e]], mn);
    }
    for (int i = 0; i < p; i++) {
        int pr = S->pos[n + i];
        for (int e = S->Ap[i]; e < S->Ap[i + 1]; e++) touch(S, pr, S->pos[S->Aj[e]]);
    }
    S->rowptr = (int *)xcalloc(N + 1, sizeof(int));
    for (int i = 0; i < N; i++) S->rowptr[i + 1] = S->rowptr[i] + (i - S->first[i] + 1);
    S->L = (double *)xcalloc(S->rowptr[N], sizeof(double));
    S->Kenv = (double *)xcalloc(S->rowptr[N], sizeof(double));
    S->D = (double *)xcalloc(N, sizeof(double));
}

/* assemble reduced KKT  [H+eps I, A', Ge'; A, -eps I, 0; Ge, 0, -We^2]  (lower envelope) */
static void assemble_kkt(ipm_t *S, int identity_scaling)
{
    const int n = S->n, p = S->p, N = S->N;
    double *K = S->Kenv;
    memset(K, 0, sizeof(double) * S->rowptr[N]);
    for (int i = 0; i < S->l; i++) {
        const int a = S->Gp[i], b = S->Gp[i + 1];
        const double w2 = identity_scaling ? 1. : S->wl[i];
        if (S->exp_of_row[i] >= 0) {
            int pr = S->pos[n + p + S->exp_of_row[i]];
            for (int e = a; e < b; e++) { int pc = S->pos[S->Gj[e]]; *ent(S, K, pr > pc ? pr : pc, pr > pc ? pc : pr) += S->Gv[e]; }
            *ent(S, K, pr, pr) -= w2;
        } else {
            const double dinv = 1. / w2;
            for (int e = a; e < b; e++)
                for (int f = a; f < b; f++) {
                    int pi = S->pos[S->Gj[e]], pj = S->pos[S->Gj[f]];
                    if (pi >= pj) *ent(S, K, pi, pj) += dinv * S->Gv[e] * S->Gv[f];
                }
        }
    }
    for (int k = 0; k < S->ncones; k++) {
        const int o = S->coff[k], d = S->q[k];
        const double e2 = identity_scaling ? 1. : 1. / (S->eta[k] * S->eta[k]);
        /* H += e2 * ( 2 v v' - g0 g0' + sum_{r>=1} g_r g_r' ),  v = sum_r what_r g_r ; identity: sum_r g_r g_r' */
        /* gather union of columns */
        int cols[256]; double v[256]; int nc = 0;
        for (int r = 0; r < d; r++)
            for (int e = S->Gp[o + r]; e < S->Gp[o + r + 1]; e++) {
                int cj = S->Gj[e], q;
                for (q = 0; q < nc; q++) if (cols[q] == cj) break;
                if (q == nc) { if (nc >= 256) { fprintf(stderr, "orc: cone touches too many columns\n"); abort(); } cols[nc] = cj; v[nc] = 0; nc++; }
                if (!identity_scaling) { double wh = (r == 0) ? S->wbar[o] : -S->wbar[o + r]; v[q] += wh * S->Gv[e]; }
            }
        if (!identity_scaling)
            for (int a = 0; a < nc; a++)
                for (int b = 0; b < nc; b++) {
                    int pi = S->pos[cols[a]], pj = S->pos[cols[b]];
                    if (pi >= pj) *ent(S, K, pi, pj) += e2 * 2 * v[a] * v[b];
                }
        for (int r = 0; r < d; r++) {
            const double sg = (identity_scaling || r > 0) ? 1. : -1.;
            for (int e = S->Gp[o + r]; e < S->Gp[o + r + 1]; e++)
                for (int f = S->Gp[o + r]; f < S->Gp[o + r + 1]; f++) {
                    int pi = S->pos[S->Gj[e]], pj = S->pos[S->Gj[f]];
                    if (pi >= pj) *ent(S, K, pi, pj) += e2 * sg * S->Gv[e] * S->Gv[f];
                }
        }
    }
    for (int i = 0; i < p; i++) {
        int pr = S->pos[n + i];
        for (int e = S->Ap[i]; e < S->Ap[i + 1]; e++) { int pc = S->pos[S->Aj[e]]; *ent(S, K, pr > pc ? pr : pc, pr > pc ? pc : pr) += S->Av[e]; }
    }
    for (int r = 0; r < N; r++) *ent(S, K, r, r) += S->sign[r] * STATIC_REG;
}

static void factor_kkt(ipm_t *S)
{
    const int N = S->N;
    memcpy(S->L, S->Kenv, sizeof(double) * S->rowptr[N]);
    for (int i = 0; i < N; i++) {
        const int fi = S->first[i];
        double *Li = S->L + S->rowptr[i] - fi; /* Li[j] valid for j in [fi,i] */
        for (int j = fi; j < i; j++) {
            const int fj = S->first[j];
            const double *Lj = S->L + S->rowptr[j] - fj;
            int k0 = fi > fj ? fi : fj;
            double acc = Li[j];
            for (int k = k0; k < j; k++) acc -= Li[k] * Lj[k]; /* Li[k] holds y_k = L_ik D_k */
            Li[j] = acc;
        }
        double d = Li[i];
        for (int j = fi; j < i; j++) { double y = Li[j]; double lij = y / S->D[j]; d -= y * lij; Li[j] = lij; }
        if (S->sign[i] > 0 ? !(d > 1e-14) : !(d < -1e-14)) d = S->sign[i] * 1e-10; /* dynamic regularisation */
        S->D[i] = d;
        Li[i] = 1.;
    }
}
static void solve_factored(const ipm_t *S, double *x /* permuted, in/out */)
{
    const int N = S->N;
    for (int i = 0; i < N; i++) {
        const int fi = S->first[i];
        const double *Li = S->L + S->rowptr[i] - fi;
        double acc = x[i];
        for (int j = fi; j < i; j++) acc -= Li[j] * x[j];
        x[i] = acc;
    }
    for (int i = 0; i < N; i++) x[i] /= S->D[i];
    for (int i = N - 1; i >= 0; i--) {
        const int fi = S->first[i];
        const double *Li = S->L + S->rowptr[i] - fi;
        const double xi = x[i];
        for (int j = fi; j < i; j++) x[j] -= Li[j] * xi;
    }
}
/* y = K_true * x (no static regularisation), unpermuted index space [vars|eqs|expanded] */
static void kkt_mult(const ipm_t *S, const double *x, double *y, int identity_scaling, double *tmp_m, double *tmp_m2)
{
    const int n = S->n, p = S->p, m = S->m;
    for (int i = 0; i < S->N; i++) y[i] = 0;
    /* G x */
    for (int i = 0; i < m; i++) { double acc = 0; for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) acc += S->Gv[e] * x[S->Gj[e]]; tmp_m[i] = acc; }
    if (identity_scaling) memcpy(tmp_m2, tmp_m, sizeof(double) * m); else apply_Winv2(S, tmp_m, tmp_m2);
    for (int i = 0; i < m; i++) {
        if (i < S->l && S->exp_of_row[i] >= 0) {
            int ei = S->exp_of_row[i];
            double ze = x[n + p + ei];
            for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) y[S->Gj[e]] += S->Gv[e] * ze;
            y[n + p + ei] = tmp_m[i] - (identity_scaling ? 1. : S->wl[i]) * ze;
        } else
            for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) y[S->Gj[e]] += S->Gv[e] * tmp_m2[i];
    }
    for (int i = 0; i < p; i++) {
        double acc = 0;
        for (int e = S->Ap[i]; e < S->Ap[i + 1]; e++) { acc += S->Av[e] * x[S->Aj[e]]; y[S->Aj[e]] += S->Av[e] * x[n + i]; }
        y[n + i] = acc;
    }
}

/* Solve  A'dy + G'dz = rx ; A dx = ry ; G dx - W^2 dz = rz   (W = I when identity_scaling) */
static void kkt_solve_inner(ipm_t *S, const double *rx, const double *ry, const double *rz,
                      double *dx, double *dy, double *dz, int identity_scaling)
{
    const int n = S->n, p = S->p, m = S->m, N = S->N;
    double *rhs = S->w_rhs, *sol = S->w_sol, *wrk = S->w_wrk, *res = S->w_res, *t1 = S->w_t1, *t2 = S->w_t2;
    memset(rhs, 0, sizeof(double) * N);
    /* rhs_x = rx + G_ne' W^-2 rz_ne */
    if (identity_scaling) memcpy(t1, rz, sizeof(double) * m); else apply_Winv2(S, rz, t1);
    for (int j = 0; j < n; j++) rhs[j] = rx[j];
    for (int i = 0; i < m; i++) {
        if (i < S->l && S->exp_of_row[i] >= 0) { rhs[n + p + S->exp_of_row[i]] = rz[i]; continue; }
        for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) rhs[S->Gj[e]] += S->Gv[e] * t1[i];
    }
    for (int i = 0; i < p; i++) rhs[n + i] = ry[i];
    /* solve + refine */
    for (int i = 0; i < N; i++) wrk[S->pos[i]] = rhs[i];
    solve_factored(S, wrk);
    for (int i = 0; i < N; i++) sol[i] = wrk[S->pos[i]];
    double best = 1e300, nrhs = 0;
    for (int i = 0; i < N; i++) if (fabs(rhs[i]) > nrhs) nrhs = fabs(rhs[i]);
    for (int it = 0; it < 8; it++) {
        kkt_mult(S, sol, res, identity_scaling, t1, t2);
        double nr = 0;
        for (int i = 0; i < N; i++) { res[i] = rhs[i] - res[i]; if (fabs(res[i]) > nr) nr = fabs(res[i]); }
        nr /= fmax(1e-300, nrhs);
        if (nr >= 0.5 * best && it > 0) { if (nr < best) best = nr; break; }
        best = nr;
        if (nr < 1e-15 || it == 7) break;
        for (int i = 0; i < N; i++) wrk[S->pos[i]] = res[i];
        solve_factored(S, wrk);
        for (int i = 0; i < N; i++) sol[i] += wrk[S->pos[i]];
    }
    memcpy(dx, sol, sizeof(double) * n);
    memcpy(dy, sol + n, sizeof(double) * p);
    /* dz = W^-2 (G dx - rz) ; expanded rows from the solve */
    for (int i = 0; i < m; i++) { double acc = -rz[i]; for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) acc += S->Gv[e] * dx[S->Gj[e]]; t1[i] = acc; }
    if (identity_scaling) memcpy(dz, t1, sizeof(double) * m); else apply_Winv2(S, t1, dz);
    for (int e = 0; e < S->ne; e++) dz[S->exp_rows[e]] = sol[n + p + e];
}

/* outer refinement against the unreduced system [0 A' G'; A 0 0; G 0 -W^2] (the reduced normal-equation
 * form loses accuracy in rx when W^-2 spans many orders of magnitude) */
static void kkt_solve(ipm_t *S, const double *rx, const double *ry, const double *rz,
                      double *dx, double *dy, double *dz, int identity_scaling)
{
    const int n = S->n, p = S->p, m = S->m;
    kkt_solve_inner(S, rx, ry, rz, dx, dy, dz, identity_scaling);
    double *ex = S->w_ex, *ey = S->w_ey, *ez = S->w_ez, *cx = S->w_cx, *cy = S->w_cy, *cz = S->w_cz, *w1 = S->w_w1, *w2 = S->w_w2;
    double prev = 1e300;
    for (int it = 0; it < 3; it++) {
        for (int j = 0; j < n; j++) ex[j] = rx[j];
        for (int i = 0; i < p; i++) { double acc = ry[i]; for (int e = S->Ap[i]; e < S->Ap[i + 1]; e++) { acc -= S->Av[e] * dx[S->Aj[e]]; ex[S->Aj[e]] -= S->Av[e] * dy[i]; } ey[i] = acc; }
        if (identity_scaling) memcpy(w2, dz, sizeof(double) * m); else { apply_W(S, dz, w1, 0); apply_W(S, w1, w2, 0); }
        for (int i = 0; i < m; i++) { double acc = rz[i] + w2[i]; for (int e = S->Gp[i]; e < S->Gp[i + 1]; e++) { acc -= S->Gv[e] * dx[S->Gj[e]]; ex[S->Gj[e]] -= S->Gv[e] * dz[i]; } ez[i] = acc; }
        double nr = 0, nb = 0;
        for (int j = 0; j < n; j++) { if (fabs(ex[j]) > nr) nr = fabs(ex[j]); if (fabs(rx[j]) > nb) nb = fabs(rx[j]); }
        for (int i = 0; i < p; i++) { if (fabs(ey[i]) > nr) nr = fabs(ey[i]); if (fabs(ry[i]) > nb) nb = fabs(ry[i]); }
        nr /= fmax(nb, 1e-300);
        S->kkt_resid = nr;
        if (nr < 1e-14 || nr > 0.5 * prev) break;
        prev = nr;
        kkt_solve_inner(S, ex, ey, ez, cx, cy, cz, identity_scaling);
        for (int j = 0; j < n; j++) dx[j] += cx[j];
        for (int i = 0; i < p; i++) dy[i] += cy[i];
        for (int i = 0; i < m; i++) dz[i] += cz[i];
    }
}

static double nrm2(const double *v, int n) { double a = 0; for (int i = 0; i < n; i++) a += v[i] * v[i]; return sqrt(a); }
static double dot(const double *a, const double *b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }

int orc_conic_solve_keys(int n, int p, int m, int l, int ncones, const int *q,
                         const double *c, const double *b, const double *h,
                         int nnzA, const int *Ai, const int *Aj, const double *Av,
                         int nnzG, const int *Gi, const int *Gj, const double *Gv,
                         const double *keys_var, const double *keys_eq,
                         double *x, double *y, double *s, double *z, orc_ipm_info *info)
{
    ipm_t S;
    memset(&S, 0, sizeof(S));
    S.n = n; S.p = p; S.m = m; S.l = l; S.ncones = ncones; S.q = q; S.c = c; S.b = b; S.h = h;
    S.coff = (int *)xcalloc(ncones + 1, sizeof(int));
    S.coff[0] = l;
    for (int k = 0; k < ncones; k++) S.coff[k + 1] = S.coff[k] + q[k];
    coo_to_csr(p, nnzA, Ai, Aj, Av, &S.Ap, &S.Aj, &S.Av);
    coo_to_csr(m, nnzG, Gi, Gj, Gv, &S.Gp, &S.Gj, &S.Gv);
    /* presolve: exact duplicates among singleton equality rows (rocketQuat.cpp:79,86-88,141-142 fix some
     * variables twice; ECOS absorbs the rank deficiency through regularisation) are neutralised */
    double *bb = (double *)xcalloc(p, sizeof(double));
    memcpy(bb, b, sizeof(double) * p);
    {
        int *seen = (int *)xcalloc(n, sizeof(int));
        for (int j = 0; j < n; j++) seen[j] = -1;
        for (int i = 0; i < p; i++) {
            if (S.Ap[i + 1] - S.Ap[i] != 1) continue;
            int j = S.Aj[S.Ap[i]];
            if (seen[j] < 0) { seen[j] = i; continue; }
            int i0 = seen[j];
            double v0 = bb[i0] / S.Av[S.Ap[i0]], v1 = bb[i] / S.Av[S.Ap[i]];
            if (fabs(v0 - v1) <= 1e-12 * fmax(1., fabs(v0))) { S.Av[S.Ap[i]] = 0.; bb[i] = 0.; }
        }
        free(seen);
    }
    b = bb; S.b = bb;
    build_structure(&S, keys_var, keys_eq);
    S.wl = (double *)xcalloc(l, sizeof(double));
    S.eta = (double *)xcalloc(ncones, sizeof(double));
    S.wbar = (double *)xcalloc(m, sizeof(double));
    S.lambda = (double *)xcalloc(m, sizeof(double));
    S.w_rhs = (double *)xcalloc(S.N, sizeof(double)); S.w_sol = (double *)xcalloc(S.N, sizeof(double));
    S.w_wrk = (double *)xcalloc(S.N, sizeof(double)); S.w_res = (double *)xcalloc(S.N, sizeof(double));
    S.w_t1 = (double *)xcalloc(m, sizeof(double)); S.w_t2 = (double *)xcalloc(m, sizeof(double));
    S.w_ex = (double *)xcalloc(n, sizeof(double)); S.w_ey = (double *)xcalloc(p, sizeof(double)); S.w_ez = (double *)xcalloc(m, sizeof(double));
    S.w_cx = (double *)xcalloc(n, sizeof(double)); S.w_cy = (double *)xcalloc(p, sizeof(double)); S.w_cz = (double *)xcalloc(m, sizeof(double));
    S.w_w1 = (double *)xcalloc(m, sizeof(double)); S.w_w2 = (double *)xcalloc(m, sizeof(double));

    double *rx = (double *)xcalloc(n, sizeof(double)), *ry = (double *)xcalloc(p, sizeof(double)), *rz = (double *)xcalloc(m, sizeof(double));
    double *dx = (double *)xcalloc(n, sizeof(double)), *dy = (double *)xcalloc(p, sizeof(double)), *dz = (double *)xcalloc(m, sizeof(double)), *ds = (double *)xcalloc(m, sizeof(double));
    double *dxa = (double *)xcalloc(n, sizeof(double)), *dya = (double *)xcalloc(p, sizeof(double)), *dza = (double *)xcalloc(m, sizeof(double)), *dsa = (double *)xcalloc(m, sizeof(double));
    double *t1 = (double *)xcalloc(m, sizeof(double)), *t2 = (double *)xcalloc(m, sizeof(double)), *t3 = (double *)xcalloc(m, sizeof(double));
    double *zero_n = (double *)xcalloc(n, sizeof(double)), *zero_p = (double *)xcalloc(p, sizeof(double)), *zero_m = (double *)xcalloc(m, sizeof(double));
    double *negc = (double *)xcalloc(n, sizeof(double));

    /* ---- initial point (CVXOPT conelp / ECOS style) ---- */
    assemble_kkt(&S, 1);
    factor_kkt(&S);
    kkt_solve(&S, zero_n, b, h, x, dy, t1, 1); /* x ; t1 = G x - h */
    for (int i = 0; i < m; i++) s[i] = -t1[i];
    { double mg = cone_min_margin(&S, s); if (mg <= 1e-8 * fmax(1., nrm2(s, m))) cone_add_e(&S, s, 1. - mg); }
    for (int j = 0; j < n; j++) negc[j] = -c[j];
    kkt_solve(&S, negc, zero_p, zero_m, dx, y, z, 1);
    { double mg = cone_min_margin(&S, z); if (mg <= 1e-8 * fmax(1., nrm2(z, m))) cone_add_e(&S, z, 1. - mg); }

    const double resx0 = fmax(1., nrm2(c, n)), resy0 = fmax(1., nrm2(b, p)), resz0 = fmax(1., nrm2(h, m));
    const int degree = l + ncones;
    int status = 1, it;
    double pres = 0, dres = 0, gap = 0, relgap = 0, pcost = 0, dcost = 0;
    S.kkt_resid = 0;
    double *bx = (double *)xcalloc(n, sizeof(double)), *by = (double *)xcalloc(p, sizeof(double)), *bs = (double *)xcalloc(m, sizeof(double)), *bz = (double *)xcalloc(m, sizeof(double));
    double best_score = 1e300, b_pres = 0, b_dres = 0, b_gap = 0, b_relgap = 0, b_pcost = 0, b_dcost = 0;
    int have_best = 0;
    for (it = 0; it <= MAXIT; it++) {
        /* residuals */
        for (int j = 0; j < n; j++) rx[j] = c[j];
        for (int i = 0; i < p; i++) { double acc = -b[i]; for (int e = S.Ap[i]; e < S.Ap[i + 1]; e++) { acc += S.Av[e] * x[S.Aj[e]]; rx[S.Aj[e]] += S.Av[e] * y[i]; } ry[i] = acc; }
        for (int i = 0; i < m; i++) { double acc = s[i] - h[i]; for (int e = S.Gp[i]; e < S.Gp[i + 1]; e++) { acc += S.Gv[e] * x[S.Gj[e]]; rx[S.Gj[e]] += S.Gv[e] * z[i]; } rz[i] = acc; }
        gap = dot(s, z, m);
        pcost = dot(c, x, n);
        dcost = -dot(b, y, p) - dot(h, z, m);
        pres = fmax(nrm2(ry, p) / resy0, nrm2(rz, m) / resz0);
        dres = nrm2(rx, n) / resx0;
        if (pcost < 0) relgap = gap / -pcost; else if (dcost > 0) relgap = gap / dcost; else relgap = 1e300;
        if (pres <= FEASTOL && dres <= FEASTOL && (gap <= ABSTOL || relgap <= RELTOL)) { status = 0; break; }
        {
            double score = fmax(fmax(pres, dres) / FEASTOL, fmin(gap / ABSTOL, relgap / RELTOL));
            if (score < best_score) {
                best_score = score; have_best = 1;
                memcpy(bx, x, sizeof(double) * n); memcpy(by, y, sizeof(double) * p); memcpy(bs, s, sizeof(double) * m); memcpy(bz, z, sizeof(double) * m);
                b_pres = pres; b_dres = dres; b_gap = gap; b_relgap = relgap; b_pcost = pcost; b_dcost = dcost;
            } else if (score > 1e3 * best_score && best_score < 1e4) { status = 2; break; } /* diverging after near-convergence */
        }
        if (it == MAXIT) break;
        if (compute_scaling(&S, s, z)) { status = 2; break; }
        assemble_kkt(&S, 0);
        factor_kkt(&S);
        /* affine direction: rhs_z_eff = -rz + s */
        for (int j = 0; j < n; j++) dx[j] = -rx[j];
        for (int i = 0; i < p; i++) dy[i] = -ry[i];
        for (int i = 0; i < m; i++) t1[i] = -rz[i] + s[i];
        kkt_solve(&S, dx, dy, t1, dxa, dya, dza, 0);
        /* scaled directions: dz~ = W dz ; ds~ = W^-1 ds = -lambda - W dz */
        apply_W(&S, dza, t2, 0);
        for (int i = 0; i < m; i++) t1[i] = -S.lambda[i] - t2[i];
        double ts = max_step_t(&S, t1), tz = max_step_t(&S, t2);
        double tt = fmax(ts, tz);
        double alpha_aff = tt <= 1. ? 1. : 1. / tt;
        double sigma = pow(1. - alpha_aff, 3), mu = gap / degree;
        /* combined: d_s = -lambda o lambda - (W^-1 ds_a) o (W dz_a) + sigma mu e */
        jprod(&S, S.lambda, S.lambda, t3);
        double *cross = ds; /* reuse */
        jprod(&S, t1, t2, cross);
        for (int i = 0; i < m; i++) t3[i] = -t3[i] - cross[i];
        cone_add_e(&S, t3, sigma * mu);
        jdiv(&S, S.lambda, t3, t1);   /* lambda \ d_s */
        apply_W(&S, t1, t2, 0);       /* W (lambda \ d_s) */
        for (int j = 0; j < n; j++) dxa[j] = -(1. - sigma) * rx[j];
        for (int i = 0; i < p; i++) dya[i] = -(1. - sigma) * ry[i];
        for (int i = 0; i < m; i++) t3[i] = -(1. - sigma) * rz[i] - t2[i];
        kkt_solve(&S, dxa, dya, t3, dx, dy, dz, 0);
        apply_W(&S, dz, t2, 0);                              /* dz~ */
        for (int i = 0; i < m; i++) t1[i] = t1[i] - t2[i];   /* ds~ = lambda\d_s - W dz */
        ts = max_step_t(&S, t1); tz = max_step_t(&S, t2);
        tt = fmax(ts, tz);
        double alpha = tt <= STEP_FRAC ? 1. : STEP_FRAC / tt;
        if (getenv("ORC_DEBUG")) fprintf(stderr, "it %2d pres %.2e dres %.2e gap %.2e relgap %.2e aff %.3f sig %.2e alpha %.4f kkt %.1e\n", it, pres, dres, gap, relgap, alpha_aff, sigma, alpha, S.kkt_resid);
        /* additive update (keeps the primal residual linear); back off if rounding leaves the cone */
        apply_W(&S, t1, ds, 0); /* ds = W ds~ */
        for (int tries = 0; tries < 20; tries++) {
            for (int i = 0; i < m; i++) { t1[i] = s[i] + alpha * ds[i]; t2[i] = z[i] + alpha * dz[i]; }
            if (cone_min_margin(&S, t1) > 0 && cone_min_margin(&S, t2) > 0) break;
            alpha *= 0.8;
        }
        for (int j = 0; j < n; j++) x[j] += alpha * dx[j];
        for (int i = 0; i < p; i++) y[i] += alpha * dy[i];
        memcpy(s, t1, sizeof(double) * m); memcpy(z, t2, sizeof(double) * m);
    }
    if (status != 0 && have_best) { /* fall back to the best iterate seen */
        memcpy(x, bx, sizeof(double) * n); memcpy(y, by, sizeof(double) * p); memcpy(s, bs, sizeof(double) * m); memcpy(z, bz, sizeof(double) * m);
        pres = b_pres; dres = b_dres; gap = b_gap; relgap = b_relgap; pcost = b_pcost; dcost = b_dcost;
        if (best_score <= 10.) status = 0;          /* within 10x of the requested tolerances (<= ECOS defaults 1e-8) */
        else if (best_score <= 1e4) status = 3;     /* reduced accuracy (ECOS "close to optimal") */
    }
    free(bx); free(by); free(bs); free(bz);
    if (info) {
        info->status = status; info->iterations = it; info->pres = pres; info->dres = dres; info->gap = gap; info->relgap = relgap;
        info->pcost = pcost; info->dcost = dcost; info->kkt_resid = S.kkt_resid;
        double mg = fmin(cone_min_margin(&S, s), cone_min_margin(&S, z));
        info->cone_viol = mg < 0 ? -mg : 0.;
    }
    free(S.coff); free(S.Ap); free(S.Aj); free(S.Av); free(S.Gp); free(S.Gj); free(S.Gv);
    free(S.exp_rows); free(S.exp_of_row); free(S.pos); free(S.first); free(S.rowptr); free(S.L); free(S.D); free(S.Kenv); free(S.sign);
    free(S.wl); free(S.eta); free(S.wbar); free(S.lambda);
    free(S.w_rhs); free(S.w_sol); free(S.w_wrk); free(S.w_res); free(S.w_t1); free(S.w_t2);
    free(S.w_ex); free(S.w_ey); free(S.w_ez); free(S.w_cx); free(S.w_cy); free(S.w_cz); free(S.w_w1); free(S.w_w2);
    free(rx); free(ry); free(rz); free(dx); free(dy); free(dz); free(ds); free(dxa); free(dya); free(dza); free(dsa);
    free(bb); free(t1); free(t2); free(t3); free(zero_n); free(zero_p); free(zero_m); free(negc);
    return status;
}

int orc_conic_solve(int n, int p, int m, int l, int ncones, const int *q,
                    const double *c, const double *b, const double *h,
                    int nnzA, const int *Ai, const int *Aj, const double *Av,
                    int nnzG, const int *Gi, const int *Gj, const double *Gv,
                    double *x, double *y, double *s, double *z, orc_ipm_info *info)
{
    return orc_conic_solve_keys(n, p, m, l, ncones, q, c, b, h, nnzA, Ai, Aj, Av, nnzG, Gi, Gj, Gv, NULL, NULL, x, y, s, z, info);
}
