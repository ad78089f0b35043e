/*
 * oracle/orc_lqr.c — CPU ORACLE (test infrastructure): LQR gains along a trajectory, restated literally.  PARITY UNPINNED (see orc.h).
 *
 * Reference: scpp_core/src/LQR.cpp:7-109 (solveSchurIterative, careSolve, ComputeLQR), scpp_core/src/LQRTracker.cpp:6-28 (one gain per
 * node from the Jacobians at (x_k, u_k)), :41-65 (getInput / interpolateGains), trajectoryData.hpp:41-78 (inputAtTime, approxStateAtTime).
 * Third-party pieces restated: Eigen's fixed-size .inverse() (PartialPivLU), isApprox, FullPivLU::solve of a rectangular system.
 */
#include "orc.h"
#include <math.h>
#include <string.h>

#define NMAX (2 * ORC_MAX_NX)

/* inverse by LU with partial pivoting, solve for the identity (row-major n x n) */
static void inverse_pplu(int n, const double *M, double *Minv)
{
    double a[NMAX * NMAX], b[NMAX * NMAX];
    memcpy(a, M, sizeof(double) * n * n);
    for (int i = 0; i < n * n; i++) b[i] = 0.;
    for (int i = 0; i < n; i++) b[i * n + i] = 1.;
    for (int c = 0; c < n; c++) {
        int piv = c; double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; r++) if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
        if (piv != c) for (int j = 0; j < n; j++) {
            double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
            t = b[c * n + j]; b[c * n + j] = b[piv * n + j]; b[piv * n + j] = t;
        }
        const double d = 1. / a[c * n + c];
        for (int r = c + 1; r < n; r++) {
            const double l = a[r * n + c] * d;
            if (l == 0.) continue;
            for (int j = c; j < n; j++) a[r * n + j] -= l * a[c * n + j];
            for (int j = 0; j < n; j++) b[r * n + j] -= l * b[c * n + j];
        }
    }
    for (int j = 0; j < n; j++)
        for (int r = n - 1; r >= 0; r--) {
            double acc = b[r * n + j];
            for (int c = r + 1; c < n; c++) acc -= a[r * n + c] * Minv[c * n + j];
            Minv[r * n + j] = acc / a[r * n + r];
        }
}

/* x = FullPivLU(U).solve(rhs) for U (m x n, m >= n, row-major), rhs (m x nr): complete pivoting; the solution is read off the n pivot rows
 * (Eigen: forward substitution with the unit-lower factor, back substitution on the leading nonzero-pivot block, undo the column
 * permutation, zeros elsewhere) */
static void fullpivlu_solve(int m, int n, const double *U, int nr, const double *rhs, double *x)
{
    double a[NMAX * ORC_MAX_NX], c[NMAX * ORC_MAX_NX];
    int colperm[ORC_MAX_NX];
    memcpy(a, U, sizeof(double) * m * n); memcpy(c, rhs, sizeof(double) * m * nr);
    for (int j = 0; j < n; j++) colperm[j] = j;
    double maxpivot = 0.;
    int nonzero = n;
    for (int k = 0; k < n; k++) {
        int pr = k, pc = k; double best = -1.;
        for (int i = k; i < m; i++) for (int j = k; j < n; j++) if (fabs(a[i * n + j]) > best) { best = fabs(a[i * n + j]); pr = i; pc = j; }
        if (best == 0.) { nonzero = k; break; }
        if (best > maxpivot) maxpivot = best;
        if (pr != k) { for (int j = 0; j < n; j++) { double t = a[k * n + j]; a[k * n + j] = a[pr * n + j]; a[pr * n + j] = t; }
                       for (int j = 0; j < nr; j++) { double t = c[k * nr + j]; c[k * nr + j] = c[pr * nr + j]; c[pr * nr + j] = t; } }
        if (pc != k) { for (int i = 0; i < m; i++) { double t = a[i * n + k]; a[i * n + k] = a[i * n + pc]; a[i * n + pc] = t; }
                       int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t; }
        for (int i = k + 1; i < m; i++) {
            const double l = a[i * n + k] / a[k * n + k];
            a[i * n + k] = l;
            for (int j = k + 1; j < n; j++) a[i * n + j] -= l * a[k * n + j];
            for (int j = 0; j < nr; j++) c[i * nr + j] -= l * c[k * nr + j];     /* forward substitution of the right-hand side */
        }
    }
    /* rank: pivots above epsilon * max(m, n) * |largest pivot| (Eigen's default threshold) */
    int rank = 0;
    for (int k = 0; k < nonzero; k++) if (fabs(a[k * n + k]) > maxpivot * 2.220446049250313e-16 * (m > n ? m : n)) rank++;
    for (int j = 0; j < nr; j++) {
        double y[ORC_MAX_NX];
        for (int r = rank - 1; r >= 0; r--) {
            double acc = c[r * nr + j];
            for (int q = r + 1; q < rank; q++) acc -= a[r * n + q] * y[q];
            y[r] = acc / a[r * n + r];
        }
        for (int i = 0; i < n; i++) x[i * nr + j] = 0.;
        for (int r = 0; r < rank; r++) x[colperm[r] * nr + j] = y[r];
    }
}

/* solveSchurIterative, LQR.cpp:7-54: matrix sign function of the Hamiltonian by Newton's iteration, then P from [M12; M22 + I] P = -[M11 + I; M21] */
static int solve_schur_iterative(int n, const double *M, double *P, double epsilon, int max_iterations)
{
    const int N2 = 2 * n;
    double Ml[NMAX * NMAX], Minv[NMAX * NMAX], Mnew[NMAX * NMAX];
    memset(Mnew, 0, sizeof(Mnew));
    memcpy(Ml, M, sizeof(double) * N2 * N2);
    int iterations = 0, converged = 0;
    while (!converged) {
        if (iterations > max_iterations) return 0;                         /* :19-20 */
        inverse_pplu(N2, Ml, Minv);
        double d2 = 0., a2 = 0., b2 = 0.;
        for (int i = 0; i < N2 * N2; i++) {
            const double mdiff = Ml[i] - Minv[i];                           /* :22 */
            Mnew[i] = Ml[i] - 0.5 * mdiff;                                  /* :24 */
            const double d = Mnew[i] - Ml[i];
            d2 += d * d; a2 += Mnew[i] * Mnew[i]; b2 += Ml[i] * Ml[i];
        }
        converged = d2 <= epsilon * epsilon * (a2 < b2 ? a2 : b2);         /* isApprox  :26 */
        memcpy(Ml, Mnew, sizeof(double) * N2 * N2);
        iterations++;
    }
    double U[NMAX * ORC_MAX_NX], V[NMAX * ORC_MAX_NX];
    memset(U, 0, sizeof(U)); memset(V, 0, sizeof(V));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            U[i * n + j] = Ml[i * N2 + n + j];                              /* M12 */
            U[(n + i) * n + j] = Ml[(n + i) * N2 + n + j] + (i == j);       /* M22 + I */
            V[i * n + j] = -(Ml[i * N2 + j] + (i == j));                    /* -(M11 + I) */
            V[(n + i) * n + j] = -Ml[(n + i) * N2 + j];                     /* -M21 */
        }
    fullpivlu_solve(N2, n, U, n, V, P);                                     /* :49-51 */
    return 1;
}

/* ComputeLQR (LQR.cpp:80-109) with careSolve (:56-78); Q, R diagonal (LQRTracker.cpp:30-40).  A (nx x nx), B (nx x nu) row-major.
 * K (nu x nx) row-major.  Returns the success flag of careSolve. */
int orc_lqr_gain(int nx, int nu, const double *q_diag, const double *r_diag, const double *A, const double *B, double *K)
{
    const int N2 = 2 * nx;
    double M[NMAX * NMAX], P[ORC_MAX_NX * ORC_MAX_NX];
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < nx; j++) {
            double brb = 0.;
            for (int l = 0; l < nu; l++) brb += B[i * nu + l] * (1. / r_diag[l]) * B[j * nu + l];
            M[i * N2 + j] = A[i * nx + j];
            M[i * N2 + nx + j] = -brb;
            M[(nx + i) * N2 + j] = (i == j) ? -q_diag[i] : 0.;
            M[(nx + i) * N2 + nx + j] = -A[j * nx + i];
        }
    const int ok = solve_schur_iterative(nx, M, P, 1e-8, 100);
    if (!ok) { for (int i = 0; i < nu * nx; i++) K[i] = NAN; return 0; }
    for (int l = 0; l < nu; l++)
        for (int j = 0; j < nx; j++) {
            double acc = 0.;
            for (int i = 0; i < nx; i++) acc += B[i * nu + l] * P[i * nx + j];
            K[l * nx + j] = acc / r_diag[l];                                /* :103 */
        }
    return 1;
}

/* LQRTracker::LQRTracker, LQRTracker.cpp:6-28: one gain per node, Jacobians at (X_k, U_k) (first-order hold: U has K nodes).
 * gains [K][nu][nx], ok [K]. */
void orc_lqr_tracker_gains(int model, int K, const double *X, const double *U, const double *par,
                           const double *q_diag, const double *r_diag, double *gains, int *ok)
{
    int nx, nu, np;
    orc_model_dims(model, &nx, &nu, &np);
    for (int k = 0; k < K; k++) {
        double Ac[ORC_MAX_NX * ORC_MAX_NX], Bc[ORC_MAX_NX * ORC_MAX_NU], A[ORC_MAX_NX * ORC_MAX_NX], B[ORC_MAX_NX * ORC_MAX_NU];
        orc_jac(model, X + nx * k, U + nu * k, par, Ac, Bc);                /* column-major */
        for (int i = 0; i < nx; i++) { for (int j = 0; j < nx; j++) A[i * nx + j] = Ac[i + nx * j]; for (int j = 0; j < nu; j++) B[i * nu + j] = Bc[i + nx * j]; }
        ok[k] = orc_lqr_gain(nx, nu, q_diag, r_diag, A, B, gains + (size_t)k * nu * nx);
    }
}
