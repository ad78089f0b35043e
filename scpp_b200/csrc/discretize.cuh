// scpp_b200/csrc/discretize.cuh — hot path 1: multiple-shooting discretisation (kernel K1 body).
//
// Replaces discretization::multipleShooting (scpp_core/include/discretizationImplementation.hpp:122-181).
// The reference integrates the 14x25 augmented system [x | Phi | Phi^-1 B a | Phi^-1 B b | Phi^-1 f | Phi^-1(-Ax-Bu)]
// with Boost RKF78 (5 fixed steps) and inverts Phi in every RHS evaluation (:65).  Here the mathematically
// identical FORWARD sensitivity form is integrated:  with P = Phi * (bar quantity)
//      x'  = sigma f(x,u(tau))
//      P'  = sigma A(x,u) P + F(tau),   F = 0 | sigma B alpha | sigma B beta | f | sigma(-A x - B u)
// so that at tau = d_tau the columns are directly A_k | B_k | C_k | s_k | z_k (no inverse, no final multiply).
// One THREAD integrates one COLUMN of one interval of one instance (plus its own copy of x, which is free in
// SIMT terms), so there is no communication at all; classical RK4 with NSUB sub-steps (north_star names RK4;
// NSUB is chosen against the RKF78 oracle).
//
// Output layout per interval, row-major [NX][NC], NC = NX + 2 NU + 2:  row r = [A(r,:) | B(r,:) | C(r,:) | s(r) | z(r)]
// => the NC threads of one interval store to consecutive addresses (coalesced).
// K2's stage-parallel passes read the tiles with one lane per stage; for them the tile is stored a second time
// stage-minor ([NX*NC][KS], KS = K rounded up to 4) so that 32 lanes load 32 consecutive doubles.
#pragma once
#include "models.cuh"

namespace scpp {

template <class M>
SCPP_HD void discretize_rhs(const double *x, const double *col, const double *u, const double *du_dtau, const double *par, double sigma,
                            int ctype, int cidx, double alpha, double beta, double *dx, double *dcol)
{
    constexpr int NX = M::NX, NU = M::NU;
    typename M::Lin L;
    M::linearize(x, u, par, L);
    (void)du_dtau;
    double Av[NX];
    M::A_apply(L, col, Av);
#pragma unroll
    for (int i = 0; i < NX; i++) { dx[i] = sigma * L.f[i]; dcol[i] = sigma * Av[i]; }
    if (ctype == 1 || ctype == 2) {            // B / C column: sigma * B e_j * (alpha | beta)
        double e[NU], Be[NX];
#pragma unroll
        for (int j = 0; j < NU; j++) e[j] = (j == cidx) ? 1. : 0.;
        M::B_apply(L, e, Be);
        const double w = sigma * (ctype == 1 ? alpha : beta);
#pragma unroll
        for (int i = 0; i < NX; i++) dcol[i] += w * Be[i];
    } else if (ctype == 3) {                   // s column: f
#pragma unroll
        for (int i = 0; i < NX; i++) dcol[i] += L.f[i];
    } else if (ctype == 4) {                   // z column: sigma (-A x - B u)
        double Ax[NX], Bu[NX];
        M::A_apply(L, x, Ax);
        M::B_apply(L, u, Bu);
#pragma unroll
        for (int i = 0; i < NX; i++) dcol[i] -= sigma * (Ax[i] + Bu[i]);
    }
}

// The same right-hand side through forward-mode AD of the model's generic-scalar flow map (cfg.jacobian = 1, the default): every column's
// derivative is sigma J (dx, du) [+ f for the s column] with ONE tangent direction,
//      A column:  (col, 0)      B / C column:  (col, alpha e_j) / (col, beta e_j)      s column:  (col, 0)      z column:  (col - x, -u)
// so one dual-number pass of flow_map gives f and the product; no Jacobian entries, no hand-derived code (what CppAD gives the reference,
// scpp_core/include/systemDynamics.hpp:206-235).
template <class M>
SCPP_HD void discretize_rhs_ad(const double *x, const double *col, const double *u, const double *par, double sigma,
                               int ctype, int cidx, double alpha, double beta, double *dx, double *dcol)
{
    constexpr int NX = M::NX, NU = M::NU;
    Dual xd[NX], ud[NU], fd[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) xd[i] = Dual(x[i], ctype == 4 ? col[i] - x[i] : col[i]);
#pragma unroll
    for (int j = 0; j < NU; j++) {
        double t = 0.;
        if (ctype == 1 && j == cidx) t = alpha;
        if (ctype == 2 && j == cidx) t = beta;
        if (ctype == 4) t = -u[j];
        ud[j] = Dual(u[j], t);
    }
    M::template flow_map<Dual>(xd, ud, par, fd);
#pragma unroll
    for (int i = 0; i < NX; i++) { dx[i] = sigma * fd[i].v; dcol[i] = sigma * fd[i].d + (ctype == 3 ? fd[i].v : 0.); }
}

// X:[K][NX] U:[K][NU] of one instance; writes column `c` of interval `k` into dd_k (row-major [NX][NC])
template <class M, bool AD = false>
SCPP_HD void discretize_column(const double *X, const double *U, double sigma, const double *par, int K, int k, int c,
                               int nsub, int free_time /* bit 0: free final time (unused here), bit 1: zero-order hold */, double *ddk, double *ddT = nullptr, int KS = 0)
{
    constexpr int NX = M::NX, NU = M::NU, NC = NX + 2 * NU + 2;
    int ctype, cidx;
    if (c < NX) { ctype = 0; cidx = c; }
    else if (c < NX + NU) { ctype = 1; cidx = c - NX; }
    else if (c < NX + 2 * NU) { ctype = 2; cidx = c - NX - NU; }
    else if (c == NX + 2 * NU) { ctype = 3; cidx = 0; }
    else { ctype = 4; cidx = 0; }
    double x0[NX], col[NX], u0[NU], du[NU];
#pragma unroll
    for (int i = 0; i < NX; i++) x0[i] = X[NX * k + i];
#pragma unroll
    // zero-order hold (flags bit 1; discretizationImplementation.hpp:41-50,96-101): the input stays u_k over the interval, the B columns
    // integrate the whole input matrix (alpha = 1) and the C columns nothing (beta = 0: C_k == 0 exactly)
    const bool zoh = (free_time & 2) != 0;
    for (int j = 0; j < NU; j++) { u0[j] = U[NU * k + j]; du[j] = zoh ? 0. : U[NU * (k + 1) + j] - u0[j]; }
    const double dtau = 1. / double(K - 1);
    // nsub > 0: classical RK4 with nsub sub-steps.  nsub < 0: RK4 with n = -nsub and with 2n sub-steps, Richardson-extrapolated
    // (y = y_2n + (y_2n - y_n) / 15 removes the h^4 term): 3n sub-steps for an error below RK4 x 5n (tests/test_host.py).
    const int passes = nsub < 0 ? 2 : 1;
    double colc[NX];
#pragma unroll 1
    for (int pass = 0; pass < passes; pass++) {
        const int ns = nsub < 0 ? (pass == 0 ? -nsub : -2 * nsub) : nsub;
        const double h = dtau / ns;
        double x[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) { x[i] = x0[i]; col[i] = (ctype == 0 && i == cidx) ? 1. : 0.; }
#pragma unroll 1
        for (int st = 0; st < ns; st++) {
            const double t0 = st * h;
            double xa[NX], ca[NX], xt[NX], ct[NX], kx[NX], kc[NX], u[NU];
            // classical RK4: stage times t0, t0+h/2, t0+h/2, t0+h ; u(tau) = u0 + tau/dtau (u1-u0)  (FOH, :45)
#pragma unroll
            for (int i = 0; i < NX; i++) { xa[i] = x[i]; ca[i] = col[i]; xt[i] = x[i]; ct[i] = col[i]; }
#pragma unroll 1
            for (int sgi = 0; sgi < 4; sgi++) {
                const double tau = t0 + (sgi == 0 ? 0. : (sgi == 3 ? h : 0.5 * h));
                const double beta = zoh ? 0. : tau / dtau, alpha = zoh ? 1. : (dtau - tau) / dtau;
#pragma unroll
                for (int j = 0; j < NU; j++) u[j] = u0[j] + beta * du[j];
                if (AD) discretize_rhs_ad<M>(xt, ct, u, par, sigma, ctype, cidx, alpha, beta, kx, kc);
                else discretize_rhs<M>(xt, ct, u, du, par, sigma, ctype, cidx, alpha, beta, kx, kc);
                const double wgt = (sgi == 0 || sgi == 3) ? h / 6. : h / 3.;
                const double nxt = (sgi == 2) ? h : 0.5 * h;
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    xa[i] += wgt * kx[i]; ca[i] += wgt * kc[i];
                    xt[i] = x[i] + nxt * kx[i]; ct[i] = col[i] + nxt * kc[i];
                }
            }
#pragma unroll
            for (int i = 0; i < NX; i++) { x[i] = xa[i]; col[i] = ca[i]; }
        }
        if (passes == 2) {
            if (pass == 0) {
#pragma unroll
                for (int i = 0; i < NX; i++) colc[i] = col[i];
            } else {
#pragma unroll
                for (int i = 0; i < NX; i++) col[i] += (col[i] - colc[i]) * (1. / 15.);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NX; i++) ddk[i * NC + c] = col[i];
    if (ddT) {   // the same tile, stage-minor: element (i, c) of interval k at (i*NC + c)*KS + k   (stage-parallel passes of K2)
#pragma unroll
        for (int i = 0; i < NX; i++) ddT[(size_t)(i * NC + c) * KS + k] = col[i];
    }
}

} // namespace scpp
