// scpp_b200/csrc/mpc.cuh — kernel K6 body: the receding-horizon MPC sub-problem of scpp_core/src/MPCProblem.cpp:6-87 for one instance.
//
// Reference path: MPCAlgorithm::initialize (scpp_core/src/MPCAlgorithm.cpp:34-66: operating point, exactLinearDiscretization
// scpp_core/src/discretization.cpp:9-40, buildMPCProblem, Model::addApplicationConstraints, ECOSSolver) and MPCAlgorithm::solve (:94-116).
// The problem is LINEAR TIME-INVARIANT (constant_dynamics, one A, B, z for the whole horizon) and small (Rocket2D MPC.info: K = 7), and a
// Monte-Carlo batch differs only in x_init / x_final.  So the dynamics equalities are eliminated on the host once per engine,
//      x_k = Phi_k x_0 + S_k U + zhat_k ,      U = (u_0 .. u_{K-2}),
// which leaves a dense conic program in  y = (U, error_cost, input_cost)  whose matrix G is THE SAME for every instance; only
// h = hc + Hx x_init + Hf x_final differs.  One THREAD solves one instance with a dense primal-dual interior-point method (Mehrotra
// predictor-corrector, Nesterov-Todd scaling: the method of ipm.cuh without any structure to exploit); G is read through the cache by
// all threads alike, the per-instance vectors live in thread-local memory (interleaved by the hardware: coalesced).
#pragma once
#include "ipm.cuh"

namespace scpp {

// MPC.info (MPCAlgorithm::loadParameters, scpp_core/src/MPCAlgorithm.cpp:17-32)
struct MpcConfig {
    int K;
    int nondimensionalize, constant_dynamics, intermediate_cost_active;
    double time_horizon;
    double state_weights_intermediate[16], state_weights_terminal[16], input_weights[8];
    IpmSettings ipm;
};

// the condensed conic program shared by all instances (built on the host, mpc.cu)
struct MpcProblem {
    int nv, nl, ncones, nr;      // variables, LP rows, second-order cones, rows in total
    const int *cdim;             // [ncones]
    const double *G;             // [nr][nv] row-major
    const double *c;             // [nv]
    const double *hc;            // [nr]
    const double *Hx, *Hf;       // [nr][nx]  h = hc + Hx x_init + Hf x_final
    const double *Phi, *S, *zh;  // states from the solution: x_k = Phi[k] x0 + S[k] U + zh[k];  Phi [K][nx][nx], S [K][nx][nu (K-1)], zh [K][nx]
};

// ---- second-order-cone primitives, any dimension (contiguous arrays) ------------------------------------------------------------------
namespace gsoc {
SCPP_HD double jn2(const double *u, int d) { double n = 0; for (int i = 1; i < d; i++) n += u[i] * u[i]; return u[0] * u[0] - n; }
// Nesterov-Todd scaling of (s, z): wbar (in w), e2i = |z|_J / |s|_J, lambda = W z = W^-1 s
SCPP_HD bool scale(const double *s, const double *z, int d, double *w, double &e2i, double *lm)
{
    const double ss = jn2(s, d), zz = jn2(z, d);
    if (!(ss > 0.) || !(zz > 0.) || !(s[0] > 0.) || !(z[0] > 0.)) return false;
    const double sn = sqrt(ss), zn = sqrt(zz);
    double sz = 0;
    for (int i = 0; i < d; i++) sz += s[i] * z[i];
    const double i2g = 1. / (2. * sqrt((1. + sz / (sn * zn)) / 2.)), isn = i2g / sn, izn = i2g / zn;
    w[0] = s[0] * isn + z[0] * izn;
    double w1z1 = 0;
    for (int i = 1; i < d; i++) { w[i] = s[i] * isn - z[i] * izn; w1z1 += w[i] * z[i]; }
    e2i = zn / sn;
    const double eta = sqrt(sn / zn), f = z[0] + w1z1 / (1. + w[0]);
    lm[0] = eta * (w[0] * z[0] + w1z1);
    for (int i = 1; i < d; i++) lm[i] = eta * (z[i] + f * w[i]);
    return true;
}
SCPP_HD void Mv(const double *w, double e2i, const double *v, int d, double *o)     // o = W^-2 v (o may alias v)
{
    double dot = w[0] * v[0];
    for (int i = 1; i < d; i++) dot -= w[i] * v[i];
    const double v0 = v[0];
    for (int i = 1; i < d; i++) o[i] = e2i * (-2. * dot * w[i] + v[i]);
    o[0] = e2i * (2. * dot * w[0] - v0);
}
SCPP_HD void Wv(const double *w, double e2i, const double *v, int d, double *o, bool inv)    // o = W v | W^-1 v (o may alias v)
{
    const double eta = 1. / sqrt(e2i), sg = inv ? -1. : 1., sc = inv ? 1. / eta : eta;
    double w1v1 = 0;
    for (int i = 1; i < d; i++) w1v1 += w[i] * v[i];
    const double o0 = w[0] * v[0] + sg * w1v1, f = sg * v[0] + w1v1 / (1. + w[0]);
    for (int i = 1; i < d; i++) o[i] = sc * (v[i] + f * w[i]);
    o[0] = sc * o0;
}
SCPP_HD void jdiv(const double *lm, const double *dv, int d, double *o)      // o = lm \ dv (o may alias dv)
{
    double l1d1 = 0;
    for (int i = 1; i < d; i++) l1d1 += lm[i] * dv[i];
    const double x0 = (lm[0] * dv[0] - l1d1) / jn2(lm, d), il0 = 1. / lm[0];
    for (int i = 1; i < d; i++) o[i] = (dv[i] - x0 * lm[i]) * il0;
    o[0] = x0;
}
SCPP_HD double step(const double *lm, const double *dk, int d)       // largest t with lm + t^-1 ... : the scaled step-length bound of the cone
{
    const double ia = 1. / sqrt(jn2(lm, d)), l0 = lm[0] * ia;
    double ld = l0 * dk[0];
    for (int i = 1; i < d; i++) ld -= lm[i] * ia * dk[i];
    const double rho0 = ld * ia, f = (ld + dk[0]) / (l0 + 1.) * ia;
    double n1 = 0;
    for (int i = 1; i < d; i++) { const double r = dk[i] - f * lm[i]; n1 += r * r; }
    return sqrt(n1) * ia - rho0;
}
} // namespace gsoc

// dense conic interior-point solver, one thread per problem:  min c'y  s.t.  h - G y in R+^nl x Q^d1 x ... ;  NVM / NRM / NCM: compile-time maxima
template <int NVM, int NRM, int NCM>
struct DenseConic {
    int nv, nl, ncones, nr;
    const int *cdim;
    const double *G, *c;
    double y[NVM], s[NRM], z[NRM], w[NRM], lam[NRM], rz[NRM], ds[NRM], dz[NRM], cr[NRM], e2[NCM];
    double H[NVM * NVM], g[NVM], dy[NVM];

    SCPP_HD bool chol()      // H (lower) -> L in place
    {
        for (int j = 0; j < nv; j++) {
            double d = H[j * nv + j];
            for (int q = 0; q < j; q++) d -= H[j * nv + q] * H[j * nv + q];
            if (!(d > 0.)) return false;
            const double l = sqrt(d), il = 1. / l;
            H[j * nv + j] = l;
            for (int i = j + 1; i < nv; i++) {
                double v = H[i * nv + j];
                for (int q = 0; q < j; q++) v -= H[i * nv + q] * H[j * nv + q];
                H[i * nv + j] = v * il;
            }
        }
        return true;
    }
    SCPP_HD void chol_solve(double *b) const    // b <- H^-1 b with the factor
    {
        for (int i = 0; i < nv; i++) { double v = b[i]; for (int q = 0; q < i; q++) v -= H[i * nv + q] * b[q]; b[i] = v / H[i * nv + i]; }
        for (int i = nv - 1; i >= 0; i--) { double v = b[i]; for (int q = i + 1; q < nv; q++) v -= H[q * nv + i] * b[q]; b[i] = v / H[i * nv + i]; }
    }
    // v <- W^-2 v over all rows (LP rows: w = z/s; cones: Nesterov-Todd)
    SCPP_HD void apply_M(double *v) const
    {
        for (int r = 0; r < nl; r++) v[r] *= w[r];
        for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) gsoc::Mv(w + o, e2[k], v + o, cdim[k], v + o);
    }
    // H = G' W^-2 G (lower triangle), identity scaling when `ident`
    SCPP_HD void assemble(bool ident)
    {
        for (int e = 0; e < nv * nv; e++) H[e] = 0.;
        for (int r = 0; r < nl; r++) {
            const double wr = ident ? 1. : w[r];
            const double *gr = G + (size_t)r * nv;
            for (int i = 0; i < nv; i++) { const double a = wr * gr[i]; if (a != 0.) for (int j = 0; j <= i; j++) H[i * nv + j] += a * gr[j]; }
        }
        for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) {
            const int d = cdim[k];
            const double e2i = ident ? 1. : e2[k];
            // W^-2 = e2i (2 (J wbar)(J wbar)' - J): rank one in q = G_c' J wbar, plus -e2i G_c' J G_c
            double q[NVM];
            for (int i = 0; i < nv; i++) q[i] = 0.;
            for (int t = 0; t < d; t++) {
                const double wt = ident ? (t == 0 ? 1. : 0.) : (t == 0 ? w[o] : -w[o + t]);      // J wbar
                const double *gr = G + (size_t)(o + t) * nv;
                const double js = t == 0 ? -e2i : e2i;
                for (int i = 0; i < nv; i++) { q[i] += wt * gr[i]; const double a = js * gr[i]; if (a != 0.) for (int j = 0; j <= i; j++) H[i * nv + j] += a * gr[j]; }
            }
            for (int i = 0; i < nv; i++) { const double a = 2. * e2i * q[i]; if (a != 0.) for (int j = 0; j <= i; j++) H[i * nv + j] += a * q[j]; }
        }
    }
    // Newton solve:  G'dz = rxv ,  G dy - W^2 dz = rzv   ->  dy (in this->dy), dz ;  rzv given in dz (overwritten)
    SCPP_HD void newton(const double *rxv)
    {
        // g = rxv + G' W^-2 rzv
        double t[NRM];
        for (int r = 0; r < nr; r++) t[r] = dz[r];
        apply_M(t);
        for (int i = 0; i < nv; i++) g[i] = rxv[i];
        for (int r = 0; r < nr; r++) { const double tr = t[r]; if (tr != 0.) { const double *gr = G + (size_t)r * nv; for (int i = 0; i < nv; i++) g[i] += gr[i] * tr; } }
        for (int i = 0; i < nv; i++) dy[i] = g[i];
        chol_solve(dy);
        // dz = W^-2 (G dy - rzv)
        for (int r = 0; r < nr; r++) { const double *gr = G + (size_t)r * nv; double a = -dz[r]; for (int i = 0; i < nv; i++) a += gr[i] * dy[i]; dz[r] = a; }
        apply_M(dz);
    }
    SCPP_HD double margin(const double *u) const
    {
        double m = 1e300;
        for (int r = 0; r < nl; r++) m = fmin(m, u[r]);
        for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) { double tq = 0; for (int i = 1; i < cdim[k]; i++) tq += u[o + i] * u[o + i]; m = fmin(m, u[o] - sqrt(tq)); }
        return m;
    }
    SCPP_HD void shift(double *u, double a) const
    {
        for (int r = 0; r < nl; r++) u[r] += a;
        for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) u[o] += a;
    }

    SCPP_HD IpmResult solve(const double *h, const IpmSettings &st)
    {
        IpmResult res;
        res.status = 1; res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.; res.point_ok = 1; res.pad_ = 0;
        double hn = 0, cn = 0;
        for (int r = 0; r < nr; r++) hn += h[r] * h[r];
        for (int i = 0; i < nv; i++) cn += c[i] * c[i];
        const double resz0 = fmax(1., sqrt(hn)), resx0 = fmax(1., sqrt(cn));
        const int degree = nl + ncones;
        // ---- starting point (CVXOPT conelp / ECOS): least-squares primal and dual points with W = I
        assemble(true);
        if (!chol()) { res.status = 2; return res; }
        for (int i = 0; i < nv; i++) { double a = 0; for (int r = 0; r < nr; r++) a += G[(size_t)r * nv + i] * h[r]; y[i] = a; }
        chol_solve(y);
        for (int r = 0; r < nr; r++) { const double *gr = G + (size_t)r * nv; double a = h[r]; for (int i = 0; i < nv; i++) a -= gr[i] * y[i]; s[r] = a; }
        { const double mg = margin(s); if (mg <= 1e-8 * resz0) shift(s, 1. - mg); }
        for (int i = 0; i < nv; i++) dy[i] = -c[i];
        chol_solve(dy);                                                     // z = -G (G'G)^-1 c  ... sign: G'z + c = 0  =>  z = G dy
        for (int r = 0; r < nr; r++) { const double *gr = G + (size_t)r * nv; double a = 0; for (int i = 0; i < nv; i++) a += gr[i] * dy[i]; z[r] = a; }
        { const double mg = margin(z); if (mg <= 1e-8 * resz0) shift(z, 1. - mg); }
        double rx[NVM], best = 1e300, ybest[NVM];
        for (int it = 0; it <= st.maxit; it++) {
            // ---- residuals, scaling
            double gap = 0, rz2 = 0, rx2 = 0, pcost = 0, zrz = 0, xrx = 0;
            bool bad = false;
            for (int i = 0; i < nv; i++) { double a = c[i]; for (int r = 0; r < nr; r++) a += G[(size_t)r * nv + i] * z[r]; rx[i] = a; rx2 += a * a; xrx += y[i] * a; pcost += c[i] * y[i]; }
            for (int r = 0; r < nr; r++) {
                const double *gr = G + (size_t)r * nv;
                double a = s[r] - h[r];
                for (int i = 0; i < nv; i++) a += gr[i] * y[i];
                rz[r] = a; rz2 += a * a; zrz += z[r] * a; gap += s[r] * z[r];
            }
            for (int r = 0; r < nl; r++) { if (!(s[r] > 0.) || !(z[r] > 0.)) bad = true; w[r] = z[r] / s[r]; lam[r] = sqrt(s[r] * z[r]); }
            for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) if (!gsoc::scale(s + o, z + o, cdim[k], w + o, e2[k], lam + o)) bad = true;
            const double pres = sqrt(rz2) / resz0, dres = sqrt(rx2) / resx0;
            const double dcost = pcost - gap + zrz - xrx;
            double relgap = 1e300;
            if (pcost < 0.) relgap = gap / -pcost; else if (dcost > 0.) relgap = gap / dcost;
            const double score = fmax(fmax(pres, dres) / st.feastol, fmin(gap / st.abstol, relgap / st.reltol));
            if (!bad && score < best) {
                best = score;
                res.pres = pres; res.dres = dres; res.gap = gap; res.relgap = relgap; res.pcost = pcost; res.iterations = it;
                for (int i = 0; i < nv; i++) ybest[i] = y[i];
            }
            if (!bad && pres <= st.feastol && dres <= st.feastol && (gap <= st.abstol || relgap <= st.reltol)) { res.status = 0; break; }
            if (bad || it == st.maxit || (score > 1e3 * best && best < 1e4) || (best <= 10. && score > best)) { res.status = bad ? 2 : (it == st.maxit ? 1 : 2); break; }
            assemble(false);
            if (!chol()) { res.status = 2; break; }
            // ---- affine direction:  rxv = -rx , rzv = -rz + s
            double rxv[NVM];
            for (int i = 0; i < nv; i++) rxv[i] = -rx[i];
            for (int r = 0; r < nr; r++) dz[r] = -rz[r] + s[r];
            newton(rxv);
            double tmax = 0;
            auto steps = [&](double rzs, bool keep_cr) {      // ds = rzs rz - G dy ; scaled directions, step-length bound, corrector term
                double tm = 0;
                for (int r = 0; r < nr; r++) { const double *gr = G + (size_t)r * nv; double a = rzs * rz[r]; for (int i = 0; i < nv; i++) a -= gr[i] * dy[i]; ds[r] = a; }
                for (int r = 0; r < nl; r++) {
                    const double iw = sqrt(w[r]), dzt = dz[r] / iw, dst = ds[r] * iw;
                    tm = fmax(tm, fmax(-dst, -dzt) / lam[r]);
                    if (keep_cr) cr[r] = dst * dzt;
                }
                for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) {
                    const int d = cdim[k];
                    double a[NRM > 64 ? 64 : NRM], b[NRM > 64 ? 64 : NRM];
                    gsoc::Wv(w + o, e2[k], dz + o, d, a, false);           // dz~ = W dz
                    gsoc::Wv(w + o, e2[k], ds + o, d, b, true);            // ds~ = W^-1 ds
                    tm = fmax(tm, fmax(gsoc::step(lam + o, b, d), gsoc::step(lam + o, a, d)));
                    if (keep_cr) {
                        double dot = 0;
                        for (int i = 0; i < d; i++) dot += a[i] * b[i];
                        for (int i = 1; i < d; i++) cr[o + i] = b[0] * a[i] + a[0] * b[i];
                        cr[o] = dot;
                    }
                }
                return tm;
            };
            tmax = steps(-1., true);
            const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
            const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = gap / degree, csig = 1. - sig, sigmu = sig * mu;
            // ---- combined direction:  rxv = -(1-sig) rx ,  rzv = -(1-sig) rz - W (lam \ d_s) ,  d_s = -lam o lam - cr + sig mu e
            for (int i = 0; i < nv; i++) rxv[i] = -csig * rx[i];
            for (int r = 0; r < nl; r++) dz[r] = -csig * rz[r] - sqrt(1. / w[r]) * ((-lam[r] * lam[r] - cr[r] + sigmu) / lam[r]);
            for (int k = 0, o = nl; k < ncones; o += cdim[k], k++) {
                const int d = cdim[k];
                double t1[NRM > 64 ? 64 : NRM];
                double ll = 0;
                for (int i = 0; i < d; i++) ll += lam[o + i] * lam[o + i];
                t1[0] = -ll - cr[o] + sigmu;
                for (int i = 1; i < d; i++) t1[i] = -2. * lam[o] * lam[o + i] - cr[o + i];
                gsoc::jdiv(lam + o, t1, d, t1);
                gsoc::Wv(w + o, e2[k], t1, d, t1, false);
                for (int i = 0; i < d; i++) dz[o + i] = -csig * rz[o + i] - t1[i];
            }
            newton(rxv);
            tmax = steps(-csig, false);
            const double alpha = tmax <= 0.99 ? 1. : 0.99 / tmax;
            for (int i = 0; i < nv; i++) y[i] += alpha * dy[i];
            for (int r = 0; r < nr; r++) { s[r] += alpha * ds[r]; z[r] += alpha * dz[r]; }
        }
        if (res.status != 0) {
            if (best <= 1e4) { for (int i = 0; i < nv; i++) y[i] = ybest[i]; res.status = 3; }
        }
        return res;
    }
};

// ---- K6: one MPC sub-problem.  h from the instance's x_init / x_final, solve, then X, U of the horizon ----------------------------------
template <class M, int KM>
SCPP_HD void mpc_solve_instance(const MpcProblem &P, int K, const IpmSettings &st, const double *x0, const double *xf, double *X /* [K][NX] */,
                                double *U /* [K-1][NU] */, int *status, int *iters)
{
    constexpr int NX = M::NX, NU = M::NU;
    constexpr int NVM = NU * (KM - 1) + 2, NCM = (M::NCONE + 1) * KM + 2, NRM = (M::NLP + M::NCR) * KM + (1 + NX) + (1 + NU * (KM - 1));
    DenseConic<NVM, NRM, NCM> S;
    S.nv = P.nv; S.nl = P.nl; S.ncones = P.ncones; S.nr = P.nr; S.cdim = P.cdim; S.G = P.G; S.c = P.c;
    double h[NRM];
    for (int r = 0; r < P.nr; r++) {
        double a = P.hc[r];
        for (int i = 0; i < NX; i++) a += P.Hx[r * NX + i] * x0[i] + P.Hf[r * NX + i] * xf[i];
        h[r] = a;
    }
    const IpmResult res = S.solve(h, st);
    *status = res.status; *iters = res.iterations;
    const int nuu = NU * (K - 1);
    for (int k = 0; k < K; k++)
        for (int i = 0; i < NX; i++) {
            double a = P.zh[k * NX + i];
            for (int j = 0; j < NX; j++) a += P.Phi[(k * NX + i) * NX + j] * x0[j];
            for (int j = 0; j < nuu; j++) a += P.S[(size_t)(k * NX + i) * nuu + j] * S.y[j];
            X[k * NX + i] = a;
        }
    for (int j = 0; j < nuu; j++) U[j] = S.y[j];
}

} // namespace scpp
