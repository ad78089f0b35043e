// scpp_b200/csrc/ipm.cuh — hot path 2: the convex sub-problem of one SC iteration (kernel K2 body).
//
// Replaces  buildSCProblem (scpp_core/src/SCProblem.cpp:6-138) + Model::addApplicationConstraints
// (scpp_models/src/rocketQuat.cpp:70-144, rocket2d.cpp:46-84) + cvx::ecos::ECOSSolver::solve
// (call site scpp_core/src/SCAlgorithm.cpp:78) by a structure-exploiting primal-dual interior-point method
// (Mehrotra predictor-corrector, Nesterov-Todd scaling — the algorithm family of ECOS) that one WARP runs
// for one problem instance.
//
// Reformulation (exact, same optimal X,U,sigma):
//   * nu_k is eliminated through the dynamics rows:  nu_k = r_k(y) = x_{k+1} - A_k x_k - B_k u_k - C_k u_{k+1} - s_k sigma - z_k
//   * w_vc * norm1_nu with  -nu_bound <= nu <= nu_bound, sum(nu_bound) <= norm1_nu   becomes   w_vc * sum t_ki,  t_ki >= |r_ki|
//   * variables pinned by single-variable equalities (x_0 = x_init, final-state rows, ...) are removed (masked)
// leaving an inequality-only conic program  min c'x  s.t.  h - G x in K  over  y = (xi_0..xi_{K-1}, sigma)  and the
// "local" epigraph variables delta_k, t_ki, delta_sigma, each of which is eliminated analytically from the Newton
// system.  What remains is  H dy = g  with H symmetric positive definite, BLOCK-TRIDIAGONAL in the nodes
// (NB x NB blocks, NB = nx + nu) plus one dense border row/column for sigma; it is factored by a block Cholesky
// sweep over the K nodes.
//
// Work split inside the warp: cone arithmetic is done lane-per-node / lane-per-row; everything touching the
// discretisation tensors [A|B|C|s|z]_k and the NB x NB blocks is warp-cooperative on tiles staged in shared memory.
#pragma once
#include "models.cuh"
#if !defined(__CUDACC__)
#include <cstdio>
#include <cstdlib>
#endif

namespace scpp {

struct IpmSettings {
    double feastol, abstol, reltol;
    int maxit;
};

struct IpmResult {
    int status;      // 0 optimal, 1 max iterations, 2 numerical failure, 3 reduced accuracy
    int iterations;
    double pres, dres, gap, relgap, pcost;
};

template <class M>
struct Ipm {
    static constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2, NCP = NC + 1;
    static constexpr int NLP = M::NLP, NCONE = M::NCONE, NCR = M::NCR;
    static constexpr int MN = NLP + NCR + 1 + NB;   // cone rows per node: model LP | model cones | trust region
    static constexpr int PN = NB + 1;               // primal per node: xi, delta
    static constexpr int NCN = NCONE + 1;           // second-order cones per node (model + trust region)
    static constexpr int NRK = NCONE + 1;           // rank-1 terms of the model Hessian: cones + one multi-entry LP row
    static constexpr int TRO = NLP + NCR;           // offset of the trust-region cone inside a node block
    static constexpr int BLK = NB * NB;
    static constexpr int FACK = 2 * BLK + NB;       // per node: Linv_kk | L_{k+1,k} | border l_k

    SCPP_HD static int m_rows(int K) { return K * MN + (K - 1) * 2 * NX + 4; }
    SCPP_HD static int n_prim(int K) { return K * PN + 2 + (K - 1) * NX; }
    SCPP_HD static int n_cones(int K) { return K * NCN + 1; }
    SCPP_HD static int n_y(int K) { return K * NB + 1; }
    SCPP_HD static int ws_doubles(int K) { return 4 * n_prim(K) + 8 * m_rows(K) + n_cones(K) + K * FACK + n_y(K) + 16; }
    SCPP_HD static int sm_doubles() { return NX * NCP + 5 * BLK + 2 * NRK * NB + 12 * NB + 4 * NX + 32; }

    // ---- problem data (read only) -----------------------------------------------------------------------------
    int K;
    const double *dd;      // [K-1][NX][NC]
    const double *Xbar;    // [K][NX]
    const double *Ubar;    // [K][NU]
    double sigbar;
    const double *cst;     // per-instance constants of the row table
    const double *tdir;    // [K][3]
    const uint32_t *fixm;  // [K]
    const double *fixv;    // [K][NB]
    double w_time, w_trs, w_tr, w_vc;
    // ---- workspace (global memory, per instance) --------------------------------------------------------------
    double *prim, *dprim, *rx;
    double *s, *z, *wb, *lam, *rz, *cr, *dz, *ds;
    double *ce;            // eta^-2 per second-order cone
    double *fac, *gy;
    double *best_;         // best primal iterate seen
    double *sm;            // per-warp shared scratch

    SCPP_HD void bind(double *ws, double *smem)
    {
        const int np = n_prim(K), m = m_rows(K);
        double *p = ws;
        prim = p; p += np; dprim = p; p += np; rx = p; p += np; best_ = p; p += np;
        s = p; p += m; z = p; p += m; wb = p; p += m; lam = p; p += m; rz = p; p += m; cr = p; p += m; dz = p; p += m; ds = p; p += m;
        ce = p; p += n_cones(K);
        fac = p; p += K * FACK;
        gy = p; p += n_y(K);
        sm = smem;
    }
    // index helpers
    SCPP_HD int pn(int k) const { return k * PN; }
    SCPP_HD int p_sigma() const { return K * PN; }
    SCPP_HD int p_dsig() const { return K * PN + 1; }
    SCPP_HD int p_t(int k) const { return K * PN + 2 + k * NX; }
    SCPP_HD int rn(int k) const { return k * MN; }
    SCPP_HD int ri(int k) const { return K * MN + k * 2 * NX; }
    SCPP_HD int rg() const { return K * MN + (K - 1) * 2 * NX; }
    SCPP_HD bool fixed(int k, int i) const { return (fixm[k] >> i) & 1u; }
    SCPP_HD double xibar(int k, int i) const { return i < NX ? Xbar[k * NX + i] : Ubar[k * NU + (i - NX)]; }

    // shared scratch layout
    SCPP_HD double *sm_dd() const { return sm; }
    SCPP_HD double *sm_H() const { return sm + NX * NCP; }
    SCPP_HD double *sm_O() const { return sm_H() + BLK; }
    SCPP_HD double *sm_Lp() const { return sm_O() + BLK; }
    SCPP_HD double *sm_Li() const { return sm_Lp() + BLK; }
    SCPP_HD double *sm_Hn() const { return sm_Li() + BLK; }
    SCPP_HD double *sm_rk() const { return sm_Hn() + BLK; }            // [NRK][NB] rank-1 vectors, then [NRK][NB] diagonals
    SCPP_HD double *sm_v(int i) const { return sm_rk() + 2 * NRK * NB + i * NB; }   // 12 NB-vectors
    SCPP_HD double *sm_x(int i) const { return sm_v(12) + i * NX; }                 // 4 NX-vectors
    SCPP_HD double *sm_sc() const { return sm_x(4); }                               // 32 scalars

    // ---- second-order-cone primitives (one lane, one cone) ------------------------------------------------------
    SCPP_HD static double jn2(const double *u, int d) { double n = 0; for (int i = 1; i < d; i++) n += u[i] * u[i]; return u[0] * u[0] - n; }
    SCPP_HD static bool soc_scale(const double *sk, const double *zk, int d, double *w, double &e2i, double *lm)
    {
        double ss = jn2(sk, d), zz = jn2(zk, d);
        if (!(ss > 0.) || !(zz > 0.) || !(sk[0] > 0.) || !(zk[0] > 0.)) return false;
        double sn = sqrt(ss), zn = sqrt(zz), sz = 0;
        for (int i = 0; i < d; i++) sz += sk[i] * zk[i];
        double gam = sqrt((1. + sz / (sn * zn)) / 2.);
        double i2g = 1. / (2. * gam);
        w[0] = (sk[0] / sn + zk[0] / zn) * i2g;
        for (int i = 1; i < d; i++) w[i] = (sk[i] / sn - zk[i] / zn) * i2g;
        e2i = zn / sn;
        double eta = sqrt(sn / zn), w1z1 = 0;
        for (int i = 1; i < d; i++) w1z1 += w[i] * zk[i];
        double f = zk[0] + w1z1 / (1. + w[0]);
        lm[0] = eta * (w[0] * zk[0] + w1z1);
        for (int i = 1; i < d; i++) lm[i] = eta * (zk[i] + f * w[i]);
        return true;
    }
    SCPP_HD static void soc_M(const double *w, double e2i, const double *v, int d, double *o)   // o = W^-2 v
    {
        double dot = w[0] * v[0];
        for (int i = 1; i < d; i++) dot -= w[i] * v[i];
        o[0] = e2i * (2. * dot * w[0] - v[0]);
        for (int i = 1; i < d; i++) o[i] = e2i * (-2. * dot * w[i] + v[i]);
    }
    SCPP_HD static void soc_W(const double *w, double e2i, const double *v, int d, double *o, bool inv) // o = W v | W^-1 v
    {
        const double eta = 1. / sqrt(e2i);
        const double sg = inv ? -1. : 1., sc = inv ? 1. / eta : eta;
        double w1v1 = 0;
        for (int i = 1; i < d; i++) w1v1 += w[i] * v[i];
        const double o0 = w[0] * v[0] + sg * w1v1, f = sg * v[0] + w1v1 / (1. + w[0]);
        for (int i = 1; i < d; i++) o[i] = sc * (v[i] + f * w[i]);
        o[0] = sc * o0;
    }
    SCPP_HD static void soc_jprod(const double *u, const double *v, int d, double *o)
    {
        double dot = 0;
        for (int i = 0; i < d; i++) dot += u[i] * v[i];
        const double u0 = u[0], v0 = v[0];
        for (int i = 1; i < d; i++) o[i] = u0 * v[i] + v0 * u[i];
        o[0] = dot;
    }
    SCPP_HD static void soc_jdiv(const double *lm, const double *dv, int d, double *o)
    {
        double den = jn2(lm, d), l1d1 = 0;
        for (int i = 1; i < d; i++) l1d1 += lm[i] * dv[i];
        const double x0 = (lm[0] * dv[0] - l1d1) / den;
        for (int i = 1; i < d; i++) o[i] = (dv[i] - x0 * lm[i]) / lm[0];
        o[0] = x0;
    }
    SCPP_HD static double soc_step(const double *lm, const double *dk, int d)
    {
        const double a = sqrt(jn2(lm, d)), l0 = lm[0] / a;
        double ld = l0 * dk[0];
        for (int i = 1; i < d; i++) ld -= lm[i] / a * dk[i];
        const double rho0 = ld / a, f = (ld + dk[0]) / (l0 + 1.);
        double n1 = 0;
        for (int i = 1; i < d; i++) { double r = (dk[i] - f * lm[i] / a) / a; n1 += r * r; }
        return sqrt(n1) - rho0;
    }

    // row r of the node table evaluated at xi:  returns h - sum coef xi
    SCPP_HD double row_slack(int r, int k, const double *xi) const
    {
        const RowDesc rd = M::row(r);
        double v = cst[rd.hs];
        for (int j = 0; j < rd.n; j++) v -= coef(rd, j, k) * xi[rd.idx[j]];
        return v;
    }
    SCPP_HD double coef(const RowDesc &rd, int j, int k) const { return rd.cs[j] >= 0 ? cst[rd.cs[j]] : -tdir[3 * k + (-rd.cs[j] - 1)]; }

    // stage the discretisation tile of interval k into shared memory (row stride padded to NCP)
    SCPP_HD void stage_dd(int k) const
    {
        const double *src = dd + (size_t)k * NX * NC;
        double *t = sm_dd();
        FOR_LANE(e, NX * NC) { int r = e / NC, c = e - r * NC; t[r * NCP + c] = src[e]; }
        warp_sync();
    }
    // r_i = x_{k+1,i} - (A xi_k)_i - (B u_k)_i - (C u_{k+1})_i - s_i sigma - z_i  for lane-owned i (tile staged), y from `p`
    SCPP_HD double dyn_resid(int k, int i, const double *p, bool with_const) const
    {
        const double *t = sm_dd() + i * NCP;
        const double *xk = p + pn(k), *xn = p + pn(k + 1);
        double acc = xn[i];
        for (int j = 0; j < NB; j++) acc -= t[j] * xk[j];
        for (int j = 0; j < NU; j++) acc -= t[NB + j] * xn[NX + j];
        acc -= t[NB + NU] * p[p_sigma()];
        if (with_const) acc -= t[NB + NU + 1];
        return acc;
    }
    // out (indexed like y) += J' w  for interval k, w in sm_x(0)[NX] ; tile staged.  out_k = node k slice, etc.
    SCPP_HD void dyn_JT(int k, const double *w, double *out_k, double *out_n, double &acc_sigma) const
    {
        const double *t = sm_dd();
        FOR_LANE(j, NB) { double a = 0; for (int i = 0; i < NX; i++) a += t[i * NCP + j] * w[i]; out_k[j] -= a; }
        FOR_LANE(j, NB) {
            if (j < NX) out_n[j] += w[j];
            else { double a = 0; for (int i = 0; i < NX; i++) a += t[i * NCP + NB + (j - NX)] * w[i]; out_n[j] -= a; }
        }
        FOR_LANE(i, NX) acc_sigma -= t[i * NCP + NB + NU] * w[i];
    }

    // =============================================================================================================
    //  Phase R : residuals, Nesterov-Todd scaling, termination quantities
    // =============================================================================================================
    struct Norms { double gap, rz2, rx2, pcost, zrz, xrx, h2; int bad; };

    SCPP_HD void phase_residuals(Norms &nm, bool identity)
    {
        double gap = 0, rz2 = 0, pcost = 0, zrz = 0, h2 = 0;
        int bad = 0;
        // ---- nodes: one lane per node
        for (int k = lane_id(); k < K; k += LANES) {
            const double *xi = prim + pn(k);
            const double dl = xi[NB];
            double *rxk = rx + pn(k);
            for (int i = 0; i < NB; i++) rxk[i] = 0.;
            const int r0 = rn(k);
            // model rows
            for (int r = 0; r < NLP + NCR; r++) {
                const RowDesc rd = M::row(r);
                double sl = cst[rd.hs];
                for (int j = 0; j < rd.n; j++) { const double c = coef(rd, j, k); sl -= c * xi[rd.idx[j]]; rxk[rd.idx[j]] += c * z[r0 + r]; }
                rz[r0 + r] = s[r0 + r] - sl;
                h2 += cst[rd.hs] * cst[rd.hs];
            }
            // trust region: s = (delta ; xibar - xi)
            rz[r0 + TRO] = s[r0 + TRO] - dl;
            for (int i = 0; i < NB; i++) {
                const double xb = xibar(k, i);
                rz[r0 + TRO + 1 + i] = s[r0 + TRO + 1 + i] - (xb - xi[i]);
                rxk[i] += z[r0 + TRO + 1 + i];
                h2 += xb * xb;
            }
            rxk[NB] = w_tr - z[r0 + TRO];
            pcost += w_tr * dl;
            // scaling
            for (int r = 0; r < NLP; r++) {
                const double sv = s[r0 + r], zv = z[r0 + r];
                if (!(sv > 0.) || !(zv > 0.)) bad = 1;
                wb[r0 + r] = identity ? 1. : zv / sv;
                lam[r0 + r] = identity ? 1. : sqrt(sv * zv);
            }
            for (int c = 0; c < NCN; c++) {
                const int o = r0 + NLP + (c < NCONE ? M::cone_off(c) : NCR), d = c < NCONE ? M::cone_dim(c) : 1 + NB;
                if (identity) { ce[k * NCN + c] = 1.; for (int i = 0; i < d; i++) { wb[o + i] = i == 0; lam[o + i] = i == 0; } }
                else if (!soc_scale(s + o, z + o, d, wb + o, ce[k * NCN + c], lam + o)) bad = 1;
            }
            for (int r = 0; r < MN; r++) { gap += s[r0 + r] * z[r0 + r]; rz2 += rz[r0 + r] * rz[r0 + r]; zrz += z[r0 + r] * rz[r0 + r]; }
        }
        warp_sync();
        // ---- intervals: warp-cooperative, tile staged
        double acc_sig = 0;
        for (int k = 0; k < K - 1; k++) {
            stage_dd(k);
            const int r0 = ri(k);
            double *w = sm_x(0);
            FOR_LANE(i, NX) {
                const double r = dyn_resid(k, i, prim, true);
                const double t = prim[p_t(k) + i];
                const double sm_ = s[r0 + i], sp = s[r0 + NX + i], zm = z[r0 + i], zp = z[r0 + NX + i];
                const double rm = sm_ - (t - r), rp = sp - (t + r);
                rz[r0 + i] = rm; rz[r0 + NX + i] = rp;
                if (!(sm_ > 0.) || !(sp > 0.) || !(zm > 0.) || !(zp > 0.)) bad = 1;
                wb[r0 + i] = identity ? 1. : zm / sm_; wb[r0 + NX + i] = identity ? 1. : zp / sp;
                lam[r0 + i] = identity ? 1. : sqrt(sm_ * zm); lam[r0 + NX + i] = identity ? 1. : sqrt(sp * zp);
                rx[p_t(k) + i] = w_vc - zm - zp;
                w[i] = zm - zp;
                gap += sm_ * zm + sp * zp; rz2 += rm * rm + rp * rp; zrz += zm * rm + zp * rp;
                pcost += w_vc * t;
                const double zc = sm_dd()[i * NCP + NB + NU + 1];
                h2 += 2. * zc * zc;
            }
            warp_sync();
            dyn_JT(k, w, rx + pn(k), rx + pn(k + 1), acc_sig);
            warp_sync();
        }
        acc_sig = warp_sum(acc_sig);
        // ---- globals: lane 0
        if (lane_id() == 0) {
            const int r0 = rg();
            const double sg = prim[p_sigma()], dsg = prim[p_dsig()];
            double rxs = w_time + acc_sig;
            // sigma >= 0.001
            rz[r0] = s[r0] - (sg - 0.001);
            rxs -= z[r0];
            if (!(s[r0] > 0.) || !(z[r0] > 0.)) bad = 1;
            wb[r0] = identity ? 1. : z[r0] / s[r0]; lam[r0] = identity ? 1. : sqrt(s[r0] * z[r0]);
            // ((1+dsg)/2 ; (1-dsg)/2 ; sigma - sigbar)
            rz[r0 + 1] = s[r0 + 1] - (0.5 + 0.5 * dsg);
            rz[r0 + 2] = s[r0 + 2] - (0.5 - 0.5 * dsg);
            rz[r0 + 3] = s[r0 + 3] - (sg - sigbar);
            rxs -= z[r0 + 3];
            rx[p_dsig()] = w_trs - 0.5 * z[r0 + 1] + 0.5 * z[r0 + 2];
            rx[p_sigma()] = rxs;
            if (identity) { ce[K * NCN] = 1.; for (int i = 0; i < 3; i++) { wb[r0 + 1 + i] = i == 0; lam[r0 + 1 + i] = i == 0; } }
            else if (!soc_scale(s + r0 + 1, z + r0 + 1, 3, wb + r0 + 1, ce[K * NCN], lam + r0 + 1)) bad = 1;
            for (int r = 0; r < 4; r++) { gap += s[r0 + r] * z[r0 + r]; rz2 += rz[r0 + r] * rz[r0 + r]; zrz += z[r0 + r] * rz[r0 + r]; }
            pcost += w_time * sg + w_trs * dsg;
            h2 += 0.001 * 0.001 + 0.5 + sigbar * sigbar;
        }
        warp_sync();
        // ---- dual residual norm over free variables
        double rx2 = 0, xrx = 0;
        FOR_LANE(e, n_prim(K)) {
            bool fx = false;
            if (e < K * PN) { const int k = e / PN, i = e - k * PN; fx = i < NB && fixed(k, i); }
            if (fx) rx[e] = 0.;
            rx2 += rx[e] * rx[e]; xrx += prim[e] * rx[e];
        }
        nm.gap = warp_sum(gap); nm.rz2 = warp_sum(rz2); nm.pcost = warp_sum(pcost); nm.zrz = warp_sum(zrz);
        nm.rx2 = warp_sum(rx2); nm.xrx = warp_sum(xrx); nm.h2 = warp_sum(h2); nm.bad = warp_or(bad);
    }

    // =============================================================================================================
    //  Phase F : assemble the reduced Hessian stage by stage and factor it (block-tridiagonal Cholesky + border)
    // =============================================================================================================
    // model-cone part of H_kk as rank-1 terms + diagonals in shared memory (lanes over cones)
    SCPP_HD void build_model_terms(int k, double *alpha)
    {
        double *rk = sm_rk(), *dg = sm_rk() + NRK * NB;
        FOR_LANE(e, 2 * NRK * NB) rk[e] = 0.;
        warp_sync();
        const int r0 = rn(k);
        FOR_LANE(c, NRK) {
            double *a = rk + c * NB, *d = dg + c * NB;
            if (c < NCONE) {
                const int o = NLP + M::cone_off(c), dim = M::cone_dim(c);
                const double e2i = ce[k * NCN + c];
                for (int r = 0; r < dim; r++) {
                    const RowDesc rd = M::row(o + r);
                    const double wh = (r == 0) ? wb[r0 + o] : -wb[r0 + o + r];
                    for (int j = 0; j < rd.n; j++) {
                        const double cf = coef(rd, j, k);
                        a[rd.idx[j]] += wh * cf;
                        d[rd.idx[j]] += (r == 0 ? -e2i : e2i) * cf * cf;    // -J_rr g g' (single-entry rows)
                    }
                }
                alpha[c] = 2. * e2i;
            } else {   // LP rows: single-entry rows go to the diagonal, the multi-entry row is a rank-1 term
                double al = 0.;
                for (int r = 0; r < NLP; r++) {
                    const RowDesc rd = M::row(r);
                    const double dv = wb[r0 + r];
                    if (rd.n == 1) { const double cf = coef(rd, 0, k); d[rd.idx[0]] += dv * cf * cf; }
                    else { for (int j = 0; j < rd.n; j++) a[rd.idx[j]] = coef(rd, j, k); al = dv; }
                }
                alpha[c] = al;
            }
        }
        warp_sync();
    }

    SCPP_HD bool phase_factor()
    {
        double *H = sm_H(), *O = sm_O(), *Lp = sm_Lp(), *Li = sm_Li(), *Hn = sm_Hn();
        double *bk = sm_v(0), *bn = sm_v(1), *lk = sm_v(2), *lprev = sm_v(3), *wt = sm_v(4);   // NB-vectors
        double *Dt = sm_x(1);
        double *alpha = sm_sc();
        double corner = 0.;       // lane-partial accumulation of H_sigma,sigma
        int bad = 0;
        FOR_LANE(e, BLK) { Hn[e] = 0.; Lp[e] = 0.; }
        FOR_LANE(j, NB) { bn[j] = 0.; lprev[j] = 0.; }
        warp_sync();
        for (int k = 0; k < K; k++) {
            const int r0 = rn(k);
            build_model_terms(k, alpha);
            // ---- node part of H_kk, trust region included, plus the carry from interval k-1
            {
                const double *rk = sm_rk(), *dg = sm_rk() + NRK * NB;
                const double *wt_ = wb + r0 + TRO;           // wbar of the trust-region cone (global memory)
                const double e2i = ce[k * NCN + NCONE];
                const double w0 = wt_[0];
                const double kap = e2i * (2. * w0 * w0 - 1.);   // M_00
                // p1 = -M e0 (tail) = e2i * 2 w0 w1  ;  Mtilde = e2i (I + 2 w1 w1') - p1 p1'/kap
                FOR_LANE(e, BLK) {
                    const int i = e / NB, j = e - i * NB;
                    double v = Hn[e];
                    for (int c = 0; c < NRK; c++) v += alpha[c] * rk[c * NB + i] * rk[c * NB + j];
                    if (i == j) for (int c = 0; c < NRK; c++) v += dg[c * NB + i];
                    const double wi = wt_[1 + i], wj = wt_[1 + j];
                    v += e2i * ((i == j ? 1. : 0.) + 2. * wi * wj) - (e2i * 2. * w0 * wi) * (e2i * 2. * w0 * wj) / kap;
                    H[e] = v;
                }
                FOR_LANE(j, NB) bk[j] = bn[j];
            }
            warp_sync();
            // ---- interval k: H_kk += A~' D A~ ; O = [-D A~ ; C' D A~] ; Hn = [[D, -D C],[-C' D, C' D C]] ; borders
            if (k < K - 1) {
                stage_dd(k);
                const double *t = sm_dd();
                const int q0 = ri(k);
                FOR_LANE(i, NX) { const double dm = wb[q0 + i], dp = wb[q0 + NX + i]; Dt[i] = 4. * dm * dp / (dm + dp); }
                warp_sync();
                FOR_LANE(e, BLK) {
                    const int a = e / NB, b = e - a * NB;
                    double v = 0;
                    for (int i = 0; i < NX; i++) v += t[i * NCP + a] * Dt[i] * t[i * NCP + b];
                    H[e] += v;
                    // O[a][b]: row a of node k+1, column b of node k
                    double o;
                    if (a < NX) o = -Dt[a] * t[a * NCP + b];
                    else { o = 0; for (int i = 0; i < NX; i++) o += t[i * NCP + NB + (a - NX)] * Dt[i] * t[i * NCP + b]; }
                    O[e] = o;
                    double hn;
                    if (a < NX && b < NX) hn = (a == b) ? Dt[a] : 0.;
                    else if (a < NX) hn = -Dt[a] * t[a * NCP + NB + (b - NX)];
                    else if (b < NX) hn = -Dt[b] * t[b * NCP + NB + (a - NX)];
                    else { hn = 0; for (int i = 0; i < NX; i++) hn += t[i * NCP + NB + (a - NX)] * Dt[i] * t[i * NCP + NB + (b - NX)]; }
                    Hn[e] = hn;
                }
                FOR_LANE(j, NB) {
                    double v = 0, vn;
                    for (int i = 0; i < NX; i++) v += t[i * NCP + j] * Dt[i] * t[i * NCP + NB + NU];
                    bk[j] += v;
                    if (j < NX) vn = -Dt[j] * t[j * NCP + NB + NU];
                    else { vn = 0; for (int i = 0; i < NX; i++) vn += t[i * NCP + NB + (j - NX)] * Dt[i] * t[i * NCP + NB + NU]; }
                    bn[j] = vn;
                }
                FOR_LANE(i, NX) { const double sv = t[i * NCP + NB + NU]; corner += Dt[i] * sv * sv; }
            } else {
                FOR_LANE(e, BLK) O[e] = 0.;
            }
            warp_sync();
            // ---- pinned variables: identity rows/columns
            {
                const uint32_t mk = fixm[k], mn = (k < K - 1) ? fixm[k + 1] : 0u;
                FOR_LANE(e, BLK) {
                    const int a = e / NB, b = e - a * NB;
                    if (((mk >> a) & 1u) || ((mk >> b) & 1u)) H[e] = (a == b) ? 1. : 0.;
                    if (((mn >> a) & 1u) || ((mk >> b) & 1u)) O[e] = 0.;
                }
                FOR_LANE(j, NB) if ((mk >> j) & 1u) bk[j] = 0.;
            }
            warp_sync();
            // ---- Schur update with the previous off-diagonal factor: H -= Lp Lp' ; bk -= Lp lprev
            if (k > 0) {
                FOR_LANE(e, BLK) {
                    const int a = e / NB, b = e - a * NB;
                    if (b <= a) { double v = 0; for (int c = 0; c < NB; c++) v += Lp[a * NB + c] * Lp[b * NB + c]; H[e] -= v; }
                }
                FOR_LANE(j, NB) { double v = 0; for (int c = 0; c < NB; c++) v += Lp[j * NB + c] * lprev[c]; bk[j] -= v; }
            }
            warp_sync();
            // ---- Cholesky of H (lower, in place), column by column
            for (int j = 0; j < NB; j++) {
                const double djj = H[j * NB + j];
                if (!(djj > 0.)) bad = 1;
                const double inv = 1. / sqrt(djj > 0. ? djj : 1.);
                warp_sync();
                FOR_LANE(i, NB) if (i >= j) H[i * NB + j] *= inv;   // L_jj = sqrt, L_ij = H_ij / L_jj
                warp_sync();
                // trailing update of the lower triangle
                const int rem = NB - 1 - j;
                FOR_LANE(e, rem * rem) {
                    const int a = j + 1 + e / rem, b = j + 1 + e % rem;
                    if (b <= a) H[a * NB + b] -= H[a * NB + j] * H[b * NB + j];
                }
                warp_sync();
            }
            // ---- Linv = L^-1 (lower): lane per column
            FOR_LANE(c, NB) {
                for (int i = 0; i < NB; i++) {
                    if (i < c) { Li[i * NB + c] = 0.; continue; }
                    double v = (i == c) ? 1. : 0.;
                    for (int q = c; q < i; q++) v -= H[i * NB + q] * Li[q * NB + c];
                    Li[i * NB + c] = v / H[i * NB + i];
                }
            }
            warp_sync();
            // ---- L_{k+1,k} = O L^-T = O Linv' ;  l_k = Linv bk ; corner -= l_k' l_k
            FOR_LANE(e, BLK) {
                const int a = e / NB, b = e - a * NB;
                double v = 0;
                for (int c = 0; c <= b; c++) v += O[a * NB + c] * Li[b * NB + c];
                Lp[e] = v;
            }
            FOR_LANE(j, NB) { double v = 0; for (int c = 0; c <= j; c++) v += Li[j * NB + c] * bk[c]; lk[j] = v; wt[j] = v; corner -= v * v; }
            warp_sync();
            // ---- store
            double *f = fac + (size_t)k * FACK;
            FOR_LANE(e, BLK) { f[e] = Li[e]; f[BLK + e] = Lp[e]; }
            FOR_LANE(j, NB) { f[2 * BLK + j] = lk[j]; lprev[j] = wt[j]; }
            warp_sync();
        }
        corner = warp_sum(corner);
        // ---- globals: sigma >= 0.001 row and the sigma trust-region cone with delta_sigma eliminated
        {
            const int r0 = rg();
            const double d = wb[r0];
            const double *w = wb + r0 + 1;
            const double e2i = ce[K * NCN];
            double g[3] = {-0.5, 0.5, 0.}, p[3];
            soc_M(w, e2i, g, 3, p);
            const double kap = g[0] * p[0] + g[1] * p[1];
            double e2[3] = {0., 0., 1.}, m2[3];
            soc_M(w, e2i, e2, 3, m2);
            corner += d + (m2[2] - p[2] * p[2] / kap);
        }
        if (!(corner > 0.)) bad = 1;
        if (lane_id() == 0) sm_sc()[16] = sqrt(corner > 0. ? corner : 1.);
        warp_sync();
        return !warp_or(bad);
    }

    // =============================================================================================================
    //  Phase S : solve the Newton system for one right-hand side.
    //     in : dprim = rx-like vector (per primal variable), ds = rz-like vector (per cone row)
    //     out: dprim = primal direction (xi, delta, sigma, delta_sigma, t), dz = dual direction
    // =============================================================================================================
    //     rzs: on exit ds = rzs * rz - G dx  (the primal Newton equation; keeps the primal residual contracting exactly)
    SCPP_HD void phase_solve(double rzs)
    {
        const double l_ss = sm_sc()[16];
        // ---- S1a: nodes (lane per node): gy_k = rx_k + sum_c G' v_c
        for (int k = lane_id(); k < K; k += LANES) {
            double g[NB];
            for (int i = 0; i < NB; i++) g[i] = dprim[pn(k) + i];
            const int r0 = rn(k);
            for (int r = 0; r < NLP; r++) {
                const RowDesc rd = M::row(r);
                const double v = wb[r0 + r] * ds[r0 + r];
                for (int j = 0; j < rd.n; j++) g[rd.idx[j]] += coef(rd, j, k) * v;
            }
            for (int c = 0; c < NCONE; c++) {
                const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                double v[M::MAXDIM];
                soc_M(wb + r0 + o, ce[k * NCN + c], ds + r0 + o, d, v);
                for (int r = 0; r < d; r++) { const RowDesc rd = M::row(o + r); for (int j = 0; j < rd.n; j++) g[rd.idx[j]] += coef(rd, j, k) * v[r]; }
            }
            {   // trust region with delta eliminated: v = M rz - p (p'rz + rx_delta)/kap , p = -M e0
                const double *w = wb + r0 + TRO, *rzv = ds + r0 + TRO;
                const double e2i = ce[k * NCN + NCONE];
                double v[1 + NB], p[1 + NB], e0[1 + NB];
                for (int i = 0; i <= NB; i++) e0[i] = (i == 0) ? -1. : 0.;
                soc_M(w, e2i, rzv, 1 + NB, v);
                soc_M(w, e2i, e0, 1 + NB, p);
                const double kap = -p[0];
                double prz = 0;
                for (int i = 0; i <= NB; i++) prz += p[i] * rzv[i];
                const double rho = (prz + dprim[pn(k) + NB]) / kap;
                for (int i = 0; i < NB; i++) g[i] += v[1 + i] - p[1 + i] * rho;
            }
            for (int i = 0; i < NB; i++) gy[k * NB + i] = g[i];
        }
        warp_sync();
        // ---- S1b: intervals
        double gsig = 0;
        for (int k = 0; k < K - 1; k++) {
            stage_dd(k);
            const int q0 = ri(k);
            double *w = sm_x(0);
            FOR_LANE(i, NX) {
                const double dm = wb[q0 + i], dp = wb[q0 + NX + i], rm = ds[q0 + i], rp = ds[q0 + NX + i];
                const double rho = (-(dm * rm + dp * rp) + dprim[p_t(k) + i]) / (dm + dp);
                w[i] = dm * (rm + rho) - dp * (rp + rho);
            }
            warp_sync();
            dyn_JT(k, w, gy + k * NB, gy + (k + 1) * NB, gsig);
            warp_sync();
        }
        gsig = warp_sum(gsig);
        // ---- S1c: globals (every lane computes the same scalars)
        double kap_s, p_s[3];
        {
            const int r0 = rg();
            const double *w = wb + r0 + 1, *rzv = ds + r0 + 1;
            const double e2i = ce[K * NCN];
            double g3[3] = {-0.5, 0.5, 0.}, v[3];
            soc_M(w, e2i, g3, 3, p_s);
            kap_s = g3[0] * p_s[0] + g3[1] * p_s[1];
            soc_M(w, e2i, rzv, 3, v);
            const double prz = p_s[0] * rzv[0] + p_s[1] * rzv[1] + p_s[2] * rzv[2];
            const double rho = (prz + dprim[p_dsig()]) / kap_s;
            gsig += dprim[p_sigma()] - wb[r0] * ds[r0] - (v[2] - p_s[2] * rho);
        }
        // ---- S2: forward sweep  f_k = Linv_k (g_k - L_{k,k-1} f_{k-1}) ; pinned entries are zero
        double *fprev = sm_v(5), *tmp = sm_v(6);
        double ldot = 0.;      // lane-partial of sum_k l_k' f_k
        for (int k = 0; k < K; k++) {
            const double *f = fac + (size_t)k * FACK;
            const double *fm = fac + (size_t)(k - 1) * FACK;
            FOR_LANE(j, NB) {
                double v = fixed(k, j) ? 0. : gy[k * NB + j];
                if (k > 0) for (int c = 0; c < NB; c++) v -= fm[BLK + j * NB + c] * fprev[c];
                tmp[j] = v;
            }
            warp_sync();
            FOR_LANE(j, NB) {
                double v = 0;
                for (int c = 0; c <= j; c++) v += f[j * NB + c] * tmp[c];
                gy[k * NB + j] = v;
                ldot += f[2 * BLK + j] * v;
            }
            warp_sync();
            FOR_LANE(j, NB) fprev[j] = gy[k * NB + j];
            warp_sync();
        }
        const double fsig = (gsig - warp_sum(ldot)) / l_ss;
        const double ysig = fsig / l_ss;
        // ---- S3: backward sweep  y_k = Linv_k' (f_k - L_{k+1,k}' y_{k+1} - l_k y_sigma)
        double *ynext = sm_v(5);
        for (int k = K - 1; k >= 0; k--) {
            const double *f = fac + (size_t)k * FACK;
            FOR_LANE(j, NB) {
                double v = gy[k * NB + j] - f[2 * BLK + j] * ysig;
                if (k < K - 1) for (int c = 0; c < NB; c++) v -= f[BLK + c * NB + j] * ynext[c];
                tmp[j] = v;
            }
            warp_sync();
            FOR_LANE(j, NB) {
                double v = 0;
                for (int c = j; c < NB; c++) v += f[c * NB + j] * tmp[c];
                gy[k * NB + j] = fixed(k, j) ? 0. : v;
            }
            warp_sync();
            FOR_LANE(j, NB) ynext[j] = gy[k * NB + j];
            warp_sync();
        }
        if (lane_id() == 0) gy[K * NB] = ysig;
        warp_sync();
        // ---- S4a: recovery at the nodes
        for (int k = lane_id(); k < K; k += LANES) {
            double dx[NB];
            for (int i = 0; i < NB; i++) dx[i] = gy[k * NB + i];
            const int r0 = rn(k);
            for (int r = 0; r < NLP; r++) {
                const RowDesc rd = M::row(r);
                double gdx = 0;
                for (int j = 0; j < rd.n; j++) gdx += coef(rd, j, k) * dx[rd.idx[j]];
                dz[r0 + r] = wb[r0 + r] * (gdx - ds[r0 + r]);
                ds[r0 + r] = rzs * rz[r0 + r] - gdx;
            }
            for (int c = 0; c < NCONE; c++) {
                const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                double q[M::MAXDIM];
                for (int r = 0; r < d; r++) {
                    const RowDesc rd = M::row(o + r);
                    double gdx = 0;
                    for (int j = 0; j < rd.n; j++) gdx += coef(rd, j, k) * dx[rd.idx[j]];
                    q[r] = gdx - ds[r0 + o + r];
                    ds[r0 + o + r] = rzs * rz[r0 + o + r] - gdx;
                }
                soc_M(wb + r0 + o, ce[k * NCN + c], q, d, dz + r0 + o);
            }
            {
                const double *w = wb + r0 + TRO;
                const double e2i = ce[k * NCN + NCONE];
                double q[1 + NB], p[1 + NB], e0[1 + NB], mq[1 + NB];
                q[0] = -ds[r0 + TRO];
                for (int i = 0; i < NB; i++) q[1 + i] = dx[i] - ds[r0 + TRO + 1 + i];
                for (int i = 0; i <= NB; i++) e0[i] = (i == 0) ? -1. : 0.;
                soc_M(w, e2i, e0, 1 + NB, p);
                const double kap = -p[0];
                double pq = 0;
                for (int i = 0; i <= NB; i++) pq += p[i] * q[i];
                const double ddl = (dprim[pn(k) + NB] - pq) / kap;
                soc_M(w, e2i, q, 1 + NB, mq);
                for (int i = 0; i <= NB; i++) dz[r0 + TRO + i] = mq[i] + p[i] * ddl;
                ds[r0 + TRO] = rzs * rz[r0 + TRO] + ddl;
                for (int i = 0; i < NB; i++) ds[r0 + TRO + 1 + i] = rzs * rz[r0 + TRO + 1 + i] - dx[i];
                for (int i = 0; i < NB; i++) dprim[pn(k) + i] = dx[i];
                dprim[pn(k) + NB] = ddl;
            }
        }
        warp_sync();
        // ---- S4b: recovery on the intervals
        for (int k = 0; k < K - 1; k++) {
            stage_dd(k);
            const int q0 = ri(k);
            FOR_LANE(i, NX) {
                // a' dy = J dy  (no constant): uses dprim node parts and sigma from gy
                const double *t = sm_dd() + i * NCP;
                const double *xk = dprim + pn(k), *xn = dprim + pn(k + 1);
                double ady = xn[i];
                for (int j = 0; j < NB; j++) ady -= t[j] * xk[j];
                for (int j = 0; j < NU; j++) ady -= t[NB + j] * xn[NX + j];
                ady -= t[NB + NU] * ysig;
                const double dm = wb[q0 + i], dp = wb[q0 + NX + i];
                const double qm = ady - ds[q0 + i], qp = -ady - ds[q0 + NX + i];
                const double dt = (dprim[p_t(k) + i] + dm * qm + dp * qp) / (dm + dp);
                dz[q0 + i] = dm * (qm - dt);
                dz[q0 + NX + i] = dp * (qp - dt);
                ds[q0 + i] = rzs * rz[q0 + i] - (ady - dt);
                ds[q0 + NX + i] = rzs * rz[q0 + NX + i] - (-ady - dt);
                dprim[p_t(k) + i] = dt;
            }
            warp_sync();
        }
        // ---- S4c: globals
        if (lane_id() == 0) {
            const int r0 = rg();
            dz[r0] = wb[r0] * (-ysig - ds[r0]);
            const double *w = wb + r0 + 1;
            const double e2i = ce[K * NCN];
            double q[3] = {-ds[r0 + 1], -ds[r0 + 2], -ysig - ds[r0 + 3]}, mq[3];
            const double pq = p_s[0] * q[0] + p_s[1] * q[1] + p_s[2] * q[2];
            const double dds = (dprim[p_dsig()] - pq) / kap_s;
            soc_M(w, e2i, q, 3, mq);
            for (int i = 0; i < 3; i++) dz[r0 + 1 + i] = mq[i] + p_s[i] * dds;
            ds[r0] = rzs * rz[r0] + ysig;
            ds[r0 + 1] = rzs * rz[r0 + 1] + 0.5 * dds;
            ds[r0 + 2] = rzs * rz[r0 + 2] - 0.5 * dds;
            ds[r0 + 3] = rzs * rz[r0 + 3] + ysig;
            dprim[p_dsig()] = dds;
            dprim[p_sigma()] = ysig;
        }
        warp_sync();
    }

    // ---- cone-wise helpers over all rows (lanes over cones / rows) ----------------------------------------------
    // visit every cone: f(offset, dim, cone_index or -1 for an LP row)
    template <class F>
    SCPP_HD void for_cones(F &&f) const
    {
        for (int k = lane_id(); k < K; k += LANES) {
            const int r0 = rn(k);
            for (int r = 0; r < NLP; r++) f(r0 + r, 1, -1);
            for (int c = 0; c < NCONE; c++) f(r0 + NLP + M::cone_off(c), M::cone_dim(c), k * NCN + c);
            f(r0 + TRO, 1 + NB, k * NCN + NCONE);
        }
        FOR_LANE(e, (K - 1) * 2 * NX) f(ri(0) + e, 1, -1);
        if (lane_id() == 0) { f(rg(), 1, -1); f(rg() + 1, 3, K * NCN); }
    }

    // s (or z) <- slack margins; returns min over cones of (u0 - |u1|) and |u|^2
    SCPP_HD void cone_margin(const double *u, double &mn, double &nrm2) const
    {
        double lmn = 1e300, n2 = 0;
        for_cones([&](int o, int d, int) {
            double t = 0;
            for (int i = 1; i < d; i++) t += u[o + i] * u[o + i];
            const double mg = u[o] - sqrt(t);
            if (mg < lmn) lmn = mg;
            n2 += t + u[o] * u[o];
        });
        mn = -warp_max(-lmn); nrm2 = warp_sum(n2);
    }
    SCPP_HD void cone_shift(double *u, double a) const { for_cones([&](int o, int, int) { u[o] += a; }); }

    // slack(prim) = h - G x into `out`
    SCPP_HD void eval_slack(double *out)
    {
        for (int k = lane_id(); k < K; k += LANES) {
            const double *xi = prim + pn(k);
            const int r0 = rn(k);
            for (int r = 0; r < NLP + NCR; r++) out[r0 + r] = row_slack(r, k, xi);
            out[r0 + TRO] = xi[NB];
            for (int i = 0; i < NB; i++) out[r0 + TRO + 1 + i] = xibar(k, i) - xi[i];
        }
        for (int k = 0; k < K - 1; k++) {
            stage_dd(k);
            const int q0 = ri(k);
            FOR_LANE(i, NX) { const double r = dyn_resid(k, i, prim, true), t = prim[p_t(k) + i]; out[q0 + i] = t - r; out[q0 + NX + i] = t + r; }
            warp_sync();
        }
        if (lane_id() == 0) {
            const int r0 = rg();
            const double sg = prim[p_sigma()], dsg = prim[p_dsig()];
            out[r0] = sg - 0.001; out[r0 + 1] = 0.5 + 0.5 * dsg; out[r0 + 2] = 0.5 - 0.5 * dsg; out[r0 + 3] = sg - sigbar;
        }
        warp_sync();
    }

    // =============================================================================================================
    //  driver
    // =============================================================================================================
    SCPP_HD IpmResult solve(const IpmSettings &st)
    {
        IpmResult res;
        res.status = 1; res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.;
        const int np = n_prim(K), m = m_rows(K);
        // ---- starting point (CVXOPT conelp / ECOS style): least-squares primal and dual points, W = I
        for (int k = lane_id(); k < K; k += LANES) {
            for (int i = 0; i < NB; i++) prim[pn(k) + i] = fixed(k, i) ? fixv[k * NB + i] : xibar(k, i);
            prim[pn(k) + NB] = 0.;
        }
        FOR_LANE(e, (K - 1) * NX) prim[p_t(0) + e] = 0.;
        if (lane_id() == 0) { prim[p_sigma()] = sigbar; prim[p_dsig()] = 0.; }
        FOR_LANE(e, m) { s[e] = 0.; z[e] = 0.; }
        warp_sync();
        cone_shift(s, 1.); cone_shift(z, 1.);
        warp_sync();
        Norms nm;
        phase_residuals(nm, true);
        if (!phase_factor()) { res.status = 2; return res; }
        // primal: min |G x - h|  ->  G dx - dz = slack(x0)
        eval_slack(ds);
        FOR_LANE(e, np) dprim[e] = 0.;
        warp_sync();
        phase_solve(0.);
        FOR_LANE(e, np) prim[e] += dprim[e];
        warp_sync();
        eval_slack(s);
        {
            double mg, n2; cone_margin(s, mg, n2);
            if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift(s, 1. - mg); }
            warp_sync();
        }
        // dual: min |z| s.t. G'z + c = 0  ->  rx = -c, rz = 0
        FOR_LANE(e, np) dprim[e] = 0.;
        FOR_LANE(e, m) ds[e] = 0.;
        warp_sync();
        for (int k = lane_id(); k < K; k += LANES) dprim[pn(k) + NB] = -w_tr;
        FOR_LANE(e, (K - 1) * NX) dprim[p_t(0) + e] = -w_vc;
        if (lane_id() == 0) { dprim[p_sigma()] = -w_time; dprim[p_dsig()] = -w_trs; }
        warp_sync();
        phase_solve(0.);
        FOR_LANE(e, m) z[e] = dz[e];
        warp_sync();
        {
            double mg, n2; cone_margin(z, mg, n2);
            if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift(z, 1. - mg); }
            warp_sync();
        }
        const double cnorm = sqrt(w_time * w_time + w_trs * w_trs + K * w_tr * w_tr + (K - 1) * NX * w_vc * w_vc);
        const double resx0 = fmax(1., cnorm);
        const int degree = K * (NLP + NCN) + (K - 1) * 2 * NX + 2;
        double best = 1e300;
        int it;
        for (it = 0; it <= st.maxit; it++) {
            phase_residuals(nm, false);
            const double resz0 = fmax(1., sqrt(nm.h2));
            const double pres = sqrt(nm.rz2) / resz0, dres = sqrt(nm.rx2) / resx0, gap = nm.gap, pcost = nm.pcost;
            const double dcost = pcost - gap + nm.zrz - nm.xrx;
            double relgap = 1e300;
            if (pcost < 0.) relgap = gap / -pcost; else if (dcost > 0.) relgap = gap / dcost;
            const double score = fmax(fmax(pres, dres) / st.feastol, fmin(gap / st.abstol, relgap / st.reltol));
#if !defined(__CUDACC__)
            if (getenv("SCPP_DEBUG")) fprintf(stderr, "it %2d pres %.2e dres %.2e gap %.2e relgap %.2e pcost %.6e bad %d\n", it, pres, dres, gap, relgap, pcost, nm.bad);
#endif
            if (!nm.bad && score < best) {
                best = score;
                res.pres = pres; res.dres = dres; res.gap = gap; res.relgap = relgap; res.pcost = pcost; res.iterations = it;
                // keep the best iterate (primal only matters downstream): cr is free at this point of the iteration
                FOR_LANE(e, np) rxbest()[e] = prim[e];
                warp_sync();
            }
            if (!nm.bad && pres <= st.feastol && dres <= st.feastol && (gap <= st.abstol || relgap <= st.reltol)) { res.status = 0; break; }
            if (nm.bad || it == st.maxit || (score > 1e3 * best && best < 1e4)) { res.status = nm.bad ? 2 : (it == st.maxit ? 1 : 2); break; }
            if (!phase_factor()) { res.status = 2; break; }
            // ---- affine direction: rx-like = -rx ; rz-like = -rz + s
            FOR_LANE(e, np) dprim[e] = -rx[e];
            FOR_LANE(e, m) ds[e] = -rz[e] + s[e];
            warp_sync();
            phase_solve(-1.);
            // scaled directions dz~ = W dz, ds~ = W^-1 ds ; step to the boundary ; cr = ds~ o dz~
            double tmax = 0;
            for_cones([&](int o, int d, int ci) {
                if (d == 1) {
                    const double w = sqrt(1. / wb[o]);           // W = sqrt(s/z)
                    const double dzt = w * dz[o], dst = ds[o] / w;
                    tmax = fmax(tmax, fmax(-dst / lam[o], -dzt / lam[o]));
                    cr[o] = dst * dzt;
                } else {
                    double dzt[1 + NB], dst[1 + NB];
                    soc_W(wb + o, ce[ci], dz + o, d, dzt, false);
                    soc_W(wb + o, ce[ci], ds + o, d, dst, true);
                    tmax = fmax(tmax, fmax(soc_step(lam + o, dst, d), soc_step(lam + o, dzt, d)));
                    soc_jprod(dst, dzt, d, cr + o);
                }
            });
            tmax = warp_max(tmax);
            const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
            const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = gap / degree;
            // ---- combined direction: d_s = -lam o lam - cr + sig mu e ; rz-like = -(1-sig) rz - W (lam \ d_s)
            FOR_LANE(e, np) dprim[e] = -(1. - sig) * rx[e];
            for_cones([&](int o, int d, int ci) {
                if (d == 1) {
                    const double w = sqrt(1. / wb[o]);
                    const double dsv = -lam[o] * lam[o] - cr[o] + sig * mu;
                    const double t1 = dsv / lam[o];
                    ds[o] = -(1. - sig) * rz[o] - w * t1;
                } else {
                    double dsv[1 + NB], t1[1 + NB], wt[1 + NB];
                    soc_jprod(lam + o, lam + o, d, dsv);
                    for (int i = 0; i < d; i++) dsv[i] = -dsv[i] - cr[o + i];
                    dsv[0] += sig * mu;
                    soc_jdiv(lam + o, dsv, d, t1);
                    soc_W(wb + o, ce[ci], t1, d, wt, false);
                    for (int i = 0; i < d; i++) ds[o + i] = -(1. - sig) * rz[o + i] - wt[i];
                }
            });
            warp_sync();
            phase_solve(-(1. - sig));
            tmax = 0;
            for_cones([&](int o, int d, int ci) {
                if (d == 1) {
                    const double w = sqrt(1. / wb[o]);
                    tmax = fmax(tmax, fmax(-(ds[o] / w) / lam[o], -(w * dz[o]) / lam[o]));
                } else {
                    double dzt[1 + NB], dst[1 + NB];
                    soc_W(wb + o, ce[ci], dz + o, d, dzt, false);
                    soc_W(wb + o, ce[ci], ds + o, d, dst, true);
                    tmax = fmax(tmax, fmax(soc_step(lam + o, dst, d), soc_step(lam + o, dzt, d)));
                }
            });
            tmax = warp_max(tmax);
            double alpha = tmax <= 0.99 ? 1. : 0.99 / tmax;
            // additive update; back off if rounding leaves the cone
            for (int tries = 0; tries < 20; tries++) {
                double lmn = 1e300;
                for_cones([&](int o, int d, int) {
                    double ts = 0, tz = 0;
                    for (int i = 1; i < d; i++) { const double a = s[o + i] + alpha * ds[o + i], b = z[o + i] + alpha * dz[o + i]; ts += a * a; tz += b * b; }
                    const double ms = s[o] + alpha * ds[o] - sqrt(ts), mz = z[o] + alpha * dz[o] - sqrt(tz);
                    lmn = fmin(lmn, fmin(ms, mz));
                });
                lmn = -warp_max(-lmn);
                if (lmn > 0.) break;
                alpha *= 0.8;
            }
            FOR_LANE(e, np) prim[e] += alpha * dprim[e];
            FOR_LANE(e, m) { s[e] += alpha * ds[e]; z[e] += alpha * dz[e]; }
            warp_sync();
        }
        if (res.status != 0) {
            // fall back to the best iterate seen
            FOR_LANE(e, np) prim[e] = rxbest()[e];
            warp_sync();
            if (best <= 10.) res.status = 0; else if (best <= 1e4) res.status = 3;
        } else res.iterations = it;
        return res;
    }
    SCPP_HD double *rxbest() const { return best_; }
};

} // namespace scpp
