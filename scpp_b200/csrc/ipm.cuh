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
// Execution shape (v3).  The work of one interior-point iteration is of two kinds and each gets its own mapping:
//   * STAGE-PARALLEL passes (residuals + Nesterov-Todd scaling, right-hand sides, recovery of dz/ds and the step
//     lengths, the update):  nothing couples the K stages (stage k = node k + shooting interval k) except two
//     nearest-neighbour reads, so ONE LANE runs the scalar cone arithmetic of ONE STAGE (K = 50: two rounds of the
//     warp).  All per-stage state lives in HBM in STAGE-MINOR arrays  a[row * KS + k]  so that the 32 lanes of a
//     load/store touch 32 consecutive doubles (fully coalesced), there are no shuffles, no divergence between the
//     cone types and no shared-memory staging in these passes.
//   * the CHAIN (block-tridiagonal Cholesky, forward and backward substitution) is sequential in k; there the warp
//     cooperates on one stage at a time: per-stage records are staged in shared memory with 16-byte asynchronous
//     copies (prefetched one stage ahead), block products run on the FP64 tensor cores (blockops.cuh), the 18x18
//     Cholesky and the triangular inverse keep one matrix row / column per lane.
#pragma once
#include "models.cuh"
#include "blockops.cuh"
#if !defined(__CUDACC__)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#endif

namespace scpp {

// lanes stride over the warp's stage range (members k_lo, k_hi of Ipm)
#define FOR_STAGE(k) for (int k = k_lo + lane_id(); k < k_hi; k += LANES)

struct IpmSettings {
    double feastol, abstol, reltol;
    int maxit;
    int stalled_step;   // see sc.cuh: sc_step_fraction
    double warm;     // 0: cold start of every sub-problem (what ECOS does).  0 < warm < 1: when the instance has a previous sub-problem
                     // solution, start from that interior point pulled back from the boundary, (s,z) <- warm*(s,z) + (1-warm)*e,
                     // and skip the least-squares start.  Same optimum (parity-tested), ~2.8x fewer interior-point iterations.
};

struct IpmResult {
    int status;      // 0 optimal, 1 max iterations, 2 numerical failure, 3 reduced accuracy
    int iterations;
    double pres, dres, gap, relgap, pcost;
    int point_ok;    // the (s, z) left in the workspace are a valid interior point (false when the last iterate left the cones: no warm start from it)
    int pad_;
};

// ---- second-order-cone primitives on small local arrays (dimension <= 4: the model cones) ------------------------------
namespace soc {
constexpr int SOC_MAXD = 4;
SCPP_HD double jn2(const double *u, int d)
{
    double n = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) n += u[i] * u[i];
    return u[0] * u[0] - n;
}
SCPP_HD bool scale(const double *sk, const double *zk, int d, double *w, double &e2i, double *lm)
{
    double ss = jn2(sk, d), zz = jn2(zk, d);
    if (!(ss > 0.) || !(zz > 0.) || !(sk[0] > 0.) || !(zk[0] > 0.)) return false;
    double sn = sqrt(ss), zn = sqrt(zz), sz = 0;
#pragma unroll
    for (int i = 0; i < SOC_MAXD; i++) if (i < d) sz += sk[i] * zk[i];
    double gam = sqrt((1. + sz / (sn * zn)) / 2.);
    double i2g = 1. / (2. * gam);
    const double isn = i2g / sn, izn = i2g / zn;
    w[0] = sk[0] * isn + zk[0] * izn;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) w[i] = sk[i] * isn - zk[i] * izn;
    e2i = zn / sn;
    double eta = sqrt(sn / zn), w1z1 = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) w1z1 += w[i] * zk[i];
    double f = zk[0] + w1z1 / (1. + w[0]);
    lm[0] = eta * (w[0] * zk[0] + w1z1);
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) lm[i] = eta * (zk[i] + f * w[i]);
    return true;
}
// o = W^-2 v  (o may alias v)
SCPP_HD void Mv(const double *w, double e2i, const double *v, int d, double *o)
{
    double dot = w[0] * v[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) dot -= w[i] * v[i];
    const double v0 = v[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = e2i * (-2. * dot * w[i] + v[i]);
    o[0] = e2i * (2. * dot * w[0] - v0);
}
// o = W v | W^-1 v  (o may alias v)
SCPP_HD void Wv(const double *w, double e2i, const double *v, int d, double *o, bool inv)
{
    const double eta = 1. / sqrt(e2i);
    const double sg = inv ? -1. : 1., sc = inv ? 1. / eta : eta;
    double w1v1 = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) w1v1 += w[i] * v[i];
    const double o0 = w[0] * v[0] + sg * w1v1, f = sg * v[0] + w1v1 / (1. + w[0]);
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = sc * (v[i] + f * w[i]);
    o[0] = sc * o0;
}
SCPP_HD void jprod(const double *u, const double *v, int d, double *o)   // o may alias u or v
{
    double dot = 0;
#pragma unroll
    for (int i = 0; i < SOC_MAXD; i++) if (i < d) dot += u[i] * v[i];
    const double u0 = u[0], v0 = v[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = u0 * v[i] + v0 * u[i];
    o[0] = dot;
}
SCPP_HD void jdiv(const double *lm, const double *dv, int d, double *o)   // o = lm \ dv  (o may alias dv)
{
    double den = jn2(lm, d), l1d1 = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) l1d1 += lm[i] * dv[i];
    const double x0 = (lm[0] * dv[0] - l1d1) / den, il0 = 1. / lm[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = (dv[i] - x0 * lm[i]) * il0;
    o[0] = x0;
}
SCPP_HD double step(const double *lm, const double *dk, int d)
{
    const double ia = 1. / sqrt(jn2(lm, d)), l0 = lm[0] * ia;
    double ld = l0 * dk[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) ld -= lm[i] * ia * dk[i];
    const double rho0 = ld * ia, f = (ld + dk[0]) / (l0 + 1.) * ia;
    double n1 = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) { const double r = dk[i] - f * lm[i]; n1 += r * r; }
    return sqrt(n1) * ia - rho0;
}
} // namespace soc

SCPP_HD constexpr int pad2(int x) { return (x + 1) & ~1; }

template <class M>
struct Ipm {
    static constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2, NCP = NC + 2;
    static constexpr int NLP = M::NLP, NCONE = M::NCONE, NCR = M::NCR;
    static constexpr int MN = NLP + NCR + 1 + NB;   // cone rows per node: model LP | model cones | trust region
    static constexpr int PN = NB + 1;               // primal per node: xi, delta
    static constexpr int NCN = NCONE + 1;           // second-order cones per node (model + trust region)
    static constexpr int NRK = NCONE + 1;           // rank-1 terms of the model Hessian: cones + one multi-entry LP row
    static constexpr int TRO = NLP + NCR;           // offset of the trust-region cone inside a node block
    static constexpr int BLK = NB * NB;
    static constexpr int D = 1 + NB;                // dimension of the trust-region cone
    static constexpr int RS = MN + 2 * NX;          // cone rows of one stage: node rows | interval rows (s-: NX, s+: NX)
    static constexpr int PSN = PN + NX;             // primal rows of one stage: xi | delta | t
    static constexpr int NROW = NLP + NCR;          // model rows of one node
    static constexpr int FS = pad2(2 * BLK + NB);   // factor record of one stage (stage-major): Linv_kk | L_{k+1,k} | l_k
    static constexpr int OFF_LN = BLK, OFF_L = 2 * BLK;
    static_assert(M::MAXDIM <= soc::SOC_MAXD, "single-lane cone primitives are unrolled for dimension <= 4");
    static_assert(NB % 2 == 0 && NC % 2 == 0 && NX % 2 == 0, "16-byte record alignment needs even nx, nx+nu and tile width");

    // stage-minor arrays:  element (row r, stage k) at  r * KS + k ;  the 4 global rows / 2 global primals follow the stage part
    SCPP_HD static int ks(int K) { return (K + 3) & ~3; }
    SCPP_HD static int m_rows(int K) { return RS * ks(K) + 4; }
    SCPP_HD static int n_prim(int K) { return PSN * ks(K) + 2; }
    SCPP_HD static int n_ce(int K) { return NCN * ks(K) + 2; }
    static constexpr int PARTS = 8, PSTR = 16;   // split pipeline: per-warp partial sums of the stage-parallel passes (K <= 32 * PARTS)
    SCPP_HD static int ws_doubles(int K) { return 4 * n_prim(K) + 8 * m_rows(K) + n_ce(K) + (NB + 2 * NX + 1) * ks(K) + PARTS * PSTR + K * FS + 16; }
    SCPP_HD static int ddt_doubles(int K) { return NX * NC * ks(K); }
    // Shared window of a warp.  Front: per-solve tables and small vectors (every kernel).  Then a union:
    //   inline factorisation (monolithic kernel):  two tile buffers | record out | L_{k,k-1} | H | O | model terms
    //   substitutions / chain factorisation:       ring of RING records | L_{k,k-1} | Linv scratch
    //   assembly kernel (one warp per stage):      tile | carry data of interval k-1 | H | O | model terms
    static constexpr int HNC = pad2(NX + NX * NU + NU * NU);      // D | D C | C' D C  of the previous interval
#ifndef SCPP_RING
#define SCPP_RING 3
#endif
    static constexpr int RING = SCPP_RING;      // factor records in flight in the substitutions (3 or 4).  4 was measured (profiles/README, r03h): no faster by itself, and the
                                                // 213 KB of shared memory it needs per CTA leave 28 KB of L1 instead of 60 KB: K2 11 % slower
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int W_CST = 0, W_RCQ = W_CST + pad2(MAX_CST + 4), W_RIDX = W_RCQ + 4 * NROW, W_REV = W_RIDX + pad2(2 * NROW),
                         W_SC = W_REV + pad2(2 * NB), W_VEC = W_SC + 32, W_X = W_VEC + 6 * NB, W_WB = W_X + 2 * pad2(NX), W_HN = W_WB + pad2(RS),
                         W_U = W_HN + HNC,
                         W_DD = W_U, W_FAC = W_DD + 2 * pad2(NX * NCP), W_LP = W_FAC + FS, W_MAT = W_LP + BLK, W_RK = W_MAT + 2 * BLK,
                         W_F_END = W_RK + 2 * NRK * NB,
                         W_RING = W_U, W_CLP = W_RING + RING * FS, W_CLI = W_CLP + BLK, W_C_END = W_CLI + BLK,
                         AW_DD = W_U, AW_PC = AW_DD + pad2(NX * NCP), AW_MAT = AW_PC + pad2(NX * NU + 3 * NX), AW_RK = AW_MAT + 2 * BLK,
                         AW_END = AW_RK + 2 * NRK * NB,
                         W_END = cmax(W_F_END, W_C_END);
    SCPP_HD static int asm_doubles() { return AW_END; }
    SCPP_HD static int sm_doubles() { return W_END; }

    // ---- problem data (read only) -----------------------------------------------------------------------------
    int K, KS;
    int k_lo, k_hi;        // stage range of the stage-parallel passes run by this warp ([0,K) unless the pass is split over warps)
    const double *dd;      // [K-1][NX][NC]   stage-major tiles (chain)
    const double *ddT;     // [NX*NC][KS]     the same tiles, stage-minor (stage-parallel passes)
    const double *Xbar;    // [K][NX]
    const double *Ubar;    // [K][NU]
    double sigbar;
    const double *cst;     // per-instance constants of the row table
    const double *tdir;    // [K][3]
    const uint32_t *fixm;  // [K]
    const double *fixv;    // [K][NB]
    double w_time, w_trs, w_tr, w_vc;
    // SCvx variant (buildSCvxProblem, scpp_core/src/SCvxProblem.cpp:6-71): fixed final time (sigma pinned: the global rows are skipped),
    // hard trust region |ubar_k - u_k| <= tr_rad on the INPUT rows of the node cone only (delta_k pinned to tr_rad, state rows inactive:
    // their s, z, w, lambda stay exactly zero), cost w_vc |nu|_1 alone (caller sets w_time = w_trs = w_tr = 0)
    bool scvx = false;
    double tr_rad = 0.;
    // sigma pinned to sigbar, the global rows (sigma >= 0.001, the sigma trust region) and the border of the Newton system skipped: SCvx, and SC
    // with free_final_time = false (SCProblem.cpp:27-35,49-56,82-100: no sigma / delta_sigma variables; the fixed-time z_k of
    // discretizationImplementation.hpp:111-116 equals z_k + s_k sigbar of the free-time tile, so K1 is unchanged).  scvx implies sig_fixed.
    bool sig_fixed = false;
    // fraction of the step to the cone boundary taken by an iterate (ECOS / CVXOPT: 0.99); sc.cuh: sc_step_fraction may raise it for the
    // sub-problem after a stalled outer iteration (experiment knob ipm.stalled_step, default off)
    double step_frac = 0.99;
    // ---- workspace (global memory, per instance) -------------------------------------------------------------------
    double *prim, *dprim, *rx, *best_;
    double *s, *z, *wb, *lam, *rz, *cr, *dz, *ds;
    double *ce;
    double *gv;            // [NB][KS]  right-hand side g -> forward solution f -> solution y
    double *wv;            // [NX][KS]  w of interval k (coupling to node k+1)
    double *cpart;         // [KS]      corner contribution of interval k (assembly kernel)
    double *dtv;           // [NX][KS]  D of interval k (CTA-per-instance solver)
    double *part;          // [PARTS][PSTR] partial sums of the split stage-parallel passes
    double *fac;           // [K][FS]
    double *sm;            // per-warp shared window
    double l_ss;           // Cholesky pivot of the sigma border

    SCPP_HD void bind(double *ws, double *smem)
    {
        KS = ks(K); k_lo = 0; k_hi = K;
        const int np = n_prim(K), m = m_rows(K);
        double *p = ws;
        prim = p; p += np; dprim = p; p += np; rx = p; p += np; best_ = p; p += np;
        s = p; p += m; z = p; p += m; wb = p; p += m; lam = p; p += m; rz = p; p += m; cr = p; p += m; dz = p; p += m; ds = p; p += m;
        ce = p; p += n_ce(K);
        gv = p; p += NB * KS;
        wv = p; p += NX * KS;
        cpart = p; p += KS;
        dtv = p; p += NX * KS;
        part = p; p += PARTS * PSTR;
        fac = p;
        sm = smem;
    }
    // accessors used by the SC glue (sc.cuh)
    SCPP_HD double xi_at(int k, int i) const { return prim[i * KS + k]; }
    SCPP_HD double delta_at(int k) const { return prim[NB * KS + k]; }
    SCPP_HD double t_at(int k, int i) const { return prim[(PN + i) * KS + k]; }
    SCPP_HD double sigma_val() const { return prim[PSN * KS]; }
    SCPP_HD double dsigma_val() const { return prim[PSN * KS + 1]; }

    SCPP_HD bool fixed(int k, int i) const { return (fixm[k] >> i) & 1u; }
    SCPP_HD bool tr_row(int i) const { return !scvx || i > NX; }      // is tail row i (1..NB) of the node cone active?
    SCPP_HD double xibar(int k, int i) const { return i < NX ? Xbar[k * NX + i] : Ubar[k * NU + (i - NX)]; }
    SCPP_HD double T(int i, int j, int k) const { return ddT[(i * NC + j) * KS + k]; }     // tile element, stage-minor

    // ---- shared-window plumbing (chain phases) ----------------------------------------------------------------------
    SCPP_HD void ld(double *dst, const double *src, int n) const   // n even, 16 bytes per lane per step, asynchronous on the device
    {
#if defined(__CUDA_ARCH__)
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        for (int c = lane_id(); c < n / 2; c += LANES)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * c), "l"(src + 2 * c) : "memory");
#else
        memcpy(dst, src, sizeof(double) * n);
#endif
    }
    SCPP_HD void ld_dd(int k) const   // the [A|B|C|s|z]_k tile: NX rows of NC doubles into rows of stride NCP (buffer k & 1)
    {
        const double *src = dd + (size_t)k * NX * NC;
        double *t = tile(k);
#if defined(__CUDA_ARCH__)
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(t);
        for (int c = lane_id(); c < NX * (NC / 2); c += LANES) {
            const int r = c / (NC / 2), q = c - r * (NC / 2);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 8u * (r * NCP + 2 * q)), "l"(src + r * NC + 2 * q) : "memory");
        }
#else
        for (int r = 0; r < NX; r++) memcpy(t + r * NCP, src + r * NC, sizeof(double) * NC);
#endif
    }
    SCPP_HD void ld_commit() const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    }
    SCPP_HD void ld_wait(int pending = 0) const   // all but the `pending` most recent groups have landed
    {
#if defined(__CUDA_ARCH__)
        if (pending == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
        else if (pending == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (pending == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
        warp_sync();
    }
    SCPP_HD void st(double *dst, const double *src, int n) const { FOR_LANE(e, n) dst[e] = src[e]; }
    SCPP_HD void copy4(double *dst, const double *src, int n) const      // global -> global, four independent loads in flight per lane
    {
        int e = lane_id();
        for (; e + 3 * LANES < n; e += 4 * LANES) {
            const double a = src[e], b = src[e + LANES], c = src[e + 2 * LANES], d = src[e + 3 * LANES];
            dst[e] = a; dst[e + LANES] = b; dst[e + 2 * LANES] = c; dst[e + 3 * LANES] = d;
        }
        for (; e < n; e += LANES) dst[e] = src[e];
    }

    SCPP_HD double *vec(int i) const { return sm + W_VEC + i * NB; }
    SCPP_HD double *xv(int i) const { return sm + W_X + i * pad2(NX); }
    SCPP_HD double *sc() const { return sm + W_SC; }
    SCPP_HD double *tile(int k) const { return sm + W_DD + (k & 1) * pad2(NX * NCP); }
    SCPP_HD double *facw() const { return sm + W_FAC; }
    SCPP_HD double *fbuf(int b) const { return sm + W_RING + b * FS; }

    // per-stage row coefficients in shared memory (chain: assembly of the node Hessian):  RCQ[r][0..2] = coefficients,
    // RCQ[r][3] = h ; RIDX[r][q] = variable index (or -1).  CST = per-instance constants (all phases).
    SCPP_HD double *rcq() const { return sm + W_RCQ; }
    SCPP_HD int *ridx() const { return reinterpret_cast<int *>(sm + W_RIDX); }
    SCPP_HD int *sup() const { return reinterpret_cast<int *>(sm + W_REV); }
    SCPP_HD double *cstw() const { return sm + W_CST; }
    SCPP_HD void cst_init() const { FOR_LANE(i, MAX_CST) cstw()[i] = cst[i]; warp_sync(); }   // all a stage-parallel pass needs of the tables
    SCPP_HD void tables_init() const
    {
        FOR_LANE(i, MAX_CST) cstw()[i] = cst[i];
        warp_sync();
        FOR_LANE(r, NROW) {
            const RowDesc rd = M::row(r);
            for (int q = 0; q < 4; q++) ridx()[r * 4 + q] = (q < rd.n) ? rd.idx[q] : -1;
            for (int q = 0; q < 3; q++) rcq()[r * 4 + q] = (q < rd.n && rd.cs[q] >= 0) ? cstw()[rd.cs[q]] : 0.;
            rcq()[r * 4 + 3] = cstw()[rd.hs];
        }
        warp_sync();
        // support (variable indices, at most 4) of each rank-1 term of the model Hessian: cones, then the multi-entry LP row
        FOR_LANE(c, NRK) {
            int n = 0, *sp = sup() + 4 * c;
            for (int q = 0; q < 4; q++) sp[q] = -1;
            const int r0 = c < NCONE ? NLP + M::cone_off(c) : 0, r1 = c < NCONE ? r0 + M::cone_dim(c) : NLP;
            for (int r = r0; r < r1; r++) {
                if (c == NCONE && ridx()[r * 4 + 1] < 0) continue;      // single-entry LP rows are diagonal terms
                for (int q = 0; q < 3; q++) {
                    const int i = ridx()[r * 4 + q];
                    bool seen = i < 0;
                    for (int t = 0; t < n; t++) seen = seen || sp[t] == i;
                    if (!seen && n < 4) sp[n++] = i;
                }
            }
        }
        warp_sync();
        if (lane_id() == 0) {      // entries of the coefficient table that take the stage's minimum-thrust direction (tables_stage)
            int n = 0, *tl = sup() + TDL;
            for (int r = 0; r < NROW; r++) {
                const RowDesc rd = M::row(r);
                for (int q = 0; q < 3; q++) if (q < rd.n && rd.cs[q] < 0 && n < TDL_MAX) { tl[1 + 2 * n] = r * 4 + q; tl[2 + 2 * n] = -rd.cs[q] - 1; n++; }
            }
            tl[0] = n;
        }
        warp_sync();
    }
    // once per stage of the chain: only the linearised minimum-thrust row depends on k (coefficient slots < 0 take -tdir[k])
    // (the list of those entries is built once by tables_init: walking the row table here cost 1.6 % of the kernel in local-memory copies
    // of RowDesc, ncu r02q)
    static constexpr int TDL = 4 * NRK, TDL_MAX = (2 * pad2(2 * NB) - TDL - 1) / 2;      // list behind sup(): count, then (rcq entry, direction component) pairs
    SCPP_HD void tables_stage(const double *td) const   // td: the stage's direction in the shared window
    {
        const int *tl = sup() + TDL;
        const int n = tl[0];
        FOR_LANE(e, n) rcq()[tl[1 + 2 * e]] = -td[tl[2 + 2 * e]];
    }
    // ---- model rows in the stage-parallel passes: the row index is a compile-time constant after unrolling, so the row table
    //      folds into immediates and the scatter targets are registers
    SCPP_HD double coef(const RowDesc &rd, int q, int k) const { return rd.cs[q] >= 0 ? cstw()[rd.cs[q]] : -tdir[3 * k + (-rd.cs[q] - 1)]; }
    SCPP_HD double row_h(const RowDesc &rd) const { return cstw()[rd.hs]; }
    // (G x)_r with x read from a stage-minor array (rows 0..NB-1 = xi)
    SCPP_HD double row_dot(int r, int k, const double *x) const
    {
        const RowDesc rd = M::crow(r);
        double a = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) if (q < rd.n) a += coef(rd, q, k) * x[rd.idx[q] * KS + k];
        return a;
    }
    // acc[idx] += coef * v  for the entries of row r
    SCPP_HD void row_scatter(int r, int k, double v, double *acc) const
    {
        const RowDesc rd = M::crow(r);
#pragma unroll
        for (int q = 0; q < 3; q++) if (q < rd.n) acc[rd.idx[q]] += coef(rd, q, k) * v;
    }

    struct Norms { double gap, rz2, rx2, pcost, zrz, xrx, h2, acc_sig; int bad; };
    SCPP_HD static double nudge(double u0, double n1) { const double thr = 4e-16 * (fabs(u0) + n1) + 1e-300; return (u0 - n1 > thr) ? u0 : n1 + thr; }

    // =============================================================================================================
    //  stage-parallel pass U : (prim, s, z) += a (dprim, ds, dz).  The step length keeps every cone 1 % inside in exact
    //  arithmetic; a cone whose margin u0 - |u1| is lost to rounding (active to ~1e-16 relative) is nudged back inside by a
    //  few ulps of u0 so the next Nesterov-Todd scaling stays defined (perturbation << the 1e-8 tolerances).
    // =============================================================================================================
    // Memory-level parallelism: the compiler may not move a load above a store to another workspace array (the arrays are
    // not provably disjoint to it), so every block of a stage-parallel pass is written  LOAD everything -> compute -> STORE:
    // dozens of independent, coalesced loads are in flight per lane instead of two, and a stage costs a handful of DRAM
    // round trips instead of a hundred (ncu: these passes were 80-96 % long-scoreboard stalls).
    template <int R0, int NR>
    SCPP_HD void upd_rows(double a, int k, double (&S)[NR], double (&Z)[NR]) const      // S,Z <- (s,z) + a (ds,dz) for rows R0..R0+NR-1
    {
        double dS[NR], dZ[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) { S[r] = s[(R0 + r) * KS + k]; dS[r] = ds[(R0 + r) * KS + k]; Z[r] = z[(R0 + r) * KS + k]; dZ[r] = dz[(R0 + r) * KS + k]; }
#pragma unroll
        for (int r = 0; r < NR; r++) { S[r] += a * dS[r]; Z[r] += a * dZ[r]; }
    }
    template <int R0, int NR>
    SCPP_HD void put_rows(int k, const double (&S)[NR], const double (&Z)[NR])
    {
#pragma unroll
        for (int r = 0; r < NR; r++) { s[(R0 + r) * KS + k] = S[r]; z[(R0 + r) * KS + k] = Z[r]; }
    }
    SCPP_HD void pass_update(double a)
    {
        FOR_STAGE(k) {
            const bool hasint = k < K - 1;
            {
                double p[PSN], d[PSN];
#pragma unroll
                for (int e = 0; e < PSN; e++) { p[e] = prim[e * KS + k]; d[e] = dprim[e * KS + k]; }
#pragma unroll
                for (int e = 0; e < PSN; e++) prim[e * KS + k] = p[e] + a * d[e];
            }
            {   // model rows: LP rows, then the model cones
                double S[NROW], Z[NROW];
                upd_rows<0, NROW>(a, k, S, Z);
#pragma unroll
                for (int r = 0; r < NLP; r++) { S[r] = nudge(S[r], 0.); Z[r] = nudge(Z[r], 0.); }
#pragma unroll
                for (int c = 0; c < NCONE; c++) {
                    const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                    double ts = 0, tz = 0;
#pragma unroll
                    for (int i = 1; i < soc::SOC_MAXD; i++) if (i < d) { ts += S[o + i] * S[o + i]; tz += Z[o + i] * Z[o + i]; }
                    S[o] = nudge(S[o], sqrt(ts)); Z[o] = nudge(Z[o], sqrt(tz));
                }
                put_rows<0, NROW>(k, S, Z);
            }
            {   // trust-region cone
                double S[D], Z[D];
                upd_rows<TRO, D>(a, k, S, Z);
                double ts = 0, tz = 0;
#pragma unroll
                for (int i = 1; i < D; i++) { ts += S[i] * S[i]; tz += Z[i] * Z[i]; }
                S[0] = nudge(S[0], sqrt(ts)); Z[0] = nudge(Z[0], sqrt(tz));
                put_rows<TRO, D>(k, S, Z);
            }
            if (hasint) {   // virtual-control pairs: s- rows, then s+ rows of the interval
                {
                    double S[NX], Z[NX];
                    upd_rows<MN, NX>(a, k, S, Z);
#pragma unroll
                    for (int r = 0; r < NX; r++) { S[r] = nudge(S[r], 0.); Z[r] = nudge(Z[r], 0.); }
                    put_rows<MN, NX>(k, S, Z);
                }
                {
                    double S[NX], Z[NX];
                    upd_rows<MN + NX, NX>(a, k, S, Z);
#pragma unroll
                    for (int r = 0; r < NX; r++) { S[r] = nudge(S[r], 0.); Z[r] = nudge(Z[r], 0.); }
                    put_rows<MN + NX, NX>(k, S, Z);
                }
            }
        }
        if (lane_id() == 0 && k_lo == 0 && !sig_fixed) {
            const int r0 = RS * KS, p0 = PSN * KS;
            double s4[4], z4[4], d4[4], e4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { s4[i] = s[r0 + i]; d4[i] = ds[r0 + i]; z4[i] = z[r0 + i]; e4[i] = dz[r0 + i]; }
            const double p0v = prim[p0], p1v = prim[p0 + 1], d0v = dprim[p0], d1v = dprim[p0 + 1];
#pragma unroll
            for (int i = 0; i < 4; i++) { s4[i] += a * d4[i]; z4[i] += a * e4[i]; }
            s4[0] = nudge(s4[0], 0.); z4[0] = nudge(z4[0], 0.);
            s4[1] = nudge(s4[1], sqrt(s4[2] * s4[2] + s4[3] * s4[3]));
            z4[1] = nudge(z4[1], sqrt(z4[2] * z4[2] + z4[3] * z4[3]));
#pragma unroll
            for (int i = 0; i < 4; i++) { s[r0 + i] = s4[i]; z[r0 + i] = z4[i]; }
            prim[p0] = p0v + a * d0v; prim[p0 + 1] = p1v + a * d1v;
        }
        warp_sync();
    }

    // =============================================================================================================
    //  stage-parallel pass R : residuals, Nesterov-Todd scaling, termination quantities
    // =============================================================================================================
    // model rows on register copies of xi (the row index is a compile-time constant after unrolling: immediates and registers);
    // td = linearised minimum-thrust direction of the stage
    SCPP_HD double coef_reg(const RowDesc &rd, int q, const double *td) const { return rd.cs[q] >= 0 ? cstw()[rd.cs[q]] : -td[-rd.cs[q] - 1]; }
    SCPP_HD double row_dot_reg(int r, const double *td, const double *P) const
    {
        const RowDesc rd = M::crow(r);
        double a = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) if (q < rd.n) a += coef_reg(rd, q, td) * P[rd.idx[q]];
        return a;
    }
    SCPP_HD void row_scatter_reg(int r, const double *td, double v, double *acc) const
    {
        const RowDesc rd = M::crow(r);
#pragma unroll
        for (int q = 0; q < 3; q++) if (q < rd.n) acc[rd.idx[q]] += coef_reg(rd, q, td) * v;
    }
    // one row of interval k for the stage-parallel passes: the tile row [A~ | C | s | z] (NC) followed by `NA` per-row scalars
    static constexpr int IR_SM = NC, IR_SP = NC + 1, IR_ZM = NC + 2, IR_ZP = NC + 3;
    SCPP_HD void ld_tile_row(int i, int k, double *row) const
    {
#pragma unroll
        for (int j = 0; j < NC; j++) row[j] = T(i, j, k);
    }

    // split == true (split pipeline): no coupling from stage k-1 and no global rows here; residual_couple / residual_globals
    // run in the next kernel; nm then holds the warp's partial sums
    SCPP_HD void pass_residuals(Norms &nm, bool identity, bool split = false)
    {
        double gap = 0, rz2 = 0, pcost = 0, zrz = 0, h2 = 0, rx2 = 0, xrx = 0, acc_sig = 0;
        int bad = 0;
        const double sg = prim[PSN * KS];
        FOR_STAGE(k) {
            const bool hasint = k < K - 1;
            double P[NB], RXa[NB], td[3];
            // ---- block 1 (loads): xi, delta, the trust-region cone, the linearisation point
            double S[D], Z[D], XB[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) { P[j] = prim[j * KS + k]; RXa[j] = 0.; }
            const double delta = prim[NB * KS + k];
#pragma unroll
            for (int i = 0; i < D; i++) { S[i] = s[(TRO + i) * KS + k]; Z[i] = z[(TRO + i) * KS + k]; }
#pragma unroll
            for (int j = 0; j < NB; j++) XB[j] = xibar(k, j);
#pragma unroll
            for (int q = 0; q < 3; q++) td[q] = tdir[3 * k + q];
            // ---- trust-region cone  (delta ; xibar - xi) in Q^{1+NB}:  S <- wb, Z <- lam, RZ
            {
                double RZ[D];
                double a = 0, b = 0, c = S[0] * Z[0];
                const double z0 = Z[0];
                RZ[0] = S[0] - delta; rz2 += RZ[0] * RZ[0]; zrz += Z[0] * RZ[0];
#pragma unroll
                for (int i = 1; i < D; i++) {
                    const double xb = XB[i - 1];
                    const double rv = tr_row(i) ? S[i] - (xb - P[i - 1]) : 0.;
                    RZ[i] = rv;
                    if (tr_row(i)) h2 += xb * xb;
                    rz2 += rv * rv; zrz += Z[i] * rv;
                    a += S[i] * S[i]; b += Z[i] * Z[i]; c += S[i] * Z[i];
                    RXa[i - 1] += Z[i];                                    // G'z of the trust-region rows
                }
                if (scvx) h2 += delta * delta;                             // the head row's h is the radius
                gap += c; pcost += w_tr * delta;
                double cev = 1.;
                const double ss = S[0] * S[0] - a, zz = Z[0] * Z[0] - b;
                const bool okc = (ss > 0.) && (zz > 0.) && (S[0] > 0.) && (Z[0] > 0.);
                if (identity || !okc) {
                    if (!identity) bad = 1;
#pragma unroll
                    for (int i = 0; i < D; i++) { S[i] = i == 0; Z[i] = i == 0; }
                } else {
                    const double sn = sqrt(ss), zn = sqrt(zz);
                    const double i2g = 1. / (2. * sqrt((1. + c / (sn * zn)) / 2.));
                    const double isn = i2g / sn, izn = i2g / zn;
                    const double w0 = S[0] * isn + Z[0] * izn;
                    double w1z1 = 0;
#pragma unroll
                    for (int i = 1; i < D; i++) { const double wi = S[i] * isn - Z[i] * izn; w1z1 += wi * Z[i]; S[i] = wi; }
                    const double eta = sqrt(sn / zn), f = Z[0] + w1z1 / (1. + w0);
                    S[0] = w0; Z[0] = eta * (w0 * Z[0] + w1z1);
#pragma unroll
                    for (int i = 1; i < D; i++) Z[i] = eta * (Z[i] + f * S[i]);
                    cev = zn / sn;
                }
#pragma unroll
                for (int i = 0; i < D; i++) { rz[(TRO + i) * KS + k] = RZ[i]; wb[(TRO + i) * KS + k] = S[i]; lam[(TRO + i) * KS + k] = Z[i]; }
                ce[NCONE * KS + k] = cev;
                const double rxdl = scvx ? 0. : w_tr - z0;                 // delta is not a variable of the SCvx problem
                rx[NB * KS + k] = rxdl;
                rx2 += rxdl * rxdl; xrx += delta * rxdl;
            }
            // ---- block 2: model rows (LP rows, model cones):  S2 <- wb, Z2 <- lam, RZ2
            {
                double S2[NROW], Z2[NROW], RZ2[NROW], CEv[NCONE];
#pragma unroll
                for (int r = 0; r < NROW; r++) { S2[r] = s[r * KS + k]; Z2[r] = z[r * KS + k]; }
#pragma unroll
                for (int c = 0; c < NCONE; c++) {
                    const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                    double sk[soc::SOC_MAXD], zk[soc::SOC_MAXD], w[soc::SOC_MAXD], lm[soc::SOC_MAXD];
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) {
                        const RowDesc rd = M::crow(o + r);
                        const double hh = row_h(rd), sl = hh - row_dot_reg(o + r, td, P);
                        sk[r] = S2[o + r]; zk[r] = Z2[o + r];
                        const double rv = sk[r] - sl;
                        RZ2[o + r] = rv;
                        h2 += hh * hh; gap += sk[r] * zk[r]; rz2 += rv * rv; zrz += zk[r] * rv;
                        row_scatter_reg(o + r, td, zk[r], RXa);
                    }
                    double e2i = 1.;
                    if (identity) {
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) { w[r] = r == 0; lm[r] = r == 0; }
                    } else if (!soc::scale(sk, zk, d, w, e2i, lm)) {
                        bad = 1;
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) { w[r] = r == 0; lm[r] = r == 0; }
                    }
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) { S2[o + r] = w[r]; Z2[o + r] = lm[r]; }
                    CEv[c] = e2i;
                }
#pragma unroll
                for (int r = 0; r < NLP; r++) {
                    const RowDesc rd = M::crow(r);
                    const double hh = row_h(rd), sl = hh - row_dot_reg(r, td, P);
                    const double sv = S2[r], zv = Z2[r];
                    const double rv = sv - sl;
                    RZ2[r] = rv;
                    h2 += hh * hh;
                    if (!(sv > 0.) || !(zv > 0.)) bad = 1;
                    S2[r] = identity ? 1. : zv / sv; Z2[r] = identity ? 1. : sqrt(sv * zv);
                    gap += sv * zv; rz2 += rv * rv; zrz += zv * rv;
                    row_scatter_reg(r, td, zv, RXa);
                }
#pragma unroll
                for (int r = 0; r < NROW; r++) { rz[r * KS + k] = RZ2[r]; wb[r * KS + k] = S2[r]; lam[r * KS + k] = Z2[r]; }
#pragma unroll
                for (int c = 0; c < NCONE; c++) ce[c * KS + k] = CEv[c];
            }
            // ---- block 3: interval k: virtual-control pairs  t_i >= |r_i| ,  r = x_{k+1} - A~ xi_k - C u_{k+1} - s sigma - z ;
            //      the loads of row i+1 are issued before the stores of row i
            if (hasint) {
                constexpr int NR = NC + 6;
                double UN[NU], cur[NR], nxt[NR];
#pragma unroll
                for (int a = 0; a < NU; a++) UN[a] = prim[(NX + a) * KS + k + 1];
                auto ld_row = [&](int i, double *row) {
                    ld_tile_row(i, k, row);
                    const int o = MN + i;
                    row[IR_SM] = s[o * KS + k]; row[IR_SP] = s[(o + NX) * KS + k]; row[IR_ZM] = z[o * KS + k]; row[IR_ZP] = z[(o + NX) * KS + k];
                    row[NC + 4] = prim[(PN + i) * KS + k]; row[NC + 5] = prim[i * KS + k + 1];
                };
                auto do_row = [&](int i, const double *cur) {
                    const int o = MN + i;
                    const double sm_ = cur[IR_SM], sp = cur[IR_SP], zm = cur[IR_ZM], zp = cur[IR_ZP];
                    const double wi = zm - zp;
                    double acc = cur[NC + 5], acc2 = 0;
#pragma unroll
                    for (int j = 0; j < NB; j += 2) {
                        const double t0 = cur[j], t1 = cur[j + 1];
                        acc -= t0 * P[j]; acc2 -= t1 * P[j + 1];
                        RXa[j] -= t0 * wi; RXa[j + 1] -= t1 * wi;
                    }
#pragma unroll
                    for (int a = 0; a < NU; a++) acc2 -= cur[NB + a] * UN[a];
                    const double tsg = cur[NB + NU], zc = cur[NB + NU + 1];
                    const double r = acc + acc2 - tsg * sg - zc;
                    acc_sig -= tsg * wi;
                    const double t = cur[NC + 4];
                    const double rm = sm_ - (t - r), rp = sp - (t + r);
                    if (!(sm_ > 0.) || !(sp > 0.) || !(zm > 0.) || !(zp > 0.)) bad = 1;
                    const double rxt = w_vc - zm - zp;
                    rz[o * KS + k] = rm; rz[(o + NX) * KS + k] = rp;
                    wb[o * KS + k] = identity ? 1. : zm / sm_; wb[(o + NX) * KS + k] = identity ? 1. : zp / sp;
                    lam[o * KS + k] = identity ? 1. : sqrt(sm_ * zm); lam[(o + NX) * KS + k] = identity ? 1. : sqrt(sp * zp);
                    rx[(PN + i) * KS + k] = rxt;
                    rx2 += rxt * rxt; xrx += t * rxt;
                    gap += sm_ * zm + sp * zp; rz2 += rm * rm + rp * rp; zrz += zm * rm + zp * rp;
                    pcost += w_vc * t;
                    h2 += 2. * zc * zc;
                };
                // two-row software pipeline (rolled: the fully unrolled form thrashes the instruction cache): the loads of the
                // next row are issued before the stores of the current one; the two buffers alternate by code position
                ld_row(0, cur);
#pragma unroll 1
                for (int i = 0; i < NX; i += 2) {
                    ld_row(i + 1, nxt);
                    do_row(i, cur);
                    if (i + 2 < NX) ld_row(i + 2, cur);
                    do_row(i + 1, nxt);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    const int o = MN + i;
                    rz[o * KS + k] = 0.; rz[(o + NX) * KS + k] = 0.; wb[o * KS + k] = 1.; wb[(o + NX) * KS + k] = 1.;
                    lam[o * KS + k] = 1.; lam[(o + NX) * KS + k] = 1.; rx[(PN + i) * KS + k] = 0.;
                }
            }
            // ---- block 4: coupling from interval k-1:  [w ; -C' w],  w = z- - z+ of that interval (two load groups)
            if (k > 0 && !split) couple_prev(k, RXa);
            const uint32_t mk = fixm[k];
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const double v = (((mk >> j) & 1u) && !split) ? 0. : RXa[j];
                rx[j * KS + k] = v;
                if (!split) { rx2 += v * v; xrx += P[j] * v; }
            }
        }
        nm.gap = warp_sum(gap); nm.rz2 = warp_sum(rz2); nm.pcost = warp_sum(pcost); nm.zrz = warp_sum(zrz);
        nm.rx2 = warp_sum(rx2); nm.xrx = warp_sum(xrx); nm.h2 = warp_sum(h2); nm.bad = warp_or(bad);
        nm.acc_sig = warp_sum(acc_sig);
        if (!split && !sig_fixed) residual_globals(nm, identity);
    }
    // RXa += [w ; -C' w] of interval k-1
    SCPP_HD void couple_prev(int k, double *RXa) const
    {
        constexpr int HX = NX / 2;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double wp[HX], Cc[HX][NU];
#pragma unroll
            for (int i = 0; i < HX; i++) {
                wp[i] = z[(MN + h * HX + i) * KS + k - 1] - z[(MN + NX + h * HX + i) * KS + k - 1];
#pragma unroll
                for (int a = 0; a < NU; a++) Cc[i][a] = T(h * HX + i, NB + a, k - 1);
            }
#pragma unroll
            for (int i = 0; i < HX; i++) {
                RXa[h * HX + i] += wp[i];
#pragma unroll
                for (int a = 0; a < NU; a++) RXa[NX + a] -= Cc[i][a] * wp[i];
            }
        }
    }
    // split pipeline: finishes rx (coupling, pinned variables) over ALL stages and adds its norms to nm (warp per instance)
    SCPP_HD void residual_couple(Norms &nm)
    {
        double rx2 = 0, xrx = 0;
        FOR_LANE(k, K) {
            double R[NB], P[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) { R[j] = rx[j * KS + k]; P[j] = prim[j * KS + k]; }
            const uint32_t mk = fixm[k];
            if (k > 0) couple_prev(k, R);
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const double v = ((mk >> j) & 1u) ? 0. : R[j];
                rx[j * KS + k] = v;
                rx2 += v * v; xrx += P[j] * v;
            }
        }
        nm.rx2 += warp_sum(rx2); nm.xrx += warp_sum(xrx);
        warp_sync();
    }
    // the four global rows (sigma >= 0.001 and the sigma trust-region cone): every lane computes the same scalars from
    // nm.acc_sig (already summed over the stages), lane 0 stores; the norms are added to nm
    SCPP_HD void residual_globals(Norms &nm, bool identity)
    {
        const int r0 = RS * KS, c0 = NCN * KS, p0 = PSN * KS;
        double s4[4], z4[4], w4[4], l4[4], r4[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { s4[i] = s[r0 + i]; z4[i] = z[r0 + i]; }
        const double sg = prim[p0], dsg = prim[p0 + 1];
        double rxs = w_time + nm.acc_sig, cev = 1.;
        int bad = 0;
        r4[0] = s4[0] - (sg - 0.001);                           // sigma >= 0.001   (SCProblem.cpp:34)
        rxs -= z4[0];
        if (!(s4[0] > 0.) || !(z4[0] > 0.)) bad = 1;
        w4[0] = identity ? 1. : z4[0] / s4[0]; l4[0] = identity ? 1. : sqrt(s4[0] * z4[0]);
        r4[1] = s4[1] - (0.5 + 0.5 * dsg);                      // ((1+dsg)/2 ; (1-dsg)/2 ; sigma - sigbar)   (:92-96)
        r4[2] = s4[2] - (0.5 - 0.5 * dsg);
        r4[3] = s4[3] - (sg - sigbar);
        rxs -= z4[3];
        const double rxd = w_trs - 0.5 * z4[1] + 0.5 * z4[2];
        if (identity) { for (int i = 0; i < 3; i++) { w4[1 + i] = i == 0; l4[1 + i] = i == 0; } }
        else if (!soc::scale(s4 + 1, z4 + 1, 3, w4 + 1, cev, l4 + 1)) { bad = 1; for (int i = 0; i < 3; i++) { w4[1 + i] = i == 0; l4[1 + i] = i == 0; } }
        if (lane_id() == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) { rz[r0 + i] = r4[i]; wb[r0 + i] = w4[i]; lam[r0 + i] = l4[i]; }
            ce[c0] = cev;
            rx[p0 + 1] = rxd; rx[p0] = rxs;
        }
        for (int r = 0; r < 4; r++) { nm.gap += s4[r] * z4[r]; nm.rz2 += r4[r] * r4[r]; nm.zrz += z4[r] * r4[r]; }
        nm.pcost += w_time * sg + w_trs * dsg;
        nm.h2 += 0.001 * 0.001 + 0.5 + sigbar * sigbar;
        nm.rx2 += rxs * rxs + rxd * rxd;
        nm.xrx += sg * rxs + dsg * rxd;
        nm.bad |= bad;
        warp_sync();
    }

    // =============================================================================================================
    //  chain phase F : assemble the reduced Hessian stage by stage and factor it (block-tridiagonal Cholesky + border)
    // =============================================================================================================
    SCPP_HD void build_model_terms(const double *WB, const double *CE, double *alpha, double *rk)
    {
        double *dg = rk + NRK * NB;
        FOR_LANE(e, 2 * NRK * NB) rk[e] = 0.;
        warp_sync();
        FOR_LANE(c, NRK) {
            double *a = rk + c * NB, *d = dg + c * NB;
            if (c < NCONE) {
                const int o = NLP + M::cone_off(c), dim = M::cone_dim(c);
                const double e2i = CE[c];
#pragma unroll 1
                for (int r = 0; r < dim; r++) {
                    const double wh = (r == 0) ? WB[o] : -WB[o + r];
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const int i = ridx()[(o + r) * 4 + q];
                        if (i >= 0) { const double cf = rcq()[(o + r) * 4 + q]; a[i] += wh * cf; d[i] += (r == 0 ? -e2i : e2i) * cf * cf; }   // -J_rr g g' (single-entry rows)
                    }
                }
                alpha[c] = 2. * e2i;
            } else {   // LP rows: single-entry rows go to the diagonal, the multi-entry row is a rank-1 term
                double al = 0.;
#pragma unroll 1
                for (int r = 0; r < NLP; r++) {
                    const double dv = WB[r];
                    if (ridx()[r * 4 + 1] < 0) { const double cf = rcq()[r * 4]; d[ridx()[r * 4]] += dv * cf * cf; }
                    else { for (int q = 0; q < 3; q++) { const int i = ridx()[r * 4 + q]; if (i >= 0) a[i] = rcq()[r * 4 + q]; } al = dv; }
                }
                alpha[c] = al;
            }
        }
        warp_sync();
    }

    // Adaptive cap on the scalings z/s of the virtual-control pairs where they enter the MATRIX of the condensed Newton system (assembly and
    // the elimination / recovery of t).  Near a sub-problem optimum every pair with nu_ki = 0 is active on both sides and z/s grows like
    // 1/mu.  When the optimal cost is zero (SCvx: the linearised dynamics can be met exactly, scpp_core/src/SCvxProblem.cpp:27) nothing else
    // bounds the reduced Hessian from below in the input directions and the block-tridiagonal Schur complements have to cancel
    // D a'a ~ 1e13 down to the O(1) curvature of the node terms: a pivot turns non-positive.  ECOS survives this case because it factors
    // the un-condensed quasi-definite KKT matrix with static regularisation delta: W^2 -> W^2 + delta, i.e. z/s -> 1/(s/z + delta).  The
    // same modification is applied here ON DEMAND: a failed factorisation is repeated with z/s capped at dcap = 1e10, 1e9, ... (DCAP_MIN),
    // and the cap stays for the rest of that sub-problem.  Right-hand sides, the primal step ds (taken from the primal Newton equation)
    // and the step lengths keep the true scalings, so the iterate stays interior and the residuals stay exact; only dz of the capped rows
    // is damped.  (Capping unconditionally doubles the iteration count of ordinary sub-problems: measured, DESIGN.md.)
    static constexpr double DCAP_FIRST = 1e10, DCAP_MIN = 1e7;
    double dcap = 0.;        // 0: no cap
#if defined(SCPP_NO_DCAP)      // A/B experiment: what the cap costs in the stage passes
    SCPP_HD double capd(double d) const { return d; }
#elif defined(SCPP_DCAP_SMEM) && defined(__CUDA_ARCH__)      // A/B experiment: the cap re-read from the shared window at every use (no long-lived register)
    SCPP_HD double capd(double d) const { const double c = sc()[24]; return (c > 0. && d > c) ? c : d; }
#else
    SCPP_HD double capd(double d) const { return (dcap > 0. && d > dcap) ? dcap : d; }
#endif
    SCPP_HD bool tighten_cap()
    {
        dcap = (dcap == 0.) ? DCAP_FIRST : dcap * 0.1;
#if defined(SCPP_DCAP_SMEM) && defined(__CUDA_ARCH__)
        if (lane_id() == 0) sc()[24] = dcap;
        warp_sync();
#endif
        return dcap >= DCAP_MIN;
    }
    // Cholesky factor of the NB x NB block H (lower triangle valid, row-major in the shared window) and its inverse Li = L^-1.
    // Device: lane i keeps row i of H in registers; column j is scaled by rsqrt of the pivot (broadcast by shuffle) and the
    // trailing update takes L[c][j] from lane c by shuffle, fully unrolled (153 DFMA).  L goes back to the window, then lane c
    // forward-substitutes column c of the inverse with uniform (broadcast) reads of L.  Host: the same arithmetic as plain loops.
    SCPP_HD_CHOL bool chol_inv(double *H, double *Li)
    {
#if defined(__CUDA_ARCH__)
        const int lane = lane_id();
        const bool own = lane < NB;
        double h[NB], invd[NB];
#pragma unroll
        for (int c = 0; c < NB; c++) h[c] = own ? H[lane * NB + c] : 0.;
        bool ok = true;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const double d = __shfl_sync(0xffffffffu, h[j], j);
            ok = ok && (d > 0.);
            const double inv = rsqrt(d > 0. ? d : 1.);
            invd[j] = inv;
            const double l = h[j] * inv;
            h[j] = l;
#pragma unroll
            for (int c = j + 1; c < NB; c++) { const double lc = __shfl_sync(0xffffffffu, l, c); h[c] = fma(-l, lc, h[c]); }
        }
        if (own) {
#pragma unroll
            for (int c = 0; c < NB; c++) H[lane * NB + c] = h[c];
        }
        __syncwarp();
        double x[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) {
            double a0 = (i == lane) ? 1. : 0., a1 = 0.;
#pragma unroll
            for (int q = 0; q < i; q += 2) { a0 = fma(-H[i * NB + q], x[q], a0); if (q + 1 < i) a1 = fma(-H[i * NB + q + 1], x[q + 1], a1); }
            x[i] = (a0 + a1) * invd[i];
        }
        if (own) {
#pragma unroll
            for (int i = 0; i < NB; i++) Li[i * NB + lane] = x[i];
        }
        return ok;
#else
        bool ok = true;
        for (int j = 0; j < NB; j++) {
            for (int i = j; i < NB; i++) { double v = H[i * NB + j]; for (int c = 0; c < j; c++) v -= H[i * NB + c] * H[j * NB + c]; H[i * NB + j] = v; }
            const double djj = H[j * NB + j];
            if (!(djj > 0.)) ok = false;
            const double inv = 1. / sqrt(djj > 0. ? djj : 1.);
            for (int i = j; i < NB; i++) H[i * NB + j] *= inv;
        }
        for (int c = 0; c < NB; c++)
            for (int i = 0; i < NB; i++) {
                if (i < c) { Li[i * NB + c] = 0.; continue; }
                double v = (i == c) ? 1. : 0.;
                for (int q = c; q < i; q++) v -= H[i * NB + q] * Li[q * NB + c];
                Li[i * NB + c] = v / H[i * NB + i];
            }
        return ok;
#endif
    }

    // H_kk, O_k = H_{k+1,k} and the sigma-border b_k of stage k BEFORE the Schur update of the chain; all buffers in the shared
    // window.  In: carry (D | D C | C' D C in Dp, border bn) of interval k-1.  Out: the same for interval k, corner += the
    // lane-partial of s_k' D s_k.
    SCPP_HD void assemble_body(int k, bool hasint, double *H, double *O, double *rk, const double *t, const double *WB, const double *CE,
                               double *alpha, double *bk, double *bn, double *dsum, double *Dp, double *Dt, double &corner)
    {
        double *DCp = Dp + NX, *CDCp = DCp + NX * NU;
        build_model_terms(WB, CE, alpha, rk);
        // ---- node part of H_kk (trust region with delta eliminated) + carry from interval k-1: dense part, then the
        //      model cones / rows, each of which touches at most 4 variables (sparse rank-1 terms, one cone at a time)
        {
            const double *dg = rk + NRK * NB;
            const double *wt_ = WB + TRO;
            const double e2i = CE[NCONE], w0 = wt_[0];
            const double kap = e2i * (2. * w0 * w0 - 1.), c2 = scvx ? 0. : 4. * e2i * e2i * w0 * w0 / kap, beta = 2. * e2i - c2;
            FOR_LANE(i, NB) {
                double v = tr_row(i + 1) ? e2i : 0.;
#pragma unroll
                for (int c = 0; c < NRK; c++) v += dg[c * NB + i];
                dsum[i] = v;
            }
            warp_sync();
#pragma unroll 2
            FOR_LANE(e, BLK) {
                const int i = e / NB, j = e - i * NB;
                double v;
                if (i < NX && j < NX) v = (i == j) ? Dp[i] : 0.;
                else if (i < NX) v = -DCp[i * NU + (j - NX)];
                else if (j < NX) v = -DCp[j * NU + (i - NX)];
                else v = CDCp[(i - NX) * NU + (j - NX)];
                if (i == j) v += dsum[i];
                H[e] = v + beta * wt_[1 + i] * wt_[1 + j];
            }
            FOR_LANE(j, NB) bk[j] = bn[j];
            warp_sync();
#pragma unroll 1
            for (int c = 0; c < NRK; c++) {
                FOR_LANE(t2, 16) {
                    const int i = sup()[4 * c + (t2 >> 2)], j = sup()[4 * c + (t2 & 3)];
                    if (i >= 0 && j >= 0) H[i * NB + j] += alpha[c] * rk[c * NB + i] * rk[c * NB + j];
                }
                warp_sync();
            }
        }
        warp_sync();
        // ---- interval k: H_kk += A~' D A~ ; O = [-D A~ ; C' D A~] ; carry = (D, D C, C' D C) ; borders
        if (hasint) {
            FOR_LANE(i, NX) { const double dm = capd(WB[MN + i]), dp = capd(WB[MN + NX + i]); Dt[i] = 4. * dm * dp / (dm + dp); }
            warp_sync();
            // O rows of the x_{k+1} block are -D A~ ; keep D A~ for the products: O[a][b], a < NX
            FOR_LANE(e, NX * NB) { const int a = e / NB, b = e - a * NB; O[a * NB + b] = -Dt[a] * t[a * NCP + b]; }
            FOR_LANE(e, NX * NU) { const int i = e / NU, j = e - i * NU; DCp[e] = Dt[i] * t[i * NCP + NB + j]; }
            FOR_LANE(i, NX) Dp[i] = Dt[i];
            warp_sync();
            // H += A~' (D A~) = -A~' O_x   and   O_u = C' D A~ = -C' O_x   on the FP64 tensor cores (lower tiles of H suffice)
            blk::mm<NB, NB, NX, true>([&](int m, int kk) { return (m < NB && kk < NX) ? t[kk * NCP + m] : 0.; },
                                      [&](int kk, int n) { return (kk < NX && n < NB) ? -O[kk * NB + n] : 0.; },
                                      [&](int m, int n, double v) { if (m < NB && n < NB) H[m * NB + n] += v; });
            blk::mm<NU, NB, NX, false>([&](int m, int kk) { return (m < NU && kk < NX) ? t[kk * NCP + NB + m] : 0.; },
                                       [&](int kk, int n) { return (kk < NX && n < NB) ? -O[kk * NB + n] : 0.; },
                                       [&](int m, int n, double v) { if (m < NU && n < NB) O[(NX + m) * NB + n] = v; });
            FOR_LANE(e, NU * NU) {
                const int a = e / NU, b = e - a * NU;
                double v = 0;
#pragma unroll 2
                for (int i = 0; i < NX; i++) v += t[i * NCP + NB + a] * DCp[i * NU + b];
                CDCp[e] = v;
            }
            FOR_LANE(j, NB) {
                double v = 0, vn;
#pragma unroll 2
                for (int i = 0; i < NX; i++) v += t[i * NCP + j] * Dt[i] * t[i * NCP + NB + NU];
                bk[j] += v;
                if (j < NX) vn = -Dt[j] * t[j * NCP + NB + NU];
                else {
                    vn = 0;
#pragma unroll 2
                    for (int i = 0; i < NX; i++) vn += t[i * NCP + NB + (j - NX)] * Dt[i] * t[i * NCP + NB + NU];
                }
                bn[j] = vn;
            }
            FOR_LANE(i, NX) { const double sv = t[i * NCP + NB + NU]; corner += Dt[i] * sv * sv; }
        } else {
            FOR_LANE(e, BLK) O[e] = 0.;
        }
        warp_sync();
        // ---- pinned variables: identity rows/columns
        {
            const uint32_t mk = fixm[k], mn = hasint ? fixm[k + 1] : 0u;
            FOR_LANE(e, BLK) {
                const int a = e / NB, b = e - a * NB;
                if (((mk >> a) & 1u) || ((mk >> b) & 1u)) H[e] = (a == b) ? 1. : 0.;
                if (((mn >> a) & 1u) || ((mk >> b) & 1u)) O[e] = 0.;
            }
            FOR_LANE(j, NB) if ((mk >> j) & 1u) bk[j] = 0.;
        }
        warp_sync();
    }
    // one step of the block Cholesky chain on stage k: Schur update with L_{k,k-1} (in Lp), factor, Linv -> Li, L_{k+1,k} -> Ln,
    // l_k -> lk.  H, O, bk are consumed.  Returns false when H is not positive definite.
    SCPP_HD bool chain_step(int k, double *H, const double *O, double *bk, const double *Lp, const double *lprev, double *Li, double *Ln,
                            double *lk, double &corner)
    {
        if (k > 0) {
            blk::mm<NB, NB, NB, true>([&](int m, int kk) { return (m < NB && kk < NB) ? Lp[m * NB + kk] : 0.; },
                                      [&](int kk, int n) { return (kk < NB && n < NB) ? Lp[n * NB + kk] : 0.; },
                                      [&](int m, int n, double v) { if (m < NB && n < NB) H[m * NB + n] -= v; });
            FOR_LANE(j, NB) {
                double v = 0;
#pragma unroll 2
                for (int c = 0; c < NB; c++) v += Lp[j * NB + c] * lprev[c];
                bk[j] -= v;
            }
        }
        warp_sync();
        const bool ok = chol_inv(H, Li);
        warp_sync();
        // ---- L_{k+1,k} = O Linv' ;  l_k = Linv bk ; corner -= l_k' l_k
        blk::mm<NB, NB, NB, false>([&](int m, int kk) { return (m < NB && kk < NB) ? O[m * NB + kk] : 0.; },
                                   [&](int kk, int n) { return (kk < NB && n < NB) ? Li[n * NB + kk] : 0.; },
                                   [&](int m, int n, double v) { if (m < NB && n < NB) Ln[m * NB + n] = v; });
        FOR_LANE(j, NB) {
            double v = 0, v2 = 0;
#pragma unroll
            for (int c = 0; c < NB; c += 2) { v += Li[j * NB + c] * bk[c]; v2 += Li[j * NB + c + 1] * bk[c + 1]; }
            v += v2;
            lk[j] = v; corner -= v * v;
        }
        warp_sync();
        return ok;
    }
    // corner of the sigma border: Schur complement of the chain + the global rows (sigma >= 0.001, sigma trust region with
    // delta_sigma eliminated); sets l_ss
    SCPP_HD bool finish_corner(double corner)
    {
        if (sig_fixed) { l_ss = 1.; return true; }             // sigma is pinned: no border
        const int r0 = RS * KS;
        const double d = wb[r0];
        double w3[3] = {wb[r0 + 1], wb[r0 + 2], wb[r0 + 3]};
        const double e2i = ce[NCN * KS];
        double g[3] = {-0.5, 0.5, 0.}, p[3], e2[3] = {0., 0., 1.}, m2[3];
        soc::Mv(w3, e2i, g, 3, p);
        const double kap = g[0] * p[0] + g[1] * p[1];
        soc::Mv(w3, e2i, e2, 3, m2);
        corner += d + (m2[2] - p[2] * p[2] / kap);
        l_ss = sqrt(corner > 0. ? corner : 1.);
        return corner > 0.;
    }

    // Inline factorisation (monolithic kernel, cold start): assembly and chain interleaved stage by stage.
    SCPP_HD bool phase_factor()
    {
        double *H = sm + W_MAT, *O = H + BLK, *Lp = sm + W_LP, *rk = sm + W_RK;
        double *Dp = sm + W_HN;             // carry of interval k-1: D | D C | C' D C
        double *F = facw();                 // Linv | Lnext | l   (the record written for this stage)
        double *Li = F, *Ln = F + OFF_LN, *lk = F + OFF_L;
        double *WB = sm + W_WB, *CE = sc() + 8, *alpha = sc();
        double *bk = vec(0), *bn = vec(1), *lprev = vec(2);
        double *Dt = xv(1);
        double corner = 0.;
        int bad = 0;
        FOR_LANE(e, BLK) Lp[e] = 0.;
        FOR_LANE(e, HNC) Dp[e] = 0.;
        FOR_LANE(j, NB) { bn[j] = 0.; lprev[j] = 0.; }
        warp_sync();
        // prefetch: the tile of interval k+1 is copied while stage k is factored; the scaling rows, cone scalars and the
        // minimum-thrust direction of stage k+1 wait in registers (strided global loads issued one stage ahead)
        double *TD = sc() + 16;
        if (K > 1) ld_dd(0);
        ld_commit();
#if defined(__CUDA_ARCH__)
        constexpr int WBN = (RS + LANES - 1) / LANES;
        double wbn[WBN], cen = 0., tdn = 0.;
        {
            const int l = lane_id();
#pragma unroll
            for (int i = 0; i < WBN; i++) wbn[i] = (l + LANES * i < RS) ? wb[(l + LANES * i) * KS] : 0.;
            if (l < NCN) cen = ce[l * KS];
            if (l < 3) tdn = tdir[l];
        }
#endif
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            if (k + 1 < K - 1) ld_dd(k + 1);
            ld_commit();
#if defined(__CUDA_ARCH__)
            {
                const int l = lane_id();
#pragma unroll
                for (int i = 0; i < WBN; i++) if (l + LANES * i < RS) WB[l + LANES * i] = wbn[i];
                if (l < NCN) CE[l] = cen;
                if (l < 3) TD[l] = tdn;
                if (k + 1 < K) {
#pragma unroll
                    for (int i = 0; i < WBN; i++) wbn[i] = (l + LANES * i < RS) ? wb[(l + LANES * i) * KS + k + 1] : 0.;
                    if (l < NCN) cen = ce[l * KS + k + 1];
                    if (l < 3) tdn = tdir[3 * (k + 1) + l];
                }
            }
#else
            FOR_LANE(r, RS) WB[r] = wb[r * KS + k];
            FOR_LANE(c, NCN) CE[c] = ce[c * KS + k];
            FOR_LANE(q, 3) TD[q] = tdir[3 * k + q];
#endif
            warp_sync();
            tables_stage(TD);
            ld_wait(1);
            assemble_body(k, hasint, H, O, rk, tile(k), WB, CE, alpha, bk, bn, vec(3), Dp, Dt, corner);
            if (!chain_step(k, H, O, bk, Lp, lprev, Li, Ln, lk, corner)) bad = 1;
            st(fac + (size_t)k * FS, F, FS);
            FOR_LANE(e, BLK) Lp[e] = Ln[e];
            FOR_LANE(j, NB) lprev[j] = lk[j];
            warp_sync();
        }
        corner = warp_sum(corner);
        if (!finish_corner(corner)) bad = 1;
        return !warp_or(bad);
    }

    // Split pipeline, step A (one warp per (instance, stage)): assemble H_kk | O_k | b_k of stage k into its factor record and
    // the corner contribution of interval k into cpart[k].  Needs tables_init() on this warp's window.
    SCPP_HD void assemble_stage(int k)
    {
        const bool hasint = k < K - 1;
        double *t = sm + AW_DD, *PC = sm + AW_PC, *H = sm + AW_MAT, *O = H + BLK, *rk = sm + AW_RK;
        double *Cp = PC, *sp = PC + NX * NU, *WBp = sp + NX;      // interval k-1: C | s column | scaling rows of its pairs
        double *Dp = sm + W_HN, *DCp = Dp + NX, *CDCp = DCp + NX * NU;
        double *WB = sm + W_WB, *CE = sc() + 8, *alpha = sc(), *TD = sc() + 16;
        double *bk = vec(0), *bn = vec(1), *Dt = xv(1);
        // ---- loads: tile k (asynchronous), then the strided pieces
        if (hasint) {
            const double *src = dd + (size_t)k * NX * NC;
#if defined(__CUDA_ARCH__)
            const unsigned d0 = (unsigned)__cvta_generic_to_shared(t);
            for (int c = lane_id(); c < NX * (NC / 2); c += LANES) {
                const int r = c / (NC / 2), q = c - r * (NC / 2);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 8u * (r * NCP + 2 * q)), "l"(src + r * NC + 2 * q) : "memory");
            }
#else
            for (int r = 0; r < NX; r++) memcpy(t + r * NCP, src + r * NC, sizeof(double) * NC);
#endif
        }
        ld_commit();
        FOR_LANE(r, RS) WB[r] = wb[r * KS + k];
        FOR_LANE(c, NCN) CE[c] = ce[c * KS + k];
        FOR_LANE(q, 3) TD[q] = tdir[3 * k + q];
        if (k > 0) {
            const double *tp = dd + (size_t)(k - 1) * NX * NC;
            FOR_LANE(e, NX * NU) { const int i = e / NU, j = e - i * NU; Cp[e] = tp[i * NC + NB + j]; }
            FOR_LANE(i, NX) sp[i] = tp[i * NC + NB + NU];
            FOR_LANE(r, 2 * NX) WBp[r] = wb[(MN + r) * KS + k - 1];
        }
        warp_sync();
        tables_stage(TD);
        // ---- carry of interval k-1
        if (k > 0) {
            FOR_LANE(i, NX) { const double dm = capd(WBp[i]), dp = capd(WBp[NX + i]); Dp[i] = 4. * dm * dp / (dm + dp); }
            warp_sync();
            FOR_LANE(e, NX * NU) DCp[e] = Dp[e / NU] * Cp[e];
            warp_sync();
            FOR_LANE(e, NU * NU) {
                const int a = e / NU, b = e - a * NU;
                double v = 0;
#pragma unroll 2
                for (int i = 0; i < NX; i++) v += Cp[i * NU + a] * DCp[i * NU + b];
                CDCp[e] = v;
            }
            FOR_LANE(j, NB) {
                double vn;
                if (j < NX) vn = -Dp[j] * sp[j];
                else {
                    vn = 0;
#pragma unroll 2
                    for (int i = 0; i < NX; i++) vn += Cp[i * NU + (j - NX)] * Dp[i] * sp[i];
                }
                bn[j] = vn;
            }
        } else {
            FOR_LANE(e, HNC) Dp[e] = 0.;
            FOR_LANE(j, NB) bn[j] = 0.;
        }
        ld_wait();
        double corner = 0.;
        assemble_body(k, hasint, H, O, rk, t, WB, CE, alpha, bk, bn, vec(3), Dp, Dt, corner);
        corner = warp_sum(corner);
        double *rec = fac + (size_t)k * FS;
        st(rec, H, BLK); st(rec + OFF_LN, O, BLK);
        FOR_LANE(j, NB) rec[OFF_L + j] = bk[j];
        if (lane_id() == 0) cpart[k] = corner;
        warp_sync();
    }
    // Split pipeline, step F (one warp per instance): the chain over the assembled records, in place (H|O|b -> Linv|Ln|l),
    // three records in flight.
    SCPP_HD bool chain_factor()
    {
        double *Lp = sm + W_CLP, *Li = sm + W_CLI, *lprev = vec(2), *lk = vec(3);
        double corner = 0.;
        int bad = 0;
        FOR_LANE(k, K) corner += cpart[k];
        FOR_LANE(j, NB) lprev[j] = 0.;
        ld(fbuf(0), fac, FS); ld_commit();
        if (K > 1) ld(fbuf(1), fac + FS, FS);
        ld_commit();
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            if (k + 2 < K) ld(fbuf((k + 2) % 3), fac + (size_t)(k + 2) * FS, FS);
            ld_commit();
            ld_wait(2);
            double *F = fbuf(k % 3);
            // L_{k+1,k} goes straight into Lp (free once the Schur update of this stage is done): it is the next stage's L_{k,k-1}
            double *rec = fac + (size_t)k * FS;
            if (!chain_step_inplace(k, F, F + OFF_LN, F + OFF_L, Lp, lprev, Li, lk, corner)) bad = 1;
            st(rec, Li, BLK); st(rec + OFF_LN, Lp, BLK);
            FOR_LANE(j, NB) { rec[OFF_L + j] = lk[j]; lprev[j] = lk[j]; }
            warp_sync();
        }
        ld_wait();
        corner = warp_sum(corner);
        if (!finish_corner(corner)) bad = 1;
        return !warp_or(bad);
    }
    // chain_step with L_{k+1,k} written over Lp: the Schur update reads Lp first, the product O Linv' is formed in registers
    // and stored after a warp barrier
    SCPP_HD bool chain_step_inplace(int k, double *H, const double *O, double *bk, double *Lp, const double *lprev, double *Li, double *lk,
                                    double &corner)
    {
        if (k > 0) {
            blk::mm<NB, NB, NB, true>([&](int m, int kk) { return (m < NB && kk < NB) ? Lp[m * NB + kk] : 0.; },
                                      [&](int kk, int n) { return (kk < NB && n < NB) ? Lp[n * NB + kk] : 0.; },
                                      [&](int m, int n, double v) { if (m < NB && n < NB) H[m * NB + n] -= v; });
            FOR_LANE(j, NB) {
                double v = 0;
#pragma unroll 2
                for (int c = 0; c < NB; c++) v += Lp[j * NB + c] * lprev[c];
                bk[j] -= v;
            }
        }
        warp_sync();
        const bool ok = chol_inv(H, Li);
        warp_sync();
        blk::mm<NB, NB, NB, false>([&](int m, int kk) { return (m < NB && kk < NB) ? O[m * NB + kk] : 0.; },
                                   [&](int kk, int n) { return (kk < NB && n < NB) ? Li[n * NB + kk] : 0.; },
                                   [&](int m, int n, double v) { if (m < NB && n < NB) Lp[m * NB + n] = v; });
        FOR_LANE(j, NB) {
            double v = 0, v2 = 0;
#pragma unroll
            for (int c = 0; c < NB; c += 2) { v += Li[j * NB + c] * bk[c]; v2 += Li[j * NB + c + 1] * bk[c + 1]; }
            v += v2;
            lk[j] = v; corner -= v * v;
        }
        warp_sync();
        return ok;
    }

    // =============================================================================================================
    //  Phase S : solve the Newton system  G'dz = rxv ,  G dx - W^2 dz = rzv  with the local variables eliminated.
    //     mode 0: rxv = dprim (array), rzv = ds (array)                        [starting point]
    //     mode 1: rxv = -rx,            rzv = -rz + s                          [affine direction]
    //     mode 2: rxv = -(1-sig) rx,    rzv = -(1-sig) rz - W (lam \ d_s),  d_s = -lam o lam - cr + sig mu e   [combined]
    //  pass A (stage-parallel): right-hand sides rzv (kept in ds) and g ; chain: forward and backward substitution ;
    //  pass B (stage-parallel): recovery of the local variables, dz and ds = rzs*rz - G dx, the scaled step lengths (tmax)
    //  and, for mode 1, cr = ds~ o dz~.
    // =============================================================================================================
    SCPP_HD double rxv_of(int mode, double csig, int e, int k) const
    {
        return mode == 0 ? dprim[e * KS + k] : (mode == 1 ? -rx[e * KS + k] : -csig * rx[e * KS + k]);
    }

    // rzv of rows R0..R0+NR-1 in scaled-residual form needs, per mode:  0: ds   1: rz, s   2: lam, cr, rz  -> in0, in1, in2
    template <int NR>
    SCPP_HD void ld_rhs_rows(int mode, int r0, int k, double *in0, double *in1, double *in2) const
    {
        if (mode == 0) {
#pragma unroll
            for (int r = 0; r < NR; r++) in0[r] = ds[(r0 + r) * KS + k];
        } else if (mode == 1) {
#pragma unroll
            for (int r = 0; r < NR; r++) { in0[r] = rz[(r0 + r) * KS + k]; in1[r] = s[(r0 + r) * KS + k]; }
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++) { in0[r] = lam[(r0 + r) * KS + k]; in1[r] = cr[(r0 + r) * KS + k]; in2[r] = rz[(r0 + r) * KS + k]; }
        }
    }
    // LP row: rzv from the loaded inputs (dv = W^-2 of the row)
    SCPP_HD static double lp_rzv(int mode, double csig, double sigmu, double dv, double i0, double i1, double i2)
    {
        if (mode == 0) return i0;
        if (mode == 1) return -i0 + i1;
        return -csig * i2 - sqrt(1. / dv) * ((-i0 * i0 - i1 + sigmu) / i0);
    }

    SCPP_HD_PASS double pass_rhs(int mode, double csig, double sigmu, bool split = false)   // returns the lane-partial of the sigma right-hand side
    {
        double gsig = 0;
        FOR_STAGE(k) {
            const bool hasint = k < K - 1;
            double G[NB], td[3];
            // ---- block 1 (loads): rxv of xi and delta, the trust-region cone
            double w[D], q[D], i1[D], i2[D];
#pragma unroll
            for (int j = 0; j < NB; j++) G[j] = rxv_of(mode, csig, j, k);
            const double rxd = rxv_of(mode, csig, NB, k);
            const double e2i = ce[NCONE * KS + k];
#pragma unroll
            for (int i = 0; i < D; i++) w[i] = wb[(TRO + i) * KS + k];
            ld_rhs_rows<D>(mode, TRO, k, q, i1, i2);
#pragma unroll
            for (int t = 0; t < 3; t++) td[t] = tdir[3 * k + t];
            // ---- trust region: rzv, then v = M rzv - p (p'rzv + rx_delta)/kap ,  p = M(-e0)
            {
                if (mode == 1) {
#pragma unroll
                    for (int i = 0; i < D; i++) q[i] = -q[i] + i1[i];
                } else if (mode == 2) {      // q = lam, i1 = cr, i2 = rz
                    double ll = 0;
#pragma unroll
                    for (int i = 0; i < D; i++) ll += q[i] * q[i];
                    const double l0 = q[0], den = 2. * l0 * l0 - ll;
                    double l1d1 = 0;
#pragma unroll
                    for (int i = 1; i < D; i++) { const double dv = -2. * l0 * q[i] - i1[i]; i1[i] = dv; l1d1 += q[i] * dv; }
                    const double dv0 = -ll - i1[0] + sigmu;
                    const double x0 = (l0 * dv0 - l1d1) / den, il0 = 1. / l0;
                    double w1v1 = 0;
#pragma unroll
                    for (int i = 1; i < D; i++) { i1[i] = (i1[i] - x0 * q[i]) * il0; w1v1 += w[i] * i1[i]; }
                    // W (lam \ d_s)
                    const double eta = 1. / sqrt(e2i), f = x0 + w1v1 / (1. + w[0]);
                    q[0] = -csig * i2[0] - eta * (w[0] * x0 + w1v1);
#pragma unroll
                    for (int i = 1; i < D; i++) q[i] = -csig * i2[i] - eta * (i1[i] + f * w[i]);
                }
                if (mode != 0) {
#pragma unroll
                    for (int i = 0; i < D; i++) ds[(TRO + i) * KS + k] = q[i];
                }
                const double w0 = w[0], kap = e2i * (2. * w0 * w0 - 1.);
                double dot = w0 * q[0], prz = -kap * q[0];
#pragma unroll
                for (int i = 1; i < D; i++) { dot -= w[i] * q[i]; prz += 2. * e2i * w0 * w[i] * q[i]; }
                const double rho = scvx ? 0. : (prz + rxd) / kap;
#pragma unroll
                for (int i = 1; i < D; i++) G[i - 1] += e2i * (-2. * dot * w[i] + q[i]) - 2. * e2i * w0 * w[i] * rho;
            }
            // ---- block 2: model rows (cones, LP rows)
            {
                double W2[NROW], a0[NROW], a1[NROW], a2[NROW], CEv[NCONE];
#pragma unroll
                for (int r = 0; r < NROW; r++) W2[r] = wb[r * KS + k];
#pragma unroll
                for (int c = 0; c < NCONE; c++) CEv[c] = ce[c * KS + k];
                ld_rhs_rows<NROW>(mode, 0, k, a0, a1, a2);
#pragma unroll
                for (int c = 0; c < NCONE; c++) {
                    const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                    double wc[soc::SOC_MAXD], t1[soc::SOC_MAXD];
                    const double e2c = CEv[c];
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) wc[r] = W2[o + r];
                    if (mode == 0) {
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) t1[r] = a0[o + r];
                    } else if (mode == 1) {
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) t1[r] = -a0[o + r] + a1[o + r];
                    } else {
                        double lm[soc::SOC_MAXD];
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) lm[r] = a0[o + r];
                        soc::jprod(lm, lm, d, t1);
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) t1[r] = -t1[r] - a1[o + r];
                        t1[0] += sigmu;
                        soc::jdiv(lm, t1, d, t1);
                        soc::Wv(wc, e2c, t1, d, t1, false);
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) t1[r] = -csig * a2[o + r] - t1[r];
                    }
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) a0[o + r] = t1[r];          // rzv (kept for the store below)
                    soc::Mv(wc, e2c, t1, d, t1);
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) row_scatter_reg(o + r, td, t1[r], G);
                }
#pragma unroll
                for (int r = 0; r < NLP; r++) {
                    const double dv = W2[r];
                    const double rzv = lp_rzv(mode, csig, sigmu, dv, a0[r], a1[r], a2[r]);
                    a0[r] = rzv;
                    row_scatter_reg(r, td, dv * rzv, G);
                }
                if (mode != 0) {
#pragma unroll
                    for (int r = 0; r < NROW; r++) ds[r * KS + k] = a0[r];
                }
            }
            // ---- block 3: interval pairs: t eliminated, w_i couples to nodes k and k+1 (row i+1 is loaded before row i is stored)
            if (hasint) {
                constexpr int NR = NB + 1 + 9;       // A~ row | s column | dm dp | rxv | six mode inputs
                double cur[NR], nxt[NR];
                auto ld_row = [&](int i, double *row) {
                    const int o = MN + i;
#pragma unroll
                    for (int j = 0; j < NB; j++) row[j] = T(i, j, k);
                    row[NB] = T(i, NB + NU, k);
                    row[NB + 1] = wb[o * KS + k]; row[NB + 2] = wb[(o + NX) * KS + k];
                    row[NB + 3] = rxv_of(mode, csig, PN + i, k);
                    if (mode == 0) { row[NB + 4] = ds[o * KS + k]; row[NB + 5] = ds[(o + NX) * KS + k]; }
                    else if (mode == 1) { row[NB + 4] = rz[o * KS + k]; row[NB + 5] = rz[(o + NX) * KS + k]; row[NB + 6] = s[o * KS + k]; row[NB + 7] = s[(o + NX) * KS + k]; }
                    else {
                        row[NB + 4] = lam[o * KS + k]; row[NB + 5] = lam[(o + NX) * KS + k]; row[NB + 6] = cr[o * KS + k]; row[NB + 7] = cr[(o + NX) * KS + k];
                        row[NB + 8] = rz[o * KS + k]; row[NB + 9] = rz[(o + NX) * KS + k];
                    }
                };
                auto do_row = [&](int i, const double *cur) {
                    const int o = MN + i;
                    const double rm = lp_rzv(mode, csig, sigmu, cur[NB + 1], cur[NB + 4], cur[NB + 6], cur[NB + 8]);
                    const double rp = lp_rzv(mode, csig, sigmu, cur[NB + 2], cur[NB + 5], cur[NB + 7], cur[NB + 9]);
                    const double dm = capd(cur[NB + 1]), dp = capd(cur[NB + 2]);      // matrix side: capped (see dcap)
                    if (mode != 0) { ds[o * KS + k] = rm; ds[(o + NX) * KS + k] = rp; }
                    const double rho = (-(dm * rm + dp * rp) + cur[NB + 3]) / (dm + dp);
                    const double wi = dm * (rm + rho) - dp * (rp + rho);
                    wv[i * KS + k] = wi;
#pragma unroll
                    for (int j = 0; j < NB; j++) G[j] -= cur[j] * wi;
                    gsig -= cur[NB] * wi;
                };
                ld_row(0, cur);
#pragma unroll 1
                for (int i = 0; i < NX; i += 2) {
                    ld_row(i + 1, nxt);
                    do_row(i, cur);
                    if (i + 2 < NX) ld_row(i + 2, cur);
                    do_row(i + 1, nxt);
                }
            } else if (mode != 0) {
#pragma unroll
                for (int i = 0; i < NX; i++) { ds[(MN + i) * KS + k] = 0.; ds[(MN + NX + i) * KS + k] = 0.; }
            }
#pragma unroll
            for (int j = 0; j < NB; j++) gv[j * KS + k] = G[j];
        }
        warp_sync();
        if (!split) rhs_couple();
        return gsig;
    }
    // coupling from interval k-1 and the pinned variables, over ALL stages (split pipeline: first step of the chain kernel)
    SCPP_HD void rhs_couple()
    {
        FOR_LANE(k, K) {
            const uint32_t mk = fixm[k];
            double g[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) g[j] = gv[j * KS + k];
            if (k > 0) {
                double wp[NX], Cc[NX][NU];
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    wp[i] = wv[i * KS + k - 1];
#pragma unroll
                    for (int a = 0; a < NU; a++) Cc[i][a] = T(i, NB + a, k - 1);
                }
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    g[i] += wp[i];
#pragma unroll
                    for (int a = 0; a < NU; a++) g[NX + a] -= Cc[i][a] * wp[i];
                }
            }
#pragma unroll
            for (int j = 0; j < NB; j++) gv[j * KS + k] = ((mk >> j) & 1u) ? 0. : g[j];
        }
        warp_sync();
    }

    // forward substitution  f_k = Linv_k (g_k - L_{k,k-1} f_{k-1})  in place in gv ; returns the lane-partial of  sum_k l_k' f_k
    SCPP_HD void ld_col(double *dst, const double *src, int ks) const   // one stage column of a stage-minor [NB][KS] array: NB 8-byte asynchronous copies
    {
#if defined(__CUDA_ARCH__)
        const int j = lane_id();
        if (j < NB) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst + j)), "l"(src + (size_t)j * ks) : "memory");
#else
        for (int j = 0; j < NB; j++) dst[j] = src[(size_t)j * ks];
#endif
    }
    SCPP_HD double chain_forward()
    {
        // the solver object lives in local memory and every asynchronous copy is a compiler memory barrier: members read inside the loop
        // are re-loaded from the stack in every stage (ncu r02i: 10 % of the kernel's stall samples waited for `fac` to come back before
        // the copy addresses could be formed) -> loop-invariant members are taken into registers here
        const double *const fac_ = fac;
        double *const gv_ = gv;
        const int K_ = K, KS_ = KS;
        double *fcur = vec(0), *tmp = vec(1);
        static_assert(RING == 3 || RING == 4, "ring of three or four records");
        constexpr int R = RING, AH = RING - 1;      // AH records ahead of the stage being processed
        double *const v0 = vec(2), *const v1 = vec(3), *const v2 = vec(4), *const v3 = vec(5);
        double *const fb0 = fbuf(0), *const fb1 = fbuf(1), *const fb2 = fbuf(2), *const fb3 = fbuf(R - 1);
        auto vecr = [&](int i) { return i == 0 ? v0 : (i == 1 ? v1 : (i == 2 ? v2 : v3)); };
        auto fbr = [&](int i) { return i == 0 ? fb0 : (i == 1 ? fb1 : (i == 2 ? fb2 : fb3)); };
        double ldot = 0;
        // group of record r carries the right-hand side of stage r+1 (needed at the end of stage r)
#pragma unroll
        for (int r = 0; r < AH; r++) {
            if (r < K_) { ld(fbr(r), fac_ + (size_t)r * FS, FS); if (r + 1 < K_) ld_col(vecr((r + 1) % R), gv_ + r + 1, KS_); }
            ld_commit();
        }
        FOR_LANE(jj, NB) tmp[jj] = gv_[jj * KS_];
#pragma unroll 1
        for (int k = 0; k < K_; k++) {
            if (k + AH < K_) { ld(fbr((k + AH) % R), fac_ + (size_t)(k + AH) * FS, FS); if (k + AH + 1 < K_) ld_col(vecr((k + AH + 1) % R), gv_ + k + AH + 1, KS_); }
            ld_commit();
            ld_wait(AH);
            const double *F = fbr(k % R), *Ln = F + OFF_LN, *gn = vecr((k + 1) % R);
            FOR_LANE(jj, NB) {
                double v = 0, v2_ = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { v += F[jj * NB + c] * tmp[c]; v2_ += F[jj * NB + c + 1] * tmp[c + 1]; }
                v += v2_;
                fcur[jj] = v;
                gv_[jj * KS_ + k] = v;
                ldot += F[OFF_L + jj] * v;
            }
            warp_sync();
            if (k + 1 < K_) {
                FOR_LANE(jj, NB) {
                    double v = gn[jj], v2_ = 0;
#pragma unroll
                    for (int c = 0; c < NB; c += 2) { v -= Ln[jj * NB + c] * fcur[c]; v2_ -= Ln[jj * NB + c + 1] * fcur[c + 1]; }
                    v += v2_;
                    tmp[jj] = v;
                }
            }
            warp_sync();
        }
        ld_wait();
        return ldot;
    }
    // back substitution  y_k = Linv_k' (f_k - L_{k+1,k}' y_{k+1} - l_k y_sigma)  in place in gv
    SCPP_HD void chain_backward(double ysig)
    {
        // loop-invariant members in registers (see chain_forward); the pinned-variable mask of the NEXT stage is loaded one stage ahead
        // (ncu r02i: 4 % of the stall samples waited for fixm[k] at its first use)
        const double *const fac_ = fac;
        double *const gv_ = gv;
        const uint32_t *const fixm_ = fixm;
        const int K_ = K, KS_ = KS;
        double *ynext = vec(0), *tmp = vec(1);
        constexpr int R = RING, AH = RING - 1;
        double *const v0 = vec(2), *const v1 = vec(3), *const v2 = vec(4), *const v3 = vec(5);
        double *const fb0 = fbuf(0), *const fb1 = fbuf(1), *const fb2 = fbuf(2), *const fb3 = fbuf(R - 1);
        auto vecr = [&](int i) { return i == 0 ? v0 : (i == 1 ? v1 : (i == 2 ? v2 : v3)); };
        auto fbr = [&](int i) { return i == 0 ? fb0 : (i == 1 ? fb1 : (i == 2 ? fb2 : fb3)); };
#pragma unroll
        for (int r = 1; r <= AH; r++) {
            if (K_ - r >= 0) { ld(fbr((K_ - r) % R), fac_ + (size_t)(K_ - r) * FS, FS); ld_col(vecr((K_ - r) % R), gv_ + K_ - r, KS_); }
            ld_commit();
        }
        uint32_t mk_next = fixm_[K_ - 1];
#pragma unroll 1
        for (int k = K_ - 1; k >= 0; k--) {
            const bool hasint = k < K_ - 1;
            if (k >= AH) { ld(fbr((k - AH) % R), fac_ + (size_t)(k - AH) * FS, FS); ld_col(vecr((k - AH) % R), gv_ + k - AH, KS_); }
            ld_commit();
            const uint32_t mk = mk_next;
            if (k > 0) mk_next = fixm_[k - 1];
            ld_wait(AH);
            const double *F = fbr(k % R), *fk = vecr(k % R);
            FOR_LANE(jj, NB) {
                double v = fk[jj] - F[OFF_L + jj] * ysig;
                if (hasint) {
                    double v2_ = 0;
#pragma unroll
                    for (int c = 0; c < NB; c += 2) { v -= F[OFF_LN + c * NB + jj] * ynext[c]; v2_ -= F[OFF_LN + (c + 1) * NB + jj] * ynext[c + 1]; }
                    v += v2_;
                }
                tmp[jj] = v;
            }
            warp_sync();
            FOR_LANE(jj, NB) {
                double v = 0, v2_ = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { v += F[c * NB + jj] * tmp[c]; v2_ += F[(c + 1) * NB + jj] * tmp[c + 1]; }
                v += v2_;
                if ((mk >> jj) & 1u) v = 0.;
                ynext[jj] = v;
                gv_[jj * KS_ + k] = v;
            }
            warp_sync();
        }
        ld_wait();
    }

    SCPP_HD_PASS double pass_recover(int mode, double csig, double rzs, double ysig)   // returns the lane-partial of tmax
    {
        double tmax = 0;
        FOR_STAGE(k) {
            const bool hasint = k < K - 1;
            double yk[NB], td[3];
            // ---- block 1 (loads): the solution of the stage, the trust-region cone
            double w[D], q[D], RZ[D], LM[D];
#pragma unroll
            for (int j = 0; j < NB; j++) yk[j] = gv[j * KS + k];
            const double e2i = ce[NCONE * KS + k];
            const double rxd = rxv_of(mode, csig, NB, k);
#pragma unroll
            for (int i = 0; i < D; i++) { w[i] = wb[(TRO + i) * KS + k]; q[i] = ds[(TRO + i) * KS + k]; }
            if (mode != 0) {
#pragma unroll
                for (int i = 0; i < D; i++) { RZ[i] = rz[(TRO + i) * KS + k]; LM[i] = lam[(TRO + i) * KS + k]; }
            }
#pragma unroll
            for (int t = 0; t < 3; t++) td[t] = tdir[3 * k + t];
#pragma unroll
            for (int j = 0; j < NB; j++) dprim[j * KS + k] = yk[j];
            // ---- trust region
            {
                const double w0 = w[0], kap = e2i * (2. * w0 * w0 - 1.);
                q[0] = -q[0];
                double pq = -kap * q[0], dot = w0 * q[0];
#pragma unroll
                for (int i = 1; i < D; i++) { q[i] = tr_row(i) ? yk[i - 1] - q[i] : 0.; pq += 2. * e2i * w0 * w[i] * q[i]; dot -= w[i] * q[i]; }
                const double ddl = scvx ? 0. : (rxd - pq) / kap;
                dprim[NB * KS + k] = ddl;
                // dz = M q + p ddl   (in q)
                q[0] = e2i * (2. * dot * w0 - q[0]) - kap * ddl;
#pragma unroll
                for (int i = 1; i < D; i++) q[i] = e2i * (-2. * dot * w[i] + q[i]) + 2. * e2i * w0 * w[i] * ddl;
#pragma unroll
                for (int i = 0; i < D; i++) dz[(TRO + i) * KS + k] = q[i];
                if (mode != 0) {
                    double dsv[D];
                    dsv[0] = rzs * RZ[0] + ddl;
                    double w1z = 0, w1s = 0;
#pragma unroll
                    for (int i = 1; i < D; i++) { dsv[i] = tr_row(i) ? rzs * RZ[i] - yk[i - 1] : 0.; w1z += w[i] * q[i]; w1s += w[i] * dsv[i]; }
#pragma unroll
                    for (int i = 0; i < D; i++) ds[(TRO + i) * KS + k] = dsv[i];
                    // scaled directions  dz~ = W dz (in q),  ds~ = W^-1 ds (in dsv)
                    const double eta = 1. / sqrt(e2i), ieta = 1. / eta;
                    const double fz = q[0] + w1z / (1. + w0), fs = -dsv[0] + w1s / (1. + w0);
                    const double z0 = eta * (w0 * q[0] + w1z), s0 = ieta * (w0 * dsv[0] - w1s);
                    double l1 = 0, a1 = 0, a2 = 0, cr0 = s0 * z0;
                    const double lm0 = LM[0];
#pragma unroll
                    for (int i = 1; i < D; i++) {
                        const double lmi = LM[i];
                        q[i] = eta * (q[i] + fz * w[i]); dsv[i] = ieta * (dsv[i] + fs * w[i]);
                        l1 += lmi * lmi; a1 += lmi * dsv[i]; a2 += lmi * q[i];
                        cr0 += dsv[i] * q[i];
                    }
                    const double ia = 1. / sqrt(lm0 * lm0 - l1), l0 = lm0 * ia;
                    const double ld1 = l0 * s0 - a1 * ia, ld2 = l0 * z0 - a2 * ia;
                    const double il = ia / (l0 + 1.);
                    const double f1 = (ld1 + s0) * il, f2 = (ld2 + z0) * il;
                    double n1 = 0, n2 = 0;
#pragma unroll
                    for (int i = 1; i < D; i++) { const double r1 = dsv[i] - f1 * LM[i], r2 = q[i] - f2 * LM[i]; n1 += r1 * r1; n2 += r2 * r2; }
                    tmax = fmax(tmax, fmax((sqrt(n1) - ld1) * ia, (sqrt(n2) - ld2) * ia));
                    if (mode == 1) {
                        cr[TRO * KS + k] = cr0;
#pragma unroll
                        for (int i = 1; i < D; i++) cr[(TRO + i) * KS + k] = s0 * q[i] + z0 * dsv[i];
                    }
                }
            }
            // ---- block 2: model rows (cones, LP rows): DZ, DS, CR are stored at the end of the block
            {
                double W2[NROW], D2[NROW], R2[NROW], L2[NROW], CEv[NCONE], DZ[NROW], CRv[NROW];
#pragma unroll
                for (int r = 0; r < NROW; r++) { W2[r] = wb[r * KS + k]; D2[r] = ds[r * KS + k]; }
#pragma unroll
                for (int c = 0; c < NCONE; c++) CEv[c] = ce[c * KS + k];
                if (mode != 0) {
#pragma unroll
                    for (int r = 0; r < NROW; r++) { R2[r] = rz[r * KS + k]; L2[r] = lam[r * KS + k]; }
                }
#pragma unroll
                for (int c = 0; c < NCONE; c++) {
                    const int o = NLP + M::cone_off(c), d = M::cone_dim(c);
                    double wc[soc::SOC_MAXD], qc[soc::SOC_MAXD], gdx[soc::SOC_MAXD], lm[soc::SOC_MAXD];
                    const double e2c = CEv[c];
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) {
                        wc[r] = W2[o + r];
                        gdx[r] = row_dot_reg(o + r, td, yk);
                        qc[r] = gdx[r] - D2[o + r];
                    }
                    soc::Mv(wc, e2c, qc, d, qc);                                   // dz
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) DZ[o + r] = qc[r];
                    if (mode != 0) {
#pragma unroll
                        for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) {
                            gdx[r] = rzs * R2[o + r] - gdx[r];                     // ds
                            D2[o + r] = gdx[r];
                            lm[r] = L2[o + r];
                        }
                        soc::Wv(wc, e2c, qc, d, qc, false);                        // dz~
                        soc::Wv(wc, e2c, gdx, d, gdx, true);                       // ds~
                        tmax = fmax(tmax, fmax(soc::step(lm, gdx, d), soc::step(lm, qc, d)));
                        if (mode == 1) {
                            soc::jprod(gdx, qc, d, qc);
#pragma unroll
                            for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) CRv[o + r] = qc[r];
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < NLP; r++) {
                    const double gdx = row_dot_reg(r, td, yk), dv = W2[r];
                    const double dzv = dv * (gdx - D2[r]);
                    DZ[r] = dzv;
                    if (mode != 0) {
                        const double dsv = rzs * R2[r] - gdx;
                        D2[r] = dsv;
                        const double iw = sqrt(dv), il = 1. / L2[r];                // W = 1/sqrt(wb)
                        const double dzt = dzv / iw, dst = dsv * iw;
                        tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                        if (mode == 1) CRv[r] = dst * dzt;
                    }
                }
#pragma unroll
                for (int r = 0; r < NROW; r++) dz[r * KS + k] = DZ[r];
                if (mode != 0) {
#pragma unroll
                    for (int r = 0; r < NROW; r++) ds[r * KS + k] = D2[r];
                    if (mode == 1) {
#pragma unroll
                        for (int r = 0; r < NROW; r++) cr[r * KS + k] = CRv[r];
                    }
                }
            }
            // ---- block 3: interval pairs (row i+1 is loaded before row i is stored)
            if (hasint) {
                constexpr int NT = NB + NU + 1, NR = NT + 10;   // A~ | C | s column ; x_{k+1,i} dm dp dsm dsp rxv | rz- rz+ lam- lam+
                double XNu[NU], cur[NR], nxt[NR];
#pragma unroll
                for (int a = 0; a < NU; a++) XNu[a] = gv[(NX + a) * KS + k + 1];
                auto ld_row = [&](int i, double *row) {
                    const int o = MN + i;
#pragma unroll
                    for (int j = 0; j < NT; j++) row[j] = T(i, j, k);
                    row[NT] = gv[i * KS + k + 1];
                    row[NT + 1] = wb[o * KS + k]; row[NT + 2] = wb[(o + NX) * KS + k];
                    row[NT + 3] = ds[o * KS + k]; row[NT + 4] = ds[(o + NX) * KS + k];
                    row[NT + 5] = rxv_of(mode, csig, PN + i, k);
                    if (mode != 0) {
                        row[NT + 6] = rz[o * KS + k]; row[NT + 7] = rz[(o + NX) * KS + k];
                        row[NT + 8] = lam[o * KS + k]; row[NT + 9] = lam[(o + NX) * KS + k];
                    }
                };
                auto do_row = [&](int i, const double *cur) {
                    const int o = MN + i;
                    double acc = cur[NT], acc2 = 0;
#pragma unroll
                    for (int j = 0; j < NB; j += 2) { acc -= cur[j] * yk[j]; acc2 -= cur[j + 1] * yk[j + 1]; }
#pragma unroll
                    for (int a = 0; a < NU; a++) acc2 -= cur[NB + a] * XNu[a];
                    const double ady = acc + acc2 - cur[NB + NU] * ysig;
                    const double dm = capd(cur[NT + 1]), dp = capd(cur[NT + 2]);
                    const double qm = ady - cur[NT + 3], qp = -ady - cur[NT + 4];
                    const double dt = (cur[NT + 5] + dm * qm + dp * qp) / (dm + dp);
                    const double dzm = dm * (qm - dt), dzp = dp * (qp - dt);
                    dz[o * KS + k] = dzm; dz[(o + NX) * KS + k] = dzp;
                    dprim[(PN + i) * KS + k] = dt;
                    if (mode != 0) {
                        const double dsm = rzs * cur[NT + 6] - (ady - dt), dsp = rzs * cur[NT + 7] - (-ady - dt);
                        ds[o * KS + k] = dsm; ds[(o + NX) * KS + k] = dsp;
                        {
                            const double iw = sqrt(cur[NT + 1]), il = 1. / cur[NT + 8];
                            const double dzt = dzm / iw, dst = dsm * iw;
                            tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                            if (mode == 1) cr[o * KS + k] = dst * dzt;
                        }
                        {
                            const double iw = sqrt(cur[NT + 2]), il = 1. / cur[NT + 9];
                            const double dzt = dzp / iw, dst = dsp * iw;
                            tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                            if (mode == 1) cr[(o + NX) * KS + k] = dst * dzt;
                        }
                    }
                };
                ld_row(0, cur);
#pragma unroll 1
                for (int i = 0; i < NX; i += 2) {
                    ld_row(i + 1, nxt);
                    do_row(i, cur);
                    if (i + 2 < NX) ld_row(i + 2, cur);
                    do_row(i + 1, nxt);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    const int o = MN + i;
                    dprim[(PN + i) * KS + k] = 0.; dz[o * KS + k] = 0.; dz[(o + NX) * KS + k] = 0.;
                    if (mode != 0) { ds[o * KS + k] = 0.; ds[(o + NX) * KS + k] = 0.; if (mode == 1) { cr[o * KS + k] = 0.; cr[(o + NX) * KS + k] = 0.; } }
                }
            }
        }
        return tmax;
    }

    // scalars of the sigma rows shared by the forward/backward halves of a solve
    struct Glob { double rzg[4], rxg[2], kap_s, p_s[3], ysig; };
    // globals: rhs and local elimination for the sigma rows, then y_sigma (every lane computes the same scalars)
    SCPP_HD void globals_mid(int mode, double csig, double sigmu, double gsig, double ldot, Glob &g) const
    {
        if (sig_fixed) { for (int i = 0; i < 4; i++) g.rzg[i] = 0.; g.rxg[0] = g.rxg[1] = 0.; g.kap_s = 1.; g.p_s[0] = g.p_s[1] = g.p_s[2] = 0.; g.ysig = 0.; return; }
        const int r0g = RS * KS, p0g = PSN * KS;
        double w3[3] = {wb[r0g + 1], wb[r0g + 2], wb[r0g + 3]};
        const double e2i = ce[NCN * KS], d0 = wb[r0g];
        if (mode == 0) { for (int i = 0; i < 4; i++) g.rzg[i] = ds[r0g + i]; g.rxg[0] = dprim[p0g]; g.rxg[1] = dprim[p0g + 1]; }
        else if (mode == 1) { for (int i = 0; i < 4; i++) g.rzg[i] = -rz[r0g + i] + s[r0g + i]; g.rxg[0] = -rx[p0g]; g.rxg[1] = -rx[p0g + 1]; }
        else {
            const double l0 = lam[r0g];
            g.rzg[0] = -csig * rz[r0g] - sqrt(1. / d0) * ((-l0 * l0 - cr[r0g] + sigmu) / l0);
            double lm3[3] = {lam[r0g + 1], lam[r0g + 2], lam[r0g + 3]}, t1[3];
            soc::jprod(lm3, lm3, 3, t1);
            for (int i = 0; i < 3; i++) t1[i] = -t1[i] - cr[r0g + 1 + i];
            t1[0] += sigmu;
            soc::jdiv(lm3, t1, 3, t1);
            soc::Wv(w3, e2i, t1, 3, t1, false);
            for (int i = 0; i < 3; i++) g.rzg[1 + i] = -csig * rz[r0g + 1 + i] - t1[i];
            g.rxg[0] = -csig * rx[p0g]; g.rxg[1] = -csig * rx[p0g + 1];
        }
        double g3[3] = {-0.5, 0.5, 0.}, v[3];
        soc::Mv(w3, e2i, g3, 3, g.p_s);
        g.kap_s = g3[0] * g.p_s[0] + g3[1] * g.p_s[1];
        soc::Mv(w3, e2i, g.rzg + 1, 3, v);
        const double prz = g.p_s[0] * g.rzg[1] + g.p_s[1] * g.rzg[2] + g.p_s[2] * g.rzg[3];
        const double rho = (prz + g.rxg[1]) / g.kap_s;
        gsig += g.rxg[0] - d0 * g.rzg[0] - (v[2] - g.p_s[2] * rho);
        const double fsig = (gsig - ldot) / l_ss;
        g.ysig = fsig / l_ss;
    }
    // recovery of the global rows: dz, ds, dsigma, ddelta_sigma, cr (mode 1); returns their scaled step-length bound.
    // Call from lane 0 only (it stores).
    SCPP_HD double globals_recover(int mode, double rzs, const Glob &g)
    {
        if (sig_fixed) return 0.;
        const int r0g = RS * KS, p0g = PSN * KS;
        double tmax = 0;
        double w3[3] = {wb[r0g + 1], wb[r0g + 2], wb[r0g + 3]};
        const double e2i = ce[NCN * KS], d0 = wb[r0g], ysig = g.ysig;
        const double dz0 = d0 * (-ysig - g.rzg[0]);
        double q[3] = {-g.rzg[1], -g.rzg[2], -ysig - g.rzg[3]}, mq[3], dzq[3];
        const double pq = g.p_s[0] * q[0] + g.p_s[1] * q[1] + g.p_s[2] * q[2];
        const double dds = (g.rxg[1] - pq) / g.kap_s;
        soc::Mv(w3, e2i, q, 3, mq);
        for (int i = 0; i < 3; i++) dzq[i] = mq[i] + g.p_s[i] * dds;
        double rz4[4] = {0., 0., 0., 0.}, lm4[4] = {1., 1., 0., 0.};
        if (mode != 0) { for (int i = 0; i < 4; i++) { rz4[i] = rz[r0g + i]; lm4[i] = lam[r0g + i]; } }
        dz[r0g] = dz0; for (int i = 0; i < 3; i++) dz[r0g + 1 + i] = dzq[i];
        dprim[p0g] = ysig; dprim[p0g + 1] = dds;
        if (mode != 0) {
            double dsg[4] = {rzs * rz4[0] + ysig, rzs * rz4[1] + 0.5 * dds, rzs * rz4[2] - 0.5 * dds, rzs * rz4[3] + ysig};
            for (int i = 0; i < 4; i++) ds[r0g + i] = dsg[i];
            const double wv_ = sqrt(1. / d0), l0 = lm4[0];
            const double dzt0 = wv_ * dz0, dst0 = dsg[0] / wv_;
            tmax = fmax(tmax, fmax(-dst0 / l0, -dzt0 / l0));
            double lm3[3] = {lm4[1], lm4[2], lm4[3]}, dzt[3], dst[3], pr[3];
            soc::Wv(w3, e2i, dzq, 3, dzt, false);
            soc::Wv(w3, e2i, dsg + 1, 3, dst, true);
            tmax = fmax(tmax, fmax(soc::step(lm3, dst, 3), soc::step(lm3, dzt, 3)));
            if (mode == 1) { cr[r0g] = dst0 * dzt0; soc::jprod(dst, dzt, 3, pr); for (int i = 0; i < 3; i++) cr[r0g + 1 + i] = pr[i]; }
        }
        return tmax;
    }

    SCPP_HD void phase_solve(int mode, double csig, double sigmu, double rzs, double &tmax_out)
    {
        const double gsig = warp_sum(pass_rhs(mode, csig, sigmu));
        const double ldot = warp_sum(chain_forward());
        Glob g;
        globals_mid(mode, csig, sigmu, gsig, ldot, g);
        warp_sync();
        chain_backward(g.ysig);
        double tmax = pass_recover(mode, csig, rzs, g.ysig);
        if (lane_id() == 0) tmax = fmax(tmax, globals_recover(mode, rzs, g));
        warp_sync();
        tmax_out = warp_max(tmax);
    }

    // =============================================================================================================
    //  light stage-parallel passes used by the cold start: slack evaluation, cone margins / shifts
    // =============================================================================================================
    SCPP_HD void eval_slack(double *out)   // out = h - G x
    {
        const double sg = prim[PSN * KS], dsg = prim[PSN * KS + 1];
        FOR_LANE(k, K) {
            const bool hasint = k < K - 1;
#pragma unroll
            for (int r = 0; r < NROW; r++) { const RowDesc rd = M::crow(r); out[r * KS + k] = row_h(rd) - row_dot(r, k, prim); }
            out[TRO * KS + k] = prim[NB * KS + k];
#pragma unroll
            for (int j = 0; j < NB; j++) out[(TRO + 1 + j) * KS + k] = tr_row(j + 1) ? xibar(k, j) - prim[j * KS + k] : 0.;
#pragma unroll 2
            for (int i = 0; i < NX; i++) {
                double tm = 0., tp = 0.;
                if (hasint) {
                    double acc = prim[i * KS + k + 1];
#pragma unroll
                    for (int j = 0; j < NB; j++) acc -= T(i, j, k) * prim[j * KS + k];
#pragma unroll
                    for (int a = 0; a < NU; a++) acc -= T(i, NB + a, k) * prim[(NX + a) * KS + k + 1];
                    acc -= T(i, NB + NU, k) * sg + T(i, NB + NU + 1, k);
                    const double t = prim[(PN + i) * KS + k];
                    tm = t - acc; tp = t + acc;
                }
                out[(MN + i) * KS + k] = tm; out[(MN + NX + i) * KS + k] = tp;
            }
        }
        if (lane_id() == 0 && !sig_fixed) { const int r0 = RS * KS; out[r0] = sg - 0.001; out[r0 + 1] = 0.5 + 0.5 * dsg; out[r0 + 2] = 0.5 - 0.5 * dsg; out[r0 + 3] = sg - sigbar; }
        warp_sync();
    }
    // visit every cone of a row array: f(first index, stride, dimension)
    template <class F>
    SCPP_HD void for_cones(F &&f) const
    {
        FOR_LANE(k, K) {
            for (int r = 0; r < NLP; r++) f(r * KS + k, KS, 1);
            for (int c = 0; c < NCONE; c++) f((NLP + M::cone_off(c)) * KS + k, KS, M::cone_dim(c));
            f(TRO * KS + k, KS, D);
            if (k < K - 1) for (int r = MN; r < RS; r++) f(r * KS + k, KS, 1);
        }
        if (lane_id() == 0 && !sig_fixed) { f(RS * KS, 1, 1); f(RS * KS + 1, 1, 3); }
    }
    SCPP_HD void cone_margin(const double *u, double &mn, double &nrm2) const
    {
        double lmn = 1e300, n2 = 0;
        for_cones([&](int o, int st_, int d) {
            double t = 0;
            for (int i = 1; i < d; i++) t += u[o + i * st_] * u[o + i * st_];
            const double mg = u[o] - sqrt(t);
            if (mg < lmn) lmn = mg;
            n2 += t + u[o] * u[o];
        });
        mn = -warp_max(-lmn); nrm2 = warp_sum(n2);
    }
    SCPP_HD void cone_shift(double *u, double a) const { for_cones([&](int o, int, int) { u[o] += a; }); }
    // Cold start: every cone on its own -- a head is raised only where the cone's margin is below INIT_MARGIN.  ECOS / CVXOPT add ONE shift
    // (1 - worst margin) to every cone; with 3300 cone rows of very different size (virtual-control pairs with duals ~w_vc next to O(1e-2)
    // thrust cones) that makes the starting gap ~1e6 times the final one.  Measured on the kernel source (4 instances x 15 cold sub-problems):
    // 21.9 instead of 24.2 interior-point iterations per sub-problem with 0.1 (1: 24.6, 0.01: 22.0, 0.001: 24.2).  Only the starting point changes.
    static constexpr double INIT_MARGIN = 0.1;
    SCPP_HD void cone_shift_each(double *u, double target) const
    {
        for_cones([&](int o, int st_, int d) {
            double t = 0;
            for (int i = 1; i < d; i++) t += u[o + i * st_] * u[o + i * st_];
            const double mg = u[o] - sqrt(t);
            if (mg < target) u[o] += target - mg;
        });
    }

    // =============================================================================================================
    //  driver
    // =============================================================================================================
    // starting point of a sub-problem: 1 = previous interior point pulled back from the boundary, 2 = least-squares (cold) start,
    // 0 = the cold start's factorisation failed
    SCPP_HD int init_point(const IpmSettings &st_, bool have_prev)
    {
        const int np = n_prim(K), m = m_rows(K);
        Norms nm;
        double tm;
        if (have_prev && st_.warm > 0. && st_.warm < 1.) {
            // previous interior point of this instance, pulled back from the boundary; pinned variables keep their values
            const double lw = st_.warm, lc = 1. - st_.warm;
            FOR_LANE(k, K) { for (int i = 0; i < NB; i++) if (fixed(k, i)) prim[i * KS + k] = fixv[k * NB + i]; if (scvx) prim[NB * KS + k] = tr_rad; }
            if (sig_fixed && lane_id() == 0) { prim[PSN * KS] = sigbar; prim[PSN * KS + 1] = 0.; }
            FOR_LANE(e, m) { s[e] *= lw; z[e] *= lw; }
            warp_sync();
            cone_shift(s, lc); cone_shift(z, lc);
            warp_sync();
            return 1;
        } else {
            // ---- starting point (CVXOPT conelp / ECOS style): least-squares primal and dual points, W = I
            FOR_LANE(e, np) prim[e] = 0.;
            FOR_LANE(e, m) { s[e] = 0.; z[e] = 0.; }
            warp_sync();
            FOR_LANE(k, K) { for (int i = 0; i < NB; i++) prim[i * KS + k] = fixed(k, i) ? fixv[k * NB + i] : xibar(k, i); if (scvx) prim[NB * KS + k] = tr_rad; }
            if (lane_id() == 0) { prim[PSN * KS] = sigbar; prim[PSN * KS + 1] = 0.; }
            warp_sync();
            cone_shift(s, 1.); cone_shift(z, 1.);
            warp_sync();
            pass_residuals(nm, true);
            if (!phase_factor()) return 0;
            // primal: min |G x - h|  ->  G dx - dz = slack(x0)
            eval_slack(ds);
            FOR_LANE(e, np) dprim[e] = 0.;
            warp_sync();
            phase_solve(0, 0., 0., 0., tm);
            FOR_LANE(e, np) prim[e] += dprim[e];
            warp_sync();
            eval_slack(s);
            {
                double mg, n2; cone_margin(s, mg, n2);
                if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift_each(s, INIT_MARGIN); }
                warp_sync();
            }
            // dual: min |z| s.t. G'z + c = 0  ->  rx = -c, rz = 0
            FOR_LANE(e, m) ds[e] = 0.;
            FOR_LANE(e, np) dprim[e] = 0.;
            warp_sync();
            FOR_LANE(k, K) { dprim[NB * KS + k] = -w_tr; if (k < K - 1) for (int i = 0; i < NX; i++) dprim[(PN + i) * KS + k] = -w_vc; }
            if (lane_id() == 0) { dprim[PSN * KS] = -w_time; dprim[PSN * KS + 1] = -w_trs; }
            warp_sync();
            phase_solve(0, 0., 0., 0., tm);
            FOR_LANE(e, m) z[e] = dz[e];
            warp_sync();
            {
                double mg, n2; cone_margin(z, mg, n2);
                if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift_each(z, INIT_MARGIN); }
                warp_sync();
            }
            return 2;
        }
    }
    // Sliced driver.  `state` (IPM_STATE doubles, per instance, global memory) carries the solver across kernel launches:
    // the engine runs at most `budget` interior-point iterations per launch and re-balances the batch between launches
    // (instances need very different iteration counts).  state[0] == 0: begin a new sub-problem.
    // Returns with finished == false when the budget ran out; everything else lives in the workspace already.
    static constexpr int IPM_STATE = 32;
    SCPP_HD IpmResult solve(const IpmSettings &st_, bool have_prev, int budget, double *state, bool &finished)
    {
        IpmResult res;
        res.status = 1; res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.; res.point_ok = 1; res.pad_ = 0;
        const int np = n_prim(K), m = m_rows(K);
        finished = true;
        tables_init();
        const bool resume = state[0] != 0.;
#if defined(SCPP_DCAP_SMEM) && defined(__CUDA_ARCH__)
        if (lane_id() == 0) sc()[24] = resume ? state[ST_DCAP] : 0.;
        warp_sync();
#endif
        Norms nm;
        double tm;
        double best = 1e300, pending = 0.;
        int it = 0;
        if (resume) {
            it = (int)state[1]; pending = state[2]; best = state[3];
#if !defined(SCPP_R01_SOLVE)
            dcap = state[ST_DCAP];
#endif
            res.pres = state[4]; res.dres = state[5]; res.gap = state[6]; res.relgap = state[7]; res.pcost = state[8]; res.iterations = (int)state[9];
        } else {
            const int how = init_point(st_, have_prev);
            if (how == 0) { res.status = 2; if (lane_id() == 0) state[0] = 0.; warp_sync(); return res; }
            if (how == 2) budget -= 1;           // the least-squares start costs about one iteration
        }
        const double cnorm = sqrt(w_time * w_time + w_trs * w_trs + K * w_tr * w_tr + (K - 1) * NX * w_vc * w_vc);
        const double resx0 = fmax(1., cnorm);
        const int degree = K * (NLP + NCN) + (K - 1) * 2 * NX + (sig_fixed ? 0 : 2);
        // One slice = factorisation + the two solves of iteration `it`, then update + residuals + termination test of `it+1`,
        // so that the cheap final test never occupies a launch of its own.  A resumed solver re-enters after the test.
        bool past_test = resume;
        double gap_cur = resume ? state[10] : 0.;
#pragma unroll 1
        for (; it <= st_.maxit; it++) {
            if (!past_test) {
                if (pending != 0.) pass_update(pending);
                pass_residuals(nm, false);
                const double resz0 = fmax(1., sqrt(nm.h2));
                const double pres = sqrt(nm.rz2) / resz0, dres = sqrt(nm.rx2) / resx0, gap = nm.gap, pcost = nm.pcost;
                const double dcost = pcost - gap + nm.zrz - nm.xrx;
                double relgap = 1e300;
                if (pcost < 0.) relgap = gap / -pcost; else if (dcost > 0.) relgap = gap / dcost;
                const double score = fmax(fmax(pres, dres) / st_.feastol, fmin(gap / st_.abstol, relgap / st_.reltol));
#if !defined(__CUDACC__)
                if (getenv("SCPP_DEBUG")) fprintf(stderr, "it %2d pres %.2e dres %.2e gap %.2e relgap %.2e pcost %.6e bad %d\n", it, pres, dres, gap, relgap, pcost, nm.bad);
#endif
                if (!nm.bad && score < best) {
                    best = score;
                    res.pres = pres; res.dres = dres; res.gap = gap; res.relgap = relgap; res.pcost = pcost; res.iterations = it;
                    if (score <= 1e4) copy4(best_, prim, np);      // a fallback iterate only matters inside the accuracy band
                    warp_sync();
                }
                if (!nm.bad && pres <= st_.feastol && dres <= st_.feastol && (gap <= st_.abstol || relgap <= st_.reltol)) { res.status = 0; break; }
                // stop on failure, on the iteration limit, on divergence, or at the accuracy floor of the condensed system: once an
                // iterate within 10x of the tolerances exists and the next one is worse (the dual residual grows with the
                // conditioning as the gap closes), later iterates only get worse; the best iterate is returned below
                if (nm.bad || it == st_.maxit || (score > 1e3 * best && best < 1e4) || (best <= 10. && score > best)) { res.status = nm.bad ? 2 : (it == st_.maxit ? 1 : 2); res.point_ok = !nm.bad; break; }
                gap_cur = gap;
                if (budget <= 0) {                       // out of budget: park the solver state, continue in the next launch
                    if (lane_id() == 0) {
                        state[0] = 1.; state[1] = it; state[2] = pending; state[3] = best;
                        state[4] = res.pres; state[5] = res.dres; state[6] = res.gap; state[7] = res.relgap; state[8] = res.pcost; state[9] = res.iterations;
                        state[10] = gap_cur;
#if !defined(SCPP_R01_SOLVE)
                        state[ST_DCAP] = dcap;
#endif
                    }
                    warp_sync();
                    finished = false;
                    return res;
                }
            }
            past_test = false;
            budget--;
#if defined(SCPP_R01_SOLVE)      // A/B experiment: the round-1 driver (no retry with a tightened cap)
            if (!phase_factor()) { res.status = 2; break; }
#else
            // numerically indefinite: regularise (see dcap) and repeat THIS iteration in the next launch.  No loop around the factorisation and
            // no branch back into the main loop: with either the compiler schedules the whole inlined iteration 8-11 % slower (43.3 k vs
            // 39.5-40.1 k instance-iterations/s, profiles/r02o_*, r02p_*).
            if (!phase_factor()) {
                if (tighten_cap()) {
                    // park exactly as when the budget runs out (the state is that of "after the test of iterate `it`"): the NEXT launch re-enters
                    // at the factorisation with the tightened cap.  An exit, not a branch back into the loop: dcap stays loop-invariant.
                    if (lane_id() == 0) {
                        state[0] = 1.; state[1] = it; state[2] = pending; state[3] = best;
                        state[4] = res.pres; state[5] = res.dres; state[6] = res.gap; state[7] = res.relgap; state[8] = res.pcost; state[9] = res.iterations;
                        state[10] = gap_cur; state[ST_DCAP] = dcap;
                    }
                    warp_sync();
                    finished = false;
                    return res;
                }
                res.status = 2; break;
            }
#endif
            double tmax;
            phase_solve(1, 1., 0., -1., tmax);                               // affine direction
            const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
            const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = gap_cur / degree;
            phase_solve(2, 1. - sig, sig * mu, -(1. - sig), tmax);           // combined direction
            pending = tmax <= step_frac ? 1. : step_frac / tmax;
        }
        if (res.status != 0) {
            if (best <= 1e4) copy4(prim, best_, np);
            warp_sync();
            if (best <= 1e4) res.status = 3;      // best iterate inside the accuracy band but short of the tolerances: reduced accuracy
        } else res.iterations = it;
        if (lane_id() == 0) state[0] = 0.;
        warp_sync();
        return res;
    }

    // =============================================================================================================
    //  Split pipeline: one interior-point iteration as a SEQUENCE OF KERNELS, each with the mapping its work wants
    //     start    (warp per new sub-problem)   starting point, residuals, first termination test
    //     assemble (warp per (instance, stage)) H_kk | O_k | b_k of every stage, fully parallel
    //     factor   (warp per instance)          block Cholesky chain over the assembled records
    //     rhs / recover (warp per 32 stages)    stage-parallel passes, several warps per instance
    //     chain    (warp per instance)          coupling, forward substitution, sigma rows, back substitution
    //     update   (warp per 32 stages)         step + residuals of the next iterate
    //     test     (warp per instance)          coupling + global rows of the residual, termination, bookkeeping
    //  Scalars travel in `state` (per instance), per-warp partial sums in `part`.
    //  state: [0] 0 idle / 1 mid-solve  [1] it  [2] pending (monolithic slices)  [3] best  [4..9] best iterate's result
    //         [10] gap of the current iterate  [11] l_ss  [12] factorisation failed  [13..23] Glob of the running solve
    // =============================================================================================================
    static constexpr int ST_LSS = 11, ST_FAIL = 12, ST_GLOB = 13, ST_DCAP = 24, ST_NOPOINT = 25;
    static constexpr int PT_ACC = 8, PT_GSIG = 9, PT_TAFF = 10, PT_TCMB = 11;
    SCPP_HD int nparts() const { return (K + 31) / 32; }
    SCPP_HD void set_part(int w) { k_lo = 32 * w; k_hi = (32 * w + 32 < K) ? 32 * w + 32 : K; }
    SCPP_HD double part_sum(int slot) const { double v = 0; for (int w = 0; w < nparts(); w++) v += part[w * PSTR + slot]; return v; }
    SCPP_HD double part_max(int slot) const { double v = 0; for (int w = 0; w < nparts(); w++) v = fmax(v, part[w * PSTR + slot]); return v; }
    SCPP_HD int degree() const { return K * (NLP + NCN) + (K - 1) * 2 * NX + (sig_fixed ? 0 : 2); }
    // parameters of a solve: mode 1 = affine direction, mode 2 = combined direction (centering from the affine step length)
    SCPP_HD void solve_params(int mode, const double *state, double &csig, double &sigmu, double &rzs) const
    {
        if (mode == 1) { csig = 1.; sigmu = 0.; rzs = -1.; return; }
        const double tmax = part_max(PT_TAFF);
        const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
        const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = state[10] / degree();
        csig = 1. - sig; sigmu = sig * mu; rzs = -(1. - sig);
    }
    SCPP_HD void glob_store(double *state, const Glob &g) const
    {
        if (lane_id() == 0) {
            double *q = state + ST_GLOB;
            for (int i = 0; i < 4; i++) q[i] = g.rzg[i];
            q[4] = g.rxg[0]; q[5] = g.rxg[1]; q[6] = g.kap_s; q[7] = g.p_s[0]; q[8] = g.p_s[1]; q[9] = g.p_s[2]; q[10] = g.ysig;
        }
    }
    SCPP_HD void glob_load(const double *state, Glob &g) const
    {
        const double *q = state + ST_GLOB;
        for (int i = 0; i < 4; i++) g.rzg[i] = q[i];
        g.rxg[0] = q[4]; g.rxg[1] = q[5]; g.kap_s = q[6]; g.p_s[0] = q[7]; g.p_s[1] = q[8]; g.p_s[2] = q[9]; g.ysig = q[10];
    }

    SCPP_HD void sp_factor(double *state)
    {
        const bool ok = chain_factor();
        // numerically indefinite: tighten the cap (see dcap); 2 = the rest of this round is skipped and the next round repeats the iteration
        const bool retry = !ok && tighten_cap();
        if (lane_id() == 0) { state[ST_LSS] = l_ss; state[ST_FAIL] = ok ? 0. : (retry ? 2. : 1.); state[ST_DCAP] = dcap; }
        warp_sync();
    }
    SCPP_HD void sp_rhs(int mode, const double *state, int w)
    {
        double csig, sigmu, rzs;
        solve_params(mode, state, csig, sigmu, rzs);
        set_part(w);
        const double g = warp_sum(pass_rhs(mode, csig, sigmu, true));
        if (lane_id() == 0) part[w * PSTR + PT_GSIG] = g;
        warp_sync();
    }
    SCPP_HD void sp_chain(int mode, double *state)
    {
        double csig, sigmu, rzs;
        solve_params(mode, state, csig, sigmu, rzs);
        l_ss = state[ST_LSS];
        rhs_couple();
        const double gsig = part_sum(PT_GSIG);
        const double ldot = warp_sum(chain_forward());
        Glob g;
        globals_mid(mode, csig, sigmu, gsig, ldot, g);
        warp_sync();
        chain_backward(g.ysig);
        glob_store(state, g);
        warp_sync();
    }
    SCPP_HD void sp_recover(int mode, const double *state, int w)
    {
        double csig, sigmu, rzs;
        solve_params(mode, state, csig, sigmu, rzs);
        Glob g;
        glob_load(state, g);
        set_part(w);
        double tmax = pass_recover(mode, csig, rzs, g.ysig);
        if (w == 0 && lane_id() == 0) tmax = fmax(tmax, globals_recover(mode, rzs, g));
        tmax = warp_max(tmax);
        if (lane_id() == 0) part[w * PSTR + (mode == 1 ? PT_TAFF : PT_TCMB)] = tmax;
        warp_sync();
    }
    // step of the iterate; every warp of the instance must have finished it before sp_residuals reads its neighbours' stages
    SCPP_HD void sp_update(int w)
    {
        const double tmax = part_max(PT_TCMB);
        set_part(w);
        pass_update(tmax <= step_frac ? 1. : step_frac / tmax);
    }
    SCPP_HD void sp_residuals(int w)
    {
        Norms nm;
        set_part(w);
        pass_residuals(nm, false, true);
        if (lane_id() == 0) {
            double *q = part + w * PSTR;
            q[0] = nm.gap; q[1] = nm.rz2; q[2] = nm.pcost; q[3] = nm.zrz; q[4] = nm.rx2; q[5] = nm.xrx; q[6] = nm.h2; q[7] = nm.bad; q[PT_ACC] = nm.acc_sig;
        }
        warp_sync();
    }
    // termination test and bookkeeping of iterate `it` (the logic of the monolithic loop); returns true when the sub-problem is done
    SCPP_HD bool test_and_book(const IpmSettings &st_, const Norms &nm, int it, bool factor_failed, double *state, IpmResult &res)
    {
        const int np = n_prim(K);
        double best = (it == 0) ? 1e300 : state[3];
        res.status = 1; res.point_ok = 1; res.pad_ = 0;
        if (it == 0) { res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.; }
        else { res.pres = state[4]; res.dres = state[5]; res.gap = state[6]; res.relgap = state[7]; res.pcost = state[8]; res.iterations = (int)state[9]; }
        bool done = false;
        double gap_cur = 0.;
        if (factor_failed) { res.status = 2; done = true; }
        else {
            const double cnorm = sqrt(w_time * w_time + w_trs * w_trs + K * w_tr * w_tr + (K - 1) * NX * w_vc * w_vc);
            const double resx0 = fmax(1., cnorm), resz0 = fmax(1., sqrt(nm.h2));
            const double pres = sqrt(nm.rz2) / resz0, dres = sqrt(nm.rx2) / resx0, gap = nm.gap, pcost = nm.pcost;
            const double dcost = pcost - gap + nm.zrz - nm.xrx;
            double relgap = 1e300;
            if (pcost < 0.) relgap = gap / -pcost; else if (dcost > 0.) relgap = gap / dcost;
            const double score = fmax(fmax(pres, dres) / st_.feastol, fmin(gap / st_.abstol, relgap / st_.reltol));
            if (!nm.bad && score < best) {
                best = score;
                res.pres = pres; res.dres = dres; res.gap = gap; res.relgap = relgap; res.pcost = pcost; res.iterations = it;
                if (score <= 1e4) copy4(best_, prim, np);
                warp_sync();
            }
            if (!nm.bad && pres <= st_.feastol && dres <= st_.feastol && (gap <= st_.abstol || relgap <= st_.reltol)) { res.status = 0; done = true; }
            else if (nm.bad || it == st_.maxit || (score > 1e3 * best && best < 1e4) || (best <= 10. && score > best)) {
                res.status = nm.bad ? 2 : (it == st_.maxit ? 1 : 2); res.point_ok = !nm.bad; done = true;
            }
            gap_cur = gap;
        }
        if (!done) {
            if (lane_id() == 0) {
                state[0] = 1.; state[1] = it; state[2] = 0.; state[3] = best;
                state[4] = res.pres; state[5] = res.dres; state[6] = res.gap; state[7] = res.relgap; state[8] = res.pcost; state[9] = res.iterations;
                state[10] = gap_cur; state[ST_FAIL] = 0.;
                if (it == 0) state[ST_DCAP] = 0.;
            }
            warp_sync();
            return false;
        }
        if (res.status != 0) {
            if (best <= 1e4) copy4(prim, best_, np);
            warp_sync();
            if (best <= 1e4) res.status = 3;      // best iterate inside the accuracy band but short of the tolerances: reduced accuracy
        } else res.iterations = it;
        if (lane_id() == 0) state[0] = 0.;
        warp_sync();
        return true;
    }
    // start of a sub-problem (warp per instance): starting point, residuals, test of iterate 0
    SCPP_HD bool sp_start(const IpmSettings &st_, bool have_prev, double *state, IpmResult &res)
    {
        tables_init();
        Norms nm;
        const int how = init_point(st_, have_prev);
        if (how != 0) pass_residuals(nm, false);
        return test_and_book(st_, nm, 0, how == 0, state, res);
    }
    // end of a round (warp per instance): finish the residual of the new iterate, test it
    SCPP_HD bool sp_test(const IpmSettings &st_, double *state, IpmResult &res)
    {
        Norms nm;
        nm.gap = part_sum(0); nm.rz2 = part_sum(1); nm.pcost = part_sum(2); nm.zrz = part_sum(3); nm.rx2 = part_sum(4); nm.xrx = part_sum(5);
        nm.h2 = part_sum(6); nm.bad = part_max(7) != 0.; nm.acc_sig = part_sum(PT_ACC);
        if (state[ST_FAIL] == 2.) { if (lane_id() == 0) state[ST_FAIL] = 0.; warp_sync(); return false; }
        const bool failed = state[ST_FAIL] != 0.;
        if (!failed) { residual_couple(nm); if (!sig_fixed) residual_globals(nm, false); }
        return test_and_book(st_, nm, (int)state[1] + 1, failed, state, res);
    }

#include "ipm_cta.inl"
};

} // namespace scpp
