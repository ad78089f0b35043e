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
// Execution shape (v2): every phase is a SWEEP over the K stages (stage k = node k + shooting interval k).  All
// per-instance state lives in HBM in stage-major records; for each stage the warp copies the stage's records
// (the [A|B|C|s|z]_k tile, the cone rows, the factor blocks) into a shared-memory window with coalesced 16-byte
// asynchronous copies, works on the window (cone arithmetic: one lane per cone / LP row / virtual-control pair;
// tile and block arithmetic: warp-cooperative), and writes the results back coalesced.  Global memory is never
// touched with per-lane strided or dependent accesses.
#pragma once
#include "models.cuh"
#include "blockops.cuh"
#if !defined(__CUDACC__)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#endif

namespace scpp {

struct IpmSettings {
    double feastol, abstol, reltol;
    int maxit;
    int pad_;
    double warm;     // 0: cold start of every sub-problem (what ECOS does).  0 < warm < 1: when the instance has a previous sub-problem
                     // solution, start from that interior point pulled back from the boundary, (s,z) <- warm*(s,z) + (1-warm)*e,
                     // and skip the least-squares start.  Same optimum (parity-tested), ~2.8x fewer interior-point iterations.
};

struct IpmResult {
    int status;      // 0 optimal, 1 max iterations, 2 numerical failure, 3 reduced accuracy
    int iterations;
    double pres, dres, gap, relgap, pcost;
};

#if defined(__CUDACC__)
#define SCPP_OUTLINE __host__ __device__ __forceinline__
#else
#define SCPP_OUTLINE inline
#endif

// ---- second-order-cone primitives (one lane, one cone); kept out of line to bound the code size -------------------
namespace soc {
constexpr int SOC_MAXD = 4;   // the single-lane primitives serve the small model cones (dimension <= 4); the trust region is warp-cooperative
SCPP_HD double jn2(const double *u, int d)
{
    double n = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) n += u[i] * u[i];
    return u[0] * u[0] - n;
}

SCPP_OUTLINE bool scale(const double *sk, const double *zk, int d, double *w, double &e2i, double *lm)
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
SCPP_OUTLINE void Mv(const double *w, double e2i, const double *v, int d, double *o)
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
SCPP_OUTLINE void Wv(const double *w, double e2i, const double *v, int d, double *o, bool inv)
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
SCPP_OUTLINE void jprod(const double *u, const double *v, int d, double *o)   // o may alias u or v
{
    double dot = 0;
#pragma unroll
    for (int i = 0; i < SOC_MAXD; i++) if (i < d) dot += u[i] * v[i];
    const double u0 = u[0], v0 = v[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = u0 * v[i] + v0 * u[i];
    o[0] = dot;
}
SCPP_OUTLINE void jdiv(const double *lm, const double *dv, int d, double *o)   // o = lm \ dv  (o may alias dv)
{
    double den = jn2(lm, d), l1d1 = 0;
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) l1d1 += lm[i] * dv[i];
    const double x0 = (lm[0] * dv[0] - l1d1) / den, il0 = 1. / lm[0];
#pragma unroll
    for (int i = 1; i < SOC_MAXD; i++) if (i < d) o[i] = (dv[i] - x0 * lm[i]) * il0;
    o[0] = x0;
}
SCPP_OUTLINE double step(const double *lm, const double *dk, int d)
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
    static constexpr int NTASK = NCN + NLP + NX;    // per-stage cone tasks: cones, LP rows, virtual-control pairs
    // stage-major record strides (even => 16-byte aligned records)
    static constexpr int RS = pad2(MN + 2 * NX);    // cone rows of one stage: node rows | interval rows (s-: NX, s+: NX)
    static constexpr int PS = pad2(PN + NX);        // primal of one stage: xi | delta | t
    static constexpr int CS = pad2(NCN);            // eta^-2 per cone of one stage
    static constexpr int FS = pad2(2 * BLK + 2 * NB);   // Linv_kk | L_{k+1,k} | l_k | f_k
    static constexpr int OFF_LN = BLK, OFF_L = 2 * BLK, OFF_F = 2 * BLK + NB;
    static_assert(M::MAXDIM <= soc::SOC_MAXD, "single-lane cone primitives are unrolled for dimension <= 4");
    static_assert(NB % 2 == 0 && NC % 2 == 0 && NX % 2 == 0, "16-byte record alignment needs even nx, nx+nu and tile width");

    SCPP_HD static int m_rows(int K) { return K * RS + 4; }
    SCPP_HD static int n_prim(int K) { return K * PS + 2; }
    SCPP_HD static int n_ce(int K) { return K * CS + 2; }
    SCPP_HD static int ws_doubles(int K) { return 4 * n_prim(K) + 8 * m_rows(K) + n_ce(K) + K * FS + 16; }
    // shared window: tile | factor record | L_{k,k-1} carry | UNION{ 8 row arrays + primal windows ; phase F: wb + H,O + model terms }
    //                | compact carry of interval k-1 | vectors | scalars | per-stage row coefficients, reverse map, constants
    static constexpr int NROW = NLP + NCR;
    static constexpr int HNC = pad2(NX + NX * NU + NU * NU);      // D | D C | C' D C  of the previous interval
    static constexpr int W_DD = 0, W_FAC = W_DD + pad2(NX * NCP), W_LP = W_FAC + FS, W_ROW = W_LP + BLK, W_PRIM = W_ROW + 8 * RS,
                         W_MAT = W_ROW + RS, W_RK = W_MAT + 2 * BLK,
                         W_UEND = (W_PRIM + 4 * PS + pad2(NB)) > (W_RK + 2 * NRK * NB) ? (W_PRIM + 4 * PS + pad2(NB)) : (W_RK + 2 * NRK * NB),
                         W_HN = W_UEND, W_VEC = W_HN + HNC, W_X = W_VEC + 6 * NB, W_SC = W_X + 2 * pad2(NX),
                         W_RCQ = W_SC + 32, W_RIDX = W_RCQ + 4 * NROW, W_REV = W_RIDX + pad2(2 * NROW), W_CST = W_REV + pad2(2 * NB),
                         W_END = W_CST + pad2(MAX_CST + 4);
    SCPP_HD static int sm_doubles() { return W_END; }

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
    // ---- workspace (global memory, per instance, stage-major) ---------------------------------------------------
    double *prim, *dprim, *rx, *best_;
    double *s, *z, *wb, *lam, *rz, *cr, *dz, *ds;
    double *ce;
    double *fac;
    double *sm;            // per-warp shared window
    double l_ss;           // Cholesky pivot of the sigma border

    SCPP_HD void bind(double *ws, double *smem)
    {
        const int np = n_prim(K), m = m_rows(K);
        double *p = ws;
        prim = p; p += np; dprim = p; p += np; rx = p; p += np; best_ = p; p += np;
        s = p; p += m; z = p; p += m; wb = p; p += m; lam = p; p += m; rz = p; p += m; cr = p; p += m; dz = p; p += m; ds = p; p += m;
        ce = p; p += n_ce(K);
        fac = p;
        sm = smem;
    }
    // accessors used by the SC glue (sc.cuh)
    SCPP_HD double xi_at(int k, int i) const { return prim[k * PS + i]; }
    SCPP_HD double delta_at(int k) const { return prim[k * PS + NB]; }
    SCPP_HD double t_at(int k, int i) const { return prim[k * PS + PN + i]; }
    SCPP_HD double sigma_val() const { return prim[K * PS]; }
    SCPP_HD double dsigma_val() const { return prim[K * PS + 1]; }

    SCPP_HD bool fixed(int k, int i) const { return (fixm[k] >> i) & 1u; }

    // ---- window plumbing ------------------------------------------------------------------------------------------
    // cooperative copy of n (even) doubles global -> shared, 16 bytes per lane per step, asynchronous on the device
    SCPP_HD void ld(double *dst, const double *src, int n) const
    {
#if defined(__CUDA_ARCH__)
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst);
        for (int c = lane_id(); c < n / 2; c += LANES)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * c), "l"(src + 2 * c) : "memory");
#else
        memcpy(dst, src, sizeof(double) * n);
#endif
    }
    // the [A|B|C|s|z]_k tile: NX rows of NC doubles into rows of stride NCP
    SCPP_HD void ld_dd(int k) const
    {
        const double *src = dd + (size_t)k * NX * NC;
        double *t = sm + W_DD;
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
    SCPP_HD void ld_wait() const
    {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
#endif
        warp_sync();
    }
    // cooperative coalesced store shared -> global
    SCPP_HD void st(double *dst, const double *src, int n) const { FOR_LANE(e, n) dst[e] = src[e]; }

    SCPP_HD double *row(int a) const { return sm + W_ROW + a * RS; }     // 8 row-array windows
    SCPP_HD double *pw(int a) const { return sm + W_PRIM + a * PS; }     // 4 primal windows (+ NB spare after the 4th)
    SCPP_HD double *vec(int i) const { return sm + W_VEC + i * NB; }
    SCPP_HD double *xv(int i) const { return sm + W_X + i * pad2(NX); }
    SCPP_HD double *sc() const { return sm + W_SC; }
    SCPP_HD double *tile() const { return sm + W_DD; }
    SCPP_HD double *facw() const { return sm + W_FAC; }

    // task -> (type, row offset in the stage window, dimension, cone index): 0 LP row, 1 second-order cone, 2 virtual-control pair
    SCPP_HD static void task(int tk, int &type, int &o, int &d, int &ci)
    {
        if (tk < NCONE) { type = 1; o = NLP + M::cone_off(tk); d = M::cone_dim(tk); ci = tk; }
        else if (tk == NCONE) { type = 1; o = TRO; d = 1 + NB; ci = NCONE; }
        else if (tk < NCN + NLP) { type = 0; o = tk - NCN; d = 1; ci = -1; }
        else { type = 2; o = MN + (tk - NCN - NLP); d = 1; ci = -1; }
    }
    // per-stage row coefficients in shared memory: RCQ[r][0..2] = coefficients, RCQ[r][3] = h ; RIDX[r][q] = variable index
    // (or -1); REV[j][t] = (row*4+q) of the up-to-3 entries that touch variable j (or -1).  CST = per-instance constants.
    SCPP_HD double *rcq() const { return sm + W_RCQ; }
    SCPP_HD int *ridx() const { return reinterpret_cast<int *>(sm + W_RIDX); }
    SCPP_HD int *rev() const { return reinterpret_cast<int *>(sm + W_REV); }
    SCPP_HD double *cstw() const { return sm + W_CST; }
    // once per solve: constants, index table, reverse map and the k-independent coefficients
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
        FOR_LANE(j, NB) {
            int n = 0;
            for (int t = 0; t < 4; t++) rev()[j * 4 + t] = -1;
            for (int r = 0; r < NROW; r++) { const RowDesc rd = M::row(r); for (int q = 0; q < rd.n; q++) if (rd.idx[q] == j && n < 4) rev()[j * 4 + n++] = r * 4 + q; }
        }
        warp_sync();
    }
    // once per stage: only the linearised minimum-thrust row depends on k (coefficient slots < 0 take -tdir[k])
    SCPP_HD void tables_stage(int k) const
    {
        FOR_LANE(e, NROW * 3) {
            const int r = e / 3, q = e - 3 * r;
            const RowDesc rd = M::row(r);
            if (q < rd.n && rd.cs[q] < 0) rcq()[r * 4 + q] = -tdir[3 * k + (-rd.cs[q] - 1)];
        }
    }
    SCPP_HD double row_h(int r) const { return rcq()[r * 4 + 3]; }
    // gather of the model-row contributions G' v onto variable j of the node
    SCPP_HD double model_GT(int j, int, const double *v /* node rows window */) const
    {
        double a = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) { const int e = rev()[j * 4 + t]; if (e >= 0) a += rcq()[e] * v[e >> 2]; }
        return a;
    }
    SCPP_HD double model_G(int r, int, const double *x) const   // (G x)_r for a model row
    {
        double a = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) { const int i = ridx()[r * 4 + q]; if (i >= 0) a += rcq()[r * 4 + q] * x[i]; }
        return a;
    }
    // r_i = x_{k+1,i} - (A~ xi_k)_i - (C u_{k+1})_i - s_i sigma [- z_i]   from windows (tile staged)
    SCPP_HD double dyn_row(int i, const double *xk, const double *xn, double sg, bool with_const) const
    {
        const double *t = tile() + i * NCP;
        double acc = xn[i], acc2 = 0;
#pragma unroll 3
        for (int j = 0; j < NB; j += 2) { acc -= t[j] * xk[j]; acc2 -= t[j + 1] * xk[j + 1]; }
        acc += acc2;
#pragma unroll
        for (int j = 0; j < NU; j++) acc -= t[NB + j] * xn[NX + j];
        acc -= t[NB + NU] * sg;
        if (with_const) acc -= t[NB + NU + 1];
        return acc;
    }
    // out_k -= A~' w ; carry = [w ; -C' w] (contribution to node k+1) ; returns lane-partial of -s'w
    SCPP_HD double dyn_JT(const double *w, double *out_k, double *carry) const
    {
        const double *t = tile();
        FOR_LANE(j, NB) {
            double a = 0, a2 = 0, c = 0;
#pragma unroll 7
            for (int i = 0; i < NX; i += 2) { a += t[i * NCP + j] * w[i]; a2 += t[(i + 1) * NCP + j] * w[i + 1]; }
            out_k[j] -= a + a2;
            if (j < NX) c = w[j];
            else {
#pragma unroll 7
                for (int i = 0; i < NX; i++) c -= t[i * NCP + NB + (j - NX)] * w[i];
            }
            carry[j] = c;
        }
        double sg = 0;
        FOR_LANE(i, NX) sg -= t[i * NCP + NB + NU] * w[i];
        return sg;
    }
    SCPP_HD void load_xibar(int k, double *dst) const
    {
        FOR_LANE(i, NB) dst[i] = i < NX ? Xbar[k * NX + i] : Ubar[k * NU + (i - NX)];
    }

    // ---- trust-region cone (dimension D = 1 + NB): WARP-COOPERATIVE primitives, lane i owns element i, vectors in the
    //      shared window, reductions by shuffles.  Every lane of the warp must call them (uniform control flow).
    //      Outputs may alias inputs.
    static constexpr int D = 1 + NB;
    SCPP_HD bool tr_scale(const double *sk, const double *zk, double *w, double &e2i, double *lm) const
    {
        double a = 0, b = 0, c = 0;
        FOR_LANE(i, D) { if (i > 0) { a += sk[i] * sk[i]; b += zk[i] * zk[i]; } c += sk[i] * zk[i]; }
        warp_sum3(a, b, c);
        const double s0 = sk[0], z0 = zk[0];
        const double ss = s0 * s0 - a, zz = z0 * z0 - b;
        if (!(ss > 0.) || !(zz > 0.) || !(s0 > 0.) || !(z0 > 0.)) return false;
        const double sn = sqrt(ss), zn = sqrt(zz);
        const double i2g = 1. / (2. * sqrt((1. + c / (sn * zn)) / 2.));
        const double isn = i2g / sn, izn = i2g / zn;
        const double w0 = s0 * isn + z0 * izn;
        double w1z1 = 0;
        warp_sync();
        FOR_LANE(i, D) { const double wi = (i == 0) ? w0 : sk[i] * isn - zk[i] * izn; if (i > 0) w1z1 += wi * zk[i]; w[i] = wi; }
        w1z1 = warp_sum(w1z1);
        e2i = zn / sn;
        const double eta = sqrt(sn / zn), f = z0 + w1z1 / (1. + w0);
        FOR_LANE(i, D) lm[i] = (i == 0) ? eta * (w0 * z0 + w1z1) : eta * (zk[i] + f * w[i]);
        warp_sync();
        return true;
    }
    SCPP_HD void tr_Mv(const double *w, double e2i, const double *v, double *o) const   // o = W^-2 v
    {
        double dot = 0;
        FOR_LANE(i, D) dot += (i == 0 ? w[0] * v[0] : -w[i] * v[i]);
        dot = warp_sum(dot);
        warp_sync();
        FOR_LANE(i, D) o[i] = (i == 0) ? e2i * (2. * dot * w[0] - v[0]) : e2i * (-2. * dot * w[i] + v[i]);
        warp_sync();
    }
    SCPP_HD void tr_Wv(const double *w, double e2i, const double *v, double *o, bool inv) const   // o = W v | W^-1 v
    {
        const double eta = 1. / sqrt(e2i), sg = inv ? -1. : 1., scl = inv ? 1. / eta : eta;
        double w1v1 = 0;
        FOR_LANE(i, D) if (i > 0) w1v1 += w[i] * v[i];
        w1v1 = warp_sum(w1v1);
        const double v0 = v[0], w0 = w[0];
        const double f = sg * v0 + w1v1 / (1. + w0);
        warp_sync();
        FOR_LANE(i, D) o[i] = (i == 0) ? scl * (w0 * v0 + sg * w1v1) : scl * (v[i] + f * w[i]);
        warp_sync();
    }
    // scaled step-to-boundary measures of two directions at once: returns max(t(d1), t(d2))
    SCPP_HD double tr_step2(const double *lm, const double *d1, const double *d2) const
    {
        double l1 = 0, a1 = 0, a2 = 0;
        FOR_LANE(i, D) if (i > 0) { l1 += lm[i] * lm[i]; a1 += lm[i] * d1[i]; a2 += lm[i] * d2[i]; }
        warp_sum3(l1, a1, a2);
        const double ia = 1. / sqrt(lm[0] * lm[0] - l1), l0 = lm[0] * ia;
        const double ld1 = l0 * d1[0] - a1 * ia, ld2 = l0 * d2[0] - a2 * ia;
        const double il = ia / (l0 + 1.);
        const double f1 = (ld1 + d1[0]) * il, f2 = (ld2 + d2[0]) * il;
        double n1 = 0, n2 = 0, dummy = 0;
        FOR_LANE(i, D) if (i > 0) { const double r1 = d1[i] - f1 * lm[i], r2 = d2[i] - f2 * lm[i]; n1 += r1 * r1; n2 += r2 * r2; }
        warp_sum3(n1, n2, dummy);
        return fmax((sqrt(n1) - ld1) * ia, (sqrt(n2) - ld2) * ia);
    }
    SCPP_HD void tr_jprod(const double *u, const double *v, double *o) const
    {
        double dot = 0;
        FOR_LANE(i, D) dot += u[i] * v[i];
        dot = warp_sum(dot);
        const double u0 = u[0], v0 = v[0];
        warp_sync();
        FOR_LANE(i, D) o[i] = (i == 0) ? dot : u0 * v[i] + v0 * u[i];
        warp_sync();
    }

    // =============================================================================================================
    //  Phase R : residuals, Nesterov-Todd scaling, termination quantities  (one forward sweep)
    // =============================================================================================================
    struct Norms { double gap, rz2, rx2, pcost, zrz, xrx, h2; int bad; };
    SCPP_HD static double nudge(double u0, double n1) { const double thr = 4e-16 * (fabs(u0) + n1) + 1e-300; return (u0 - n1 > thr) ? u0 : n1 + thr; }

    //  `step` != 0 fuses the update of the previous iteration into this sweep:  (prim, s, z) += step * (dprim, ds, dz)
    //  is applied to each stage window as it is loaded (and written back) before the residuals are taken.
    SCPP_HD void phase_residuals(Norms &nm, bool identity, double step = 0.)
    {
        double gap = 0, rz2 = 0, pcost = 0, zrz = 0, h2 = 0, rx2 = 0, xrx = 0, acc_sig = 0;
        int bad = 0;
        double *S = row(0), *Z = row(1), *RZ = row(2), *WB = row(3), *LM = row(4), *DSW = row(5), *DZW = row(6), *DPW = row(7);
        double *P = pw(0), *PNX = pw(1), *RX = pw(2), *XB = pw(3), *CE = sc();
        double *carry = vec(0), *w = xv(0), *DPN = vec(2);
        const bool upd = step != 0.;
        if (upd) {
            if (lane_id() == 0) {
                const int r0 = K * RS, p0 = K * PS;
                for (int i = 0; i < 4; i++) { s[r0 + i] += step * ds[r0 + i]; z[r0 + i] += step * dz[r0 + i]; }
                prim[p0] += step * dprim[p0]; prim[p0 + 1] += step * dprim[p0 + 1];
                s[r0] = nudge(s[r0], 0.); z[r0] = nudge(z[r0], 0.);
                s[r0 + 1] = nudge(s[r0 + 1], sqrt(s[r0 + 2] * s[r0 + 2] + s[r0 + 3] * s[r0 + 3]));
                z[r0 + 1] = nudge(z[r0 + 1], sqrt(z[r0 + 2] * z[r0 + 2] + z[r0 + 3] * z[r0 + 3]));
            }
            warp_sync();
        }
        const double sg = prim[K * PS];
        FOR_LANE(j, NB) carry[j] = 0.;
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            if (hasint) { ld_dd(k); ld(PNX, prim + (k + 1) * PS, PS); if (upd) ld(DPN, dprim + (k + 1) * PS, PS); }
            ld(S, s + k * RS, RS); ld(Z, z + k * RS, RS); ld(P, prim + k * PS, PS);
            if (upd) { ld(DSW, ds + k * RS, RS); ld(DZW, dz + k * RS, RS); ld(DPW, dprim + k * PS, PS); }
            load_xibar(k, XB);
            tables_stage(k);
            ld_wait();
            if (upd) {
                update_window(step, S, Z, DSW, DZW, P, DPW, hasint);
                if (hasint) { FOR_LANE(e, PN + NX) PNX[e] += step * DPN[e]; }
                warp_sync();
                st(s + k * RS, S, RS); st(z + k * RS, Z, RS); st(prim + k * PS, P, PN + NX);
            }
            // ---- trust-region cone: warp-cooperative
            {
                const int o = TRO;
                FOR_LANE(i, D) {
                    const double sl = (i == 0) ? P[NB] : XB[i - 1] - P[i - 1];
                    RZ[o + i] = S[o + i] - sl;
                    if (i > 0) h2 += XB[i - 1] * XB[i - 1];
                }
                if (lane_id() == 0) pcost += w_tr * P[NB];
                if (identity) { FOR_LANE(i, D) { WB[o + i] = i == 0; LM[o + i] = i == 0; } if (lane_id() == 0) CE[NCONE] = 1.; }
                else {
                    double e2i;
                    if (!tr_scale(S + o, Z + o, WB + o, e2i, LM + o)) bad = 1;
                    else if (lane_id() == 0) CE[NCONE] = e2i;
                }
                FOR_LANE(i, D) { gap += S[o + i] * Z[o + i]; rz2 += RZ[o + i] * RZ[o + i]; zrz += Z[o + i] * RZ[o + i]; }
            }
            // ---- small cone tasks: one lane each
            FOR_LANE(tk, NTASK) {
                if (tk == NCONE) continue;
                int type, o, d, ci;
                task(tk, type, o, d, ci);
                if (type == 1) {
#pragma unroll 1
                    for (int r = 0; r < d; r++) { const double hh = row_h(o + r); const double sl = hh - model_G(o + r, k, P); RZ[o + r] = S[o + r] - sl; h2 += hh * hh; }
                    if (identity) { CE[ci] = 1.; for (int i = 0; i < d; i++) { WB[o + i] = i == 0; LM[o + i] = i == 0; } }
                    else if (!soc::scale(S + o, Z + o, d, WB + o, CE[ci], LM + o)) bad = 1;
#pragma unroll 1
                    for (int r = 0; r < d; r++) { gap += S[o + r] * Z[o + r]; rz2 += RZ[o + r] * RZ[o + r]; zrz += Z[o + r] * RZ[o + r]; }
                } else if (type == 0) {
                    const double hh = row_h(o);
                    const double sl = hh - model_G(o, k, P);
                    const double sv = S[o], zv = Z[o];
                    RZ[o] = sv - sl;
                    h2 += hh * hh;
                    if (!(sv > 0.) || !(zv > 0.)) bad = 1;
                    WB[o] = identity ? 1. : zv / sv; LM[o] = identity ? 1. : sqrt(sv * zv);
                    gap += sv * zv; rz2 += RZ[o] * RZ[o]; zrz += zv * RZ[o];
                } else if (hasint) {
                    const int i = o - MN;
                    const double r = dyn_row(i, P, PNX, sg, true), t = P[PN + i];
                    const double sm_ = S[o], sp = S[o + NX], zm = Z[o], zp = Z[o + NX];
                    const double rm = sm_ - (t - r), rp = sp - (t + r);
                    RZ[o] = rm; RZ[o + NX] = rp;
                    if (!(sm_ > 0.) || !(sp > 0.) || !(zm > 0.) || !(zp > 0.)) bad = 1;
                    WB[o] = identity ? 1. : zm / sm_; WB[o + NX] = identity ? 1. : zp / sp;
                    LM[o] = identity ? 1. : sqrt(sm_ * zm); LM[o + NX] = identity ? 1. : sqrt(sp * zp);
                    RX[PN + i] = w_vc - zm - zp;
                    w[i] = zm - zp;
                    gap += sm_ * zm + sp * zp; rz2 += rm * rm + rp * rp; zrz += zm * rm + zp * rp;
                    pcost += w_vc * t;
                    const double zc = tile()[i * NCP + NB + NU + 1];
                    h2 += 2. * zc * zc;
                } else {
                    const int i = o - MN;
                    RZ[o] = 0.; RZ[o + NX] = 0.; WB[o] = 1.; WB[o + NX] = 1.; LM[o] = 1.; LM[o + NX] = 1.; RX[PN + i] = 0.;
                }
            }
            warp_sync();
            // ---- dual residual of node k: carry from interval k-1 + G'z of the node cones (+ interval k below)
            FOR_LANE(j, NB) RX[j] = carry[j] + Z[TRO + 1 + j] + model_GT(j, k, Z);
            if (lane_id() == 0) RX[NB] = w_tr - Z[TRO];
            warp_sync();
            if (hasint) acc_sig += dyn_JT(w, RX, carry);
            warp_sync();
            FOR_LANE(e, PN + NX) {
                if (e < NB && fixed(k, e)) RX[e] = 0.;
                rx2 += RX[e] * RX[e]; xrx += P[e] * RX[e];
            }
            warp_sync();
            st(rz + k * RS, RZ, RS); st(wb + k * RS, WB, RS); st(lam + k * RS, LM, RS);
            st(rx + k * PS, RX, PN + NX);
            FOR_LANE(c, NCN) ce[k * CS + c] = CE[c];
            warp_sync();
        }
        acc_sig = warp_sum(acc_sig);
        // ---- globals: lane 0
        if (lane_id() == 0) {
            const int r0 = K * RS, c0 = K * CS, p0 = K * PS;
            const double dsg = prim[p0 + 1];
            double rxs = w_time + acc_sig;
            rz[r0] = s[r0] - (sg - 0.001);                          // sigma >= 0.001   (SCProblem.cpp:34)
            rxs -= z[r0];
            if (!(s[r0] > 0.) || !(z[r0] > 0.)) bad = 1;
            wb[r0] = identity ? 1. : z[r0] / s[r0]; lam[r0] = identity ? 1. : sqrt(s[r0] * z[r0]);
            rz[r0 + 1] = s[r0 + 1] - (0.5 + 0.5 * dsg);            // ((1+dsg)/2 ; (1-dsg)/2 ; sigma - sigbar)   (:92-96)
            rz[r0 + 2] = s[r0 + 2] - (0.5 - 0.5 * dsg);
            rz[r0 + 3] = s[r0 + 3] - (sg - sigbar);
            rxs -= z[r0 + 3];
            rx[p0 + 1] = w_trs - 0.5 * z[r0 + 1] + 0.5 * z[r0 + 2];
            rx[p0] = rxs;
            if (identity) { ce[c0] = 1.; for (int i = 0; i < 3; i++) { wb[r0 + 1 + i] = i == 0; lam[r0 + 1 + i] = i == 0; } }
            else if (!soc::scale(s + r0 + 1, z + r0 + 1, 3, wb + r0 + 1, ce[c0], lam + r0 + 1)) bad = 1;
            for (int r = 0; r < 4; r++) { gap += s[r0 + r] * z[r0 + r]; rz2 += rz[r0 + r] * rz[r0 + r]; zrz += z[r0 + r] * rz[r0 + r]; }
            pcost += w_time * sg + w_trs * dsg;
            h2 += 0.001 * 0.001 + 0.5 + sigbar * sigbar;
            rx2 += rx[p0] * rx[p0] + rx[p0 + 1] * rx[p0 + 1];
            xrx += sg * rx[p0] + dsg * rx[p0 + 1];
        }
        warp_sync();
        nm.gap = warp_sum(gap); nm.rz2 = warp_sum(rz2); nm.pcost = warp_sum(pcost); nm.zrz = warp_sum(zrz);
        nm.rx2 = warp_sum(rx2); nm.xrx = warp_sum(xrx); nm.h2 = warp_sum(h2); nm.bad = warp_or(bad);
    }

    // =============================================================================================================
    //  Phase F : assemble the reduced Hessian stage by stage and factor it (block-tridiagonal Cholesky + border)
    // =============================================================================================================
    SCPP_HD void build_model_terms(int k, const double *WB, const double *CE, double *alpha)
    {
        double *rk = sm + W_RK, *dg = rk + NRK * NB;
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

    SCPP_HD bool phase_factor()
    {
        double *H = sm + W_MAT, *O = H + BLK, *Lp = sm + W_LP;
        double *Dp = sm + W_HN, *DCp = Dp + NX, *CDCp = DCp + NX * NU;   // carry of interval k-1: D | D C | C' D C
        double *F = facw();                 // Linv | Lnext | l | f   (the record written for this stage)
        double *Li = F, *Ln = F + OFF_LN, *lk = F + OFF_L;
        double *WB = row(0), *CE = sc() + 8, *alpha = sc();
        double *bk = vec(0), *bn = vec(1), *lprev = vec(2);
        double *Dt = xv(1);
        double corner = 0.;
        int bad = 0;
        FOR_LANE(e, BLK) Lp[e] = 0.;
        FOR_LANE(e, HNC) Dp[e] = 0.;
        FOR_LANE(j, NB) { bn[j] = 0.; lprev[j] = 0.; }
        warp_sync();
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            if (hasint) ld_dd(k);
            ld(WB, wb + k * RS, RS); ld(CE, ce + k * CS, CS);
            tables_stage(k);
            ld_wait();
            build_model_terms(k, WB, CE, alpha);
            // ---- node part of H_kk (trust region with delta eliminated) + carry from interval k-1
            {
                const double *rk = sm + W_RK, *dg = rk + NRK * NB;
                const double *wt_ = WB + TRO;
                const double e2i = CE[NCONE], w0 = wt_[0];
                const double kap = e2i * (2. * w0 * w0 - 1.), c2 = 4. * e2i * e2i * w0 * w0 / kap;
#pragma unroll 2
                FOR_LANE(e, BLK) {
                    const int i = e / NB, j = e - i * NB;
                    double v;
                    if (i < NX && j < NX) v = (i == j) ? Dp[i] : 0.;
                    else if (i < NX) v = -DCp[i * NU + (j - NX)];
                    else if (j < NX) v = -DCp[j * NU + (i - NX)];
                    else v = CDCp[(i - NX) * NU + (j - NX)];
#pragma unroll
                    for (int c = 0; c < NRK; c++) v += alpha[c] * rk[c * NB + i] * rk[c * NB + j];
                    if (i == j) {
#pragma unroll
                        for (int c = 0; c < NRK; c++) v += dg[c * NB + i];
                        v += e2i;
                    }
                    v += (2. * e2i - c2) * wt_[1 + i] * wt_[1 + j];
                    H[e] = v;
                }
                FOR_LANE(j, NB) bk[j] = bn[j];
            }
            warp_sync();
            // ---- interval k: H_kk += A~' D A~ ; O = [-D A~ ; C' D A~] ; carry = (D, D C, C' D C) ; borders
            if (hasint) {
                double *t = tile();
                FOR_LANE(i, NX) { const double dm = WB[MN + i], dp = WB[MN + NX + i]; Dt[i] = 4. * dm * dp / (dm + dp); }
                warp_sync();
                // O rows of the x_{k+1} block are -D A~ ; keep D A~ for the products: O[a][b], a < NX
                FOR_LANE(e, NX * NB) { const int a = e / NB, b = e - a * NB; O[a * NB + b] = -Dt[a] * t[a * NCP + b]; }
                FOR_LANE(e, NX * NU) { const int i = e / NU, j = e - i * NU; DCp[e] = Dt[i] * t[i * NCP + NB + j]; }
                FOR_LANE(i, NX) Dp[i] = Dt[i];
                warp_sync();
                // H += A~' (D A~) = -A~' O_x   and   O_u = C' D A~ = -C' O_x   on the FP64 tensor cores (lower tiles of H suffice)
                blk::mm<NB, NB, NX, true>([&](int m, int k) { return (m < NB && k < NX) ? t[k * NCP + m] : 0.; },
                                          [&](int k, int n) { return (k < NX && n < NB) ? -O[k * NB + n] : 0.; },
                                          [&](int m, int n, double v) { if (m < NB && n < NB) H[m * NB + n] += v; });
                blk::mm<NU, NB, NX, false>([&](int m, int k) { return (m < NU && k < NX) ? t[k * NCP + NB + m] : 0.; },
                                           [&](int k, int n) { return (k < NX && n < NB) ? -O[k * NB + n] : 0.; },
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
            // ---- Schur update with the previous off-diagonal factor: H -= Lp Lp' ; bk -= Lp lprev
            if (k > 0) {
                blk::mm<NB, NB, NB, true>([&](int m, int k) { return (m < NB && k < NB) ? Lp[m * NB + k] : 0.; },
                                          [&](int k, int n) { return (k < NB && n < NB) ? Lp[n * NB + k] : 0.; },
                                          [&](int m, int n, double v) { if (m < NB && n < NB) H[m * NB + n] -= v; });
                FOR_LANE(j, NB) {
                    double v = 0;
#pragma unroll 2
                    for (int c = 0; c < NB; c++) v += Lp[j * NB + c] * lprev[c];
                    bk[j] -= v;
                }
            }
            warp_sync();
            // ---- Cholesky of H (lower, in place), left-looking by columns: lane i owns row i
#pragma unroll 1
            for (int j = 0; j < NB; j++) {
                FOR_LANE(i, NB) if (i >= j) {
                    double v0 = H[i * NB + j], v1 = 0;
                    int c = 0;
#pragma unroll 1
                    for (; c + 1 < j; c += 2) { v0 -= H[i * NB + c] * H[j * NB + c]; v1 -= H[i * NB + c + 1] * H[j * NB + c + 1]; }
                    if (c < j) v0 -= H[i * NB + c] * H[j * NB + c];
                    H[i * NB + j] = v0 + v1;         // unscaled column j (row j reads only columns < j of itself)
                }
                warp_sync();
                const double djj = H[j * NB + j];
                if (!(djj > 0.)) bad = 1;
                const double inv = 1. / sqrt(djj > 0. ? djj : 1.);
                warp_sync();
                FOR_LANE(i, NB) if (i >= j) H[i * NB + j] *= inv;
                warp_sync();
            }
            // ---- Linv = L^-1 (lower): lane per column
            FOR_LANE(c, NB) {
#pragma unroll 1
                for (int i = 0; i < NB; i++) {
                    if (i < c) { Li[i * NB + c] = 0.; continue; }
                    double v = (i == c) ? 1. : 0., v2 = 0.;
                    int q = c;
#pragma unroll 1
                    for (; q + 1 < i; q += 2) { v -= H[i * NB + q] * Li[q * NB + c]; v2 -= H[i * NB + q + 1] * Li[(q + 1) * NB + c]; }
                    if (q < i) v -= H[i * NB + q] * Li[q * NB + c];
                    Li[i * NB + c] = (v + v2) / H[i * NB + i];
                }
            }
            warp_sync();
            // ---- L_{k+1,k} = O Linv' ;  l_k = Linv bk ; corner -= l_k' l_k
            blk::mm<NB, NB, NB, false>([&](int m, int k) { return (m < NB && k < NB) ? O[m * NB + k] : 0.; },
                                       [&](int k, int n) { return (k < NB && n < NB) ? Li[n * NB + k] : 0.; },
                                       [&](int m, int n, double v) { if (m < NB && n < NB) Ln[m * NB + n] = v; });
            FOR_LANE(j, NB) {
                double v = 0, v2 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { v += Li[j * NB + c] * bk[c]; v2 += Li[j * NB + c + 1] * bk[c + 1]; }
                v += v2;
                lk[j] = v; corner -= v * v;
            }
            warp_sync();
            st(fac + (size_t)k * FS, F, OFF_F);
            FOR_LANE(e, BLK) Lp[e] = Ln[e];
            FOR_LANE(j, NB) lprev[j] = lk[j];
            warp_sync();
        }
        corner = warp_sum(corner);
        {   // globals: sigma >= 0.001 row and the sigma trust-region cone with delta_sigma eliminated
            const int r0 = K * RS;
            const double d = wb[r0];
            double w3[3] = {wb[r0 + 1], wb[r0 + 2], wb[r0 + 3]};
            const double e2i = ce[K * CS];
            double g[3] = {-0.5, 0.5, 0.}, p[3], e2[3] = {0., 0., 1.}, m2[3];
            soc::Mv(w3, e2i, g, 3, p);
            const double kap = g[0] * p[0] + g[1] * p[1];
            soc::Mv(w3, e2i, e2, 3, m2);
            corner += d + (m2[2] - p[2] * p[2] / kap);
        }
        if (!(corner > 0.)) bad = 1;
        l_ss = sqrt(corner > 0. ? corner : 1.);
        return !warp_or(bad);
    }

    // =============================================================================================================
    //  Phase S : solve the Newton system  G'dz = rxv ,  G dx - W^2 dz = rzv  with the local variables eliminated.
    //     mode 0: rxv = dprim (array), rzv = ds (array)                        [starting point]
    //     mode 1: rxv = -rx,            rzv = -rz + s                          [affine direction]
    //     mode 2: rxv = -(1-sig) rx,    rzv = -(1-sig) rz - W (lam \ d_s),  d_s = -lam o lam - cr + sig mu e   [combined]
    //  forward sweep: right-hand side + forward substitution; backward sweep: back substitution + recovery of the
    //  local variables, dz and ds = rzs*rz - G dx, the scaled step lengths (tmax) and, for mode 1, cr = ds~ o dz~.
    // =============================================================================================================
    SCPP_HD void gen_rhs(int mode, bool hasint, double csig, double sigmu, double *RZV, double *RXV, double *T1,
                         const double *S_, const double *RZ, const double *LM, const double *CR, const double *WB, const double *CE, const double *RXW)
    {
        if (mode == 1) {
            FOR_LANE(e, RS) RZV[e] = -RZ[e] + S_[e];
            FOR_LANE(e, PS) RXV[e] = -RXW[e];
        } else {
            FOR_LANE(e, PS) RXV[e] = -csig * RXW[e];
            {   // trust-region cone (warp-cooperative): rzv = -csig rz - W (lam \ (-lam o lam - cr + sigmu e))
                const int o = TRO;
                const double *lm = LM + o, *w = WB + o;
                double ll = 0;
                FOR_LANE(i, D) ll += lm[i] * lm[i];
                ll = warp_sum(ll);
                const double l0 = lm[0], den = 2. * l0 * l0 - ll;
                double l1d1 = 0;
                FOR_LANE(i, D) {
                    const double dv = (i == 0) ? -ll - CR[o] + sigmu : -2. * l0 * lm[i] - CR[o + i];
                    T1[o + i] = dv;
                    if (i > 0) l1d1 += lm[i] * dv;
                }
                l1d1 = warp_sum(l1d1);
                warp_sync();
                const double x0 = (l0 * T1[o] - l1d1) / den;
                warp_sync();
                FOR_LANE(i, D) T1[o + i] = (i == 0) ? x0 : (T1[o + i] - x0 * lm[i]) / l0;
                warp_sync();
                tr_Wv(w, CE[NCONE], T1 + o, T1 + o, false);
                FOR_LANE(i, D) RZV[o + i] = -csig * RZ[o + i] - T1[o + i];
            }
            FOR_LANE(tk, NTASK) {
                if (tk == NCONE) continue;
                int type, o, d, ci;
                task(tk, type, o, d, ci);
                if (type == 1) {
                    double *t1 = T1 + o;
                    soc::jprod(LM + o, LM + o, d, t1);
#pragma unroll 1
                    for (int i = 0; i < d; i++) t1[i] = -t1[i] - CR[o + i];
                    t1[0] += sigmu;
                    soc::jdiv(LM + o, t1, d, t1);
                    soc::Wv(WB + o, CE[ci], t1, d, t1, false);
#pragma unroll 1
                    for (int i = 0; i < d; i++) RZV[o + i] = -csig * RZ[o + i] - t1[i];
                } else {
                    const int reps = (type == 2) ? 2 : 1;
                    if (type == 2 && !hasint) { RZV[o] = 0.; RZV[o + NX] = 0.; continue; }
                    for (int q = 0; q < reps; q++) {
                        const int oo = o + q * NX;
                        const double wv = sqrt(1. / WB[oo]);
                        const double t1 = (-LM[oo] * LM[oo] - CR[oo] + sigmu) / LM[oo];
                        RZV[oo] = -csig * RZ[oo] - wv * t1;
                    }
                }
            }
        }
    }

    SCPP_HD void phase_solve(int mode, double csig, double sigmu, double rzs, double &tmax_out)
    {
        double *WB = row(0), *RZV = row(1), *V = row(2), *RZ = row(3), *LM = row(4), *CR = row(5), *S_ = row(6), *DZ = row(7);
        double *RXV = pw(0), *RXW = pw(1), *DP = pw(2), *CE = sc() + 8;
        double *F = facw();
        double *carry = vec(0), *g = vec(1), *tmp = vec(2), *fprev = vec(3), *Lp = sm + W_LP;   // Lp: L_{k,k-1}
        double *w = xv(0);
        const int r0g = K * RS, p0g = K * PS;
        // ---------------- forward sweep ----------------
        double gsig = 0, ldot = 0;
        FOR_LANE(j, NB) { carry[j] = 0.; fprev[j] = 0.; }
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            if (hasint) ld_dd(k);
            ld(F, fac + (size_t)k * FS, OFF_F);
            ld(WB, wb + k * RS, RS); ld(CE, ce + k * CS, CS);
            if (mode == 0) { ld(RZV, ds + k * RS, RS); ld(RXV, dprim + k * PS, PS); }
            else {
                ld(RZ, rz + k * RS, RS); ld(RXW, rx + k * PS, PS);
                if (mode == 1) ld(S_, s + k * RS, RS);
                else { ld(LM, lam + k * RS, RS); ld(CR, cr + k * RS, RS); }
            }
            tables_stage(k);
            ld_wait();
            if (mode != 0) { gen_rhs(mode, hasint, csig, sigmu, RZV, RXV, V, S_, RZ, LM, CR, WB, CE, RXW); warp_sync(); }
            // ---- v_c = Mtilde rzv per cone; pairs produce w_i
            {   // trust region (warp-cooperative): v = M rzv - p (p'rzv + rx_delta)/kap ,  p = M(-e0)
                const int o = TRO;
                const double *w = WB + o;
                const double e2i = CE[NCONE], w0 = w[0];
                tr_Mv(w, e2i, RZV + o, V + o);
                double prz = 0;
                FOR_LANE(i, D) { const double p = (i == 0) ? -e2i * (2. * w0 * w0 - 1.) : 2. * e2i * w0 * w[i]; prz += p * RZV[o + i]; }
                prz = warp_sum(prz);
                const double rho = (prz + RXV[NB]) / (e2i * (2. * w0 * w0 - 1.));
                FOR_LANE(i, D) { const double p = (i == 0) ? -e2i * (2. * w0 * w0 - 1.) : 2. * e2i * w0 * w[i]; V[o + i] -= p * rho; }
            }
            FOR_LANE(tk, NTASK) {
                if (tk == NCONE) continue;
                int type, o, d, ci;
                task(tk, type, o, d, ci);
                if (type == 1) soc::Mv(WB + o, CE[ci], RZV + o, d, V + o);
                else if (type == 0) V[o] = WB[o] * RZV[o];
                else if (hasint) {
                    const int i = o - MN;
                    const double dm = WB[o], dp = WB[o + NX], rm = RZV[o], rp = RZV[o + NX];
                    const double rho = (-(dm * rm + dp * rp) + RXV[PN + i]) / (dm + dp);
                    w[i] = dm * (rm + rho) - dp * (rp + rho);
                }
            }
            warp_sync();
            FOR_LANE(j, NB) g[j] = fixed(k, j) ? 0. : RXV[j] + carry[j] + V[TRO + 1 + j] + model_GT(j, k, V);
            warp_sync();
            if (hasint) gsig += dyn_JT(w, g, carry);
            warp_sync();
            // forward substitution: f_k = Linv (g_k - L_{k,k-1} f_{k-1})
            FOR_LANE(j, NB) {
                double v = fixed(k, j) ? 0. : g[j];
                if (k > 0) {
#pragma unroll
                    for (int c = 0; c < NB; c++) v -= Lp[j * NB + c] * fprev[c];
                }
                tmp[j] = v;
            }
            warp_sync();
            FOR_LANE(j, NB) {
                double v = 0, v2 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { v += F[j * NB + c] * tmp[c]; v2 += F[j * NB + c + 1] * tmp[c + 1]; }
                v += v2;
                F[OFF_F + j] = v;
                ldot += F[OFF_L + j] * v;
            }
            warp_sync();
            FOR_LANE(j, NB) { fprev[j] = F[OFF_F + j]; fac[(size_t)k * FS + OFF_F + j] = F[OFF_F + j]; }
            FOR_LANE(e, BLK) Lp[e] = F[OFF_LN + e];
            if (mode != 0) st(ds + k * RS, RZV, RS);
            warp_sync();
        }
        gsig = warp_sum(gsig);
        // ---- globals: rhs and local elimination for the sigma rows (every lane computes the same scalars)
        double kap_s, p_s[3], rzg[4], rxg[2];
        {
            double w3[3] = {wb[r0g + 1], wb[r0g + 2], wb[r0g + 3]};
            const double e2i = ce[K * CS], d0 = wb[r0g];
            if (mode == 0) { for (int i = 0; i < 4; i++) rzg[i] = ds[r0g + i]; rxg[0] = dprim[p0g]; rxg[1] = dprim[p0g + 1]; }
            else if (mode == 1) { for (int i = 0; i < 4; i++) rzg[i] = -rz[r0g + i] + s[r0g + i]; rxg[0] = -rx[p0g]; rxg[1] = -rx[p0g + 1]; }
            else {
                const double l0 = lam[r0g];
                rzg[0] = -csig * rz[r0g] - sqrt(1. / d0) * ((-l0 * l0 - cr[r0g] + sigmu) / l0);
                double lm3[3] = {lam[r0g + 1], lam[r0g + 2], lam[r0g + 3]}, t1[3];
                soc::jprod(lm3, lm3, 3, t1);
                for (int i = 0; i < 3; i++) t1[i] = -t1[i] - cr[r0g + 1 + i];
                t1[0] += sigmu;
                soc::jdiv(lm3, t1, 3, t1);
                soc::Wv(w3, e2i, t1, 3, t1, false);
                for (int i = 0; i < 3; i++) rzg[1 + i] = -csig * rz[r0g + 1 + i] - t1[i];
                rxg[0] = -csig * rx[p0g]; rxg[1] = -csig * rx[p0g + 1];
            }
            double g3[3] = {-0.5, 0.5, 0.}, v[3];
            soc::Mv(w3, e2i, g3, 3, p_s);
            kap_s = g3[0] * p_s[0] + g3[1] * p_s[1];
            soc::Mv(w3, e2i, rzg + 1, 3, v);
            const double prz = p_s[0] * rzg[1] + p_s[1] * rzg[2] + p_s[2] * rzg[3];
            const double rho = (prz + rxg[1]) / kap_s;
            gsig += rxg[0] - d0 * rzg[0] - (v[2] - p_s[2] * rho);
        }
        const double fsig = (gsig - warp_sum(ldot)) / l_ss;
        const double ysig = fsig / l_ss;
        warp_sync();
        // ---------------- backward sweep ----------------
        double tmax = 0;
        double *ynext = vec(3), *yk = vec(4), *DS = RZV;
        double *XN = pw(3);                 // d xi_{k+1}
        FOR_LANE(j, NB) ynext[j] = 0.;
#pragma unroll 1
        for (int k = K - 1; k >= 0; k--) {
            const bool hasint = k < K - 1;
            if (hasint) ld_dd(k);
            ld(F, fac + (size_t)k * FS, FS);
            ld(WB, wb + k * RS, RS); ld(CE, ce + k * CS, CS); ld(RZV, ds + k * RS, RS);
            if (mode == 0) ld(RXV, dprim + k * PS, PS);
            else { ld(RXW, rx + k * PS, PS); ld(RZ, rz + k * RS, RS); ld(LM, lam + k * RS, RS); }
            tables_stage(k);
            ld_wait();
            if (mode != 0) { FOR_LANE(e, PS) RXV[e] = (mode == 1 ? -1. : -csig) * RXW[e]; }
            // back substitution: y_k = Linv' (f_k - L_{k+1,k}' y_{k+1} - l_k y_sigma)
            FOR_LANE(j, NB) {
                double v = F[OFF_F + j] - F[OFF_L + j] * ysig;
                if (hasint) {
#pragma unroll
                    for (int c = 0; c < NB; c++) v -= F[OFF_LN + c * NB + j] * ynext[c];
                }
                tmp[j] = v;
                XN[j] = ynext[j];
            }
            warp_sync();
            FOR_LANE(j, NB) {
                double v = 0, v2 = 0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) { v += F[c * NB + j] * tmp[c]; v2 += F[(c + 1) * NB + j] * tmp[c + 1]; }
                v += v2;
                yk[j] = fixed(k, j) ? 0. : v;
            }
            warp_sync();
            // ---- recovery per cone: q = G dy - rzv ; local ; dz = M q + p dl ; ds = rzs rz - G dx ; scaled steps
            //      scratch: Q = V window, GDX = S_ window (both free in the backward sweep)
            double *Q = V, *GDX = S_;
            {   // trust region (warp-cooperative)
                const int o = TRO;
                const double *w = WB + o;
                const double e2i = CE[NCONE], w0 = w[0], kap = e2i * (2. * w0 * w0 - 1.);
                double pq = 0;
                FOR_LANE(i, D) {
                    const double q = (i == 0) ? -RZV[o] : yk[i - 1] - RZV[o + i];
                    const double p = (i == 0) ? -kap : 2. * e2i * w0 * w[i];
                    Q[o + i] = q; pq += p * q;
                }
                pq = warp_sum(pq);
                const double ddl = (RXV[NB] - pq) / kap;
                warp_sync();
                tr_Mv(w, e2i, Q + o, DZ + o);
                FOR_LANE(i, D) {
                    const double p = (i == 0) ? -kap : 2. * e2i * w0 * w[i];
                    DZ[o + i] += p * ddl;
                    if (mode != 0) DS[o + i] = rzs * RZ[o + i] - ((i == 0) ? -ddl : yk[i - 1]);
                }
                if (lane_id() == 0) DP[NB] = ddl;
                warp_sync();
                if (mode != 0) {
                    tr_Wv(w, e2i, DZ + o, Q + o, false);          // dz~
                    tr_Wv(w, e2i, DS + o, GDX + o, true);         // ds~
                    tmax = fmax(tmax, tr_step2(LM + o, GDX + o, Q + o));
                    if (mode == 1) tr_jprod(GDX + o, Q + o, CR + o);
                }
            }
            FOR_LANE(tk, NTASK) {
                if (tk == NCONE) continue;
                int type, o, d, ci;
                task(tk, type, o, d, ci);
                if (type == 1) {
#pragma unroll 1
                    for (int r = 0; r < d; r++) { const double gdx = model_G(o + r, k, yk); GDX[o + r] = gdx; Q[o + r] = gdx - RZV[o + r]; }
                    soc::Mv(WB + o, CE[ci], Q + o, d, DZ + o);
                    if (mode != 0) {
#pragma unroll 1
                        for (int r = 0; r < d; r++) DS[o + r] = rzs * RZ[o + r] - GDX[o + r];
                        soc::Wv(WB + o, CE[ci], DZ + o, d, Q + o, false);
                        soc::Wv(WB + o, CE[ci], DS + o, d, GDX + o, true);
                        tmax = fmax(tmax, fmax(soc::step(LM + o, GDX + o, d), soc::step(LM + o, Q + o, d)));
                        if (mode == 1) soc::jprod(GDX + o, Q + o, d, CR + o);
                    }
                } else if (type == 0) {
                    const double gdx = model_G(o, k, yk);
                    DZ[o] = WB[o] * (gdx - RZV[o]);
                    if (mode != 0) {
                        DS[o] = rzs * RZ[o] - gdx;
                        const double iw = sqrt(WB[o]), il = 1. / LM[o];      // W = 1/sqrt(wb)
                        const double dzt = DZ[o] / iw, dst = DS[o] * iw;
                        tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                        if (mode == 1) CR[o] = dst * dzt;
                    }
                } else if (hasint) {
                    const int i = o - MN;
                    const double ady = dyn_row(i, yk, XN, ysig, false);
                    const double dm = WB[o], dp = WB[o + NX];
                    const double qm = ady - RZV[o], qp = -ady - RZV[o + NX];
                    const double dt = (RXV[PN + i] + dm * qm + dp * qp) / (dm + dp);
                    DZ[o] = dm * (qm - dt); DZ[o + NX] = dp * (qp - dt);
                    DP[PN + i] = dt;
                    if (mode != 0) {
                        DS[o] = rzs * RZ[o] - (ady - dt); DS[o + NX] = rzs * RZ[o + NX] - (-ady - dt);
                        for (int q = 0; q < 2; q++) {
                            const int oo = o + q * NX;
                            const double iw = sqrt(WB[oo]), il = 1. / LM[oo];
                            const double dzt = DZ[oo] / iw, dst = DS[oo] * iw;
                            tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                            if (mode == 1) CR[oo] = dst * dzt;
                        }
                    }
                } else {
                    const int i = o - MN;
                    DP[PN + i] = 0.; DZ[o] = 0.; DZ[o + NX] = 0.;
                    if (mode != 0) { DS[o] = 0.; DS[o + NX] = 0.; if (mode == 1) { CR[o] = 0.; CR[o + NX] = 0.; } }
                }
            }
            FOR_LANE(j, NB) DP[j] = yk[j];
            warp_sync();
            st(dz + k * RS, DZ, RS); st(dprim + k * PS, DP, PN + NX);
            if (mode != 0) { st(ds + k * RS, DS, RS); if (mode == 1) st(cr + k * RS, CR, RS); }
            FOR_LANE(j, NB) ynext[j] = yk[j];
            warp_sync();
        }
        // ---- globals recovery (lane 0)
        if (lane_id() == 0) {
            double w3[3] = {wb[r0g + 1], wb[r0g + 2], wb[r0g + 3]};
            const double e2i = ce[K * CS], d0 = wb[r0g];
            const double dz0 = d0 * (-ysig - rzg[0]);
            double q[3] = {-rzg[1], -rzg[2], -ysig - rzg[3]}, mq[3], dzq[3];
            const double pq = p_s[0] * q[0] + p_s[1] * q[1] + p_s[2] * q[2];
            const double dds = (rxg[1] - pq) / kap_s;
            soc::Mv(w3, e2i, q, 3, mq);
            for (int i = 0; i < 3; i++) dzq[i] = mq[i] + p_s[i] * dds;
            dz[r0g] = dz0; for (int i = 0; i < 3; i++) dz[r0g + 1 + i] = dzq[i];
            dprim[p0g] = ysig; dprim[p0g + 1] = dds;
            if (mode != 0) {
                double dsg[4] = {rzs * rz[r0g] + ysig, rzs * rz[r0g + 1] + 0.5 * dds, rzs * rz[r0g + 2] - 0.5 * dds, rzs * rz[r0g + 3] + ysig};
                for (int i = 0; i < 4; i++) ds[r0g + i] = dsg[i];
                const double wv = sqrt(1. / d0), l0 = lam[r0g];
                const double dzt0 = wv * dz0, dst0 = dsg[0] / wv;
                tmax = fmax(tmax, fmax(-dst0 / l0, -dzt0 / l0));
                double lm3[3] = {lam[r0g + 1], lam[r0g + 2], lam[r0g + 3]}, dzt[3], dst[3], pr[3];
                soc::Wv(w3, e2i, dzq, 3, dzt, false);
                soc::Wv(w3, e2i, dsg + 1, 3, dst, true);
                tmax = fmax(tmax, fmax(soc::step(lm3, dst, 3), soc::step(lm3, dzt, 3)));
                if (mode == 1) { cr[r0g] = dst0 * dzt0; soc::jprod(dst, dzt, 3, pr); for (int i = 0; i < 3; i++) cr[r0g + 1 + i] = pr[i]; }
            }
        }
        warp_sync();
        tmax_out = warp_max(tmax);
    }

    // =============================================================================================================
    //  light sweeps: slack evaluation, cone margins / shifts, the update
    // =============================================================================================================
    // out = h - G x  (stage sweep)
    SCPP_HD void eval_slack(double *out)
    {
        double *P = pw(0), *PNX = pw(1), *XB = pw(3), *O_ = row(0);
        const double sg = prim[K * PS], dsg = prim[K * PS + 1];
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            if (hasint) { ld_dd(k); ld(PNX, prim + (k + 1) * PS, PS); }
            ld(P, prim + k * PS, PS);
            load_xibar(k, XB);
            tables_stage(k);
            ld_wait();
            FOR_LANE(r, RS) {
                double v = 0.;
                if (r < NLP + NCR) v = row_h(r) - model_G(r, k, P);
                else if (r == TRO) v = P[NB];
                else if (r < MN) v = XB[r - TRO - 1] - P[r - TRO - 1];
                else if (r < MN + 2 * NX && hasint) {
                    const int i = (r - MN) % NX;
                    const double rr = dyn_row(i, P, PNX, sg, true), t = P[PN + i];
                    v = (r - MN < NX) ? t - rr : t + rr;
                }
                O_[r] = v;
            }
            warp_sync();
            st(out + k * RS, O_, RS);
            warp_sync();
        }
        if (lane_id() == 0) { const int r0 = K * RS; out[r0] = sg - 0.001; out[r0 + 1] = 0.5 + 0.5 * dsg; out[r0 + 2] = 0.5 - 0.5 * dsg; out[r0 + 3] = sg - sigbar; }
        warp_sync();
    }
    // visit every cone of the flat row array (used only by the two start-up shifts)
    template <class F>
    SCPP_HD void for_cones(F &&f) const
    {
        FOR_LANE(e, K * NTASK) {
            const int k = e / NTASK, tk = e - k * NTASK;
            int type, o, d, ci;
            task(tk, type, o, d, ci);
            if (type == 2) { if (k < K - 1) { f(k * RS + o, 1); f(k * RS + o + NX, 1); } }
            else f(k * RS + o, d);
        }
        if (lane_id() == 0) { f(K * RS, 1); f(K * RS + 1, 3); }
    }
    SCPP_HD void cone_margin(const double *u, double &mn, double &nrm2) const
    {
        double lmn = 1e300, n2 = 0;
        for_cones([&](int o, int d) {
            double t = 0;
            for (int i = 1; i < d; i++) t += u[o + i] * u[o + i];
            const double mg = u[o] - sqrt(t);
            if (mg < lmn) lmn = mg;
            n2 += t + u[o] * u[o];
        });
        mn = -warp_max(-lmn); nrm2 = warp_sum(n2);
    }
    SCPP_HD void cone_shift(double *u, double a) const { for_cones([&](int o, int) { u[o] += a; }); }

    // prim += a dprim ; s += a ds ; z += a dz.  The step length keeps every cone 1 % inside in exact arithmetic; a cone whose
    // margin u0 - |u1| is lost to rounding (active to ~1e-16 relative) is nudged back inside by a few ulps of u0 so the next
    // Nesterov-Todd scaling stays defined (perturbation << the 1e-8 tolerances).
    // window update shared by apply_step (stand-alone) and phase_residuals (fused): S += a DS, Z += a DZ, P += a DP, then nudge
    SCPP_HD void update_window(double a, double *S_, double *Z, const double *DS, const double *DZ, double *P, const double *DP, bool hasint)
    {
        FOR_LANE(e, RS) { S_[e] += a * DS[e]; Z[e] += a * DZ[e]; }
        FOR_LANE(e, PN + NX) P[e] += a * DP[e];
        warp_sync();
        {   // trust region (warp-cooperative)
            double ts = 0, tz = 0, dummy = 0;
            FOR_LANE(i, D) if (i > 0) { ts += S_[TRO + i] * S_[TRO + i]; tz += Z[TRO + i] * Z[TRO + i]; }
            warp_sum3(ts, tz, dummy);
            if (lane_id() == 0) { S_[TRO] = nudge(S_[TRO], sqrt(ts)); Z[TRO] = nudge(Z[TRO], sqrt(tz)); }
        }
        FOR_LANE(tk, NTASK) {
            if (tk == NCONE) continue;
            int type, o, d, ci;
            task(tk, type, o, d, ci);
            if (type == 2) {
                if (hasint) { S_[o] = nudge(S_[o], 0.); S_[o + NX] = nudge(S_[o + NX], 0.); Z[o] = nudge(Z[o], 0.); Z[o + NX] = nudge(Z[o + NX], 0.); }
            } else {
                double ts = 0, tz = 0;
#pragma unroll
                for (int i = 1; i < soc::SOC_MAXD; i++) if (i < d) { ts += S_[o + i] * S_[o + i]; tz += Z[o + i] * Z[o + i]; }
                S_[o] = nudge(S_[o], sqrt(ts)); Z[o] = nudge(Z[o], sqrt(tz));
            }
        }
        warp_sync();
    }
    SCPP_HD void apply_step(double a)
    {
        double *S_ = row(0), *Z = row(1), *DS = row(2), *DZ = row(3), *P = pw(0), *DP = pw(1);
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const bool hasint = k < K - 1;
            ld(S_, s + k * RS, RS); ld(Z, z + k * RS, RS); ld(DS, ds + k * RS, RS); ld(DZ, dz + k * RS, RS);
            ld(P, prim + k * PS, PS); ld(DP, dprim + k * PS, PS);
            ld_wait();
            update_window(a, S_, Z, DS, DZ, P, DP, hasint);
            st(s + k * RS, S_, RS); st(z + k * RS, Z, RS); st(prim + k * PS, P, PN + NX);
            warp_sync();
        }
        if (lane_id() == 0) {
            const int r0 = K * RS, p0 = K * PS;
            for (int i = 0; i < 4; i++) { s[r0 + i] += a * ds[r0 + i]; z[r0 + i] += a * dz[r0 + i]; }
            prim[p0] += a * dprim[p0]; prim[p0 + 1] += a * dprim[p0 + 1];
            s[r0] = nudge(s[r0], 0.); z[r0] = nudge(z[r0], 0.);
            s[r0 + 1] = nudge(s[r0 + 1], sqrt(s[r0 + 2] * s[r0 + 2] + s[r0 + 3] * s[r0 + 3]));
            z[r0 + 1] = nudge(z[r0 + 1], sqrt(z[r0 + 2] * z[r0 + 2] + z[r0 + 3] * z[r0 + 3]));
        }
        warp_sync();
    }

    // =============================================================================================================
    //  driver
    // =============================================================================================================
    SCPP_HD IpmResult solve(const IpmSettings &st_, bool have_prev = false)
    {
        IpmResult res;
        res.status = 1; res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.;
        const int np = n_prim(K), m = m_rows(K);
        tables_init();
        const bool warm = have_prev && st_.warm > 0. && st_.warm < 1.;
        Norms nm;
        double tm;
        if (warm) {
            // previous interior point of this instance, pulled back from the boundary; pinned variables keep their values
            const double lw = st_.warm, lc = 1. - st_.warm;
            FOR_LANE(e, K * PS) { const int k = e / PS, i = e - k * PS; if (i < NB && fixed(k, i)) prim[e] = fixv[k * NB + i]; }
            FOR_LANE(e, m) { s[e] *= lw; z[e] *= lw; }
            warp_sync();
            cone_shift(s, lc); cone_shift(z, lc);
            warp_sync();
        } else {
        // ---- starting point (CVXOPT conelp / ECOS style): least-squares primal and dual points, W = I
        FOR_LANE(e, K * PS) {
            const int k = e / PS, i = e - k * PS;
            double v = 0.;
            if (i < NB) v = fixed(k, i) ? fixv[k * NB + i] : (i < NX ? Xbar[k * NX + i] : Ubar[k * NU + (i - NX)]);
            prim[e] = v;
        }
        if (lane_id() == 0) { prim[K * PS] = sigbar; prim[K * PS + 1] = 0.; }
        FOR_LANE(e, m) { s[e] = 0.; z[e] = 0.; }
        warp_sync();
        cone_shift(s, 1.); cone_shift(z, 1.);
        warp_sync();
        phase_residuals(nm, true);
        if (!phase_factor()) { res.status = 2; return res; }
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
            if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift(s, 1. - mg); }
            warp_sync();
        }
        // dual: min |z| s.t. G'z + c = 0  ->  rx = -c, rz = 0
        FOR_LANE(e, m) ds[e] = 0.;
        FOR_LANE(e, K * PS) { const int i = e % PS; dprim[e] = (i == NB) ? -w_tr : ((i >= PN && i < PN + NX && e / PS < K - 1) ? -w_vc : 0.); }
        if (lane_id() == 0) { dprim[K * PS] = -w_time; dprim[K * PS + 1] = -w_trs; }
        warp_sync();
        phase_solve(0, 0., 0., 0., tm);
        FOR_LANE(e, m) z[e] = dz[e];
        warp_sync();
        {
            double mg, n2; cone_margin(z, mg, n2);
            if (mg <= 1e-8 * fmax(1., sqrt(n2))) { cone_shift(z, 1. - mg); }
            warp_sync();
        }
        }
        const double cnorm = sqrt(w_time * w_time + w_trs * w_trs + K * w_tr * w_tr + (K - 1) * NX * w_vc * w_vc);
        const double resx0 = fmax(1., cnorm);
        const int degree = K * (NLP + NCN) + (K - 1) * 2 * NX + 2;
        double best = 1e300;
        int it;
#pragma unroll 1
        double pending = 0.;       // step of the previous iteration, applied inside the next residual sweep
        for (it = 0; it <= st_.maxit; it++) {
            phase_residuals(nm, false, pending);
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
                if (score <= 1e4) { FOR_LANE(e, np) best_[e] = prim[e]; }      // a fallback iterate only matters inside the accuracy band
                warp_sync();
            }
            if (!nm.bad && pres <= st_.feastol && dres <= st_.feastol && (gap <= st_.abstol || relgap <= st_.reltol)) { res.status = 0; break; }
            if (nm.bad || it == st_.maxit || (score > 1e3 * best && best < 1e4)) { res.status = nm.bad ? 2 : (it == st_.maxit ? 1 : 2); break; }
            if (!phase_factor()) { res.status = 2; break; }
            double tmax;
            phase_solve(1, 1., 0., -1., tmax);                               // affine direction
            const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
            const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = gap / degree;
            phase_solve(2, 1. - sig, sig * mu, -(1. - sig), tmax);           // combined direction
            pending = tmax <= 0.99 ? 1. : 0.99 / tmax;
        }
        if (res.status != 0) {
            if (best <= 1e4) { FOR_LANE(e, np) prim[e] = best_[e]; }
            warp_sync();
            if (best <= 10.) res.status = 0; else if (best <= 1e4) res.status = 3;
        } else res.iterations = it;
        return res;
    }
};

} // namespace scpp
