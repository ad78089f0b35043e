// scpp_b200/csrc/ipm_cta.inl — the same interior-point method with ONE CTA PER INSTANCE (member functions of Ipm<M>; included inside
// the struct in ipm.cuh).  Round 2 redesign of kernel K2 (k_solve_cta in kernels.cuh):
//
//   * a CTA (8 warps) takes an instance from a device-side queue and runs its WHOLE sub-problem (all interior-point iterations) before
//     it takes the next one, so the instance's row arrays (~0.3 MB) and tiles (0.14 MB) are touched by one SM only and stay in L2 for the
//     whole solve (one warp per instance with 1036 instances in flight streamed 0.84 GB through HBM in every pass);
//   * the block factor lives in SHARED MEMORY for the whole solve: per stage the packed lower triangle of Linv_kk and l_k (189 doubles,
//     76 KB at K = 50).  L_{k+1,k} = O_k Linv_kk' is never stored: the substitutions apply it as  O_k (Linv' f)  with O_k = [-D A~ ;
//     C' D A~] taken from the tile, which costs the same flops as the dense product (two triangular + one tile mat-vec instead of two
//     dense ones) and cuts the factor from 666 to 189 doubles per stage;
//   * the right-hand side / solution of the Newton system (gv) lives in shared memory too;
//   * STAGE-PARALLEL passes run on all 256 threads, one work item per (stage, row group) or (stage, variable): the cone arithmetic of a
//     stage is split into its trust-region cone, its model rows and its 14 virtual-control pairs, the G'z-type sums are GATHERS per
//     variable (deterministic, no atomics); scalars are reduced over the CTA in a fixed order;
//   * the assembly of the 18 x 18 diagonal blocks is fully parallel over (stage, row); only the Cholesky chain itself is sequential
//     (warp 0), reading the tiles through 8-byte asynchronous copies one stage ahead.
//
// All arithmetic is the method of this file's host (ipm.cuh header): same Newton systems, same local eliminations, same termination.

static constexpr int NPK = NB * (NB + 1) / 2;                 // packed lower triangle of an NB x NB block
static constexpr int FR = pad2(NPK + NB);                     // factor record of a stage in shared memory: Linv_kk (packed) | l_k
static constexpr int TW = NB + NU;                            // row stride of a staged tile: A~ | C
static constexpr int TBS = pad2(NX * TW + NX);                // staged tile + D of the interval
static constexpr int NRING = 4;                                // records in flight in the sequential sweeps of a solve
static constexpr int C_RED = W_U, C_BC = C_RED + 16 * 16, C_REVT = C_BC + 32, C_HS = C_REVT + pad2(NB * 4), C_LN = C_HS + BLK, C_MM = C_LN + BLK,
                     C_TB = C_MM + pad2(NX * NB), C_GV = C_TB + 2 * TBS;
static constexpr int C_RING = C_HS;                           // ring of B_k records (substitution sweeps): aliases the scratch of the factor chain
static_assert(C_GV - C_RING >= NRING * BLK, "the record ring must fit into the factor-chain scratch");
static constexpr int REV_MAX = 8;                             // model-row entries per variable (reverse table)
SCPP_HD static int cta_sm_doubles(int K) { return C_GV + 2 * NB * ks(K) + K * FR; }
SCPP_HD static int pidx(int i, int j) { return i * (i + 1) / 2 + j; }          // j <= i

double *c_fac, *gv2;                                                          // [K][FR], [NB][KS] shared
SCPP_HD double *slot(int k) const { return c_fac + (size_t)k * FR; }
SCPP_HD double *tb(int b) const { return sm + C_TB + (b & 1) * TBS; }

// binding for the CTA solver: the row arrays stay in the global workspace (L2-resident), gv and the factor move to shared memory
SCPP_HD void bind_cta(double *ws, double *smem)
{
    bind(ws, smem);
    gv = smem + C_GV;
    gv2 = gv + NB * KS;
    c_fac = gv2 + NB * KS;
}

// ---- CTA-wide reductions (fixed order: results are bit-identical whatever else runs on the GPU) --------------------------------------
template <int N>
SCPP_HD void cta_sum(double (&v)[N])
{
#if defined(__CUDA_ARCH__)
    double *red = sm + C_RED;
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = warp_sum(v[i]);
    if (lane_id() == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) red[cta_warp() * 16 + i] = v[i];
    }
    cta_sync();
#pragma unroll
    for (int i = 0; i < N; i++) { double a = 0.; for (int w = 0; w < cta_warps(); w++) a += red[w * 16 + i]; v[i] = a; }
    cta_sync();
#endif
}
SCPP_HD double cta_max(double v)
{
#if defined(__CUDA_ARCH__)
    double *red = sm + C_RED;
    v = warp_max(v);
    if (lane_id() == 0) red[cta_warp() * 16] = v;
    cta_sync();
    double a = red[0];
    for (int w = 1; w < cta_warps(); w++) a = fmax(a, red[w * 16]);
    cta_sync();
    return a;
#else
    return v;
#endif
}
SCPP_HD double cta_bcast(double v)      // value of thread 0
{
#if defined(__CUDA_ARCH__)
    double *bc = sm + C_BC;
    if (cta_tid() == 0) bc[0] = v;
    cta_sync();
    v = bc[0];
    cta_sync();
#endif
    return v;
}

// Reverse table of the model rows (shared memory, built once per sub-problem by cp_tables): for variable j the (row, coefficient code) pairs
// of the rows that touch it; code >= 0: cst[code], code < 0: -tdir[-code-1].  REVT[j][t] = row | (code + 64) << 8 ; count in slot REV_MAX-1... kept
// as ints: entry t of variable j at revt()[j * REV_MAX + t], -1 terminates.
SCPP_HD int *revt() const { return reinterpret_cast<int *>(sm + C_REVT); }
SCPP_HD void cp_tables()
{
    if (cta_tid() < LANES) tables_init();
    FOR_CTA(j, NB) {
        int n = 0, *e = revt() + j * REV_MAX;
        for (int r = 0; r < NROW; r++) {
            const RowDesc rd = M::row(r);
            for (int q = 0; q < rd.n; q++) if (rd.idx[q] == j && n < REV_MAX - 1) e[n++] = r | ((rd.cs[q] + 64) << 8);
        }
        e[n] = -1;
    }
    cta_sync();
}
// model rows that touch variable j of a node: sum over them of coef * v[row]  (v: a stage-minor row array)
SCPP_HD double gather_model(int j, int k, const double *td, const double *v) const
{
    double a = 0.;
    const int *e = revt() + j * REV_MAX;
#pragma unroll 1
    for (int t = 0; e[t] >= 0; t++) {
        const int r = e[t] & 255, code = (e[t] >> 8) - 64;
        const double cf = code >= 0 ? cstw()[code] : -td[-code - 1];
        a += cf * v[r * KS + k];
    }
    return a;
}

// =====================================================================================================================================
//  pass U (CTA): (prim, s, z) += a (dprim, ds, dz), cones nudged back inside when rounding ate their margin (see pass_update)
// =====================================================================================================================================
SCPP_HD void cp_update(double a)
{
    FOR_CTA(e, PSN * KS) prim[e] += a * dprim[e];
    const int items = K * 2 + 2 * (K - 1);
    FOR_CTA(it, items) {
        if (it < K) {                      // model rows of stage k
            const int k = it;
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
        } else if (it < 2 * K) {           // trust-region cone
            const int k = it - K;
            double S[D], Z[D];
            upd_rows<TRO, D>(a, k, S, Z);
            double ts = 0, tz = 0;
#pragma unroll
            for (int i = 1; i < D; i++) { ts += S[i] * S[i]; tz += Z[i] * Z[i]; }
            S[0] = nudge(S[0], sqrt(ts)); Z[0] = nudge(Z[0], sqrt(tz));
            put_rows<TRO, D>(k, S, Z);
        } else {                           // virtual-control pairs of interval k: the s- rows or the s+ rows
            const int e = it - 2 * K, h = e / (K - 1), k = e - h * (K - 1);
            double S[NX], Z[NX];
            if (h == 0) upd_rows<MN, NX>(a, k, S, Z); else upd_rows<MN + NX, NX>(a, k, S, Z);
#pragma unroll
            for (int r = 0; r < NX; r++) { S[r] = nudge(S[r], 0.); Z[r] = nudge(Z[r], 0.); }
            if (h == 0) put_rows<MN, NX>(k, S, Z); else put_rows<MN + NX, NX>(k, S, Z);
        }
    }
    if (cta_tid() == 0 && !sig_fixed) {
        const int r0 = RS * KS, p0 = PSN * KS;
        double s4[4], z4[4];
        for (int i = 0; i < 4; i++) { s4[i] = s[r0 + i] + a * ds[r0 + i]; z4[i] = z[r0 + i] + a * dz[r0 + i]; }
        s4[0] = nudge(s4[0], 0.); z4[0] = nudge(z4[0], 0.);
        s4[1] = nudge(s4[1], sqrt(s4[2] * s4[2] + s4[3] * s4[3]));
        z4[1] = nudge(z4[1], sqrt(z4[2] * z4[2] + z4[3] * z4[3]));
        for (int i = 0; i < 4; i++) { s[r0 + i] = s4[i]; z[r0 + i] = z4[i]; }
        prim[p0] += a * dprim[p0]; prim[p0 + 1] += a * dprim[p0 + 1];
    }
    cta_sync();
}

// =====================================================================================================================================
//  pass R (CTA): residuals, Nesterov-Todd scaling, termination quantities.  Items: trust-region cone (k), model rows (k), pair row (k, i),
//  dual residual of variable (k, j) [a gather over the rows that touch it]
// =====================================================================================================================================
SCPP_HD void cp_residuals(Norms &nm, bool identity)
{
    double gap = 0, rz2 = 0, pcost = 0, zrz = 0, h2 = 0, rx2 = 0, xrx = 0, acc_sig = 0, bad = 0;
    const double sg = prim[PSN * KS];
    const int n_tr = K, n_md = K, n_pr = K * NX, n_vr = K * NB;
    FOR_CTA(it, n_tr + n_md + n_pr + n_vr) {
        if (it < n_tr) {
            // ---- trust-region cone  (delta ; xibar - xi) in Q^{1+NB}
            const int k = it;
            double P[NB], S[D], Z[D], XB[NB], RZ[D];
#pragma unroll
            for (int j = 0; j < NB; j++) P[j] = prim[j * KS + k];
            const double delta = prim[NB * KS + k];
#pragma unroll
            for (int i = 0; i < D; i++) { S[i] = s[(TRO + i) * KS + k]; Z[i] = z[(TRO + i) * KS + k]; }
#pragma unroll
            for (int j = 0; j < NB; j++) XB[j] = xibar(k, j);
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
            }
            if (scvx) h2 += delta * delta;
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
            const double rxdl = scvx ? 0. : w_tr - z0;
            rx[NB * KS + k] = rxdl;
            rx2 += rxdl * rxdl; xrx += delta * rxdl;
        } else if (it < n_tr + n_md) {
            // ---- model rows (LP rows, model cones)
            const int k = it - n_tr;
            double P[NB], td[3], S2[NROW], Z2[NROW], RZ2[NROW], CEv[NCONE];
#pragma unroll
            for (int j = 0; j < NB; j++) P[j] = prim[j * KS + k];
#pragma unroll
            for (int q = 0; q < 3; q++) td[q] = tdir[3 * k + q];
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
                }
                double e2i = 1.;
                if (identity || !soc::scale(sk, zk, d, w, e2i, lm)) {
                    if (!identity) bad = 1;
                    e2i = 1.;
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
            }
#pragma unroll
            for (int r = 0; r < NROW; r++) { rz[r * KS + k] = RZ2[r]; wb[r * KS + k] = S2[r]; lam[r * KS + k] = Z2[r]; }
#pragma unroll
            for (int c = 0; c < NCONE; c++) ce[c * KS + k] = CEv[c];
        } else if (it < n_tr + n_md + n_pr) {
            // ---- pair row i of interval k:  t_i >= |r_i| ,  r = x_{k+1} - A~ xi_k - C u_{k+1} - s sigma - z
            const int e = it - n_tr - n_md, i = e / K, k = e - i * K;
            const int o = MN + i;
            if (k < K - 1) {
                double acc = prim[i * KS + k + 1], acc2 = 0;
#pragma unroll
                for (int j = 0; j < NB; j += 2) { acc -= T(i, j, k) * prim[j * KS + k]; acc2 -= T(i, j + 1, k) * prim[(j + 1) * KS + k]; }
#pragma unroll
                for (int a = 0; a < NU; a++) acc2 -= T(i, NB + a, k) * prim[(NX + a) * KS + k + 1];
                const double tsg = T(i, NB + NU, k), zc = T(i, NB + NU + 1, k);
                const double r = acc + acc2 - tsg * sg - zc;
                const double sm_ = s[o * KS + k], sp = s[(o + NX) * KS + k], zm = z[o * KS + k], zp = z[(o + NX) * KS + k];
                acc_sig -= tsg * (zm - zp);
                const double t = prim[(PN + i) * KS + k];
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
            } else {
                rz[o * KS + k] = 0.; rz[(o + NX) * KS + k] = 0.; wb[o * KS + k] = 1.; wb[(o + NX) * KS + k] = 1.;
                lam[o * KS + k] = 1.; lam[(o + NX) * KS + k] = 1.; rx[(PN + i) * KS + k] = 0.;
            }
        } else {
            // ---- dual residual of variable j of node k:  (G'z)_j  gathered over the rows that touch it
            const int e = it - n_tr - n_md - n_pr, j = e / K, k = e - j * K;
            double td[3];
#pragma unroll
            for (int q = 0; q < 3; q++) td[q] = tdir[3 * k + q];
            // every load of the item is issued before the first use (memory-level parallelism: the sums below would otherwise wait one L2
            // round trip per term)
            double tv[NX], zm[NX], zp[NX], tc[NX], zmp[NX], zpp[NX];
            const bool hasint = k < K - 1, inprev = k > 0 && j >= NX;
#pragma unroll
            for (int i = 0; i < NX; i++) {
                tv[i] = hasint ? T(i, j, k) : 0.; zm[i] = hasint ? z[(MN + i) * KS + k] : 0.; zp[i] = hasint ? z[(MN + NX + i) * KS + k] : 0.;
                tc[i] = inprev ? T(i, NB + (j - NX), k - 1) : 0.; zmp[i] = inprev ? z[(MN + i) * KS + k - 1] : 0.; zpp[i] = inprev ? z[(MN + NX + i) * KS + k - 1] : 0.;
            }
            double v = z[(TRO + 1 + j) * KS + k] + gather_model(j, k, td, z);
            if (k > 0 && j < NX) v += z[(MN + j) * KS + k - 1] - z[(MN + NX + j) * KS + k - 1];
            double v2 = 0, v3 = 0;
#pragma unroll
            for (int i = 0; i < NX; i++) { v2 += tv[i] * (zm[i] - zp[i]); v3 += tc[i] * (zmp[i] - zpp[i]); }
            v -= v2 + v3;
            if (fixed(k, j)) v = 0.;
            rx[j * KS + k] = v;
            rx2 += v * v; xrx += prim[j * KS + k] * v;
        }
    }
    double r[9] = {gap, rz2, pcost, zrz, rx2, xrx, h2, acc_sig, bad};
    cta_sum(r);
    nm.gap = r[0]; nm.rz2 = r[1]; nm.pcost = r[2]; nm.zrz = r[3]; nm.rx2 = r[4]; nm.xrx = r[5]; nm.h2 = r[6]; nm.acc_sig = r[7]; nm.bad = r[8] != 0.;
    if (!sig_fixed) {
        // the four global rows: every thread computes the same scalars, one thread stores
        Norms g = nm;
        if (cta_tid() < LANES) residual_globals(g, identity);          // warp 0 (its lane 0 stores)
        double q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cta_tid() == 0) { q[0] = g.gap; q[1] = g.rz2; q[2] = g.zrz; q[3] = g.pcost; q[4] = g.h2; q[5] = g.rx2; q[6] = g.xrx; q[7] = g.bad; }
        double *bc = sm + C_BC;
#if defined(__CUDA_ARCH__)
        if (cta_tid() == 0) for (int i = 0; i < 8; i++) bc[i] = q[i];
        cta_sync();
        for (int i = 0; i < 8; i++) q[i] = bc[i];
        cta_sync();
#else
        (void)bc;
#endif
        nm.gap = q[0]; nm.rz2 = q[1]; nm.zrz = q[2]; nm.pcost = q[3]; nm.h2 = q[4]; nm.rx2 = q[5]; nm.xrx = q[6]; nm.bad = q[7] != 0.;
    }
    cta_sync();
}

// =====================================================================================================================================
//  assembly (CTA): D of every interval, the packed diagonal blocks H_kk and the borders b_k into the shared factor records; returns the
//  corner contribution  sum_k s_k' D_k s_k  (reduced over the CTA)
// =====================================================================================================================================
SCPP_HD double cp_assemble()
{
    double corner = 0.;
    FOR_CTA(e, NX * (K - 1)) {
        const int i = e / (K - 1), k = e - i * (K - 1);
        const double dm = capd(wb[(MN + i) * KS + k]), dp = capd(wb[(MN + NX + i) * KS + k]);
        const double dt = 4. * dm * dp / (dm + dp);
        dtv[i * KS + k] = dt;
        const double sv = T(i, NB + NU, k);
        corner += dt * sv * sv;
    }
    cta_sync();
    // ---- dense part: interval k, carry of interval k-1, trust-region cone; row i of the block per item
    FOR_CTA(it, K * NB) {
        const int i = it / K, k = it - i * K;
        double *F = slot(k);
        const uint32_t mk = fixm[k];
        const bool pin_i = (mk >> i) & 1u, hasint = k < K - 1, hasprev = k > 0;
        const double e2i = ce[NCONE * KS + k], w0 = wb[TRO * KS + k];
        const double kap = e2i * (2. * w0 * w0 - 1.), c2 = scvx ? 0. : 4. * e2i * e2i * w0 * w0 / kap, beta = 2. * e2i - c2;
        const double wi = wb[(TRO + 1 + i) * KS + k];
        double ai[NX], ci[NX];            // D A~(:,i) of interval k ; D C(:,i-NX) of interval k-1 (input rows)
#pragma unroll
        for (int r = 0; r < NX; r++) {
            ai[r] = hasint ? dtv[r * KS + k] * T(r, i, k) : 0.;
            ci[r] = (hasprev && i >= NX) ? dtv[r * KS + k - 1] * T(r, NB + (i - NX), k - 1) : 0.;
        }
        // software pipeline over j: the column of j + 1 is loaded while the column of j is consumed
        double tj[NX], tn[NX], wj = wb[(TRO + 1) * KS + k], wn = 0.;
#pragma unroll
        for (int r = 0; r < NX; r++) tj[r] = hasint ? T(r, 0, k) : 0.;
#pragma unroll 1
        for (int j = 0; j <= i; j++) {
            if (j < i) {
#pragma unroll
                for (int r = 0; r < NX; r++) tn[r] = hasint ? T(r, j + 1, k) : 0.;
                wn = wb[(TRO + 2 + j) * KS + k];
            }
            double v;
            if (pin_i || ((mk >> j) & 1u)) v = (i == j) ? 1. : 0.;
            else {
                v = beta * wi * wj;
                if (i == j && tr_row(i + 1)) v += e2i;
                if (hasint) {
                    double a0 = 0, a1 = 0;
#pragma unroll
                    for (int r = 0; r < NX; r += 2) { a0 += ai[r] * tj[r]; a1 += ai[r + 1] * tj[r + 1]; }
                    v += a0 + a1;
                }
                if (hasprev) {
                    if (i < NX) { if (i == j) v += dtv[i * KS + k - 1]; }
                    else if (j < NX) v -= ci[j];
                    else {
                        double a0 = 0;
#pragma unroll
                        for (int r = 0; r < NX; r++) a0 += ci[r] * T(r, NB + (j - NX), k - 1);
                        v += a0;
                    }
                }
            }
            F[pidx(i, j)] = v;
#pragma unroll
            for (int r = 0; r < NX; r++) tj[r] = tn[r];
            wj = wn;
        }
        double b = 0.;
        if (!pin_i) {
            if (hasint) {
#pragma unroll
                for (int r = 0; r < NX; r++) b += ai[r] * T(r, NB + NU, k);
            }
            if (hasprev) {
                if (i < NX) b -= dtv[i * KS + k - 1] * T(i, NB + NU, k - 1);
                else {
#pragma unroll
                    for (int r = 0; r < NX; r++) b += ci[r] * T(r, NB + NU, k - 1);
                }
            }
        }
        F[NPK + i] = b;
    }
    cta_sync();
    // ---- model rows: diagonal terms and sparse rank-one terms (each touches at most 4 variables); one thread per stage adds them in a
    //      fixed order
    FOR_CTA(k, K) {
        double *F = slot(k);
        const uint32_t mk = fixm[k];
        double td[3];
        for (int q = 0; q < 3; q++) td[q] = tdir[3 * k + q];
#pragma unroll 1
        for (int c = 0; c < NRK; c++) {
            const int *sp = sup() + 4 * c;
            double a4[4] = {0., 0., 0., 0.}, d4[4] = {0., 0., 0., 0.}, alpha;
            if (c < NCONE) {
                const int o = NLP + M::cone_off(c), dim = M::cone_dim(c);
                const double e2c = ce[c * KS + k];
                for (int r = 0; r < dim; r++) {
                    const double wh = (r == 0) ? wb[o * KS + k] : -wb[(o + r) * KS + k];
                    const RowDesc rd = M::row(o + r);
                    for (int q = 0; q < rd.n; q++) {
                        const double cf = coef_reg(rd, q, td);
                        for (int t = 0; t < 4; t++) if (sp[t] == rd.idx[q]) { a4[t] += wh * cf; d4[t] += (r == 0 ? -e2c : e2c) * cf * cf; }
                    }
                }
                alpha = 2. * e2c;
            } else {
                alpha = 0.;
                for (int r = 0; r < NLP; r++) {
                    const double dv = wb[r * KS + k];
                    const RowDesc rd = M::row(r);
                    if (rd.n == 1) {                     // single-entry LP row: diagonal term (its variable need not be in the support list)
                        const int i = rd.idx[0];
                        const double cf = coef_reg(rd, 0, td);
                        if (!((mk >> i) & 1u)) F[pidx(i, i)] += dv * cf * cf;
                    } else {
                        for (int q = 0; q < rd.n; q++) for (int t = 0; t < 4; t++) if (sp[t] == rd.idx[q]) a4[t] = coef_reg(rd, q, td);
                        alpha = dv;
                    }
                }
            }
            for (int t = 0; t < 4; t++) {
                const int i = sp[t];
                if (i < 0 || ((mk >> i) & 1u)) continue;
                for (int u = 0; u < 4; u++) {
                    const int j = sp[u];
                    if (j < 0 || j > i || ((mk >> j) & 1u)) continue;
                    F[pidx(i, j)] += alpha * a4[t] * a4[u] + (t == u ? d4[t] : 0.);
                }
            }
        }
    }
    double r[1] = {corner};
    cta_sum(r);          // also the barrier that publishes the records to the chain warp
    return r[0];
}

// ---- chain helpers (warp 0) ---------------------------------------------------------------------------------------------------------
// ---- bulk asynchronous copies (cp.async.bulk + mbarrier, SASS UBLKCP): one elected lane moves a whole contiguous record -------------------
SCPP_HD unsigned long long *mbar(int i) const { return reinterpret_cast<unsigned long long *>(sm + C_BC + 24) + i; }
SCPP_HD void mbar_init_all() const
{
#if defined(__CUDA_ARCH__)
    if (lane_id() == 0) {
        for (int i = 0; i < NRING; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar(i))) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif
}
SCPP_HD void bulk_load(double *dst, const double *src, int n, int bar) const      // n doubles, 16-byte aligned, n even
{
#if defined(__CUDA_ARCH__)
    if (lane_id() == 0) {
        const unsigned b = (unsigned)__cvta_generic_to_shared(mbar(bar)), d = (unsigned)__cvta_generic_to_shared(dst), bytes = 8u * n;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic reads of the buffer are ordered before the async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
    }
#else
    (void)bar;
    memcpy(dst, src, sizeof(double) * n);
#endif
}
SCPP_HD void bulk_wait(int bar, unsigned parity) const
{
#if defined(__CUDA_ARCH__)
    const unsigned b = (unsigned)__cvta_generic_to_shared(mbar(bar));
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    } while (!ok);
#else
    (void)bar; (void)parity;
#endif
}
// Cholesky of the packed block in F (lower triangle) and its inverse, written back packed; HS: NB x NB scratch
SCPP_HD bool chol_inv_packed(double *F, double *HS)
{
#if defined(__CUDA_ARCH__)
    const int lane = lane_id();
    const bool own = lane < NB;
    double h[NB], invd[NB];
#pragma unroll
    for (int c = 0; c < NB; c++) h[c] = (own && c <= lane) ? F[pidx(lane, c)] : 0.;
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
        for (int c = 0; c < NB; c++) HS[lane * NB + c] = h[c];
    }
    __syncwarp();
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        double a0 = (i == lane) ? 1. : 0., a1 = 0.;
#pragma unroll
        for (int q = 0; q < i; q += 2) { a0 = fma(-HS[i * NB + q], x[q], a0); if (q + 1 < i) a1 = fma(-HS[i * NB + q + 1], x[q + 1], a1); }
        x[i] = (a0 + a1) * invd[i];
    }
    if (own) {
#pragma unroll
        for (int i = 0; i < NB; i++) if (i >= lane) F[pidx(i, lane)] = x[i];
    }
    __syncwarp();
    return ok;
#else
    bool ok = true;
    for (int i = 0; i < NB; i++) for (int j = 0; j < NB; j++) HS[i * NB + j] = (j <= i) ? F[pidx(i, j)] : 0.;
    for (int j = 0; j < NB; j++) {
        for (int i = j; i < NB; i++) { double v = HS[i * NB + j]; for (int c = 0; c < j; c++) v -= HS[i * NB + c] * HS[j * NB + c]; HS[i * NB + j] = v; }
        const double djj = HS[j * NB + j];
        if (!(djj > 0.)) ok = false;
        const double inv = 1. / sqrt(djj > 0. ? djj : 1.);
        for (int i = j; i < NB; i++) HS[i * NB + j] *= inv;
    }
    for (int c = 0; c < NB; c++)
        for (int i = c; i < NB; i++) {
            double v = (i == c) ? 1. : 0.;
            for (int q = c; q < i; q++) v -= HS[i * NB + q] * F[pidx(q, c)];
            F[pidx(i, c)] = v / HS[i * NB + i];
        }
    return ok;
#endif
}

// The block Cholesky chain over the assembled records, in place  H_kk | b_k  ->  Linv_kk | l_k ; also writes B_k = L_{k+1,k} Linv_kk (global
// record k) and sets l_ss.  The chain is sequential in k, but only the 18 x 18 Cholesky + triangular inverse of a stage is a one-warp job
// (warp 0, registers and shuffles); while it runs the other warps stage the tile of the interval into shared memory, and the block
// products that follow (M = A~ Linv', C' D M, the Schur update L L' of the next block, B_k) are split by 8 x 8 output tiles over all warps
// on the FP64 tensor cores, four block barriers per stage.
SCPP_HD bool cp_chain_factor(double corner_in)
{
    double *HS = sm + C_HS, *LN = sm + C_LN, *MM = sm + C_MM, *lv = vec(0);
    const int w = cta_warp(), nw = cta_warps(), wl = nw - 1;       // wl: the warp that also carries the border vector
    constexpr int MT_X = (NX + 7) / 8, NT_B = (NB + 7) / 8, KT_B = (NB + 3) / 4, KT_X = (NX + 3) / 4;
    double corner = 0., bad = 0.;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const bool hasint = k < K - 1;
        double *F = slot(k);
        double *t = tb(k), *Dt = t + NX * TW;
        if (w == 0) { if (!chol_inv_packed(F, HS)) bad = 1.; }
        if (w != 0 || nw == 1) {
            // meanwhile the other warps: B_{k-1} = L_{k,k-1} Linv_{k-1,k-1} (off the critical path: LN still holds L_{k,k-1}) ...
            if (k > 0) {
                const double *Fp = slot(k - 1);
                double *Bp = fac + (size_t)(k - 1) * BLK;
                const int w1 = nw == 1 ? 0 : w - 1, nw1 = nw == 1 ? 1 : nw - 1;
                for (int job = w1; job < NT_B * NT_B; job += nw1) {
                    const int mi = job / NT_B, ni = job - mi * NT_B;
                    blk::mm_tile<KT_B>(mi, ni, [&](int m, int kk) { return (m < NB && kk < NB) ? LN[m * NB + kk] : 0.; },
                                       [&](int kk, int n) { return (kk < NB && n <= kk) ? Fp[pidx(kk, n)] : 0.; },
                                       [&](int m, int n, double v) { if (m < NB && n < NB) Bp[m * NB + n] = v; });
                }
            }
            // ... and the tile A~ | C and D of interval k
            if (hasint) {
                const int nth = nw == 1 ? cta_threads() : cta_threads() - LANES, me = nw == 1 ? cta_tid() : cta_tid() - LANES;
                for (int e = me; e < NX * TW + NX; e += nth) {
                    if (e < NX * TW) { const int i = e / TW, c = e - i * TW; t[e] = T(i, c, k); }
                    else t[e] = dtv[(e - NX * TW) * KS + k];
                }
            }
        }
        cta_sync();                                                                 // [A] Linv_kk and the tile are in shared memory
        const uint32_t mk = fixm[k], mn = hasint ? fixm[k + 1] : 0u;
        if (hasint) {
            // M = A~ Linv' with the pinned columns of node k masked; epilogue: D M -> MM, -D M (rows of pinned x_{k+1} masked) -> Ln
            for (int job = w; job < MT_X * NT_B; job += nw) {
                const int mi = job / NT_B, ni = job - mi * NT_B;
                blk::mm_tile<KT_B>(mi, ni, [&](int m, int kk) { return (m < NX && kk < NB && !((mk >> kk) & 1u)) ? t[m * TW + kk] : 0.; },
                                   [&](int kk, int n) { return (n < NB && kk <= n) ? F[pidx(n, kk)] : 0.; },
                                   [&](int m, int n, double v) { if (m < NX && n < NB) { const double dm = Dt[m] * v; MM[m * NB + n] = dm; LN[m * NB + n] = ((mn >> m) & 1u) ? 0. : -dm; } });
            }
        }
        if (w == wl) {                                                              // l_k = Linv b_k
            double v = 0.;
            FOR_LANE(j, NB) { v = 0; for (int c = 0; c <= j; c++) v += F[pidx(j, c)] * F[NPK + c]; lv[j] = v; }
            warp_sync();
            FOR_LANE(j, NB) { const double u = lv[j]; F[NPK + j] = u; corner -= u * u; }
        }
        cta_sync();                                                                 // [B]
        if (hasint) {
            for (int job = w; job < NT_B; job += nw)                                // L_{k+1,k} input rows: C' (D M)
                blk::mm_tile<KT_X>(0, job, [&](int m, int kk) { return (m < NU && kk < NX) ? t[kk * TW + NB + m] : 0.; },
                                   [&](int kk, int n) { return (kk < NX && n < NB) ? MM[kk * NB + n] : 0.; },
                                   [&](int m, int n, double v) { if (m < NU && n < NB) LN[(NX + m) * NB + n] = ((mn >> (NX + m)) & 1u) ? 0. : v; });
        }
        cta_sync();                                                                 // [C] L_{k+1,k} complete
        if (hasint) {
            double *Fn = slot(k + 1);
            constexpr int NLOW = NT_B * (NT_B + 1) / 2;
            for (int job = w; job < NLOW; job += nw) {                              // Schur update of the next block (lower tiles)
                int mi = 0, r = job;
                while (r > mi) { r -= mi + 1; mi++; }
                const int ni = r;
                blk::mm_tile<KT_B>(mi, ni, [&](int m, int kk) { return (m < NB && kk < NB) ? LN[m * NB + kk] : 0.; },
                                   [&](int kk, int n) { return (kk < NB && n < NB) ? LN[n * NB + kk] : 0.; },
                                   [&](int m, int n, double v) { if (m < NB && n <= m) Fn[pidx(m, n)] -= v; });
            }
            if (w == wl) {                                                          // border: b_{k+1} -= L_{k+1,k} l_k
                FOR_LANE(j, NB) {
                    double v = 0;
#pragma unroll 2
                    for (int c = 0; c < NB; c++) v += LN[j * NB + c] * F[NPK + c];
                    Fn[NPK + j] -= v;
                }
            }
        }
        cta_sync();                                                                 // [D] the next block is ready
    }
#if defined(__CUDA_ARCH__)
    asm volatile("fence.proxy.async;" ::: "memory");          // the records were written through the generic proxy and are read by bulk copies
#endif
    double r[2] = {corner, bad};
    cta_sum(r);
    const bool okc = finish_corner(r[0] + corner_in);
    return okc && r[1] == 0.;
}

// ---- substitutions.  With B_k = L_{k+1,k} Linv_kk the block bidiagonal solves split into a SEQUENTIAL recursion that costs one dense
//      NB x NB mat-vec per stage and touches nothing but B_k (warp 0, records streamed by bulk copies NRING stages ahead), and PARALLEL
//      triangular products with the shared-memory factor (whole CTA):
//        forward   t_0 = g_0 ,  t_{k+1} = g_{k+1} - B_k t_k          (sequential, in place in gv)
//                  f_k = Linv_kk t_k ,  ldot = sum_k l_k' f_k          (parallel, gv -> gv2)
//        backward  yh_k = Linv_kk' (f_k - l_k y_sigma)                 (parallel, gv2 -> gv)
//                  y_{K-1} = yh_{K-1} ,  y_k = yh_k - B_k' y_{k+1}     (sequential, in place in gv)
SCPP_HD void cp_seq_forward()
{
    double *ring = sm + C_RING;
    const int nrec = K - 1;
    mbar_init_all();
    for (int r = 0; r < NRING && r < nrec; r++) bulk_load(ring + r * BLK, fac + (size_t)r * BLK, BLK, r);
#pragma unroll 1
    for (int k = 0; k < nrec; k++) {
        const int b = k % NRING;
        bulk_wait(b, (k / NRING) & 1);
        const double *Bk = ring + b * BLK;
        const int j = lane_id();
        if (LANES == 1) {
            double tn[NB];
            for (int a = 0; a < NB; a++) { double v = gv[a * KS + k + 1]; for (int c = 0; c < NB; c++) v -= Bk[a * NB + c] * gv[c * KS + k]; tn[a] = v; }
            for (int a = 0; a < NB; a++) gv[a * KS + k + 1] = tn[a];
        } else if (j < NB) {
            double v0 = gv[j * KS + k + 1], v1 = 0.;
#pragma unroll
            for (int c = 0; c < NB; c += 2) { v0 -= Bk[j * NB + c] * gv[c * KS + k]; v1 -= Bk[j * NB + c + 1] * gv[(c + 1) * KS + k]; }      // NB is even
            gv[j * KS + k + 1] = v0 + v1;
        }
        warp_sync();
        if (k + NRING < nrec) bulk_load(ring + b * BLK, fac + (size_t)(k + NRING) * BLK, BLK, b);
    }
}
SCPP_HD void cp_seq_backward()
{
    double *ring = sm + C_RING;
    const int nrec = K - 1;
    mbar_init_all();
    for (int r = 0; r < NRING && r < nrec; r++) bulk_load(ring + r * BLK, fac + (size_t)(nrec - 1 - r) * BLK, BLK, r);
#pragma unroll 1
    for (int q = 0; q < nrec; q++) {
        const int k = nrec - 1 - q, b = q % NRING;
        bulk_wait(b, (q / NRING) & 1);
        const double *Bk = ring + b * BLK;
        const uint32_t mk = fixm[k];
        const int c = lane_id();
        if (LANES == 1) {
            double yn[NB];
            for (int cc = 0; cc < NB; cc++) { double v = gv[cc * KS + k]; for (int a = 0; a < NB; a++) v -= Bk[a * NB + cc] * gv[a * KS + k + 1]; yn[cc] = ((mk >> cc) & 1u) ? 0. : v; }
            for (int cc = 0; cc < NB; cc++) gv[cc * KS + k] = yn[cc];
        } else if (c < NB) {
            double v0 = gv[c * KS + k], v1 = 0.;
#pragma unroll
            for (int a = 0; a < NB; a += 2) { v0 -= Bk[a * NB + c] * gv[a * KS + k + 1]; v1 -= Bk[(a + 1) * NB + c] * gv[(a + 1) * KS + k + 1]; }
            gv[c * KS + k] = ((mk >> c) & 1u) ? 0. : v0 + v1;
        }
        warp_sync();
        if (q + NRING < nrec) bulk_load(ring + b * BLK, fac + (size_t)(nrec - 1 - (q + NRING)) * BLK, BLK, b);
    }
}
// f = Linv t for every stage (gv -> gv2); returns sum_k l_k' f_k over the CTA
SCPP_HD double cp_par_forward()
{
    double ldot = 0;
    FOR_CTA(it, K * NB) {
        const int j = it / K, k = it - j * K;
        const double *F = slot(k);
        const int rb = pidx(j, 0);
        double v0 = 0, v1 = 0;
        int c = 0;
        for (; c + 1 <= j; c += 2) { v0 += F[rb + c] * gv[c * KS + k]; v1 += F[rb + c + 1] * gv[(c + 1) * KS + k]; }
        if (c <= j) v0 += F[rb + c] * gv[c * KS + k];
        const double f = v0 + v1;
        gv2[j * KS + k] = f;
        ldot += F[NPK + j] * f;
    }
    double r[1] = {ldot};
    cta_sum(r);
    return r[0];
}
// yh = Linv' (f - l y_sigma) for every stage (gv2 -> gv)
SCPP_HD void cp_par_backward(double ysig)
{
    FOR_CTA(it, K * NB) {
        const int c = it / K, k = it - c * K;
        const double *F = slot(k);
        double v0 = 0, v1 = 0;
        int j = c;
        for (; j + 1 < NB; j += 2) { v0 += F[pidx(j, c)] * (gv2[j * KS + k] - F[NPK + j] * ysig); v1 += F[pidx(j + 1, c)] * (gv2[(j + 1) * KS + k] - F[NPK + j + 1] * ysig); }
        if (j < NB) v0 += F[pidx(j, c)] * (gv2[j * KS + k] - F[NPK + j] * ysig);
        gv[c * KS + k] = fixed(k, c) ? 0. : v0 + v1;
    }
    cta_sync();
}

// =====================================================================================================================================
//  pass A of a solve (CTA): right-hand sides.  Row items leave rzv in ds and the per-row contributions to g in dz (scratch until the
//  recovery overwrites it) / wv; after a barrier the variable items gather g into gv (shared).  Returns the sigma right-hand side.
// =====================================================================================================================================
SCPP_HD double cp_rhs(int mode, double csig, double sigmu)
{
    double gsig = 0;
    const int n_tr = K, n_md = K, n_pr = K * NX;
    FOR_CTA(it, n_tr + n_md + n_pr) {
        if (it < n_tr) {
            const int k = it;
            double w[D], q[D], i1[D], i2[D];
            const double rxd = rxv_of(mode, csig, NB, k);
            const double e2i = ce[NCONE * KS + k];
#pragma unroll
            for (int i = 0; i < D; i++) w[i] = wb[(TRO + i) * KS + k];
            ld_rhs_rows<D>(mode, TRO, k, q, i1, i2);
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
            for (int i = 1; i < D; i++) dz[(TRO + i) * KS + k] = e2i * (-2. * dot * w[i] + q[i]) - 2. * e2i * w0 * w[i] * rho;
        } else if (it < n_tr + n_md) {
            const int k = it - n_tr;
            double W2[NROW], a0[NROW], a1[NROW], a2[NROW], CEv[NCONE], GV[NROW];
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
                for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) a0[o + r] = t1[r];
                soc::Mv(wc, e2c, t1, d, t1);
#pragma unroll
                for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) GV[o + r] = t1[r];
            }
#pragma unroll
            for (int r = 0; r < NLP; r++) {
                const double dv = W2[r];
                const double rzv = lp_rzv(mode, csig, sigmu, dv, a0[r], a1[r], a2[r]);
                a0[r] = rzv;
                GV[r] = dv * rzv;
            }
#pragma unroll
            for (int r = 0; r < NROW; r++) dz[r * KS + k] = GV[r];
            if (mode != 0) {
#pragma unroll
                for (int r = 0; r < NROW; r++) ds[r * KS + k] = a0[r];
            }
        } else {
            const int e = it - n_tr - n_md, i = e / K, k = e - i * K;
            const int o = MN + i;
            if (k < K - 1) {
                const double dmt = wb[o * KS + k], dpt = wb[(o + NX) * KS + k];
                double m0, m1, m2, m3, m4, m5;
                if (mode == 0) { m0 = ds[o * KS + k]; m1 = ds[(o + NX) * KS + k]; m2 = m3 = m4 = m5 = 0.; }
                else if (mode == 1) { m0 = rz[o * KS + k]; m1 = rz[(o + NX) * KS + k]; m2 = s[o * KS + k]; m3 = s[(o + NX) * KS + k]; m4 = m5 = 0.; }
                else { m0 = lam[o * KS + k]; m1 = lam[(o + NX) * KS + k]; m2 = cr[o * KS + k]; m3 = cr[(o + NX) * KS + k]; m4 = rz[o * KS + k]; m5 = rz[(o + NX) * KS + k]; }
                const double rm = lp_rzv(mode, csig, sigmu, dmt, m0, m2, m4);
                const double rp = lp_rzv(mode, csig, sigmu, dpt, m1, m3, m5);
                const double dm = capd(dmt), dp = capd(dpt);
                if (mode != 0) { ds[o * KS + k] = rm; ds[(o + NX) * KS + k] = rp; }
                const double rho = (-(dm * rm + dp * rp) + rxv_of(mode, csig, PN + i, k)) / (dm + dp);
                const double wi = dm * (rm + rho) - dp * (rp + rho);
                wv[i * KS + k] = wi;
                gsig -= T(i, NB + NU, k) * wi;
            } else if (mode != 0) { ds[o * KS + k] = 0.; ds[(o + NX) * KS + k] = 0.; }
        }
    }
    cta_sync();
    FOR_CTA(it, K * NB) {
        const int j = it / K, k = it - j * K;
        double td[3];
#pragma unroll
        for (int q = 0; q < 3; q++) td[q] = tdir[3 * k + q];
        double tv[NX], wc[NX], tc[NX], wp[NX];
        const bool hasint = k < K - 1, inprev = k > 0 && j >= NX;
#pragma unroll
        for (int i = 0; i < NX; i++) {
            tv[i] = hasint ? T(i, j, k) : 0.; wc[i] = hasint ? wv[i * KS + k] : 0.;
            tc[i] = inprev ? T(i, NB + (j - NX), k - 1) : 0.; wp[i] = inprev ? wv[i * KS + k - 1] : 0.;
        }
        double g = rxv_of(mode, csig, j, k) + dz[(TRO + 1 + j) * KS + k] + gather_model(j, k, td, dz);
        if (k > 0 && j < NX) g += wv[j * KS + k - 1];
        double v2 = 0, v3 = 0;
#pragma unroll
        for (int i = 0; i < NX; i++) { v2 += tv[i] * wc[i]; v3 += tc[i] * wp[i]; }
        g -= v2 + v3;
        gv[j * KS + k] = fixed(k, j) ? 0. : g;
    }
    double r[1] = {gsig};
    cta_sum(r);
    return r[0];
}

// =====================================================================================================================================
//  pass B of a solve (CTA): recovery of the local variables, dz, ds, the scaled step-length bound and (mode 1) the corrector term
// =====================================================================================================================================
SCPP_HD double cp_recover(int mode, double csig, double rzs, double ysig)
{
    double tmax = 0;
    const int n_tr = K, n_md = K, n_pr = K * NX;
    FOR_CTA(it, n_tr + n_md + n_pr) {
        if (it < n_tr) {
            const int k = it;
            double yk[NB], w[D], q[D], RZ[D], LM[D];
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
            for (int j = 0; j < NB; j++) dprim[j * KS + k] = yk[j];
            const double w0 = w[0], kap = e2i * (2. * w0 * w0 - 1.);
            q[0] = -q[0];
            double pq = -kap * q[0], dot = w0 * q[0];
#pragma unroll
            for (int i = 1; i < D; i++) { q[i] = tr_row(i) ? yk[i - 1] - q[i] : 0.; pq += 2. * e2i * w0 * w[i] * q[i]; dot -= w[i] * q[i]; }
            const double ddl = scvx ? 0. : (rxd - pq) / kap;
            dprim[NB * KS + k] = ddl;
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
        } else if (it < n_tr + n_md) {
            const int k = it - n_tr;
            double yk[NB], td[3];
#pragma unroll
            for (int j = 0; j < NB; j++) yk[j] = gv[j * KS + k];
#pragma unroll
            for (int t = 0; t < 3; t++) td[t] = tdir[3 * k + t];
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
                soc::Mv(wc, e2c, qc, d, qc);
#pragma unroll
                for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) DZ[o + r] = qc[r];
                if (mode != 0) {
#pragma unroll
                    for (int r = 0; r < soc::SOC_MAXD; r++) if (r < d) {
                        gdx[r] = rzs * R2[o + r] - gdx[r];
                        D2[o + r] = gdx[r];
                        lm[r] = L2[o + r];
                    }
                    soc::Wv(wc, e2c, qc, d, qc, false);
                    soc::Wv(wc, e2c, gdx, d, gdx, true);
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
                    const double iw = sqrt(dv), il = 1. / L2[r];
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
        } else {
            const int e = it - n_tr - n_md, i = e / K, k = e - i * K;
            const int o = MN + i;
            if (k < K - 1) {
                double acc = gv[i * KS + k + 1], acc2 = 0;
#pragma unroll
                for (int j = 0; j < NB; j += 2) { acc -= T(i, j, k) * gv[j * KS + k]; acc2 -= T(i, j + 1, k) * gv[(j + 1) * KS + k]; }
#pragma unroll
                for (int a = 0; a < NU; a++) acc2 -= T(i, NB + a, k) * gv[(NX + a) * KS + k + 1];
                const double ady = acc + acc2 - T(i, NB + NU, k) * ysig;
                const double dmt = wb[o * KS + k], dpt = wb[(o + NX) * KS + k];
                const double dm = capd(dmt), dp = capd(dpt);
                const double qm = ady - ds[o * KS + k], qp = -ady - ds[(o + NX) * KS + k];
                const double dt = (rxv_of(mode, csig, PN + i, k) + dm * qm + dp * qp) / (dm + dp);
                const double dzm = dm * (qm - dt), dzp = dp * (qp - dt);
                dz[o * KS + k] = dzm; dz[(o + NX) * KS + k] = dzp;
                dprim[(PN + i) * KS + k] = dt;
                if (mode != 0) {
                    const double dsm = rzs * rz[o * KS + k] - (ady - dt), dsp = rzs * rz[(o + NX) * KS + k] - (-ady - dt);
                    ds[o * KS + k] = dsm; ds[(o + NX) * KS + k] = dsp;
                    {
                        const double iw = sqrt(dmt), il = 1. / lam[o * KS + k];
                        const double dzt = dzm / iw, dst = dsm * iw;
                        tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                        if (mode == 1) cr[o * KS + k] = dst * dzt;
                    }
                    {
                        const double iw = sqrt(dpt), il = 1. / lam[(o + NX) * KS + k];
                        const double dzt = dzp / iw, dst = dsp * iw;
                        tmax = fmax(tmax, fmax(-dst, -dzt) * il);
                        if (mode == 1) cr[(o + NX) * KS + k] = dst * dzt;
                    }
                }
            } else {
                dprim[(PN + i) * KS + k] = 0.; dz[o * KS + k] = 0.; dz[(o + NX) * KS + k] = 0.;
                if (mode != 0) { ds[o * KS + k] = 0.; ds[(o + NX) * KS + k] = 0.; if (mode == 1) { cr[o * KS + k] = 0.; cr[(o + NX) * KS + k] = 0.; } }
            }
        }
    }
    return tmax;
}

// one Newton solve (see phase_solve): passes and triangular products on the whole CTA, the two recursions on warp 0
SCPP_HD void cp_solve(int mode, double csig, double sigmu, double rzs, double &tmax_out)
{
    const double gsig = cp_rhs(mode, csig, sigmu);
    if (cta_tid() < LANES) cp_seq_forward();
    cta_sync();
    const double ldot = cp_par_forward();
    Glob g;
    globals_mid(mode, csig, sigmu, gsig, ldot, g);           // every thread computes the same scalars (reads only)
    cp_par_backward(g.ysig);
    if (cta_tid() < LANES) cp_seq_backward();
    cta_sync();
    double tmax = cp_recover(mode, csig, rzs, g.ysig);
    if (cta_tid() == 0) tmax = fmax(tmax, globals_recover(mode, rzs, g));
    tmax_out = cta_max(tmax);
}

// factorisation: parallel assembly, then the chain; true when every block was positive definite
SCPP_HD bool cp_factor()
{
    const double corner = cp_assemble();
    return cp_chain_factor(corner);
}

// ---- light CTA passes of the cold start -----------------------------------------------------------------------------------------------
SCPP_HD void cp_eval_slack(double *out)
{
    const double sg = prim[PSN * KS], dsg = prim[PSN * KS + 1];
    FOR_CTA(k, K) {
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
    if (cta_tid() == 0 && !sig_fixed) { const int r0 = RS * KS; out[r0] = sg - 0.001; out[r0 + 1] = 0.5 + 0.5 * dsg; out[r0 + 2] = 0.5 - 0.5 * dsg; out[r0 + 3] = sg - sigbar; }
    cta_sync();
}
template <class F>
SCPP_HD void cp_for_cones(F &&f) const
{
    FOR_CTA(k, K) {
        for (int r = 0; r < NLP; r++) f(r * KS + k, KS, 1);
        for (int c = 0; c < NCONE; c++) f((NLP + M::cone_off(c)) * KS + k, KS, M::cone_dim(c));
        f(TRO * KS + k, KS, D);
        if (k < K - 1) for (int r = MN; r < RS; r++) f(r * KS + k, KS, 1);
    }
    if (cta_tid() == 0 && !sig_fixed) { f(RS * KS, 1, 1); f(RS * KS + 1, 1, 3); }
}
SCPP_HD void cp_cone_margin(const double *u, double &mn, double &nrm2)
{
    double lmn = 1e300, n2 = 0;
    cp_for_cones([&](int o, int st_, int d) {
        double t = 0;
        for (int i = 1; i < d; i++) t += u[o + i * st_] * u[o + i * st_];
        const double mg = u[o] - sqrt(t);
        if (mg < lmn) lmn = mg;
        n2 += t + u[o] * u[o];
    });
    mn = -cta_max(-lmn);
    double r[1] = {n2};
    cta_sum(r);
    nrm2 = r[0];
}
SCPP_HD void cp_cone_shift(double *u, double a) { cp_for_cones([&](int o, int, int) { u[o] += a; }); cta_sync(); }
SCPP_HD void cp_cone_shift_each(double *u, double target)      // see cone_shift_each
{
    cp_for_cones([&](int o, int st_, int d) {
        double t = 0;
        for (int i = 1; i < d; i++) t += u[o + i * st_] * u[o + i * st_];
        const double mg = u[o] - sqrt(t);
        if (mg < target) u[o] += target - mg;
    });
    cta_sync();
}

// starting point (see init_point): 1 = previous interior point pulled back, 2 = least-squares start, 0 = its factorisation failed
SCPP_HD int cp_init_point(const IpmSettings &st_, bool have_prev)
{
    const int np = n_prim(K), m = m_rows(K);
    Norms nm;
    double tm;
    if (have_prev && st_.warm > 0. && st_.warm < 1.) {
        const double lw = st_.warm, lc = 1. - st_.warm;
        FOR_CTA(k, K) { for (int i = 0; i < NB; i++) if (fixed(k, i)) prim[i * KS + k] = fixv[k * NB + i]; if (scvx) prim[NB * KS + k] = tr_rad; }
        if (sig_fixed && cta_tid() == 0) { prim[PSN * KS] = sigbar; prim[PSN * KS + 1] = 0.; }
        FOR_CTA(e, m) { s[e] *= lw; z[e] *= lw; }
        cta_sync();
        cp_cone_shift(s, lc); cp_cone_shift(z, lc);
        return 1;
    }
    FOR_CTA(e, np) prim[e] = 0.;
    FOR_CTA(e, m) { s[e] = 0.; z[e] = 0.; }
    cta_sync();
    FOR_CTA(k, K) { for (int i = 0; i < NB; i++) prim[i * KS + k] = fixed(k, i) ? fixv[k * NB + i] : xibar(k, i); if (scvx) prim[NB * KS + k] = tr_rad; }
    if (cta_tid() == 0) { prim[PSN * KS] = sigbar; prim[PSN * KS + 1] = 0.; }
    cta_sync();
    cp_cone_shift(s, 1.); cp_cone_shift(z, 1.);
    cp_residuals(nm, true);
    if (!cp_factor()) return 0;
    cp_eval_slack(ds);
    FOR_CTA(e, np) dprim[e] = 0.;
    cta_sync();
    cp_solve(0, 0., 0., 0., tm);
    FOR_CTA(e, np) prim[e] += dprim[e];
    cta_sync();
    cp_eval_slack(s);
    {
        double mg, n2; cp_cone_margin(s, mg, n2);
        if (mg <= 1e-8 * fmax(1., sqrt(n2))) cp_cone_shift_each(s, INIT_MARGIN);
    }
    FOR_CTA(e, m) ds[e] = 0.;
    FOR_CTA(e, np) dprim[e] = 0.;
    cta_sync();
    FOR_CTA(k, K) { dprim[NB * KS + k] = -w_tr; if (k < K - 1) for (int i = 0; i < NX; i++) dprim[(PN + i) * KS + k] = -w_vc; }
    if (cta_tid() == 0) { dprim[PSN * KS] = -w_time; dprim[PSN * KS + 1] = -w_trs; }
    cta_sync();
    cp_solve(0, 0., 0., 0., tm);
    FOR_CTA(e, m) z[e] = dz[e];
    cta_sync();
    {
        double mg, n2; cp_cone_margin(z, mg, n2);
        if (mg <= 1e-8 * fmax(1., sqrt(n2))) cp_cone_shift_each(z, INIT_MARGIN);
    }
    return 2;
}

// the whole sub-problem on one CTA (the loop of solve(), without slicing: the instance keeps its SM until it is done)
SCPP_HD IpmResult cp_solve_subproblem(const IpmSettings &st_, bool have_prev)
{
    IpmResult res;
    res.status = 1; res.iterations = 0; res.pres = res.dres = res.gap = res.relgap = res.pcost = 0.; res.point_ok = 1; res.pad_ = 0;
    const int np = n_prim(K);
    cp_tables();
    dcap = 0.;
    if (cp_init_point(st_, have_prev) == 0) { res.status = 2; return res; }
    const double cnorm = sqrt(w_time * w_time + w_trs * w_trs + K * w_tr * w_tr + (K - 1) * NX * w_vc * w_vc);
    const double resx0 = fmax(1., cnorm);
    const int deg = degree();
    Norms nm;
    double best = 1e300, pending = 0.;
    int it = 0;
#pragma unroll 1
    for (; it <= st_.maxit; it++) {
        if (pending != 0.) cp_update(pending);
        cp_residuals(nm, false);
        const double resz0 = fmax(1., sqrt(nm.h2));
        const double pres = sqrt(nm.rz2) / resz0, dres = sqrt(nm.rx2) / resx0, gap = nm.gap, pcost = nm.pcost;
        const double dcost = pcost - gap + nm.zrz - nm.xrx;
        double relgap = 1e300;
        if (pcost < 0.) relgap = gap / -pcost; else if (dcost > 0.) relgap = gap / dcost;
        const double score = fmax(fmax(pres, dres) / st_.feastol, fmin(gap / st_.abstol, relgap / st_.reltol));
#if !defined(__CUDACC__)
        if (getenv("SCPP_DEBUG")) fprintf(stderr, "cta it %2d pres %.2e dres %.2e gap %.2e relgap %.2e pcost %.6e bad %d\n", it, pres, dres, gap, relgap, pcost, nm.bad);
#endif
        if (!nm.bad && score < best) {
            best = score;
            res.pres = pres; res.dres = dres; res.gap = gap; res.relgap = relgap; res.pcost = pcost; res.iterations = it;
            if (score <= 1e4) { FOR_CTA(e, np) best_[e] = prim[e]; }
            cta_sync();
        }
        if (!nm.bad && pres <= st_.feastol && dres <= st_.feastol && (gap <= st_.abstol || relgap <= st_.reltol)) { res.status = 0; break; }
        if (nm.bad || it == st_.maxit || (score > 1e3 * best && best < 1e4) || (best <= 10. && score > best)) { res.status = nm.bad ? 2 : (it == st_.maxit ? 1 : 2); res.point_ok = !nm.bad; break; }
        bool factored = cp_factor();
        while (!factored && tighten_cap()) factored = cp_factor();
        if (!factored) { res.status = 2; break; }
        double tmax;
        cp_solve(1, 1., 0., -1., tmax);
        const double a_aff = tmax <= 1. ? 1. : 1. / tmax;
        const double sig = (1. - a_aff) * (1. - a_aff) * (1. - a_aff), mu = gap / deg;
        cp_solve(2, 1. - sig, sig * mu, -(1. - sig), tmax);
        pending = tmax <= step_frac ? 1. : step_frac / tmax;
    }
    if (res.status != 0) {
        if (best <= 1e4) { FOR_CTA(e, np) prim[e] = best_[e]; }
        cta_sync();
        if (best <= 1e4) res.status = 3;
    } else res.iterations = it;
    cta_sync();
    return res;
}
