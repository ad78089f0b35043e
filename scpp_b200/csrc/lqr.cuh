// scpp_b200/csrc/lqr.cuh — kernel K5 body: LQR tracking gains along a trajectory, one WARP per (instance, node).
//
// Replaces ComputeLQR / careSolve / solveSchurIterative (scpp_core/src/LQR.cpp:7-109) as used by LQRTracker::LQRTracker
// (scpp_core/src/LQRTracker.cpp:6-28): Jacobians at (x_k, u_k), Hamiltonian M = [A, -B R^-1 B'; -Q, -A'], matrix sign function by
// Newton's iteration M <- M - (M - M^-1)/2 until isApprox(1e-8) (at most 100 steps), then the stabilising solution from
// [M12; M22 + I] P = -[M11 + I; M21] by LU with complete pivoting (Eigen FullPivLU::solve of the rectangular system), K = R^-1 B' P.
// All matrices live in the warp's shared-memory window; the 2n x 2n inverse is a Gauss-Jordan sweep with partial pivoting in which
// the lanes share the (row, column) pairs of both halves of the augmented matrix.
#pragma once
#include "models.cuh"

namespace scpp {

template <class M>
struct Lqr {
    static constexpr int NX = M::NX, NU = M::NU, N2 = 2 * NX;
    // window (doubles): Ml | W | Inv | A | B | column | P
    static constexpr int O_ML = 0, O_W = O_ML + N2 * N2, O_INV = O_W + N2 * N2, O_A = O_INV + N2 * N2, O_B = O_A + NX * NX,
                         O_COL = O_B + NX * NU + (NX * NU & 1), O_P = O_COL + N2, O_END = O_P + NX * NX;
    SCPP_HD static int sm_doubles() { return O_END + (O_END & 1); }

    // index of the first largest |v| among candidates: every lane passes its best (value, linear index); ties go to the smaller index
    SCPP_HD static int arg_first_max(double v, int idx)
    {
        const double vm = warp_max(v);
        return warp_min_i(v == vm ? idx : 0x7fffffff);
    }

    // Inv <- W^-1 (W destroyed), Gauss-Jordan with partial pivoting
    SCPP_HD static void inverse(double *W, double *Inv, double *col)
    {
        FOR_LANE(e, N2 * N2) Inv[e] = (e / N2 == e % N2) ? 1. : 0.;
        warp_sync();
#pragma unroll 1
        for (int c = 0; c < N2; c++) {
            double best = -1.; int bi = 0x7fffffff;
            for (int r = c + lane_id(); r < N2; r += LANES) { const double a = fabs(W[r * N2 + c]); if (a > best) { best = a; bi = r; } }
            const int p = arg_first_max(best, bi);
            if (p != c) {
                FOR_LANE(j, 2 * N2) {
                    double *row_c = (j < N2) ? W + c * N2 + j : Inv + c * N2 + (j - N2), *row_p = (j < N2) ? W + p * N2 + j : Inv + p * N2 + (j - N2);
                    const double t = *row_c; *row_c = *row_p; *row_p = t;
                }
                warp_sync();
            }
            const double d = 1. / W[c * N2 + c];
            FOR_LANE(r, N2) col[r] = W[r * N2 + c];
            warp_sync();
            FOR_LANE(j, 2 * N2) { double *q = (j < N2) ? W + c * N2 + j : Inv + c * N2 + (j - N2); *q *= d; }
            warp_sync();
            FOR_LANE(e, N2 * 2 * N2) {
                const int r = e / (2 * N2), j = e - r * (2 * N2);
                if (r == c) continue;
                const double l = col[r];
                if (j < N2) W[r * N2 + j] -= l * W[c * N2 + j]; else Inv[r * N2 + (j - N2)] -= l * Inv[c * N2 + (j - N2)];
            }
            warp_sync();
        }
    }

    // K_out [NU][NX] row-major; returns the success flag of careSolve (false: the sign iteration did not converge in 100 steps)
    SCPP_HD static bool gain(const double *x, const double *u, const double *par, const double *qd, const double *rd, double *K_out, double *sm)
    {
        double *Ml = sm + O_ML, *W = sm + O_W, *Inv = sm + O_INV, *A = sm + O_A, *B = sm + O_B, *col = sm + O_COL, *P = sm + O_P;
        typename M::Lin L;
        M::linearize(x, u, par, L);
        FOR_LANE(j, NX) { double e[NX], o[NX]; for (int i = 0; i < NX; i++) e[i] = (i == j); M::A_apply(L, e, o); for (int i = 0; i < NX; i++) A[i * NX + j] = o[i]; }
        FOR_LANE(j, NU) { double e[NU], o[NX]; for (int i = 0; i < NU; i++) e[i] = (i == j); M::B_apply(L, e, o); for (int i = 0; i < NX; i++) B[i * NU + j] = o[i]; }
        warp_sync();
        FOR_LANE(e, N2 * N2) {
            const int r = e / N2, c = e - r * N2;
            double v;
            if (r < NX && c < NX) v = A[r * NX + c];
            else if (r < NX) { double brb = 0.; for (int l = 0; l < NU; l++) brb += B[r * NU + l] * (1. / rd[l]) * B[(c - NX) * NU + l]; v = -brb; }
            else if (c < NX) v = (r - NX == c) ? -qd[c] : 0.;
            else v = -A[(c - NX) * NX + (r - NX)];
            Ml[e] = v;
        }
        warp_sync();
        // solveSchurIterative, LQR.cpp:7-33
        int iterations = 0;
        bool converged = false;
#pragma unroll 1
        while (!converged) {
            if (iterations > 100) return false;
            FOR_LANE(e, N2 * N2) W[e] = Ml[e];
            warp_sync();
            inverse(W, Inv, col);
            double d2 = 0., a2 = 0., b2 = 0.;
            FOR_LANE(e, N2 * N2) {
                const double ml = Ml[e], mdiff = ml - Inv[e], mn = ml - 0.5 * mdiff, d = mn - ml;
                d2 += d * d; a2 += mn * mn; b2 += ml * ml;
                Ml[e] = mn;
            }
            d2 = warp_sum(d2); a2 = warp_sum(a2); b2 = warp_sum(b2);
            converged = d2 <= 1e-8 * 1e-8 * (a2 < b2 ? a2 : b2);
            iterations++;
            warp_sync();
        }
        // [M12; M22 + I] P = -[M11 + I; M21] : a (N2 x NX) in W, right-hand side c (N2 x NX) in Inv   (LQR.cpp:35-51)
        double *a = W, *cc = Inv;
        FOR_LANE(e, NX * NX) {
            const int i = e / NX, j = e - i * NX;
            a[i * NX + j] = Ml[i * N2 + NX + j];
            a[(NX + i) * NX + j] = Ml[(NX + i) * N2 + NX + j] + (i == j);
            cc[i * NX + j] = -(Ml[i * N2 + j] + (i == j));
            cc[(NX + i) * NX + j] = -Ml[(NX + i) * N2 + j];
        }
        int *colperm = reinterpret_cast<int *>(col);
        FOR_LANE(j, NX) colperm[j] = j;
        warp_sync();
        double maxpivot = 0.;
        int nonzero = NX;
#pragma unroll 1
        for (int k = 0; k < NX; k++) {
            double best = -1.; int bi = 0x7fffffff;
            const int nr = N2 - k, nc = NX - k;
            for (int e = lane_id(); e < nr * nc; e += LANES) {
                const int i = k + e / nc, j = k + e % nc;
                const double v = fabs(a[i * NX + j]);
                const int lin = i * NX + j;
                if (v > best || (v == best && lin < bi)) { best = v; bi = lin; }
            }
            const double vm = warp_max(best);
            const int lin = warp_min_i(best == vm ? bi : 0x7fffffff);
            if (vm == 0.) { nonzero = k; break; }
            if (vm > maxpivot) maxpivot = vm;
            const int pr = lin / NX, pc = lin - pr * NX;
            if (pr != k) {
                FOR_LANE(j, 2 * NX) { double *q0 = (j < NX) ? a + k * NX + j : cc + k * NX + (j - NX), *q1 = (j < NX) ? a + pr * NX + j : cc + pr * NX + (j - NX); const double t = *q0; *q0 = *q1; *q1 = t; }
                warp_sync();
            }
            if (pc != k) {
                FOR_LANE(i, N2) { const double t = a[i * NX + k]; a[i * NX + k] = a[i * NX + pc]; a[i * NX + pc] = t; }
                if (lane_id() == 0) { const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t; }
                warp_sync();
            }
            const double piv = a[k * NX + k];
            FOR_LANE(i, N2) if (i > k) a[i * NX + k] = a[i * NX + k] / piv;
            warp_sync();
            FOR_LANE(e, (N2 - k - 1) * 2 * NX) {
                const int i = k + 1 + e / (2 * NX), j = e % (2 * NX);
                const double l = a[i * NX + k];
                if (j < NX) { if (j > k) a[i * NX + j] -= l * a[k * NX + j]; }
                else cc[i * NX + (j - NX)] -= l * cc[k * NX + (j - NX)];
            }
            warp_sync();
        }
        int rank = 0;
        for (int k = 0; k < nonzero; k++) if (fabs(a[k * NX + k]) > maxpivot * 2.220446049250313e-16 * N2) rank++;
        FOR_LANE(j, NX) {            // one right-hand-side column per lane
            double y[NX];
            for (int r = rank - 1; r >= 0; r--) {
                double acc = cc[r * NX + j];
                for (int q = r + 1; q < rank; q++) acc -= a[r * NX + q] * y[q];
                y[r] = acc / a[r * NX + r];
            }
            for (int i = 0; i < NX; i++) P[i * NX + j] = 0.;
            for (int r = 0; r < rank; r++) P[colperm[r] * NX + j] = y[r];
        }
        warp_sync();
        FOR_LANE(e, NU * NX) {
            const int l = e / NX, j = e - l * NX;
            double acc = 0.;
            for (int i = 0; i < NX; i++) acc += B[i * NU + l] * P[i * NX + j];
            K_out[e] = acc / rd[l];
        }
        warp_sync();
        return true;
    }
};

} // namespace scpp
