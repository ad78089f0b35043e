// scpp_b200/csrc/blockops.cuh — small dense block products of the SOCP kernel on the FP64 tensor cores.
//
// The per-node contractions of the reduced KKT system (A~' D A~, the Schur update L L', L_{k+1,k} = O L^-T, C' D A~)
// are genuine (nx+nu)-sized GEMMs; on sm_100a they run as mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 — tcgen05 has
// no FP64 path).  Operands live in the shared-memory window; element accessors return 0 outside the matrix so the
// 18-wide blocks need no physical padding.  Measured on this box: DMMA 37.1 TFLOP/s vs DFMA 34.7 TFLOP/s — the win is
// not peak rate but instruction count: one DMMA replaces 8 DFMA + 16 LDS of the scalar loops.
// The host-simulation build (LANES == 1) uses plain loops with the same accessors.
#pragma once
#include "portable.cuh"

namespace scpp {
namespace blk {

#if defined(__CUDA_ARCH__)
SCPP_D void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// For every (m,n) of an M x N output (only tiles on or below the block diagonal when LOWER):
//      fc(m, n, sum_k fa(m,k) * fb(k,n)),   k < Kd
// fa / fb must return 0 outside their matrices; fc must ignore (m,n) outside the output.
// Fragment layout of m8n8k4.f64: A row-major 8x4: lane holds A[lane>>2][lane&3]; B col-major 4x8: lane holds
// B[lane&3][lane>>2]; C 8x8: lane holds C[lane>>2][2*(lane&3)] and the next column.
template <int M, int N, int Kd, bool LOWER, class FA, class FB, class FC>
SCPP_HD void mm(FA fa, FB fb, FC fc)
{
#if defined(__CUDA_ARCH__)
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    constexpr int MT = (M + 7) / 8, NT = (N + 7) / 8, KT = (Kd + 3) / 4;
#pragma unroll
    for (int mi = 0; mi < MT; mi++) {
        double af[KT];
#pragma unroll
        for (int kk = 0; kk < KT; kk++) af[kk] = fa(mi * 8 + g, kk * 4 + t);
#pragma unroll
        for (int ni = 0; ni < NT; ni++) {
            if (LOWER && ni > mi) continue;
            double c0 = 0., c1 = 0.;
#pragma unroll
            for (int kk = 0; kk < KT; kk++) dmma(c0, c1, af[kk], fb(kk * 4 + t, ni * 8 + g));
            fc(mi * 8 + g, ni * 8 + 2 * t, c0);
            fc(mi * 8 + g, ni * 8 + 2 * t + 1, c1);
        }
    }
#else
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            if (LOWER && n / 8 > m / 8) continue;
            double v = 0.;
            for (int k = 0; k < Kd; k++) v += fa(m, k) * fb(k, n);
            fc(m, n, v);
        }
#endif
}

// One 8 x 8 output tile (mi, ni) of the same product (Kd = 4 KT): the unit of work when several warps share a block product.
template <int KT, class FA, class FB, class FC>
SCPP_HD void mm_tile(int mi, int ni, FA fa, FB fb, FC fc)
{
#if defined(__CUDA_ARCH__)
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    double c0 = 0., c1 = 0.;
#pragma unroll
    for (int kk = 0; kk < KT; kk++) dmma(c0, c1, fa(mi * 8 + g, kk * 4 + t), fb(kk * 4 + t, ni * 8 + g));
    fc(mi * 8 + g, ni * 8 + 2 * t, c0);
    fc(mi * 8 + g, ni * 8 + 2 * t + 1, c1);
#else
    for (int m = mi * 8; m < mi * 8 + 8; m++)
        for (int n = ni * 8; n < ni * 8 + 8; n++) {
            double v = 0.;
            for (int k = 0; k < 4 * KT; k++) v += fa(m, k) * fb(k, n);
            fc(m, n, v);
        }
#endif
}

} // namespace blk
} // namespace scpp
