// scpp_b200/csrc/discretize_shared.cuh — K1, round-2 mapping (cfg.jacobian = 2): the linearisation of an interval is computed ONCE per
// right-hand-side evaluation and shared by its columns.
//
// Same mathematics as discretize.cuh (forward-sensitivity form of discretizationImplementation.hpp:38-181, classical RK4 with the
// Richardson pair), other mapping.  In discretize.cuh every column thread integrates its own copy of x(tau) and re-runs M::linearize:
// in SIMT terms the 24 copies cost one warp instruction each, but that instruction stream is 3/4 of what a warp executes (linearize ~300
// flop against ~50 for the column's own sparse A v), so the FP64 pipe spends most of its time on 24 identical results.
// Here a CTA owns IPB intervals and splits the ROLES between warps:
//   * one CHAIN warp, lane i <-> interval i: integrates x(tau) alone (f only), TWO RK4 steps ahead of the columns, and leaves the four
//     stage points (x, u) of a step in shared memory;
//   * a LINEARISER warp (two in the large shape), lane <-> (interval, stage), ONE step ahead: per stage point a record { Lin (f, sparse A, B), columns of B,
//     g = A x + B u } in shared memory (the Jacobian work is off the sequential chain x_s -> f -> x_s+1 and parallel over the stages);
//   * IPB CONSUMER warps (6; 13 in the one-CTA-per-SM shape), warp w <-> interval w, lane c <-> column c (coalesced tile stores as before): per stage only their own
//     sigma A v + w_c V_c out of the record (A entries by broadcast loads; the forcing vector V_c -- a column of B, f or g -- by lane,
//     so all column types run the same instructions), and the RK4 update of one 14-vector.
// Both stashes are double-buffered by step parity; one __syncthreads() per RK4 step hands the buffers on.  The step bodies are SCPP_HD functions, so the host-simulation build
// (tests/_hostsim) runs the same schedule sequentially and the arithmetic is unit-tested without a GPU.
#pragma once
#include "discretize.cuh"
#include <stddef.h>

namespace scpp {

// padding (in doubles) between Lin and the B columns such that the forcing vectors of a record -- f, the NU columns of B, g: read at the
// same element index by different lanes of a consumer warp -- start in different shared-memory banks
template <class M>
SCPP_HD constexpr int k1s_pad()
{
    constexpr int NX = M::NX, NU = M::NU, LD = int(sizeof(typename M::Lin) / 8);
    for (int pad = 1; pad < 17; pad++) {
        int bank[NU + 2] = {};
        bank[0] = 0;                                              // f: first member of every Lin
        for (int j = 0; j < NU; j++) bank[1 + j] = (2 * (LD + pad + j * NX)) % 32;
        bank[NU + 1] = (2 * (LD + pad + NU * NX)) % 32;
        bool ok = true;
        for (int a = 0; a < NU + 2; a++)
            for (int b = a + 1; b < NU + 2; b++) ok = ok && bank[a] != bank[b];
        if (ok) return pad;
    }
    return 1;
}

template <class M>
struct StageLin {
    typename M::Lin L;
    double pad_[k1s_pad<M>()];
    double Bd[M::NU][M::NX];   // columns of B (what the B and C columns integrate, weighted by the hold functions)
    double g[M::NX];           // A x + B u at the stage (the z column integrates -sigma g)
};

// numbers of doubles between the records of two intervals in the shared-memory stashes: 4 stages, odd (the lanes of a helper warp write
// the same member of different intervals' records: an odd stride in doubles spreads them over the banks)
template <class M>
SCPP_HD constexpr int k1s_stride() { return int(4 * sizeof(StageLin<M>) / 8) | 1; }
template <class M>
SCPP_HD constexpr int k1s_xstride() { return (4 * (M::NX + M::NU)) | 1; }

SCPP_HD void k1s_schedule(int nsub, int s, int &ns, int &st)      // global step s -> (sub-steps of its pass, step inside the pass)
{
    if (nsub > 0) { ns = nsub; st = s; }
    else if (s < -nsub) { ns = -nsub; st = s; }
    else { ns = -2 * nsub; st = s + nsub; }
}
SCPP_HD int k1s_steps(int nsub) { return nsub > 0 ? nsub : -3 * nsub; }
// step sizes of the first / second pass and 1 / dtau: computed once, no division inside the step loops
SCPP_HD void k1s_step_sizes(int nsub, int K, double &h0, double &h1, double &rdtau)
{
    const double dtau = 1. / double(K - 1);
    rdtau = double(K - 1);
    h0 = dtau / (nsub > 0 ? nsub : -nsub);
    h1 = dtau / (nsub > 0 ? nsub : -2 * nsub);
}

// CHAIN: RK4 step `st` (of ns) of x alone; leaves the four stage points (x, u) in xs[4][NX + NU].  Only f of M::linearize is used (the
// rest of the inlined call is dead code), so x(tau) is bit for bit the trajectory the column kernel integrates.
template <class M>
SCPP_HD void k1s_chain(double *x, const double *x0, const double *u0, const double *du, const double *par, double sigma, double h, double rdtau,
                       int st, double *xs)
{
    constexpr int NX = M::NX, NU = M::NU;
    if (st == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) x[i] = x0[i];
    }
    const double t0 = st * h, hs = sigma * h, h6 = hs * (1. / 6.);
    // x at the start of the step is the stage-0 point in xs: it is read back from there instead of being held in registers (the chain
    // keeps the accumulator x[] and the stage point xt[] only)
    double xt[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; i++) xt[i] = x[i];
#pragma unroll
    for (int sgi = 0; sgi < 4; sgi++) {      // unrolled: stage times and weights are immediates
        const double tau = t0 + (sgi == 0 ? 0. : (sgi == 3 ? h : 0.5 * h));
        const double beta = tau * rdtau;
#pragma unroll
        for (int j = 0; j < NU; j++) u[j] = u0[j] + beta * du[j];
        double *o = xs + sgi * (NX + NU);
#pragma unroll
        for (int i = 0; i < NX; i++) o[i] = xt[i];
#pragma unroll
        for (int j = 0; j < NU; j++) o[NX + j] = u[j];
        typename M::Lin L;
        M::linearize(xt, u, par, L);
        const double wgt = (sgi == 0 || sgi == 3) ? h6 : 2. * h6;      // sigma folded into the step: h6 = sigma h / 6, hs = sigma h
        const double nxt = (sgi == 2) ? hs : 0.5 * hs;
#pragma unroll
        for (int i = 0; i < NX; i++) {
            x[i] += wgt * L.f[i];
            if (sgi < 3) xt[i] = xs[i] + nxt * L.f[i];
        }
    }
}

// LINEARISER: the record of one stage point
template <class M>
SCPP_HD void k1s_linearize(const double *xu, const double *par, StageLin<M> &R)
{
    constexpr int NX = M::NX, NU = M::NU;
    double xt[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; i++) xt[i] = xu[i];
#pragma unroll
    for (int j = 0; j < NU; j++) u[j] = xu[NX + j];
    M::linearize(xt, u, par, R.L);
    double Ax[NX], Bu[NX];
    M::A_apply(R.L, xt, Ax);
    M::B_apply(R.L, u, Bu);
#pragma unroll
    for (int i = 0; i < NX; i++) R.g[i] = Ax[i] + Bu[i];
#pragma unroll
    for (int j = 0; j < NU; j++) {
        double e[NU], Be[NX];
#pragma unroll
        for (int q = 0; q < NU; q++) e[q] = (q == j) ? 1. : 0.;
        M::B_apply(R.L, e, Be);
#pragma unroll
        for (int i = 0; i < NX; i++) R.Bd[j][i] = Be[i];
    }
}

// the forcing term of a column as (offset of a 14-vector inside the stage record, weight): the same code for every lane of a warp
//   A columns: 0 * g      B column j: sigma alpha * B e_j      C column j: sigma beta * B e_j      s column: 1 * f      z column: -sigma * g
template <class M>
SCPP_HD int k1s_forcing_offset(int ctype, int cidx)
{
    using R = StageLin<M>;
    if (ctype == 1 || ctype == 2) return int(offsetof(R, Bd) / 8) + cidx * M::NX;
    if (ctype == 3) return int((offsetof(R, L) + offsetof(typename M::Lin, f)) / 8);
    return int(offsetof(R, g) / 8);
}

// CONSUMER: RK4 step `st` of one column out of the four stage records.  sigma is folded into the step:  sigma (A v + w V) h  =  (A v + w V) (sigma h),
// so a stage costs three fused multiply-adds per element on top of the sparse A v  (w: alpha | beta | 1 / sigma | -1 | 0 by column type)
template <class M>
SCPP_HD void k1s_consume(double *col, int ctype, int voff, double sigma, double rsigma, double h, double rdtau, int st, const StageLin<M> *rec, bool zoh = false)
{
    constexpr int NX = M::NX;
    const double t0 = st * h, hs = sigma * h, h6 = hs * (1. / 6.);
    const double wa = ctype == 1 ? 1. : 0., wb = ctype == 2 ? 1. : 0., wc = ctype == 3 ? rsigma : (ctype == 4 ? -1. : 0.);
    double ca[NX], ct[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) { ca[i] = col[i]; ct[i] = col[i]; }
#pragma unroll
    for (int sgi = 0; sgi < 4; sgi++) {      // unrolled: stage weights, record offsets and hold weights fold into immediates
        const double tau = t0 + (sgi == 0 ? 0. : (sgi == 3 ? h : 0.5 * h));
        const double beta = zoh ? 0. : tau * rdtau, alpha = 1. - beta;      // zero-order hold: B integrates the whole input matrix, C nothing
        const double w = wa * alpha + (wb * beta + wc);      // without a branch on the column type
        const StageLin<M> &R = rec[sgi];
        const double *V = reinterpret_cast<const double *>(&R) + voff;
        double kc[NX];
        M::A_apply(R.L, ct, kc);
        const double wgt = (sgi == 0 || sgi == 3) ? h6 : 2. * h6;
        const double nxt = (sgi == 2) ? hs : 0.5 * hs;
#pragma unroll
        for (int i = 0; i < NX; i++) {
            const double k = kc[i] + w * V[i];
            ca[i] += wgt * k;
            kc[i] = col[i] + nxt * k;
        }
#pragma unroll
        for (int i = 0; i < NX; i++) ct[i] = kc[i];
    }
#pragma unroll
    for (int i = 0; i < NX; i++) col[i] = ca[i];
}

SCPP_HD void k1s_column_type(int NX, int NU, int c, int &ctype, int &cidx)
{
    if (c < NX) { ctype = 0; cidx = c; }
    else if (c < NX + NU) { ctype = 1; cidx = c - NX; }
    else if (c < NX + 2 * NU) { ctype = 2; cidx = c - NX - NU; }
    else if (c == NX + 2 * NU) { ctype = 3; cidx = 0; }
    else { ctype = 4; cidx = 0; }
}

// what a consumer does around step s: start of a pass (unit column), the step, and after the last step the Richardson combination.
// The result of the first pass is parked in colc[i * cstride] (on the GPU: the column's own slot of the output tile, so it holds no registers).
template <class M>
SCPP_HD void k1s_consumer_step(double *col, double *colc, int cstride, int ctype, int cidx, double sigma, double rsigma, double h0, double h1, double rdtau, int nsub, int s,
                               const StageLin<M> *rec, bool zoh = false)
{
    constexpr int NX = M::NX;
    int ns, st;
    k1s_schedule(nsub, s, ns, st);
    if (st == 0) {
        if (s > 0) {
#pragma unroll
            for (int i = 0; i < NX; i++) colc[i * cstride] = col[i];
        }
#pragma unroll
        for (int i = 0; i < NX; i++) col[i] = (ctype == 0 && i == cidx) ? 1. : 0.;
    }
    k1s_consume<M>(col, ctype, k1s_forcing_offset<M>(ctype, cidx), sigma, rsigma, (nsub < 0 && s >= -nsub) ? h1 : h0, rdtau, st, rec, zoh);
    if (nsub < 0 && s == k1s_steps(nsub) - 1) {      // y = y_2n + (y_2n - y_n) / 15 removes the h^4 term (discretize.cuh)
#pragma unroll
        for (int i = 0; i < NX; i++) col[i] += (col[i] - colc[i * cstride]) * (1. / 15.);
    }
}

#if defined(__CUDACC__)
#ifndef SCPP_K1S_SMALL
#define SCPP_K1S_SMALL 1
#endif
// CTA shape.  0: 13 consumer warps + chain warp + 2 lineariser warps (two stages per warp) = 512 threads x 128 registers, one CTA per SM.
//            1 (default): 6 consumer warps + chain warp + 1 lineariser warp (four stages) = 256 threads, two CTAs per SM: while one CTA
//               waits at its step barrier for its chain warp the other one runs (measured: 22.1 against 23.9 ms of K1 per bench step)
constexpr int K1S_LIN = SCPP_K1S_SMALL ? 1 : 2;                // lineariser warps
constexpr int K1S_IPB = SCPP_K1S_SMALL ? 6 : 13;               // intervals per CTA
constexpr int K1S_LSH = SCPP_K1S_SMALL ? 3 : 4;                // a lineariser lane: interval = lane & (2^LSH - 1), stage slot = lane >> LSH
constexpr int K1S_THREADS = (K1S_IPB + 1 + K1S_LIN) * 32;

template <class M>
__host__ __device__ constexpr size_t k1s_smem_bytes() { return (size_t)2 * K1S_IPB * (k1s_stride<M>() + k1s_xstride<M>()) * sizeof(double); }
template <class M>
__host__ __device__ constexpr bool k1s_fits()      // models with a dense generated Lin keep the column kernel (their linearize is NX + NU dual-number passes: too much for one lineariser lane)
{
    return k1s_smem_bytes<M>() * (SCPP_K1S_SMALL ? 2 : 1) <= 200 * 1024 && sizeof(typename M::Lin) / 8 <= 128;
}

template <class M>
__global__ void __launch_bounds__(K1S_THREADS, SCPP_K1S_SMALL ? 2 : 1) k_discretize_shared(ScArrays<M> a, int nsub, int zoh, const int *__restrict__ active, int n_active)
{
    constexpr int NX = M::NX, NU = M::NU, NC = NX + 2 * NU + 2, IPB = K1S_IPB, STRIDE = k1s_stride<M>(), XSTRIDE = k1s_xstride<M>();
    static_assert(NC <= 32 && IPB <= (1 << K1S_LSH) && K1S_LIN * (32 >> K1S_LSH) == 4, "one lane per column; the lineariser lanes cover four stages of every interval");
    extern __shared__ __align__(16) double stash[];
    if constexpr (k1s_fits<M>()) {
        double *LS = stash;                                   // [2][IPB][STRIDE]   stage records
        double *XS = stash + (size_t)2 * IPB * STRIDE;        // [2][IPB][XSTRIDE]  stage points
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int role = warp < IPB ? 0 : (warp == IPB ? 1 : 2);                 // consumer | chain | lineariser
        const int slot = role == 0 ? warp : (role == 1 ? lane : (lane & ((1 << K1S_LSH) - 1)));   // interval of this CTA the thread works for
        const int lstage = (32 >> K1S_LSH) * (warp - IPB - 1) + (lane >> K1S_LSH);                // lineariser: its stage of the step
        const long long total = (long long)n_active * (a.K - 1);
        const long long p = (long long)blockIdx.x * IPB + slot;
        const bool on = slot < IPB && p < total && (role != 0 || lane < NC);
        const int K = a.K;
        const int ai = on ? int(p / (K - 1)) : 0, k = on ? int(p - (long long)ai * (K - 1)) : 0;
        const int n = active ? active[ai] : ai;
        const double sigma = a.sigma[n];
        double h0, h1, rdtau;
        k1s_step_sizes(nsub, K, h0, h1, rdtau);
        const int S = k1s_steps(nsub);
        double par[M::NP];                                    // in registers: what linearize derives from the parameters alone leaves the loops
#pragma unroll
        for (int i = 0; i < M::NP; i++) par[i] = a.par[(size_t)n * M::NP + i];
        // iteration t: the chain computes step t, the linearisers step t - 1, the columns step t - 2; buffers by step parity
        if (role == 1) {
            double x[NX], u0[NU], du[NU];
            const double *X = a.X + (size_t)n * K * NX, *U = a.U + (size_t)n * K * NU;
#pragma unroll
            for (int j = 0; j < NU; j++) { u0[j] = U[NU * k + j]; du[j] = zoh ? 0. : U[NU * (k + 1) + j] - u0[j]; }      // zero-order hold: u_k over the interval
#pragma unroll 1
            for (int t = 0; t < S + 2; t++) {
                if (on && t < S) {
                    int ns, st;
                    k1s_schedule(nsub, t, ns, st);
                    k1s_chain<M>(x, X + NX * k, u0, du, par, sigma, (nsub < 0 && t >= -nsub) ? h1 : h0, rdtau, st, XS + ((size_t)(t & 1) * IPB + slot) * XSTRIDE);
                }
                __syncthreads();
            }
        } else if (role == 2) {
#pragma unroll 1
            for (int t = 0; t < S + 2; t++) {
                if (on && t >= 1 && t - 1 < S) {
                    const int b = (t - 1) & 1;
                    k1s_linearize<M>(XS + ((size_t)b * IPB + slot) * XSTRIDE + lstage * (NX + NU), par,
                                     reinterpret_cast<StageLin<M> *>(LS + ((size_t)b * IPB + slot) * STRIDE)[lstage]);
                }
                __syncthreads();
            }
        } else {
            int ctype, cidx;
            k1s_column_type(NX, NU, lane < NC ? lane : 0, ctype, cidx);
            double col[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) col[i] = 0.;
            const int c = lane;
            double *ddk = a.dd + ((size_t)n * (K - 1) + k) * NX * NC;
            const double rsigma = 1. / sigma;
#pragma unroll 1
            for (int t = 0; t < S + 2; t++) {
                if (on && t >= 2)
                    k1s_consumer_step<M>(col, ddk + c, NC, ctype, cidx, sigma, rsigma, h0, h1, rdtau, nsub, t - 2,
                                         reinterpret_cast<const StageLin<M> *>(LS + ((size_t)(t & 1) * IPB + slot) * STRIDE), zoh != 0);
                __syncthreads();
            }
            if (on) {
#pragma unroll
                for (int i = 0; i < NX; i++) ddk[i * NC + c] = col[i];
                if (a.ddT) {
                    double *ddT = a.ddT + (size_t)n * Ipm<M>::ddt_doubles(K);
                    const int KS = Ipm<M>::ks(K);
#pragma unroll
                    for (int i = 0; i < NX; i++) ddT[(size_t)(i * NC + c) * KS + k] = col[i];
                }
            }
        }
    }
}
#endif

} // namespace scpp
