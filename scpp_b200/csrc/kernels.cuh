// scpp_b200/csrc/kernels.cuh — the heavy kernel templates (K1, K2 monolithic, K2 split pipeline).
// They are compiled in separate translation units (kernels_inst.cu, one group per nvcc process, see scpp_b200/build.py);
// engine.cu only sees explicit-instantiation DECLARATIONS, so the library builds in parallel.
#pragma once
#include "sc.cuh"
#include "discretize_shared.cuh"
#include <cuda_runtime.h>

namespace scpp {

#ifndef SCPP_WPB_MAX
#define SCPP_WPB_MAX 7
#endif
constexpr int WPB_MAX = SCPP_WPB_MAX;   // warps (= problem instances) per CTA of the SOCP kernels (chosen per launch, see EngineT::solve())

// K1: one thread per (active instance, interval, column)
template <class M, bool AD>
__global__ void __launch_bounds__(128) k_discretize(ScArrays<M> a, int nsub, int free_time, const int *__restrict__ active, int n_active)
{
    constexpr int NX = M::NX, NU = M::NU, NC = NX + 2 * NU + 2;
    const int per = (a.K - 1) * NC;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_active * per) return;
    const int ai = int(idx / per), rem = int(idx - (long long)ai * per);
    const int k = rem / NC, c = rem - k * NC;
    const int n = active ? active[ai] : ai;
    discretize_column<M, AD>(a.X + (size_t)n * a.K * NX, a.U + (size_t)n * a.K * NU, a.sigma[n], a.par + (size_t)n * M::NP,
                         a.K, k, c, nsub, free_time, a.dd + ((size_t)n * (a.K - 1) + k) * NX * NC,
                         a.ddT ? a.ddT + (size_t)n * Ipm<M>::ddt_doubles(a.K) : nullptr, Ipm<M>::ks(a.K));
}

// K2 (+K3 epilogue): one warp per active instance
template <class M, int MAXW, int MINB>
__global__ void __launch_bounds__(MAXW * 32, MINB) k_solve(ScArrays<M> a, ScConfig cfg, const int *__restrict__ active, int n_active)
{
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    if (gw >= n_active) return;
    sc_solve_instance<M>(a, cfg, active[gw], smem + (size_t)warp * Ipm<M>::sm_doubles());
}

// K2 (+K3), round 2 (cfg.solver == 1): one CTA per instance, persistent CTAs pulling instances from a device-side queue; a CTA keeps its
// instance for the whole sub-problem (ipm_cta.inl), so an instance that needs 25 interior-point iterations never holds back one that needs 5
// two CTAs of 4 warps per SM when two shared-memory images fit (K <= ~55), one CTA of 8 warps otherwise: 255 registers per thread either way
// (the cone arithmetic of a stage needs them: at 128 registers the kernel spills 10 KB per thread and runs 2x slower)
constexpr int cta_threads_for(int minb) { return minb == 2 ? 128 : 256; }
template <class M, int MINB>
__global__ void __launch_bounds__(cta_threads_for(MINB), MINB) k_solve_cta(ScArrays<M> a, ScConfig cfg, const int *__restrict__ active, int n_active, int *__restrict__ queue)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_idx;
    for (;;) {
        if (threadIdx.x == 0) s_idx = atomicAdd(queue, 1);
        __syncthreads();
        const int idx = s_idx;
        __syncthreads();
        if (idx >= n_active) return;
        sc_solve_instance_cta<M>(a, cfg, active[idx], smem);
    }
}

// ---- split pipeline (cfg.ipm_slice < 0): one kernel per step of the interior-point iteration (sc.cuh: sc_split_step) ----
// warp per instance, full shared window: start of a sub-problem / chain factorisation / substitutions
template <class M, int STEP, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) k_sp_warp(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, const int *__restrict__ n_list_dev,
                                                          int n_list, int mode)
{
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    const int n_l = n_list_dev ? *n_list_dev : n_list;
    if (gw >= n_l) return;
    sc_split_step<M, STEP>(a, cfg, list[gw], smem + (size_t)warp * Ipm<M>::sm_doubles(), mode, 0);
}
// warp per instance, no window: the test at the end of a round (+ K3)
template <class M>
__global__ void __launch_bounds__(128) k_sp_test(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, int n_list)
{
    __shared__ __align__(16) double smem[4 * 16];
    const int warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * 4 + warp;
    if (gw >= n_list) return;
    sc_split_step<M, SP_TEST>(a, cfg, list[gw], smem + warp * 16, 0, 0);
}
// warp per (instance, stage): assembly of the block-tridiagonal system
template <class M, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 2) k_sp_assemble(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, int n_list)
{
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * MAXW + warp;
    if (gw >= (long long)n_list * a.K) return;
    const int li = int(gw / a.K), k = int(gw - (long long)li * a.K);
    sc_split_step<M, SP_ASSEMBLE>(a, cfg, list[li], smem + (size_t)warp * Ipm<M>::asm_doubles(), 0, k);
}
// warp per (instance, 32 stages): stage-parallel passes.  A block holds ALL parts of its instances (ipb instances x parts warps) so that
// the update of the iterate can be separated from the residual pass (which reads the neighbouring stage) by a block barrier.
template <class M, int STEP>
__global__ void __launch_bounds__(128, 3) k_sp_stage(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, int n_list, int parts, int ipb, int mode)
{
    __shared__ __align__(16) double smem[4 * 16];
    const int warp = threadIdx.x >> 5;
    const int slot = warp / parts, w = warp - slot * parts;
    const int li = blockIdx.x * ipb + slot;
    const bool on = slot < ipb && li < n_list;
    const int n = on ? list[li] : 0;
    if (STEP == SP_UPDATE) {
        if (on) sc_split_step<M, SP_UPDATE>(a, cfg, n, smem + warp * 16, mode, w);
        __syncthreads();
        if (on) sc_split_step<M, SP_RESIDUALS>(a, cfg, n, smem + warp * 16, mode, w);
    } else if (on) sc_split_step<M, STEP>(a, cfg, n, smem + warp * 16, mode, w);
}


// ---- explicit instantiations: SCPP_KERNEL_GROUP(model, group) expands to definitions in kernels_inst.cu and to declarations elsewhere
#define SCPP_ARGS_SOLVE(M) (ScArrays<M>, ScConfig, const int *, int)
#define SCPP_ARGS_WARP(M) (ScArrays<M>, ScConfig, const int *, const int *, int, int)
#define SCPP_ARGS_STAGE(M) (ScArrays<M>, ScConfig, const int *, int, int, int, int)
#define SCPP_GROUP0(X, M) X template __global__ void k_solve<M, WPB_MAX, 1> SCPP_ARGS_SOLVE(M);
#define SCPP_GROUP1(X, M)                                                                  \
    X template __global__ void k_solve_cta<M, 2>(ScArrays<M>, ScConfig, const int *, int, int *);  \
    X template __global__ void k_solve_cta<M, 1>(ScArrays<M>, ScConfig, const int *, int, int *);
#define SCPP_GROUP2(X, M) X template __global__ void k_sp_warp<M, SP_START, WPB_MAX> SCPP_ARGS_WARP(M);
#define SCPP_GROUP3(X, M)                                                                  \
    X template __global__ void k_sp_warp<M, SP_FACTOR, WPB_MAX> SCPP_ARGS_WARP(M);          \
    X template __global__ void k_sp_warp<M, SP_CHAIN, WPB_MAX> SCPP_ARGS_WARP(M);           \
    X template __global__ void k_sp_assemble<M, WPB_MAX> SCPP_ARGS_SOLVE(M);                \
    X template __global__ void k_sp_test<M> SCPP_ARGS_SOLVE(M);                             \
    X template __global__ void k_discretize<M, false>(ScArrays<M>, int, int, const int *, int);   \
    X template __global__ void k_discretize<M, true>(ScArrays<M>, int, int, const int *, int);
#define SCPP_GROUP4(X, M)                                                                  \
    X template __global__ void k_sp_stage<M, SP_RHS> SCPP_ARGS_STAGE(M);                    \
    X template __global__ void k_sp_stage<M, SP_RECOVER> SCPP_ARGS_STAGE(M);                \
    X template __global__ void k_sp_stage<M, SP_UPDATE> SCPP_ARGS_STAGE(M);
#define SCPP_GROUP5(X, M) X template __global__ void k_discretize_shared<M>(ScArrays<M>, int, int, const int *, int);
#define SCPP_ALL_GROUPS(X, M) SCPP_GROUP0(X, M) SCPP_GROUP1(X, M) SCPP_GROUP2(X, M) SCPP_GROUP3(X, M) SCPP_GROUP4(X, M) SCPP_GROUP5(X, M)

#if !defined(SCPP_KERNEL_INST)
SCPP_ALL_GROUPS(extern, RocketQuat)
SCPP_ALL_GROUPS(extern, Rocket2d)
SCPP_ALL_GROUPS(extern, Rocket2dPlugin)
SCPP_ALL_GROUPS(extern, RocketQuatRollPlugin)
#endif

} // namespace scpp
