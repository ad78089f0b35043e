// scpp_b200/csrc/portable.cuh — lane abstraction shared by the CUDA kernels and the host-simulation build.
//
// The SOCP kernel is written "one warp per problem instance": 32 lanes cooperate on one instance and exchange
// data through shared memory.  The same source is compiled by g++ with LANES == 1 into a TEST-ONLY library
// (tests/_hostsim) so the algorithm can be unit-tested on a machine without a GPU; the product library
// (libscpp_b200.so) contains only the CUDA build and has no CPU execution path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SCPP_HD __host__ __device__ __forceinline__
#define SCPP_D __device__ __forceinline__
#define SCPP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SCPP_HD inline
#define SCPP_D inline
#define SCPP_HD_NOINLINE inline
#endif

namespace scpp {

#if defined(__CUDA_ARCH__)
constexpr int LANES = 32;
SCPP_D int lane_id() { return threadIdx.x & 31; }
SCPP_D void warp_sync() { __syncwarp(); }
SCPP_D double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SCPP_D double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
SCPP_D void warp_sum3(double &a, double &b, double &c)   // three interleaved butterfly reductions (one latency chain)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o), tc = __shfl_xor_sync(0xffffffffu, c, o);
        a += ta; b += tb; c += tc;
    }
}
SCPP_D int warp_or(int v) { return __any_sync(0xffffffffu, v); }
SCPP_D int warp_min_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
SCPP_D double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#else
constexpr int LANES = 1;
inline int lane_id() { return 0; }
inline void warp_sync() {}
inline double warp_sum(double v) { return v; }
inline double warp_max(double v) { return v; }
inline void warp_sum3(double &, double &, double &) {}
inline int warp_or(int v) { return v; }
inline int warp_min_i(int v) { return v; }
inline double warp_bcast(double v, int) { return v; }
#endif

// lanes stride over [0,n)
// code-size experiments (build variants, scpp_b200/build.py --variant): out-of-line Cholesky / stage passes
#if defined(SCPP_NOINLINE_CHOL)
#define SCPP_HD_CHOL SCPP_HD_NOINLINE
#else
#define SCPP_HD_CHOL SCPP_HD
#endif
#if defined(SCPP_NOINLINE_PASS)
#define SCPP_HD_PASS SCPP_HD_NOINLINE
#else
#define SCPP_HD_PASS SCPP_HD
#endif
#define FOR_LANE(i, n) for (int i = lane_id(); i < (n); i += LANES)

// ---- CTA-wide cooperation (the CTA-per-instance solver, cta_ipm.cuh); the host-simulation build is one thread ----
#if defined(__CUDA_ARCH__)
SCPP_D int cta_tid() { return threadIdx.x; }
SCPP_D int cta_threads() { return blockDim.x; }
SCPP_D int cta_warp() { return threadIdx.x >> 5; }
SCPP_D int cta_warps() { return blockDim.x >> 5; }
SCPP_D void cta_sync() { __syncthreads(); }
#else
inline int cta_tid() { return 0; }
inline int cta_threads() { return 1; }
inline int cta_warp() { return 0; }
inline int cta_warps() { return 1; }
inline void cta_sync() {}
#endif
// threads of the CTA stride over [0,n)
#define FOR_CTA(i, n) for (int i = cta_tid(); i < (n); i += cta_threads())

} // namespace scpp
