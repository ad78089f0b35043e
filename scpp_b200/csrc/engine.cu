// scpp_b200/csrc/engine.cu — CUDA kernels, the batched SC engine and the C-ABI (include/scpp_b200.h).
// Built for sm_100a only; there is no CPU execution path in this library.
#include "../../include/scpp_b200.h"
#include "sc.cuh"
#include "lqr.cuh"
#include "info_parser.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace scpp;

static_assert(sizeof(scpp_b200_model_params) == sizeof(ModelParamsHost), "ABI struct mismatch");
static_assert(sizeof(scpp_b200_sc_config) == sizeof(ScConfig), "ABI struct mismatch");
static_assert(SCPP_B200_INFO_STRIDE == INFO_STRIDE, "ABI constant mismatch");

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
int scpp_b200_fail(int code, const std::string &msg) { return fail(code, msg); }      // for the other translation units of the library (mpc.cu)
#define CU(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) return fail(SCPP_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)


// ------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------
template <class M>
__global__ void k_setup(ScArrays<M> a, ModelParamsHost P, ScConfig cfg)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < a.N) sc_setup_instance<M>(a, a.Pn ? a.Pn[n] : P, cfg, n);
}

template <class M>
__global__ void k_warm(ScArrays<M> a, ModelParamsHost P, ScConfig cfg)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < a.N && !(a.frozen && a.frozen[n])) sc_warm_instance<M>(a, a.Pn ? a.Pn[n] : P, cfg, n);
}

// K4: one closed-loop step per instance (thread per instance)
template <class M>
__global__ void k_sim_step(ScArrays<M> a, ModelParamsHost P, ScConfig cfg, double time_step, double *x_out, double *u_out, int *reached)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < a.N) sc_sim_step_instance<M>(a, a.Pn ? a.Pn[n] : P, cfg, n, time_step, x_out ? x_out + (size_t)n * M::NX : nullptr, u_out ? u_out + (size_t)n * M::NU : nullptr,
                                         reached ? reached + n : nullptr);
}
// K4 test hook: plain simulate for n independent states
template <class M>
__global__ void k_simulate(int n, double dt, double *x, const double *u0, const double *u1, const double *par)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rkf78_simulate<M>(x + (size_t)i * M::NX, u0 + (size_t)i * M::NU, u1 + (size_t)i * M::NU, par + (size_t)i * M::NP, dt, 20);
}
// K5: LQR tracking gains, one warp per (instance, node); Jacobians at the redimensionalised solution node with dimensional parameters
template <class M>
__global__ void __launch_bounds__(128) k_lqr(ScArrays<M> a, ModelParamsHost P, const double *qd, const double *rd, double *gains, int *ok)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NX = M::NX, NU = M::NU;
    const int warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * 4 + warp;
    if (gw >= (long long)a.N * a.K) return;
    const int n = int(gw / a.K), k = int(gw - (long long)n * a.K);
    double x[NX], u[NU], xi[NX], xf[NX], par[M::NP], cst[MAX_CST], sc2[2];
    for (int i = 0; i < NX; i++) { x[i] = a.X[((size_t)n * a.K + k) * NX + i]; xi[i] = a.x_init[(size_t)n * NX + i]; xf[i] = a.x_final[(size_t)n * NX + i]; }
    for (int j = 0; j < NU; j++) u[j] = a.U[((size_t)n * a.K + k) * NU + j];
    M::redim(a.scale + 2 * n, x, u);
    M::setup(a.Pn ? a.Pn[n] : P, 0, xi, xf, par, cst, sc2);
    const bool good = Lqr<M>::gain(x, u, par, qd, rd, gains + (size_t)gw * NU * NX, smem + (size_t)warp * Lqr<M>::sm_doubles());
    if ((threadIdx.x & 31) == 0) ok[gw] = good;
}

// SCvx: nonlinear cost of the candidates (thread per (listed instance, interval)) and the ratio test (thread per listed instance)
template <class M>
__global__ void k_scvx_cost(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, int n_list)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = a.K - 1;
    if (idx >= (long long)n_list * per) return;
    const int li = int(idx / per), k = int(idx - (long long)li * per);
    sc_scvx_cost<M>(a, cfg, list[li], k);
}
template <class M>
__global__ void k_scvx_decide(ScArrays<M> a, ScConfig cfg, const int *__restrict__ list, int n_list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_list) sc_scvx_decide<M>(a, cfg, list[i]);
}

// first active list of a solve: every instance that is not frozen
__global__ void k_first_list(int *list, int *count, const int *frozen, int *converged, unsigned char *flags, int first, int n)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= first + n) return;
    if (frozen && frozen[i]) { converged[i] = 8; flags[i] = 8; return; }
    list[atomicAdd(count, 1)] = i;
}

#include "kernels.cuh"

// K1 launch.  jacobian: 1 dual numbers over the flow map (one thread per column), 0 hand-derived Jacobian (same mapping),
// 2 hand-derived Jacobian with the linearisation shared by the columns of an interval (discretize_shared.cuh; models whose Lin record is
// a dense generated matrix do not fit its shared-memory stash and take path 1)
template <class M>
static cudaError_t launch_discretize(const ScArrays<M> &a, int nsub, int jacobian, int free_time /* bit 0: free final time, bit 1: zero-order hold */, const int *list, int n, cudaStream_t stream)
{
    constexpr int NC = M::NX + 2 * M::NU + 2;
    if (jacobian == 2 && !k1s_fits<M>()) jacobian = 1;
    if (jacobian == 2) {
        static bool attr_set[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 64 && !attr_set[dev]) {
            e = cudaFuncSetAttribute(k_discretize_shared<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(k1s_smem_bytes<M>()));
            if (e != cudaSuccess) return e;
            attr_set[dev] = true;
        }
        const long long pairs = (long long)n * (a.K - 1);
        k_discretize_shared<M><<<(unsigned)((pairs + K1S_IPB - 1) / K1S_IPB), K1S_THREADS, k1s_smem_bytes<M>(), stream>>>(a, nsub, (free_time & 2) != 0, list, n);
    } else {
        const long long thr = (long long)n * (a.K - 1) * NC;
        if (jacobian) k_discretize<M, true><<<(unsigned)((thr + 127) / 128), 128, 0, stream>>>(a, nsub, free_time, list, n);
        else k_discretize<M, false><<<(unsigned)((thr + 127) / 128), 128, 0, stream>>>(a, nsub, free_time, list, n);
    }
    return cudaGetLastError();
}


__global__ void k_iota(int *v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }

// after a round: the next active list, the instances among them that start a new sub-problem (to be discretised) and the
// per-instance flag bytes the ranks exchange (0 running, 1 converged, 2 failed, 4 iteration limit reached)
__global__ void k_compact(const int *__restrict__ active, int n_active, const int *__restrict__ converged, const int *__restrict__ iters, int max_it,
                          const double *__restrict__ ipm_state, int state_stride, int *__restrict__ next, int *__restrict__ disc,
                          int *__restrict__ counters, unsigned char *__restrict__ flags, int *__restrict__ cont = nullptr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_active) return;
    const int n = active[i];
    const int c = converged[n];
    const int f = c ? c : (iters[n] >= max_it ? 4 : 0);
    flags[n] = (unsigned char)f;
    if (!f) {
        atomicMin(counters + 2, iters[n]);
        next[atomicAdd(counters, 1)] = n;
        if (ipm_state[(size_t)n * state_stride] == 0.) disc[atomicAdd(counters + 1, 1)] = n;
        else if (cont) cont[atomicAdd(counters + 3, 1)] = n;               // in the middle of a sub-problem (solver 2: tail rounds)
    }
}
// longest-processing-time-first order of the work queue of the CTA-per-instance solver: the instances whose previous sub-problem took the
// most interior-point iterations are handed out first, so the last CTAs to finish a round hold short sub-problems (counting sort by that
// iteration count, descending; one block; the order inside a bucket does not matter: instances are independent)
__global__ void __launch_bounds__(1024) k_lpt_order(const int *__restrict__ in, int *__restrict__ out, int n, const int *__restrict__ iters,
                                                    const double *__restrict__ info, int max_it)
{
    __shared__ int hist[128], base[128];
    for (int b = threadIdx.x; b < 128; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    auto key = [&](int inst) {
        const int it = iters[inst];
        int kx = it > 0 ? (int)info[((size_t)inst * max_it + (it - 1)) * INFO_STRIDE + 5] : 127;
        return kx < 0 ? 0 : (kx > 127 ? 127 : kx);
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[key(in[i])], 1);
    __syncthreads();
    if (threadIdx.x == 0) { int acc = 0; for (int b = 127; b >= 0; b--) { base[b] = acc; acc += hist[b]; } }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const int inst = in[i]; out[atomicAdd(&base[key(inst)], 1)] = inst; }
}

__global__ void k_count_zero(const unsigned char *__restrict__ flags, long long n, unsigned long long *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int z = (i < n) && flags[i] == 0;
    const unsigned b = __ballot_sync(0xffffffffu, z);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}

template <class M>
__global__ void k_export(ScArrays<M> a, double *Xo, double *Uo)
{
    constexpr int NX = M::NX, NU = M::NU;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.N * a.K) return;
    const int n = int(idx / a.K);
    double x[NX], u[NU];
    for (int i = 0; i < NX; i++) x[i] = a.X[idx * NX + i];
    for (int i = 0; i < NU; i++) u[i] = a.U[idx * NU + i];
    M::redim(a.scale + 2 * n, x, u);
    for (int i = 0; i < NX; i++) Xo[idx * NX + i] = x[i];
    for (int i = 0; i < NU; i++) Uo[idx * NU + i] = u[i];
}

// self-test of the tensor-core block products (blockops.cuh) against scalar loops: one warp, the shapes used by the factorisation
__global__ void k_selftest_blockops(double *err)
{
    constexpr int NB = 18, NX = 14, NU = 4, NCP = 26;
    __shared__ double A[NB * NB], B[NB * NB], T[NX * NCP], C1[NB * NB], C2[NB * NB];
    const int lane = threadIdx.x;
    for (int e = lane; e < NB * NB; e += 32) { A[e] = sin(0.37 * e + 0.1); B[e] = cos(0.11 * e) * ((e / NB >= e % NB) ? 1. : 0.); C1[e] = 0.01 * e; C2[e] = 0.01 * e; }
    for (int e = lane; e < NX * NCP; e += 32) T[e] = sin(0.05 * e) + 0.3;
    __syncwarp();
    double worst = 0.;
    // (1) C -= A A'  (lower tiles)   (2) C = A B' with B lower triangular   (3) C += T' (A_x) with a 14-row contraction   (4) 4 x 18 output
    blk::mm<NB, NB, NB, true>([&](int m, int k) { return (m < NB && k < NB) ? A[m * NB + k] : 0.; }, [&](int k, int n) { return (k < NB && n < NB) ? A[n * NB + k] : 0.; },
                              [&](int m, int n, double v) { if (m < NB && n < NB) C1[m * NB + n] -= v; });
    __syncwarp();
    for (int e = lane; e < NB * NB; e += 32) { const int a = e / NB, b = e % NB; if (b / 8 <= a / 8) { double v = 0; for (int c = 0; c < NB; c++) v += A[a * NB + c] * A[b * NB + c]; worst = fmax(worst, fabs(C1[e] - (C2[e] - v))); } }
    __syncwarp();
    blk::mm<NB, NB, NB, false>([&](int m, int k) { return (m < NB && k < NB) ? A[m * NB + k] : 0.; }, [&](int k, int n) { return (k < NB && n < NB) ? B[n * NB + k] : 0.; },
                               [&](int m, int n, double v) { if (m < NB && n < NB) C1[m * NB + n] = v; });
    __syncwarp();
    for (int e = lane; e < NB * NB; e += 32) { const int a = e / NB, b = e % NB; double v = 0; for (int c = 0; c < NB; c++) v += A[a * NB + c] * B[b * NB + c]; worst = fmax(worst, fabs(C1[e] - v)); }
    __syncwarp();
    for (int e = lane; e < NB * NB; e += 32) C1[e] = C2[e];
    __syncwarp();
    blk::mm<NB, NB, NX, true>([&](int m, int k) { return (m < NB && k < NX) ? T[k * NCP + m] : 0.; }, [&](int k, int n) { return (k < NX && n < NB) ? -A[k * NB + n] : 0.; },
                              [&](int m, int n, double v) { if (m < NB && n < NB) C1[m * NB + n] += v; });
    __syncwarp();
    for (int e = lane; e < NB * NB; e += 32) { const int a = e / NB, b = e % NB; if (b / 8 <= a / 8) { double v = 0; for (int i = 0; i < NX; i++) v -= T[i * NCP + a] * A[i * NB + b]; worst = fmax(worst, fabs(C1[e] - (C2[e] + v))); } }
    __syncwarp();
    blk::mm<NU, NB, NX, false>([&](int m, int k) { return (m < NU && k < NX) ? T[k * NCP + NB + m] : 0.; }, [&](int k, int n) { return (k < NX && n < NB) ? -A[k * NB + n] : 0.; },
                               [&](int m, int n, double v) { if (m < NU && n < NB) C1[m * NB + n] = v; });
    __syncwarp();
    for (int e = lane; e < NU * NB; e += 32) { const int a = e / NB, b = e % NB; double v = 0; for (int i = 0; i < NX; i++) v -= T[i * NCP + NB + a] * A[i * NB + b]; worst = fmax(worst, fabs(C1[a * NB + b] - v)); }
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (lane == 0) *err = worst;
}

// ------------------------------------------------------------------------------------------------------------------
// NCCL through dlopen (only needed for multi-GPU runs)
// ------------------------------------------------------------------------------------------------------------------
struct NcclUid { char b[128]; };
struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUid /* ncclUniqueId by value */, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (h) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllGather && CommDestroy;
    }
};
static Nccl g_nccl;

// ------------------------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------------------------
struct scpp_b200_engine {
    virtual ~scpp_b200_engine() {}
    virtual int init() = 0;
    virtual int set_boundary(const double *xi, const double *xf) = 0;
    virtual int solve(int warm) = 0;
    virtual int get_solution(double *X, double *U, double *t, int *it, int *flags) = 0;
    virtual int get_iterate(int it, double *X, double *U, double *t, int redim = 0) = 0;
    virtual int get_info(double *info) = 0;
    virtual int sim_step(double time_step, double *x_new, double *u0, int *reached) = 0;
    virtual int lqr_gains(const double *q_diag, const double *r_diag, double *gains, int *ok) = 0;
    virtual int comm_buffers() = 0;
    virtual int set_instance_params(const scpp_b200_model_params *Pn) = 0;
    int model = 0, N = 0, device = 0;
    ModelParamsHost P;
    ScConfig cfg;
    double ms_disc = 0, ms_socp = 0, ms_total = 0;
    int launches = 0, outer = 0, rounds = 0;
    long long inst_iters = 0, global_active = 0, inst_rounds = 0;
    size_t bytes = 0;
    void *comm = nullptr;
    int nranks = 1, rank = 0;
};

template <class M>
struct EngineT : scpp_b200_engine {
    static constexpr int NX = M::NX, NU = M::NU, NB = NX + NU, NC = NX + 2 * NU + 2;
    ScArrays<M> a;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int *active[2] = {nullptr, nullptr}, *disc = nullptr, *cont = nullptr;
    int *counter = nullptr;               // [0] next active count, [1] instances starting a new sub-problem, [2] fewest outer iterations among the active
    std::vector<int> h_iters;
    std::vector<unsigned char> h_flags;
    unsigned long long *gcount = nullptr;
    unsigned char *flags = nullptr, *flags_all = nullptr;
    double *Xo = nullptr, *Uo = nullptr;
    double *sim_x = nullptr, *sim_u = nullptr;   // outputs of the closed-loop step (K4)
    int *sim_r = nullptr;
    double *lqr_q = nullptr, *lqr_g = nullptr;   // K5 buffers (first use)
    int *lqr_ok = nullptr;
    int *h_counter = nullptr;             // pinned
    unsigned long long *h_gcount = nullptr;
    std::vector<void *> allocs;
    bool have_states = false, solved_once = false;
    int n_sm = 148;
    ModelParamsHost *d_Pn = nullptr;       // per-instance model parameters (optional)
    size_t cta_smem = 0;
    bool cta_ok = false;                                    // the shared-memory image of the CTA solver fits (K small enough)
    int cta_per_sm = 1, *queue = nullptr, *lpt = nullptr;   // CTA-per-instance solver: shared-memory image, residency, device-side work queue
    cudaStream_t cstream = nullptr;        // communication stream (flag exchanges)
    cudaEvent_t ev_flags = nullptr;
    unsigned char *flags_tx = nullptr;
    int exchanges = 0;
    bool exchange_pending = false;
    int N_pad = 0;                         // multi-GPU: flag bytes every rank contributes (max shard size)

    template <class T>
    int dalloc(T **p, size_t n)
    {
        CU(cudaMalloc((void **)p, n * sizeof(T)));
        CU(cudaMemsetAsync(*p, 0, n * sizeof(T), stream));
        allocs.push_back(*p);
        bytes += n * sizeof(T);
        return 0;
    }
    ~EngineT() override
    {
        cudaSetDevice(device);
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        for (void *p : allocs) cudaFree(p);
        if (h_counter) cudaFreeHost(h_counter);
        if (h_gcount) cudaFreeHost(h_gcount);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        if (ev_flags) cudaEventDestroy(ev_flags);
        if (cstream) cudaStreamDestroy(cstream);
        if (stream) cudaStreamDestroy(stream);
    }
    int init() override
    {
        CU(cudaSetDevice(device));
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto &e : ev) CU(cudaEventCreate(&e));
        const int K = cfg.K;
        a.N = N; a.K = K; a.max_it = cfg.max_iterations; a.Pn = nullptr;
        a.ws_stride = Ipm<M>::ws_doubles(K);
        int rc;
#define DA(ptr, n) if ((rc = dalloc(&(ptr), (size_t)(n)))) return rc
        DA(a.x_init, (size_t)N * NX); DA(a.x_final, (size_t)N * NX); DA(a.xi, (size_t)N * NX); DA(a.xf, (size_t)N * NX);
        DA(a.par, (size_t)N * M::NP); DA(a.cst, (size_t)N * MAX_CST); DA(a.scale, (size_t)N * 2);
        DA(a.X, (size_t)N * K * NX); DA(a.U, (size_t)N * K * NU); DA(a.sigma, N);
        DA(a.tdir, (size_t)N * K * 3); DA(a.fixm, (size_t)N * K); DA(a.fixv, (size_t)N * K * NB); DA(a.w_tr, N);
        DA(a.iters, N); DA(a.status, N); DA(a.converged, N);
        DA(a.dd, (size_t)N * (K - 1) * NX * NC);
        DA(a.ddT, (size_t)N * Ipm<M>::ddt_doubles(K));
        DA(a.ws, (size_t)N * a.ws_stride);
        DA(a.info, (size_t)N * cfg.max_iterations * INFO_STRIDE);
        a.hist = nullptr;
        if (cfg.keep_history) DA(a.hist, (size_t)N * (cfg.max_iterations + 1) * a.hist_stride());
        DA(active[0], N); DA(active[1], N); DA(disc, N); DA(cont, N); DA(counter, 4); DA(a.ipm_state, (size_t)N * Ipm<M>::IPM_STATE); DA(gcount, 1); DA(flags, N); DA(a.frozen, N);
        a.trust = a.last_cost = a.n1c = a.Xc = a.Uc = a.costp = nullptr; a.have_last = a.solves = a.phase = nullptr;
        if (cfg.algorithm == 1) {
            DA(a.trust, N); DA(a.last_cost, N); DA(a.n1c, N); DA(a.Xc, (size_t)N * K * NX); DA(a.Uc, (size_t)N * K * NU); DA(a.costp, (size_t)N * K);
            DA(a.have_last, N); DA(a.solves, N); DA(a.phase, N);
        } DA(sim_x, (size_t)N * NX); DA(sim_u, (size_t)N * NU); DA(sim_r, N);
        DA(Xo, (size_t)N * K * NX); DA(Uo, (size_t)N * K * NU);
#undef DA
        CU(cudaMallocHost((void **)&h_counter, 4 * sizeof(int)));
        CU(cudaMallocHost((void **)&h_gcount, sizeof(unsigned long long)));
        CU(cudaFuncSetAttribute(k_solve<M, WPB_MAX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WPB_MAX * Ipm<M>::sm_doubles() * sizeof(double))));
        if (const char *cv = getenv("SCPP_CARVEOUT"))      // experiment knob: shared-memory share of the 256 KB L1 / shared array in percent (the rest is L1)
            CU(cudaFuncSetAttribute(k_solve<M, WPB_MAX, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv)));
        {
            const int wsm = int(WPB_MAX * Ipm<M>::sm_doubles() * sizeof(double));
            CU(cudaFuncSetAttribute(k_sp_warp<M, SP_START, WPB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsm));
            CU(cudaFuncSetAttribute(k_sp_warp<M, SP_FACTOR, WPB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsm));
            CU(cudaFuncSetAttribute(k_sp_warp<M, SP_CHAIN, WPB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsm));
            CU(cudaFuncSetAttribute(k_sp_assemble<M, WPB_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WPB_MAX * Ipm<M>::asm_doubles() * sizeof(double))));
        }
        // CTA-per-instance solver: two CTAs per SM when two shared-memory images fit (K <= ~55), otherwise one
        cta_smem = (size_t)Ipm<M>::cta_sm_doubles(K) * sizeof(double);
        cta_per_sm = (2 * (cta_smem + 1024) <= 227 * 1024) ? 2 : 1;
        if (getenv("SCPP_CTA_PER_SM")) cta_per_sm = atoi(getenv("SCPP_CTA_PER_SM")) == 1 ? 1 : cta_per_sm;      // experiments
        cta_ok = cta_smem <= 227 * 1024;
        if (cfg.solver == 1 || (cfg.solver == 2 && cta_ok)) {
            if (!cta_ok) return fail(SCPP_B200_ERR_UNSUPPORTED, "solver = 1 keeps the block factor in shared memory: K too large (use solver = 0)");
            CU(cudaFuncSetAttribute(k_solve_cta<M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cta_smem)));
            CU(cudaFuncSetAttribute(k_solve_cta<M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cta_smem)));
            int rc2;
            if ((rc2 = dalloc(&queue, 1)) || (rc2 = dalloc(&lpt, N))) return rc2;
        }
        CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
        CU(cudaStreamSynchronize(stream));
        return 0;
    }
    int set_boundary(const double *xi, const double *xf) override
    {
        CU(cudaSetDevice(device));
        CU(cudaMemcpyAsync(a.x_init, xi, (size_t)N * NX * sizeof(double), cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(a.x_final, xf, (size_t)N * NX * sizeof(double), cudaMemcpyHostToDevice, stream));
        CU(cudaMemsetAsync(a.frozen, 0, (size_t)N * sizeof(int), stream));
        CU(cudaStreamSynchronize(stream));
        have_states = true;
        return 0;
    }
    int solve(int warm) override
    {
        if (!have_states) return fail(SCPP_B200_ERR_ARG, "scpp_b200_solve: boundary states not set");
        if (warm && !solved_once) return fail(SCPP_B200_ERR_ARG, "scpp_b200_solve: warm start requested before any solve");
        CU(cudaSetDevice(device));
        const int K = cfg.K, T = 128;
        launches = 0; outer = 0; ms_disc = ms_socp = 0; inst_iters = 0; rounds = 0; inst_rounds = 0;
        CU(cudaEventRecord(ev[0], stream));
        if (warm) k_warm<M><<<(N + T - 1) / T, T, 0, stream>>>(a, P, cfg);
        else { CU(cudaMemsetAsync(a.frozen, 0, (size_t)N * sizeof(int), stream)); k_setup<M><<<(N + T - 1) / T, T, 0, stream>>>(a, P, cfg); }
        CU(cudaMemsetAsync(flags, 0, N, stream));
        launches += 1;
        global_active = (long long)N * nranks;
        exchanges = 0; exchange_pending = false;
        // Large batches are solved in chunks of `chunk` instances, one after the other: 4096 instances already fill the GPU (7 warps per SM x 4
        // waves), and a round over 16 384 instances was measured 25 % slower per instance than over 4096 (round 1: 35.8 k vs 47.5 k
        // instance-iterations/s; the scattered 0.8 MB per-instance regions of a round then span 13 GB).  Chunking keeps the per-round footprint
        // at the 4096-instance size whatever N is.
        const int chunk = chunk_size();
        for (int c0 = 0; c0 < N; c0 += chunk) {
        const int cn = (N - c0 < chunk) ? N - c0 : chunk;
        const bool last_chunk = c0 + cn >= N;
        CU(cudaMemsetAsync(counter, 0, 4 * sizeof(int), stream));
        k_first_list<<<(cn + 255) / 256, 256, 0, stream>>>(active[0], counter, warm ? a.frozen : nullptr, a.converged, flags, c0, cn);   // all but the frozen instances
        launches += 1;
        CU(cudaMemcpyAsync(h_counter, counter, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        int n_active = h_counter[0], n_disc = h_counter[0], n_cont = 0, cur = 0;
        const int *disc_list = active[0];                         // first round: every instance starts its first sub-problem
        // Rounds.  A round (1) discretises the instances that start a new sub-problem (K1), (2) advances EVERY unfinished
        // instance by one slice of cfg.ipm_slice interior-point iterations (K2; K3 runs in its epilogue when a sub-problem is
        // solved), (3) re-forms the lists and exchanges the flag bytes.  With ipm_slice == 0 a slice is a whole sub-problem and
        // the rounds are the reference's outer iterations in lock-step.
        const int slice_eff = cfg.solver == 1 ? 0 : (cfg.ipm_slice < 0 ? 1 : cfg.ipm_slice);      // solver 1: a round is an outer iteration
        const long long max_rounds = (long long)cfg.max_iterations * (cfg.algorithm == 1 ? SCVX_MAX_RESOLVE + 1 : 1) * (slice_eff > 0 ? (cfg.ipm.maxit + 3 + 8) / slice_eff + 2 : 1) + 9;   // + 8: rounds repeated after a regularised re-factorisation
        for (long long round = 0; round < max_rounds && n_active > 0; round++) {
            CU(cudaEventRecord(ev[1], stream));
            if (n_disc > 0) {
                CU(launch_discretize<M>(a, cfg.nsub, cfg.jacobian, (cfg.free_final_time ? 1 : 0) | (cfg.interpolate_input ? 0 : 2), disc_list, n_disc, stream));
                launches++;
            }
            CU(cudaEventRecord(ev[2], stream));
            const int n_active_round = n_active;
            const int parts = (K + 31) / 32;
            // ipm_slice == -1: split pipeline in every round; ipm_slice == -T (T > 1): hybrid, split pipeline only in rounds that advance
            // fewer than T instances (the tail of a solve, where one warp per instance leaves the GPU empty); both paths park the solver
            // in the same state, so they can alternate round by round
            const bool split = cfg.solver == 0 && cfg.ipm_slice < 0 && parts <= 4 && (cfg.ipm_slice == -1 || n_active < -cfg.ipm_slice);
            if (cfg.solver == 1) {
                if (n_active > 0) {
                    CU(cudaMemsetAsync(queue, 0, sizeof(int), stream));
                    const int grid = n_active < n_sm * cta_per_sm ? n_active : n_sm * cta_per_sm;
                    const int *order = active[cur];
                    if (n_active > grid && round > 0) {       // more instances than resident CTAs: hand out the long sub-problems first
                        k_lpt_order<<<1, 1024, 0, stream>>>(active[cur], lpt, n_active, a.iters, a.info, cfg.max_iterations);
                        order = lpt; launches++;
                    }
                    if (cta_per_sm == 2) k_solve_cta<M, 2><<<grid, cta_threads_for(2), cta_smem, stream>>>(a, cfg, order, n_active, queue);
                    else k_solve_cta<M, 1><<<grid, cta_threads_for(1), cta_smem, stream>>>(a, cfg, order, n_active, queue);
                    launches++;
                }
            } else if (split) {
                // split pipeline: one interior-point iteration of every unfinished instance as a sequence of kernels (sc.cuh)
                auto warp_launch = [&](auto kern, const int *list, int n, int mode) {
                    int wpb = (n + n_sm - 1) / n_sm;
                    if (wpb < 1) wpb = 1;
                    if (wpb > WPB_MAX) wpb = WPB_MAX;
                    kern<<<(n + wpb - 1) / wpb, wpb * 32, (size_t)wpb * Ipm<M>::sm_doubles() * sizeof(double), stream>>>(a, cfg, list, nullptr, n, mode);
                    launches++;
                };
                if (n_disc > 0) warp_launch(k_sp_warp<M, SP_START, WPB_MAX>, disc_list, n_disc, 0);
                if (n_active > 0) {
                    const int *lst = active[cur];
                    const long long aw = (long long)n_active * K;
                    k_sp_assemble<M, WPB_MAX><<<(unsigned)((aw + WPB_MAX - 1) / WPB_MAX), WPB_MAX * 32,
                                                (size_t)WPB_MAX * Ipm<M>::asm_doubles() * sizeof(double), stream>>>(a, cfg, lst, n_active);
                    launches++;
                    warp_launch(k_sp_warp<M, SP_FACTOR, WPB_MAX>, lst, n_active, 0);
                    const int ipb = parts >= 4 ? 1 : 4 / parts, sthreads = 32 * parts * ipb, sgrid = (n_active + ipb - 1) / ipb;
                    for (int mode = 1; mode <= 2; mode++) {
                        k_sp_stage<M, SP_RHS><<<sgrid, sthreads, 0, stream>>>(a, cfg, lst, n_active, parts, ipb, mode);
                        warp_launch(k_sp_warp<M, SP_CHAIN, WPB_MAX>, lst, n_active, mode);
                        k_sp_stage<M, SP_RECOVER><<<sgrid, sthreads, 0, stream>>>(a, cfg, lst, n_active, parts, ipb, mode);
                        launches += 2;
                    }
                    k_sp_stage<M, SP_UPDATE><<<sgrid, sthreads, 0, stream>>>(a, cfg, lst, n_active, parts, ipb, 0);
                    k_sp_test<M><<<(n_active + 3) / 4, 128, 0, stream>>>(a, cfg, lst, n_active);
                    launches += 2;
                }
            } else if (cfg.solver == 2 && cta_ok && n_active > 0 && n_active <= n_sm * cta_per_sm) {
                // solver 2, TAIL of a solve: fewer unfinished instances than CTAs the GPU holds.  A round of the warp solver lasts as long as one
                // warp needs for one interior-point iteration however few instances it advances, and the last 15 % of the rounds of a 1024-batch
                // advance a few dozen stragglers.  Here a sub-problem that STARTS in such a round runs on the CTA-per-instance solver (8 warps per
                // instance, whole sub-problem in this launch); instances in the middle of a sub-problem finish it on the warp solver.  The two
                // mappings sum in different orders: results agree to solver accuracy, not bit for bit, and which mapping an instance sees depends
                // on the batch -- hence a knob (cfg.solver = 2), not the default.
                if (n_disc > 0) {
                    CU(cudaMemsetAsync(queue, 0, sizeof(int), stream));
                    const int grid = n_disc < n_sm * cta_per_sm ? n_disc : n_sm * cta_per_sm;
                    if (cta_per_sm == 2) k_solve_cta<M, 2><<<grid, cta_threads_for(2), cta_smem, stream>>>(a, cfg, disc_list, n_disc, queue);
                    else k_solve_cta<M, 1><<<grid, cta_threads_for(1), cta_smem, stream>>>(a, cfg, disc_list, n_disc, queue);
                    launches++;
                }
                if (n_cont > 0) {
                    int wpb = (n_cont + n_sm - 1) / n_sm;
                    if (wpb > WPB_MAX) wpb = WPB_MAX;
                    k_solve<M, WPB_MAX, 1><<<(n_cont + wpb - 1) / wpb, wpb * 32, (size_t)wpb * Ipm<M>::sm_doubles() * sizeof(double), stream>>>(a, cfg, cont, n_cont);
                    launches++;
                }
            } else
            if (n_active > 0) {
                // one warp per instance.  Small batches: spread the warps evenly, one CTA per SM (a batch of 1024 on 148 SMs is
                // 7 warps per SM).  Large batches: 7-warp CTAs, one resident per SM (measured at 4096 instances: 47.7 k
                // instance-iterations/s against 43.3 k for 4-warp CTAs two per SM, whose 3.46 waves leave the last one half empty;
                // 39.5 k / 43.3 k with 5 / 6 warps per CTA; a persistent work-queue variant was slower, 42.5 k)
                int wpb = (n_active + n_sm - 1) / n_sm;
                if (wpb < 1) wpb = 1;
                if (wpb > WPB_MAX) wpb = WPB_MAX;
                // experiment knob SCPP_WPB_BIG: warps per CTA once the instances exceed one wave (fewer warps leave more of the 256 KB array to the L1)
                if (n_active > n_sm * WPB_MAX) { static const int big = getenv("SCPP_WPB_BIG") ? atoi(getenv("SCPP_WPB_BIG")) : WPB_MAX; if (big >= 1 && big < wpb) wpb = big; }
                const size_t smem = (size_t)wpb * Ipm<M>::sm_doubles() * sizeof(double);
                k_solve<M, WPB_MAX, 1><<<(n_active + wpb - 1) / wpb, wpb * 32, smem, stream>>>(a, cfg, active[cur], n_active);
                launches++;
            }
            if (cfg.algorithm == 1 && n_active > 0) {      // SCvx: simulate the candidates of the sub-problems that ended in this round, ratio test
                const long long thr = (long long)n_active * (K - 1);
                k_scvx_cost<M><<<(unsigned)((thr + 127) / 128), 128, 0, stream>>>(a, cfg, active[cur], n_active);
                k_scvx_decide<M><<<(n_active + 127) / 128, 128, 0, stream>>>(a, cfg, active[cur], n_active);
                launches += 2;
            }
            CU(cudaEventRecord(ev[3], stream));
            CU(cudaMemsetAsync(counter, 0, 2 * sizeof(int), stream));
            CU(cudaMemsetAsync(counter + 2, 0x7f, sizeof(int), stream));
            CU(cudaMemsetAsync(counter + 3, 0, sizeof(int), stream));
            if (n_active > 0) {
                k_compact<<<(n_active + 255) / 256, 256, 0, stream>>>(active[cur], n_active, a.converged, a.iters, cfg.max_iterations, a.ipm_state,
                                                                     Ipm<M>::IPM_STATE, active[cur ^ 1], disc, counter, flags, cont);
                launches++;
            }
            CU(cudaMemcpyAsync(h_counter, counter, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            // The one data-path collective: ncclAllGather of the per-instance flag bytes, ONE PER OUTER ITERATION, and never waited for inside
            // the loop.  Instances are independent, so a rank stops on the completion of its own shard; the exchange only reports the global
            // state.  Every rank issues exactly max_iterations exchanges per solve (a collective needs matching calls): the i-th is enqueued
            // on the communication stream once every running instance of this rank has passed outer iteration i; what is left when the rank
            // runs out of work is enqueued after the loop.
            if (comm) {
                CU(cudaEventRecord(ev_flags, stream));
                exchange_pending = true;
            }
            CU(cudaStreamSynchronize(stream));
            float m1 = 0, m2 = 0;
            CU(cudaEventElapsedTime(&m1, ev[1], ev[2]));
            CU(cudaEventElapsedTime(&m2, ev[2], ev[3]));
            ms_disc += m1; ms_socp += m2;
            n_active = h_counter[0]; n_disc = h_counter[1]; n_cont = h_counter[3];
            if (comm && last_chunk) {
                const int passed = n_active > 0 ? h_counter[2] : cfg.max_iterations;
                while (exchanges < passed && exchanges < cfg.max_iterations) { int rc = exchange_flags(); if (rc) return rc; }
            }
            disc_list = disc;
            cur ^= 1;
            rounds++; inst_rounds += n_active_round;
            global_active = (long long)n_active;           // this rank's shard; the global count follows after the loop
        }
        }   // chunks
        if (comm) {
            while (exchanges < cfg.max_iterations) { int rc = exchange_flags(); if (rc) return rc; }
            CU(cudaMemsetAsync(gcount, 0, sizeof(unsigned long long), cstream));
            const long long tot = (long long)N_pad * nranks;
            k_count_zero<<<(unsigned)((tot + 255) / 256), 256, 0, cstream>>>(flags_all, tot, gcount);
            CU(cudaMemcpyAsync(h_gcount, gcount, sizeof(unsigned long long), cudaMemcpyDeviceToHost, cstream));
            CU(cudaStreamSynchronize(cstream));
            global_active = (long long)*h_gcount;
            launches++;
        }
        // SC iterations done: per instance (reported as instance-iterations) and the largest count (outer iterations)
        h_iters.resize(N); h_flags.resize(N);
        CU(cudaMemcpyAsync(h_iters.data(), a.iters, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(h_flags.data(), flags, (size_t)N, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        for (int n = 0; n < N; n++) {
            if (h_flags[n] == 8) continue;                          // frozen in the closed loop: not solved, a.iters is the count of an earlier solve
            inst_iters += h_iters[n]; if (h_iters[n] > outer) outer = h_iters[n];
        }
        CU(cudaEventRecord(ev[4], stream));
        CU(cudaStreamSynchronize(stream));
        float mt = 0;
        CU(cudaEventElapsedTime(&mt, ev[0], ev[4]));
        ms_total = mt;
        CU(cudaGetLastError());
        solved_once = true;
        return 0;
    }
    int get_solution(double *X, double *U, double *t, int *it, int *flg) override
    {
        CU(cudaSetDevice(device));
        const int K = cfg.K;
        const long long tot = (long long)N * K;
        k_export<M><<<(unsigned)((tot + 127) / 128), 128, 0, stream>>>(a, Xo, Uo);
        if (X) CU(cudaMemcpyAsync(X, Xo, (size_t)N * K * NX * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (U) CU(cudaMemcpyAsync(U, Uo, (size_t)N * K * NU * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (t) CU(cudaMemcpyAsync(t, a.sigma, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (it) CU(cudaMemcpyAsync(it, a.iters, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (flg) CU(cudaMemcpyAsync(flg, a.converged, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        return 0;
    }
    int get_iterate(int it, double *X, double *U, double *t, int redim = 0) override
    {
        if (!a.hist) return fail(SCPP_B200_ERR_ARG, "scpp_b200_get_iterate: engine created without keep_history");
        if (it < 0 || it > cfg.max_iterations) return fail(SCPP_B200_ERR_ARG, "scpp_b200_get_iterate: iteration out of range");
        CU(cudaSetDevice(device));
        const int K = cfg.K;
        const size_t hs = a.hist_stride();
        std::vector<double> buf((size_t)N * hs);
        CU(cudaMemcpy2DAsync(buf.data(), hs * sizeof(double), a.hist + (size_t)it * hs, (size_t)(cfg.max_iterations + 1) * hs * sizeof(double),
                             hs * sizeof(double), N, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        std::vector<double> sc;
        if (redim) {      // SCAlgorithm::getAllSolutions redimensionalises every iterate (SCAlgorithm.cpp:217-232, model->redimensionalizeTrajectory)
            sc.resize((size_t)N * 2);
            CU(cudaMemcpy(sc.data(), a.scale, sc.size() * sizeof(double), cudaMemcpyDeviceToHost));
        }
        for (int n = 0; n < N; n++) {
            const double *h = buf.data() + (size_t)n * hs;
            for (int k = 0; k < K; k++) {
                double x[NX], u[NU];
                for (int i = 0; i < NX; i++) x[i] = h[k * NB + i];
                for (int i = 0; i < NU; i++) u[i] = h[k * NB + NX + i];
                if (redim) M::redim(sc.data() + 2 * n, x, u);
                if (X) for (int i = 0; i < NX; i++) X[((size_t)n * K + k) * NX + i] = x[i];
                if (U) for (int i = 0; i < NU; i++) U[((size_t)n * K + k) * NU + i] = u[i];
            }
            if (t) t[n] = h[K * NB];
        }
        return 0;
    }
    // one step of the closed loop of scpp/src/SC_sim.cpp for every instance: K4 advances x_init on the device; the next solve(warm) uses it
    int sim_step(double time_step, double *x_new, double *u0, int *reached) override
    {
        if (!solved_once) return fail(SCPP_B200_ERR_ARG, "scpp_b200_sim_step: no solution yet");
        if (!(time_step > 0.)) return fail(SCPP_B200_ERR_ARG, "scpp_b200_sim_step: time_step must be positive");
        CU(cudaSetDevice(device));
        k_sim_step<M><<<(N + 63) / 64, 64, 0, stream>>>(a, P, cfg, time_step, sim_x, sim_u, sim_r);
        if (x_new) CU(cudaMemcpyAsync(x_new, sim_x, (size_t)N * NX * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (u0) CU(cudaMemcpyAsync(u0, sim_u, (size_t)N * NU * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (reached) CU(cudaMemcpyAsync(reached, sim_r, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        return 0;
    }
    // LQRTracker::LQRTracker (scpp_core/src/LQRTracker.cpp:6-28) for every instance: one gain per node of the current solution
    int lqr_gains(const double *q_diag, const double *r_diag, double *gains, int *ok) override
    {
        if (!solved_once) return fail(SCPP_B200_ERR_ARG, "scpp_b200_lqr_gains: no solution yet");
        if (!q_diag || !r_diag || !gains) return fail(SCPP_B200_ERR_ARG, "scpp_b200_lqr_gains: null argument");
        CU(cudaSetDevice(device));
        const int K = cfg.K;
        const size_t ng = (size_t)N * K * NU * NX;
        if (!lqr_g) {                                            // engine-owned buffers, allocated on first use, freed with the engine
            int rc;
            if ((rc = dalloc(&lqr_q, (size_t)(NX + NU))) || (rc = dalloc(&lqr_g, ng)) || (rc = dalloc(&lqr_ok, (size_t)N * K))) return rc;
        }
        double *dq = lqr_q, *dg = lqr_g; int *dok = lqr_ok;
        CU(cudaMemsetAsync(dg, 0xff, sizeof(double) * ng, stream));       // NaN: a gain Lqr::gain does not write (sign iteration not converged, ok = 0) is never garbage
        CU(cudaMemcpyAsync(dq, q_diag, sizeof(double) * NX, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(dq + NX, r_diag, sizeof(double) * NU, cudaMemcpyHostToDevice, stream));
        const int smem = int(4 * Lqr<M>::sm_doubles() * sizeof(double));
        CU(cudaFuncSetAttribute(k_lqr<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const long long warps = (long long)N * K;
        k_lqr<M><<<(unsigned)((warps + 3) / 4), 128, smem, stream>>>(a, P, dq, dq + NX, dg, dok);
        CU(cudaMemcpyAsync(gains, dg, sizeof(double) * ng, cudaMemcpyDeviceToHost, stream));
        if (ok) CU(cudaMemcpyAsync(ok, dok, sizeof(int) * N * K, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        CU(cudaGetLastError());
        return 0;
    }
    // flag exchange buffers.  Shards may have different sizes (N_total % nranks != 0): every rank sends N_pad = max over ranks of N bytes,
    // the bytes beyond its own N are a constant non-zero 'done' flag, so ncclAllGather sees equal counts and the padding never counts as active.
    int comm_buffers() override
    {
        int *dn = nullptr;
        std::vector<int> hn(nranks, 0);
        CU(cudaMalloc((void **)&dn, sizeof(int) * (nranks + 1)));
        CU(cudaMemcpyAsync(dn + nranks, &N, sizeof(int), cudaMemcpyHostToDevice, stream));
        int rc = g_nccl.AllGather(dn + nranks, dn, 1, /*ncclInt32*/ 2, comm, stream);
        if (rc != 0) { cudaFree(dn); return fail(SCPP_B200_ERR_NCCL, std::string("ncclAllGather(N): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?")); }
        CU(cudaMemcpyAsync(hn.data(), dn, sizeof(int) * nranks, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        cudaFree(dn);
        N_pad = N;
        for (int r = 0; r < nranks; r++) if (hn[r] > N_pad) N_pad = hn[r];
        unsigned char *fp = nullptr, *fa = nullptr;
        CU(cudaMalloc((void **)&fp, (size_t)N_pad)); allocs.push_back(fp);
        CU(cudaMalloc((void **)&fa, (size_t)N_pad * nranks)); allocs.push_back(fa);
        CU(cudaMemsetAsync(fp, 1, (size_t)N_pad, stream));
        CU(cudaStreamSynchronize(stream));
        flags = fp; flags_all = fa;
        CU(cudaMalloc((void **)&flags_tx, (size_t)N_pad)); allocs.push_back(flags_tx);
        CU(cudaEventCreateWithFlags(&ev_flags, cudaEventDisableTiming));
        return 0;
    }
    // one flag exchange on the communication stream, ordered after the last flag update of the compute stream; not waited for
    int exchange_flags()
    {
        if (!cstream) { CU(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking)); }
        if (exchange_pending) CU(cudaStreamWaitEvent(cstream, ev_flags, 0));
        CU(cudaMemcpyAsync(flags_tx, flags, (size_t)N_pad, cudaMemcpyDeviceToDevice, cstream));      // snapshot: the compute stream keeps updating flags
        int rc = g_nccl.AllGather(flags_tx, flags_all, (size_t)N_pad, /*ncclUint8*/ 1, comm, cstream);
        if (rc != 0) return fail(SCPP_B200_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
        exchanges++;
        return 0;
    }
    // RocketQuat::Parameters per instance (rocketQuat.hpp:50-85): a Monte-Carlo batch may vary the vehicle, not only the boundary states
    int set_instance_params(const scpp_b200_model_params *Pn) override
    {
        CU(cudaSetDevice(device));
        if (!Pn) { a.Pn = nullptr; return 0; }
        for (int n = 0; n < N; n++) if ((Pn[n].enable_roll_control != 0) != (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL)) return fail(SCPP_B200_ERR_UNSUPPORTED, "enable_roll_control: use model SCPP_B200_MODEL_ROCKETQUAT_ROLL for true, SCPP_B200_MODEL_ROCKETQUAT for false");
        if (!d_Pn) { int rc; if ((rc = dalloc(&d_Pn, (size_t)N))) return rc; }
        CU(cudaMemcpyAsync(d_Pn, Pn, (size_t)N * sizeof(ModelParamsHost), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        a.Pn = d_Pn;
        return 0;
    }
    static int chunk_size()
    {
        const char *e = getenv("SCPP_CHUNK");
        const int c = e ? atoi(e) : 4096;
        return c > 0 ? c : 4096;
    }
    int get_info(double *info) override
    {
        CU(cudaSetDevice(device));
        CU(cudaMemcpy(info, a.info, (size_t)N * cfg.max_iterations * INFO_STRIDE * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    }
};

template <class M>
static int simulate_hook(int n, double dt, int device, double *x, const double *u0, const double *u1, const double *par)
{
    CU(cudaSetDevice(device));
    double *dx = nullptr, *d0 = nullptr, *d1 = nullptr, *dp = nullptr;
    CU(cudaMalloc((void **)&dx, sizeof(double) * n * M::NX)); CU(cudaMalloc((void **)&d0, sizeof(double) * n * M::NU));
    CU(cudaMalloc((void **)&d1, sizeof(double) * n * M::NU)); CU(cudaMalloc((void **)&dp, sizeof(double) * n * M::NP));
    CU(cudaMemcpy(dx, x, sizeof(double) * n * M::NX, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d0, u0, sizeof(double) * n * M::NU, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d1, u1, sizeof(double) * n * M::NU, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dp, par, sizeof(double) * n * M::NP, cudaMemcpyHostToDevice));
    k_simulate<M><<<(n + 63) / 64, 64>>>(n, dt, dx, d0, d1, dp);
    CU(cudaMemcpy(x, dx, sizeof(double) * n * M::NX, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(d0); cudaFree(d1); cudaFree(dp);
    CU(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" {

int scpp_b200_version(void) { return 100; }
const char *scpp_b200_last_error(void) { return g_err.c_str(); }
int scpp_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int scpp_b200_model_dims(int model, int *nx, int *nu, int *np)
{
    if (model == SCPP_B200_MODEL_ROCKETQUAT || model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) { *nx = RocketQuat::NX; *nu = RocketQuat::NU; *np = RocketQuat::NP; return 0; }
    if (model == SCPP_B200_MODEL_ROCKET2D || model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) { *nx = Rocket2d::NX; *nu = Rocket2d::NU; *np = Rocket2d::NP; return 0; }
    return fail(SCPP_B200_ERR_ARG, "unknown model");
}
extern "C++" {
template <class M>
static int model_rows_t(const ModelParamsHost &P, const double *x_init, const double *x_final, int max_rows, double *rows, int *n_lp, int *n_cones, int *cone_dims)
{
    if (M::NLP + M::NCR > max_rows) return fail(SCPP_B200_ERR_ARG, "scpp_b200_model_rows: max_rows too small");
    double xi[M::NX], xf[M::NX], par[M::NP], cst[MAX_CST], sc2[2];
    for (int i = 0; i < M::NX; i++) { xi[i] = x_init[i]; xf[i] = x_final[i]; }
    M::setup(P, 0, xi, xf, par, cst, sc2);
    for (int r = 0; r < M::NLP + M::NCR; r++) {
        const RowDesc rd = M::row(r);
        double *o = rows + 8 * r;
        o[0] = rd.n;
        for (int q = 0; q < 3; q++) { o[1 + q] = q < rd.n ? rd.idx[q] : -1; o[4 + q] = q < rd.n ? (rd.cs[q] >= 0 ? cst[rd.cs[q]] : NAN) : 0.; }
        o[7] = cst[rd.hs];
    }
    *n_lp = M::NLP; *n_cones = M::NCONE;
    for (int c = 0; c < M::NCONE; c++) cone_dims[c] = M::cone_dim(c);
    return 0;
}
}   // extern "C++"
int scpp_b200_model_rows(int model, const scpp_b200_model_params *params, const double *x_init, const double *x_final, int max_rows,
                         double *rows, int *n_lp, int *n_cones, int *cone_dims)
{
    if (!params || !x_init || !x_final || !rows || !n_lp || !n_cones || !cone_dims) return fail(SCPP_B200_ERR_ARG, "scpp_b200_model_rows: bad argument");
    ModelParamsHost P;
    memcpy(&P, params, sizeof(P));
    if (model == SCPP_B200_MODEL_ROCKETQUAT) return model_rows_t<RocketQuat>(P, x_init, x_final, max_rows, rows, n_lp, n_cones, cone_dims);
    if (model == SCPP_B200_MODEL_ROCKET2D) return model_rows_t<Rocket2d>(P, x_init, x_final, max_rows, rows, n_lp, n_cones, cone_dims);
    if (model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) return model_rows_t<Rocket2dPlugin>(P, x_init, x_final, max_rows, rows, n_lp, n_cones, cone_dims);
    if (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) return model_rows_t<RocketQuatRollPlugin>(P, x_init, x_final, max_rows, rows, n_lp, n_cones, cone_dims);
    return fail(SCPP_B200_ERR_ARG, "unknown model");
}
// scpp_models/config/RocketQuat/SC.info:1-16, scpp_models/config/Rocket2D/SC.info:1-16
void scpp_b200_default_config(int model, scpp_b200_sc_config *c)
{
    memset(c, 0, sizeof(*c));
    c->free_final_time = 1; c->interpolate_input = 1; c->nondimensionalize = 1;
    const bool rq = model == SCPP_B200_MODEL_ROCKETQUAT || model == SCPP_B200_MODEL_ROCKETQUAT_ROLL;
    c->K = rq ? 15 : 25;
    c->weight_time = 1.; c->weight_trust_region_time = 1.;
    c->weight_trust_region_trajectory = rq ? 50. : 1.;
    c->weight_virtual_control = 1000.;
    c->nu_tol = 1e-5; c->delta_tol = 1e-3; c->max_iterations = 15;
    c->nsub = -5; c->keep_history = 0; c->ipm_slice = 1;   // RK4 x 5 and x 10, Richardson-extrapolated: the accuracy of RK4 x 20 for 3/4 of the work
    c->ipm.feastol = 1e-8; c->ipm.abstol = 1e-8; c->ipm.reltol = 1e-8; c->ipm.maxit = 100;
    c->algorithm = 0;   // SCvx.info values, used when algorithm is set to 1
    c->scvx_rho_0 = 0.; c->scvx_rho_1 = 0.25; c->scvx_rho_2 = 0.9; c->scvx_alpha = 2.; c->scvx_beta = 3.2;
    c->scvx_change_threshold = rq ? 1e-3 : 1e-2; c->scvx_trust_region = 5.;
    c->jacobian = 1;
}

static void deg2rad(double &v) { v *= M_PI / 180.; }
static void quat_mul(const double *a, const double *b, double *o)
{
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    o[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}
// eulerToQuaternionXYZ, scpp_models/include/common.hpp:29-38
static void euler_xyz(const double *e, double *q)
{
    const double qx[4] = {cos(e[0] / 2), sin(e[0] / 2), 0, 0}, qy[4] = {cos(e[1] / 2), 0, sin(e[1] / 2), 0}, qz[4] = {cos(e[2] / 2), 0, 0, sin(e[2] / 2)};
    double t[4];
    quat_mul(qx, qy, t); quat_mul(t, qz, q);
}

int scpp_b200_load_model_info(const char *path, int model, scpp_b200_model_params *p, double *x_init, double *x_final)
{
    try {
        ParameterServer ps(path);
        memset(p, 0, sizeof(*p));
        if (model == SCPP_B200_MODEL_ROCKETQUAT || model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) {   // rocketQuat.cpp:234-289
            bool random_initial_state, exact, roll;
            double I_sp, m_init, m_dry, r_init[3], v_init[3], rpy_init[3], w_init[3], r_final[3], v_final[3], rpy_final[3], w_final[3];
            ps.loadMatrix("g_I", p->g_I, 3); ps.loadMatrix("J_B", p->J_B, 3); ps.loadMatrix("r_T_B", p->r_T_B, 3);
            ps.loadScalar("m_init", m_init); ps.loadMatrix("r_init", r_init, 3); ps.loadMatrix("v_init", v_init, 3);
            ps.loadMatrix("rpy_init", rpy_init, 3); ps.loadMatrix("w_init", w_init, 3); ps.loadMatrix("w_final", w_final, 3);
            ps.loadScalar("m_dry", m_dry); ps.loadMatrix("r_final", r_final, 3); ps.loadMatrix("v_final", v_final, 3);
            ps.loadMatrix("rpy_final", rpy_final, 3);
            ps.loadScalar("T_min", p->T_min); ps.loadScalar("T_max", p->T_max); ps.loadScalar("t_max", p->t_max);
            ps.loadScalar("I_sp", I_sp);
            ps.loadScalar("gimbal_max", p->gimbal_max); ps.loadScalar("theta_max", p->theta_max);
            ps.loadScalar("gamma_gs", p->gamma_gs); ps.loadScalar("w_B_max", p->w_B_max);
            ps.loadScalar("random_initial_state", random_initial_state);
            ps.loadScalar("final_time", p->final_time);
            ps.loadScalar("exact_minimum_thrust", exact); ps.loadScalar("enable_roll_control", roll);
            deg2rad(p->gimbal_max); deg2rad(p->theta_max); deg2rad(p->gamma_gs); deg2rad(p->w_B_max);
            for (int i = 0; i < 3; i++) { deg2rad(rpy_init[i]); deg2rad(rpy_final[i]); deg2rad(w_init[i]); deg2rad(w_final[i]); }
            p->alpha_m = 1. / (I_sp * fabs(p->g_I[2]));
            p->exact_minimum_thrust = exact; p->enable_roll_control = roll;
            double q0[4], q1[4];
            euler_xyz(rpy_init, q0); euler_xyz(rpy_final, q1);
            x_init[0] = m_init; x_final[0] = m_dry;
            for (int i = 0; i < 3; i++) { x_init[1 + i] = r_init[i]; x_init[4 + i] = v_init[i]; x_init[11 + i] = w_init[i];
                                          x_final[1 + i] = r_final[i]; x_final[4 + i] = v_final[i]; x_final[11 + i] = w_final[i]; }
            for (int i = 0; i < 4; i++) { x_init[7 + i] = q0[i]; x_final[7 + i] = q1[i]; }
            // random_initial_state: randomizeInitialState() is commented out in the reference (rocketQuat.cpp:203-227); batches are
            // perturbed by the caller instead
            (void)random_initial_state;
            if (roll != (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL)) return fail(SCPP_B200_ERR_UNSUPPORTED, "enable_roll_control: use model SCPP_B200_MODEL_ROCKETQUAT_ROLL for true, SCPP_B200_MODEL_ROCKETQUAT for false");
        } else if (model == SCPP_B200_MODEL_ROCKET2D || model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) {   // rocket2d.cpp:150-196
            bool cif, slack;
            double r_init[2], v_init[2], r_final[2], v_final[2], w_init, w_final, eta_init, eta_final;
            ps.loadMatrix("g_I", p->g_I, 2); ps.loadScalar("J_B", p->J_B[0]); ps.loadMatrix("r_T_B", p->r_T_B, 2);
            ps.loadMatrix("r_init", r_init, 2); ps.loadMatrix("v_init", v_init, 2); ps.loadScalar("eta_init", eta_init); ps.loadScalar("w_init", w_init);
            ps.loadMatrix("r_final", r_final, 2); ps.loadMatrix("v_final", v_final, 2); ps.loadScalar("eta_final", eta_final); ps.loadScalar("w_final", w_final);
            ps.loadScalar("final_time", p->final_time);
            ps.loadScalar("m", p->m); ps.loadScalar("T_min", p->T_min); ps.loadScalar("T_max", p->T_max);
            ps.loadScalar("gamma_gs", p->gamma_gs); ps.loadScalar("gimbal_max", p->gimbal_max);
            ps.loadScalar("theta_max", p->theta_max); ps.loadScalar("w_B_max", p->w_B_max);
            ps.loadScalar("constrain_initial_final", cif); ps.loadScalar("add_slack_variables", slack);
            deg2rad(p->gimbal_max); deg2rad(p->theta_max); deg2rad(p->gamma_gs); deg2rad(p->w_B_max);
            deg2rad(w_init); deg2rad(w_final); deg2rad(eta_init); deg2rad(eta_final);
            p->constrain_initial_final = cif;
            x_init[0] = r_init[0]; x_init[1] = r_init[1]; x_init[2] = v_init[0]; x_init[3] = v_init[1]; x_init[4] = eta_init; x_init[5] = w_init;
            x_final[0] = r_final[0]; x_final[1] = r_final[1]; x_final[2] = v_final[0]; x_final[3] = v_final[1]; x_final[4] = eta_final; x_final[5] = w_final;
        } else return fail(SCPP_B200_ERR_ARG, "unknown model");
    } catch (const std::exception &ex) { return fail(SCPP_B200_ERR_IO, ex.what()); }
    return 0;
}

int scpp_b200_load_sc_info(const char *path, scpp_b200_sc_config *c)
{
    try {   // SCAlgorithm::loadParameters, SCAlgorithm.cpp:22-46
        ParameterServer ps(path);
        bool fft, nd, ii;
        ps.loadScalar("K", c->K);
        ps.loadScalar("free_final_time", fft);
        ps.loadScalar("nondimensionalize", nd);
        ps.loadScalar("delta_tol", c->delta_tol); ps.loadScalar("max_iterations", c->max_iterations); ps.loadScalar("nu_tol", c->nu_tol);
        ps.loadScalar("weight_time", c->weight_time); ps.loadScalar("weight_virtual_control", c->weight_virtual_control);
        ps.loadScalar("weight_trust_region_trajectory", c->weight_trust_region_trajectory);
        ps.loadScalar("interpolate_input", ii);
        if (fft) ps.loadScalar("weight_trust_region_time", c->weight_trust_region_time);
        c->free_final_time = fft; c->nondimensionalize = nd; c->interpolate_input = ii;
    } catch (const std::exception &ex) { return fail(SCPP_B200_ERR_IO, ex.what()); }
    return 0;
}

int scpp_b200_load_scvx_info(const char *path, scpp_b200_sc_config *c)
{
    try {   // SCvxAlgorithm::loadParameters, SCvxAlgorithm.cpp:23-44
        ParameterServer ps(path);
        bool nd, ii;
        ps.loadScalar("K", c->K);
        ps.loadScalar("nondimensionalize", nd);
        ps.loadScalar("max_iterations", c->max_iterations);
        ps.loadScalar("alpha", c->scvx_alpha); ps.loadScalar("beta", c->scvx_beta);
        ps.loadScalar("rho_0", c->scvx_rho_0); ps.loadScalar("rho_1", c->scvx_rho_1); ps.loadScalar("rho_2", c->scvx_rho_2);
        ps.loadScalar("change_threshold", c->scvx_change_threshold);
        ps.loadScalar("weight_virtual_control", c->weight_virtual_control);
        ps.loadScalar("trust_region", c->scvx_trust_region);
        ps.loadScalar("interpolate_input", ii);
        c->nondimensionalize = nd; c->interpolate_input = ii; c->algorithm = 1;
        c->free_final_time = 1;   // engine-internal: the time column exists and sigma is pinned to final_time (SCvx has a fixed final time)
    } catch (const std::exception &ex) { return fail(SCPP_B200_ERR_IO, ex.what()); }
    return 0;
}

int scpp_b200_create(int model, const scpp_b200_model_params *params, const scpp_b200_sc_config *cfg, int n, int device, scpp_b200_engine **out)
{
    if (cfg && cfg->algorithm != 0 && cfg->algorithm != 1) return fail(SCPP_B200_ERR_ARG, "scpp_b200_create: algorithm must be 0 (SC) or 1 (SCvx)");
    if (cfg && cfg->algorithm == 1 && !(cfg->scvx_trust_region > 0. && cfg->scvx_alpha > 1. && cfg->scvx_beta > 1.))
        return fail(SCPP_B200_ERR_ARG, "scpp_b200_create: SCvx needs trust_region > 0, alpha > 1, beta > 1");
    if (!params || !cfg || !out || n <= 0) return fail(SCPP_B200_ERR_ARG, "scpp_b200_create: bad argument");
    if (cfg->K < 3 || cfg->max_iterations < 1 || cfg->nsub == 0) return fail(SCPP_B200_ERR_ARG, "scpp_b200_create: K >= 3, max_iterations >= 1, nsub != 0 required");
    if (!cfg->interpolate_input && (cfg->algorithm != 0 || (model != SCPP_B200_MODEL_ROCKETQUAT && model != SCPP_B200_MODEL_ROCKET2D)))
        return fail(SCPP_B200_ERR_UNSUPPORTED, "interpolate_input = false (zero-order hold) is built for the SC algorithm on RocketQuat and Rocket2D (models with a placeholder input, models.cuh)");
    if ((params->enable_roll_control != 0) != (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL))
        return fail(SCPP_B200_ERR_UNSUPPORTED, "enable_roll_control: use model SCPP_B200_MODEL_ROCKETQUAT_ROLL for true, SCPP_B200_MODEL_ROCKETQUAT for false");
    if (scpp_b200_device_count() <= 0) return fail(SCPP_B200_ERR_CUDA, "no CUDA device: libscpp_b200 has no CPU execution path");
    scpp_b200_engine *e = nullptr;
    if (model == SCPP_B200_MODEL_ROCKETQUAT) e = new EngineT<RocketQuat>();
    else if (model == SCPP_B200_MODEL_ROCKET2D) e = new EngineT<Rocket2d>();
    else if (model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) {
        if (!params->constrain_initial_final) return fail(SCPP_B200_ERR_UNSUPPORTED, "the plugin model's table was generated with constrain_initial_final = true");
        e = new EngineT<Rocket2dPlugin>();
    }
    else if (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) e = new EngineT<RocketQuatRollPlugin>();
    else return fail(SCPP_B200_ERR_ARG, "unknown model");
    e->model = model; e->N = n; e->device = device;
    memcpy(&e->P, params, sizeof(ModelParamsHost));
    memcpy(&e->cfg, cfg, sizeof(ScConfig));
    int rc = e->init();
    if (rc) { delete e; return rc; }
    *out = e;
    return 0;
}
void scpp_b200_destroy(scpp_b200_engine *e) { delete e; }
int scpp_b200_set_boundary_states(scpp_b200_engine *e, const double *xi, const double *xf) { return (e && xi && xf) ? e->set_boundary(xi, xf) : fail(SCPP_B200_ERR_ARG, "null argument"); }
int scpp_b200_set_instance_params(scpp_b200_engine *e, const scpp_b200_model_params *Pn) { return e ? e->set_instance_params(Pn) : fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_solve(scpp_b200_engine *e, int warm) { return e ? e->solve(warm) : fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_get_solution(scpp_b200_engine *e, double *X, double *U, double *t, int *it, int *fl) { return e ? e->get_solution(X, U, t, it, fl) : fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_get_iterate(scpp_b200_engine *e, int it, double *X, double *U, double *t) { return e ? e->get_iterate(it, X, U, t) : fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_get_iterate_dimensional(scpp_b200_engine *e, int it, double *X, double *U, double *t) { return e ? e->get_iterate(it, X, U, t, 1) : fail(SCPP_B200_ERR_ARG, "null engine"); }
int scpp_b200_get_info(scpp_b200_engine *e, double *info) { return (e && info) ? e->get_info(info) : fail(SCPP_B200_ERR_ARG, "null argument"); }
int scpp_b200_last_timing(scpp_b200_engine *e, double *a, double *b, double *c, int *l, int *o, long long *ii)
{
    if (!e) return fail(SCPP_B200_ERR_ARG, "null engine");
    if (a) *a = e->ms_disc; if (b) *b = e->ms_socp; if (c) *c = e->ms_total; if (l) *l = e->launches; if (o) *o = e->outer; if (ii) *ii = e->inst_iters;
    return 0;
}
int scpp_b200_last_rounds(scpp_b200_engine *e, int *r, long long *ir)
{
    if (!e) return fail(SCPP_B200_ERR_ARG, "null engine");
    if (r) *r = e->rounds; if (ir) *ir = e->inst_rounds;
    return 0;
}
size_t scpp_b200_device_bytes(scpp_b200_engine *e) { return e ? e->bytes : 0; }
long long scpp_b200_global_active(scpp_b200_engine *e) { return e ? e->global_active : -1; }

} // extern "C"

template <class M>
static int discretize_hook(int K, int n, int nsub, int jacobian, int device, const double *X, const double *U, const double *sigma, const double *par,
                           double *A, double *B, double *C, double *s, double *z)
{
    constexpr int NX = M::NX, NU = M::NU, NC = NX + 2 * NU + 2;
    CU(cudaSetDevice(device));
    ScArrays<M> a;
    memset(&a, 0, sizeof(a));
    a.N = n; a.K = K;
    const size_t nd = (size_t)n * (K - 1) * NX * NC;
    CU(cudaMalloc((void **)&a.X, (size_t)n * K * NX * 8)); CU(cudaMalloc((void **)&a.U, (size_t)n * K * NU * 8));
    CU(cudaMalloc((void **)&a.sigma, (size_t)n * 8)); CU(cudaMalloc((void **)&a.par, (size_t)n * M::NP * 8)); CU(cudaMalloc((void **)&a.dd, nd * 8));
    CU(cudaMemcpy(a.X, X, (size_t)n * K * NX * 8, cudaMemcpyHostToDevice)); CU(cudaMemcpy(a.U, U, (size_t)n * K * NU * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(a.sigma, sigma, (size_t)n * 8, cudaMemcpyHostToDevice)); CU(cudaMemcpy(a.par, par, (size_t)n * M::NP * 8, cudaMemcpyHostToDevice));
    CU(launch_discretize<M>(a, nsub, jacobian, 1, nullptr, n, 0));
    CU(cudaGetLastError());
    std::vector<double> dd(nd);
    CU(cudaMemcpy(dd.data(), a.dd, nd * 8, cudaMemcpyDeviceToHost));
    cudaFree(a.X); cudaFree(a.U); cudaFree(a.sigma); cudaFree(a.par); cudaFree(a.dd);
    for (size_t b = 0; b < (size_t)n * (K - 1); b++) {
        const double *t = dd.data() + b * NX * NC;
        for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) A[b * NX * NX + i + NX * j] = t[i * NC + j];
            for (int j = 0; j < NU; j++) { B[b * NX * NU + i + NX * j] = t[i * NC + NX + j]; C[b * NX * NU + i + NX * j] = t[i * NC + NX + NU + j]; }
            s[b * NX + i] = t[i * NC + NX + 2 * NU]; z[b * NX + i] = t[i * NC + NX + 2 * NU + 1];
        }
    }
    return 0;
}
extern "C" {

int scpp_b200_discretize2(int model, int K, int n, int nsub, int jacobian, int device, const double *X, const double *U, const double *sigma, const double *par,
                          double *A, double *B, double *C, double *s, double *z)
{
    if (K < 2 || n <= 0 || nsub == 0 || !X || !U || !sigma || !par || !A || !B || !C || !s || !z) return fail(SCPP_B200_ERR_ARG, "scpp_b200_discretize: bad argument");
    if (scpp_b200_device_count() <= 0) return fail(SCPP_B200_ERR_CUDA, "no CUDA device: libscpp_b200 has no CPU execution path");
    if (model == SCPP_B200_MODEL_ROCKETQUAT) return discretize_hook<RocketQuat>(K, n, nsub, jacobian, device, X, U, sigma, par, A, B, C, s, z);
    if (model == SCPP_B200_MODEL_ROCKET2D) return discretize_hook<Rocket2d>(K, n, nsub, jacobian, device, X, U, sigma, par, A, B, C, s, z);
    if (model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) return discretize_hook<Rocket2dPlugin>(K, n, nsub, jacobian, device, X, U, sigma, par, A, B, C, s, z);
    if (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) return discretize_hook<RocketQuatRollPlugin>(K, n, nsub, jacobian, device, X, U, sigma, par, A, B, C, s, z);
    return fail(SCPP_B200_ERR_ARG, "unknown model");
}
int scpp_b200_discretize(int model, int K, int n, int nsub, int device, const double *X, const double *U, const double *sigma, const double *par,
                         double *A, double *B, double *C, double *s, double *z)
{
    return scpp_b200_discretize2(model, K, n, nsub, 1, device, X, U, sigma, par, A, B, C, s, z);
}

int scpp_b200_lqr_gains(scpp_b200_engine *e, const double *q_diag, const double *r_diag, double *gains, int *ok)
{
    return e ? e->lqr_gains(q_diag, r_diag, gains, ok) : fail(SCPP_B200_ERR_ARG, "null engine");
}
int scpp_b200_sim_step(scpp_b200_engine *e, double time_step, double *x_new, double *u0, int *reached)
{
    return e ? e->sim_step(time_step, x_new, u0, reached) : fail(SCPP_B200_ERR_ARG, "null engine");
}

int scpp_b200_simulate(int model, int n, double dt, int device, double *x, const double *u0, const double *u1, const double *par)
{
    if (n <= 0 || !(dt > 0.) || !x || !u0 || !u1 || !par) return fail(SCPP_B200_ERR_ARG, "scpp_b200_simulate: bad argument");
    if (scpp_b200_device_count() <= 0) return fail(SCPP_B200_ERR_CUDA, "no CUDA device: libscpp_b200 has no CPU execution path");
    if (model == SCPP_B200_MODEL_ROCKETQUAT) return simulate_hook<RocketQuat>(n, dt, device, x, u0, u1, par);
    if (model == SCPP_B200_MODEL_ROCKET2D) return simulate_hook<Rocket2d>(n, dt, device, x, u0, u1, par);
    if (model == SCPP_B200_MODEL_ROCKET2D_PLUGIN) return simulate_hook<Rocket2dPlugin>(n, dt, device, x, u0, u1, par);
    if (model == SCPP_B200_MODEL_ROCKETQUAT_ROLL) return simulate_hook<RocketQuatRollPlugin>(n, dt, device, x, u0, u1, par);
    return fail(SCPP_B200_ERR_ARG, "unknown model");
}

int scpp_b200_selftest_blockops(int device, double *max_abs_err)
{
    if (!max_abs_err) return fail(SCPP_B200_ERR_ARG, "null argument");
    if (scpp_b200_device_count() <= 0) return fail(SCPP_B200_ERR_CUDA, "no CUDA device: libscpp_b200 has no CPU execution path");
    CU(cudaSetDevice(device));
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, sizeof(double)));
    k_selftest_blockops<<<1, 32>>>(d);
    CU(cudaMemcpy(max_abs_err, d, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d);
    CU(cudaGetLastError());
    return 0;
}

int scpp_b200_comm_unique_id(char id[128])
{
    if (!g_nccl.load()) return fail(SCPP_B200_ERR_NCCL, "libnccl.so.2 not found");
    int rc = g_nccl.GetUniqueId(id);
    return rc ? fail(SCPP_B200_ERR_NCCL, "ncclGetUniqueId failed") : 0;
}
int scpp_b200_comm_init(scpp_b200_engine *e, int nranks, int rank, const char id[128])
{
    if (!e || nranks < 1 || rank < 0 || rank >= nranks) return fail(SCPP_B200_ERR_ARG, "scpp_b200_comm_init: bad argument");
    if (nranks == 1) return 0;
    if (!g_nccl.load()) return fail(SCPP_B200_ERR_NCCL, "libnccl.so.2 not found");
    CU(cudaSetDevice(e->device));
    NcclUid uid;
    memcpy(uid.b, id, 128);
    int rc = g_nccl.CommInitRank(&e->comm, nranks, uid, rank);
    if (rc) return fail(SCPP_B200_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    e->nranks = nranks; e->rank = rank;
    return e->comm_buffers();
}

} // extern "C"
