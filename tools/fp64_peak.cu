// tools/fp64_peak.cu — measures the FP64 peaks this path is bounded by (MEASURED_PEAKS.json has none):
//   DFMA  : dependent-chain-free fused multiply-add throughput on the CUDA cores
//   DMMA  : mma.sync.m8n8k4.f64 throughput on the tensor cores (tcgen05 has no FP64)
// Writes one JSON line to stdout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_dmma(double *out, int iters)
{
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[0] + c0[1] + c1[1] + c0[2] + c1[2] + c0[3] + c1[3];
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    double best_fma = 0, best_mma = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best_fma) best_fma = tf;
        cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double tm = 2.0 * 256 * 4 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
        if (tm > best_mma) best_mma = tm;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"how\": \"8 independent DFMA chains x 256 thr x %d CTAs; 4 independent mma.sync.m8n8k4.f64 chains per warp; best of 5, CUDA events\", \"err\": \"%s\"}\n",
           p.name, p.multiProcessorCount, best_fma, best_mma, blocks, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
