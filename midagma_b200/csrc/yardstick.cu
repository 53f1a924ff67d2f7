// FP64 pipe yardsticks: dependent-chain-free DFMA and DMMA.8x8x4 loops whose only
// purpose is to measure the machine's FP64 issue rate for the roofline denominator.
#include "common.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

__global__ void fma_yardstick(int iters, double* sink) {
    double acc[16];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) sink[0] = s;
}

__global__ void dmma_yardstick(int iters, double* sink) {
    double c[8][2];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-12 * (blockIdx.x + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace dagma

using namespace dagma;

// flops per launch: ctas*threads*iters*16*2 (fma); ctas*(threads/32)*iters*8*512 (dmma)
extern "C" int dagma_bench_fp64_fma(dagma_stream_t stream, int ctas, int threads, int iters, double* sink_dev) {
    fma_yardstick<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_bench_fp64_dmma(dagma_stream_t stream, int ctas, int threads, int iters, double* sink_dev) {
    dmma_yardstick<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
