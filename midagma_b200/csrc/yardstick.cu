// FP64 pipe yardsticks: dependent-chain-free DFMA and DMMA.8x8x4 loops whose only
// purpose is to measure the machine's FP64 issue rate for the roofline denominator.
#include "common.cuh"
#include "small_dmma.cuh"
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

// the GEMM's register pattern: 4 x 4 accumulator tiles fed by 4 A and 4 B fragments, optionally
// reloaded from shared memory every step (mode 1) as in the k-loop of gemm_f64_kernel
__global__ void dmma_yardstick_tiles(int iters, double* sink, int mode) {
    __shared__ double sh[2][8][36];
    double c[4][4][2], a[4], b[4];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * 8 * 36; i += blockDim.x) (&sh[0][0][0])[i] = 1.0 + 1e-9 * i;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        b[i] = 1e-12 * (blockIdx.x + 1 + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = i + j;
    }
    for (int it = 0; it < iters; ++it) {
        if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = sh[it & 1][i][lane];
                b[i] = sh[it & 1][4 + i][lane];
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace dagma

using namespace dagma;

// flops per launch: ctas*(threads/32)*iters*16*512
extern "C" int dagma_bench_fp64_dmma_tiles(dagma_stream_t stream, int ctas, int threads, int iters, int mode,
                                           double* sink_dev) {
    dmma_yardstick_tiles<<<ctas, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev, mode);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

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

// ---- dependent-issue latencies of the instructions on the serial chain of the on-chip sweep
// (one warp, clock64 around a chain of N dependent operations); out[i] = cycles per operation.
namespace dagma {
__global__ void latency_probe(double* out, double seed) {
    constexpr int N = 256;
    __shared__ double sh[64];
    const int lane = threadIdx.x;
    sh[lane] = seed + lane;
    sh[32 + lane] = 0.0;
    __syncthreads();
    double x = seed, y = 1.0 + 1e-9 * lane;
    long long t0, t1;
    // 0: DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-12);
    t1 = clock64();
    if (lane == 0) out[0] = double(t1 - t0) / N;
    // 1: DMMA chain through the accumulator
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    t1 = clock64();
    if (lane == 0) out[1] = double(t1 - t0) / N;
    // 2: DMMA chain through the A operand (CS -> update dependency)
    double a0 = c0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        double d0 = 0.0, d1 = 0.0;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(d0), "+d"(d1) : "d"(a0), "d"(y));
        a0 = d0;
    }
    t1 = clock64();
    if (lane == 0) out[2] = double(t1 - t0) / N;
    // 3: 64-bit shuffle chain
    double s = a0 + c1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) out[3] = double(t1 - t0) / N;
    // 4: MUFU.RCP64H + one DFMA (dependent)
    double r = s + 3.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        double q;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r));
        r = fma(q, 1e-3, 2.0);
    }
    t1 = clock64();
    if (lane == 0) out[4] = double(t1 - t0) / N;
    // 5: shared-memory pointer chase (LDS.64 -> address)
    int idx = lane;
    double acc = r;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        const double v = sh[32 + (idx & 31)];
        idx = idx + (int)v + 1;
        acc += v;
    }
    t1 = clock64();
    if (lane == 0) out[5] = double(t1 - t0) / N;
    // 6: STS -> __syncwarp -> LDS round trip (publish / read back)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        sh[lane] = acc;
        __syncwarp();
        acc = sh[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64();
    if (lane == 0) out[6] = double(t1 - t0) / N;
    // 7: DADD chain, 8: DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) acc = acc + y;
    t1 = clock64();
    if (lane == 0) out[7] = double(t1 - t0) / N;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) acc = acc * y;
    t1 = clock64();
    if (lane == 0) out[8] = double(t1 - t0) / N;
    // 9, 10: worst relative error of the MUFU seeds rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64
    double e_rcp = 0.0, e_rsq = 0.0;
    for (int i = 0; i < 20000; ++i) {
        const double xx = (1.0 + (lane * 20000 + i) * (3.0 / 640000.0)) * ((i & 1) ? 1e-7 : 37.0);
        double q, z;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(xx));
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(xx));
        e_rcp = fmax(e_rcp, fabs(q * xx - 1.0));
        e_rsq = fmax(e_rsq, fabs(z * z * xx - 1.0) * 0.5);
    }
    for (int off = 16; off > 0; off >>= 1) {
        e_rcp = fmax(e_rcp, __shfl_xor_sync(0xffffffffu, e_rcp, off));
        e_rsq = fmax(e_rsq, __shfl_xor_sync(0xffffffffu, e_rsq, off));
    }
    if (lane == 0) { out[9] = e_rcp; out[10] = e_rsq; }
    if (acc + idx == 123.456) out[15] = acc;
}
}  // namespace dagma

extern "C" int dagma_bench_latency(dagma_stream_t stream, double* out_dev) {
    dagma::latency_probe<<<1, 32, 0, (cudaStream_t)stream>>>(out_dev, 1.5);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- serial chain of the sweep in isolation: one warp inverts an 8 x 8 pivot block over and over
// (P -> P^{-1} -> P ...), through the same shared-memory hand-over the block step uses.
// out[0]: clk per stage_pivot_block + read-back; out[1]: the same with 10 independent DMMAs issued inside
// (the diagonal warp's publish-critical tiles); out[2]: max |P - inv(inv(P))| as a sanity check
namespace dagma {
__global__ void stage_probe_kernel(double* out) {
    extern __shared__ __align__(16) double sm[];
    const DmmaPos ps(threadIdx.x);
    double a[2][4][2];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) a[ti][tj][0] = a[ti][tj][1] = 0.0;
    const int i = ps.qr;
    auto p_of = [&](int r, int c) { return (r == c) ? 1.0 : -0.01 * (1 + ((r * 8 + c) % 5)); };
    a[0][0][0] = p_of(i, 2 * ps.qc);
    a[0][0][1] = p_of(i, 2 * ps.qc + 1);
    const int n0 = 2 * ps.qc, n1 = 2 * ps.qc + 1;
    const double* qd = sm + DmmaSmem::qbuf + 32 * (i >> 2) + (i & 3);
    constexpr int N = 512;
    double g[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int variant = 0; variant < 2; ++variant) {
        __syncwarp();
        const long long t0 = clock64();
        for (int it = 0; it < N; ++it) {
            if (variant == 0)
                stage_pivot_block<0, 0>(a, ps, sm, 0, 0, [](int) {});
            else
                stage_pivot_block<0, 0>(a, ps, sm, 0, 0, [&](int k) {
                    if (k < 5) {
                        dmma(g[k & 3][0], g[k & 3][1], a[0][0][0], a[0][0][1]);
                        dmma(g[(k + 1) & 3][0], g[(k + 1) & 3][1], a[0][0][1], a[0][0][0]);
                    }
                });
            __syncwarp();
            a[0][0][0] = -qd[4 * (2 * (n0 & 3) + (n0 >> 2))];
            a[0][0][1] = -qd[4 * (2 * (n1 & 3) + (n1 >> 2))];
            __syncwarp();
        }
        const long long t1 = clock64();
        if (ps.lane == 0) out[variant] = double(t1 - t0) / N;
    }
    double e = fmax(fabs(a[0][0][0] - p_of(i, n0)), fabs(a[0][0][1] - p_of(i, n1)));
    for (int off = 16; off > 0; off >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, off));
    if (ps.lane == 0) out[2] = e;
    if (g[0][0] + g[1][1] + g[2][0] + g[3][1] == 123.456) out[3] = g[0][0];
}
}  // namespace dagma

extern "C" int dagma_bench_stage(dagma_stream_t stream, double* out_dev) {
    DAGMA_CUDA_OK(cudaFuncSetAttribute(dagma::stage_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)dagma::DmmaSmem::bytes));
    dagma::stage_probe_kernel<<<1, 32, dagma::DmmaSmem::bytes, (cudaStream_t)stream>>>(out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
