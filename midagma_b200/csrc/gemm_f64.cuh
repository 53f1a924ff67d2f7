// FP64 tensor-core GEMM for sm_100a:  C = alpha * op(A) * B + beta * C   (row-major).
//
// The native FP64 tensor instruction on Blackwell is DMMA.8x8x4 (mma.sync.m8n8k4.f64);
// tcgen05 has no f64 kind.  CTA tile 128 x 128 x 16, 256 threads = 8 warps in a 4 x 2
// grid, warp tile 32 x 64 = 4 x 8 DMMA tiles (64 accumulator doubles per thread), so
// one k4 step needs 12 fragment loads for 32 DMMAs.  Operand slabs are staged in
// shared memory by cp.async (16-byte, zero-filled at the edges), double buffered.
// Shared layouts are padded so that every fragment load is bank-conflict free:
//   A (non-transposed): As[m][k], row stride 20 doubles; fragment a = As[r0 + lane/4][k0 + lane%4]
//   A (transposed, A given as K x M): As[k][m], row stride BM + 4; a = As[k0 + lane%4][r0 + lane/4]
//   B: Bs[k][n], row stride BN + 4; fragment b = Bs[k0 + lane%4][c0 + lane/4]
// Split-K (gridDim.z > 1) writes partial tiles to a workspace that a second kernel
// reduces in a fixed order, so results are deterministic.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dagma {

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;
constexpr int LDA_N = GBK + 4;      // As[m][k]
constexpr int LDA_T = GBM + 4;      // As[k][m]
constexpr int LDB_S = GBN + 4;      // Bs[k][n]
constexpr int A_STAGE = (GBM * LDA_N > GBK * LDA_T) ? GBM * LDA_N : GBK * LDA_T;   // doubles
constexpr int B_STAGE = GBK * LDB_S;
constexpr size_t GEMM_SMEM_BYTES = (size_t)2 * (A_STAGE + B_STAGE) * sizeof(double);

enum GemmEpilogue { EPI_NONE = 0, EPI_SIGMOID = 1 };

struct GemmArgs {
    int M, N, K;
    const double* A; int lda;    // M x K (or K x M when TRANS_A)
    const double* B; int ldb;    // K x N
    double* C; int ldc;          // M x N
    double alpha, beta;
    int k_chunk;                 // K range per z-slice (multiple of GBK); K when no split
    double* partial;             // [gridDim.z][M][N] when split-K, else nullptr
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int n = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, bool pred) {
    const int n = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// load a [rows x cols] slab (row-major source with leading dimension ld) into shared
// memory with row stride lds; elements outside [max_r, max_c) are zero-filled.
template <int ROWS, int COLS>
__device__ __forceinline__ void load_slab(uint32_t smem, int lds, const double* src, int ld, int r0, int c0,
                                          int max_r, int max_c, bool vec2, int tid) {
    if (vec2) {
        constexpr int CH = COLS / 2;
        for (int e = tid; e < ROWS * CH; e += GTHREADS) {
            const int r = e / CH, c = (e - r * CH) * 2;
            const bool ok = (r0 + r < max_r) && (c0 + c + 1 < max_c);
            const bool ok1 = (r0 + r < max_r) && (c0 + c < max_c);
            const double* p = src + (size_t)(r0 + r) * ld + c0 + c;
            if (ok || !ok1)
                cp_async16(smem + (r * lds + c) * 8, ok ? p : src, ok);
            else {   // last odd column: 8 bytes valid
                cp_async8(smem + (r * lds + c) * 8, p, true);
                cp_async8(smem + (r * lds + c + 1) * 8, src, false);
            }
        }
    } else {
        for (int e = tid; e < ROWS * COLS; e += GTHREADS) {
            const int r = e / COLS, c = e - r * COLS;
            const bool ok = (r0 + r < max_r) && (c0 + c < max_c);
            cp_async8(smem + (r * lds + c) * 8, ok ? src + (size_t)(r0 + r) * ld + c0 + c : src, ok);
        }
    }
}

template <bool TRANS_A, int EPI>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_f64_kernel(const GemmArgs P) {
    extern __shared__ __align__(16) double gsm[];
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(gsm));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;               // 4 x 2 warps
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int kbeg = blockIdx.z * P.k_chunk;
    const int kend = min(P.K, kbeg + P.k_chunk);
    const int nk = (kend - kbeg + GBK - 1) / GBK;

    const bool vecA = ((P.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0);
    const bool vecB = ((P.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0);

    auto stageA = [&](int s) { return sbase + (uint32_t)(s * (A_STAGE + B_STAGE)) * 8; };
    auto stageB = [&](int s) { return sbase + (uint32_t)(s * (A_STAGE + B_STAGE) + A_STAGE) * 8; };
    auto issue = [&](int kt, int s) {
        const int k0 = kbeg + kt * GBK;
        if (TRANS_A)   // A is K x M: slab rows = k, cols = m
            load_slab<GBK, GBM>(stageA(s), LDA_T, P.A, P.lda, k0, m0, kend, P.M, vecA, tid);
        else
            load_slab<GBM, GBK>(stageA(s), LDA_N, P.A, P.lda, m0, k0, P.M, kend, vecA, tid);
        load_slab<GBK, GBN>(stageB(s), LDB_S, P.B, P.ldb, k0, n0, kend, P.N, vecB, tid);
        cp_async_commit();
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (nk > 0) issue(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        if (kt + 1 < nk) {
            issue(kt + 1, s ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double* As = gsm + s * (A_STAGE + B_STAGE);
        const double* Bs = As + A_STAGE;
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = wm * 32 + i * 8 + (lane >> 2);
                a[i] = TRANS_A ? As[(kk + (lane & 3)) * LDA_T + r] : As[r * LDA_N + kk + (lane & 3)];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Bs[(kk + (lane & 3)) * LDB_S + wn * 64 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    // ---- epilogue: thread holds C[r][c], C[r][c+1] with r = lane/4, c = 2*(lane%4) in each 8x8 tile
    const bool split = (P.partial != nullptr);
    double* out = split ? P.partial + (size_t)blockIdx.z * P.M * P.N : P.C;
    const int ldo = split ? P.N : P.ldc;
    const bool vecC = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + wm * 32 + i * 8 + (lane >> 2);
        if (r >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n0 + wn * 64 + j * 8 + 2 * (lane & 3);
            if (c >= P.N) continue;
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            double* p = out + (size_t)r * ldo + c;
            const bool two = (c + 1 < P.N);
            if (!split) {
                v0 *= P.alpha;
                v1 *= P.alpha;
                if (P.beta != 0.0) {
                    v0 = fma(P.beta, p[0], v0);
                    if (two) v1 = fma(P.beta, p[1], v1);
                }
                if (EPI == EPI_SIGMOID) {
                    v0 = 1.0 / (1.0 + exp(-v0));
                    v1 = 1.0 / (1.0 + exp(-v1));
                }
            }
            if (two && vecC)
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
            else {
                p[0] = v0;
                if (two) p[1] = v1;
            }
        }
    }
}

// C = alpha * sum_z partial[z] + beta * C   (fixed summation order)
template <int EPI>
__global__ void splitk_reduce_kernel(const double* __restrict__ partial, int splits, int M, int N, double* C,
                                     int ldc, double alpha, double beta) {
    const size_t total = (size_t)M * N;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + e];
        const int r = (int)(e / N), c = (int)(e - (size_t)r * N);
        double* p = C + (size_t)r * ldc + c;
        double v = alpha * s;
        if (beta != 0.0) v = fma(beta, *p, v);
        if (EPI == EPI_SIGMOID) v = 1.0 / (1.0 + exp(-v));
        *p = v;
    }
}

}  // namespace dagma
