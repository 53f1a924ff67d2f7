// FP64 tensor-core GEMM for sm_100a:  C = alpha * op(A) * B + beta * C   (row-major).
//
// The native FP64 tensor instruction on Blackwell is DMMA.8x8x4 (mma.sync.m8n8k4.f64);
// tcgen05 has no f64 kind.  The FP64 pipe takes one DMMA per 16 clk per SM sub-partition, so
// the kernel is built for overlap rather than for operand reuse: CTA tile BM x BN x 16 with
// 256 threads = 8 warps in a 4 x 2 grid, warp tile (BM/4) x (BN/2); the 128 x 64 tile keeps 32
// accumulator doubles per thread so that TWO CTAs are resident per SM and one CTA's barrier /
// cp.async waits are covered by the other's DMMAs.  Operand slabs are staged in shared memory
// by a 3-stage cp.async pipeline (16-byte copies, zero-filled at the edges), one barrier per
// k-tile.  When alpha = beta = 1 (the rank-k updates of the blocked inverse) the accumulators
// are initialised from C, so the epilogue is a plain store.
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

constexpr int GBK = 16, GSTAGES = 3;

template <int BM, int BN, int WM, int WN>      // CTA tile and warp grid
struct GemmTile {
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int LDA_N = GBK + 4;      // As[m][k]
    static constexpr int LDA_T = BM + 4;       // As[k][m]
    static constexpr int LDB_S = BN + 4;       // Bs[k][n]
    static constexpr int A_STAGE = (BM * LDA_N > GBK * LDA_T) ? BM * LDA_N : GBK * LDA_T;   // doubles
    static constexpr int B_STAGE = GBK * LDB_S;
    static constexpr size_t SMEM_BYTES = (size_t)GSTAGES * (A_STAGE + B_STAGE) * sizeof(double);
    static constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;   // DMMA tiles per warp
    static constexpr int MINB = (MI * NJ <= 16) ? (512 / THREADS) : 1;
    static_assert(BM % (8 * WM) == 0 && BN % (8 * WN) == 0, "warp tiles are whole DMMA tiles");
};

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
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// load a [rows x cols] slab (row-major source with leading dimension ld) into shared
// memory with row stride lds; elements outside [max_r, max_c) are zero-filled.
template <int ROWS, int COLS, int NTHREADS>
__device__ __forceinline__ void load_slab(uint32_t smem, int lds, const double* src, int ld, int r0, int c0,
                                          int max_r, int max_c, bool vec2, int tid) {
    if (vec2) {
        constexpr int CH = COLS / 2;
        for (int e = tid; e < ROWS * CH; e += NTHREADS) {
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
        for (int e = tid; e < ROWS * COLS; e += NTHREADS) {
            const int r = e / COLS, c = e - r * COLS;
            const bool ok = (r0 + r < max_r) && (c0 + c < max_c);
            cp_async8(smem + (r * lds + c) * 8, ok ? src + (size_t)(r0 + r) * ld + c0 + c : src, ok);
        }
    }
}

template <bool TRANS_A, int EPI, int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(GemmTile<BM, BN, WM, WN>::THREADS, GemmTile<BM, BN, WM, WN>::MINB)
gemm_f64_kernel(const GemmArgs P) {
    using T = GemmTile<BM, BN, WM, WN>;
    constexpr int NTH = T::THREADS;
    constexpr int MI = T::MI, NJ = T::NJ, STG = T::A_STAGE + T::B_STAGE;
    extern __shared__ __align__(16) double gsm[];
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(gsm));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * P.k_chunk;
    const int kend = min(P.K, kbeg + P.k_chunk);
    const int nk = (kend - kbeg + GBK - 1) / GBK;

    const bool vecA = ((P.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15) == 0);
    const bool vecB = ((P.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.B) & 15) == 0);

    auto issue = [&](int kt) {
        if (kt < nk) {
            const int s = kt % GSTAGES;
            const uint32_t sa = sbase + (uint32_t)(s * STG) * 8, sb = sa + (uint32_t)T::A_STAGE * 8;
            const int k0 = kbeg + kt * GBK;
            if (TRANS_A)   // A is K x M: slab rows = k, cols = m
                load_slab<GBK, BM, NTH>(sa, T::LDA_T, P.A, P.lda, k0, m0, kend, P.M, vecA, tid);
            else
                load_slab<BM, GBK, NTH>(sa, T::LDA_N, P.A, P.lda, m0, k0, P.M, kend, vecA, tid);
            load_slab<GBK, BN, NTH>(sb, T::LDB_S, P.B, P.ldb, k0, n0, kend, P.N, vecB, tid);
        }
        cp_async_commit();           // empty groups keep the wait arithmetic uniform
    };
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) issue(s);

    const bool split = (P.partial != nullptr);
    double* out = split ? P.partial + (size_t)blockIdx.z * P.M * P.N : P.C;
    const int ldo = split ? P.N : P.ldc;
    const bool vecC = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    // rank-k update fast path: acc starts as C (read while the first slabs are in flight)
    const bool acc_from_c = !split && EPI == EPI_NONE && P.alpha == 1.0 && P.beta == 1.0;

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = m0 + wm * (BM / WM) + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = n0 + wn * (BN / WN) + j * 8 + 2 * (lane & 3);
            double v0 = 0.0, v1 = 0.0;
            if (acc_from_c && r < P.M && c < P.N) {
                const double* p = P.C + (size_t)r * P.ldc + c;
                if (c + 1 < P.N && vecC) {
                    const double2 t = *reinterpret_cast<const double2*>(p);
                    v0 = t.x;
                    v1 = t.y;
                } else {
                    v0 = p[0];
                    if (c + 1 < P.N) v1 = p[1];
                }
            }
            acc[i][j][0] = v0;
            acc[i][j][1] = v1;
        }
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();                       // slab kt landed; slab kt-1 is free for reuse
        issue(kt + GSTAGES - 1);
        const double* As = gsm + (kt % GSTAGES) * STG;
        const double* Bs = As + T::A_STAGE;
#pragma unroll
        for (int kk = 0; kk < GBK; kk += 4) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int r = wm * (BM / WM) + i * 8 + (lane >> 2);
                a[i] = TRANS_A ? As[(kk + (lane & 3)) * T::LDA_T + r] : As[r * T::LDA_N + kk + (lane & 3)];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                b[j] = Bs[(kk + (lane & 3)) * T::LDB_S + wn * (BN / WN) + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: thread holds C[r][c], C[r][c+1] with r = lane/4, c = 2*(lane%4) in each 8x8 tile
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = m0 + wm * (BM / WM) + i * 8 + (lane >> 2);
        if (r >= P.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = n0 + wn * (BN / WN) + j * 8 + 2 * (lane & 3);
            if (c >= P.N) continue;
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            double* p = out + (size_t)r * ldo + c;
            const bool two = (c + 1 < P.N);
            if (!split && !acc_from_c) {
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
