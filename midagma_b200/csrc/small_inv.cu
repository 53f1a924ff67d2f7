// Batched on-chip fused slogdet + inverse for d <= 128 (API (i) of dagma_b200.h).
// One CTA per problem; the matrix lives in registers (RM x RN tile per thread) and is
// inverted by the Gauss-Jordan sweep of small_gj.cuh; log|det| is the sum of the logs
// of the pivots of the same sweep.  M is pre-scaled by a power of two so that s / 2^e
// lies in (0.5, 1] (exact), which keeps the sweep's publish identity in its accurate range.
#include "common.cuh"
#include "small_gj.cuh"
#include "small_dmma.cuh"
#include "small_inv2.cuh"
#include <cstdlib>
#include "../../include/dagma_b200.h"

namespace dagma {

struct InvArgs {
    int batch, d, lda, ldo, square;
    double s, inv_scale, log_scale;
    const double* a;
    double *logabsdet, *h, *minv, *grad, *min_entry;
    int* info;
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1) logdet_inv_small_kernel(const InvArgs P) {
    constexpr int RM = C::RM, RN = C::RN, NT = C::NT;
    __shared__ __align__(16) double linebuf[SweepSmem<C>::doubles + 2];
    __shared__ double red[96];
    const double* pivots = linebuf + SweepSmem<C>::piv_off;
    const uint32_t line_a = smem_u32(linebuf);
    const int tid = threadIdx.x;
    const ThreadPos<C> pos(tid);
    const int ty = pos.ty, tx = pos.tx, d = P.d;

    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        const double* A = P.a + (size_t)b * d * P.lda;
        double a[RM][RN], dummy[RM][RN];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) {
                const int r = C::grow(ty, i), c = C::gcol(tx, j);
                double x = (r < d && c < d) ? A[(size_t)r * P.lda + c] : 0.0;
                if (P.square) x *= x;
                a[i][j] = (((r == c) ? P.s : 0.0) - x) * P.inv_scale;
                dummy[i][j] = 0.0;
            }
        gj_sweep<C, false>(a, dummy, 0u, 0u, line_a, d, ty, tx);

        double ld = 0.0, zero1 = 0.0, zero2 = 0.0;
        bool badpiv = false;
        if (tid < d) {
            ld = log(fabs(pivots[tid]));
            badpiv = !(pivots[tid] > 0.0);
        }
        double mn = INFINITY;
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) {
                const int r = C::grow(ty, i), c = C::gcol(tx, j);
                if (r < d && c < d) {
                    const double mi = a[i][j] * P.inv_scale;
                    mn = fmin(mn, mi);
                    if (P.minv) P.minv[((size_t)b * d + r) * P.ldo + c] = mi;
                    if (P.grad) {   // grad[c][r] = (square ? 2 A[c][r] : 1) * Minv[r][c]
                        const double f = P.square ? 2.0 * A[(size_t)c * P.lda + r] : 1.0;
                        P.grad[((size_t)b * d + c) * P.ldo + r] = f * mi;
                    }
                }
            }
        block_sum3<NT>(ld, zero1, zero2, red, tid);
        mn = block_min<NT>(mn, red, tid);
        const int anybad = __syncthreads_or(badpiv);
        if (tid == 0) {
            const double lad = ld + (double)d * P.log_scale;
            if (P.logabsdet) P.logabsdet[b] = lad;
            if (P.h) P.h[b] = -lad + (double)d * log(P.s);
            if (P.min_entry) P.min_entry[b] = mn;
            if (P.info) P.info[b] = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
        }
        __syncthreads();
    }
}

// 32 < d <= 64: the tensor-core sweep of small_dmma.cuh (8 x 8 pivot blocks, rank-8 DMMA updates, one barrier per
// block step instead of one per pivot) -- the sweep the fit kernel and the tile steps of the blocked inverse use.
// One 256-thread CTA per problem, two per SM; the matrix lives in the DMMA accumulator layout
// a[ti][tj][e] = M[16 wr + 8 ti + lane/4][32 wc + 8 tj + 2 (lane%4) + e], identity padded to 64 x 64.
__global__ void __launch_bounds__(DM_NT, 2) logdet_inv_dmma_kernel(const InvArgs P) {
    extern __shared__ __align__(16) double psm[];
    using S = DmmaSmem;
    const int tid = threadIdx.x, d = P.d;
    const DmmaPos ps(tid);
    double* red = psm + S::red;
    SweepSync sy{smem_u32(psm + S::mbar), 0u};
    if (tid == 0) mbar_init(sy.bar, DM_NT / 32);
    __syncthreads();
    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        const double* A = P.a + (size_t)b * d * P.lda;
        double a[2][4][2];
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = ps.row(ti), c = ps.col(tj) + e;
                    double v = (r == c) ? 1.0 : 0.0;
                    if (r < d && c < d) {
                        double x = A[(size_t)r * P.lda + c];
                        if (P.square) x *= x;
                        v = (((r == c) ? P.s : 0.0) - x) * P.inv_scale;
                    }
                    a[ti][tj][e] = v;
                }
        dmma_sweep(a, ps, psm, d, sy);
        // log|det| from the fraction-free pivots: sum_k ((k & 3) - 2) log|p_k| over whole groups of four
        double ld = 0.0, zero1 = 0.0, zero2 = 0.0;
        bool badpiv = false;
        if (tid < ((d + 3) & ~3)) {
            const double p = psm[S::pinfo + tid];
            ld = (double)((tid & 3) - 2) * log(fabs(p));
            badpiv = !(p > 0.0);
        }
        double mn = INFINITY;
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = ps.row(ti), c = ps.col(tj) + e;
                    if (r < d && c < d) {
                        const double mi = a[ti][tj][e] * P.inv_scale;
                        mn = fmin(mn, mi);
                        if (P.minv) P.minv[((size_t)b * d + r) * P.ldo + c] = mi;
                        if (P.grad) {   // grad[c][r] = (square ? 2 A[c][r] : 1) * Minv[r][c]
                            const double f = P.square ? 2.0 * A[(size_t)c * P.lda + r] : 1.0;
                            P.grad[((size_t)b * d + c) * P.ldo + r] = f * mi;
                        }
                    }
                }
        block_sum3<DM_NT>(ld, zero1, zero2, red, tid);
        mn = block_min<DM_NT>(mn, red, tid);
        const int anybad = __syncthreads_or(badpiv);
        if (tid == 0) {
            const double lad = ld + (double)d * P.log_scale;
            if (P.logabsdet) P.logabsdet[b] = lad;
            if (P.h) P.h[b] = -lad + (double)d * log(P.s);
            if (P.min_entry) P.min_entry[b] = mn;
            if (P.info) P.info[b] = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
        }
        __syncthreads();
    }
}
// 64 < d <= 128: two-by-two block Gauss-Jordan on chip (small_inv2.cuh), one 256-thread CTA per problem and SM
__global__ void __launch_bounds__(DM_NT, 1) logdet_inv_dmma2_kernel(const InvArgs P) {
    extern __shared__ __align__(16) double psm[];
    using S = DmmaSmem;
    constexpr int LD = DM_LD;
    const int tid = threadIdx.x, d = P.d;
    const DmmaPos ps(tid);
    double* red = psm + S::red;
    double* pv = psm + Inv2Smem::piv;
    SweepSync sy{smem_u32(psm + S::mbar), 0u};
    if (tid == 0) mbar_init(sy.bar, DM_NT / 32);
    __syncthreads();
    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        const double* A = P.a + (size_t)b * d * P.lda;
        // ---- M = (s I - A o A) / scale into the four tiles, identity padded to 128 x 128
        for (int e = tid; e < 4 * DM_DP * DM_DP; e += DM_NT) {
            const int r = e >> 7, c = e & 127;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < d && c < d) {
                double x = A[(size_t)r * P.lda + c];
                if (P.square) x *= x;
                v = (((r == c) ? P.s : 0.0) - x) * P.inv_scale;
            }
            psm[Inv2Smem::at(r >> 6, c >> 6) + (r & 63) * LD + (c & 63)] = v;
        }
        __syncthreads();
        inv2_block_gj(psm, ps, sy, d);
        // ---- outputs
        double ld = 0.0, zero1 = 0.0, zero2 = 0.0;
        bool badpiv = false;
        if (tid < 2 * DM_DP) {
            const double p = pv[tid];
            ld = (double)((tid & 3) - 2) * log(fabs(p));
            badpiv = !(p > 0.0);
        }
        double mn = INFINITY;
        for (int e = tid; e < d * d; e += DM_NT) {
            const int r = e / d, c = e - r * d;
            const double mi = psm[Inv2Smem::at(r >> 6, c >> 6) + (r & 63) * LD + (c & 63)] * P.inv_scale;
            mn = fmin(mn, mi);
            if (P.minv) P.minv[((size_t)b * d + r) * P.ldo + c] = mi;
        }
        if (P.grad)      // grad[r][c] = (square ? 2 A[r][c] : 1) * Minv[c][r]: coalesced over c, transposed read from shared
            for (int e = tid; e < d * d; e += DM_NT) {
                const int r = e / d, c = e - r * d;
                const double mt = psm[Inv2Smem::at(c >> 6, r >> 6) + (c & 63) * LD + (r & 63)] * P.inv_scale;
                const double f = P.square ? 2.0 * A[(size_t)r * P.lda + c] : 1.0;
                P.grad[((size_t)b * d + r) * P.ldo + c] = f * mt;
            }
        block_sum3<DM_NT>(ld, zero1, zero2, red, tid);
        mn = block_min<DM_NT>(mn, red, tid);
        const int anybad = __syncthreads_or(badpiv);
        if (tid == 0) {
            const double lad = ld + (double)d * P.log_scale;
            if (P.logabsdet) P.logabsdet[b] = lad;
            if (P.h) P.h[b] = -lad + (double)d * log(P.s);
            if (P.min_entry) P.min_entry[b] = mn;
            if (P.info) P.info[b] = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
        }
        __syncthreads();
    }
}
// DAGMA_SMALL_INV_DMMA (A-B timing): 1 (default) = tensor-core sweeps for 32 < d <= 128, 0 = scalar rank-1 sweep
static bool small_inv_dmma() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DAGMA_SMALL_INV_DMMA");
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

template <class C>
static int launch_inv(cudaStream_t stream, const InvArgs& a, int ctas) {
    logdet_inv_small_kernel<C><<<ctas, C::NT, 0, stream>>>(a);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

int logdet_inv_small(cudaStream_t stream, int batch, int d, double s, const double* a_dev, int lda,
                     int square, double* logabsdet, double* h, double* minv, double* grad, int ldo,
                     double* min_entry, int* info) {
    int dev = 0, sms = 0;
    DAGMA_CUDA_OK(cudaGetDevice(&dev));
    DAGMA_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double scale = 1.0;
    if (s > 0.0 && isfinite(s)) {
        int e = 0;
        const double f = frexp(s, &e);      // s = f * 2^e, f in [0.5, 1)
        if (f == 0.5) --e;                  // keep s / scale in (0.5, 1]
        scale = ldexp(1.0, e);
    }
    InvArgs P{batch, d, lda, ldo, square, s, 1.0 / scale, log(scale), a_dev, logabsdet, h, minv, grad,
              min_entry, info};
    if (d > 32 && d <= 64 && small_inv_dmma()) {
        static bool attr = false;
        if (!attr) {
            DAGMA_CUDA_OK(cudaFuncSetAttribute(logdet_inv_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)DmmaSmem::bytes));
            attr = true;
        }
        logdet_inv_dmma_kernel<<<batch < 2 * sms ? batch : 2 * sms, DM_NT, DmmaSmem::bytes, stream>>>(P);
        DAGMA_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (d > 64 && d <= 128 && small_inv_dmma()) {
        static bool attr2 = false;
        if (!attr2) {
            DAGMA_CUDA_OK(cudaFuncSetAttribute(logdet_inv_dmma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)Inv2Smem::bytes));
            attr2 = true;
        }
        logdet_inv_dmma2_kernel<<<batch < sms ? batch : sms, DM_NT, Inv2Smem::bytes, stream>>>(P);
        DAGMA_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int ctas = batch < 4 * sms ? batch : 4 * sms;
    if (d <= 16) return launch_inv<Cfg<1, 1, 16, 16>>(stream, P, ctas);
    if (d <= 32) return launch_inv<Cfg<2, 2, 16, 16>>(stream, P, ctas);
    if (d <= 48) return launch_inv<Cfg<4, 2, 12, 24>>(stream, P, ctas);
    if (d <= 64) return launch_inv<Cfg<4, 2, 16, 32>>(stream, P, ctas);
    if (d <= 96) return launch_inv<Cfg<8, 4, 12, 24>>(stream, P, ctas);
    return launch_inv<Cfg<8, 4, 16, 32>>(stream, P, ctas);
}

}  // namespace dagma
