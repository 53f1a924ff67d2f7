// Batched on-chip fused slogdet + inverse for d <= 128 (API (i) of dagma_b200.h).
// One CTA per problem; the matrix lives in registers (RM x RN tile per thread) and is
// inverted by the Gauss-Jordan sweep of small_gj.cuh; log|det| is the sum of the logs
// of the pivots of the same sweep.  M is pre-scaled by a power of two so that s / 2^e
// lies in (0.5, 1] (exact), which keeps the sweep's publish identity in its accurate range.
#include "common.cuh"
#include "small_gj.cuh"
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
    const int ctas = batch < 4 * sms ? batch : 4 * sms;
    if (d <= 16) return launch_inv<Cfg<1, 1, 16, 16>>(stream, P, ctas);
    if (d <= 32) return launch_inv<Cfg<2, 2, 16, 16>>(stream, P, ctas);
    if (d <= 48) return launch_inv<Cfg<4, 2, 12, 24>>(stream, P, ctas);
    if (d <= 64) return launch_inv<Cfg<4, 2, 16, 32>>(stream, P, ctas);
    if (d <= 96) return launch_inv<Cfg<8, 4, 12, 24>>(stream, P, ctas);
    return launch_inv<Cfg<8, 4, 16, 32>>(stream, P, ctas);
}

}  // namespace dagma
