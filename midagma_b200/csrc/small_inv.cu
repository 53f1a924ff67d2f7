// Batched on-chip fused slogdet + inverse for d <= 128 (API (i) of dagma_b200.h).
// One 256-thread CTA per problem; the matrix lives in registers (R x R tile per
// thread, R = ceil(d / 16) <= 8) and is inverted by the Gauss-Jordan sweep of
// small_gj.cuh; log|det| is the sum of the logs of the pivots of the same sweep.
#include "common.cuh"
#include "small_gj.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

struct InvArgs {
    int batch, d, lda, ldo, square;
    double s;
    const double* a;
    double *logabsdet, *h, *minv, *grad, *min_entry;
    int* info;
};

template <int R>
__global__ void __launch_bounds__(NT, 1) logdet_inv_small_kernel(const InvArgs P) {
    using T = Tile<R>;
    constexpr int DP = T::DP;
    __shared__ __align__(16) double rowbuf[2 * DP];
    __shared__ __align__(16) double colbuf[2 * DP];
    __shared__ double pinvbuf[2];
    __shared__ double pivots[DP];
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const ThreadPos pos(tid);
    const int ty = pos.ty, tx = pos.tx, d = P.d;

    for (int b = blockIdx.x; b < P.batch; b += gridDim.x) {
        const double* A = P.a + (size_t)b * d * P.lda;
        double a[R][R], dummy[R][R];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int r = T::g(ty, i), c = T::g(tx, j);
                double x = (r < d && c < d) ? A[(size_t)r * P.lda + c] : 0.0;
                if (P.square) x *= x;
                a[i][j] = ((r == c) ? P.s : 0.0) - x;
                dummy[i][j] = 0.0;
            }
        gj_sweep<R, false>(a, dummy, nullptr, nullptr, rowbuf, colbuf, pinvbuf, pivots, d, ty, tx);

        double ld = 0.0, zero1 = 0.0, zero2 = 0.0;
        bool badpiv = false;
        if (tid < d) {
            ld = log(fabs(pivots[tid]));
            badpiv = !(pivots[tid] > 0.0);
        }
        double mn = INFINITY;
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int r = T::g(ty, i), c = T::g(tx, j);
                if (r < d && c < d) {
                    mn = fmin(mn, a[i][j]);
                    if (P.minv) P.minv[((size_t)b * d + r) * P.ldo + c] = a[i][j];
                    if (P.grad) {   // grad[c][r] = (square ? 2 A[c][r] : 1) * Minv[r][c]
                        const double f = P.square ? 2.0 * A[(size_t)c * P.lda + r] : 1.0;
                        P.grad[((size_t)b * d + c) * P.ldo + r] = f * a[i][j];
                    }
                }
            }
        block_sum3(ld, zero1, zero2, red, tid);
        mn = block_min(mn, red, tid);
        const int anybad = __syncthreads_or(badpiv);
        if (tid == 0) {
            if (P.logabsdet) P.logabsdet[b] = ld;
            if (P.h) P.h[b] = -ld + (double)d * log(P.s);
            if (P.min_entry) P.min_entry[b] = mn;
            if (P.info) P.info[b] = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
        }
        __syncthreads();
    }
}

template <int R>
static int launch_inv(cudaStream_t stream, const InvArgs& a, int ctas) {
    logdet_inv_small_kernel<R><<<ctas, NT, 0, stream>>>(a);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

int logdet_inv_small(cudaStream_t stream, int batch, int d, double s, const double* a_dev, int lda,
                     int square, double* logabsdet, double* h, double* minv, double* grad, int ldo,
                     double* min_entry, int* info) {
    int dev = 0, sms = 0;
    DAGMA_CUDA_OK(cudaGetDevice(&dev));
    DAGMA_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    InvArgs P{batch, d, lda, ldo, square, s, a_dev, logabsdet, h, minv, grad, min_entry, info};
    const int ctas = batch < 4 * sms ? batch : 4 * sms;
    switch ((d + TG - 1) / TG) {
        case 1: return launch_inv<1>(stream, P, ctas);
        case 2: return launch_inv<2>(stream, P, ctas);
        case 3: return launch_inv<3>(stream, P, ctas);
        case 4: return launch_inv<4>(stream, P, ctas);
        case 5: return launch_inv<5>(stream, P, ctas);
        case 6: return launch_inv<6>(stream, P, ctas);
        case 7: return launch_inv<7>(stream, P, ctas);
        default: return launch_inv<8>(stream, P, ctas);
    }
}

}  // namespace dagma
