// Perron pair of a non-negative matrix by power iteration, on device -- the spectral trek-cycle-coupling
// penalty of the fork (reference: src/notreks/notreks.py:156-239 perron_eig_with_gradA, :340-378 spectral branch).
//
// The work is HBM/L2 bound FP64 streaming (one pass over the n x n matrix per iteration, n = 2d), so ONE pass serves
// both chains of the reference (right vector v <- A v and left vector u <- A^T u): a CTA owns PW_ROWS rows, thread t
// owns the columns t, t + 256, ...; per element it does one coalesced load and two FMAs --
//   racc[r]   += A[r][c] * xr[c]   (row sums of the right product, reduced over the CTA once at the end)
//   lacc      += A[r][c] * xl[r]   (this row tile's share of column c of the left product, written to a partial row)
// The partial rows of the left product are summed in tile order by a second small kernel, which also produces the two
// squared norms (fixed-order, no atomics on doubles: bit-reproducible).  The normalisation x / (||x|| + eps) of the
// reference is not a separate pass: the next product reads the un-normalised vector and the device-resident norm.
#include "common.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

constexpr int PW_ROWS = 16;      // rows per CTA (250 CTAs at n = 4000)
constexpr int PW_NT = 256;

__device__ __forceinline__ double pw_inv_norm(const double* norm2, double eps) {
    return norm2 ? 1.0 / (sqrt(*norm2) + eps) : 1.0;
}

// yr = (A + shift I) (xr * ir),   partial[tile][c] = sum_{r in tile} A[r][c] (xl[r] * il)      (il, ir: see above)
// square != 0: the matrix is A o A (the Hadamard square W o W of the "exact_original_graph" baseline).
__global__ void __launch_bounds__(PW_NT) power_pass_kernel(int n, const double* __restrict__ A, int lda, int square,
                                                           double shift, const double* __restrict__ xr,
                                                           const double* __restrict__ xr_norm2,
                                                           const double* __restrict__ xl,
                                                           const double* __restrict__ xl_norm2, double eps,
                                                           double* __restrict__ yr, double* __restrict__ partial) {
    __shared__ double s_xl[PW_ROWS];
    __shared__ double s_red[PW_ROWS][PW_NT / 32];
    const int tid = threadIdx.x, r0 = blockIdx.x * PW_ROWS;
    const double ir = xr ? pw_inv_norm(xr_norm2, eps) : 0.0;
    const double il = xl ? pw_inv_norm(xl_norm2, eps) : 0.0;
    if (tid < PW_ROWS) s_xl[tid] = (xl && r0 + tid < n) ? xl[r0 + tid] * il : 0.0;
    __syncthreads();
    double racc[PW_ROWS];
#pragma unroll
    for (int r = 0; r < PW_ROWS; ++r) racc[r] = 0.0;
    for (int c = tid; c < n; c += PW_NT) {
        const double xc = xr ? xr[c] * ir : 0.0;
        double a[PW_ROWS];
#pragma unroll
        for (int r = 0; r < PW_ROWS; ++r) a[r] = (r0 + r < n) ? A[(size_t)(r0 + r) * lda + c] : 0.0;
        double lacc = 0.0;
#pragma unroll
        for (int r = 0; r < PW_ROWS; ++r) {
            const double v = square ? a[r] * a[r] : a[r];
            racc[r] = fma(v, xc, racc[r]);
            lacc = fma(v, s_xl[r], lacc);
        }
        if (partial) partial[(size_t)blockIdx.x * n + c] = lacc;
    }
    if (!yr) return;
#pragma unroll
    for (int r = 0; r < PW_ROWS; ++r) {
        double v = racc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[r][tid >> 5] = v;
    }
    __syncthreads();
    if (tid < PW_ROWS && r0 + tid < n) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < PW_NT / 32; ++w) v += s_red[tid][w];
        if (shift != 0.0) v = fma(shift, xr[r0 + tid] * ir, v);
        yr[r0 + tid] = v;
    }
}

// yl[c] = sum_tiles partial[tile][c] + shift * xl[c] * il;  norms2[0] = ||yr||^2, norms2[1] = ||yl||^2.
// One CTA (n is a few thousand): every sum has a fixed order.
__global__ void __launch_bounds__(1024) power_reduce_kernel(int n, int tiles, const double* __restrict__ partial,
                                                            double shift, const double* __restrict__ xl,
                                                            const double* __restrict__ xl_norm2, double eps,
                                                            const double* __restrict__ yr, double* __restrict__ yl,
                                                            double* __restrict__ norms2) {
    __shared__ double s_a[32], s_b[32];
    const int tid = threadIdx.x;
    const double il = (xl && shift != 0.0) ? pw_inv_norm(xl_norm2, eps) : 0.0;
    double sr = 0.0, sl = 0.0;
    for (int c = tid; c < n; c += 1024) {
        if (yl) {
            double v = 0.0;
            for (int t = 0; t < tiles; ++t) v += partial[(size_t)t * n + c];
            if (shift != 0.0) v = fma(shift, xl[c] * il, v);
            yl[c] = v;
            sl = fma(v, v, sl);
        }
        if (yr) sr = fma(yr[c], yr[c], sr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if ((tid & 31) == 0) { s_a[tid >> 5] = sr; s_b[tid >> 5] = sl; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) { a += s_a[w]; b += s_b[w]; }
        norms2[0] = a;
        norms2[1] = b;
    }
}

// x <- x / (sqrt(norm2) + eps)
__global__ void __launch_bounds__(1024) vec_finish_kernel(int n, double* __restrict__ x, const double* __restrict__ norm2,
                                                          double eps) {
    const double inv = pw_inv_norm(norm2, eps);
    for (int c = blockIdx.x * 1024 + threadIdx.x; c < n; c += gridDim.x * 1024) x[c] *= inv;
}

__global__ void __launch_bounds__(1024) vec_dots_kernel(int n, const double* __restrict__ x, const double* __restrict__ y,
                                                        double* __restrict__ out4) {
    __shared__ double s[4][32];
    const int tid = threadIdx.x;
    double a = 0.0, b = 0.0, c = 0.0, e = 0.0;
    for (int i = tid; i < n; i += 1024) {
        const double xv = x[i], yv = y[i], df = xv - yv;
        a = fma(xv, yv, a);
        b = fma(xv, xv, b);
        c = fma(yv, yv, c);
        e = fma(df, df, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        e += __shfl_xor_sync(0xffffffffu, e, o);
    }
    if ((tid & 31) == 0) { s[0][tid >> 5] = a; s[1][tid >> 5] = b; s[2][tid >> 5] = c; s[3][tid >> 5] = e; }
    __syncthreads();
    if (tid < 4) {
        double v = 0.0;
        for (int w = 0; w < 32; ++w) v += s[tid][w];
        out4[tid] = v;
    }
}

// out (+)= num / (den[0] + eps) * 2 W o (a1 b1^T + a2 b2^T)  -- the fold-back of a rank-one d rho / dA = u v^T / (u.v)
// onto W through the blocks A11 = W o W and A22 = (W o W)^T (notreks.py:278-287, 343-344), without ever forming the
// 2d x 2d outer product; the same kernel carries the Rayleigh baseline (u1 u1^T + u2 u2^T) / (u.u) of :365-371.
__global__ void __launch_bounds__(256) rank2_fold_kernel(int d, const double* __restrict__ W, const double* __restrict__ a1,
                                                         const double* __restrict__ b1, const double* __restrict__ a2,
                                                         const double* __restrict__ b2, double num,
                                                         const double* __restrict__ den, double eps, int accumulate,
                                                         double* __restrict__ out) {
    const double scale = num / (den[0] + eps);
    const size_t total = (size_t)d * d;
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const int i = (int)(idx / d), j = (int)(idx - (size_t)i * d);
        double g = a1[i] * b1[j];
        if (a2) g = fma(a2[i], b2[j], g);
        const double v = W ? scale * 2.0 * W[idx] * g : scale * g;
        out[idx] = accumulate ? out[idx] + v : v;
    }
}

__global__ void fill_ones_kernel(int n, double* a, double* b) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) { a[i] = 1.0; b[i] = 1.0; }
}

// scal = [rho = v.Av / (v.v + eps), u.v, u.u, v.v]
__global__ void perron_scalars_kernel(const double* d3, double eps, double* scal) {
    scal[0] = d3[0] / (d3[1] + eps);
    scal[1] = d3[4];
    scal[2] = d3[5];
    scal[3] = d3[1];
}

}  // namespace dagma

using namespace dagma;

extern "C" size_t dagma_power_workspace_bytes(int n) {
    const size_t tiles = (size_t)(n + PW_ROWS - 1) / PW_ROWS;
    return (tiles * (size_t)n + 4 * (size_t)n + 16) * sizeof(double);
}

// n_iter steps of  v <- A v / (||A v|| + eps)  and  u <- A^T u / (||A^T u|| + eps)  from the all-ones vectors
// (notreks.py:178-192), both chains in one pass over A per step.  start != 0: continue from the given v_dev / u_dev
// (unit vectors) instead of ones.  Outputs: v_dev, u_dev (normalised), scal_dev[0..3] = rho = v.(A v) / (v.v + eps),
// u.v, u.u, v.v.  shift: iterate with A + shift I (same Perron vectors, no periodicity) -- rho is of A itself.
extern "C" int dagma_perron_power_f64(dagma_stream_t stream_, int n, const double* a_dev, int lda, int square,
                                      double shift, int n_iter, double eps, int start, double* v_dev, double* u_dev,
                                      double* scal_dev, double* ws_dev, size_t ws_bytes) {
    DAGMA_REQUIRE(n >= 1 && lda >= n && n_iter >= 0, "bad shape");
    DAGMA_REQUIRE(a_dev && v_dev && u_dev && scal_dev && ws_dev, "null device pointer");
    DAGMA_REQUIRE(ws_bytes >= dagma_power_workspace_bytes(n), "workspace too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int tiles = (n + PW_ROWS - 1) / PW_ROWS;
    double* partial = ws_dev;
    double* buf[2][2] = {{partial + (size_t)tiles * n, partial + (size_t)tiles * n + n},
                         {partial + (size_t)tiles * n + 2 * (size_t)n, partial + (size_t)tiles * n + 3 * (size_t)n}};
    double* norms = partial + (size_t)tiles * n + 4 * (size_t)n;      // [2 buffers][2] squared norms, then scratch
    // iteration k reads vectors (xr, xl) with norms `nin` and writes buf[k & 1] / norms + 2 (k & 1)
    const double* xr = v_dev;
    const double* xl = u_dev;
    const double* nin = nullptr;
    if (!start) {
        fill_ones_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, v_dev, u_dev);
    }
    for (int k = 0; k < n_iter; ++k) {
        double* yr = buf[k & 1][0];
        double* yl = buf[k & 1][1];
        double* nout = norms + 2 * (k & 1);
        power_pass_kernel<<<tiles, PW_NT, 0, stream>>>(n, a_dev, lda, square, shift, xr, nin, xl, nin ? nin + 1 : nullptr,
                                                       eps, yr, partial);
        power_reduce_kernel<<<1, 1024, 0, stream>>>(n, tiles, partial, shift, xl, nin ? nin + 1 : nullptr, eps, yr, yl,
                                                    nout);
        xr = yr; xl = yl; nin = nout;
    }
    if (n_iter > 0) {
        DAGMA_CUDA_OK(cudaMemcpyAsync(v_dev, xr, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        DAGMA_CUDA_OK(cudaMemcpyAsync(u_dev, xl, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        const int g = (n + 1023) / 1024;
        vec_finish_kernel<<<g, 1024, 0, stream>>>(n, v_dev, nin, eps);
        vec_finish_kernel<<<g, 1024, 0, stream>>>(n, u_dev, nin + 1, eps);
    }
    // rho = v.(A v) / (v.v + eps)  (notreks.py:194), u.v for the gradient's denominator (:236)
    double* av = buf[n_iter & 1][0];
    double* d3 = norms + 8;                                               // two blocks of 4
    power_pass_kernel<<<tiles, PW_NT, 0, stream>>>(n, a_dev, lda, square, 0.0, v_dev, nullptr, nullptr, nullptr, eps, av,
                                                   nullptr);
    vec_dots_kernel<<<1, 1024, 0, stream>>>(n, v_dev, av, d3);            // [v.Av, v.v, .]
    vec_dots_kernel<<<1, 1024, 0, stream>>>(n, u_dev, v_dev, d3 + 4);     // [u.v, u.u, v.v]
    perron_scalars_kernel<<<1, 1, 0, stream>>>(d3, eps, scal_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_matvec_f64(dagma_stream_t stream_, int n, const double* a_dev, int lda, int square,
                                const double* x_dev, double* y_dev) {
    DAGMA_REQUIRE(n >= 1 && lda >= n && a_dev && x_dev && y_dev, "bad arguments");
    const int tiles = (n + PW_ROWS - 1) / PW_ROWS;
    power_pass_kernel<<<tiles, PW_NT, 0, (cudaStream_t)stream_>>>(n, a_dev, lda, square, 0.0, x_dev, nullptr, nullptr,
                                                                  nullptr, 0.0, y_dev, nullptr);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_vec_dots_f64(dagma_stream_t stream_, int n, const double* x_dev, const double* y_dev,
                                  double* out4_dev) {
    DAGMA_REQUIRE(n >= 1 && x_dev && y_dev && out4_dev, "bad arguments");
    vec_dots_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(n, x_dev, y_dev, out4_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_rank2_fold_f64(dagma_stream_t stream_, int d, const double* w_dev, const double* a1_dev,
                                    const double* b1_dev, const double* a2_dev, const double* b2_dev, double num,
                                    const double* den_dev, double eps, int accumulate, double* out_dev) {
    DAGMA_REQUIRE(d >= 1 && a1_dev && b1_dev && den_dev && out_dev, "bad arguments");
    DAGMA_REQUIRE((a2_dev == nullptr) == (b2_dev == nullptr), "a2 / b2 come together");
    const size_t total = (size_t)d * d;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    rank2_fold_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(d, w_dev, a1_dev, b1_dev, a2_dev, b2_dev, num, den_dev,
                                                               eps, accumulate, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
