// Pairwise independence tests of the fork's pre-processing step (reference:
// src/notreks/mi_tests.py:19-135, 165-203, CR-delimited line numbers): HSIC with RBF kernels and distance
// correlation, each with a permutation p-value.
//
// The reference rebuilds both n x n Gram matrices (exp, median heuristic, double centring) for EVERY
// permutation of every pair: O(pairs * perms * n^2) transcendental work.  A permutation only relabels the
// second variable, L'[a][b] = L[pi(a)][pi(b)], and centring commutes with the relabelling, so here every
// variable's centred Gram matrix is built ONCE (d * n^2 doubles in HBM) and a permuted statistic is the
// gathered dot product
//     sum_ab Kc_i[a][b] * Lc_j[pi(a)][pi(b)]
// -- HBM / L2-bound FP64 work: one coalesced read of Kc_i and one gathered read of Lc_j (rows of Lc_j are
// 8n bytes, L1/L2 resident) per (pair, permutation).  Grid = (row chunks, permutations, pairs); partial sums
// are reduced in a fixed order (no atomics), so statistics and p-values are reproducible.
#include "common.cuh"
#include "small_gj.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

constexpr int MI_ROWS = 16;      // rows of Kc per block of the gathered dot product

// G[v][a][b] = exp(-(x_a - x_b)^2 / (2 sigma2_v))  (kind 0, mi_tests.py:31-50)  or  |x_a - x_b|  (kind 1, :87-88);
// X is n x d row-major, column cols[v]
__global__ void __launch_bounds__(256) mi_gram_kernel(const double* __restrict__ X, int n, int d,
                                                      const int* __restrict__ cols, const double* __restrict__ sigma2,
                                                      int kind, double* __restrict__ G) {
    const int v = blockIdx.z, a = blockIdx.y;
    const int c = cols[v];
    const double xa = X[(size_t)a * d + c];
    const double inv = kind == 0 ? 1.0 / (2.0 * sigma2[v]) : 0.0;
    double* row = G + ((size_t)v * n + a) * n;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
        const double df = xa - X[(size_t)b * d + c];
        row[b] = kind == 0 ? exp(-(df * df) * inv) : fabs(df);
    }
}
// squared pairwise distances of the strict upper triangle (input of the median heuristic, mi_tests.py:41-43)
__global__ void __launch_bounds__(256) mi_upper_d2_kernel(const double* __restrict__ X, int n, int d, int c,
                                                          double* __restrict__ out) {
    const int a = blockIdx.y;
    const double xa = X[(size_t)a * d + c];
    const size_t base = (size_t)a * n - (size_t)a * (a + 1) / 2 - a - 1;     // index of (a, a + 1) minus (a + 1)
    for (int b = a + 1 + blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
        const double df = xa - X[(size_t)b * d + c];
        out[base + b] = df * df;
    }
}
// row means of G[v] (one block per row)
__global__ void __launch_bounds__(256) mi_row_mean_kernel(const double* __restrict__ G, int n, double* __restrict__ rm) {
    __shared__ double red[96];
    const double* row = G + (size_t)blockIdx.x * n;      // blockIdx.x = v * n + a
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    for (int b = threadIdx.x; b < n; b += 256) acc += row[b];
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) rm[blockIdx.x] = acc / (double)n;
}
// all-mean per variable from the row means; G <- G - rm[a] - rm[b] + all (the matrices are symmetric, so the
// column means of mi_tests.py:22-27 are the row means)
__global__ void __launch_bounds__(256) mi_all_mean_kernel(const double* __restrict__ rm, int n, double* __restrict__ am) {
    __shared__ double red[96];
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    for (int a = threadIdx.x; a < n; a += 256) acc += rm[(size_t)blockIdx.x * n + a];
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) am[blockIdx.x] = acc / (double)n;
}
__global__ void __launch_bounds__(256) mi_center_kernel(double* __restrict__ G, int n, const double* __restrict__ rm,
                                                        const double* __restrict__ am) {
    const int v = blockIdx.z, a = blockIdx.y;
    const double ra = rm[(size_t)v * n + a], all = am[v];
    double* row = G + ((size_t)v * n + a) * n;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x)
        row[b] = row[b] - ra - rm[(size_t)v * n + b] + all;
}
// partial[(q * P + p) * chunks + chunk] = sum over the chunk's rows a and all b of K_{iq}[a][b] L_{jq}[pi(a)][pi(b)]
//   pair q = blockIdx.z: variables vi[q], vj[q] (indices into G); permutation p = blockIdx.y of that pair
__global__ void __launch_bounds__(256) mi_perm_dot_kernel(const double* __restrict__ G, int n,
                                                          const int* __restrict__ vi, const int* __restrict__ vj,
                                                          const int* __restrict__ perms, int P,
                                                          double* __restrict__ partial) {
    extern __shared__ int pi_s[];
    __shared__ double red[96];
    const int q = blockIdx.z, p = blockIdx.y, chunk = blockIdx.x;
    const int* pi = perms + ((size_t)q * P + p) * n;
    for (int e = threadIdx.x; e < n; e += 256) pi_s[e] = pi[e];
    __syncthreads();
    const double* K = G + (size_t)vi[q] * n * n;
    const double* L = G + (size_t)vj[q] * n * n;
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    const int a1 = min(n, (chunk + 1) * MI_ROWS);
    for (int a = chunk * MI_ROWS; a < a1; ++a) {
        const double* kr = K + (size_t)a * n;
        const double* lr = L + (size_t)pi_s[a] * n;
        for (int b = threadIdx.x; b < n; b += 256) acc = fma(kr[b], lr[pi_s[b]], acc);
    }
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) partial[((size_t)q * P + p) * gridDim.x + chunk] = acc;
}
// out[q * P + p] = scale * sum over chunks (fixed order)
__global__ void mi_perm_finish_kernel(const double* __restrict__ partial, int chunks, size_t total, double scale,
                                      double* __restrict__ out) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= total) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[e * chunks + c];
    out[e] = s * scale;
}

}  // namespace dagma

using namespace dagma;

extern "C" int dagma_mi_upper_d2_f64(dagma_stream_t stream, int n, int d, int col, const double* x_dev, double* out_dev) {
    DAGMA_REQUIRE(x_dev && out_dev && n >= 2 && col >= 0 && col < d, "bad arguments");
    dim3 grid((n + 255) / 256, n - 1);
    mi_upper_d2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, col, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mi_centered_gram_f64(dagma_stream_t stream, int n, int d, int nvars, const int* cols_dev,
                                          const double* x_dev, const double* sigma2_dev, int kind, double* g_dev,
                                          double* rowmean_dev, double* allmean_dev) {
    DAGMA_REQUIRE(x_dev && cols_dev && g_dev && rowmean_dev && allmean_dev && n >= 1 && nvars >= 1, "bad arguments");
    DAGMA_REQUIRE(kind == 1 || sigma2_dev, "the RBF kernel needs sigma^2 per variable");
    cudaStream_t st = (cudaStream_t)stream;
    const int bx = (n + 255) / 256 < 8 ? (n + 255) / 256 : 8;
    dim3 grid(bx, n, nvars);
    mi_gram_kernel<<<grid, 256, 0, st>>>(x_dev, n, d, cols_dev, sigma2_dev, kind, g_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    mi_row_mean_kernel<<<nvars * n, 256, 0, st>>>(g_dev, n, rowmean_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    mi_all_mean_kernel<<<nvars, 256, 0, st>>>(rowmean_dev, n, allmean_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    mi_center_kernel<<<grid, 256, 0, st>>>(g_dev, n, rowmean_dev, allmean_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" size_t dagma_mi_perm_workspace_bytes(int n, int npairs, int nperm) {
    return (size_t)npairs * nperm * ((n + MI_ROWS - 1) / MI_ROWS) * sizeof(double);
}

extern "C" int dagma_mi_perm_dots_f64(dagma_stream_t stream, int n, const double* g_dev, int npairs,
                                      const int* vi_dev, const int* vj_dev, const int* perms_dev, int nperm,
                                      double scale, double* ws_dev, size_t ws_bytes, double* out_dev) {
    DAGMA_REQUIRE(g_dev && vi_dev && vj_dev && perms_dev && ws_dev && out_dev && n >= 1 && npairs >= 1 && nperm >= 1,
                  "bad arguments");
    DAGMA_REQUIRE(ws_bytes >= dagma_mi_perm_workspace_bytes(n, npairs, nperm), "workspace too small");
    DAGMA_REQUIRE(nperm <= 65535 && npairs <= 65535, "at most 65535 permutations / pairs per call");
    DAGMA_REQUIRE((size_t)n * sizeof(int) <= 200 * 1024, "n too large for the shared-memory permutation");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (n + MI_ROWS - 1) / MI_ROWS;
    const size_t smem = (size_t)n * sizeof(int);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        DAGMA_CUDA_OK(cudaFuncSetAttribute(mi_perm_dot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    mi_perm_dot_kernel<<<dim3(chunks, nperm, npairs), 256, smem, st>>>(g_dev, n, vi_dev, vj_dev, perms_dev, nperm, ws_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    const size_t total = (size_t)npairs * nperm;
    mi_perm_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws_dev, chunks, total, scale, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
