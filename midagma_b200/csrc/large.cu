// Multi-CTA path for d > 64 and for the logistic loss (SURVEY.md 8a rows a1-a5 at
// C2 / C5 sizes): FP64 DMMA GEMM, blocked Gauss-Jordan inverse, score reductions and
// the fused Adam / feasibility update, all driven by a small device-resident state
// block so that a whole inner iteration is a fixed launch sequence (CUDA-graph
// replayable, no host synchronisation until the next convergence checkpoint).
//
// Blocked inverse (d > 128), block size 64, in place, no pivoting (M-matrix):
//   step kb:  P = A[K,K];  Q = P^{-1} (on-chip sweep, redundantly in every panel CTA);
//             CS = -Cpub Q  (d x 64)  with Cpub = A[:,K] and its K block replaced by P - I;
//             Rpub = A[K,:]           with its K block replaced by I + P;
//             A += CS * Rpub          (one d x d x 64 DMMA GEMM, no special tiles).
// The two replacements make the single GEMM produce Q in the pivot block, Q A[K,:] in
// the pivot rows and -A[:,K] Q in the pivot columns (block form of the identity used
// by the on-chip sweep, see small_gj.cuh).  log|det| = sum of the logs of all pivots.
#include "common.cuh"
#include "gemm_f64.cuh"
#include "small_gj.cuh"
#include "../../include/dagma_b200.h"

namespace dagma {

int logdet_inv_small(cudaStream_t, int, int, double, const double*, int, int, double*, double*,
                     double*, double*, int, double*, int*);

// ------------------------------------------------------------------ GEMM host side
static int gemm_launch(cudaStream_t stream, int transA, int M, int N, int K, double alpha, const double* A, int lda,
                       const double* B, int ldb, double beta, double* C, int ldc, int epi, double* ws,
                       size_t ws_bytes) {
    if (M <= 0 || N <= 0) return 0;
    static bool attr_done = false;
    if (!attr_done) {
        DAGMA_CUDA_OK(cudaFuncSetAttribute(gemm_f64_kernel<false, EPI_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
        DAGMA_CUDA_OK(cudaFuncSetAttribute(gemm_f64_kernel<false, EPI_SIGMOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
        DAGMA_CUDA_OK(cudaFuncSetAttribute(gemm_f64_kernel<true, EPI_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
        DAGMA_CUDA_OK(cudaFuncSetAttribute(gemm_f64_kernel<true, EPI_SIGMOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
        attr_done = true;
    }
    const int tm = (M + GBM - 1) / GBM, tn = (N + GBN - 1) / GBN;
    const int ktiles = (K + GBK - 1) / GBK;
    // split K when the output grid cannot fill the machine and K is long
    int splits = 1;
    if (tm * tn < 74 && ktiles >= 8 && ws != nullptr) {
        splits = 148 / (tm * tn);
        if (splits > ktiles / 4) splits = ktiles / 4;
        while (splits > 1 && (size_t)splits * M * N * sizeof(double) > ws_bytes) --splits;
        if (splits < 1) splits = 1;
    }
    int kchunk_tiles = (ktiles + splits - 1) / splits;
    splits = (ktiles + kchunk_tiles - 1) / kchunk_tiles;
    GemmArgs P{M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, kchunk_tiles * GBK, splits > 1 ? ws : nullptr};
    dim3 grid(tn, tm, splits);
    if (splits > 1) {   // epilogue is applied by the reduce kernel
        if (transA) gemm_f64_kernel<true, EPI_NONE><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
        else gemm_f64_kernel<false, EPI_NONE><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
        DAGMA_CUDA_OK(cudaGetLastError());
        const int blocks = (int)(((size_t)M * N + 255) / 256);
        if (epi == EPI_SIGMOID)
            splitk_reduce_kernel<EPI_SIGMOID><<<blocks < 1184 ? blocks : 1184, 256, 0, stream>>>(ws, splits, M, N, C, ldc, alpha, beta);
        else
            splitk_reduce_kernel<EPI_NONE><<<blocks < 1184 ? blocks : 1184, 256, 0, stream>>>(ws, splits, M, N, C, ldc, alpha, beta);
    } else if (epi == EPI_SIGMOID) {
        if (transA) gemm_f64_kernel<true, EPI_SIGMOID><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
        else gemm_f64_kernel<false, EPI_SIGMOID><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
    } else {
        if (transA) gemm_f64_kernel<true, EPI_NONE><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
        else gemm_f64_kernel<false, EPI_NONE><<<grid, GTHREADS, GEMM_SMEM_BYTES, stream>>>(P);
    }
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ blocked inverse
constexpr int NB = 64;
using CP = Cfg<4, 2, 16, 32>;     // panel CTA: 512 threads, 64 x 64 on chip

// build M = s I - (square ? A o A : A), scaled by inv_scale, into out (d x d, ld = d)
__global__ void build_m_kernel(const double* __restrict__ A, int lda, double* __restrict__ out, int d, double s,
                               double inv_scale, int square) {
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / d), c = (int)(e - (size_t)r * d);
        double x = A[(size_t)r * lda + c];
        if (square) x *= x;
        out[e] = (((r == c) ? s : 0.0) - x) * inv_scale;
    }
}

constexpr int PANEL_LINE = ((SweepSmem<CP>::doubles + 3) / 2) * 2;
constexpr size_t PANEL_SMEM_BYTES = (size_t)(PANEL_LINE + 2 * NB * (NB + 2)) * sizeof(double);

struct PanelArgs {
    double* A;      // d x d in place (ld = d)
    int d, kb;
    double* CS;     // d x NB  (ld = NB)
    double* Rbuf;   // NB x d  (ld = d)
    double* pivots; // d
};

__global__ void __launch_bounds__(CP::NT, 1) inv_panel_kernel(const PanelArgs P) {
    constexpr int RM = CP::RM, RN = CP::RN, NT = CP::NT, LQ = NB + 2;
    extern __shared__ __align__(16) double psm[];
    double* linebuf = psm;
    double* Qs = psm + PANEL_LINE;
    double* Cs = Qs + NB * LQ;
    const int tid = threadIdx.x;
    const ThreadPos<CP> pos(tid);
    const int ty = pos.ty, tx = pos.tx;
    const int d = P.d, k0 = P.kb * NB;
    const int kn = min(NB, d - k0);
    const int blk = blockIdx.x;                 // row block (for CS) and column block (for Rbuf)
    const double* pivots_s = linebuf + SweepSmem<CP>::piv_off;

    // ---- Q = P^{-1} on chip (every CTA redundantly; P is 32 KB and L2 resident)
    double a[RM][RN], dummy[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int r = CP::grow(ty, i), c = CP::gcol(tx, j);
            a[i][j] = (r < kn && c < kn) ? P.A[(size_t)(k0 + r) * d + k0 + c] : ((r == c) ? 1.0 : 0.0);
            dummy[i][j] = 0.0;
        }
    // stash P - I (for Cpub) before the sweep destroys it
    if (blk == P.kb) {
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) {
                const int r = CP::grow(ty, i), c = CP::gcol(tx, j);
                Cs[r * LQ + c] = (r < kn && c < kn) ? a[i][j] - ((r == c) ? 1.0 : 0.0) : 0.0;
            }
    }
    gj_sweep<CP, false>(a, dummy, 0u, 0u, smem_u32(linebuf), kn, ty, tx);
    if (blk == 0 && tid < kn) P.pivots[k0 + tid] = pivots_s[tid];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int r = CP::grow(ty, i), c = CP::gcol(tx, j);
            Qs[r * LQ + c] = (r < kn && c < kn) ? a[i][j] : 0.0;
        }

    // ---- Cpub rows of this block: A[I, K]  (K block: P - I, stashed above)
    const int r0 = blk * NB;
    if (blk != P.kb) {
        for (int e = tid; e < NB * NB; e += NT) {
            const int r = e / NB, c = e - r * NB;
            Cs[r * LQ + c] = (r0 + r < d && c < kn) ? P.A[(size_t)(r0 + r) * d + k0 + c] : 0.0;
        }
    }
    // ---- Rpub columns of this block: A[K, J]  (K block: I + P)
    for (int e = tid; e < NB * NB; e += NT) {
        const int r = e / NB, c = e - r * NB;
        if (r < kn && r0 + c < d) {
            double v = P.A[(size_t)(k0 + r) * d + r0 + c];
            if (blk == P.kb && r == c) v += 1.0;
            P.Rbuf[(size_t)r * d + r0 + c] = v;
        }
    }
    __syncthreads();
    // ---- CS[I, :] = -Cpub Q   (64 x 64 x kn on chip)
    double acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.0;
    for (int k = 0; k < kn; ++k) {
        double cv[RM], qv[RN];
#pragma unroll
        for (int i = 0; i < RM; ++i) cv[i] = Cs[CP::grow(ty, i) * LQ + k];
#pragma unroll
        for (int j = 0; j < RN; ++j) qv[j] = Qs[k * LQ + CP::gcol(tx, j)];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) acc[i][j] = fma(-cv[i], qv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int r = r0 + CP::grow(ty, i), c = CP::gcol(tx, j);
            if (r < d && c < kn) P.CS[(size_t)r * NB + c] = acc[i][j];
        }
}

// after the sweep: log|det|, min entry, info; optional outputs (scaled back)
__global__ void inv_finish_kernel(const double* __restrict__ Minv, const double* __restrict__ pivots, int d,
                                  double s, double inv_scale, double log_scale, double* logabsdet, double* h,
                                  double* min_entry, int* info) {
    __shared__ double red[96];
    const int tid = threadIdx.x;
    double ld = 0.0, z1 = 0.0, z2 = 0.0;
    bool bad = false;
    for (int k = tid; k < d; k += blockDim.x) {
        const double p = pivots[k];
        ld += log(fabs(p));
        bad |= !(p > 0.0);
    }
    double mn = INFINITY;
    const size_t total = (size_t)d * d;
    for (size_t e = tid; e < total; e += blockDim.x) mn = fmin(mn, Minv[e]);
    block_sum3<1024>(ld, z1, z2, red, tid);
    mn = block_min<1024>(mn, red, tid) * inv_scale;
    const int anybad = __syncthreads_or(bad);
    if (tid == 0) {
        const double lad = ld + (double)d * log_scale;
        if (logabsdet) *logabsdet = lad;
        if (h) *h = -lad + (double)d * log(s);
        if (min_entry) *min_entry = mn;
        if (info) *info = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
    }
}

// out = scale * Minv  and/or  grad = (square ? 2 A^T-indexed : 1) * Minv^T
__global__ void inv_outputs_kernel(const double* __restrict__ Minv, const double* __restrict__ A, int lda, int d,
                                   double inv_scale, int square, double* minv_out, double* grad_out, int ldo) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int x = threadIdx.x, y0 = threadIdx.y;     // 32 x 8
    for (int y = y0; y < 32; y += 8) {
        const int r = by + y, c = bx + x;
        double v = (r < d && c < d) ? Minv[(size_t)r * d + c] * inv_scale : 0.0;
        tile[y][x] = v;
        if (minv_out && r < d && c < d) minv_out[(size_t)r * ldo + c] = v;
    }
    __syncthreads();
    if (grad_out) {
        for (int y = y0; y < 32; y += 8) {
            const int r = bx + y, c = by + x;       // transposed position
            if (r < d && c < d) {
                const double f = square ? 2.0 * A[(size_t)r * lda + c] : 1.0;
                grad_out[(size_t)r * ldo + c] = f * tile[x][y];
            }
        }
    }
}

static size_t large_ws_bytes(int d) {
    return ((size_t)d * d + (size_t)d * NB + (size_t)NB * d + d + 64) * sizeof(double);
}

// one problem; ws holds: Mwork (d*d) | CS (d*NB) | Rbuf (NB*d) | pivots (d)
static int logdet_inv_blocked(cudaStream_t stream, int d, double s, const double* a_dev, int lda, int square,
                              double* logabsdet, double* h, double* minv, double* grad, int ldo,
                              double* min_entry, int* info, double* ws) {
    double scale = 1.0;
    if (s > 0.0 && isfinite(s)) {
        int e = 0;
        const double f = frexp(s, &e);
        if (f == 0.5) --e;
        scale = ldexp(1.0, e);
    }
    const double inv_scale = 1.0 / scale;
    // work in place in the caller's minv buffer when it is dense (ldo == d), else in ws
    double* Mw = (minv != nullptr && ldo == d) ? minv : ws;
    double* CS = ws + (size_t)d * d;
    double* Rbuf = CS + (size_t)d * NB;
    double* piv = Rbuf + (size_t)NB * d;
    static bool panel_attr = false;
    if (!panel_attr) {
        DAGMA_CUDA_OK(cudaFuncSetAttribute(inv_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM_BYTES));
        panel_attr = true;
    }
    build_m_kernel<<<592, 256, 0, stream>>>(a_dev, lda, Mw, d, s, inv_scale, square);
    DAGMA_CUDA_OK(cudaGetLastError());
    const int nblk = (d + NB - 1) / NB;
    for (int kb = 0; kb < nblk; ++kb) {
        PanelArgs P{Mw, d, kb, CS, Rbuf, piv};
        inv_panel_kernel<<<nblk, CP::NT, PANEL_SMEM_BYTES, stream>>>(P);
        DAGMA_CUDA_OK(cudaGetLastError());
        const int kn = (d - kb * NB) < NB ? (d - kb * NB) : NB;
        int rc = gemm_launch(stream, 0, d, d, kn, 1.0, CS, NB, Rbuf, d, 1.0, Mw, d, EPI_NONE, nullptr, 0);
        if (rc) return rc;
    }
    inv_finish_kernel<<<1, 1024, 0, stream>>>(Mw, piv, d, s, inv_scale, log(scale), logabsdet, h, min_entry, info);
    DAGMA_CUDA_OK(cudaGetLastError());
    if (grad != nullptr || (minv != nullptr && (minv != Mw || inv_scale != 1.0))) {
        dim3 grid((d + 31) / 32, (d + 31) / 32);
        // note: when minv == Mw and a rescale is needed the tile kernel rewrites it in place
        inv_outputs_kernel<<<grid, dim3(32, 8), 0, stream>>>(Mw, a_dev, lda, d, inv_scale, square,
                                                             (minv == Mw && inv_scale == 1.0) ? nullptr : minv, grad, ldo);
        DAGMA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int logdet_inv_large(cudaStream_t stream, int batch, int d, double s, const double* a_dev, int lda, int square,
                     double* logabsdet, double* h, double* minv, double* grad, int ldo, double* min_entry,
                     int* info) {
    double* ws = nullptr;
    DAGMA_CUDA_OK(cudaMallocAsync((void**)&ws, large_ws_bytes(d), stream));
    int rc = 0;
    for (int b = 0; b < batch && rc == 0; ++b) {
        rc = logdet_inv_blocked(stream, d, s, a_dev + (size_t)b * d * lda, lda, square,
                                logabsdet ? logabsdet + b : nullptr, h ? h + b : nullptr,
                                minv ? minv + (size_t)b * d * ldo : nullptr, grad ? grad + (size_t)b * d * ldo : nullptr,
                                ldo, min_entry ? min_entry + b : nullptr, info ? info + b : nullptr, ws);
    }
    cudaFreeAsync(ws, stream);
    return rc;
}

// ------------------------------------------------------------------ iteration state
struct LinState {           // mirrored by midagma_b200/_large.py (all 8-byte fields first)
    double mu, s, lr, lambda1, beta1, beta2;
    double p1_hi, p1_lo, p2_hi, p2_lo;      // beta^it as double-double
    double logabsdet, h, min_entry;
    double score_acc, l1_acc, loss_acc;     // reduction outputs
    double gscale;                          // l2: 1 (T = cov W), logistic: 1/n (T = X^T sigmoid(XW))
    int32_t it, halted, info, pad;
};

__device__ __forceinline__ void dd_mul(double& hi, double& lo, double b) {
    const double ph = hi * b;
    const double pl = fma(hi, b, -ph) + lo * b;
    const double s = ph + pl;
    lo = pl - (s - ph);
    hi = s;
}

// Gobj, Adam, step, masks for one inner iteration (linear.py:248, 158-162, 275-276).
// No-op (and latches `halted`) when the inverse of this iteration was infeasible.
__global__ void __launch_bounds__(256) linear_update_kernel(LinState* st, int d, double* __restrict__ W,
                                                             const double* __restrict__ Minv,
                                                             const double* __restrict__ T,
                                                             const double* __restrict__ cov, double* __restrict__ m,
                                                             double* __restrict__ v, const uint8_t* mask_exc,
                                                             const uint8_t* mask_inc) {
    __shared__ double tile[32][33];
    const bool stop = (st->halted != 0) || (st->info != 0);
    if (stop) return;                        // latching is done by linear_advance_kernel
    const double mu = st->mu, lr = st->lr, lambda1 = st->lambda1, b1 = st->beta1, b2 = st->beta2;
    double p1h = st->p1_hi, p1l = st->p1_lo, p2h = st->p2_hi, p2l = st->p2_lo;
    dd_mul(p1h, p1l, b1);
    dd_mul(p2h, p2l, b2);
    const double c1 = 1.0 / ((1.0 - p1h) - p1l), c2 = 1.0 / ((1.0 - p2h) - p2l);
    const double gscale = st->gscale;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int x = threadIdx.x, y0 = threadIdx.y;
    // tile of Minv at the transposed block position
    for (int y = y0; y < 32; y += 8) {
        const int r = bx + y, c = by + x;
        tile[y][x] = (r < d && c < d) ? Minv[(size_t)r * d + c] : 0.0;
    }
    __syncthreads();
    for (int y = y0; y < 32; y += 8) {
        const int r = by + y, c = bx + x;
        if (r >= d || c >= d) continue;
        const size_t e = (size_t)r * d + c;
        const double w = W[e];
        const double minvT = tile[x][y];
        const double sg = (w > 0.0) ? 1.0 : ((w < 0.0) ? -1.0 : 0.0);
        const double gsc = fma(gscale, T[e], -cov[e]);          // l2: cov W - cov ; logistic: X^T sigma / n - cov
        double go = fma(mu, gsc, mu * lambda1 * sg);
        go = fma(2.0 * w, minvT + 1e-16, go);
        if (mask_inc && mask_inc[e]) go = fma(-2.0 * mu * lambda1, sg, go);
        const double mn = fma(m[e], b1, (1.0 - b1) * go);
        const double vn = fma(v[e], b2, (1.0 - b2) * (go * go));
        m[e] = mn;
        v[e] = vn;
        const double dir = fast_div(mn * c1, fast_sqrt_nonneg(vn * c2) + 1e-8);
        double wn = w - lr * dir;
        if (mask_exc && mask_exc[e]) wn = 0.0;
        W[e] = wn;
    }
}

__global__ void linear_advance_kernel(LinState* st) {
    if (st->halted != 0) return;
    if (st->info != 0) {
        st->halted = 1;
        return;
    }
    double p1h = st->p1_hi, p1l = st->p1_lo, p2h = st->p2_hi, p2l = st->p2_lo;
    dd_mul(p1h, p1l, st->beta1);
    dd_mul(p2h, p2l, st->beta2);
    st->p1_hi = p1h; st->p1_lo = p1l; st->p2_hi = p2h; st->p2_lo = p2l;
    st->it += 1;
}

// W += sign * lr * dir(m, v, it)   (back-tracking, linear.py:235, 239)
__global__ void linear_apply_dir_kernel(const LinState* st, int d, double* __restrict__ W,
                                        const double* __restrict__ m, const double* __restrict__ v, double sign) {
    const double c1 = 1.0 / ((1.0 - st->p1_hi) - st->p1_lo), c2 = 1.0 / ((1.0 - st->p2_hi) - st->p2_lo);
    const double scale = sign * st->lr;
    const size_t total = (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const double dir = fast_div(m[e] * c1, fast_sqrt_nonneg(v[e] * c2) + 1e-8);
        W[e] = __dadd_rn(W[e], __dmul_rn(scale, dir));
    }
}

// l2: score_acc = 1/2 sum (I - W) o (cov - T), l1_acc = sum |W|   (T = cov W)
// logistic: only l1_acc (the loss comes from logistic_loss_kernel)
__global__ void __launch_bounds__(1024) linear_objective_kernel(LinState* st, int d, const double* __restrict__ W,
                                                                 const double* __restrict__ T,
                                                                 const double* __restrict__ cov, int l2) {
    __shared__ double red[96];
    const int tid = threadIdx.x;
    double sc = 0.0, l1 = 0.0, z = 0.0;
    const size_t total = (size_t)d * d;
    for (size_t e = tid; e < total; e += blockDim.x) {
        const int r = (int)(e / d), c = (int)(e - (size_t)r * d);
        const double w = W[e];
        l1 += fabs(w);
        if (l2) sc = fma(((r == c) ? 1.0 : 0.0) - w, cov[e] - T[e], sc);
    }
    block_sum3<1024>(sc, l1, z, red, tid);
    if (tid == 0) {
        st->score_acc = 0.5 * sc;
        st->l1_acc = l1;
    }
}

// partial[b] = sum over a slice of (logaddexp(0, R) - X o R)   (linear.py:91); finished by a 1-block pass
__global__ void __launch_bounds__(256) logistic_loss_kernel(const double* __restrict__ X, const double* __restrict__ R,
                                                             size_t total, double* partial) {
    __shared__ double red[96];
    double acc = 0.0, z1 = 0.0, z2 = 0.0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const double r = R[e];
        const double sp = fmax(r, 0.0) + log1p(exp(-fabs(r)));
        acc += sp - X[e] * r;
    }
    block_sum3<256>(acc, z1, z2, red, threadIdx.x);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void sum_partials_kernel(const double* partial, int n, double scale, double* out) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = s * scale;
}

// the reference's stand-alone Adam step (linear.py:158-163): updates m, v, writes the direction
__global__ void adam_direction_kernel(size_t n, const double* __restrict__ g, double* __restrict__ m,
                                      double* __restrict__ v, double b1, double b2, double bc1, double bc2,
                                      double* __restrict__ out) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const double go = g[e];
        const double mn = m[e] * b1 + (1.0 - b1) * go;
        const double vn = v[e] * b2 + (1.0 - b2) * (go * go);
        m[e] = mn;
        v[e] = vn;
        out[e] = (mn / bc1) / (sqrt(vn / bc2) + 1e-8);
    }
}

}  // namespace dagma

using namespace dagma;

// =============================================================================== C ABI
extern "C" int dagma_adam_direction_f64(dagma_stream_t stream, size_t n, const double* grad_dev, double* m_dev,
                                        double* v_dev, double beta1, double beta2, double bias1, double bias2,
                                        double* out_dev) {
    DAGMA_REQUIRE(grad_dev && m_dev && v_dev && out_dev, "null pointer");
    const int blocks = (int)((n + 255) / 256);
    adam_direction_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, (cudaStream_t)stream>>>(n, grad_dev, m_dev, v_dev, beta1,
                                                                                         beta2, bias1, bias2, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" size_t dagma_large_workspace_bytes(int d) { return large_ws_bytes(d); }

extern "C" int dagma_gemm_f64(dagma_stream_t stream, int trans_a, int M, int N, int K, double alpha,
                              const double* a_dev, int lda, const double* b_dev, int ldb, double beta,
                              double* c_dev, int ldc, int epilogue, double* ws_dev, size_t ws_bytes) {
    DAGMA_REQUIRE(a_dev && b_dev && c_dev, "null pointer");
    DAGMA_REQUIRE(K >= 0 && lda >= 1 && ldb >= N && ldc >= N, "bad leading dimension");
    return gemm_launch((cudaStream_t)stream, trans_a, M, N, K, alpha, a_dev, lda, b_dev, ldb, beta, c_dev, ldc,
                       epilogue, ws_dev, ws_bytes);
}

extern "C" int dagma_logdet_inv_ws_f64(dagma_stream_t stream, int d, double s, const double* a_dev, int lda,
                                       int square_input, double* logabsdet_dev, double* h_dev, double* minv_dev,
                                       double* grad_dev, int ldo, double* min_entry_dev, int* info_dev,
                                       double* ws_dev, size_t ws_bytes) {
    DAGMA_REQUIRE(d >= 1 && a_dev, "bad arguments");
    if (d <= DAGMA_ONCHIP_INV_MAX_D)
        return logdet_inv_small((cudaStream_t)stream, 1, d, s, a_dev, lda, square_input, logabsdet_dev, h_dev,
                                minv_dev, grad_dev, ldo, min_entry_dev, info_dev);
    DAGMA_REQUIRE(ws_dev && ws_bytes >= large_ws_bytes(d), "workspace too small (dagma_large_workspace_bytes)");
    return logdet_inv_blocked((cudaStream_t)stream, d, s, a_dev, lda, square_input, logabsdet_dev, h_dev, minv_dev,
                              grad_dev, ldo, min_entry_dev, info_dev, ws_dev);
}

extern "C" int dagma_linear_update_f64(dagma_stream_t stream, int d, void* state_dev, double* w_dev,
                                       const double* minv_dev, const double* t_dev, const double* cov_dev,
                                       double* m_dev, double* v_dev, const uint8_t* mask_exc_dev,
                                       const uint8_t* mask_inc_dev) {
    DAGMA_REQUIRE(state_dev && w_dev && minv_dev && t_dev && cov_dev && m_dev && v_dev, "null pointer");
    dim3 grid((d + 31) / 32, (d + 31) / 32);
    linear_update_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((LinState*)state_dev, d, w_dev, minv_dev, t_dev,
                                                                        cov_dev, m_dev, v_dev, mask_exc_dev, mask_inc_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    linear_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((LinState*)state_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_linear_apply_dir_f64(dagma_stream_t stream, int d, const void* state_dev, double* w_dev,
                                          const double* m_dev, const double* v_dev, double sign) {
    DAGMA_REQUIRE(state_dev && w_dev && m_dev && v_dev, "null pointer");
    linear_apply_dir_kernel<<<592, 256, 0, (cudaStream_t)stream>>>((const LinState*)state_dev, d, w_dev, m_dev, v_dev, sign);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_linear_objective_f64(dagma_stream_t stream, int d, void* state_dev, const double* w_dev,
                                          const double* t_dev, const double* cov_dev, int l2) {
    DAGMA_REQUIRE(state_dev && w_dev, "null pointer");
    linear_objective_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((LinState*)state_dev, d, w_dev, t_dev, cov_dev, l2);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_logistic_loss_f64(dagma_stream_t stream, int n, int d, const double* x_dev, const double* r_dev,
                                       double scale, double* partial_dev, int n_partial, double* out_dev) {
    DAGMA_REQUIRE(x_dev && r_dev && partial_dev && out_dev && n_partial >= 1, "bad arguments");
    const int blocks = n_partial < 592 ? n_partial : 592;
    logistic_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, r_dev, (size_t)n * d, partial_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    sum_partials_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(partial_dev, blocks, scale, out_dev);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
