// On-chip Gauss-Jordan sweep for one d x d FP64 problem held by one 256-thread CTA.
//
// Layout.  The CTA is a 16 x 16 thread grid; thread (ty, tx) owns an R x R register
// tile of the (padded, DP = 16 R) matrix.  Rows and columns use the same interleaved
// map  g(t, i) = (i / CW) * 16 CW + CW t + i % CW  (CW = 2 for even R), so a thread's
// row/column fragments are 16-byte chunks that a warp reads from shared memory
// without bank conflicts, and the transpose of thread (ty, tx)'s tile is exactly the
// tile of thread (tx, ty).  Warps are 4 (ty) x 8 (tx) patches so that the row and
// column fragments broadcast inside a warp.
//
// Sweep.  In-place Gauss-Jordan without pivoting (stable for the M-matrices
// sI - W o W of DAGMA: every off-diagonal update adds numbers of one sign), pivots in
// natural order.  Step k: the owners publish row k, column k and 1/pivot through a
// double-buffered shared-memory line (one __syncthreads per step), zero their copies,
// and every thread applies  a_ij += (-a_ik / p) * a_kj  to its whole tile; publishing
// row[k] = 1, col[k] = -1 makes the same FMA produce the scaled pivot row, the
// negated pivot column and 1/p, so there is no per-element special case.  The next
// pivot's reciprocal is started one step early by its owner to take the FP64
// reciprocal off the serial chain.  With GEMM = true the k-th rank-1 update of
// G -= cov[:, k] W[k, :] is interleaved into the same step: it is independent of the
// sweep, has the same shape, and fills the latency of the publish/barrier/read chain.
#pragma once
#include <cuda_runtime.h>

namespace dagma {

constexpr int TG = 16;        // thread grid side
constexpr int NT = TG * TG;   // threads per CTA

template <int R>
struct Tile {
    static constexpr int CW = (R % 2 == 0) ? 2 : 1;   // chunk width (doubles)
    static constexpr int NCH = R / CW;                // chunks per fragment
    static constexpr int DP = TG * R;                 // padded dimension
    static constexpr int XLD = TG * (TG + 1);         // exchange-buffer plane stride
    __host__ __device__ static constexpr int g(int t, int i) {
        return (i / CW) * (TG * CW) + CW * t + (i % CW);
    }
};

struct ThreadPos {
    int ty, tx;
    __device__ __forceinline__ explicit ThreadPos(int tid) {
        const int w = tid >> 5, lane = tid & 31;
        ty = (w >> 1) * 4 + (lane >> 3);
        tx = (w & 1) * 8 + (lane & 7);
    }
};

// fragment = the R entries base[g(t, 0..R-1)] of one shared-memory line
template <int R>
__device__ __forceinline__ void load_frag(double (&f)[R], const double* base, int t) {
    using T = Tile<R>;
#pragma unroll
    for (int ch = 0; ch < T::NCH; ++ch) {
        if constexpr (T::CW == 2) {
            const double2 v = *reinterpret_cast<const double2*>(base + ch * (TG * 2) + 2 * t);
            f[2 * ch] = v.x;
            f[2 * ch + 1] = v.y;
        } else {
            f[ch] = base[ch * TG + t];
        }
    }
}

template <int R>
__device__ __forceinline__ void store_frag(const double (&f)[R], double* base, int t) {
    using T = Tile<R>;
#pragma unroll
    for (int ch = 0; ch < T::NCH; ++ch) {
        if constexpr (T::CW == 2) {
            *reinterpret_cast<double2*>(base + ch * (TG * 2) + 2 * t) = make_double2(f[2 * ch], f[2 * ch + 1]);
        } else {
            base[ch * TG + t] = f[ch];
        }
    }
}

// tile[i][j] = src[g(ty,i) * DP + g(tx,j)]
template <int R>
__device__ __forceinline__ void load_tile(double (&tile)[R][R], const double* src, int ty, int tx) {
    using T = Tile<R>;
#pragma unroll
    for (int i = 0; i < R; ++i) load_frag<R>(tile[i], src + T::g(ty, i) * T::DP, tx);
}

template <int R>
__device__ __forceinline__ void store_tile(const double (&tile)[R][R], double* dst, int ty, int tx) {
    using T = Tile<R>;
#pragma unroll
    for (int i = 0; i < R; ++i) store_frag<R>(tile[i], dst + T::g(ty, i) * T::DP, tx);
}

// out[i][j] = in_of_thread(tx,ty)[j][i]  (pairwise exchange through shared memory)
template <int R>
__device__ __forceinline__ void transpose_tile(double (&out)[R][R], const double (&in)[R][R],
                                               double* xch, int ty, int tx) {
    using T = Tile<R>;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) xch[(i * R + j) * T::XLD + ty * (TG + 1) + tx] = in[i][j];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) out[i][j] = xch[(j * R + i) * T::XLD + tx * (TG + 1) + ty];
}

// Gauss-Jordan sweep over pivots 0..d-1 of the register-resident matrix `a`
// (on exit a = inverse on the leading d x d block; pivots[k] = k-th pivot).
// GEMM: additionally g -= covT[k][rows] * Ws[k][cols] for k < d.
template <int R, bool GEMM>
__device__ __forceinline__ void gj_sweep(double (&a)[R][R], double (&g)[R][R],
                                         const double* __restrict__ covT,
                                         const double* __restrict__ Ws,
                                         double* rowbuf, double* colbuf, double* pinvbuf,
                                         double* pivots, int d, int ty, int tx) {
    using T = Tile<R>;
    constexpr int DP = T::DP, CW = T::CW;
    double pinv_early = 0.0;
    int cur = 0;
#pragma unroll
    for (int c = 0; c < T::NCH; ++c) {
        const int kbase = c * TG * CW;
        int tmax = (d - kbase + CW - 1) / CW;
        tmax = tmax < 0 ? 0 : (tmax > TG ? TG : tmax);
#pragma unroll 1
        for (int t = 0; t < tmax; ++t) {
#pragma unroll
            for (int e = 0; e < CW; ++e) {
                const int IK = c * CW + e;          // local index of pivot k (compile time)
                const int k = kbase + t * CW + e;
                if (k < d) {
                    double* rb = rowbuf + cur * DP;
                    double* cb = colbuf + cur * DP;
                    const bool own_r = (ty == t), own_c = (tx == t);
                    // ---- publish pivot row / column / reciprocal ----
                    if (own_r) store_frag<R>(a[IK], rb, tx);
                    if (own_c) {
                        double colv[R];
#pragma unroll
                        for (int i = 0; i < R; ++i) colv[i] = a[i][IK];
                        store_frag<R>(colv, cb, ty);
                    }
                    if (own_r && own_c) {
                        const double p = a[IK][IK];
                        const bool late = (e == 0) && (t == 0);
                        const double pinv = late ? __drcp_rn(p) : pinv_early;
                        pivots[k] = p;
                        pinvbuf[cur] = pinv;
                        rb[k] = 1.0;
                        cb[k] = -1.0;
                    }
                    if (own_r) {
#pragma unroll
                        for (int j = 0; j < R; ++j) a[IK][j] = 0.0;
                    }
                    if (own_c) {
#pragma unroll
                        for (int i = 0; i < R; ++i) a[i][IK] = 0.0;
                    }
                    __syncthreads();
                    // ---- read the published line ----
                    const double pinv = pinvbuf[cur];
                    double r[R], cs[R];
                    load_frag<R>(r, rb, tx);
                    load_frag<R>(cs, cb, ty);
#pragma unroll
                    for (int i = 0; i < R; ++i) cs[i] = -cs[i] * pinv;
                    // ---- start the next pivot's reciprocal early ----
                    if (k + 1 < d) {
                        if (e + 1 < CW) {
                            const int IN = IK + 1 < R ? IK + 1 : IK;
                            if (own_r && own_c)
                                pinv_early = __drcp_rn(fma(cs[IN], r[IN], a[IN][IN]));
                        } else if (t + 1 < TG) {
                            const int IN = c * CW;
                            if (ty == t + 1 && tx == t + 1)
                                pinv_early = __drcp_rn(fma(cs[IN], r[IN], a[IN][IN]));
                        }
                    }
                    // ---- rank-1 updates ----
                    if constexpr (GEMM) {
                        double af[R], bf[R];
                        load_frag<R>(af, covT + k * DP, ty);
                        load_frag<R>(bf, Ws + k * DP, tx);
#pragma unroll
                        for (int i = 0; i < R; ++i)
#pragma unroll
                            for (int j = 0; j < R; ++j) g[i][j] = fma(-af[i], bf[j], g[i][j]);
                    }
#pragma unroll
                    for (int i = 0; i < R; ++i)
#pragma unroll
                        for (int j = 0; j < R; ++j) a[i][j] = fma(cs[i], r[j], a[i][j]);
                    cur ^= 1;
                }
            }
        }
    }
    __syncthreads();
}

// deterministic block-wide sum of up to 3 values (all threads get the totals)
__device__ __forceinline__ void block_sum3(double& x, double& y, double& z, double* red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, off);
        y += __shfl_xor_sync(0xffffffffu, y, off);
        z += __shfl_xor_sync(0xffffffffu, z, off);
    }
    const int w = tid >> 5;
    if ((tid & 31) == 0) {
        red[w] = x;
        red[8 + w] = y;
        red[16 + w] = z;
    }
    __syncthreads();
    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        sx += red[i];
        sy += red[8 + i];
        sz += red[16 + i];
    }
    __syncthreads();
    x = sx;
    y = sy;
    z = sz;
}

__device__ __forceinline__ double block_min(double x, double* red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, off));
    const int w = tid >> 5;
    if ((tid & 31) == 0) red[w] = x;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int i = 1; i < NT / 32; ++i) m = fmin(m, red[i]);
    __syncthreads();
    return m;
}

}  // namespace dagma
