// On-chip Gauss-Jordan sweep for one d x d FP64 problem held by one CTA.
//
// Layout.  The CTA is a TY x TX thread grid; thread (ty, tx) owns an RM x RN register
// tile of the padded matrix (DP = TY*RM = TX*RN).  Rows use the interleaved map
// grow(ty, i) = (i / CW) * TY*CW + CW*ty + i % CW and columns the same with TX (CW = 2
// doubles = one 16-byte chunk), so the row / column fragments a thread needs in every
// step are 16-byte shared-memory chunks read without bank conflicts.  Warps are
// 4 (ty) x 8 (tx) patches so fragments broadcast inside a warp.
//
// Sweep.  In-place Gauss-Jordan without pivoting (stable for the M-matrices
// sI - W o W of DAGMA: off-diagonal updates add numbers of one sign), pivots in
// natural order.  Step k: the owners publish row k, column k and 1/pivot through a
// double-buffered shared-memory line (ONE __syncthreads per step) and every thread
// applies  a_ij += (-a_ik / p) * a_kj  to its whole tile.  Publishing row[k] = 1 + p
// and col[k] = p - 1 makes that same FMA turn the pivot row into a_kj / p, the pivot
// column into -a_ik / p and the pivot into 1 / p, so no element is special-cased and
// nothing is zeroed.  (The identity costs a relative error eps*max(1, p); p <= s,
// which the callers keep <= 2 -- by a power-of-two pre-scale where s is arbitrary.)
// The reciprocal of the NEXT pivot is started by its owner inside the current step,
// which takes the FP64 reciprocal off the serial publish/barrier/read chain.
// With GEMM = true the k-th rank-1 update of  G -= cov[:, k] W[k, :]  is interleaved
// into the same step (operands prefetched one step ahead): it is independent of the
// sweep, has the same shape, and fills the latency of the chain.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dagma {

// ---------------------------------------------------------------- shared memory access
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lds128(uint32_t a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts64_if(bool p, uint32_t a, double x) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.shared.f64 [%1], %2;\n\t}"
                 ::"r"((int)p), "r"(a), "d"(x) : "memory");
}
__device__ __forceinline__ void sts128_if(bool p, uint32_t a, double x, double y) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.shared.v2.f64 [%1], {%2, %3};\n\t}"
                 ::"r"((int)p), "r"(a), "d"(x), "d"(y) : "memory");
}

// full-precision reciprocal without the slow-path branch of 1.0/x (inputs are pivots:
// finite, normal range; 0 gives inf, which the callers flag as "not an M-matrix")
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);      // 2^-20 -> 2^-40
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);             // -> 2^-80
    r = fma(r, e, r);
    return r;
}

// sqrt(x) for x >= 0 (0 allowed) to ~1 ulp, branch free
__device__ __forceinline__ double fast_sqrt_nonneg(double x) {
    const double xs = fmax(x, 1e-290);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    double g = xs * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double res = fma(-g, g, xs);
    return fma(res, h, g);
}

// n / dn for dn in the normal range, ~1 ulp, branch free
__device__ __forceinline__ double fast_div(double n, double dn) {
    const double r = fast_rcp(dn);
    const double q = n * r;
    const double rem = fma(-q, dn, n);
    return fma(rem, r, q);
}

// ---------------------------------------------------------------- tiling
template <int RM_, int RN_, int TY_, int TX_>
struct Cfg {
    static constexpr int RM = RM_, RN = RN_, TY = TY_, TX = TX_;
    static constexpr int DP = TY * RM;
    static constexpr int NT = TY * TX;
    static constexpr int CW = (RM % 2 == 0 && RN % 2 == 0) ? 2 : 1;
    static constexpr int TMIN = TY < TX ? TY : TX;
    static constexpr int NU = DP / (TMIN * CW);           // unrolled super-chunks
    static_assert(TY * RM == TX * RN, "tile grid must be square in elements");
    static_assert(TY % 4 == 0 && TX % 8 == 0, "warps are 4 x 8 thread patches");
    static_assert(TY % TMIN == 0 && TX % TMIN == 0, "thread grid sides must nest");
    static_assert(RM % CW == 0 && RN % CW == 0, "chunking");
    static_assert(NT % 32 == 0 && NT <= 1024, "whole warps");
    __host__ __device__ static constexpr int grow(int ty, int i) {
        return (i / CW) * (TY * CW) + CW * ty + (i % CW);
    }
    __host__ __device__ static constexpr int gcol(int tx, int j) {
        return (j / CW) * (TX * CW) + CW * tx + (j % CW);
    }
};

template <class C>
struct ThreadPos {
    int ty, tx;
    __device__ __forceinline__ explicit ThreadPos(int tid) {
        const int w = tid >> 5, lane = tid & 31;
        constexpr int WX = C::TX / 8;
        ty = (w / WX) * 4 + (lane >> 3);
        tx = (w % WX) * 8 + (lane & 7);
    }
};

// row fragment: the RN entries line[gcol(tx, 0..RN-1)] of a shared-memory line (byte address)
template <class C>
__device__ __forceinline__ void load_rowfrag(double (&f)[C::RN], uint32_t line, int tx) {
#pragma unroll
    for (int ch = 0; ch < C::RN / C::CW; ++ch) {
        if constexpr (C::CW == 2)
            lds128(line + (ch * C::TX * 2 + 2 * tx) * 8, f[2 * ch], f[2 * ch + 1]);
        else
            f[ch] = lds64(line + (ch * C::TX + tx) * 8);
    }
}
// column fragment: the RM entries line[grow(ty, 0..RM-1)]
template <class C>
__device__ __forceinline__ void load_colfrag(double (&f)[C::RM], uint32_t line, int ty) {
#pragma unroll
    for (int ch = 0; ch < C::RM / C::CW; ++ch) {
        if constexpr (C::CW == 2)
            lds128(line + (ch * C::TY * 2 + 2 * ty) * 8, f[2 * ch], f[2 * ch + 1]);
        else
            f[ch] = lds64(line + (ch * C::TY + ty) * 8);
    }
}
template <class C>
__device__ __forceinline__ void store_rowfrag(const double (&f)[C::RN], uint32_t line, int tx) {
#pragma unroll
    for (int ch = 0; ch < C::RN / C::CW; ++ch) {
        if constexpr (C::CW == 2)
            sts128(line + (ch * C::TX * 2 + 2 * tx) * 8, f[2 * ch], f[2 * ch + 1]);
        else
            sts64(line + (ch * C::TX + tx) * 8, f[ch]);
    }
}

// Gauss-Jordan sweep over pivots 0..d-1 of the register-resident matrix `a`
// (on exit a = inverse on the leading d x d block; pivots[k] = k-th pivot).
// Shared-memory operands are byte addresses in the shared window:
//   line  : 2 x (DP row line, DP column line) doubles, then 2 reciprocals, then DP pivots
//   covT  : cov transposed, [k][row] (GEMM only);   Ws : W, [k][col] (GEMM only)
// GEMM: additionally g -= covT[k][rows] * Ws[k][cols] for k < d.
template <class C>
struct SweepSmem {
    static constexpr int row_off = 0;                  // doubles
    static constexpr int col_off = 2 * C::DP;
    static constexpr int pinv_off = 4 * C::DP;
    static constexpr int piv_off = 4 * C::DP + 2;
    static constexpr int doubles = 5 * C::DP + 2;
};

template <class C, bool GEMM>
__device__ __forceinline__ void gj_sweep(double (&a)[C::RM][C::RN], double (&g)[C::RM][C::RN],
                                         uint32_t covT, uint32_t Ws, uint32_t line, int d, int ty, int tx) {
    constexpr int RM = C::RM, RN = C::RN, DP = C::DP, CW = C::CW, TMIN = C::TMIN, NU = C::NU;
    using SS = SweepSmem<C>;
    const uint32_t rowbuf = line + SS::row_off * 8, colbuf = line + SS::col_off * 8;
    const uint32_t pinvbuf = line + SS::pinv_off * 8, pivots = line + SS::piv_off * 8;

    // reciprocal of pivot 0 (owner = thread (0,0), local (0,0))
    double pinv_early = fast_rcp(a[0][0]);
    double af[RM], bf[RN];
    if constexpr (GEMM) {
        load_colfrag<C>(af, covT, ty);
        load_rowfrag<C>(bf, Ws, tx);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int kbase = u * TMIN * CW;
        const int cr = (u * TMIN) / C::TY, cc = (u * TMIN) / C::TX;     // chunk of rows / cols
        const int tr0 = (u * TMIN) % C::TY, tc0 = (u * TMIN) % C::TX;   // first owner in chunk
        int tmax = (d - kbase + CW - 1) / CW;
        tmax = tmax < 0 ? 0 : (tmax > TMIN ? TMIN : tmax);
#pragma unroll 1
        for (int t = 0; t < tmax; ++t) {
#pragma unroll
            for (int e = 0; e < CW; ++e) {
                const int IK = cr * CW + e, JK = cc * CW + e;   // local indices (compile time)
                const int k = kbase + t * CW + e;
                if (k < d) {
                    const int cur = (CW == 2) ? e : (t & 1);
                    const uint32_t rb = rowbuf + cur * DP * 8, cb = colbuf + cur * DP * 8;
                    const bool own_r = (ty == tr0 + t), own_c = (tx == tc0 + t);
                    // ---- publish pivot row / column / reciprocal ----
#pragma unroll
                    for (int ch = 0; ch < RN / CW; ++ch) {
                        if constexpr (CW == 2)
                            sts128_if(own_r, rb + (ch * C::TX * 2 + 2 * tx) * 8, a[IK][2 * ch], a[IK][2 * ch + 1]);
                        else
                            sts64_if(own_r, rb + (ch * C::TX + tx) * 8, a[IK][ch]);
                    }
#pragma unroll
                    for (int i = 0; i < RM; ++i) sts64_if(own_c, cb + C::grow(ty, i) * 8, a[i][JK]);
                    {
                        const bool own = own_r && own_c;
                        const double p = a[IK][JK];
                        sts64_if(own, pivots + k * 8, p);
                        sts64_if(own, pinvbuf + cur * 8, pinv_early);
                        sts64_if(own, rb + k * 8, 1.0 + p);
                        sts64_if(own, cb + k * 8, p - 1.0);
                    }
                    __syncthreads();
                    // ---- read the published line ----
                    const double pinv = lds64(pinvbuf + cur * 8);
                    double r[RN], cs[RM];
                    load_rowfrag<C>(r, rb, tx);
                    load_colfrag<C>(cs, cb, ty);
                    // ---- GEMM rank-1 update with the operands fetched one step ago ----
                    if constexpr (GEMM) {
#pragma unroll
                        for (int i = 0; i < RM; ++i)
#pragma unroll
                            for (int j = 0; j < RN; ++j) g[i][j] = fma(-af[i], bf[j], g[i][j]);
                        const int kn = (k + 1 < d) ? k + 1 : k;
                        load_colfrag<C>(af, covT + kn * DP * 8, ty);
                        load_rowfrag<C>(bf, Ws + kn * DP * 8, tx);
                    }
#pragma unroll
                    for (int i = 0; i < RM; ++i) cs[i] = -cs[i] * pinv;
                    // ---- reciprocal of the next pivot, by its owner ----
                    {
                        double cand;
                        bool own_next;
                        if (e + 1 < CW) {
                            const int IN = (IK + 1 < RM) ? IK + 1 : IK, JN = (JK + 1 < RN) ? JK + 1 : JK;
                            cand = fma(cs[IN], r[JN], a[IN][JN]);
                            own_next = own_r && own_c;
                        } else {
                            const int IA = cr * CW, JA = cc * CW;                     // next t, same chunk
                            const double candA = fma(cs[IA], r[JA], a[IA][JA]);
                            double candB = 1.0;
                            bool ownB = false;
                            if (u + 1 < NU) {                                         // first of next chunk
                                const int crn = ((u + 1) * TMIN) / C::TY, ccn = ((u + 1) * TMIN) / C::TX;
                                const int IB = crn * CW, JB = ccn * CW;
                                candB = fma(cs[IB], r[JB], a[IB][JB]);
                                ownB = (ty == ((u + 1) * TMIN) % C::TY) && (tx == ((u + 1) * TMIN) % C::TX);
                            }
                            const bool last = (t == TMIN - 1);
                            cand = last ? candB : candA;
                            own_next = last ? ownB : ((ty == tr0 + t + 1) && (tx == tc0 + t + 1));
                        }
                        // warp-uniform branch: only the owner's warp pays for the Newton chain
                        if (__any_sync(0xffffffffu, own_next)) {
                            if (own_next) pinv_early = fast_rcp(cand);
                        }
                    }
                    // ---- Gauss-Jordan rank-1 update ----
#pragma unroll
                    for (int i = 0; i < RM; ++i)
#pragma unroll
                        for (int j = 0; j < RN; ++j) a[i][j] = fma(cs[i], r[j], a[i][j]);
                }
            }
        }
    }
    __syncthreads();
}

// deterministic block-wide sum of 3 values (all threads get the totals); red: 3*32 doubles
template <int NTHREADS>
__device__ __forceinline__ void block_sum3(double& x, double& y, double& z, double* red, int tid) {
    constexpr int NW = NTHREADS / 32;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, off);
        y += __shfl_xor_sync(0xffffffffu, y, off);
        z += __shfl_xor_sync(0xffffffffu, z, off);
    }
    const int w = tid >> 5;
    if ((tid & 31) == 0) {
        red[w] = x;
        red[32 + w] = y;
        red[64 + w] = z;
    }
    __syncthreads();
    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        sx += red[i];
        sy += red[32 + i];
        sz += red[64 + i];
    }
    __syncthreads();
    x = sx;
    y = sy;
    z = sz;
}

template <int NTHREADS>
__device__ __forceinline__ double block_min(double x, double* red, int tid) {
    constexpr int NW = NTHREADS / 32;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, off));
    const int w = tid >> 5;
    if ((tid & 31) == 0) red[w] = x;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) m = fmin(m, red[i]);
    __syncthreads();
    return m;
}

template <int NTHREADS>
__device__ __forceinline__ double block_max(double x, double* red, int tid) {
    constexpr int NW = NTHREADS / 32;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, off));
    const int w = tid >> 5;
    if ((tid & 31) == 0) red[w] = x;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) m = fmax(m, red[i]);
    __syncthreads();
    return m;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Telemetry of a checkpoint iteration (the fork's `minimize.checkpoint` row, src/dagma/linear.py:262-273, 307-324):
// per-thread partial sums over the thread's entries, reduced over the CTA by `finish` into one row of
// DAGMA_DIAG_COLS doubles.  Only the iterations that end a checkpoint interval pay for it (a uniform branch).
struct FitDiag {
    double go2 = 0.0, gs2 = 0.0, gh2 = 0.0, nl1 = 0.0, ninc = 0.0, dir2 = 0.0, w2 = 0.0, wabs = 0.0, wmax = 0.0,
           wmin = 1.7976931348623157e308;
    // gradient pieces of one entry: total, score part, acyclicity part; w != 0 / included edge for the l1 parts
    __device__ __forceinline__ void grad(double go, double gs, double gh, bool nonzero, bool inc) {
        go2 = fma(go, go, go2);
        gs2 = fma(gs, gs, gs2);
        gh2 = fma(gh, gh, gh2);
        nl1 += nonzero ? 1.0 : 0.0;
        ninc += (nonzero && inc) ? 1.0 : 0.0;
    }
    // Adam direction of the entry and the entry of W after the step (and the exclusion mask)
    __device__ __forceinline__ void step(double dir, double w_new) {
        dir2 = fma(dir, dir, dir2);
        w2 = fma(w_new, w_new, w2);
        const double aw = fabs(w_new);
        wabs += aw;
        wmax = fmax(wmax, aw);
        if (aw > 0.0) wmin = fmin(wmin, aw);
    }
    template <int NT>
    __device__ __forceinline__ void finish(double* red, int tid, double l1c, double incc, double seconds, double* row) {
        block_sum3<NT>(go2, gs2, gh2, red, tid);
        block_sum3<NT>(nl1, ninc, dir2, red, tid);
        double z = 0.0;
        block_sum3<NT>(w2, wabs, z, red, tid);
        wmax = block_max<NT>(wmax, red, tid);
        wmin = block_min<NT>(wmin, red, tid);
        if (tid == 0 && row) {
            row[0] = sqrt(go2);
            row[1] = sqrt(gs2);
            row[2] = sqrt(gh2);
            row[3] = fabs(l1c) * sqrt(nl1);
            row[4] = fabs(incc) * sqrt(ninc);
            row[5] = sqrt(dir2);
            row[6] = sqrt(w2);
            row[7] = wabs;
            row[8] = wmax;
            row[9] = (wmax > 0.0) ? wmin : 0.0;
            row[10] = seconds;
        }
    }
};

struct DD {   // double-double running power beta^k
    double hi, lo;
    __device__ __forceinline__ void mul(double b) {
        const double ph = hi * b;
        const double pl = fma(hi, b, -ph) + lo * b;
        const double s = ph + pl;
        lo = pl - (s - ph);
        hi = s;
    }
    __device__ __forceinline__ double one_minus() const { return (1.0 - hi) - lo; }
};

}  // namespace dagma
