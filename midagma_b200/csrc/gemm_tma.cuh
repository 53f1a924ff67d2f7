// TMA-fed slab pipeline for the FP64 DMMA tile engine (64 x 64 output tile, 128 threads = 4 warps in a
// 2 x 2 grid, warp tile 32 x 32 = 4 x 4 DMMA.8x8x4 accumulator tiles).
//
// The cp.async engine of gemm_f64.cuh spends one CTA-wide barrier and 8 address computations + copies
// per thread on every 64 x 16 / 16 x 64 operand slab; its four warps therefore move in lock step and a
// warp that is late on its sub-partition's FP64 pipe stalls the other three.  Here a slab is TWO bulk
// tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG) issued by one elected thread and tracked by
// mbarriers: full[s] (count 1 + transaction bytes) is what a warp waits for before it reads stage s,
// empty[s] (one arrive per warp, after a fence.proxy.async of every thread: generic-proxy reads before the
// async-proxy refill) is what the elected thread waits for before it refills the stage -- one slab LATER
// than the warps released it, so that no warp waits for another warp unless it is a whole slab ahead.
// The pipeline state (stage, phase bits) runs on across tiles.
//
// Shared layout of a stage (16 KB, 1024-byte aligned):
//   A box  [64 rows m][16 k]  = 64 x 128 B, CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk c of row r
//          lands at chunk c ^ (r & 7).  A DMMA fragment load a = A[8i + lane/4][kk + lane%4] touches
//          8 rows x 32 B; the swizzle spreads them over all banks (two wavefronts, the minimum for 256 B).
//   B box  [16 rows k][64 n]  = 16 x 512 B, dense.  b = B[kk + lane%4][8j + lane/4] touches 4 rows x 64 B
//          on the same banks (4 wavefronts instead of 2) -- at 16 DMMA (256 pipe clocks) per 8 fragment
//          loads the shared-memory pipe stays under 40 % busy, so the B operand keeps its row-major
//          global layout and needs no transposed copy.
// Out-of-range rows / columns are zero-filled by the TMA unit, so edge tiles need no predicates.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "common.cuh"

namespace dagma {

constexpr int TM_BK = 16;
// DAGMA_TM_FIX (A-B experiments, scripts/check_tma_repeat2.py): bit 2 = every thread executes fence.proxy.async
// before its warp releases a stage -- REQUIRED: the stage is read through the generic proxy (ld.shared) and
// refilled through the async proxy (TMA), and without the cross-proxy fence the refill can overtake the last
// reads (measured on B200: 7 % of 2000^3 products wrong in the balanced schedule without it, 0 of 300 with it).
// bit 0 / bit 1 = proxy fence / warp sync after the full wait (no effect), bit 3 = a spare stage (also hides it).
#ifndef DAGMA_TM_FIX
#define DAGMA_TM_FIX 4
#endif
// CTA (or engine) tile BM x BN, warps in a (BM / 32) x (BN / 32) grid, STAGES slabs of BM x 16 + 16 x BN doubles
template <int BM_, int BN_, int STAGES_, int DIST_ = STAGES_ - 1>
struct TmaCfg {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_;
    static constexpr int DIST = DIST_;      // slabs requested ahead of the one being multiplied (<= STAGES - 1)
    static constexpr int WM = BM / 32, WN = BN / 32, WARPS = WM * WN, THREADS = 32 * WARPS;
    static constexpr int A_BYTES = BM * TM_BK * 8, B_BYTES = TM_BK * BN * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;                   // (+ up to 1023 bytes of alignment slack)
    static_assert(BM % 32 == 0 && BN % 32 == 0 && BM <= 256 && BN <= 256, "whole 32 x 32 warp tiles, box <= 256");
    static_assert(A_BYTES % 1024 == 0 && STAGE_BYTES % 1024 == 0, "swizzled boxes start on 1024-byte boundaries");
};
#if DAGMA_TM_FIX & 8
using Tm64 = TmaCfg<64, 64, 4, 2>;    // debug: one spare stage between release and refill
#else
using Tm64 = TmaCfg<64, 64, 3>;       // the 128-thread engine of the persistent inverse kernels (16 KB stages)
#endif
using Tm128 = TmaCfg<128, 128, 6>;    // one 512-thread CTA per SM (32 KB stages, 192 KB pipeline)
constexpr int TM_STAGES = Tm64::STAGES;
constexpr int TM_PIPE_BYTES = Tm64::PIPE_BYTES;

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*TmEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmEncodeTiledFn tm_encode_fn() {
    static TmEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TmEncodeTiledFn>(p);
    }
    return fn;
}
// row-major FP64 matrix `rows x cols` (leading dimension ld doubles, ld even, base 16-byte aligned);
// box = box_rows x box_cols; swizzle128 needs box_cols * 8 == 128
inline int tm_make_map(CUtensorMap* map, const double* base, int rows, int cols, int ld, int box_rows, int box_cols,
                       bool swizzle128) {
    TmEncodeTiledFn fn = tm_encode_fn();
    if (!fn) return set_error(-3, "cuTensorMapEncodeTiled is not available from the driver");
    if ((ld & 1) || (reinterpret_cast<uintptr_t>(base) & 15)) return set_error(-1, "TMA operand is not 16-byte aligned");
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[64];
        snprintf(buf, sizeof(buf), "CUresult %d", (int)r);
        return set_error(-4, "cuTensorMapEncodeTiled failed", buf);
    }
    return 0;
}
inline int tm_make_a_map(CUtensorMap* map, const double* A, int M, int K, int lda, int bm = 64) {   // A: M x K, box bm x 16, swizzled
    return tm_make_map(map, A, M, K, lda, bm, TM_BK, true);
}
inline int tm_make_b_map(CUtensorMap* map, const double* B, int K, int N, int ldb, int bn = 64) {   // B: K x N, box 16 x bn, dense
    return tm_make_map(map, B, K, N, ldb, TM_BK, bn, false);
}

// ------------------------------------------------------------------ device: mbarrier / TMA primitives
__device__ __forceinline__ void tm_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tm_mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool tm_mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tm_mbar_wait(uint32_t bar, uint32_t parity) {
    while (!tm_mbar_test(bar, parity)) { }
}
__device__ __forceinline__ void tm_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tm_fence_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tm_load_2d(uint32_t dst, const CUtensorMap* map, int c_inner, int c_row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner), "r"(c_row), "r"(bar) : "memory");
}
__device__ __forceinline__ void tm_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ double tm_lds(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

// ------------------------------------------------------------------ the pipeline of one engine (Cfg::THREADS threads)
template <class Cfg>
struct TmaPipe {
    uint32_t stage0;        // shared address of stage 0 (1024-byte aligned)
    uint32_t bars;          // shared address of full[0..S), then empty[0..S)
    uint32_t is, iph;       // issue side (elected thread): next stage, parity to wait for on empty[is] (valid once wrapped)
    uint32_t wrapped;       // issue side: every stage has been filled at least once
    uint32_t cs, cph;       // consume side (every thread): next stage, parity of full[cs]
    __device__ __forceinline__ uint32_t full(uint32_t s) const { return bars + 8u * s; }
    __device__ __forceinline__ uint32_t empty(uint32_t s) const { return bars + 8u * (Cfg::STAGES + s); }
};
// `region`: >= Cfg::PIPE_BYTES + 1023 bytes of shared memory owned by the engine; `bars`: 2 * STAGES mbarriers.
// Called by all threads of the engine; the caller synchronises the engine afterwards.
template <class Cfg>
__device__ __forceinline__ void tm_pipe_init(TmaPipe<Cfg>& p, uint32_t region, uint32_t bars, bool elected) {
    p.stage0 = (region + 1023u) & ~1023u;
    p.bars = bars;
    p.is = p.iph = p.wrapped = p.cs = p.cph = 0u;
    if (elected) {
#pragma unroll
        for (int s = 0; s < Cfg::STAGES; ++s) {
            tm_mbar_init(p.full(s), 1);
            tm_mbar_init(p.empty(s), Cfg::WARPS);
        }
        tm_fence_init();
    }
}
// elected thread: request the slab A[r0.., k..k+16) / B[k.., c0..) into the next stage
template <class Cfg>
__device__ __forceinline__ void tm_issue(TmaPipe<Cfg>& p, const CUtensorMap* mapA, const CUtensorMap* mapB, int r0, int c0,
                                         int k) {
    const uint32_t s = p.is;
    if (p.wrapped) tm_mbar_wait(p.empty(s), p.iph);
    const uint32_t dst = p.stage0 + s * Cfg::STAGE_BYTES, bar = p.full(s);
    tm_mbar_expect_tx(bar, Cfg::STAGE_BYTES);
    tm_load_2d(dst, mapA, k, r0, bar);
    tm_load_2d(dst + Cfg::A_BYTES, mapB, c0, k, bar);
    if (++p.is == Cfg::STAGES) {
        p.is = 0;
        if (p.wrapped) p.iph ^= 1u;
        p.wrapped = 1u;
    }
}

template <class Cfg>
struct TmaFrag {            // per-thread constants of the fragment addressing
    uint32_t a_row[4];      // byte offset of row 32 wm + 8 i + qr in the A box, + (qc & 1) * 8
    uint32_t a_x;           // (qc >> 1) ^ qr : chunk index of k = kk + qc is (kk >> 1) ^ a_x
    uint32_t b_off;         // byte offset of B[qc][32 wn + qr]
    __device__ __forceinline__ TmaFrag(int wm, int wn, int qr, int qc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a_row[i] = (uint32_t)((32 * wm + 8 * i + qr) * 128 + (qc & 1) * 8);
        a_x = (uint32_t)((qc >> 1) ^ qr);
        b_off = (uint32_t)((qc * Cfg::BN + 32 * wn + qr) * 8);
    }
};
// every thread of the engine: wait for the next slab, multiply it into acc, release the stage
template <class Cfg>
__device__ __forceinline__ void tm_consume(TmaPipe<Cfg>& p, const TmaFrag<Cfg>& f, double (&acc)[4][4][2], int lane) {
    const uint32_t s = p.cs;
    tm_mbar_wait(p.full(s), p.cph);
#if DAGMA_TM_FIX & 1
    tm_fence_proxy();
#endif
#if DAGMA_TM_FIX & 2
    __syncwarp();
#endif
    const uint32_t sa = p.stage0 + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
#pragma unroll
    for (int kk = 0; kk < TM_BK; kk += 4) {
        double a[4], b[4];
        const uint32_t ch = (((uint32_t)(kk >> 1)) ^ f.a_x) << 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = tm_lds(sa + f.a_row[i] + ch);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = tm_lds(sb + f.b_off + (uint32_t)(kk * Cfg::BN + 8 * j) * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(acc[i][j][0]), "+d"(acc[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
#if DAGMA_TM_FIX & 4
    tm_fence_proxy();
#endif
    __syncwarp();
    if (lane == 0) tm_mbar_arrive(p.empty(s));
    if (++p.cs == Cfg::STAGES) {
        p.cs = 0;
        p.cph ^= 1u;
    }
}

// acc += A[r0..r0+BM, kbeg..kbeg+16 nk) B[.., c0..c0+BN) for one tile.  `primed` slabs of this tile have already
// been requested (by the previous tile's tail; only the elected thread's value matters).  After it has
// multiplied a slab the elected thread requests ONE more slab -- of this tile while it has any, then the
// first S - 1 of the next tile -- which goes into the stage the warps released one slab ago.  Returns the
// number of slabs of `nxt` that have been requested (elected thread).
struct TmaTile {
    int r0, c0, kbeg, nk;       // nk = 0: no tile
};
template <class Cfg>
__device__ __forceinline__ int tm_tile_gemm(TmaPipe<Cfg>& p, const TmaFrag<Cfg>& f, double (&acc)[4][4][2],
                                            const CUtensorMap* mapA, const CUtensorMap* mapB, const TmaTile& t,
                                            int primed, const TmaTile& nxt, bool elected, int lane) {
    constexpr int D = Cfg::DIST;
    int issued = primed, nissued = 0;
    const int nlim = nxt.nk < D ? nxt.nk : D;
    if (elected)
        while (issued < D && issued < t.nk) {
            tm_issue(p, mapA, mapB, t.r0, t.c0, t.kbeg + issued * TM_BK);
            ++issued;
        }
    for (int kt = 0; kt < t.nk; ++kt) {
        tm_consume(p, f, acc, lane);
        if (elected) {
            if (issued < t.nk) {
                tm_issue(p, mapA, mapB, t.r0, t.c0, t.kbeg + issued * TM_BK);
                ++issued;
            } else if (nissued < nlim) {
                tm_issue(p, mapA, mapB, nxt.r0, nxt.c0, nxt.kbeg + nissued * TM_BK);
                ++nissued;
            }
        }
    }
    return nissued;
}

}  // namespace dagma
