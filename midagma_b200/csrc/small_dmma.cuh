// Tensor-core version of the on-chip sweep for 32 < d <= 64: block Gauss-Jordan with
// 4 x 4 pivot blocks whose rank-4 updates are DMMA.8x8x4 (mma.sync.m8n8k4.f64).
//
// Why.  The scalar rank-1 sweep (small_gj.cuh) is bound by shared->register bandwidth
// (16 doubles/clk/SM; a 4 x 2 register tile gets 1.2 FMA per loaded double, the FP64 pipe
// needs 4) and by one barrier per pivot.  A 16 x 32 warp tile updated by DMMA gets
// 10.7 FMA per loaded double and needs ONE barrier per 4 pivots.  The remaining serial
// chain (publish -> barrier -> 4 x 4 pivot-block inverse) is hidden by running TWO
// independent problems (CTAs) per SM: 256 threads and <= 128 registers each, with the
// Adam moments parked in tensor memory (TMEM) -- the 256 KB per SM that tcgen05 kernels
// use for accumulators and that is otherwise idle here (FP64 has no tcgen05 kind).
//
// Layout.  256 threads = 8 warps in a 4 (wr) x 2 (wc) grid (wr = warp / 2: the two warps that
// own the pivot rows of a step sit on different SM sub-partitions, and the diagonal warp shares
// its FP64 pipe with a warp that has little to do before the publish); warp (wr, wc) owns rows
// 16wr.., columns 32wc.. of the 64 x 64 padded matrix as 2 x 4 DMMA accumulator tiles:
//   a[ti][tj][e] = A[16wr + 8ti + lane/4][32wc + 8tj + 2(lane%4) + e].
// The score accumulator g and the Adam moments use the same ownership.
//
// Block step b (pivots K = 4b..4b+3) -- block form of the publish identity of small_gj.cuh:
//   publish  Rpub = A[K, :] + [I at the K block],  Cpub = A[:, K] - [I at the K block];
//            -Q = -(A[K,K])^{-1} was computed one step EARLY by the diagonal warp,
//   barrier, then every warp:  CS = Cpub (-Q)      (one DMMA per 8 rows, see cs_fragment),
//            A += CS Rpub  (8 DMMA),   G += (-cov)[:, K] W[K, :]   (8 DMMA).
// With the two I's the same product turns the pivot block into Q, the pivot rows into
// Q A[K,:] and the pivot columns into -A[:,K] Q: no tile is special and nothing is zeroed.
// The 4 x 4 pivot blocks are themselves inverted by scalar Gauss-Jordan in natural order,
// so the sweep yields the same d scalar pivots as small_gj.cuh: log|det| = sum log|pivot|,
// and "M-matrix" = all pivots > 0.
#pragma once
#include "small_gj.cuh"

namespace dagma {

constexpr int DM_DP = 64;          // padded dimension
constexpr int DM_NT = 256;         // threads per CTA
constexpr int DM_LD = 68;          // row stride (doubles) of the padded smem matrices / row lines
constexpr int DM_TMEM_COLS = 128;  // 32-bit TMEM columns per CTA: 2 warps per lane quarter x 64

struct DmmaSmem {                  // offsets in doubles
    static constexpr int ncov = 0;                         // -cov, [64][68]
    static constexpr int W = ncov + DM_DP * DM_LD;         // W,    [64][68]
    static constexpr int rbuf = W + DM_DP * DM_LD;         // 2 x [4][68]   published pivot rows
    static constexpr int cbuf = rbuf + 2 * 4 * DM_LD;      // 2 x [64][4]   published pivot columns
    static constexpr int qbuf = cbuf + 2 * DM_DP * 4;      // 2 x 16        -Q, row-major
    static constexpr int pinfo = qbuf + 32;                // 64            scalar pivots in natural order
    static constexpr int red = pinfo + 64;                 // 96
    static constexpr int mbar = red + 96;                  // 1 (the step barrier, 8 bytes)
    static constexpr int total = mbar + 2;
    static constexpr size_t bytes = (size_t)total * sizeof(double);
};
static_assert(DmmaSmem::rbuf % 2 == 0 && DmmaSmem::cbuf % 2 == 0 && DmmaSmem::qbuf % 2 == 0 &&
              DmmaSmem::pinfo % 2 == 0 && DmmaSmem::W % 2 == 0, "16-byte alignment of vector accesses");

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct DmmaPos {
    int warp, lane, wr, wc, qr, qc;
    __device__ __forceinline__ explicit DmmaPos(int tid) {
        warp = tid >> 5; lane = tid & 31; wr = warp >> 1; wc = warp & 1; qr = lane >> 2; qc = lane & 3;
    }
    __device__ __forceinline__ int row(int ti) const { return 16 * wr + 8 * ti + qr; }
    __device__ __forceinline__ int col(int tj) const { return 32 * wc + 8 * tj + 2 * qc; }   // + e
};

// ---------------------------------------------------------------- mbarrier (split arrive / wait)
__device__ __forceinline__ void mbar_init(uint32_t addr, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    // one cheap test first (the last arriver of a phase is usually the warp on the serial chain),
    // then try_wait, which suspends the warp in hardware
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
                 "WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(addr), "r"(parity) : "memory");
}

// The diagonal warp of block `b` inverts the pivot block P = A[K,K] (it sits in tile (TI,TJ),
// lanes with qr/4 == HALF and qc/2 == HALF), one element per lane, by a FRACTION-FREE 4 x 4
// Gauss-Jordan in natural pivot order: rows i != k become p_k row_i - x_ik row_k, so no
// reciprocal sits between two pivots; the row scales D_i = p_i prod_{m>i} p_m are divided out
// once at the end (one reciprocal chain instead of four on the serial path of the sweep).
// Broadcasts are warp shuffles.  For the Z-matrices of DAGMA the off-diagonal updates still
// add numbers of one sign, exactly as in the scaled elimination.
// Writes -Q for the next block step and the four fraction-free pivots p_k (all > 0 <=> all
// scalar pivots > 0; log|det P| = sum_k (k - 2) log|p_k|).
// `extra(k)` is issued right after the shuffles of pivot k: independent tensor work of the same warp that
// fits in the shuffle latency (the chain is latency bound, the warp's issue slots are mostly free).
template <int TI, int TJ, int HALF, class Extra>
__device__ __forceinline__ void stage_pivot_block(const double (&a)[2][4][2], const DmmaPos& ps, double* sm, int b,
                                                  int qslot, Extra&& extra) {
    constexpr unsigned FULL = 0xffffffffu;
    const int L = ps.lane & 15, i = L >> 2, j = L & 3;
    const int src = ((4 * HALF + i) << 2) | (2 * HALF + (j >> 1));
    const double v0 = __shfl_sync(FULL, a[TI][TJ][0], src), v1 = __shfl_sync(FULL, a[TI][TJ][1], src);
    double x = (j & 1) ? v1 : v0;
    double D = 1.0, E = 1.0;                  // E = prod_{m<k} p_m: scale of a row not yet pivoted
    double* pi = sm + DmmaSmem::pinfo + b * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double p = __shfl_sync(FULL, x, 5 * k);
        const double r = __shfl_sync(FULL, x, 4 * k + j), c = __shfl_sync(FULL, x, 4 * i + k);
        extra(k);
        if (ps.lane == 0) pi[k] = p;
        const bool ik = (i == k), jk = (j == k);
        // off the chain: what the pivot row / pivot column lanes become
        const double alt = ik ? (jk ? E : x) : (-c * E);
        const double Dp = D * p;
        D = (k >= i) ? Dp : D;                // D_i = prod_{m >= i} p_m
        E *= p;
        // on the chain: one multiply, one fma, one select
        const double t = fma(-c, r, p * x);
        x = (ik || jk) ? alt : t;
    }
    const double q = x * fast_rcp(D);
    if (ps.lane < 16) sm[DmmaSmem::qbuf + qslot * 16 + L] = -q;
}

// publish the pivot rows (+I) / pivot columns (-I) of block b = 8 bo + BQ into line buffer BQ & 1;
// the +-I touches only the pivot block itself, i.e. tile (TI, TJ) of the diagonal warp
template <int BQ>
__device__ __forceinline__ void publish_block(const double (&a)[2][4][2], const DmmaPos& ps, double* sm, int b) {
    constexpr int TI = (BQ & 3) >> 1, TJ = BQ >> 1, HALF = BQ & 1, CUR = BQ & 1;
    double* rb = sm + DmmaSmem::rbuf + CUR * 4 * DM_LD;
    double* cb = sm + DmmaSmem::cbuf + CUR * DM_DP * 4;
    const bool rown = ps.wr == (b >> 2), cown = ps.wc == (b >> 3);
    double p0 = a[TI][TJ][0], p1 = a[TI][TJ][1], m0 = p0, m1 = p1;      // pivot-block tile: row / column copy
    if (rown && cown) {
        const int dlt = ps.row(TI) - ps.col(TJ);
        const double i0 = (dlt == 0) ? 1.0 : 0.0, i1 = (dlt == 1) ? 1.0 : 0.0;
        p0 += i0; p1 += i1; m0 -= i0; m1 -= i1;
    }
    if (rown && (ps.qr >> 2) == HALF) {
        double* dst = rb + (ps.qr & 3) * DM_LD + ps.col(0);
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
            *reinterpret_cast<double2*>(dst + 8 * tj) = (tj == TJ) ? make_double2(p0, p1) : make_double2(a[TI][tj][0], a[TI][tj][1]);
    }
    if (cown && (ps.qc >> 1) == HALF) {
        double* dst = cb + ps.row(0) * 4 + 2 * (ps.qc & 1);
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
            *reinterpret_cast<double2*>(dst + 32 * ti) = (ti == TI) ? make_double2(m0, m1) : make_double2(a[ti][TJ][0], a[ti][TJ][1]);
    }
}

// G += (-cov)[:, K] W[K, :] for the k-block kb (8 DMMA); independent of the elimination
__device__ __forceinline__ void gemm_chunk(double (&g)[2][4][2], const DmmaPos& ps, const double* sm, int kb) {
    const double* nc = sm + DmmaSmem::ncov;
    const double* Ws = sm + DmmaSmem::W;
    const int k0 = 4 * kb;
    double an[2], bw[4];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) an[ti] = nc[ps.row(ti) * DM_LD + k0 + ps.qc];
#pragma unroll
    for (int tj = 0; tj < 4; ++tj) bw[tj] = Ws[(k0 + ps.qc) * DM_LD + 32 * ps.wc + 8 * tj + ps.qr];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) dmma(g[ti][tj][0], g[ti][tj][1], an[ti], bw[tj]);
}

#ifdef DAGMA_SWEEP_TRACE
// debug build only: cycle stamps of the serial chain (CTA 0, first 64 block steps after reset)
__device__ long long g_sweep_trace[64 * 8];
__device__ int g_sweep_trace_n;
#define SWEEP_STAMP(slot, cond) do { if (blockIdx.x == 0 && (cond) && ps.lane == 0 && sy.trace >= 0 && sy.trace < 64) g_sweep_trace[sy.trace * 8 + (slot)] = clock64(); } while (0)
#else
#define SWEEP_STAMP(slot, cond) do { } while (0)
#endif

struct SweepSync {        // per-thread view of the step barrier
    uint32_t bar;         // shared address of the mbarrier (count = warps)
    uint32_t phase;       // parity to wait for next
#ifdef DAGMA_SWEEP_TRACE
    int trace = -1;
#endif
};

// One block step; BQ = b % 8 is compile time, bo = b / 8.  On entry the lines of block b are
// published and every warp has arrived on the step barrier; the step
//   waits for it,
//   forms CS, updates FIRST the five tiles that hold the pivot rows / columns of block b + 1,
//   lets the diagonal warp invert the next pivot block, publishes the lines of block b + 1 and
//   arrives -- and only then updates its remaining three tiles, off the serial chain.
template <int BQ, bool GEMM>
__device__ __forceinline__ void dmma_block_step(double (&a)[2][4][2], double (&g)[2][4][2], const DmmaPos& ps,
                                                double* sm, int bo, int nb, SweepSync& sy) {
    constexpr int CUR = BQ & 1;
    constexpr int BN = (BQ + 1) & 7;                                   // next block (mod 8)
    constexpr int TIN = (BN & 3) >> 1, TJN = BN >> 1, HALFN = BN & 1;
    const int b = 8 * bo + BQ;
    const double* rb = sm + DmmaSmem::rbuf + CUR * 4 * DM_LD;
    const double* cb = sm + DmmaSmem::cbuf + CUR * DM_DP * 4;
    const double* qb = sm + DmmaSmem::qbuf + CUR * 16;
#ifdef DAGMA_SWEEP_TRACE
    const bool trc_diag = (b + 1 < nb) && (ps.wr == ((b + 1) >> 2)) && (ps.wc == ((b + 1) >> 3));
    const bool trc_other = (ps.warp == ((((b + 1) >> 2) + 2) & 3) + 4 * (1 - ((b + 1) >> 3)));   // a warp far from the diagonal
#endif
    SWEEP_STAMP(0, trc_diag);
    SWEEP_STAMP(5, trc_other);
    mbar_wait(sy.bar, sy.phase);
    sy.phase ^= 1u;
    SWEEP_STAMP(1, trc_diag);
    SWEEP_STAMP(6, trc_other);
    // ---- CS = Cpub (-Q) as ONE DMMA per 8 rows: B[k][n] = -Q[k][n/2] on even n, so the C
    //      fragment element c0 of lane (qr, qc) is CS[row qr][qc] = exactly its A fragment.
    const double bq = (ps.qr & 1) ? 0.0 : qb[ps.qc * 4 + (ps.qr >> 1)];
    double acs[2], ac[2], br[4];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti) ac[ti] = cb[ps.row(ti) * 4 + ps.qc];
#pragma unroll
    for (int tj = 0; tj < 4; ++tj) br[tj] = rb[ps.qc * DM_LD + 32 * ps.wc + 8 * tj + ps.qr];
    auto form_cs = [&](int ti) {
        double c0 = 0.0, c1 = 0.0;
        dmma(c0, c1, ac[ti], bq);
        acs[ti] = c0;
    };
    // ---- the serial chain first: CS of the next pivot rows, the next pivot block itself
    form_cs(TIN);
    dmma(a[TIN][TJN][0], a[TIN][TJN][1], acs[TIN], br[TJN]);
    const bool has_next = b + 1 < nb;
    SWEEP_STAMP(2, trc_diag);
    // the rest of the next pivot rows / columns (what the publish needs), in four portions
    constexpr int O0 = (TJN == 0) ? 1 : 0, O1 = (TJN <= 1) ? 2 : 1, O2 = (TJN <= 2) ? 3 : 2;   // tj != TJN
    auto portion = [&](int k) {
        if (k == 0) form_cs(TIN ^ 1);
        else if (k == 1) {
            dmma(a[TIN][O0][0], a[TIN][O0][1], acs[TIN], br[O0]);
            dmma(a[TIN][O1][0], a[TIN][O1][1], acs[TIN], br[O1]);
        } else if (k == 2) dmma(a[TIN][O2][0], a[TIN][O2][1], acs[TIN], br[O2]);
        else dmma(a[TIN ^ 1][TJN][0], a[TIN ^ 1][TJN][1], acs[TIN ^ 1], br[TJN]);
    };
    if (has_next && (ps.wr == ((b + 1) >> 2)) && (ps.wc == ((b + 1) >> 3)))
        stage_pivot_block<TIN, TJN, HALFN>(a, ps, sm, b + 1, CUR ^ 1, [](int) {});
#pragma unroll
    for (int k = 0; k < 4; ++k) portion(k);
    SWEEP_STAMP(3, trc_diag);
    if (has_next) {
        publish_block<BN>(a, ps, sm, b + 1);
        __syncwarp();
        if (ps.lane == 0) mbar_arrive(sy.bar);
    }
    SWEEP_STAMP(4, trc_diag);
    // ---- the three tiles nobody is waiting for
#pragma unroll
    for (int tj = 0; tj < 4; ++tj)
        if (tj != TJN) dmma(a[TIN ^ 1][tj][0], a[TIN ^ 1][tj][1], acs[TIN ^ 1], br[tj]);
    SWEEP_STAMP(7, trc_other);
#ifdef DAGMA_SWEEP_TRACE
    if (sy.trace >= 0) ++sy.trace;
#endif
}

// a := a^{-1} on the leading 4*ceil(d/4) block (padding inside the last pivot block must
// carry a unit diagonal); GEMM: g += (-cov) W over the same k range.
// pinfo[k] = fraction-free pivots (see stage_pivot_block), k < 4*ceil(d/4).
// All warps must have passed a __syncthreads since the last use of the line buffers.
template <bool GEMM>
__device__ __forceinline__ void dmma_sweep(double (&a)[2][4][2], double (&g)[2][4][2], const DmmaPos& ps, double* sm,
                                           int d, SweepSync& sy) {
    const int nb = (d + 3) >> 2;
    // the score GEMM first: 16 independent DMMA per k-block and warp, no serial chain -- this is
    // the phase that keeps the FP64 pipe busy while the other CTA of the SM is inside its sweep
    if constexpr (GEMM) {
#pragma unroll 1
        for (int kb = 0; kb < nb; ++kb) gemm_chunk(g, ps, sm, kb);
    }
    if (ps.wr == 0 && ps.wc == 0) stage_pivot_block<0, 0, 0>(a, ps, sm, 0, 0, [](int) {});
    publish_block<0>(a, ps, sm, 0);
    __syncwarp();
    if (ps.lane == 0) mbar_arrive(sy.bar);
    const int nbo = (nb + 7) >> 3;
#pragma unroll 1
    for (int bo = 0; bo < nbo; ++bo) {
        const int left = nb - 8 * bo;
        dmma_block_step<0, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 1) dmma_block_step<1, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 2) dmma_block_step<2, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 3) dmma_block_step<3, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 4) dmma_block_step<4, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 5) dmma_block_step<5, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 6) dmma_block_step<6, GEMM>(a, g, ps, sm, bo, nb, sy);
        if (left > 7) dmma_block_step<7, GEMM>(a, g, ps, sm, bo, nb, sy);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- tensor memory (TMEM) scratch
// Thread-private spill space: warp w owns TMEM lanes 32(w%4)..+31 (hardware rule for
// tcgen05.ld/st), thread = lane; the two warps of a lane quarter use disjoint column ranges.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot_addr) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot_addr),
                 "n"(DM_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr) {             // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(DM_TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 doubles <-> 16 consecutive 32-bit columns of the thread's own TMEM lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&x)[8]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r[2 * i] = (uint32_t)__double2loint(x[i]);
        r[2 * i + 1] = (uint32_t)__double2hiint(x[i]);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
struct TmemLoad8 {   // issue now, finish() later (after which get() is valid)
    uint32_t r[16];
    __device__ __forceinline__ void issue(uint32_t taddr) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
    }
    // the wait names the registers so no use of them can be scheduled above it
    __device__ __forceinline__ void finish() {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                       "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                       "+r"(r[15])
                     :
                     : "memory");
    }
    __device__ __forceinline__ double get(int i) const { return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]); }
};

}  // namespace dagma
