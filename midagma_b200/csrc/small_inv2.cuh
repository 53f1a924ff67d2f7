// Two-by-two block Gauss-Jordan of an identity-padded 128 x 128 matrix held as four 64 x 64 tiles in shared memory
// (shared by small_inv.cu and lin_iter.cu).
#pragma once
#include "small_dmma.cuh"

namespace dagma {

// 64 < d <= 128: two-by-two block Gauss-Jordan on chip.  The four 64 x 64 tiles of the (identity padded) matrix live
// in shared memory ([64][68] doubles each, the row stride the DMMA fragment loads want); per block step kb
//   Q = T[kb][kb]^{-1}            by the tensor-core sweep above (in registers, then to a fifth tile buffer),
//   T[o][kb]  = -T[o][kb] Q       (new pivot column, o = 1 - kb; in place: its old value is not needed afterwards)
//   T[o][o]  +=  T[o][kb] T[kb][o]
//   T[kb][o]  =  Q T[kb][o]       (new pivot row, in place),       T[kb][kb] = Q
// three 64^3 DMMA products per step instead of 64 barrier-separated rank-1 updates of the scalar sweep.
// One 256-thread CTA per problem and SM (194 KB of shared memory).
struct Inv2Smem {                  // offsets in doubles; the sweep's buffers keep their DmmaSmem offsets
    static constexpr int tile = DM_DP * DM_LD;
    static constexpr int t0 = DmmaSmem::ncov, t1 = DmmaSmem::W;
    static constexpr int t2 = (DmmaSmem::total + 1) & ~1, t3 = t2 + tile, q = t3 + tile, piv = q + tile;
    static constexpr int total = piv + 2 * DM_DP;
    static constexpr size_t bytes = (size_t)total * sizeof(double);
    __device__ static constexpr int at(int bi, int bj) { return bi == 0 ? (bj == 0 ? t0 : t1) : (bj == 0 ? t2 : t3); }
};
// The two block steps on the tiles in shared memory (all threads of a DM_NT CTA; the tiles are complete and a barrier
// has been passed).  On return the tiles hold the inverse, pv[0 .. 127] the fraction-free pivots (1 in the padding), and
// a barrier has been passed.
// `final_tile(bi, bj, a, sign)` is called by every thread with the accumulators of tile (bi, bj) of the INVERSE (entry =
// sign * a[ti][tj][e] at row 64 bi + ps.row(ti), column 64 bj + ps.col(tj) + e) when that tile is final -- results can
// leave the CTA without a second pass over shared memory.
struct Inv2NoOut {
    __device__ __forceinline__ void operator()(int, int, const double (&)[2][4][2], double) const {}
};
template <class Out>
__device__ __forceinline__ void inv2_block_gj(double* psm, const DmmaPos& ps, SweepSync& sy, int d, Out&& final_tile) {
    using S = DmmaSmem;
    constexpr int LD = DM_LD;
    const int tid = threadIdx.x;
    double* Qs = psm + Inv2Smem::q;
    double* pv = psm + Inv2Smem::piv;
    double a[2][4][2];
    // a += Am Bm on the leading mrows x ncols of the output tile over k < kdepth (multiples of 8 / 8 / 4): whatever lies
    // beyond is padding -- exact zeros in one of the operands -- so the tiles and k-steps that are left out contribute
    // nothing.  Warp tile 16 x 32.
    auto product = [&](const double* Am, const double* Bm, int mrows, int ncols, int kdepth) {
        bool mv[2], nv[4];
#pragma unroll
        for (int ti = 0; ti < 2; ++ti) mv[ti] = 16 * ps.wr + 8 * ti < mrows;
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) nv[tj] = 32 * ps.wc + 8 * tj < ncols;
        if (!mv[0] || !nv[0]) return;
#pragma unroll 4
        for (int kk = 0; kk < kdepth; kk += 4) {
            double an[2], bw[4];
#pragma unroll
            for (int ti = 0; ti < 2; ++ti) an[ti] = Am[ps.row(ti) * LD + kk + ps.qc];
#pragma unroll
            for (int tj = 0; tj < 4; ++tj) bw[tj] = Bm[(kk + ps.qc) * LD + 32 * ps.wc + 8 * tj + ps.qr];
#pragma unroll
            for (int ti = 0; ti < 2; ++ti)
#pragma unroll
                for (int tj = 0; tj < 4; ++tj)
                    if (mv[ti] && nv[tj]) dmma(a[ti][tj][0], a[ti][tj][1], an[ti], bw[tj]);
        }
    };
    auto load_acc = [&](const double* T) {
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj) {
                const double2 v = *reinterpret_cast<const double2*>(T + ps.row(ti) * LD + ps.col(tj));
                a[ti][tj][0] = v.x;
                a[ti][tj][1] = v.y;
            }
    };
    auto store_acc = [&](double* T, double sign) {
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj)
                *reinterpret_cast<double2*>(T + ps.row(ti) * LD + ps.col(tj)) =
                    make_double2(sign * a[ti][tj][0], sign * a[ti][tj][1]);
    };
    auto zero_acc = [&]() {
#pragma unroll
        for (int ti = 0; ti < 2; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj) a[ti][tj][0] = a[ti][tj][1] = 0.0;
    };
#pragma unroll 1
    for (int kb = 0; kb < 2; ++kb) {
        const int o = 1 - kb, kn = min(DM_DP, d - DM_DP * kb), on = min(DM_DP, d - DM_DP * o);
        const int kn8 = (kn + 7) & ~7, on8 = (on + 7) & ~7, kn4 = (kn + 3) & ~3;      // valid extents of the two blocks
        double* Tkk = psm + Inv2Smem::at(kb, kb);
        double* Tok = psm + Inv2Smem::at(o, kb);
        double* Tko = psm + Inv2Smem::at(kb, o);
        double* Too = psm + Inv2Smem::at(o, o);
        load_acc(Tkk);
        __syncthreads();
        dmma_sweep(a, ps, psm, kn, sy);                       // a = Q (identity padded); ends with a barrier
        if (tid < DM_DP) pv[DM_DP * kb + tid] = (tid < ((kn + 3) & ~3)) ? psm[S::pinfo + tid] : 1.0;
        store_acc(Qs, 1.0);
        store_acc(Tkk, 1.0);
        if (kb == 1) final_tile(1, 1, a, 1.0);
        __syncthreads();
        zero_acc();
        product(Tok, Qs, on8, kn8, kn4);                      // T[o][kb] Q
        __syncthreads();
        store_acc(Tok, -1.0);                                 // new pivot column
        if (kb == 1) final_tile(0, 1, a, -1.0);
        __syncthreads();
        load_acc(Too);
        product(Tok, Tko, on8, on8, kn4);
        store_acc(Too, 1.0);
        if (kb == 1) final_tile(0, 0, a, 1.0);
        zero_acc();
        product(Qs, Tko, kn8, on8, kn4);                      // Q T[kb][o]
        __syncthreads();
        store_acc(Tko, 1.0);                                  // new pivot row
        if (kb == 1) final_tile(1, 0, a, 1.0);
        __syncthreads();
    }
}
__device__ __forceinline__ void inv2_block_gj(double* psm, const DmmaPos& ps, SweepSync& sy, int d) {
    inv2_block_gj(psm, ps, sy, d, Inv2NoOut{});
}

}  // namespace dagma
