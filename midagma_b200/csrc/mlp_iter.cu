// DagmaMLP / DagmaNonlinear, dims = [d, m1, 1], d <= 64: the WHOLE inner iteration -- h, forward, score, closed-form
// backward, torch-Adam step (reference: src/dagma/nonlinear.py:68-97, 139-159, 208-225) -- and any number of
// consecutive iterations as ONE persistent kernel.
//
// Why.  The launch sequence of mlp.cu is ~10 dependent graph nodes of 5-10 us for 1.3e8 flop of work: the iteration
// is bound by launch / drain latency, not by anything on the chip (C3: 69 us = 5 % of the FP64 peak).  Here an
// iteration is two phases separated by two grid barriers (bounded spins on a counter in L2; one CTA per SM, so the
// grid is co-resident), and the samples never leave the SM that loaded them:
//
//   phase A   The MLP is locally connected: out_j depends only on the m1 hidden units of node j.  A worker CTA owns
//             a SLICE of nodes (<= 40 rows of fc1.weight) and a GROUP of samples (MiPlan); its rows of W1 are staged
//             in shared memory once per iteration, its samples once per LAUNCH.
//               Z = W1 Xg^T                   DMMA.8x8x4, both operands in shared memory (row strides = 4 mod 16:
//                                             conflict-free fragments), unit = one m-tile x two n-tiles
//               H = sigmoid(Z + b1), out = sum_k H W2 + b2, res = out - x, S += res^2
//               dZ = res W2 H (1 - H), gb1 += dZ, gW2 += res H, gb2 += res      (one warp per row, fixed shuffle tree)
//               gW1 += dZ Xg                  DMMA over the samples: warps = 2 m-groups x 4 k-slices, every warp
//                                             keeps all tiles of its m-group in registers (0.4 fragment loads per
//                                             DMMA); the k-slices are added in a fixed order through shared memory
//             all sums UN-SCALED (the factor d / S needs the S of every group) into row `sample group` of `part`.
//             The last CTA meanwhile builds M = sI - A(W1) in the accumulator layout of the tensor-core sweep
//             (small_dmma.cuh), inverts it on chip and leaves h, log|det|, M^{-1} and sum |W1|.
//   phase B   every CTA adds the SG rows of `part` for its share of the parameters in a FIXED order (bit-reproducible),
//             applies d / S, l1, dh/dW1 and weight decay and takes the Adam step exactly as mlp_adam_kernel does.
//             The LAST CTA to reach the closing barrier writes the state block (S, score, objective, h < 0 latch,
//             step counter, ExponentialLR) before it opens it.
//
// Rows of X sharded over the GPUs of one box (one process per GPU): the same kernel on every GPU; between the two
// phases every CTA stores the fixed-order sums of ITS GPU's sample groups for its share of the parameters (and the
// GPU's S) into slot `rank` of every GPU's exchange buffer over NVLink peer memory (csrc/peer.cu), one more grid barrier,
// sequence-number flags, and phase B adds the slots in rank order -- the same order everywhere, so the replicas stay
// bit-identical and there is no collective call in the loop (protocol and memory ordering: csrc/lin_iter.cu).
//
// Stacks other than [d, m1, 1] stay on the launch sequence of mlp.cu.
#include "common.cuh"
#include "small_dmma.cuh"
#include "gemm_f64.cuh"
#include "mlp_state.h"
#include "../../include/dagma_b200.h"

namespace dagma {

constexpr int MI_NT = 256;        // threads per CTA
constexpr int MI_PR = 40;         // rows of fc1.weight (hidden units) per worker slice: <= 5 DMMA m-tiles
constexpr int MI_MTG = 3;         // m-tiles per warp in the gW1 product (two m-groups cover 6 >= 5)
constexpr int MI_NTT = 8;         // n-tiles of the gW1 product: d <= 64
constexpr int MI_KSL = 4;         // k-slices (sample ranges) of the gW1 product: warps = 2 m-groups x 4 k-slices
constexpr unsigned MI_SPIN_MAX = 1u << 22;
constexpr int MI_MAX_RANKS = 8;   // GPUs of one box that may share the rows of X
constexpr size_t MI_SMEM_CAP = 200 * 1024;

// How the work of an iteration is cut (host and device agree through this struct).  The MLP is locally connected:
// out_j depends only on the m1 hidden units of node j, so a worker CTA owns a SLICE of nodes (JN nodes = JN m1 <= 40
// rows of fc1.weight) and a GROUP of samples -- its rows of W1 are read once per iteration, its samples stay in
// shared memory across iterations, and a parameter is summed over only SG partial rows in phase B.
struct MiPlan {
    int JN, JS, SG, NSG, NSUB, G;       // nodes per slice, slices, sample groups, samples per group, per sub-group, CTAs
    int ldx, ldh;                       // row strides (doubles) of Xs / W1s and Hs: = 4 mod 16, conflict-free fragments
    int o_w1, o_xs, o_hs, o_rs, o_b1, o_w2, o_b2, o_tail, o_scr, o_red, total;   // shared-memory offsets (doubles)
    bool ok;
};
__host__ __device__ inline MiPlan mi_plan(int n, int d, int m1, int sms) {
    MiPlan L{};
    L.ok = false;
    if (d < 1 || d > 64 || m1 < 1 || m1 > MI_PR || n < 1 || sms < 2) return L;
    L.JN = MI_PR / m1 < d ? MI_PR / m1 : d;
    L.JS = (d + L.JN - 1) / L.JN;
    if (L.JS > sms - 1) return L;
    L.SG = (sms - 1) / L.JS;
    L.NSG = 16 * (((n + L.SG - 1) / L.SG + 15) / 16);
    L.SG = (n + L.NSG - 1) / L.NSG;
    L.ldx = 16 * ((d + 15) / 16) + 4;
    const int NT = (d + 7) / 8, MT = (L.JN * m1 + 7) / 8;
    const int fixed = 8 * MT * L.ldx + 2 * MI_PR + L.JN + (2 * MI_PR + L.JN + 2) + MI_KSL * MT * NT * 64 + 96 + 16;
    const int per_sample = L.ldx + 8 * MT + L.JN;
    int cap = ((int)(MI_SMEM_CAP / sizeof(double)) - fixed - 8 * MT * 4) / per_sample;
    cap = (cap / 16) * 16;
    if (cap < 16) return L;
    L.NSUB = L.NSG < cap ? L.NSG : cap;
    L.ldh = L.NSUB + 4;
    L.G = L.JS * L.SG + 1;
    L.o_w1 = 0;
    L.o_xs = L.o_w1 + 8 * MT * L.ldx;
    L.o_hs = L.o_xs + L.NSUB * L.ldx;
    L.o_rs = L.o_hs + 8 * MT * L.ldh;
    L.o_b1 = L.o_rs + L.JN * L.NSUB;
    L.o_w2 = L.o_b1 + MI_PR;
    L.o_b2 = L.o_w2 + MI_PR;
    L.o_tail = L.o_b2 + L.JN;                                   // [gb1 (40) | gW2 (40) | gb2 (JN) | S]
    L.o_scr = (L.o_tail + 2 * MI_PR + L.JN + 1 + 1) & ~1;
    L.o_red = L.o_scr + MI_KSL * MT * NT * 64;
    L.total = L.o_red + 96;
    L.ok = (size_t)L.total * sizeof(double) <= MI_SMEM_CAP;
    return L;
}

struct MlpIterArgs {
    MlpState* st;
    double *theta, *m, *v;                 // [total] parameters and Adam moments
    const double* X;                       // [n][d] row-major
    int n, n_total, d, m1;
    double* part;                          // [SG][total] un-scaled gradient sums per sample group, then [SG * JS] partial S
    double* Minv;                          // [d][d]
    unsigned* sync;                        // [0] arrivals at the grid barriers of this launch [2] error (zeroed by the host per launch)
    int iters, sms;
    int h_stage;                           // doubles of the h CTA's staging buffer for fc1.weight
    // rows of X sharded over `nranks` GPUs of one box (nranks = 1: none of this is touched) -- the protocol of
    // csrc/lin_iter.cu: xchg[r] = rank r's exchange buffer as mapped into THIS process, [2 (parity)][nranks (writer)]
    // [total + 1] doubles (the un-scaled gradient sums of the writer's rows, then its S); flags[r]: rank r's [nranks]
    // sequence numbers; seq: this rank's count of exchanged iterations
    int rank, nranks;
    double* xchg[MI_MAX_RANKS];
    unsigned* flags[MI_MAX_RANKS];
    unsigned* seq;
};

#ifdef DAGMA_MLP_TRACE
// debug build only: %globaltimer stamps (ns) of the LAST iteration of a launch.  [0] worker 0: iteration starts
// [1] worker 0: parameters staged  [2] forward done  [3] res / dZ done  [4] gW1 done  [5] partial row written
// [6] h CTA: A built  [7] h CTA: sweep done  [8] h CTA: done  [9] worker 0 past barrier 1  [10] S known  [11] Adam done
// [12] past barrier 2
__device__ unsigned long long g_mlp_trace[16];
__device__ __forceinline__ unsigned long long mi_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MI_STAMP(slot, cond) do { if ((cond) && threadIdx.x == 0) g_mlp_trace[slot] = mi_gtime(); } while (0)
#else
#define MI_STAMP(slot, cond) do { } while (0)
#endif

// grid barrier number `k` of this launch (k = 0, 1, ...; the host zeroes the counter before every launch): ONE atomic
// per CTA and no reset -- the arrival of the last CTA is what the others are polling for
__device__ __forceinline__ void mi_grid_barrier(unsigned* sync, unsigned G, unsigned& k) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = G * (k + 1);
        __threadfence();
        atomicAdd(sync, 1u);
        unsigned spins = 0;
        while (*(volatile unsigned*)sync < target) {
            if ((++spins & 63u) == 0u) {
                if (*(volatile unsigned*)(sync + 2)) break;
                if (spins > MI_SPIN_MAX) { atomicExch(sync + 2, 1u); break; }
            }
        }
        __threadfence();
    }
    ++k;
    __syncthreads();
}

// ---------------------------------------------------------------- the CTA that owns h
// Shared memory of this CTA: the sweep's DmmaSmem (A sits where the fit kernel keeps -cov, which the sweep does not
// touch), then a staging buffer for fc1.weight, filled by cp.async in chunks of whole nodes.
constexpr int MI_H_A = DmmaSmem::ncov;                       // A, [64][68]
constexpr int MI_H_STAGE = (DmmaSmem::total + 1) & ~1;       // fc1.weight rows of a chunk of nodes
__device__ __forceinline__ void mi_h_role(const MlpIterArgs& P, double* psm, SweepSync& sy, int stage_doubles) {
    using S = DmmaSmem;
    const int tid = threadIdx.x, d = P.d, m1 = P.m1;
    const DmmaPos ps(tid);
    double* red = psm + S::red;
    double* As = psm + MI_H_A;
    double* stg = psm + MI_H_STAGE;
    const double* W1 = P.theta;
    const double s = P.st->s;
    double scale = 1.0;
    if (s > 0.0 && isfinite(s)) {              // s / 2^e in (0.5, 1]: exact, keeps the publish identity in its accurate range
        int e = 0;
        const double f = frexp(s, &e);
        if (f == 0.5) --e;
        scale = ldexp(1.0, e);
    }
    const double inv_scale = 1.0 / scale;
    // A[i][j] = sum_k W1[j m1 + k][i]^2 and sum |W1|, node chunk by node chunk (d is even or the rows are copied singly)
    double l1 = 0.0;
    const int rowlen = m1 * d;                                   // doubles of one node's rows (contiguous in theta)
    const int jc_max = max(1, stage_doubles / rowlen);
    const bool vec = ((rowlen & 1) == 0);
    for (int jb = 0; jb < d; jb += jc_max) {
        const int jc = min(jc_max, d - jb), cnt = jc * rowlen;
        const double* src = W1 + (size_t)jb * rowlen;
        __syncthreads();                                         // the previous chunk has been consumed
        if (vec) {
            for (int e = 2 * tid; e < cnt; e += 2 * MI_NT) cp_async16(smem_u32(stg + e), src + e, true);
            cp_async_commit();
            cp_async_wait<0>();
        } else {
            for (int e = tid; e < cnt; e += MI_NT) stg[e] = __ldcg(src + e);
        }
        __syncthreads();
        for (int e = tid; e < jc * d; e += MI_NT) {
            const int jl = e / d, i = e - jl * d;
            double x = 0.0;
            for (int k = 0; k < m1; ++k) {
                const double w = stg[(jl * m1 + k) * d + i];
                x = fma(w, w, x);
                l1 += fabs(w);
            }
            As[i * DM_LD + jb + jl] = x;
        }
    }
    __syncthreads();
    // M = (s I - A) / scale in the accumulator layout, identity padded
    double a[2][4][2];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = ps.row(ti), c = ps.col(tj) + e;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < d && c < d) v = (((r == c) ? s : 0.0) - As[r * DM_LD + c]) * inv_scale;
                a[ti][tj][e] = v;
            }
    MI_STAMP(6, true);
    dmma_sweep(a, ps, psm, d, sy);
    MI_STAMP(7, true);
    double ld = 0.0, zero = 0.0;
    bool badpiv = false;
    if (tid < ((d + 3) & ~3)) {
        const double p = psm[S::pinfo + tid];
        ld = (double)((tid & 3) - 2) * log(fabs(p));
        badpiv = !(p > 0.0);
    }
    double mn = INFINITY;
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = ps.row(ti), c = ps.col(tj) + e;
                if (r < d && c < d) {
                    const double mi = a[ti][tj][e] * inv_scale;
                    mn = fmin(mn, mi);
                    P.Minv[(size_t)r * d + c] = mi;
                }
            }
    block_sum3<MI_NT>(ld, l1, zero, red, tid);
    mn = block_min<MI_NT>(mn, red, tid);
    const int anybad = __syncthreads_or(badpiv);
    if (tid == 0) {
        const double lad = ld + (double)d * log(scale);
        P.st->logabsdet = lad;
        P.st->h = -lad + (double)d * log(s);
        P.st->min_entry = mn;
        P.st->info = anybad ? 1 : ((mn + 1e-16 < 0.0) ? 2 : 0);
        P.st->l1 = l1;
    }
    MI_STAMP(8, true);
}

// 1 / (1 + exp(-z)) without the slow-path branches of the IEEE division (1 + e is in the normal range after the clamp)
__device__ __forceinline__ double mi_sigmoid(double z) { return fast_rcp(1.0 + exp(fmin(-z, 700.0))); }

// gW1 tiles of one warp (m-tiles mbase .. mbase + mcnt, all NTC n-tiles) += dZ[:, k-slice] X[k-slice, :]
template <int NTC>
__device__ __forceinline__ void mi_gw1(double (&g)[MI_MTG][MI_NTT][2], const double* Hs, const double* Xs, int ldh, int ldx,
                                       int mbase, int mcnt, int ks0, int nks, int qr, int qc) {
    const double* ap = Hs + (8 * mbase + qr) * ldh + qc;
    const double* bp = Xs + qc * ldx + qr;
    for (int ks = ks0; ks < nks; ks += MI_KSL) {
        double bv[NTC];
#pragma unroll
        for (int j = 0; j < NTC; ++j) bv[j] = bp[4 * ks * ldx + 8 * j];
#pragma unroll
        for (int i = 0; i < MI_MTG; ++i)
            if (i < mcnt) {
                const double av = ap[8 * i * ldh + 4 * ks];
#pragma unroll
                for (int j = 0; j < NTC; ++j) dmma(g[i][j][0], g[i][j][1], av, bv[j]);
            }
    }
}

// ---------------------------------------------------------------- a worker CTA: node slice js, sample group sg
__device__ __forceinline__ void mi_load_x(const MlpIterArgs& P, const MiPlan& L, double* Xs, int s_lo, int cnt, int rows) {
    const int d = P.d;
    for (int e = threadIdx.x; e < rows * L.ldx; e += MI_NT) {
        const int sr = e / L.ldx, c = e - sr * L.ldx;
        Xs[e] = (sr < cnt && c < d) ? __ldg(P.X + (size_t)(s_lo + sr) * d + c) : 0.0;
    }
}

__device__ __forceinline__ void mi_worker_role(const MlpIterArgs& P, const MiPlan& L, double* sm, int js, int sg,
                                               bool x_resident) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, qr = lane >> 2, qc = lane & 3;
    const int d = P.d, m1 = P.m1, n = P.n, PP = d * m1, W = PP * d + 2 * PP + d;
    const int j0 = js * L.JN, jn = min(L.JN, d - j0), p0 = j0 * m1, PR = jn * m1;
    const int MT = (PR + 7) / 8, NT = (d + 7) / 8, KTd = (d + 3) / 4;
    const int ldx = L.ldx, ldh = L.ldh;
    double *W1s = sm + L.o_w1, *Xs = sm + L.o_xs, *Hs = sm + L.o_hs, *Rs = sm + L.o_rs, *b1s = sm + L.o_b1;
    double *W2s = sm + L.o_w2, *b2s = sm + L.o_b2, *tail = sm + L.o_tail, *scr = sm + L.o_scr, *red = sm + L.o_red;
    const int s_beg = sg * L.NSG, s_end = min(n, s_beg + L.NSG);

    // ---- this iteration's parameters of the slice (theta changes between iterations: L2 loads)
    for (int e = tid; e < 8 * MT * ldx; e += MI_NT) {
        const int r = e / ldx, c = e - r * ldx;
        W1s[e] = (r < PR && c < d) ? __ldcg(P.theta + (size_t)(p0 + r) * d + c) : 0.0;
    }
    for (int e = tid; e < PR; e += MI_NT) {
        b1s[e] = __ldcg(P.theta + (size_t)PP * d + p0 + e);
        W2s[e] = __ldcg(P.theta + (size_t)PP * d + PP + p0 + e);
    }
    for (int e = tid; e < jn; e += MI_NT) b2s[e] = __ldcg(P.theta + (size_t)PP * d + 2 * PP + j0 + e);
    for (int e = tid; e < 2 * MI_PR + L.JN + 1; e += MI_NT) tail[e] = 0.0;

    // gW1 accumulators: warp = (m-group, k-slice); m-group 0 owns the m-tiles [0, mg0), group 1 the rest
    const int mgrp = warp >> 2, ksl = warp & 3;            // the two warps of a sub-partition (w, w + 4) are one of each m-group
    const int mg0 = (MT + 1) / 2;
    const int mbase = mgrp ? mg0 : 0, mcnt = mgrp ? MT - mg0 : mg0;
    double g[MI_MTG][MI_NTT][2];
#pragma unroll
    for (int i = 0; i < MI_MTG; ++i)
#pragma unroll
        for (int j = 0; j < MI_NTT; ++j) g[i][j][0] = g[i][j][1] = 0.0;

    for (int s_lo = s_beg; s_lo < s_end; s_lo += L.NSUB) {
        const int cnt = min(L.NSUB, s_end - s_lo), cntp = (cnt + 15) & ~15;      // samples of the sub-group, padded
        __syncthreads();                                   // parameters staged; the previous sub-group is done with Xs / Hs / Rs
        if (!x_resident) {
            mi_load_x(P, L, Xs, s_lo, cnt, cntp);
            __syncthreads();
        }
        MI_STAMP(1, blockIdx.x == 0);
        // ---- Z = W1 Xc^T (K = d), H = sigmoid(Z + b1): unit = one m-tile x two n-tiles
        const int ntp = cntp / 16, nun = MT * ntp;
        for (int u0 = warp; u0 < nun; u0 += 2 * (MI_NT / 32)) {        // two units per round: four accumulator chains
            const int u1 = u0 + MI_NT / 32;
            const bool two = u1 < nun;
            const int mt0 = u0 % MT, np0 = u0 / MT, mt1 = two ? u1 % MT : mt0, np1 = two ? u1 / MT : np0;
            const double* ap0 = W1s + (8 * mt0 + qr) * ldx + qc;
            const double* bp0 = Xs + (16 * np0 + qr) * ldx + qc;
            const double* ap1 = W1s + (8 * mt1 + qr) * ldx + qc;
            const double* bp1 = Xs + (16 * np1 + qr) * ldx + qc;
            double z[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll 2
            for (int ks = 0; ks < KTd; ++ks) {
                const double a0 = ap0[4 * ks], b00 = bp0[4 * ks], b01 = bp0[8 * ldx + 4 * ks];
                const double a1 = ap1[4 * ks], b10 = bp1[4 * ks], b11 = bp1[8 * ldx + 4 * ks];
                dmma(z[0][0], z[0][1], a0, b00);
                dmma(z[0][2], z[0][3], a0, b01);
                dmma(z[1][0], z[1][1], a1, b10);
                dmma(z[1][2], z[1][3], a1, b11);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int p = 8 * (q ? mt1 : mt0) + qr, np = q ? np1 : np0;
                if (p < PR && (q == 0 || two)) {
                    const double bb = b1s[p];
                    double* hp = Hs + p * ldh + 16 * np + 2 * qc;
                    *reinterpret_cast<double2*>(hp) = make_double2(mi_sigmoid(z[q][0] + bb), mi_sigmoid(z[q][1] + bb));
                    *reinterpret_cast<double2*>(hp + 8) = make_double2(mi_sigmoid(z[q][2] + bb), mi_sigmoid(z[q][3] + bb));
                }
            }
        }
        __syncthreads();
        MI_STAMP(2, blockIdx.x == 0);
        // ---- out, res, S
        double sq = 0.0, z1 = 0.0, z2 = 0.0;
        for (int e = tid; e < jn * cntp; e += MI_NT) {
            const int jl = e / cntp, sr = e - jl * cntp;
            double o = b2s[jl];
            for (int k = 0; k < m1; ++k) o = fma(Hs[(jl * m1 + k) * ldh + sr], W2s[jl * m1 + k], o);
            const double r = (sr < cnt) ? o - Xs[sr * ldx + j0 + jl] : 0.0;
            Rs[jl * L.NSUB + sr] = r;
            sq = fma(r, r, sq);
        }
        block_sum3<MI_NT>(sq, z1, z2, red, tid);          // (its barriers publish Rs)
        if (tid == 0) tail[2 * MI_PR + L.JN] += sq;
        // ---- dZ (over H), gb1, gW2: four threads per hidden unit, each every fourth sample (no shuffles, no divergent
        // trip counts); the four partial sums are added in a fixed order
        {
            constexpr int SEG = 4;
            double* segp = scr;                                 // [PR][SEG][2]: the scratch of the gW1 reduction is free here
            if (tid < PR * SEG) {
                const int p = tid / SEG, sgm = tid - p * SEG, jl = p / m1;
                const double w2 = W2s[p];
                double* hrow = Hs + p * ldh;
                const double* rrow = Rs + jl * L.NSUB;
                double gw2 = 0.0, gb1 = 0.0;
#pragma unroll 4
                for (int sr = sgm; sr < cntp; sr += SEG) {
                    const double hh = hrow[sr], r = rrow[sr];
                    const double dz = r * w2 * hh * (1.0 - hh);
                    hrow[sr] = dz;
                    gw2 = fma(r, hh, gw2);
                    gb1 += dz;
                }
                segp[2 * tid] = gb1;
                segp[2 * tid + 1] = gw2;
            }
            __syncthreads();
            if (tid < PR) {
                double gb1 = 0.0, gw2 = 0.0;
#pragma unroll
                for (int q = 0; q < SEG; ++q) {
                    gb1 += segp[2 * (tid * SEG + q)];
                    gw2 += segp[2 * (tid * SEG + q) + 1];
                }
                tail[tid] += gb1;
                tail[MI_PR + tid] += gw2;
            }
        }
        for (int jl = warp; jl < jn; jl += MI_NT / 32) {
            double gsum = 0.0;
            for (int sr = lane; sr < cntp; sr += 32) gsum += Rs[jl * L.NSUB + sr];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, off);
            if (lane == 0) tail[2 * MI_PR + jl] += gsum;
        }
        __syncthreads();
        MI_STAMP(3, blockIdx.x == 0);
        // ---- gW1 += dZ Xc (K = samples): every warp multiplies its k-slice into all tiles of its m-group
        switch (NT) {
            case 1: mi_gw1<1>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 2: mi_gw1<2>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 3: mi_gw1<3>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 4: mi_gw1<4>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 5: mi_gw1<5>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 6: mi_gw1<6>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            case 7: mi_gw1<7>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
            default: mi_gw1<8>(g, Hs, Xs, ldh, ldx, mbase, mcnt, ksl, cntp / 4, qr, qc); break;
        }
    }
    // ---- k-slices summed in a fixed order -> this sample group's row of `part`
    __syncthreads();
    MI_STAMP(4, blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < MI_MTG; ++i)
#pragma unroll
        for (int j = 0; j < MI_NTT; ++j)
            if (i < mcnt && j < NT)
                *reinterpret_cast<double2*>(scr + ((size_t)(ksl * MT + mbase + i) * NT + j) * 64 + 2 * lane) = make_double2(g[i][j][0], g[i][j][1]);
    __syncthreads();
    double* prow = P.part + (size_t)sg * W;
    for (int idx = tid; idx < MT * NT * 64; idx += MI_NT) {
        const int tile = idx >> 6, w = idx & 63, ln = w >> 1, e = w & 1;
        const int mt = tile / NT, j = tile - mt * NT;
        const int pl = 8 * mt + (ln >> 2), i = 8 * j + 2 * (ln & 3) + e;
        double sum = scr[idx];
#pragma unroll
        for (int k = 1; k < MI_KSL; ++k) sum += scr[(size_t)k * MT * NT * 64 + idx];
        if (pl < PR && i < d) prow[(size_t)(p0 + pl) * d + i] = sum;
    }
    for (int e = tid; e < PR; e += MI_NT) {
        prow[(size_t)PP * d + p0 + e] = tail[e];
        prow[(size_t)PP * d + PP + p0 + e] = tail[MI_PR + e];
    }
    for (int e = tid; e < jn; e += MI_NT) prow[(size_t)PP * d + 2 * PP + j0 + e] = tail[2 * MI_PR + e];
    if (tid == 0) P.part[(size_t)L.SG * W + sg * L.JS + js] = tail[2 * MI_PR + L.JN];
    MI_STAMP(5, blockIdx.x == 0);
}

// the state block as every thread reads it at the START of an iteration (it is rewritten after the first barrier)
struct MiScalars {
    double mu, lr, lambda1, lambda2, beta1, beta2, gamma;
    int step;
};
__device__ __forceinline__ MiScalars mi_read_state(const MlpState* stp) {
    const volatile MlpState* st = stp;
    return MiScalars{st->mu, st->lr, st->lambda1, st->lambda2, st->beta1, st->beta2, st->lr_gamma, st->step};
}

// ---------------------------------------------------------------- phase B: fixed-order sums + Adam (mlp_adam_kernel)
// the SG rows of `part` for one entry, added in order (sixteen loads in flight)
__device__ __forceinline__ double mi_group_sum(const double* src, int SG, int W) {
    double acc = 0.0;
    int c = 0;
    for (; c + 16 <= SG; c += 16) {
        double t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = __ldcg(src + (size_t)(c + u) * W);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += t[u];
    }
    {
        double t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = (c + u < SG) ? __ldcg(src + (size_t)(c + u) * W) : 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (c + u < SG) acc += t[u];
    }
    return acc;
}

// sharded rows: this GPU's sums for the CTA's share of the parameters (and, by CTA 0, its S) to every GPU's buffer
__device__ __forceinline__ void mi_publish_sums(const MlpIterArgs& P, const MiPlan& L, int cta, double S_local, unsigned seq) {
    const int tid = threadIdx.x, d = P.d, PP = d * P.m1, W = PP * d + 2 * PP + d;
    const size_t slot = ((size_t)(seq & 1u) * P.nranks + P.rank) * (W + 1);
    for (int e = cta * MI_NT + tid; e < W; e += L.G * MI_NT) {
        const double acc = mi_group_sum(P.part + e, L.SG, W);
        for (int r = 0; r < P.nranks; ++r) P.xchg[r][slot + e] = acc;
    }
    if (cta == 0 && tid == 0)
        for (int r = 0; r < P.nranks; ++r) P.xchg[r][slot + W] = S_local;
    __threadfence_system();
}

__device__ __forceinline__ void mi_update(const MlpIterArgs& P, const MiPlan& L, int cta, double S, const MiScalars& sc,
                                          double bc1, double bc2, unsigned seq) {
    const int tid = threadIdx.x, d = P.d, m1 = P.m1, PP = d * m1;
    const int W = PP * d + 2 * PP + d, SG = L.SG;
    const double mu = sc.mu, lr = sc.lr, b1 = sc.beta1, b2 = sc.beta2;
    const double wd = mu * sc.lambda2, l1c = mu * sc.lambda1;
    const double gs = mu * (double)d / S;
    const double step_size = lr / bc1, bc2s = sqrt(bc2);
    const int nW1 = PP * d;
    const double* xs = (P.nranks > 1) ? P.xchg[P.rank] + (size_t)(seq & 1u) * P.nranks * (W + 1) : nullptr;
    for (int e = cta * MI_NT + tid; e < W; e += L.G * MI_NT) {
        // everything the step needs travels together with the partial sums
        const double p = P.theta[e], mo = P.m[e], vo = P.v[e];
        double hterm = 0.0;
        if (e < nW1) {
            const int row = e / d, i = e - row * d, j = row / m1;
            hterm = __ldcg(P.Minv + (size_t)j * d + i);
        }
        double acc;
        if (P.nranks == 1) acc = mi_group_sum(P.part + e, SG, W);
        else {                                                    // the GPUs' sums in rank order: the same on every GPU
            acc = 0.0;
            for (int q = 0; q < P.nranks; ++q) acc += __ldcg(xs + (size_t)q * (W + 1) + e);
        }
        double g = gs * acc;
        if (e < nW1) {
            const double sg = (p > 0.0) ? 1.0 : ((p < 0.0) ? -1.0 : 0.0);
            g = fma(l1c, sg, g);
            g = fma(2.0 * p, hterm, g);
        }
        g = fma(wd, p, g);
        const double mn = mo + (g - mo) * (1.0 - b1);            // exp_avg.lerp_(grad, 1 - beta1)
        const double vn = fma(vo, b2, (1.0 - b2) * g * g);
        P.m[e] = mn;
        P.v[e] = vn;
        const double denom = sqrt(vn) / bc2s + 1e-8;
        P.theta[e] = p - step_size * (mn / denom);
    }
}

__global__ void __launch_bounds__(MI_NT, 1) mlp_iter_kernel(const MlpIterArgs P) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_S, s_bc[2];
    __shared__ MiScalars s_sc;
    __shared__ unsigned s_seq;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    const MiPlan L = mi_plan(P.n, P.d, P.m1, P.sms);
    const int d = P.d, PP = d * P.m1, W = PP * d + 2 * PP + d;
    const bool hcta = (cta == G - 1);
    const int js = cta % L.JS, sg = cta / L.JS;
    const bool x_resident = (L.NSG <= L.NSUB);
    SweepSync sy{smem_u32(sm + DmmaSmem::mbar), 0u};
    if (hcta) {
        if (tid == 0) mbar_init(sy.bar, DM_NT / 32);
    } else {
        const int MT = (min(L.JN, d - js * L.JN) * P.m1 + 7) / 8;
        for (int e = tid; e < 8 * MT * L.ldh; e += MI_NT) sm[L.o_hs + e] = 0.0;      // the padding rows of Hs stay zero
        if (x_resident) {
            const int s_beg = sg * L.NSG, cnt = min(L.NSG, P.n - s_beg);
            mi_load_x(P, L, sm + L.o_xs, s_beg, cnt, (cnt + 15) & ~15);
        }
    }
    __syncthreads();
    unsigned bar_k = 0;
    for (int it = 0; it < P.iters; ++it) {
        if (*(volatile int32_t*)&P.st->halted != 0 || *(volatile unsigned*)(P.sync + 2) != 0u) break;   // uniform over the grid
        MI_STAMP(0, cta == 0);
        if (hcta) mi_h_role(P, sm, sy, P.h_stage);
        else mi_worker_role(P, L, sm, js, sg, x_resident);
#ifdef DAGMA_MLP_TRACE
        if (tid == 0) {                                               // [13] last arrival at barrier 1, [14] its CTA, [15] first arrival
            const unsigned long long t = mi_gtime();
            if (cta == 0) { g_mlp_trace[13] = 0ull; g_mlp_trace[15] = ~0ull; }
            if (atomicMax(&g_mlp_trace[13], t) < t) g_mlp_trace[14] = (unsigned long long)cta;
            atomicMin(&g_mlp_trace[15], t);
        }
#endif
        if (tid == 0) {
            s_sc = mi_read_state(P.st);                               // before the barrier: CTA 0 rewrites it after it;
                                                                      // one thread per CTA (36 k readers of one line slow the barrier down)
            s_seq = (P.nranks > 1) ? *(volatile unsigned*)P.seq + 1u : 0u;
        }
        mi_grid_barrier(P.sync, (unsigned)G, bar_k);
        MI_STAMP(9, cta == 0);
        const MiScalars sc = s_sc;
        // ---- S in a fixed order (every CTA for itself); h < 0: no step (nonlinear.py:215-217)
        if (tid < 32) {
            const double* sp = P.part + (size_t)L.SG * W;
            const int ns = L.SG * L.JS;                   // <= SMs - 1 partial sums: at most eight per lane
            double t[8], acc = 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = (tid + 32 * u < ns) ? __ldcg(sp + tid + 32 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += t[u];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (tid == 0) s_S = acc;
        } else if (tid == 32 || tid == 64) {
            // Adam bias corrections 1 - beta^t as -expm1(t log beta), one lane each, beside the loads of S: pow(double,
            // double) is ~1000 instructions, this is ~80 and agrees with it to a few ulp
            s_bc[tid == 64] = -expm1((double)(sc.step + 1) * log(tid == 64 ? sc.beta2 : sc.beta1));
        }
        const double h = *(volatile double*)&P.st->h;
        __syncthreads();
        const unsigned seq = s_seq;
        if (P.nranks > 1) {
            // ---- rows sharded over GPUs: exchange the sums of this GPU's rows (see the file header)
            mi_publish_sums(P, L, cta, s_S, seq);
            mi_grid_barrier(P.sync, (unsigned)G, bar_k);              // every CTA's peer stores are fenced
            if (cta == 0 && tid < P.nranks && tid != P.rank) {
                __threadfence_system();                               // release
                *(volatile unsigned*)(P.flags[tid] + P.rank) = seq;
            }
            if (tid < P.nranks && tid != P.rank) {
                unsigned spins = 0;
                while ((int)(*(volatile unsigned*)(P.flags[P.rank] + tid) - seq) < 0) {
                    if ((++spins & 63u) == 0u) {
                        if (*(volatile unsigned*)(P.sync + 2)) break;
                        if (spins > MI_SPIN_MAX) { atomicExch(P.sync + 2, 2u); break; }
                    }
                }
                __threadfence_system();
            }
            __syncthreads();
            if (tid == 0) {                                           // S over all GPUs, rank order
                const int Wt = PP * d + 2 * PP + d;
                const double* xs = P.xchg[P.rank] + (size_t)(seq & 1u) * P.nranks * (Wt + 1);
                double S_all = 0.0;
                for (int q = 0; q < P.nranks; ++q) S_all += __ldcg(xs + (size_t)q * (Wt + 1) + Wt);
                s_S = S_all;
            }
            __syncthreads();
        }
        const double S = s_S;
        const bool neg = h < 0.0;
        MI_STAMP(10, cta == 0);
        // the state block of the next iteration: everybody read the current one before the first barrier, nobody reads
        // it again before the second
        if (cta == 0 && tid == 0) {
            volatile MlpState* st = P.st;
            const double score = 0.5 * (double)d * log(S / (double)P.n_total);
            if (P.nranks > 1) *(volatile unsigned*)P.seq = seq;
            st->S = S;
            st->score = score;
            st->obj = sc.mu * (score + sc.lambda1 * st->l1) + h;
            if (neg) st->halted = 1;
            else {
                st->step = sc.step + 1;
                if (sc.gamma != 1.0 && ((sc.step + 1) % 1000) == 0) st->lr = sc.lr * sc.gamma;
            }
        }
        if (!neg) mi_update(P, L, cta, S, sc, s_bc[0], s_bc[1], seq);
        MI_STAMP(11, cta == 0);
        mi_grid_barrier(P.sync, (unsigned)G, bar_k);
        MI_STAMP(12, cta == 0);
    }
    if (tid == 0 && cta == 0 && *(volatile unsigned*)(P.sync + 2) != 0u) P.st->info = 99;   // a barrier timed out
}

}  // namespace dagma

using namespace dagma;

static int mi_sms() {
    static int sms = -1;
    if (sms < 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
            sms = v;
        else
            return 0;
    }
    return sms;
}

// 1 when the fused iteration covers dims = [d, m1, 1] with n rows on this device
extern "C" int dagma_mlp_iter_supported(int n, int d, int m1) { return mi_plan(n, d, m1, mi_sms()).ok ? 1 : 0; }

// doubles of the partial-sum workspace `part_dev` of dagma_mlp_iter_f64
extern "C" size_t dagma_mlp_iter_workspace_doubles(int n, int d, int m1) {
    const MiPlan L = mi_plan(n, d, m1, mi_sms());
    if (!L.ok) return 0;
    const size_t W = (size_t)d * m1 * d + 2 * (size_t)d * m1 + d;
    return (size_t)L.SG * W + (size_t)L.SG * L.JS + 8;
}

static int mlp_iter_launch(cudaStream_t stream, int n, int n_total, int d, int m1, int iters, void* state_dev,
                           double* theta_dev, double* m_dev, double* v_dev, const double* x_dev, double* part_dev,
                           double* minv_dev, unsigned* sync_dev, int rank, int nranks, void* const* xchg_ptrs,
                           void* const* flag_ptrs, unsigned* seq_dev) {
    DAGMA_REQUIRE(state_dev && theta_dev && m_dev && v_dev && x_dev && part_dev && minv_dev && sync_dev, "null pointer");
    const MiPlan L = mi_plan(n, d, m1, mi_sms());
    DAGMA_REQUIRE(iters >= 0 && L.ok, "shape not supported by the fused iteration (dagma_mlp_iter_supported)");
    size_t smem = (size_t)L.total * sizeof(double);
    // the h CTA: the sweep's buffers + as much of fc1.weight as fits beside them (whole nodes, >= one)
    const size_t w1 = (size_t)d * m1 * d, node = (size_t)m1 * d;
    size_t stage = (MI_SMEM_CAP - MI_H_STAGE * sizeof(double)) / sizeof(double);
    if (stage > w1) stage = w1;
    if (stage < node) stage = node;
    stage = (stage + 1) & ~(size_t)1;
    const size_t hbytes = (MI_H_STAGE + stage) * sizeof(double);
    if (smem < hbytes) smem = hbytes;
    DAGMA_REQUIRE(smem <= 227 * 1024, "fc1.weight rows of one node do not fit beside the sweep");
    static size_t attr = 0;
    if (smem > attr) {
        DAGMA_CUDA_OK(cudaFuncSetAttribute(mlp_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    MlpIterArgs A{};
    A.st = (MlpState*)state_dev;
    A.theta = theta_dev; A.m = m_dev; A.v = v_dev; A.X = x_dev;
    A.n = n; A.n_total = n_total; A.d = d; A.m1 = m1;
    A.part = part_dev; A.Minv = minv_dev; A.sync = sync_dev;
    A.iters = iters; A.sms = mi_sms(); A.h_stage = (int)stage;
    A.rank = rank; A.nranks = nranks; A.seq = seq_dev;
    for (int r = 0; r < nranks && nranks > 1; ++r) {
        A.xchg[r] = (double*)xchg_ptrs[r];
        A.flags[r] = (unsigned*)flag_ptrs[r];
    }
    // 32-bit arrival counter: up to G * 3 * iters arrivals per launch
    DAGMA_REQUIRE((double)L.G * 3.0 * (double)iters < 4.0e9, "too many iterations for one launch");
    DAGMA_CUDA_OK(cudaMemsetAsync(sync_dev, 0, 4 * sizeof(unsigned), stream));
    mlp_iter_kernel<<<L.G, MI_NT, smem, stream>>>(A);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dagma_mlp_iter_f64(dagma_stream_t stream, int n, int n_total, int d, int m1, int iters, void* state_dev,
                                  double* theta_dev, double* m_dev, double* v_dev, const double* x_dev, double* part_dev,
                                  double* minv_dev, unsigned* sync_dev) {
    return mlp_iter_launch((cudaStream_t)stream, n, n_total, d, m1, iters, state_dev, theta_dev, m_dev, v_dev, x_dev,
                           part_dev, minv_dev, sync_dev, 0, 1, nullptr, nullptr, nullptr);
}

static size_t mi_exchange_data_bytes(int d, int m1, int nranks) {
    const size_t W = (size_t)d * m1 * d + 2 * (size_t)d * m1 + d;
    return (((size_t)2 * nranks * (W + 1) * sizeof(double)) + 127) & ~(size_t)127;
}

extern "C" size_t dagma_mlp_iter_exchange_bytes(int d, int m1, int nranks) {
    if (d < 1 || d > 64 || m1 < 1 || m1 > MI_PR || nranks < 2 || nranks > MI_MAX_RANKS) return 0;
    return mi_exchange_data_bytes(d, m1, nranks) + 256;
}

extern "C" int dagma_mlp_iter_sharded_f64(dagma_stream_t stream, int n_local, int n_total, int d, int m1, int iters,
                                          void* state_dev, double* theta_dev, double* m_dev, double* v_dev,
                                          const double* x_dev, double* part_dev, double* minv_dev, unsigned* sync_dev,
                                          int rank, int nranks, void* const* exchange_ptrs) {
    DAGMA_REQUIRE(nranks >= 2 && nranks <= MI_MAX_RANKS && rank >= 0 && rank < nranks && exchange_ptrs, "bad rank arguments");
    // layout of a GPU's exchange allocation (dagma_mlp_iter_exchange_bytes): [2][nranks][total + 1] doubles, then
    // [nranks] flags, then this GPU's own sequence counter
    void* xp[MI_MAX_RANKS];
    void* fp[MI_MAX_RANKS];
    const size_t data = mi_exchange_data_bytes(d, m1, nranks);
    for (int r = 0; r < nranks; ++r) {
        DAGMA_REQUIRE(exchange_ptrs[r], "null exchange pointer");
        xp[r] = exchange_ptrs[r];
        fp[r] = (char*)exchange_ptrs[r] + data;
    }
    unsigned* seq = (unsigned*)((char*)exchange_ptrs[rank] + data) + 32;
    return mlp_iter_launch((cudaStream_t)stream, n_local, n_total, d, m1, iters, state_dev, theta_dev, m_dev, v_dev, x_dev,
                           part_dev, minv_dev, sync_dev, rank, nranks, xp, fp, seq);
}

#ifdef DAGMA_MLP_TRACE
extern "C" int dagma_debug_mlp_trace(unsigned long long* out_host) {
    DAGMA_CUDA_OK(cudaDeviceSynchronize());
    DAGMA_CUDA_OK(cudaMemcpyFromSymbol(out_host, g_mlp_trace, sizeof(unsigned long long) * 16));
    return 0;
}
#endif
