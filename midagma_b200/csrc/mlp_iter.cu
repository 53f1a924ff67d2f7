// DagmaMLP / DagmaNonlinear, dims = [d, m1, 1], d <= 64: the WHOLE inner iteration -- h, forward, score, closed-form
// backward, torch-Adam step (reference: src/dagma/nonlinear.py:68-97, 139-159, 208-225) -- and any number of
// consecutive iterations as ONE persistent kernel.
//
// Why.  The launch sequence of mlp.cu is ~10 dependent graph nodes of 5-10 us for 1.3e8 flop of work: the iteration
// is bound by launch / drain latency, not by anything on the chip (C3: 69 us = 5 % of the FP64 peak).  Here an
// iteration is two phases separated by two grid barriers (bounded spins on a counter in L2; one CTA per SM, so the
// grid is co-resident), and the samples never leave the SM that loaded them:
//
//   phase A   worker CTA c owns the sample chunks c, c + workers, ... of MI_NS = 16 rows of X.  Per chunk:
//               Z = W1 Xc^T                   DMMA.8x8x4, A fragments straight from L2 (prefetched one m-tile pair ahead),
//                                             B fragments from the chunk in shared memory
//               H = sigmoid(Z + b1), out = sum_k H W2 + b2, res = out - x, S += res^2
//               dZ = res W2 H (1 - H), gb1 += dZ, gW2 += res H, gb2 += res        (shared memory, fixed order)
//               gW1 += dZ Xc                  DMMA, both operands in shared memory, 2 x 2 tile blocks per warp
//             all sums UN-SCALED (the factor d / S needs the S of every chunk) into the CTA's own row of `part`.
//             The last CTA meanwhile builds M = sI - A(W1) in the accumulator layout of the tensor-core sweep
//             (small_dmma.cuh), inverts it on chip and leaves h, log|det|, M^{-1} and sum |W1|.
//   phase B   every CTA adds the rows of `part` for its share of the parameters in a FIXED order (R lanes per
//             parameter, combined by shuffles: bit-reproducible), applies d / S, l1, dh/dW1 and weight decay and
//             takes the Adam step exactly as mlp_adam_kernel does.  The LAST CTA to reach the closing barrier writes
//             the state block (S, score, objective, h < 0 latch, step counter, ExponentialLR) before it opens it.
//
// Rows sharded over GPUs and stacks other than [d, m1, 1] stay on the launch sequence of mlp.cu.
#include "common.cuh"
#include "small_dmma.cuh"
#include "mlp_state.h"
#include "../../include/dagma_b200.h"

namespace dagma {

constexpr int MI_NT = 256;        // threads per CTA
constexpr int MI_NS = 16;         // samples per chunk (two DMMA n-tiles forward, four k-steps backward)
constexpr int MI_LDH = 20;        // row stride of Hs[P][16]: conflict-free A fragments
constexpr unsigned MI_SPIN_MAX = 1u << 22;

struct MlpIterArgs {
    MlpState* st;
    double *theta, *m, *v;                 // [total] parameters and Adam moments
    const double* X;                       // [n][d] row-major
    int n, n_total, d, m1;
    double* part;                          // [G][total + 1] per-CTA un-scaled gradient sums, then S
    double* Minv;                          // [d][d]
    unsigned* sync;                        // [0] arrivals [1] generation [2] error
    int iters;
};

__device__ __forceinline__ int mi_ldx(int d) { return 16 * ((d + 15) / 16) + 4; }   // row stride of Xs[16][.]: = 4 mod 16

// grid barrier; `last()` runs in the last CTA to arrive, before anybody is released
template <class F>
__device__ __forceinline__ void mi_grid_barrier(unsigned* sync, unsigned G, F&& last) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned* gen = sync + 1;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(sync, 1u) == G - 1) {
            last();
            *(volatile unsigned*)sync = 0u;
            __threadfence();
            atomicAdd(sync + 1, 1u);
        } else {
            unsigned spins = 0;
            while (*gen == g) {
                if (*(volatile unsigned*)(sync + 2)) break;
                if (++spins > MI_SPIN_MAX) { atomicExch(sync + 2, 1u); break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------- the CTA that owns h
__device__ __forceinline__ void mi_h_role(const MlpIterArgs& P, double* psm, SweepSync& sy) {
    using S = DmmaSmem;
    const int tid = threadIdx.x, d = P.d, m1 = P.m1;
    const DmmaPos ps(tid);
    double* red = psm + S::red;
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
    // M = (s I - A) / scale, A[r][c] = sum_k W1[c m1 + k][r]^2, identity padded
    double a[2][4][2];
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = ps.row(ti), c = ps.col(tj) + e;
                double v = (r == c) ? 1.0 : 0.0;
                if (r < d && c < d) {
                    double x = 0.0;
                    for (int k = 0; k < m1; ++k) {
                        const double w = __ldcg(W1 + (size_t)(c * m1 + k) * d + r);
                        x = fma(w, w, x);
                    }
                    v = (((r == c) ? s : 0.0) - x) * inv_scale;
                }
                a[ti][tj][e] = v;
            }
    double l1 = 0.0;
    for (int e = tid; e < d * m1 * d; e += MI_NT) l1 += fabs(__ldcg(W1 + e));
    dmma_sweep(a, ps, psm, d, sy);
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
}

// ---------------------------------------------------------------- a worker CTA: its sample chunks
struct MiSmem {                   // offsets in doubles
    int xs, hs, rs, b1, w2, b2, tail, red, total;
    __host__ __device__ MiSmem(int d, int m1) {
        const int P = d * m1, ldx = 16 * ((d + 15) / 16) + 4;
        xs = 0;
        hs = xs + MI_NS * ldx;
        rs = hs + 8 * ((P + 7) / 8) * MI_LDH;          // whole m-tiles: the padding rows stay zero
        b1 = rs + d * MI_NS;
        w2 = b1 + P;
        b2 = w2 + P;
        tail = b2 + d;                                  // [gb1 (P) | gW2 (P) | gb2 (d) | S]
        red = (tail + 2 * P + d + 1 + 1) & ~1;
        total = red + 96;
    }
};

template <int KT>                 // k-steps of the forward product: 4 KT >= d
__device__ __forceinline__ void mi_worker_role(const MlpIterArgs& P, double* sm, int cta, int workers) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, qr = lane >> 2, qc = lane & 3;
    const int d = P.d, m1 = P.m1, n = P.n, PP = d * m1;
    const int W = PP * d + 2 * PP + d;
    const int ldx = mi_ldx(d);
    const MiSmem L(d, m1);
    double *Xs = sm + L.xs, *Hs = sm + L.hs, *Rs = sm + L.rs, *b1s = sm + L.b1, *W2s = sm + L.w2, *b2s = sm + L.b2;
    double *tail = sm + L.tail, *red = sm + L.red;
    const double* W1 = P.theta;
    const int MT = (PP + 7) / 8, NT = (d + 7) / 8;
    const int chunks = (n + MI_NS - 1) / MI_NS;
    double* prow = P.part + (size_t)cta * (W + 1);

    for (int e = tid; e < PP; e += MI_NT) {
        b1s[e] = __ldcg(P.theta + (size_t)PP * d + e);
        W2s[e] = __ldcg(P.theta + (size_t)PP * d + PP + e);
    }
    for (int e = tid; e < d; e += MI_NT) b2s[e] = __ldcg(P.theta + (size_t)PP * d + 2 * PP + e);
    for (int e = tid; e < 2 * PP + d + 1; e += MI_NT) tail[e] = 0.0;
    for (int e = tid; e < (L.rs - L.hs); e += MI_NT) Hs[e] = 0.0;

    bool first = true;
    for (int chunk = cta; chunk < chunks; chunk += workers) {
        const int s0 = chunk * MI_NS, ns = min(MI_NS, n - s0);
        __syncthreads();                                   // the previous chunk is done with Xs / Hs / Rs
        for (int e = tid; e < MI_NS * ldx; e += MI_NT) {
            const int sr = e / ldx, c = e - sr * ldx;
            Xs[e] = (sr < ns && c < d) ? __ldg(P.X + (size_t)(s0 + sr) * d + c) : 0.0;
        }
        __syncthreads();
        // ---- Z = W1 Xc^T: warp w owns the m-tiles w, w + 8, ...; two at a time, the next pair's A fragments in flight
        {
            auto load_a = [&](double (&af)[2][KT], int mt0) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int row = 8 * (mt0 + 8 * u) + qr;
#pragma unroll
                    for (int ks = 0; ks < KT; ++ks) {
                        const int k = 4 * ks + qc;
                        af[u][ks] = (row < PP && k < d) ? __ldcg(W1 + (size_t)row * d + k) : 0.0;
                    }
                }
            };
            double af[2][KT], an[2][KT];
            if (warp < MT) load_a(af, warp);
            for (int mt0 = warp; mt0 < MT; mt0 += 16) {
                const bool more = mt0 + 16 < MT;
                if (more) load_a(an, mt0 + 16);
                double z[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                for (int ks = 0; ks < KT; ++ks) {
                    const double b0 = Xs[qr * ldx + 4 * ks + qc], b1 = Xs[(8 + qr) * ldx + 4 * ks + qc];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        dmma(z[u][0][0], z[u][0][1], af[u][ks], b0);
                        dmma(z[u][1][0], z[u][1][1], af[u][ks], b1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int p = 8 * (mt0 + 8 * u) + qr;
                    if (p < PP) {
                        const double bb = b1s[p];
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt) {
                            const double h0 = 1.0 / (1.0 + exp(-(z[u][nt][0] + bb))), h1 = 1.0 / (1.0 + exp(-(z[u][nt][1] + bb)));
                            *reinterpret_cast<double2*>(Hs + p * MI_LDH + 8 * nt + 2 * qc) = make_double2(h0, h1);
                        }
                    }
                }
                if (more) {
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int ks = 0; ks < KT; ++ks) af[u][ks] = an[u][ks];
                }
            }
        }
        __syncthreads();
        // ---- out, res, S
        double sq = 0.0, z1 = 0.0, z2 = 0.0;
        for (int e = tid; e < d * MI_NS; e += MI_NT) {
            const int j = e / MI_NS, sr = e - j * MI_NS;
            double o = b2s[j];
            for (int k = 0; k < m1; ++k) o = fma(Hs[(j * m1 + k) * MI_LDH + sr], W2s[j * m1 + k], o);
            const double r = (sr < ns) ? o - Xs[sr * ldx + j] : 0.0;
            Rs[e] = r;
            sq = fma(r, r, sq);
        }
        block_sum3<MI_NT>(sq, z1, z2, red, tid);          // (has the barriers that publish Rs)
        if (tid == 0) tail[2 * PP + d] += sq;
        // ---- dZ (over H), gb1, gW2, gb2
        for (int p = tid; p < PP; p += MI_NT) {
            const int j = p / m1;
            const double w2 = W2s[p];
            double gw2 = 0.0, gb1 = 0.0;
#pragma unroll
            for (int sr = 0; sr < MI_NS; ++sr) {
                const double hh = Hs[p * MI_LDH + sr], r = Rs[j * MI_NS + sr];
                const double dz = r * w2 * hh * (1.0 - hh);
                Hs[p * MI_LDH + sr] = dz;
                gw2 = fma(r, hh, gw2);
                gb1 += dz;
            }
            tail[p] += gb1;
            tail[PP + p] += gw2;
        }
        for (int j = tid; j < d; j += MI_NT) {
            double g = 0.0;
#pragma unroll
            for (int sr = 0; sr < MI_NS; ++sr) g += Rs[j * MI_NS + sr];
            tail[2 * PP + j] += g;
        }
        __syncthreads();
        // ---- gW1 += dZ Xc: 2 x 2 tile blocks
        {
            const int MB = (MT + 1) / 2, NB2 = (NT + 1) / 2;
            for (int u = warp; u < MB * NB2; u += MI_NT / 32) {
                const int mb = u / NB2, nb = u - mb * NB2;
                double g[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                for (int kk = 0; kk < MI_NS; kk += 4) {
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int row = 8 * (2 * mb + i) + qr;
                        av[i] = (row < 8 * MT) ? Hs[row * MI_LDH + kk + qc] : 0.0;
                        bv[i] = Xs[(kk + qc) * ldx + 8 * (2 * nb + i) + qr];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma(g[i][j][0], g[i][j][1], av[i], bv[j]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int p = 8 * (2 * mb + i) + qr, c = 8 * (2 * nb + j) + 2 * qc;
                        if (p < PP) {
                            double* q = prow + (size_t)p * d + c;
                            if (c < d) q[0] = first ? g[i][j][0] : q[0] + g[i][j][0];
                            if (c + 1 < d) q[1] = first ? g[i][j][1] : q[1] + g[i][j][1];
                        }
                    }
            }
        }
        first = false;
    }
    __syncthreads();
    if (!first)
        for (int e = tid; e < 2 * PP + d + 1; e += MI_NT) prow[(size_t)PP * d + e] = tail[e];
}

// ---------------------------------------------------------------- phase B: fixed-order sums + Adam (mlp_adam_kernel)
__device__ __forceinline__ void mi_update(const MlpIterArgs& P, int cta, int G, int nact, double S, int R) {
    const MlpState* st = P.st;
    const int tid = threadIdx.x, d = P.d, m1 = P.m1, PP = d * m1;
    const int W = PP * d + 2 * PP + d;
    const double mu = st->mu, lr = st->lr, b1 = st->beta1, b2 = st->beta2;
    const double wd = mu * st->lambda2, l1c = mu * st->lambda1;
    const double gs = mu * (double)d / S;
    const int step = st->step + 1;
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    const double step_size = lr / bc1, bc2s = sqrt(bc2);
    const int nW1 = PP * d;
    const int per_cta = MI_NT / R;                       // parameters per CTA and pass
    const int r = tid % R, q = tid / R;
    for (int base = cta * per_cta; base < W; base += G * per_cta) {
        const int e = base + q;
        double acc = 0.0;
        if (e < W) {
            const double* src = P.part + e;
            int c = r;
            for (; c + 7 * R < nact; c += 8 * R) {        // eight loads in flight per lane, added in order
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = __ldcg(src + (size_t)(c + u * R) * (W + 1));
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += t[u];
            }
            for (; c < nact; c += R) acc += __ldcg(src + (size_t)c * (W + 1));
        }
        for (int off = 1; off < R; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (e < W && r == 0) {
            const double p = P.theta[e];
            double g = gs * acc;
            if (e < nW1) {
                const int row = e / d, i = e - row * d, j = row / m1;
                const double sg = (p > 0.0) ? 1.0 : ((p < 0.0) ? -1.0 : 0.0);
                g = fma(l1c, sg, g);
                g = fma(2.0 * p, __ldcg(P.Minv + (size_t)j * d + i), g);
            }
            g = fma(wd, p, g);
            const double mo = P.m[e], vo = P.v[e];
            const double mn = mo + (g - mo) * (1.0 - b1);            // exp_avg.lerp_(grad, 1 - beta1)
            const double vn = fma(vo, b2, (1.0 - b2) * g * g);
            P.m[e] = mn;
            P.v[e] = vn;
            const double denom = sqrt(vn) / bc2s + 1e-8;
            P.theta[e] = p - step_size * (mn / denom);
        }
    }
}

template <int KT>
__global__ void __launch_bounds__(MI_NT, 1) mlp_iter_kernel(const MlpIterArgs P) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_S;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x;
    const int workers = G - 1;
    const int d = P.d, PP = d * P.m1, W = PP * d + 2 * PP + d;
    const int chunks = (P.n + MI_NS - 1) / MI_NS;
    const int nact = min(workers, chunks);
    int R = 1;                                            // lanes per parameter in phase B (power of two <= 8)
    while (R < 8 && 2 * R * W <= G * MI_NT) R *= 2;
    SweepSync sy{smem_u32(sm + DmmaSmem::mbar), 0u};
    if (cta == G - 1) {
        if (tid == 0) mbar_init(sy.bar, DM_NT / 32);
        __syncthreads();
    }
    for (int it = 0; it < P.iters; ++it) {
        if (*(volatile int32_t*)&P.st->halted != 0 || *(volatile unsigned*)(P.sync + 2) != 0u) break;   // uniform over the grid
        if (cta == G - 1) mi_h_role(P, sm, sy);
        else mi_worker_role<KT>(P, sm, cta, workers);
        mi_grid_barrier(P.sync, (unsigned)G, [] {});
        // ---- S in chunk-owner order; h < 0: no step (nonlinear.py:215-217)
        if (tid == 0) {
            double S = 0.0;
            for (int c = 0; c < nact; ++c) S += __ldcg(P.part + (size_t)c * (W + 1) + W);
            s_S = S;
        }
        __syncthreads();
        const double S = s_S;
        const double h = *(volatile double*)&P.st->h;
        const bool neg = h < 0.0;
        if (!neg) mi_update(P, cta, G, nact, S, R);
        mi_grid_barrier(P.sync, (unsigned)G, [&] {
            MlpState* st = P.st;
            const double l1 = *(volatile double*)&st->l1;
            const double score = 0.5 * (double)d * log(S / (double)P.n_total);
            st->S = S;
            st->score = score;
            st->obj = st->mu * (score + st->lambda1 * l1) + h;
            if (neg) st->halted = 1;
            else {
                st->step += 1;
                if (st->lr_gamma != 1.0 && (st->step % 1000) == 0) st->lr *= st->lr_gamma;
            }
        });
    }
    if (tid == 0 && cta == 0 && *(volatile unsigned*)(P.sync + 2) != 0u) P.st->info = 99;   // a barrier timed out
}

}  // namespace dagma

using namespace dagma;

static size_t mi_smem_bytes(int d, int m1) {
    const MiSmem L(d, m1);
    const size_t w = (size_t)L.total * sizeof(double);
    return w > DmmaSmem::bytes ? w : DmmaSmem::bytes;
}

extern "C" int dagma_mlp_iter_supported(int d, int m1) {
    return (d >= 1 && d <= 64 && m1 >= 1 && mi_smem_bytes(d, m1) <= 200 * 1024) ? 1 : 0;
}

extern "C" int dagma_mlp_iter_grid(int n) {
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return 0;
    const int chunks = (n + MI_NS - 1) / MI_NS;
    return chunks + 1 < sms ? chunks + 1 : sms;
}

extern "C" int dagma_mlp_iter_f64(dagma_stream_t stream, int n, int n_total, int d, int m1, int iters, void* state_dev,
                                  double* theta_dev, double* m_dev, double* v_dev, const double* x_dev, double* part_dev,
                                  double* minv_dev, unsigned* sync_dev) {
    DAGMA_REQUIRE(state_dev && theta_dev && m_dev && v_dev && x_dev && part_dev && minv_dev && sync_dev, "null pointer");
    DAGMA_REQUIRE(n >= 1 && iters >= 0 && dagma_mlp_iter_supported(d, m1), "shape not supported by the fused iteration");
    const int G = dagma_mlp_iter_grid(n);
    DAGMA_REQUIRE(G >= 2, "no device");
    const size_t smem = mi_smem_bytes(d, m1);
    MlpIterArgs A{(MlpState*)state_dev, theta_dev, m_dev, v_dev, x_dev, n, n_total, d, m1, part_dev, minv_dev, sync_dev, iters};
    const int kt = (d + 3) / 4;
#define MI_LAUNCH(KT)                                                                                            \
    do {                                                                                                         \
        DAGMA_CUDA_OK(cudaFuncSetAttribute(mlp_iter_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        mlp_iter_kernel<KT><<<G, MI_NT, smem, (cudaStream_t)stream>>>(A);                                        \
    } while (0)
    if (kt <= 4) MI_LAUNCH(4);
    else if (kt <= 8) MI_LAUNCH(8);
    else if (kt <= 12) MI_LAUNCH(12);
    else MI_LAUNCH(16);
#undef MI_LAUNCH
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}
