// Small-d (d <= 64) batched DAGMA-linear path: one persistent 256-thread CTA per
// problem runs whole minimize() stages -- or the whole path-following fit -- on chip.
//
// Reference semantics reproduced (src/dagma/linear.py, quirk numbers from SURVEY.md 8):
//   :226      M = inv(sI - W o W) + 1e-16, the shifted inverse enters the gradient (Q2)
//   :230-241  feasibility = any(M < 0); exit (W, False) at iter 1 or s <= 0.9, otherwise
//             undo / halve lr / redo with the previous Adam direction, Adam state and
//             mask_exc untouched (Q3)
//   :244,248  Gobj = -mu cov (I - W) + mu l1 sign(W) + 2 W o M^T + mask_inc sign(W) (Q1,Q5)
//   :158-162  Adam with 1-based bias correction, eps after the square root (Q4)
//   :275-276  W -= lr dir ; W *= mask_exc
//   :279,328  objective every `checkpoint` iterations and at max_iter, relative test (Q6)
//   :446-453  stage loop with retry (lr/2, s + 0.1) and warm start
//   :456-457  h_final with s = 1 and score_final on the un-thresholded W (Q11)
//
// Per iteration the CTA does ONE fused sweep (Gauss-Jordan inverse of M interleaved
// with the rank-1 updates of cov @ W, see small_gj.cuh) and one element-wise pass.
// The sweep that follows an update serves both the checkpoint objective of that
// update (log-det from the pivots, score from cov @ (I - W)) and the next iteration.
#include "common.cuh"
#include "small_gj.cuh"
#include "small_fit_internal.h"
#include <cstdlib>
#include "../../include/dagma_b200.h"

namespace dagma {

template <class C>
struct FitSmem {
    static constexpr int DP = C::DP;
    static constexpr int XLD = DP + 1;
    // offsets in doubles
    static constexpr size_t off_covT = 0;
    static constexpr size_t off_cov = off_covT + (size_t)DP * DP;
    static constexpr size_t off_W = off_cov + (size_t)DP * DP;
    static constexpr size_t off_xch = off_W + (size_t)DP * DP;
    static constexpr size_t off_line = ((off_xch + (size_t)DP * XLD + 1) / 2) * 2;
    static constexpr size_t off_red = off_line + ((SweepSmem<C>::doubles + 1) / 2) * 2;
    static constexpr size_t total = off_red + 96;
    static constexpr size_t bytes = total * sizeof(double);
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1) fit_small_kernel(const dagma_small_fit_args P) {
    using S = FitSmem<C>;
    constexpr int DP = C::DP, RM = C::RM, RN = C::RN, NT = C::NT, XLD = S::XLD;
    extern __shared__ __align__(16) double smem[];
    double* covT = smem + S::off_covT;
    double* covS = smem + S::off_cov;
    double* Ws = smem + S::off_W;
    double* xch = smem + S::off_xch;
    double* linebuf = smem + S::off_line;
    double* pivots = linebuf + SweepSmem<C>::piv_off;
    double* red = smem + S::off_red;
    const uint32_t covT_a = smem_u32(covT), cov_a = smem_u32(covS), W_a = smem_u32(Ws), line_a = smem_u32(linebuf);
    __shared__ unsigned s_prob;

    const int tid = threadIdx.x;
    const ThreadPos<C> pos(tid);
    const int ty = pos.ty, tx = pos.tx;
    const int d = P.d;
    const size_t dd = (size_t)d * d;

    // exclusion / inclusion bit masks of this thread's tile (shared by the batch)
    unsigned excbits = 0, incbits = 0;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int r = C::grow(ty, i), c = C::gcol(tx, j);
            if (r < d && c < d) {
                if (P.mask_exc_dev && P.mask_exc_dev[r * d + c]) excbits |= 1u << (i * RN + j);
                if (P.mask_inc_dev && P.mask_inc_dev[r * d + c]) incbits |= 1u << (i * RN + j);
            }
        }

    for (;;) {
        if (tid == 0) s_prob = atomicAdd(P.work_counter_dev, 1u);
        __syncthreads();
        const unsigned b = s_prob;
        __syncthreads();
        if (b >= (unsigned)P.batch) break;

        const double* cov_g = P.cov_dev + (size_t)b * dd;
        double* W_g = P.w_dev + (size_t)b * dd;
        const double lambda1 = P.lambda1_dev[b];

        for (int idx = tid; idx < DP * DP; idx += NT) {
            const int r = idx / DP, c = idx - r * DP;
            const bool in = (r < d) && (c < d);
            const double cv = in ? cov_g[r * d + c] : 0.0;
            covS[idx] = cv;
            covT[c * DP + r] = cv;
            Ws[idx] = in ? W_g[r * d + c] : 0.0;
        }
        __syncthreads();

        double a[RM][RN], g[RM][RN], m[RM][RN], v[RM][RN];
        int status = 0;
        int n_ckpt = 0;

        // ---- state of the path-following loop ----
        int stage = 0;
        bool final_phase = (P.n_stages == 0);
        double mu = 0, s_cur = 1, lr = 0, lr_adam = P.lr, obj_prev = 1e16;
        double last_obj = 0, last_score = 0, last_h = 0;
        int iters_max = 0, it = 0, retries = 0, backtracks = 0;
        bool in_backtrack = false;
        DD p1{1.0, 0.0}, p2{1.0, 0.0};

        unsigned long long t_attempt = 0;        // %globaltimer at the start of the minimize call (telemetry)
        auto start_attempt = [&]() {
            it = 0;
            t_attempt = global_ns();
            lr = lr_adam;
            obj_prev = 1e16;
            in_backtrack = false;
            backtracks = 0;
            p1 = DD{1.0, 0.0};
            p2 = DD{1.0, 0.0};
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) m[i][j] = v[i][j] = 0.0;
        };
        auto start_stage = [&]() {
            mu = P.mu[stage];
            s_cur = P.s[stage];
            iters_max = P.iters[stage];
            lr_adam = P.lr;
            retries = 0;
            start_attempt();
        };
        auto write_W = [&]() {   // Ws -> global (stage result / restart point)
            __syncthreads();
            for (int idx = tid; idx < d * d; idx += NT) {
                const int r = idx / d, c = idx - r * d;
                W_g[idx] = Ws[r * DP + c];
            }
        };
        auto read_W = [&]() {
            __syncthreads();
            for (int idx = tid; idx < d * d; idx += NT) {
                const int r = idx / d, c = idx - r * d;
                Ws[r * DP + c] = W_g[idx];
            }
            __syncthreads();
        };
        auto write_stage_stats = [&]() {
            if (tid == 0 && P.stage_stats_dev) {
                double* st = P.stage_stats_dev + ((size_t)b * P.n_stages + stage) * 8;
                st[0] = (double)it;
                st[1] = lr;
                st[2] = s_cur;
                st[3] = last_obj;
                st[4] = last_score;
                st[5] = last_h;
                st[6] = (double)retries;
                st[7] = (double)backtracks;
            }
        };
        // Ws += scale * dir(m, v, it): the previous Adam direction rebuilt from the moments
        auto apply_dir = [&](double scale) {
            const double c1 = 1.0 / p1.one_minus(), c2 = 1.0 / p2.one_minus();
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                double wrow[RN];
                const uint32_t wp = W_a + C::grow(ty, i) * DP * 8;
                load_rowfrag<C>(wrow, wp, tx);
#pragma unroll
                for (int j = 0; j < RN; ++j) {
                    const double dir = fast_div(m[i][j] * c1, fast_sqrt_nonneg(v[i][j] * c2) + 1e-8);
                    wrow[j] = __dadd_rn(wrow[j], __dmul_rn(scale, dir));
                }
                store_rowfrag<C>(wrow, wp, tx);
            }
            __syncthreads();
        };

        if (!final_phase) start_stage();

        for (;;) {
            // ================= build M and run the fused sweep =================
            const double s_use = final_phase ? 1.0 : s_cur;
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                double wrow[RN], crow[RN];
                const int r = C::grow(ty, i);
                load_rowfrag<C>(wrow, W_a + r * DP * 8, tx);
                load_rowfrag<C>(crow, cov_a + r * DP * 8, tx);
#pragma unroll
                for (int j = 0; j < RN; ++j) {
                    const int c = C::gcol(tx, j);
                    a[i][j] = ((r == c) ? s_use : 0.0) - wrow[j] * wrow[j];
                    g[i][j] = crow[j];
                }
            }
            gj_sweep<C, true>(a, g, covT_a, W_a, line_a, d, ty, tx);
            // now: a = M^{-1},  g = cov - cov W = cov (I - W),  pivots[0..d)

            // ================= objective pieces (checkpoint / final) =================
            const bool at_ckpt = !final_phase && !in_backtrack && it >= 1 &&
                                 (it % P.checkpoint == 0 || it == iters_max);
            if (at_ckpt || final_phase) {
                double sc = 0.0, l1 = 0.0, ld = 0.0;
#pragma unroll
                for (int i = 0; i < RM; ++i) {
                    double wrow[RN];
                    const int r = C::grow(ty, i);
                    load_rowfrag<C>(wrow, W_a + r * DP * 8, tx);
#pragma unroll
                    for (int j = 0; j < RN; ++j) {
                        const int c = C::gcol(tx, j);
                        const double dif = ((r == c && r < d) ? 1.0 : 0.0) - wrow[j];
                        sc = fma(dif, g[i][j], sc);
                        l1 += fabs(wrow[j]);
                    }
                }
                if (tid < d) ld = log(fabs(pivots[tid]));
                block_sum3<NT>(sc, l1, ld, red, tid);
                const double score = 0.5 * sc;
                const double h = -ld + (double)d * log(s_use);
                if (final_phase) {
                    if (tid == 0 && P.final_dev) {
                        P.final_dev[2 * (size_t)b] = h;
                        P.final_dev[2 * (size_t)b + 1] = score;
                    }
                    break;
                }
                const double obj = mu * (score + lambda1 * l1) + h;
                last_obj = obj;
                last_score = score;
                last_h = h;
                if (tid == 0 && P.ckpt_log_dev && n_ckpt < P.ckpt_log_cap) {
                    double* row = P.ckpt_log_dev + ((size_t)b * P.ckpt_log_cap + n_ckpt) * 6;
                    row[0] = stage; row[1] = it; row[2] = obj; row[3] = score; row[4] = h; row[5] = lr;
                }
                ++n_ckpt;
                const bool converged = fabs((obj_prev - obj) / obj_prev) <= P.tol;
                obj_prev = obj;
                if (converged || it == iters_max) goto stage_done;
            } else if (!final_phase && !in_backtrack && it == iters_max) {
                goto stage_done;   // only reachable for iters_max == 0
            }

            {
                // ================= feasibility of iteration it + 1 =================
                bool bad = false;
                if (tid < d) bad = !(pivots[tid] > 0.0);
#pragma unroll
                for (int i = 0; i < RM; ++i)
#pragma unroll
                    for (int j = 0; j < RN; ++j) bad |= (a[i][j] + 1e-16 < 0.0);
                bad = __syncthreads_or(bad);
                if (bad) {
                    if (it == 0 || s_cur <= 0.9) {            // linear.py:231-233
                        if (P.retry_on_fail) {
                            if (++retries > 64) {
                                // the reference would keep retrying; give up on the stage's restart point, but
                                // leave complete outputs: stage stats, status, and the final h / score of that W
                                status |= DAGMA_ST_RETRY_LIMIT;
                                --retries;
                                read_W();
                                write_stage_stats();
                                if (!P.final_dev) goto problem_done;
                                final_phase = true;
                                continue;
                            }
                            lr_adam *= 0.5;                    // linear.py:450-451
                            s_cur += 0.1;
                            read_W();
                            start_attempt();
                            continue;
                        }
                        status |= DAGMA_ST_OUT_OF_DOMAIN;
                        write_W();
                        write_stage_stats();
                        goto problem_done;
                    }
                    apply_dir(lr);                             // W += lr * grad   :235
                    lr *= 0.5;                                 //                  :236
                    if (lr <= 1e-16) {                         //                  :237-238
                        status |= DAGMA_ST_LR_UNDERFLOW;
                        goto stage_done;
                    }
                    apply_dir(-lr);                            // W -= lr * grad   :239
                    ++backtracks;
                    in_backtrack = true;
                    continue;                                  // re-invert        :240
                }
                in_backtrack = false;
            }

            {
                // ================= gradient, Adam, step (iteration it + 1) =================
                ++it;
                p1.mul(P.beta1);
                p2.mul(P.beta2);
                const double c1 = 1.0 / p1.one_minus(), c2 = 1.0 / p2.one_minus();
                const double ob1 = 1.0 - P.beta1, ob2 = 1.0 - P.beta2;
                const double l1c = mu * lambda1, incc = -2.0 * mu * lambda1;
                const bool diag_now = P.ckpt_diag_dev != nullptr && (it % P.checkpoint == 0 || it == iters_max);
                FitDiag dg;
                // transpose M^{-1} through shared memory: aT[i][j] = Minv[col][row]
#pragma unroll
                for (int i = 0; i < RM; ++i)
#pragma unroll
                    for (int j = 0; j < RN; ++j) xch[C::grow(ty, i) * XLD + C::gcol(tx, j)] = a[i][j];
                __syncthreads();
#pragma unroll
                for (int i = 0; i < RM; ++i) {
                    double wrow[RN];
                    const uint32_t wp = W_a + C::grow(ty, i) * DP * 8;
                    load_rowfrag<C>(wrow, wp, tx);
#pragma unroll
                    for (int j = 0; j < RN; ++j) {
                        const double w = wrow[j];
                        const double minvT = xch[C::gcol(tx, j) * XLD + C::grow(ty, i)];
                        const double sg = (w > 0.0) ? 1.0 : ((w < 0.0) ? -1.0 : 0.0);
                        double go = fma(-mu, g[i][j], l1c * sg);
                        go = fma(2.0 * w, minvT + 1e-16, go);
                        if (incbits >> (i * RN + j) & 1u) go = fma(incc, sg, go);
                        m[i][j] = fma(m[i][j], P.beta1, ob1 * go);
                        v[i][j] = fma(v[i][j], P.beta2, ob2 * (go * go));
                        const double dir = fast_div(m[i][j] * c1, fast_sqrt_nonneg(v[i][j] * c2) + 1e-8);
                        double wn = w - lr * dir;
                        if (excbits >> (i * RN + j) & 1u) wn = 0.0;
                        wrow[j] = wn;
                        if (diag_now) {
                            dg.grad(go, -mu * g[i][j], 2.0 * w * (minvT + 1e-16), w != 0.0,
                                    (incbits >> (i * RN + j) & 1u) != 0u);
                            dg.step(dir, wn);
                        }
                    }
                    store_rowfrag<C>(wrow, wp, tx);
                }
                __syncthreads();
                if (diag_now)
                    dg.finish<NT>(red, tid, l1c, incc, 1e-9 * (double)(global_ns() - t_attempt),
                                  n_ckpt < P.ckpt_log_cap
                                      ? P.ckpt_diag_dev + ((size_t)b * P.ckpt_log_cap + n_ckpt) * DAGMA_DIAG_COLS : nullptr);
                continue;
            }

        stage_done:
            write_stage_stats();
            write_W();
            ++stage;
            if (stage < P.n_stages) {
                start_stage();
                __syncthreads();
                continue;
            }
            if (!P.final_dev) goto problem_done;
            final_phase = true;
            __syncthreads();
        }
    problem_done:
        if (tid == 0) {
            if (P.status_dev) P.status_dev[b] = status;
            if (P.ckpt_count_dev) P.ckpt_count_dev[b] = n_ckpt < P.ckpt_log_cap ? n_ckpt : P.ckpt_log_cap;
        }
        __syncthreads();
    }
}

using C16 = Cfg<1, 1, 16, 16>;     // d <= 16 : 256 threads, 1 element each
using C32 = Cfg<2, 2, 16, 16>;     // d <= 32 : 256 threads, 2 x 2
using C48 = Cfg<4, 2, 12, 24>;     // d <= 48 : 288 threads, 4 x 2
using C64 = Cfg<4, 2, 16, 32>;     // d <= 64 : 512 threads, 4 x 2
using C64B = Cfg<4, 4, 16, 16>;    // d <= 64 : 256 threads, 4 x 4 (fewer shared-memory loads per FMA)
using C64C = Cfg<2, 4, 32, 16>;    // d <= 64 : 512 threads, 2 x 4

// DAGMA_SMALL_CFG (debug / A-B timing only): unset or 0 = DMMA kernel for 32 < d <= 64;
// 10 / 11 / 12 = scalar rank-1 kernels C64 / C64B / C64C
static int c64_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DAGMA_SMALL_CFG");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <class C>
static int fit_geometry(int batch, int sms, int* ctas, int* threads, size_t* smem_bytes) {
    constexpr size_t bytes = FitSmem<C>::bytes;
    DAGMA_CUDA_OK(cudaFuncSetAttribute(fit_small_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int per_sm = 0;
    DAGMA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_small_kernel<C>, C::NT, bytes));
    if (per_sm < 1) return set_error(-5, "fit kernel does not fit on an SM");
    long n = (long)sms * per_sm;
    if (n > batch) n = batch;
    if (ctas) *ctas = (int)n;
    if (threads) *threads = C::NT;
    if (smem_bytes) *smem_bytes = bytes;
    return 0;
}

template <class C>
static int launch_fit(cudaStream_t stream, const dagma_small_fit_args& a, int sms) {
    int ctas = 0;
    int rc = fit_geometry<C>(a.batch, sms, &ctas, nullptr, nullptr);
    if (rc) return rc;
    fit_small_kernel<C><<<ctas, C::NT, FitSmem<C>::bytes, stream>>>(a);
    DAGMA_CUDA_OK(cudaGetLastError());
    return 0;
}

static int sm_count(int* sms) {
    int dev = 0;
    DAGMA_CUDA_OK(cudaGetDevice(&dev));
    DAGMA_CUDA_OK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

}  // namespace dagma

using namespace dagma;

extern "C" int dagma_linear_fit_small_geometry(int d, int batch, int* ctas, int* threads, size_t* smem_bytes) {
    DAGMA_REQUIRE(d >= 1 && d <= DAGMA_SMALL_MAX_D, "d out of range for the on-chip fit path");
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    if (d <= 16) return fit_geometry<C16>(batch, sms, ctas, threads, smem_bytes);
    if (d <= 32) return fit_geometry<C32>(batch, sms, ctas, threads, smem_bytes);
    if (c64_variant() == 0) return fit_dmma_geometry(batch, sms, ctas, threads, smem_bytes);
    if (d <= 48) return fit_geometry<C48>(batch, sms, ctas, threads, smem_bytes);
    if (c64_variant() == 11) return fit_geometry<C64B>(batch, sms, ctas, threads, smem_bytes);
    if (c64_variant() == 12) return fit_geometry<C64C>(batch, sms, ctas, threads, smem_bytes);
    return fit_geometry<C64>(batch, sms, ctas, threads, smem_bytes);
}

extern "C" int dagma_linear_fit_small_f64(dagma_stream_t stream_, const dagma_small_fit_args* args) {
    DAGMA_REQUIRE(args != nullptr, "null args");
    const dagma_small_fit_args& a = *args;
    DAGMA_REQUIRE(a.batch >= 1, "batch must be positive");
    DAGMA_REQUIRE(a.d >= 1 && a.d <= DAGMA_SMALL_MAX_D, "d out of range for the on-chip fit path");
    DAGMA_REQUIRE(a.n_stages >= 0 && a.n_stages <= DAGMA_MAX_STAGES, "too many stages");
    DAGMA_REQUIRE(a.checkpoint >= 1, "checkpoint must be >= 1");
    DAGMA_REQUIRE(a.cov_dev && a.w_dev && a.lambda1_dev && a.work_counter_dev, "null device pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    DAGMA_CUDA_OK(cudaMemsetAsync(a.work_counter_dev, 0, sizeof(uint32_t), stream));
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    if (a.d <= 16) return launch_fit<C16>(stream, a, sms);
    if (a.d <= 32) return launch_fit<C32>(stream, a, sms);
    if (c64_variant() == 0) return launch_fit_dmma(stream, a, sms);
    if (a.d <= 48) return launch_fit<C48>(stream, a, sms);
    if (c64_variant() == 11) return launch_fit<C64B>(stream, a, sms);
    if (c64_variant() == 12) return launch_fit<C64C>(stream, a, sms);
    return launch_fit<C64>(stream, a, sms);
}
